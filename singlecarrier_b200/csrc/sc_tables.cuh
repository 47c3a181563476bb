// sc_tables.cuh -- the modem's numeric tables for device code.
//
// The tables are local constexpr arrays inside constexpr accessors: with the fully unrolled loops
// of the kernels every index is a compile-time constant after unrolling, so the values fold into
// immediates / constant-bank operands.  Values come from include/sc_tables.inc.
#pragma once
#include "sc_common.cuh"

namespace sc {

__host__ __device__ constexpr float tap35(int k) {
#define SC_TABLE_RRC35 constexpr float t[NTAPS]
#include "../../include/sc_tables.inc"
#undef SC_TABLE_RRC35
    return t[k];
}
__host__ __device__ constexpr float tap50(int k) {
#define SC_TABLE_RRC50 constexpr float t[NTAPS]
#include "../../include/sc_tables.inc"
#undef SC_TABLE_RRC50
    return t[k];
}
template <bool WIDE>
__host__ __device__ constexpr float tap(int k) { return WIDE ? tap50(k) : tap35(k); }

__host__ __device__ constexpr int preamble_value(int i) {
#define SC_TABLE_PREAMBLE constexpr int8_t t[PRE]
#include "../../include/sc_tables.inc"
#undef SC_TABLE_PREAMBLE
    return t[i];
}
__host__ __device__ constexpr bool pre_neg(int i) { return preamble_value(i) < 0; }

// bit b of word w set <=> preamblevalues[32w + b] == -1
__host__ __device__ constexpr uint32_t pre_neg_word(int w) {
    uint32_t m = 0;
    for (int b = 0; b < 32; b++)
        if (preamble_value(w * 32 + b) < 0) m |= (1u << b);
    return m;
}

}  // namespace sc
