// sc_search.cuh -- preamble search of one stream by one warp (src/qpsk.c:88-96, 172-183).
//
// pre[i] = v(1+1i), v = +-1  =>  pre[i]*s = v*(s.r - s.i) + i*v*(s.i + s.r) exactly (products by
// +-1 are exact and round-to-nearest is sign symmetric), so d = s.r - s.i and e = s.i + s.r are
// formed once per symbol and +-(d,e) is accumulated in the reference's order, 2 adds per tap
// instead of 8 operations.  Lane l evaluates lags 4l..4l+3 (128 sequential packed adds each) from
// a shared-memory copy of (d,e) laid out at pos(x) = x + x/4, which gives the lanes a stride of 5
// slots (odd => conflict-free) and compile-time offsets.  The argmax is the reference's: strict
// '>' against a running maximum that starts at 0.0f, first maximum wins.
#pragma once
#include "sc_common.cuh"
#include "sc_tables.cuh"

namespace sc {

constexpr int SEARCH_SYMS = 2 * PRE - 1;            // 255 symbols are read by the 128 lags
constexpr int SEARCH_DE_SLOTS = 320;                // >= pos(254) + 1 = 318

__device__ __forceinline__ int de_pos(int x) { return x + (x >> 2); }

// sym -> (d,e) operand (one rounding each, exactly the two adds the reference's product needs)
__device__ __forceinline__ float2 de_from_symbol(float2 w) {
    return make_float2(__fsub_rn(w.x, w.y), __fadd_rn(w.y, w.x));
}

// All 32 lanes must call; DE must be visible to the warp (__syncwarp before).  Returns the
// reference's (max_index, max_value) in every lane.
__device__ __forceinline__ void search_warp(const float2 *__restrict__ DE, int lane, int &best_idx, float &best_val) {
    u64 a[4] = {0ull, 0ull, 0ull, 0ull};
    const u64 *dp = reinterpret_cast<const u64 *>(DE) + 5 * lane;     // lags 4*lane .. 4*lane+3
#pragma unroll
    for (int j = 0; j < PRE + 3; j++) {
        const u64 v = dp[j + (j >> 2)];
#pragma unroll
        for (int qd = 0; qd < 4; qd++) {
            const int i = j - qd;
            if (i >= 0 && i < PRE) a[qd] = pre_neg(i) ? pk_sub(a[qd], v) : pk_add(a[qd], v);
        }
    }
    best_idx = 0;
    best_val = 0.0f;
#pragma unroll
    for (int qd = 0; qd < 4; qd++) {
        float re, im;
        unpk(a[qd], re, im);
        const float val = __fadd_rn(__fmul_rn(re, re), __fmul_rn(im, im));   // cnormf, qpsk.c:75-80
        if (val > best_val) {
            best_val = val;
            best_idx = 4 * lane + qd;
        }
    }
    // strict '>' with the first maximum winning == largest value, smallest lag among ties
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best_val, off);
        const int oi = __shfl_xor_sync(0xffffffffu, best_idx, off);
        if (ov > best_val || (ov == best_val && oi < best_idx)) {
            best_val = ov;
            best_idx = oi;
        }
    }
    if (!(best_val > 0.0f)) best_idx = 0;     // nothing ever exceeded the initial 0.0f
}

}  // namespace sc
