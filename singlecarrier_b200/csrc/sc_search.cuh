// sc_search.cuh -- preamble search of one stream by one warp (src/qpsk.c:88-96, 172-183).
//
// pre[i] = v(1+1i), v = +-1  =>  pre[i]*s = v*(s.r - s.i) + i*v*(s.i + s.r) exactly (products by
// +-1 are exact and round-to-nearest is sign symmetric), so d = s.r - s.i and e = s.i + s.r are
// formed once per symbol and +-d, +-e are accumulated in the reference's order: 2 adds per tap
// instead of 8 operations.
//
// The front-end is bound by the shared-memory (LSU) data pipe, so the operands are kept as two
// float arrays (d[] and e[]) and the warp is split: lanes 0-15 accumulate the real parts (d) and
// lanes 16-31 the imaginary parts (e) of lags 8g..8g+7 (g = lane & 15), 128 sequential adds each.
// Every shared-memory load is then a conflict-free 32-bit access of all 32 lanes (one wavefront):
// pos(x) = x + x/8 gives the lanes a stride of 9 words and compile-time offsets, and the e[] array
// starts 16 banks after d[] so the two half-warps hit complementary banks.  The argmax is the
// reference's: strict '>' against a running maximum that starts at 0.0f, first maximum wins.
#pragma once
#include "sc_common.cuh"
#include "sc_tables.cuh"

namespace sc {

constexpr int SEARCH_SYMS = 2 * PRE - 1;            // 255 symbols are read by the 128 lags
constexpr int SEARCH_E_OFF = 304;                   // e[] base: >= pos(254) + 1 = 286 and == 16 (mod 32)
constexpr int SEARCH_WORDS = SEARCH_E_OFF + 288;    // floats of shared memory per stream
constexpr int SEARCH_LAGS_PER_LANE = 8;

__device__ __forceinline__ int de_pos(int x) { return x + (x >> 3); }

// symbol x -> the two operands (one rounding each: exactly the adds the reference's product needs)
__device__ __forceinline__ void de_store(float *__restrict__ DE, int x, float2 w) {
    DE[de_pos(x)] = __fsub_rn(w.x, w.y);
    DE[SEARCH_E_OFF + de_pos(x)] = __fadd_rn(w.y, w.x);
}

// All 32 lanes must call; DE must be visible to the warp (__syncwarp before).  Returns the
// reference's (max_index, max_value) in every lane.
__device__ __forceinline__ void search_warp(const float *__restrict__ DE, int lane, int &best_idx, float &best_val) {
    const int comp = lane >> 4, g = lane & 15;
    const float *p = DE + comp * SEARCH_E_OFF + 9 * g;            // pos(8g + j) = 9g + j + j/8
    float a[SEARCH_LAGS_PER_LANE];
#pragma unroll
    for (int q = 0; q < SEARCH_LAGS_PER_LANE; q++) a[q] = 0.0f;
#pragma unroll
    for (int j = 0; j < PRE + SEARCH_LAGS_PER_LANE - 1; j++) {
        const float v = p[j + (j >> 3)];
#pragma unroll
        for (int q = 0; q < SEARCH_LAGS_PER_LANE; q++) {
            const int i = j - q;
            if (i >= 0 && i < PRE) a[q] = pre_neg(i) ? __fsub_rn(a[q], v) : __fadd_rn(a[q], v);
        }
    }
    best_idx = 0;
    best_val = 0.0f;
#pragma unroll
    for (int q = 0; q < SEARCH_LAGS_PER_LANE; q++) {
        const float o = __shfl_xor_sync(0xffffffffu, a[q], 16);
        const float re = comp ? o : a[q], im = comp ? a[q] : o;
        const float val = __fadd_rn(__fmul_rn(re, re), __fmul_rn(im, im));   // cnormf, qpsk.c:75-80
        if (val > best_val) {
            best_val = val;
            best_idx = SEARCH_LAGS_PER_LANE * g + q;
        }
    }
    // strict '>' with the first maximum winning == largest value, smallest lag among ties
#pragma unroll
    for (int off = 8; off > 0; off >>= 1) {
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
