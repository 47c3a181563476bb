// sc_search_mma.cuh -- preamble search with a tensor-core PROPOSER and an exact VERIFIER.
//
// The reference's search (src/qpsk.c:88-96, 172-183) is 128 lags x 128 taps of sequential float32 adds;
// its outputs (max_index, max_value) must be reproduced to the last bit, and the exact sums cost 1,024
// dependent-order FADDs per lane (sc_search.cuh).  Only ONE lag wins, though.  So:
//
//   propose   all 128 correlations are first computed approximately on the tensor cores.  With
//             d = s.r - s.i, e = s.i + s.r (the exact operands of the reference's product, SURVEY A-4)
//             the correlation is  out[L] = sum_x P[L][x] * (d,e)[x],  P[L][x] = pre[x - L]  (a constant
//             128 x 256 Toeplitz matrix of +-1/0, exact in bf16) -- a GEMM  OUT = P * X  with the data as
//             the B operand.  d and e are split by truncation into two bf16 pieces each (16 significant
//             bits), one piece per B column, accumulated in fp32: mma.sync m16n8k16, 72 tiles per
//             window (9 non-zero 16x16 Toeplitz tiles per row block), A fragments from a constant table.
//   bound     |approx - reference| <= delta = 2^-13 * sum(|d| + |e|) per component (truncation 2^-14,
//             tensor-core accumulation < 2^-15, the reference's own rounding 127 * 2^-24), hence
//             |v_approx - v_ref| <= mu = 2 delta (m + delta) + 2^-20 v_max with m = sqrt(2 v_max).
//   verify    every lag with v_approx >= v_max - 2 mu is a candidate (almost always exactly one); each
//             candidate's sum is then evaluated EXACTLY -- the reference's 128 sequential adds per
//             component, one lane per (candidate, component) -- and the reference's argmax rule (strict
//             '>', first maximum wins, initial maximum 0.0f) is applied to the exact values.  More than
//             MAX_CAND candidates (silence, degenerate inputs) or none at all (NaN / Inf samples) fall back
//             to the full exact search.
//
// The result is therefore bit-identical to sc_search.cuh by construction; the tensor cores only decide
// WHICH of the exact sums are worth evaluating.
#pragma once
#include "sc_common.cuh"
#include "sc_tables.cuh"
#include "sc_search.cuh"

namespace sc {

constexpr int SM_KBLOCKS = 16;                       // 256 x-values in blocks of 16
constexpr int SM_ROWBLOCKS = 8;                      // 128 lags in blocks of 16
constexpr int SM_TILES = 9;                          // distinct Toeplitz tiles: t = kblock - rowblock = 0..8
constexpr int SM_COL_WORDS = 136;                    // 128 words of bf16 pairs + 8: conflict-free 64-bit fragment loads
constexpr int SM_MAX_CAND = 8;                       // candidates verified in one go per window
constexpr int SM_DE_FLOATS = 2 * PRE + 8;            // d[256] / e[256] (index 255 is zero) + 8 words: the four arrays of a
                                                     // window pair start 8 banks apart, so the verifier's lanes (which
                                                     // walk all four in step) never collide

// A-fragment table (device memory, built once per device by search_mma_table()): a_table[t * 32 + lane] = the
// four .b32 registers of lane for Toeplitz tile t, value(r, c) = pre[16 t + c - r] as bf16, 0 outside 0..127.
static __constant__ uint32_t c_search_pre_neg[4] = {pre_neg_word(0), pre_neg_word(1), pre_neg_word(2), pre_neg_word(3)};

// Shared memory of one window PAIR, in two parts so that the fused front-end can place them in the (dead) sample
// buffers of the two warps that own the windows.
struct SearchMmaB {
    uint32_t bp[8][SM_COL_WORDS];                    // B operand: column n, bf16 pairs in fragment order (4,352 bytes)
};
struct SearchMmaDE {
    float de[2][2][SM_DE_FLOATS];                    // [window][d/e][x]
    int cand[2][SM_MAX_CAND];
    int n_cand[2];
    float s_abs[2];                                  // sum(|d| + |e|) of each window (the fused front-end passes it here)
};                                                   // 4,312 bytes
struct SearchMmaSmem {
    SearchMmaB b;
    SearchMmaDE d;
};

__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint4 &a, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b0), "r"(b1));
}

// two-piece truncation split of v: hi = top 16 bits of v, mid = top 16 bits of (v - hi); both exact bf16
__device__ __forceinline__ void split2(float v, uint32_t &hi, uint32_t &mid) {
    hi = __float_as_uint(v) & 0xffff0000u;
    const float r = __fsub_rn(v, __uint_as_float(hi));            // exact: the low 16 significand bits
    mid = __float_as_uint(r) & 0xffff0000u;
}

// word offset (inside a column) of the bf16 pair (x, x+1), x even: K-block x/16, fragment order inside
__device__ __forceinline__ int bp_word(int x) {
    const int j = (x & 15) >> 1;
    return (x >> 4) * 8 + 2 * (j & 3) + (j >> 2);
}

// Stage one symbol pair (x, x+1) of window w (0/1): d/e floats for the verifier, bf16 pieces for the proposer.
// Returns |d|+|e| of both symbols (for the error bound).
// The bf16 pieces of one (d, e) pair of symbols (x, x+1) -> B operand columns of window w.
__device__ __forceinline__ void search_mma_stage_pieces(SearchMmaB &sb, int w, int x, float d0, float d1, float e0, float e1) {
    uint32_t dh0, dm0, dh1, dm1, eh0, em0, eh1, em1;
    split2(d0, dh0, dm0);
    split2(d1, dh1, dm1);
    split2(e0, eh0, em0);
    split2(e1, eh1, em1);
    const int word = bp_word(x);
    // columns: 4w + {0: d_hi, 1: d_mid, 2: e_hi, 3: e_mid}; low half = the lower k (x), high half = x + 1
    sb.bp[4 * w + 0][word] = __byte_perm(dh0, dh1, 0x7632);
    sb.bp[4 * w + 1][word] = __byte_perm(dm0, dm1, 0x7632);
    sb.bp[4 * w + 2][word] = __byte_perm(eh0, eh1, 0x7632);
    sb.bp[4 * w + 3][word] = __byte_perm(em0, em1, 0x7632);
}

__device__ __forceinline__ float search_mma_stage_pair(SearchMmaSmem &smem, int w, int x, float2 s0, float2 s1) {
    SearchMmaDE &sm = smem.d;
    const float d0 = __fsub_rn(s0.x, s0.y), e0 = __fadd_rn(s0.y, s0.x);       // qpsk.c:88-96 with pre = v(1+i)
    const float d1 = __fsub_rn(s1.x, s1.y), e1 = __fadd_rn(s1.y, s1.x);
    *reinterpret_cast<float2 *>(&sm.de[w][0][x]) = make_float2(d0, d1);
    *reinterpret_cast<float2 *>(&sm.de[w][1][x]) = make_float2(e0, e1);
    search_mma_stage_pieces(smem.b, w, x, d0, d1, e0, e1);
    return __fadd_rn(__fadd_rn(fabsf(d0), fabsf(e0)), __fadd_rn(fabsf(d1), fabsf(e1)));
}

// One lane's exact sum for lag L of the array X (d or e): the reference's order, 128 sequential adds.
__device__ __forceinline__ float search_exact_sum(const float *__restrict__ X, int L) {
    const float *p = X + L;
    float a = 0.0f;
#pragma unroll
    for (int i = 0; i < PRE; i++) a = pre_neg(i) ? __fsub_rn(a, p[i]) : __fadd_rn(a, p[i]);
    return a;
}

// Full exact search on unpadded d[]/e[] arrays (the rare fallback; same arithmetic as search_warp()).
__device__ __forceinline__ void search_warp_unpadded(const float *__restrict__ D, const float *__restrict__ E, int lane,
                                                     int &best_idx, float &best_val) {
    const int comp = lane >> 4, g = lane & 15;
    const float *p = (comp ? E : D) + 8 * g;
    float a[8];
#pragma unroll
    for (int q = 0; q < 8; q++) a[q] = 0.0f;
#pragma unroll 1
    for (int j0 = 0; j0 < PRE + 8; j0 += 8) {                          // rolled: code size matters more than speed here
#pragma unroll
        for (int jj = 0; jj < 8; jj++) {
            const int j = j0 + jj;
            if (j >= PRE + 7) break;
            const float v = p[j];
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const int i = j - q;
                if (i >= 0 && i < PRE) {
                    const bool neg = (c_search_pre_neg[i >> 5] >> (i & 31)) & 1u;
                    a[q] = neg ? __fsub_rn(a[q], v) : __fadd_rn(a[q], v);
                }
            }
        }
    }
    best_idx = 0;
    best_val = 0.0f;
#pragma unroll
    for (int q = 0; q < 8; q++) {
        const float o = __shfl_xor_sync(0xffffffffu, a[q], 16);
        const float re = comp ? o : a[q], im = comp ? a[q] : o;
        const float val = __fadd_rn(__fmul_rn(re, re), __fmul_rn(im, im));
        if (val > best_val) {
            best_val = val;
            best_idx = 8 * g + q;
        }
    }
#pragma unroll
    for (int off = 8; off > 0; off >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best_val, off);
        const int oi = __shfl_xor_sync(0xffffffffu, best_idx, off);
        if (ov > best_val || (ov == best_val && oi < best_idx)) {
            best_val = ov;
            best_idx = oi;
        }
    }
    if (!(best_val > 0.0f)) best_idx = 0;
}

// Exact verification by one 16-lane half of a warp: lane pair (2k, 2k+1) of the half evaluates the reference's sums
// (re, im) of candidate k (up to SM_MAX_CAND), then the reference's argmax rule -- strict '>' scanning the lags
// upwards, initial maximum 0.0f -- is applied to the exact values.  n_cand > SM_MAX_CAND gives (0, -1): the caller
// falls back to the full exact search.  All 32 lanes call; result in every lane of the half.
__device__ __forceinline__ void search_verify16(const float *__restrict__ D, const float *__restrict__ E,
                                                const int *__restrict__ cand, int nc, int lane, int &ei, float &ev) {
    const int vk = (lane & 15) >> 1, vc = lane & 1;
    const bool have = vk < nc && nc <= SM_MAX_CAND;
    const int L = have ? cand[vk] : 0;
    const float part = search_exact_sum(vc ? E : D, L);
    const float sq = __fmul_rn(part, part);
    // cnormf: re*re + im*im (qpsk.c:75-80); a float add commutes, so both lanes of the pair get the same bits
    ev = __fadd_rn(sq, __shfl_xor_sync(0xffffffffu, sq, 1));
    ei = L;
    if (!have) {
        ev = -1.0f;
        ei = 1 << 20;
    }
    // largest exact value, smallest lag among equals == strict '>' scanning lags upwards
#pragma unroll
    for (int off = 2; off < 16; off <<= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, ev, off);
        const int oi = __shfl_xor_sync(0xffffffffu, ei, off);
        if (ov > ev || (ov == ev && oi < ei)) {
            ev = ov;
            ei = oi;
        }
    }
    if (!(ev > 0.0f)) ei = 0, ev = fmaxf(ev, 0.0f);
}

// Error bound -> candidate threshold.  |approx - reference| <= delta per component, so
// |v_approx - v_ref| <= mu = 2 delta (m + delta) + 2^-20 v_max with m = sqrt(2 v_max) >= |re| + |im| of every lag,
// and every lag whose approximate value reaches v_max - 2 mu may be the true maximum.
__device__ __forceinline__ float search_candidate_threshold(float vmax, float delta) {
    const float m = __fmul_rn(sqrtf(__fmul_rn(2.0f, vmax)), 1.0001f);
    const float mu = __fadd_rn(__fmul_rn(__fmul_rn(2.0f, delta), __fadd_rn(m, delta)), __fmul_rn(vmax, 0x1p-20f));
    return __fsub_rn(vmax, __fmul_rn(mu, 2.002f));
}

// Proposer + verifier for the TWO windows staged in (sb, sd) (window 1 may be all zeros).  All 32 lanes call;
// sd.s_abs[w] = sum(|d|+|e|) of window w.  Results in every lane.
// B_IN_REGS: the 128 lags are done in two halves of 64 with the 12 B fragments of a half held in registers
// (24 fragment loads instead of 72: for callers that are short of shared-memory bandwidth, i.e. the front-end).
template <bool B_IN_REGS>
__device__ __forceinline__ void search_mma_pair(const SearchMmaB &sb, SearchMmaDE &sd, const uint4 *__restrict__ a_table,
                                                int lane, int (&best_idx)[2], float (&best_val)[2]) {
    const int g = lane >> 2, tid = lane & 3;
    const uint2 *bfrag = reinterpret_cast<const uint2 *>(&sb.bp[g][2 * tid]);
    // lane (g, tid): tid 0/1 = re/im of window 0, tid 2/3 = re/im of window 1; rows g and g + 8 of every row block
    const int w = tid >> 1;
    float v[2 * SM_ROWBLOCKS];
    float vmax = 0.0f;
    int imax = 0;
    // ---- propose: OUT[lag][col] = P * X, 72 HMMA in two halves of 4 row blocks
#pragma unroll
    for (int half = 0; half < 2; half++) {
        constexpr int RB = SM_ROWBLOCKS / 2;
        float acc[RB][4];
#pragma unroll
        for (int a = 0; a < RB; a++) acc[a][0] = acc[a][1] = acc[a][2] = acc[a][3] = 0.0f;
        uint2 bf[RB + SM_TILES - 1];
        if (B_IN_REGS) {
#pragma unroll
            for (int b = 0; b < RB + SM_TILES - 1; b++) bf[b] = bfrag[4 * (RB * half + b)];
        }
#pragma unroll
        for (int t = 0; t < SM_TILES; t++) {
            const uint4 af = __ldg(a_table + t * 32 + lane);
#pragma unroll
            for (int a = 0; a < RB; a++) {
                const uint2 b2 = B_IN_REGS ? bf[a + t] : bfrag[4 * (RB * half + a + t)];
                mma_bf16_16816(acc[a], af, b2.x, b2.y);
            }
        }
#pragma unroll
        for (int a = 0; a < RB; a++) {
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const float part = __fadd_rn(acc[a][2 * h], acc[a][2 * h + 1]);        // hi + mid piece
                const float sq = __fmul_rn(part, part);
                const float val = __fadd_rn(sq, __shfl_xor_sync(0xffffffffu, sq, 1));   // re^2 + im^2
                v[2 * (RB * half + a) + h] = val;
                if (val > vmax) {
                    vmax = val;
                    imax = 16 * (RB * half + a) + 8 * h + g;
                }
            }
        }
    }
    // maximum over the 8 row groups (lanes with the same tid): the warp-reduced argmax of the proposal
#pragma unroll
    for (int off = 4; off < 32; off <<= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, vmax, off);
        const int oi = __shfl_xor_sync(0xffffffffu, imax, off);
        if (ov > vmax || (ov == vmax && oi < imax)) {
            vmax = ov;
            imax = oi;
        }
    }
    // ---- bound and candidates
    const float thr = search_candidate_threshold(vmax, __fmul_rn(sd.s_abs[w], 0x1.004p-13f));
    if (lane < 2) sd.n_cand[lane] = 0;
    __syncwarp();
    if ((tid & 1) == 0) {
#pragma unroll
        for (int k = 0; k < 2 * SM_ROWBLOCKS; k++) {
            if (v[k] >= thr) {
                const int pos = atomicAdd(&sd.n_cand[w], 1);
                if (pos < SM_MAX_CAND) sd.cand[w][pos] = 16 * (k >> 1) + 8 * (k & 1) + g;
            }
        }
    }
    __syncwarp();
    // ---- verify: lanes 0..15 take window 0, 16..31 window 1
    const int vw = lane >> 4;
    float ev;
    int ei;
    search_verify16(sd.de[vw][0], sd.de[vw][1], sd.cand[vw], sd.n_cand[vw], lane, ei, ev);
    // publish window results to every lane
#pragma unroll
    for (int ww = 0; ww < 2; ww++) {
        best_val[ww] = __shfl_sync(0xffffffffu, ev, 16 * ww);
        best_idx[ww] = __shfl_sync(0xffffffffu, ei, 16 * ww);
    }
    // ---- fallback: too many candidates (silence, ties over many lags) or none (a NaN or Inf sample poisons the whole
    // proposal, 0 x NaN, while the reference only loses the lags that contain it): the full exact search
#pragma unroll 1
    for (int ww = 0; ww < 2; ww++) {
        if (sd.n_cand[ww] > SM_MAX_CAND || sd.n_cand[ww] == 0)
            search_warp_unpadded(sd.de[ww][0], sd.de[ww][1], lane, best_idx[ww], best_val[ww]);
    }
}

}  // namespace sc
