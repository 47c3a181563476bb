// sc_frontend.cuh -- the sample staging and the 49-tap FIR of the fused front-end, shared by frontend_kernel
// (sc_rx_kernels.cu) and frontend_umma_kernel (sc_frontend_umma.cu).
#pragma once
#include "sc_common.cuh"
#include "sc_tables.cuh"
#include "sc_search.cuh"
#include "sc_search_mma.cuh"

namespace sc {

// ------------------------------------------------------------------------------------------------
// frontend_kernel: one warp per stream, FE_WARPS streams per CTA.
//
// Needed input: decimated instant i of the older half is filtered sample 5i + T (T = rx_timing at
// entry, qpsk.c:161), which depends on mixed samples 5i+T-48 .. 5i+T (src/fir.c: memory[] holds raw
// inputs, so any output can be evaluated on its own, bit-exactly).  For i < 290 that is the sample
// range [T-48, T+1445] of ONE frame (T >= 128 after the first call), 1494 samples.
//
// The 290 outputs are produced in two passes of 145 (29 lanes x 5 consecutive outputs), so only
// 769 mixed samples (6 KB) are staged in shared memory at a time: 25 KB per 4-warp CTA; with 80
// registers 6 CTAs = 24 warps are resident per SM, which is what hides the latencies of the staging
// and epilogue phases behind other warps' FIR/search arithmetic.  Lane l reads samples 25l .. 25l+68 of the pass: lane stride
// 25 slots (odd => conflict-free 64-bit reads), compile-time offsets, taps as immediates.
// All global loads of a stream-frame (24 coalesced 4-byte loads per lane) are issued up front.
// ------------------------------------------------------------------------------------------------
constexpr int FE_WARPS = 4;
constexpr int FE_NSAMP = NTAPS + CYC * (WIN - 1);           // 1494
constexpr int FE_R = 5;                                     // outputs per lane per pass
constexpr int FE_FIR_LANES = 29;
constexpr int FE_PASS_OUT = FE_FIR_LANES * FE_R;            // 145
constexpr int FE_PASS_SAMP = NTAPS + CYC * (FE_PASS_OUT - 1);   // 769
constexpr int FE_FRONT = 2;                                 // slack slots in front (pair alignment), +1 for parity
constexpr int FE_BUF = 840;                                 // >= FE_FRONT + 1 + 769 + pair slack; after the FIR passes the
                                                            // buffer holds W[290] and, behind it, one half of the window
                                                            // pair's tensor-core search structures (sc_search_mma.cuh)
static_assert((FE_BUF - WIN) * sizeof(float2) >= sizeof(SearchMmaB) && (FE_BUF - WIN) * sizeof(float2) >= sizeof(SearchMmaDE),
              "search structures fit behind W");
static_assert((FE_BUF * sizeof(float2)) % 16 == 0 && (WIN * sizeof(float2)) % 16 == 0, "16-byte aligned regions");
constexpr int FE_NPAIR = (FE_NSAMP + 2 + 1) / 2;            // 748 int16 pairs cover any alignment
constexpr int FE_PAIRS_PER_LANE = (FE_NPAIR + 31) / 32;     // 24
constexpr int FE_DE_SLOTS = (SEARCH_WORDS + 1) / 2;           // the search operands (floats) in float2 slots
static_assert(WIN == 2 * FE_PASS_OUT, "290 = 2 x 145");
static_assert(WIN + FE_DE_SLOTS <= FE_BUF, "W and (d,e) reuse the mixed-sample region");
static_assert(FE_FRONT + 1 + FE_PASS_SAMP + 2 <= FE_BUF, "pass buffer");
static_assert((WIN * 2) % 32 == 4, "search arrays start 4 banks after W: fine, only their relative offset matters");

// Stage the mixed samples rel in [h0, h0 + 769) of the stream-frame into buf (slot FE_FRONT + rel - h0).
// Pair p holds samples rel = 2p - shift and 2p + 1 - shift; out-of-range slots land in the slack.
constexpr int FE_KA_LO = 0, FE_KA_HI = 12;                  // pair rounds of pass A (rel 0..768)
constexpr int FE_KB_LO = 11, FE_KB_HI = FE_PAIRS_PER_LANE - 1;   // pair rounds of pass B (rel 725..1493)
constexpr int FE_KN = 13;
static_assert(FE_KA_HI - FE_KA_LO + 1 == FE_KN && FE_KB_HI - FE_KB_LO + 1 == FE_KN, "13 rounds per pass");

template <int K_LO>
__device__ __forceinline__ void fe_load(uint32_t (&raw)[FE_KN], const uint32_t *__restrict__ fp, int lane, int base2) {
#pragma unroll
    for (int k = 0; k < FE_KN; k++) {
        const int p = lane + 32 * (k + K_LO);
        // base2 <= 206 and p <= 747 keep every pair inside the 1880-sample frame; only the last round
        // of pass B runs past the 748 pairs that exist
        if (k + K_LO == FE_PAIRS_PER_LANE - 1) raw[k] = p < FE_NPAIR ? __ldg(fp + p) : 0u;
        else raw[k] = __ldg(fp + p);
    }
}

// TAB_SHARED: the phasor table lies in shared memory (frontend_umma_kernel stages the call's table once per CTA)
template <int K_LO, bool TAB_SHARED = false>
__device__ __forceinline__ void fe_stage(float2 *__restrict__ buf, const uint32_t (&rawk)[FE_KN],
                                         const float2 *__restrict__ tab, int lane, int off) {
    // off = front - shift - h0 is even (front is chosen per pass to make it so), hence every pair
    // lands on a 16-byte boundary and is written with one conflict-free 128-bit store
#pragma unroll
    for (int kk = 0; kk < FE_KN; kk++) {
        const int k = kk + K_LO;
        const uint32_t (&raw)[FE_KN] = rawk;
        const int p = lane + 32 * k;
        const int d = 2 * p + off;                            // slot of the pair's first sample
        // off is +2 in pass A and -722/-724 in pass B, so only three rounds can fall outside the buffer:
        // the last of pass A (slots >= FE_BUF), the first of pass B (slots < 0), the last of pass B (p >= 748)
        bool ok = true;
        if (K_LO == FE_KA_LO && k == FE_KA_HI) ok = d + 1 < FE_BUF;
        if (K_LO == FE_KB_LO && k == FE_KB_LO) ok = d >= 0;
        if (K_LO == FE_KB_LO && k == FE_KB_HI) ok = p < FE_NPAIR;
        if (ok) {
            const float4 ph = TAB_SHARED ? *reinterpret_cast<const float4 *>(tab + 2 * p)
                                         : __ldg(reinterpret_cast<const float4 *>(tab + 2 * p));
            const float v0 = (float) (int16_t) (raw[kk] & 0xffffu);
            const float v1 = (float) (int16_t) (raw[kk] >> 16);
            *reinterpret_cast<float4 *>(buf + d) = make_float4(__fmul_rn(ph.x, v0), __fmul_rn(ph.y, v0),     // qpsk.c:141
                                                               __fmul_rn(ph.z, v1), __fmul_rn(ph.w, v1));
        }
    }
}

// 49-tap RRC at 5 consecutive decimated instants per lane (src/fir.c:36-42): y += mem[i]*coeff[i]
// left to right, both components at once (packed f32x2, see sc_exact.cuh).
template <bool WIDE>
__device__ __forceinline__ void fe_fir(const float2 *__restrict__ buf, int front, int lane, u64 (&acc)[FE_R]) {
#pragma unroll
    for (int r = 0; r < FE_R; r++) acc[r] = 0ull;
    const u64 *mp = reinterpret_cast<const u64 *>(buf) + front + CYC * FE_R * lane;
#pragma unroll
    for (int j = 0; j < NTAPS + CYC * (FE_R - 1); j++) {
        const u64 x = mp[j];
#pragma unroll
        for (int r = 0; r < FE_R; r++) {
            const int k = j - CYC * r;
            if (k >= 0 && k < NTAPS) acc[r] = pk_add(acc[r], pk_mul_bcast_pz(x, tap<WIDE>(k)));
        }
    }
}

}  // namespace sc
