// sc_frontend_umma.cu -- the fused front-end with the preamble search PROPOSED on the 5th-generation tensor cores
// (SC_OPT_FE_SEARCH = SC_FE_SEARCH_TCGEN05).
//
// Same call as frontend_kernel (sc_rx_kernels.cu): int16 -> mix -> 49-tap RRC at the <= 290 decimated instants
// (src/qpsk.c:138-166, src/fir.c:29-43) -> 128-lag preamble correlation + first-maximum argmax (qpsk.c:88-96,
// 172-183) -> tracker window.  The staging and the FIR are the same code (sc_frontend.cuh).  The search is the
// proposer / verifier scheme of sc_search_mma.cuh -- all 128 correlations approximately as OUT = P * X on the tensor
// cores, a rigorous bound, the few lags that can still be the maximum evaluated with the reference's exact 128-term
// sequential sums -- so max_index / max_value are bit-identical to the all-exact search.  What differs from
// frontend_kernel<.., MMA> (mma.sync, bound by the shared-memory data pipe because every fragment passes through
// registers) is how the GEMM is fed:
//
//   * a CTA is 8 warps = 8 stream-frames; after the FIR passes each window W[290] lies in its warp's buffer;
//   * the 8 windows' d = s.r - s.i, e = s.i + s.r are split into two bf16 pieces each and written as ONE B operand,
//     N = 32 columns (piece p of window w = column 8 p + w) x K = 256 symbols, K-major no-swizzle core matrices, into
//     the dead part of the sample buffers: a warp stages four 8-symbol chunks of all 8 windows, so its 128-bit loads
//     and stores are conflict-free (the buffers are 16 bytes (mod 128) apart);
//   * the A operand is the 11.5 KB Toeplitz master (K-step s of P is the master read from row 240 - 16 s on: another
//     start address in the shared-memory descriptor), brought in by one TMA bulk copy at the start of the CTA;
//   * two threads issue 8 tcgen05.mma each (M = 128 lags, N = 32, K = 16), accumulating the two K halves in two
//     32-column accumulators in tensor memory; tcgen05.commit on an mbarrier;
//   * epilogue: thread = lag (warp % 4 = tensor-memory lane quarter, warp / 4 = which four windows): tcgen05.ld,
//     |re|^2 + |im|^2, redux.sync maxima, the bound's threshold, candidate lists by ballot;
//   * ONE warp verifies all 8 windows at once (lane = rank x window x component): 128 shared-memory loads per CTA
//     instead of 135 per window, which is what takes the search off the LSU pipe;
//   * the window is handed to the tracker in 64-byte segments (8 adjacent streams per row).
#include "sc_common.cuh"
#include "sc_tables.cuh"
#include "sc_frontend.cuh"
#include "sc_umma.cuh"
#include "sc_kernels.h"

namespace sc {

constexpr int FU_WARPS = 8;
constexpr int FU_THREADS = 32 * FU_WARPS;
constexpr int FU_BUF = 850;                          // float2 per warp: 6,800 bytes = 16 (mod 128)
constexpr int FU_BUF_BYTES = FU_BUF * 8;
static_assert(FU_BUF >= FE_BUF && FU_BUF_BYTES % 128 == 16, "buffer stride");
constexpr int FU_N = 4 * FU_WARPS;                   // MMA N: column 8 p + w, p = d_hi, d_mid, e_hi, e_mid
constexpr int FU_KSTEPS = 2 * PRE / 16;              // 16 MMAs of K = 16
constexpr int FU_ISSUERS = 2;                        // threads issuing MMAs: K-steps 0..7 and 8..15, an accumulator each
constexpr int FU_B_LBO = (FU_N / 8) * 128;           // 512: bytes between the 8-symbol chunks of B
constexpr int FU_B_SLICE = 2 * FU_B_LBO;             // 1,024 bytes per K-step
constexpr int FU_SBO = 128;
constexpr int FU_A_LBO = ((SU_A_ROWS + 7) / 8) * 128;
constexpr int FU_A_BYTES = 2 * FU_A_LBO;             // 11,776
static_assert(FU_A_BYTES == 16 * SU_A_WORDS4 && FU_A_BYTES % 128 == 0, "master size");
constexpr int FU_TMEM_COLS = FU_ISSUERS * FU_N;      // 64
constexpr int FU_W_BYTES = WIN * 8;                  // 2,320: W[290]; behind it (rounded up to 128) two K-steps of B
static_assert(FU_W_BYTES + 127 + 2 * FU_B_SLICE <= FU_BUF_BYTES, "B slices fit behind W");
constexpr uint32_t FU_IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t) (FU_N >> 3) << 17) | ((128u >> 4) << 24);

struct FuShared {
    unsigned long long a_full;                       // TMA -> issuers: the master has landed (1 + bytes)
    unsigned long long mma_done;                     // tcgen05.commit x 2 -> everybody
    uint32_t tmem_base;
    uint32_t warp_max[4][FU_WARPS];                  // per lane quarter and window
    float s_abs[FU_WARPS];                           // sum(|d| + |e|) per window
    int maxidx[FU_WARPS];
    int t2[FU_WARPS];
    int n_cand[4][FU_WARPS];
    unsigned char cand[4][FU_WARPS][SM_MAX_CAND];
};
constexpr int FU_SMEM = FU_A_BYTES + FU_WARPS * FU_BUF_BYTES + (int) sizeof(FuShared) + 1024;
static_assert(3 * (FU_SMEM + 1024) <= 228 * 1024, "three CTAs per SM");

__device__ __forceinline__ void umma_bf16_ss(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(a_desc), "l"(b_desc), "r"(FU_IDESC), "r"(accumulate)
        : "memory");
}
// this thread's TMEM lane, 4 consecutive columns
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float (&v)[4]) {
    uint32_t r[4];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(taddr));
#pragma unroll
    for (int i = 0; i < 4; i++) v[i] = __uint_as_float(r[i]);
}

template <bool WIDE>
__global__ void __launch_bounds__(FU_THREADS, 3)
frontend_umma_kernel(const int16_t *__restrict__ in, long stream_stride, const float2 *__restrict__ mix_table,
                     const int *__restrict__ timing_cur, const int *__restrict__ timing_next,
                     float2 *__restrict__ win, int *__restrict__ max_index_out, float *__restrict__ max_value_out,
                     int n_streams, const uint4 *__restrict__ a_master) {
    extern __shared__ unsigned char fu_smem_raw[];
    // the operands of the tensor core want 128-byte aligned core matrices: align the whole carve-up
    unsigned char *fu_smem = fu_smem_raw + ((1024u - (smem_u32(fu_smem_raw) & 1023u)) & 1023u);
    unsigned char *sA = fu_smem;
    unsigned char *sMix = fu_smem + FU_A_BYTES;
    FuShared &sh = *reinterpret_cast<FuShared *>(sMix + FU_WARPS * FU_BUF_BYTES);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long s0 = (long) blockIdx.x * FU_WARPS;
    const long s = s0 + warp;
    const bool active = s < n_streams;
    float2 *mix = reinterpret_cast<float2 *>(sMix + warp * FU_BUF_BYTES);

    // ---- set-up: barriers, tensor memory, the master of the A operand (TMA, lands during the FIR)
    if (tid == 0) {
        mbar_init(&sh.a_full, 1);
        mbar_init(&sh.mma_done, FU_ISSUERS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sh.tmem_base)),
                     "r"((uint32_t) FU_TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (tid == 0) {
        mbar_expect_tx(&sh.a_full, (uint32_t) FU_A_BYTES);
        tma_bulk_g2s(sA, a_master, (uint32_t) FU_A_BYTES, &sh.a_full);
    }
    const uint32_t tmem = sh.tmem_base;

    u64 accA[FE_R], accB[FE_R];
#pragma unroll
    for (int r = 0; r < FE_R; r++) accA[r] = accB[r] = 0ull;

    // warp-uniform set-up (frontend_kernel, GENERIC = false)
    int shift = 0, base2 = 0;
    const float2 *tab = mix_table;
    const uint32_t *fp = nullptr;
    uint32_t raw[FE_KN];
#pragma unroll
    for (int k = 0; k < FE_KN; k++) raw[k] = 0u;
    if (active) {
        const int T = timing_cur[s];
        const int base = max(min(T, 2 * PRE - 1) - (NTAPS - 1), 0);
        const int16_t *frame = in + s * stream_stride;
        base2 = base & ~1;
        shift = base - base2;
        fp = reinterpret_cast<const uint32_t *>(frame + base2);
        tab = mix_table + base2;
        fe_load<FE_KA_LO>(raw, fp, lane, base2);
        if (lane == 0) sh.t2[warp] = timing_next[s];
    }

#pragma unroll 1
    for (int h = 0; h < 2; h++) {
        const int h0 = h * CYC * FE_PASS_OUT;
        const int front = FE_FRONT + ((shift + h0) & 1);
        if (active) {
            if (h == 0) {
                fe_stage<FE_KA_LO>(mix, raw, tab, lane, front - shift - h0);
                fe_load<FE_KB_LO>(raw, fp, lane, base2);
            } else {
                fe_stage<FE_KB_LO>(mix, raw, tab, lane, front - shift - h0);
            }
        }
        __syncwarp();
        u64 acc[FE_R];
#pragma unroll
        for (int r = 0; r < FE_R; r++) acc[r] = 0ull;
        if (active && lane < FE_FIR_LANES) fe_fir<WIDE>(mix, front, lane, acc);
        __syncwarp();
#pragma unroll
        for (int r = 0; r < FE_R; r++) {
            if (h == 0) accA[r] = acc[r];
            else accB[r] = acc[r];
        }
    }

    // ---- W into the warp's own buffer; sum(|d| + |e|) over the 255 symbols the lags read (the bound's scale)
    {
        float2 *W = mix;
        float part = 0.0f;
        if (lane < FE_FIR_LANES) {
#pragma unroll
            for (int r = 0; r < FE_R; r++) {
                float yr, yi;
                unpk(accA[r], yr, yi);
                const float2 wa = make_float2(__fmul_rn(yr, FIR_GAIN), __fmul_rn(yi, FIR_GAIN));       // src/fir.c:42
                unpk(accB[r], yr, yi);
                const float2 wb = make_float2(__fmul_rn(yr, FIR_GAIN), __fmul_rn(yi, FIR_GAIN));
                const int xa = FE_R * lane + r, xb = FE_PASS_OUT + xa;
                W[xa] = wa;
                W[xb] = wb;
                part = __fadd_rn(part, __fadd_rn(fabsf(__fsub_rn(wa.x, wa.y)), fabsf(__fadd_rn(wa.y, wa.x))));
                if (xb < SEARCH_SYMS) part = __fadd_rn(part, __fadd_rn(fabsf(__fsub_rn(wb.x, wb.y)), fabsf(__fadd_rn(wb.y, wb.x))));
            }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) part = __fadd_rn(part, __shfl_xor_sync(0xffffffffu, part, off));
        if (lane == 0) sh.s_abs[warp] = part;
    }
    __syncthreads();                                         // every warp is done with its sample buffer

    // ---- B operand: this warp stages 8-symbol chunks 4 warp .. 4 warp + 3 of all 8 windows
    {
        const int w = lane & 7, c = 4 * warp + (lane >> 3);
        const uint4 *src = reinterpret_cast<const uint4 *>(sMix + w * FU_BUF_BYTES + c * 64);
        uint32_t dh[4], dm[4], eh[4], em[4];                            // bf16 pairs (x, x + 1): low half = x
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const uint4 v = src[q];                                     // symbols 8c + 2q, 8c + 2q + 1
            float d0 = __fsub_rn(__uint_as_float(v.x), __uint_as_float(v.y));      // qpsk.c:88-96, pre = v(1+i)
            float e0 = __fadd_rn(__uint_as_float(v.y), __uint_as_float(v.x));
            float d1 = __fsub_rn(__uint_as_float(v.z), __uint_as_float(v.w));
            float e1 = __fadd_rn(__uint_as_float(v.w), __uint_as_float(v.z));
            if (c == 31 && q == 3) d1 = e1 = 0.0f;                      // x = 255 is outside every lag's sum: P[.][255] = 0
            split2_pair(d0, d1, dh[q], dm[q]);
            split2_pair(e0, e1, eh[q], em[q]);
        }
        // K-step c / 2 lives behind W in the buffer of warp c / 4 (two K-steps per buffer), 128-byte aligned
        const int ks = c >> 1;
        const uint32_t hole = (smem_u32(sMix + (ks >> 1) * FU_BUF_BYTES + FU_W_BYTES) + 127u) & ~127u;
        const uint32_t dst = hole + (uint32_t) ((ks & 1) * FU_B_SLICE + (c & 1) * FU_B_LBO + w * 16);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + 0 * 128), "r"(dh[0]), "r"(dh[1]), "r"(dh[2]), "r"(dh[3]) : "memory");
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + 1 * 128), "r"(dm[0]), "r"(dm[1]), "r"(dm[2]), "r"(dm[3]) : "memory");
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + 2 * 128), "r"(eh[0]), "r"(eh[1]), "r"(eh[2]), "r"(eh[3]) : "memory");
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + 3 * 128), "r"(em[0]), "r"(em[1]), "r"(em[2]), "r"(em[3]) : "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // B is read by the tensor core
    tc_fence_before();
    __syncthreads();

    // ---- 16 MMAs: two issuing threads, K-steps 8 i .. 8 i + 7 into accumulator i
    if (lane == 0 && warp < FU_ISSUERS) {
        mbar_wait(&sh.a_full, 0u);
        tc_fence_after();
        const uint32_t a0 = smem_u32(sA);
#pragma unroll 1
        for (int k = 0; k < FU_KSTEPS / FU_ISSUERS; k++) {
            const int ks = warp * (FU_KSTEPS / FU_ISSUERS) + k;
            const uint32_t hole = (smem_u32(sMix + (ks >> 1) * FU_BUF_BYTES + FU_W_BYTES) + 127u) & ~127u;
            umma_bf16_ss(tmem + (uint32_t) warp * FU_N, umma_desc(a0 + (30 - 2 * ks) * 128, FU_A_LBO, FU_SBO),
                         umma_desc(hole + (ks & 1) * FU_B_SLICE, FU_B_LBO, FU_SBO), k > 0);
        }
        umma_commit(&sh.mma_done);
    }
    __syncwarp();
    mbar_wait(&sh.mma_done, 0u);
    tc_fence_after();

    // ---- epilogue: thread = lag 32 q + lane for the four windows 4 g .. 4 g + 3
    const int q4 = warp & 3, g4 = warp >> 2, lag = 32 * q4 + lane;
    float val[4];
    {
        const uint32_t taddr = tmem + ((uint32_t) (32 * q4) << 16) + (uint32_t) (4 * g4);
        float a[4], b[4], re[4];
        // re: d_hi (column 8 * 0 + w) + d_mid (8 * 1 + w), both accumulators
        tmem_ld4(taddr + 0, a);
        tmem_ld4(taddr + FU_N + 0, b);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 4; i++) re[i] = __fadd_rn(a[i], b[i]);
        tmem_ld4(taddr + 8, a);
        tmem_ld4(taddr + FU_N + 8, b);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 4; i++) re[i] = __fadd_rn(re[i], __fadd_rn(a[i], b[i]));
        tmem_ld4(taddr + 16, a);
        tmem_ld4(taddr + FU_N + 16, b);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 4; i++) val[i] = __fadd_rn(a[i], b[i]);
        tmem_ld4(taddr + 24, a);
        tmem_ld4(taddr + FU_N + 24, b);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const float im = __fadd_rn(val[i], __fadd_rn(a[i], b[i]));
            val[i] = __fadd_rn(__fmul_rn(re[i], re[i]), __fmul_rn(im, im));
        }
    }
    tc_fence_before();
    {
        // warp maxima (non-negative floats order like their bit patterns; a NaN sorts above everything, gives a NaN
        // threshold, no candidate, and ends in the verifier's fallback)
        uint32_t wm = 0;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const uint32_t m = __reduce_max_sync(0xffffffffu, __float_as_uint(val[i]));
            if (lane == i) wm = m;
        }
        if (lane < 4) sh.warp_max[q4][4 * g4 + lane] = wm;
    }
    __syncthreads();
    {
        float thr = 0.0f;                                               // lane i < 4: the threshold of window 4 g + i
        if (lane < 4) {
            const int w = 4 * g4 + lane;
            const uint32_t m = max(max(sh.warp_max[0][w], sh.warp_max[1][w]), max(sh.warp_max[2][w], sh.warp_max[3][w]));
            // |approx - reference| per component <= delta = 2^-13 sum(|d| + |e|): the bound of sc_search_mma.cuh /
            // sc_search_umma.cu (truncation of the split 2^-16 per piece pair, fp32 accumulation of 128 non-zero terms
            // <= 2^-15 even with truncating adders, one more fp32 addition for the two K halves, the reference's own
            // rounding 127 * 2^-24)
            thr = su_candidate_threshold(__uint_as_float(m), __fmul_rn(sh.s_abs[w], 0x1.004p-13f));
        }
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const float t = __shfl_sync(0xffffffffu, thr, i);
            const bool is = val[i] >= t;
            const unsigned m = __ballot_sync(0xffffffffu, is);
            const int pos = __popc(m & ((1u << lane) - 1u));
            if (is && pos < SM_MAX_CAND) sh.cand[q4][4 * g4 + i][pos] = (unsigned char) lag;
            if (lane == 0) sh.n_cand[q4][4 * g4 + i] = __popc(m);
        }
    }
    __syncthreads();

    // ---- verify: warp 0, lane = (rank parity k, window w, component): the reference's exact sums for the candidates
    if (warp == 0) {
        const int k2 = lane >> 4, w = (lane >> 1) & 7, comp = lane & 1;
        const bool exists = s0 + w < n_streams;
        const float2 *Ww = reinterpret_cast<const float2 *>(sMix + w * FU_BUF_BYTES);
        int cnt[4], nc = 0;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            cnt[q] = sh.n_cand[q][w];
            nc += cnt[q];
        }
        const bool direct = exists && nc >= 1 && nc <= SM_MAX_CAND;
        auto kth = [&](int k) {                                         // the k-th candidate of this lane's window, in lag order
            int q = 0, p = k;
#pragma unroll
            for (int qq = 0; qq < 3; qq++)
                if (q == qq && p >= cnt[qq]) {
                    p -= cnt[qq];
                    q = qq + 1;
                }
            return (int) sh.cand[q][w][p];
        };
        // largest exact value, smallest lag among equals == the reference's strict '>' scanning the lags upwards
        float ev = -1.0f;
        int ei = 1 << 20;
        int rounds = direct ? (nc + 1) >> 1 : 0;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) rounds = max(rounds, __shfl_xor_sync(0xffffffffu, rounds, off));
#pragma unroll 1
        for (int r = 0; r < rounds; r++) {
            const int k = 2 * r + k2;
            const bool have = direct && k < nc;
            const int L = have ? kth(k) : 0;
            const float part = su_exact_sum(Ww + L, comp);
            const float sq = __fmul_rn(part, part);
            const float v = __fadd_rn(sq, __shfl_xor_sync(0xffffffffu, sq, 1));    // cnormf, qpsk.c:75-80
            if (have && (v > ev || (v == ev && L < ei))) {
                ev = v;
                ei = L;
            }
        }
        {
            const float ov = __shfl_xor_sync(0xffffffffu, ev, 16);
            const int oi = __shfl_xor_sync(0xffffffffu, ei, 16);
            if (ov > ev || (ov == ev && oi < ei)) {
                ev = ov;
                ei = oi;
            }
        }
        if (!(ev > 0.0f)) ei = 0, ev = fmaxf(ev, 0.0f);
        // no candidate (NaNs) or too many (silence, ties over many lags): the full exact search, the warp per window
        unsigned fb = __ballot_sync(0xffffffffu, exists && !direct && comp == 0 && k2 == 0);
        while (fb) {
            const int l2 = __ffs(fb) - 1;
            fb &= fb - 1;
            const int w2 = (l2 >> 1) & 7;
            int bi;
            float bv;
            su_search_warp(reinterpret_cast<const float2 *>(sMix + w2 * FU_BUF_BYTES), lane, bi, bv);
            if (w == w2) {
                ei = bi;
                ev = bv;
            }
        }
        if (exists && comp == 0 && k2 == 0) {
            max_index_out[s0 + w] = ei;
            max_value_out[s0 + w] = ev;
            sh.maxidx[w] = ei;
        }
    }
    __syncthreads();

    // ---- hand the tracker its window: 64-byte segments (8 adjacent streams per row)
    {
        const int j = tid & (FU_WARPS - 1);
        const long sj = s0 + j;
        if (sj < n_streams) {
            const int mi = sh.maxidx[j], t2 = sh.t2[j];
            const float2 *Wj = reinterpret_cast<const float2 *>(sMix + j * FU_BUF_BYTES);
            float2 *dst = win + ((sj >> 5) * WIN_ROWS) * 32 + (sj & 31);
            for (int row = tid / FU_WARPS; row < WIN_ROWS; row += FU_THREADS / FU_WARPS) {
                const int src = row < X_ROWS ? mi + row : t2 + (row - X_ROWS);
                float2 v = make_float2(0.f, 0.f);
                if (src >= 0 && src < WIN) v = Wj[src];
                dst[row * 32] = v;
            }
        }
    }
    // every tcgen05.ld of this CTA was waited for before the barriers above
    if (warp == 1)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t) FU_TMEM_COLS) : "memory");
}

// The master of the A operand, host side: M[r][k] = pre[k - r] for r = -240 .. 127 (row index r + 240), k < 16, as bf16
// in the core-matrix layout: byte (k / 8) * LBO + (row / 8) * 128 + (row % 8) * 16 + (k % 8) * 2.  K-step s of the
// 128 x 256 Toeplitz matrix P[L][x] = pre[x - L] is rows 240 - 16 s .. 367 - 16 s of it.
void search_umma_make_master(uint16_t *table /* [SU_A_WORDS4 * 8] */) {
    for (int i = 0; i < SU_A_WORDS4 * 8; i++) table[i] = 0;
    for (int row = 0; row < SU_A_ROWS; row++) {
        for (int k = 0; k < 16; k++) {
            const int i = k - (row - 240);
            if (i < 0 || i >= PRE) continue;
            const int byte = (k / 8) * FU_A_LBO + (row / 8) * 128 + (row % 8) * 16 + (k % 8) * 2;
            table[byte / 2] = pre_neg(i) ? 0xBF80 : 0x3F80;             // -1.0 / +1.0
        }
    }
}

cudaError_t launch_frontend_umma(bool wide, const int16_t *in, long stream_stride, const float2 *mix_table,
                                 const int *timing_cur, const int *timing_next, float2 *win, int *max_index,
                                 float *max_value, int n_streams, cudaStream_t st, const void *a_master) {
    static std::atomic<unsigned long long> configured{0};              // bit per device
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev >= 64 || !((configured.load() >> dev) & 1ull)) {
        e = cudaFuncSetAttribute(frontend_umma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FU_SMEM);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(frontend_umma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FU_SMEM);
        if (e != cudaSuccess) return e;
        if (dev < 64) configured.fetch_or(1ull << dev);
    }
    const int grid = (n_streams + FU_WARPS - 1) / FU_WARPS;
    if (wide)
        frontend_umma_kernel<true><<<grid, FU_THREADS, FU_SMEM, st>>>(in, stream_stride, mix_table, timing_cur, timing_next, win,
                                                                       max_index, max_value, n_streams, (const uint4 *) a_master);
    else
        frontend_umma_kernel<false><<<grid, FU_THREADS, FU_SMEM, st>>>(in, stream_stride, mix_table, timing_cur, timing_next, win,
                                                                        max_index, max_value, n_streams, (const uint4 *) a_master);
    g_launch_count++;
    return cudaGetLastError();
}

bool frontend_umma_eligible(const int16_t *in, long stream_stride) {
    // the staging reads the samples as aligned 32-bit pairs (frontend_kernel's GENERIC = false case)
    return ((((uintptr_t) in) & 3) == 0) && ((stream_stride & 1) == 0);
}

}  // namespace sc
