// sc_frontend_umma.cu -- the fused front-end with the preamble search PROPOSED on the 5th-generation tensor cores
// (SC_OPT_FE_SEARCH = SC_FE_SEARCH_TCGEN05), as a persistent warp-specialised kernel.
//
// Same call as frontend_kernel (sc_rx_kernels.cu): int16 -> mix -> 49-tap RRC at the <= 290 decimated instants
// (src/qpsk.c:138-166, src/fir.c:29-43) -> 128-lag preamble correlation + first-maximum argmax (qpsk.c:88-96,
// 172-183) -> tracker window.  The staging and the FIR are the same code (sc_frontend.cuh).  The search is the
// proposer / verifier scheme of sc_search_mma.cuh -- all 128 correlations approximately as OUT = P * X on the tensor
// cores, a rigorous bound, the few lags that can still be the maximum evaluated with the reference's exact 128-term
// sequential sums -- so max_index / max_value are bit-identical to the all-exact search.
//
// One persistent CTA per SM (a kernel that allocates tensor memory is given one CTA per SM whatever it needs: measured,
// profiles/r02_fe_tcgen05.md), 20 warps, working on batches of 16 stream-frames:
//
//   * warps 0..15 (FIR): one stream-frame each per batch -- global loads, mixing, the two FIR passes in the warp's own
//     sample buffer, then W[290] into the batch's window buffer and sum(|d| + |e|) for the bound.  With four warps per
//     scheduler nothing hides a warp's own latencies, so rx_timing is fetched two batches ahead, pass A's samples
//     during the previous frame's second FIR pass, and the call's phasor table (15 KB, the same for every stream) lies
//     in shared memory (one TMA bulk copy per CTA).  The only hand-overs to the search warps are two mbarriers
//     (windows full / windows free again);
//   * warps 16..19 (search), one batch behind the FIR warps:
//       - the 16 windows' d = s.r - s.i, e = s.i + s.r split into two bf16 pieces each and written as ONE B operand,
//         N = 64 columns (piece p of window w = column 16 p + w) x K = 256 symbols, K-major no-swizzle core matrices; a
//         quarter-warp handles one 8-symbol chunk of 8 windows, so its 128-bit loads and stores are conflict-free
//         (the windows lie 16 bytes (mod 128) apart);
//       - the A operand is the 11.5 KB Toeplitz master (K-step s of P is the master read from row 240 - 16 s on:
//         another start address in the shared-memory descriptor), brought in once per CTA by a TMA bulk copy;
//       - one thread issues 16 tcgen05.mma (M = 128 lags, N = 64, K = 16), D fp32 in tensor memory; tcgen05.commit
//         on an mbarrier;
//       - epilogue, thread = lag (warp % 4 = tensor-memory lane quarter): tcgen05.ld, |re|^2 + |im|^2 for the 16
//         windows, redux.sync maxima, the bound's threshold, candidate lists by ballot;
//       - verification, lane = (window, component): search warp q takes the candidates of rank q, q + 4, ... of every
//         window (128 shared-memory loads per round for 16 windows, against 135 per window in the all-exact search);
//       - the windows are handed to the tracker in 128-byte segments (16 adjacent streams per row).
//
// Measured (profiles/r02_fe_tcgen05.md): 0.469 ms per 131,072 stream-frames against 0.478 ms for frontend_kernel, with
// a third fewer instructions (270 M against 407 M); the search is off the FP32 and LSU pipes, what remains is the FIR
// with 16 warps per SM (6.7 KB of private buffer each limits their number).  Earlier forms: one role per CTA
// (62d71ad, 0.530 ms), 8 + 4 warps (0.537), before the latency fixes (0.526 / 0.502).
#include "sc_common.cuh"
#include "sc_tables.cuh"
#include "sc_frontend.cuh"
#include "sc_umma.cuh"
#include "sc_kernels.h"

#include <stdio.h>
#include <stdlib.h>

namespace sc {

constexpr int FU_WIN = 16;                           // stream-frames per batch = FIR warps
constexpr int FU_SRCH_WARP0 = FU_WIN, FU_SRCH_WARPS = 4;
constexpr int FU_THREADS = 32 * (FU_WIN + FU_SRCH_WARPS);
constexpr int FU_MIX_BYTES = FE_BUF * 8;             // a FIR warp's sample buffer
constexpr int FU_W_BYTES = WIN * 8;                  // a window: 2,320 bytes = 16 (mod 128)
static_assert(FU_W_BYTES % 128 == 16 && FU_MIX_BYTES % 16 == 0, "window stride");
constexpr int FU_N = 4 * FU_WIN;                     // MMA N: column 16 p + w, p = d_hi, d_mid, e_hi, e_mid
constexpr int FU_KSTEPS = 2 * PRE / 16;              // 16 MMAs of K = 16
constexpr int FU_B_LBO = (FU_N / 8) * 128;           // 1,024: bytes between the 8-symbol chunks of B
constexpr int FU_B_BYTES = (2 * PRE / 8) * FU_B_LBO; // 32 KB
constexpr int FU_SBO = 128;
constexpr int FU_A_LBO = ((SU_A_ROWS + 7) / 8) * 128;
constexpr int FU_A_BYTES = 2 * FU_A_LBO;             // 11,776
static_assert(FU_A_BYTES == 16 * SU_A_WORDS4 && FU_A_BYTES % 128 == 0, "master size");
constexpr int FU_TMEM_COLS = FU_N;                   // 64: one fp32 accumulator, 128 lags x 64 columns
constexpr uint32_t FU_IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t) (FU_N >> 3) << 17) | ((128u >> 4) << 24);

struct FuShared {
    unsigned long long a_full;                       // TMA -> issuers: the master has landed (1 + bytes)
    unsigned long long tab_full;                     // TMA -> FIR warps: the phasor table has landed (1 + bytes)
    unsigned long long w_full;                       // FIR warps -> search warps: the batch's windows are written (16)
    unsigned long long w_empty;                      // search warps -> FIR warps: ... and have been consumed (4)
    unsigned long long mma_done;                     // tcgen05.commit -> search warps
    uint32_t tmem_base;
    uint32_t warp_max[FU_SRCH_WARPS][FU_WIN];        // per lane quarter and window
    float s_abs[FU_WIN];                             // sum(|d| + |e|) per window
    int maxidx[FU_WIN];
    int t2[FU_WIN];
    int n_cand[FU_SRCH_WARPS][FU_WIN];
    float ver_v[FU_SRCH_WARPS][FU_WIN];              // verify: the best exact value / lag among the ranks of each search warp
    int ver_i[FU_SRCH_WARPS][FU_WIN];
    unsigned char cand[FU_SRCH_WARPS][FU_WIN][SM_MAX_CAND];
};
constexpr int FU_OFF_B = FU_A_BYTES;
constexpr int FU_OFF_W = FU_OFF_B + FU_B_BYTES;
constexpr int FU_OFF_MIX = FU_OFF_W + FU_WIN * FU_W_BYTES;
constexpr int FU_TAB_BYTES = FRAME * 8;               // the call's phasor table, 15,040 bytes (shared by every stream)
constexpr int FU_OFF_TAB = FU_OFF_MIX + FU_WIN * FU_MIX_BYTES;
constexpr int FU_OFF_CTRL = FU_OFF_TAB + FU_TAB_BYTES;
static_assert(FU_OFF_TAB % 16 == 0 && FU_TAB_BYTES % 16 == 0, "TMA bulk copy of the table");
static_assert(FU_OFF_B % 128 == 0 && FU_OFF_W % 128 == 0 && FU_OFF_MIX % 16 == 0 && FU_OFF_CTRL % 16 == 0, "alignment");
constexpr int FU_SMEM = FU_OFF_CTRL + (int) sizeof(FuShared) + 1024;
static_assert(FU_SMEM <= 227 * 1024, "one CTA per SM");

__device__ __forceinline__ void umma_bf16_ss(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(a_desc), "l"(b_desc), "r"(FU_IDESC), "r"(accumulate)
        : "memory");
}
// this thread's TMEM lane (row of D), 16 consecutive columns; the values may be used after tmem_ld_wait()
__device__ __forceinline__ void fu_tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
#pragma unroll
    for (int i = 0; i < 16; i++) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void srch_bar() { asm volatile("bar.sync 1, 128;" ::: "memory"); }   // the four search warps

// development aid (make SU_DEFS=-DFU_PROFILE): cycles per phase of CTA 0, printed at the end of the kernel
#ifdef FU_PROFILE
#define FU_T(acc) { const long long t_ = clock64(); acc += t_ - t_last; t_last = t_; }
#else
#define FU_T(acc)
#endif

// a global load that stays where it is written (the compiler otherwise sinks the rx_timing loads, which are issued a
// frame ahead on purpose, down to their use)
__device__ __forceinline__ int ld_pinned(const int *p) {
    int v;
    asm volatile("ld.global.nc.b32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

template <bool WIDE>
__global__ void __launch_bounds__(FU_THREADS, 1)
frontend_umma_kernel(const int16_t *__restrict__ in, long stream_stride, const float2 *__restrict__ mix_table,
                     const int *__restrict__ timing_cur, const int *__restrict__ timing_next,
                     float2 *__restrict__ win, int *__restrict__ max_index_out, float *__restrict__ max_value_out,
                     int n_streams, const uint4 *__restrict__ a_master) {
    extern __shared__ unsigned char fu_smem_raw[];
    // the operands of the tensor core want 128-byte aligned core matrices: align the whole carve-up
    unsigned char *fu_smem = fu_smem_raw + ((1024u - (smem_u32(fu_smem_raw) & 1023u)) & 1023u);
    unsigned char *sA = fu_smem;
    unsigned char *sB = fu_smem + FU_OFF_B;
    unsigned char *sW = fu_smem + FU_OFF_W;
    FuShared &sh = *reinterpret_cast<FuShared *>(fu_smem + FU_OFF_CTRL);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long n_batches = ((long) n_streams + FU_WIN - 1) / FU_WIN;
    const long my_batches = blockIdx.x < n_batches ? (n_batches - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

    // ---- set-up: barriers, tensor memory, the master of the A operand (TMA, lands during the first FIR)
    if (tid == 0) {
        mbar_init(&sh.a_full, 1);
        mbar_init(&sh.tab_full, 1);
        mbar_init(&sh.w_full, FU_WIN);
        mbar_init(&sh.w_empty, FU_SRCH_WARPS);
        mbar_init(&sh.mma_done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == FU_SRCH_WARP0) {
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
        mbar_expect_tx(&sh.tab_full, (uint32_t) FU_TAB_BYTES);
        tma_bulk_g2s(fu_smem + FU_OFF_TAB, mix_table, (uint32_t) FU_TAB_BYTES, &sh.tab_full);
    }
    const uint32_t tmem = sh.tmem_base;
#ifdef FU_PROFILE
    long long t_last = clock64(), p0 = 0, p1 = 0, p2 = 0, p3 = 0, p4 = 0, p5 = 0, p6 = 0, p7 = 0;
    const long long t_begin = t_last;
#endif

    if (warp < FU_WIN) {
        // ================= FIR warps: one stream-frame per batch (frontend_kernel with GENERIC = false)
        float2 *mix = reinterpret_cast<float2 *>(fu_smem + FU_OFF_MIX + warp * FU_MIX_BYTES);
        float2 *W = reinterpret_cast<float2 *>(sW + warp * FU_W_BYTES);
        // With four warps per scheduler nothing hides a frame's two dependent global latencies (rx_timing, then the
        // samples it selects), so both are fetched ahead: rx_timing two batches early, pass A's samples during the
        // previous frame's second FIR pass.
        auto stream_of = [&](long n) { return (blockIdx.x + n * gridDim.x) * FU_WIN + warp; };
        int shift = 0, base2 = 0;
        const float2 *const tab_s = reinterpret_cast<const float2 *>(fu_smem + FU_OFF_TAB);
        const float2 *tab = tab_s;
        const uint32_t *fp = nullptr;
        uint32_t raw[FE_KN];
#pragma unroll
        for (int k = 0; k < FE_KN; k++) raw[k] = 0u;
        auto setup = [&](long sx, int T) {                  // where stream sx's samples start for rx_timing T
            const int base = max(min(T, 2 * PRE - 1) - (NTAPS - 1), 0);
            base2 = base & ~1;
            shift = base - base2;
            fp = reinterpret_cast<const uint32_t *>(in + sx * stream_stride + base2);
            tab = tab_s + base2;
        };
        long s = stream_of(0);
        bool active = my_batches > 0 && s < n_streams;
        if (active) {
            setup(s, ld_pinned(timing_cur + s));
            fe_load<FE_KA_LO>(raw, fp, lane, base2);
        }
        long s_nx = stream_of(1);
        bool act_nx = my_batches > 1 && s_nx < n_streams;
        int T_nx = act_nx ? ld_pinned(timing_cur + s_nx) : 0;
        mbar_wait(&sh.tab_full, 0u);
#pragma unroll 1
        for (long n = 0; n < my_batches; n++) {
            const int t2 = active ? ld_pinned(timing_next + s) : 0;
            u64 accA[FE_R], accB[FE_R];
            {   // pass A: outputs 0..144
                const int front = FE_FRONT + (shift & 1);
                if (active) {
                    fe_stage<FE_KA_LO, true>(mix, raw, tab, lane, front - shift);
                    fe_load<FE_KB_LO>(raw, fp, lane, base2);
                }
                __syncwarp();
#pragma unroll
                for (int r = 0; r < FE_R; r++) accA[r] = 0ull;
                if (active && lane < FE_FIR_LANES) fe_fir<WIDE>(mix, front, lane, accA);
                __syncwarp();
            }
            {   // pass B: outputs 145..289
                constexpr int h0 = CYC * FE_PASS_OUT;
                const int front = FE_FRONT + ((shift + h0) & 1);
                if (active) fe_stage<FE_KB_LO, true>(mix, raw, tab, lane, front - shift - h0);
                const bool active_cur = active;
                s = s_nx;
                active = act_nx;
                if (active) {                               // the next frame's pass-A samples fly during this FIR pass
                    setup(s, T_nx);
                    fe_load<FE_KA_LO>(raw, fp, lane, base2);
                }
                s_nx = stream_of(n + 2);
                act_nx = n + 2 < my_batches && s_nx < n_streams;
                T_nx = act_nx ? ld_pinned(timing_cur + s_nx) : 0;
                __syncwarp();
#pragma unroll
                for (int r = 0; r < FE_R; r++) accB[r] = 0ull;
                if (active_cur && lane < FE_FIR_LANES) fe_fir<WIDE>(mix, front, lane, accB);
                __syncwarp();
            }
            // ---- W into the batch's window buffer once the search warps are done with the previous batch;
            // sum(|d| + |e|) over the 255 symbols the lags read (the bound's scale)
            FU_T(p0)
            mbar_wait(&sh.w_empty, ((uint32_t) n & 1u) ^ 1u);
            FU_T(p1)
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
                    if (xb < SEARCH_SYMS)
                        part = __fadd_rn(part, __fadd_rn(fabsf(__fsub_rn(wb.x, wb.y)), fabsf(__fadd_rn(wb.y, wb.x))));
                }
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) part = __fadd_rn(part, __shfl_xor_sync(0xffffffffu, part, off));
            if (lane == 0) {
                sh.s_abs[warp] = part;
                sh.t2[warp] = t2;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&sh.w_full);
            FU_T(p2)
        }
    } else {
        // ================= search warps: proposer on tcgen05, exact verification, window hand-over
        const int q4 = warp - FU_SRCH_WARP0, st = tid - 32 * FU_SRCH_WARP0, lag = st;
#pragma unroll 1
        for (long n = 0; n < my_batches; n++) {
            const long s0 = (blockIdx.x + n * gridDim.x) * FU_WIN;
            const uint32_t ph = (uint32_t) n & 1u;
            mbar_wait_relaxed(&sh.w_full, ph);          // ~20 % of a search warp's time: no hurry, the FIR warps are the pace
            FU_T(p0)
            // ---- B operand: thread t stages the 8-symbol chunks (t >> 4) + 8 i, i < 4, of window t & 15
#pragma unroll 1
            for (int i = 0; i < 4; i++) {
                const int w = st & 15, c = 8 * i + (st >> 4);
                const uint4 *src = reinterpret_cast<const uint4 *>(sW + w * FU_W_BYTES + c * 64);
                uint32_t dh[4], dm[4], eh[4], em[4];                    // bf16 pairs (x, x + 1): low half = x
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const uint4 v = src[q];                             // symbols 8c + 2q, 8c + 2q + 1
                    float d0 = __fsub_rn(__uint_as_float(v.x), __uint_as_float(v.y));      // qpsk.c:88-96, pre = v(1+i)
                    float e0 = __fadd_rn(__uint_as_float(v.y), __uint_as_float(v.x));
                    float d1 = __fsub_rn(__uint_as_float(v.z), __uint_as_float(v.w));
                    float e1 = __fadd_rn(__uint_as_float(v.w), __uint_as_float(v.z));
                    if (c == 31 && q == 3) d1 = e1 = 0.0f;              // x = 255 is outside every lag's sum: P[.][255] = 0
                    split2_pair(d0, d1, dh[q], dm[q]);
                    split2_pair(e0, e1, eh[q], em[q]);
                }
                // column n = 16 p + w: row group 2 p + (w >> 3), row w & 7; the 8 symbols are one 16-byte row
                unsigned char *dst = sB + c * FU_B_LBO + (w >> 3) * 128 + (w & 7) * 16;
                *reinterpret_cast<uint4 *>(dst + 0 * 256) = make_uint4(dh[0], dh[1], dh[2], dh[3]);
                *reinterpret_cast<uint4 *>(dst + 1 * 256) = make_uint4(dm[0], dm[1], dm[2], dm[3]);
                *reinterpret_cast<uint4 *>(dst + 2 * 256) = make_uint4(eh[0], eh[1], eh[2], eh[3]);
                *reinterpret_cast<uint4 *>(dst + 3 * 256) = make_uint4(em[0], em[1], em[2], em[3]);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // B is read by the tensor core
            tc_fence_before();
            srch_bar();
            FU_T(p1)

            // ---- 16 MMAs (M = 128 lags, N = 64, K = 16), one issuing thread
            if (tid == 32 * FU_SRCH_WARP0) {
                if (n == 0) mbar_wait(&sh.a_full, 0u);
                tc_fence_after();
                const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
#pragma unroll 1
                for (int ks = 0; ks < FU_KSTEPS; ks++)
                    umma_bf16_ss(tmem, umma_desc(a0 + (30 - 2 * ks) * 128, FU_A_LBO, FU_SBO),
                                 umma_desc(b0 + ks * 2 * FU_B_LBO, FU_B_LBO, FU_SBO), ks > 0);
                umma_commit(&sh.mma_done);
            }
            __syncwarp();
            mbar_wait(&sh.mma_done, ph);
            tc_fence_after();
            FU_T(p2)

            // ---- epilogue: thread = lag, all 16 windows
            float val[FU_WIN];
            {
                const uint32_t taddr = tmem + ((uint32_t) (32 * q4) << 16);
                float h[16], m[16];
                fu_tmem_ld16(taddr + 0, h);
                fu_tmem_ld16(taddr + 16, m);
                tmem_ld_wait();
#pragma unroll
                for (int w = 0; w < FU_WIN; w++) val[w] = __fadd_rn(h[w], m[w]);          // re
                fu_tmem_ld16(taddr + 32, h);
                fu_tmem_ld16(taddr + 48, m);
                tmem_ld_wait();
#pragma unroll
                for (int w = 0; w < FU_WIN; w++) {
                    const float im = __fadd_rn(h[w], m[w]);
                    val[w] = __fadd_rn(__fmul_rn(val[w], val[w]), __fmul_rn(im, im));
                }
            }
            tc_fence_before();
            {
                // warp maxima (non-negative floats order like their bit patterns; a NaN sorts above everything, gives
                // a NaN threshold, no candidate, and ends in the verifier's fallback)
                uint32_t wm = 0;
#pragma unroll
                for (int i = 0; i < FU_WIN; i++) {
                    const uint32_t m = __reduce_max_sync(0xffffffffu, __float_as_uint(val[i]));
                    if (lane == i) wm = m;
                }
                if (lane < FU_WIN) sh.warp_max[q4][lane] = wm;
            }
            srch_bar();
            FU_T(p3)
            {
                float thr = 0.0f;                                       // lane w < 16: the threshold of window w
                if (lane < FU_WIN) {
                    const uint32_t m = max(max(sh.warp_max[0][lane], sh.warp_max[1][lane]),
                                           max(sh.warp_max[2][lane], sh.warp_max[3][lane]));
                    // |approx - reference| per component <= delta = 2^-13 sum(|d| + |e|): the bound of
                    // sc_search_mma.cuh / sc_search_umma.cu (truncation of the split 2^-16 per piece pair, fp32
                    // accumulation of 128 non-zero terms <= 2^-15 even with truncating adders, the reference's own
                    // rounding 127 * 2^-24)
                    thr = su_candidate_threshold(__uint_as_float(m), __fmul_rn(sh.s_abs[lane], 0x1.004p-13f));
                }
#pragma unroll
                for (int i = 0; i < FU_WIN; i++) {
                    const float t = __shfl_sync(0xffffffffu, thr, i);
                    const bool is = val[i] >= t;
                    const unsigned m = __ballot_sync(0xffffffffu, is);
                    const int pos = __popc(m & ((1u << lane) - 1u));
                    if (is && pos < SM_MAX_CAND) sh.cand[q4][i][pos] = (unsigned char) lag;
                    if (lane == 0) sh.n_cand[q4][i] = __popc(m);
                }
            }
            srch_bar();
            FU_T(p4)

            // ---- verify: lane = (window, component): the reference's exact sums; search warp q takes the candidates of
            // rank q, q + 4, ... of every window (4 % of noise-only windows have a second candidate), warp 0 picks
            {
                const int w = lane >> 1, comp = lane & 1;
                const bool exists = s0 + w < n_streams;
                const float2 *Ww = reinterpret_cast<const float2 *>(sW + w * FU_W_BYTES);
                int cnt[4], nc = 0;
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    cnt[q] = sh.n_cand[q][w];
                    nc += cnt[q];
                }
                const bool direct = exists && nc >= 1 && nc <= SM_MAX_CAND;
                auto kth = [&](int k) {                                 // the k-th candidate of this lane's window, in lag order
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
                int rounds = direct && nc > q4 ? (nc - q4 + 3) >> 2 : 0;
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) rounds = max(rounds, __shfl_xor_sync(0xffffffffu, rounds, off));
#pragma unroll 1
                for (int r = 0; r < rounds; r++) {
                    const int k = q4 + 4 * r;
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
                if (comp == 0) {
                    sh.ver_v[q4][w] = ev;
                    sh.ver_i[q4][w] = ei;
                }
                srch_bar();
                if (q4 == 0) {
#pragma unroll
                    for (int q = 1; q < 4; q++) {
                        const float ov = sh.ver_v[q][w];
                        const int oi = sh.ver_i[q][w];
                        if (ov > ev || (ov == ev && oi < ei)) {
                            ev = ov;
                            ei = oi;
                        }
                    }
                    if (!(ev > 0.0f)) ei = 0, ev = fmaxf(ev, 0.0f);
                    // no candidate (NaNs) or too many (silence, ties over many lags): the full exact search, the warp per window
                    unsigned fb = __ballot_sync(0xffffffffu, exists && !direct && comp == 0);
                    while (fb) {
                        const int l2 = __ffs(fb) - 1;
                        fb &= fb - 1;
                        const int w2 = l2 >> 1;
                        int bi;
                        float bv;
                        su_search_warp(reinterpret_cast<const float2 *>(sW + w2 * FU_W_BYTES), lane, bi, bv);
                        if (w == w2) {
                            ei = bi;
                            ev = bv;
                        }
                    }
                    if (exists && comp == 0) {
                        max_index_out[s0 + w] = ei;
                        max_value_out[s0 + w] = ev;
                        sh.maxidx[w] = ei;
                    }
                }
            }
            srch_bar();
            FU_T(p5)

            // ---- hand the tracker its windows: 128-byte segments (16 adjacent streams per row)
            {
                const int j = st & (FU_WIN - 1);
                const long sj = s0 + j;
                if (sj < n_streams) {
                    const int mi = sh.maxidx[j], t2 = sh.t2[j];
                    const float2 *Wj = reinterpret_cast<const float2 *>(sW + j * FU_W_BYTES);
                    float2 *dst = win + ((sj >> 5) * WIN_ROWS) * 32 + (sj & 31);
#pragma unroll 4
                    for (int row = st / FU_WIN; row < WIN_ROWS; row += (32 * FU_SRCH_WARPS) / FU_WIN) {
                        const int src = row < X_ROWS ? mi + row : t2 + (row - X_ROWS);
                        float2 v = make_float2(0.f, 0.f);
                        if (src >= 0 && src < WIN) v = Wj[src];
                        dst[row * 32] = v;
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&sh.w_empty);
            FU_T(p6)
        }
    }

#ifdef FU_PROFILE
    if (blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == FU_SRCH_WARP0) && my_batches > 8)
        printf("FUPROF warp %d batches %ld total %lld | %lld %lld %lld %lld %lld %lld %lld %lld\n", warp, my_batches,
               clock64() - t_begin, p0, p1, p2, p3, p4, p5, p6, p7);
#endif
    tc_fence_before();
    __syncthreads();
    if (warp == FU_SRCH_WARP0)
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
    static std::atomic<int> sm_count[64];
    int dev = 0, sms = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev >= 64 || !((configured.load() >> dev) & 1ull)) {
        e = cudaFuncSetAttribute(frontend_umma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FU_SMEM);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(frontend_umma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FU_SMEM);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(frontend_umma_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                 (int) cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(frontend_umma_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                 (int) cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return e;
        int nb = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, frontend_umma_kernel<false>, FU_THREADS, FU_SMEM);
        if (e != cudaSuccess) return e;
        if (nb < 1) return cudaErrorLaunchOutOfResources;
        if (getenv("SC_FE_UMMA_DEBUG"))
            fprintf(stderr, "frontend_umma_kernel: %d CTAs per SM, %d bytes of shared memory each\n", nb, FU_SMEM);
        e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess) return e;
        if (dev < 64) {
            sm_count[dev].store(sms);
            configured.fetch_or(1ull << dev);
        }
    } else {
        sms = sm_count[dev].load();
    }
    const long n_batches = ((long) n_streams + FU_WIN - 1) / FU_WIN;
    const int grid = (int) std::min<long>(n_batches, (long) sms);      // persistent: one CTA per SM
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
