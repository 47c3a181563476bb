// sc_search_umma.cu -- the preamble search of sc_search_mma.cuh with the PROPOSER on the 5th-generation tensor cores
// (tcgen05.mma, accumulators in tensor memory) and the samples brought in by TMA bulk copies.
//
// correlate() + argmax (src/qpsk.c:88-96, 172-183) over a batch of symbol windows.  As in sc_search_mma.cuh the 128
// correlations of a window are first computed approximately, OUT = P * X with P[L][x] = pre[x - L] a constant 128 x 256
// Toeplitz matrix of +-1 / 0 (exact in bf16) and X the window's d = s.r - s.i, e = s.i + s.r split by truncation into
// two bf16 pieces each; a rigorous bound on |approx - reference| selects the lags that can still be the maximum
// (almost always one), and those are evaluated with the reference's exact 128-term sequential sums, so the result is
// bit-identical to the all-exact search.  What changes is who does the GEMM and how the operands travel:
//
//   * one persistent CTA per SM (22 warps, each with one role) works on batches of 16 windows, every hand-over an
//     mbarrier, five batches in flight:
//       TMA warp        one 2 KB cp.async.bulk per window, issued by 16 lanes, three buffers deep;
//       8 staging warps d, e, their bf16 pieces (cvt.rz.bf16x2), sum(|d| + |e|) -> the B operand, N = 64 columns
//                       (4 pieces x 16 windows) x K = 256, in the K-major no-swizzle core-matrix layout the tensor
//                       core reads from shared memory;
//       2 MMA warps     one thread each (even / odd batches: issuing a tcgen05.mma costs its thread ~100 clocks
//                       whatever the shape) issues 16 tcgen05.mma, M = 128 lags, N = 64, K = 16, accumulating
//                       D[128 x 64] fp32 in tensor memory; tcgen05.commit signals an mbarrier;
//       2 x 4 epilogue  thread = lag: tcgen05.ld of its 64 accumulators, |re|^2 + |im|^2 for the 16 windows, redux.sync
//         warps         maxima, the bound's threshold, candidate lists by ballot;
//       3 verify warps  lane = (window, component): the candidate's 130 symbols L2 -> shared memory by one TMA bulk
//                       copy each (33 cp.async per lane cost the warp 13 % of the kernel's time), the
//                       reference's 128 sequential adds, the reference's argmax rule; second candidates are queued
//                       and verified 16 at a time; the rare fallbacks (no or too many candidates) run the full
//                       exact search;
//   * the A operand P (+-1 / 0) is written into tensor memory once per CTA (thread = lag = row, 128 columns of bf16
//     pairs) and read from there by every MMA: the shared-memory pipe then only feeds B (2 KB per MMA).  (First form:
//     A from shared memory through a descriptor -- slice s of the Toeplitz matrix is a 368 x 16 master read from row
//     240 - 16 s on, so 11.5 KB served all 16 slices -- correct, but slower.)
//
// Measured (tools/umma_bench.py, 2^20 windows): 0.51 ms with a preamble in every window, 0.56 ms on noise-only windows
// (4 % of which have a second candidate) -- 65 % / 60 % of the measured HBM peak, against 0.99 ms for the mma.sync
// kernel (shared-memory pipe at 85 %) and 1.45 ms for the all-exact one.  What bounds it now is spread over the roles
// (make SU_DEFS=-DSU_PROFILE + tools/umma_prof.py shows where each waits): every role is one warp per scheduler
// running dependent code.
#include "sc_common.cuh"
#include "sc_tables.cuh"
#include "sc_search_mma.cuh"
#include "sc_umma.cuh"
#include "sc_kernels.h"

namespace sc {

#ifndef SU_CFG_STAGES
#define SU_CFG_STAGES 3
#define SU_CFG_VERIFY 3
#endif
static_assert(SU_CFG_VERIFY % 2 == 1, "an odd number of verify warps");
constexpr int SU_WIN = 16;                           // windows per batch
constexpr int SU_N = 4 * SU_WIN;                     // MMA N: column p * 16 + w, p = d_hi, d_mid, e_hi, e_mid
constexpr int SU_KSTEPS = 2 * PRE / 16;              // 16 MMAs of K = 16
constexpr int SU_STAGES = SU_CFG_STAGES;             // raw-sample buffers (TMA in flight)
constexpr int SU_SLOTS = SU_CFG_VERIFY;               // verify warps (each with a gather scratch)
constexpr int SU_LISTS = 2 * SU_SLOTS;               // candidate-list slots: slot n % LISTS always meets the same epilogue set
                                                     // (n % 2) and the same verify warp (n % SLOTS, SLOTS odd), so every mbarrier has one
                                                     // sequential waiter on each side (a waiter two phases ahead of its
                                                     // barrier would see the parity it waits for)
constexpr int SU_EPI_WARPS = 4;                      // warps 0..3 (even batches) and 4..7 (odd): thread = lag, warp % 4 = TMEM lane quarter
constexpr int SU_MMA_WARP = 8;                       // warps 8, 9: one thread each issues the MMAs of the even / odd batches
constexpr int SU_TMA_WARP = 10;                      // warp 10: 16 lanes issue a batch's 16 bulk copies
constexpr int SU_STG_WARP0 = 11, SU_STG_WARPS = 8;   // warps 11..18: raw symbols -> B operand
constexpr int SU_VER_WARP0 = 19;                     // warps 19..21: exact verification
constexpr int SU_THREADS = 32 * (SU_VER_WARP0 + SU_SLOTS);
constexpr int SU_WIN_BYTES = 2 * PRE * 8;            // 256 complex floats
constexpr int SU_RAW_STRIDE = SU_WIN_BYTES + 16;     // per window in a raw buffer: 8 windows apart = 8 different 16-byte bank groups
constexpr int SU_RAW_BYTES = SU_WIN * SU_RAW_STRIDE;
constexpr int SU_B_LBO = (SU_N / 8) * 128;           // 1,024: bytes between 8-symbol chunks of B
constexpr int SU_B_BYTES = (2 * PRE / 8) * SU_B_LBO; // 32 KB per batch
constexpr int SU_SBO = 128;                          // 8 rows x 16 bytes: one core matrix
constexpr int SU_TMEM_D = 0;                         // tensor memory: two accumulator buffers of 64 columns ...
constexpr int SU_TMEM_A = 2 * SU_N;                  // ... and the A operand: 128 lags x 256 symbols, bf16 pairs per column
constexpr int SU_TMEM_COLS = 256;
static_assert(SU_TMEM_A + PRE <= SU_TMEM_COLS, "tensor memory columns");

// Batch n of a CTA goes through:  TMA warp -> raw[n % STAGES] -> staging warps -> B[n & 1] -> 16 MMAs (MMA warp) ->
// D[n & 1] in tensor memory -> epilogue set n & 1 -> candidate list n % LISTS -> verify warp n % SLOTS -> results.
// Every hand-over is an mbarrier; "empty" barriers are waited on with the inverted parity first (a fresh barrier
// reports its preceding phase as complete).
struct SuShared {
    unsigned long long raw_full[SU_STAGES];          // TMA -> staging            (1 arrival + bytes)
    unsigned long long raw_empty[SU_STAGES];         // staging -> TMA warp       (8 warps)
    unsigned long long b_full[2];                    // staging -> MMA warp       (8 warps)
    unsigned long long mma_done[2];                  // tcgen05.commit -> staging (B free) and epilogue (D ready)
    unsigned long long d_empty[2];                   // epilogue -> MMA warp      (1)
    unsigned long long cand_full[SU_LISTS];          // epilogue -> verify        (1)
    unsigned long long cand_empty[SU_LISTS];         // verify -> epilogue        (1)
    unsigned long long gathered[SU_SLOTS];           // TMA -> verify warp: the candidates' symbols have landed (1 + bytes)
    uint32_t tmem_base;
    uint32_t warp_max[2][SU_EPI_WARPS][SU_WIN];      // per epilogue set
    float part_abs[4][SU_STG_WARPS][SU_WIN];         // sum(|d| + |e|) per staging warp and window, slot n % 4
    // candidates of window w among the 32 lags of epilogue warp q: count (may exceed the list) and the first 8 lags
    int2 defer[SU_SLOTS][SU_WIN];                    // per verify warp: postponed second candidates (window, lag)
    int n_cand[SU_LISTS][SU_EPI_WARPS][SU_WIN];
    unsigned char cand[SU_LISTS][SU_EPI_WARPS][SU_WIN][SM_MAX_CAND];
};
constexpr int SU_SCR_STRIDE = PRE + 2;               // float2 per window in a verify warp's scratch: 130 = 65 16-byte chunks
                                                     // (a candidate's symbols from the even lag below it), windows 4 banks apart
constexpr int SU_SCR_BYTES = SU_WIN * SU_SCR_STRIDE * 8;
constexpr int SU_SMEM = 2 * SU_B_BYTES + SU_STAGES * SU_RAW_BYTES + SU_SLOTS * SU_SCR_BYTES + (int) sizeof(SuShared) + 128;
static_assert(SU_SMEM <= 227 * 1024, "shared memory");

// development aid: cycles spent in waits, per role (SU_CLK(slot) around a wait adds to clk[slot])
#ifdef SU_PROFILE
#define SU_T0 const long long t0_ = clock64();
#define SU_T1(acc) acc += clock64() - t0_;
#else
#define SU_T0
#define SU_T1(acc)
#endif
// Instruction descriptor (cute::UMMA::InstrDescriptor): D fp32, A and B bf16, both K-major, N = 64, M = 128
constexpr uint32_t SU_IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t) (SU_N >> 3) << 17) | ((128u >> 4) << 24);

// D[tmem] (+)= A[tmem] * B[smem]: the A operand in tensor memory (row = lane, 16 K elements = 8 columns of bf16 pairs)
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t b_desc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
        "r"(tmem_a), "l"(b_desc), "r"(SU_IDESC), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
        "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}

// this thread's TMEM lane (row of D), 16 consecutive columns; the values may be used after tmem_ld_wait()
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
#pragma unroll
    for (int i = 0; i < 16; i++) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void cp_async16(void *dst, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}


// One verify round over a warp's queue of postponed second candidates (see the verify role below): gather, exact
// sums, and the window's stored result corrected in place.  Kept out of line: it runs once in ~20 batches and would
// otherwise double the verify role's code.
__device__ __noinline__ void su_flush_queue(const int2 *__restrict__ queue, int qn, float2 *__restrict__ scr,
                                            const float2 *__restrict__ symbols, long symbol_stride,
                                            int *__restrict__ max_index, float *__restrict__ max_value, int lane) {
    const int w = lane >> 1, comp = lane & 1;
    const bool have = w < qn;
    const int2 it = have ? queue[w] : make_int2(0, 0);
    if (have) {
        const float2 *src = symbols + (long) it.x * symbol_stride + (it.y & ~1);
#pragma unroll
        for (int i = 0; i < 33; i++) {
            const int ch = 2 * i + comp;
            if (ch < 65) cp_async16(scr + 2 * ch, src + 2 * ch);
        }
    }
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
    __syncwarp();
    if (have) {
        const float part = su_exact_sum(scr + (it.y & 1), comp);
        const float sq = __fmul_rn(part, part);
        const float v = __fadd_rn(sq, __shfl_xor_sync(0x3u << (lane & ~1), sq, 1));   // cnormf, qpsk.c:75-80
        if (comp == 0) {                                        // strict '>' scanning the lags upwards: larger wins, and
            const float cur = max_value[it.x];                  // among equals the lower lag
            if (v > cur || (v == cur && v > 0.0f && it.y < max_index[it.x])) {
                max_value[it.x] = v;
                max_index[it.x] = it.y;
            }
        }
    }
    __syncwarp();
}

__device__ __forceinline__ void epi_bar(int set) {                     // the 128 threads of one epilogue set
    if (set) asm volatile("bar.sync 2, 128;" ::: "memory");
    else asm volatile("bar.sync 1, 128;" ::: "memory");
}

__global__ void __launch_bounds__(SU_THREADS, 1)
search_umma_batch_kernel(const float2 *__restrict__ symbols, long symbol_stride,
                         int *__restrict__ max_index, float *__restrict__ max_value, long n_streams,
                         float *__restrict__ dbg_approx) {
    extern __shared__ __align__(128) unsigned char su_smem[];
    unsigned char *sB = su_smem;
    unsigned char *sRaw = sB + 2 * SU_B_BYTES;
    unsigned char *sScr = sRaw + SU_STAGES * SU_RAW_BYTES;
    SuShared &sh = *reinterpret_cast<SuShared *>(sScr + SU_SLOTS * SU_SCR_BYTES);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long n_batches = (n_streams + SU_WIN - 1) / SU_WIN;
    const long my_batches = blockIdx.x < n_batches ? (n_batches - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

    // ---- set-up: barriers, tensor memory, the master of the A operand
    if (tid == 0) {
        for (int i = 0; i < SU_STAGES; i++) {
            mbar_init(&sh.raw_full[i], 1);
            mbar_init(&sh.raw_empty[i], SU_STG_WARPS);
        }
        for (int i = 0; i < 2; i++) {
            mbar_init(&sh.b_full[i], SU_STG_WARPS);
            mbar_init(&sh.mma_done[i], 1);
            mbar_init(&sh.d_empty[i], 1);
        }
        for (int i = 0; i < SU_LISTS; i++) {
            mbar_init(&sh.cand_full[i], 1);
            mbar_init(&sh.cand_empty[i], 1);
        }
        for (int i = 0; i < SU_SLOTS; i++) mbar_init(&sh.gathered[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sh.tmem_base)),
                     "r"((uint32_t) SU_TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = sh.tmem_base;
    long long w0 = 0, w1 = 0, w2 = 0, w3 = 0, w4 = 0, w5 = 0;          // SU_PROFILE: cycles in this role's waits / sections
    const long long t_begin = clock64();

    // ---- the A operand, once: P[L][x] = pre[x - L] (+-1 / 0, exact in bf16) into tensor memory, thread = lag = row
    if (warp < SU_EPI_WARPS) {
#pragma unroll 1
        for (int c0 = 0; c0 < PRE; c0 += 16) {
            uint32_t r[16];
#pragma unroll
            for (int q = 0; q < 16; q++) {
                uint32_t pair = 0;
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const int i = 2 * (c0 + q) + h - tid;               // preamble index of symbol x = 2 (c0 + q) + h at lag tid
                    uint32_t v = 0;
                    if (i >= 0 && i < PRE) v = ((c_search_pre_neg[i >> 5] >> (i & 31)) & 1u) ? 0xBF80u : 0x3F80u;
                    pair |= v << (16 * h);
                }
                r[q] = pair;
            }
            tmem_st16(tmem + ((uint32_t) (32 * warp) << 16) + SU_TMEM_A + c0, r);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    if (warp == SU_TMA_WARP) {
        // ================= TMA: one 2 KB bulk copy per window, issued by 16 lanes, as far ahead as buffers are free
        for (long n = 0; n < my_batches; n++) {
            const long b = blockIdx.x + n * gridDim.x;
            const int nw = (int) min((long) SU_WIN, n_streams - b * SU_WIN);
            const int st = (int) (n % SU_STAGES);
            if (lane == 0) {
                { SU_T0 mbar_wait(&sh.raw_empty[st], (uint32_t) ((n / SU_STAGES) & 1) ^ 1u); SU_T1(w0) }
                mbar_expect_tx(&sh.raw_full[st], (uint32_t) nw * SU_WIN_BYTES);
            }
            __syncwarp();
            if (lane < nw)
                tma_bulk_g2s(sRaw + st * SU_RAW_BYTES + lane * SU_RAW_STRIDE, symbols + (b * SU_WIN + lane) * symbol_stride,
                             SU_WIN_BYTES, &sh.raw_full[st]);
        }
    } else if (warp == SU_MMA_WARP || warp == SU_MMA_WARP + 1) {
        // ================= MMA: 16 x (128 lags x 64 columns x 16 symbols) per batch, A from tensor memory, B from shared.
        // Issuing one tcgen05.mma costs its thread ~100 clocks whatever N is: two issuers, one per accumulator buffer.
        if (lane == 0) {
            for (long n = warp - SU_MMA_WARP; n < my_batches; n += 2) {
                const int j = (int) (n & 1);
                { SU_T0 mbar_wait(&sh.b_full[j], (uint32_t) (n >> 1) & 1u); SU_T1(w0) }
                { SU_T0 mbar_wait(&sh.d_empty[j], ((uint32_t) (n >> 1) & 1u) ^ 1u); SU_T1(w1) }
                tc_fence_after();
                const uint32_t b0 = smem_u32(sB + j * SU_B_BYTES);
                {
                    SU_T0
#pragma unroll 1
                    for (int s = 0; s < SU_KSTEPS; s++)
                        umma_bf16_ts(tmem + SU_TMEM_D + (uint32_t) j * SU_N, tmem + SU_TMEM_A + 8 * s,
                                     umma_desc(b0 + s * 2 * SU_B_LBO, SU_B_LBO, SU_SBO), s > 0);
                    umma_commit(&sh.mma_done[j]);
                    SU_T1(w2)
                }
            }
        }
    } else if (warp >= SU_STG_WARP0 && warp < SU_STG_WARP0 + SU_STG_WARPS) {
        // ================= staging: raw symbols -> B operand (bf16 pieces, core-matrix layout)
        const int sw = warp - SU_STG_WARP0, w = lane & 15;
        for (long n = 0; n < my_batches; n++) {
            const int st = (int) (n % SU_STAGES), j = (int) (n & 1);
            const unsigned char *raw = sRaw + st * SU_RAW_BYTES;
            unsigned char *B = sB + j * SU_B_BYTES;
            { SU_T0 mbar_wait(&sh.raw_full[st], (uint32_t) (n / SU_STAGES) & 1u); SU_T1(w0) }
            if (n >= 2) { SU_T0 mbar_wait(&sh.mma_done[j], (uint32_t) ((n - 2) >> 1) & 1u); SU_T1(w1) }   // the MMAs of batch n - 2 have read B
            float sabs = 0.0f;
#pragma unroll
            for (int k = 0; k < 2; k++) {
                const int c = 16 * k + 2 * sw + (lane >> 4);            // 8-symbol chunk of the window
                const uint4 *src = reinterpret_cast<const uint4 *>(raw + w * SU_RAW_STRIDE + c * 64);
                uint32_t dh[4], dm[4], eh[4], em[4];                    // bf16 pairs (x, x + 1): low half = x
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const uint4 v = src[q];                             // symbols 8c + 2q, 8c + 2q + 1
                    float d0 = __fsub_rn(__uint_as_float(v.x), __uint_as_float(v.y));      // qpsk.c:88-96, pre = v(1+i)
                    float e0 = __fadd_rn(__uint_as_float(v.y), __uint_as_float(v.x));
                    float d1 = __fsub_rn(__uint_as_float(v.z), __uint_as_float(v.w));
                    float e1 = __fadd_rn(__uint_as_float(v.w), __uint_as_float(v.z));
                    if (c == 31 && q == 3) d1 = e1 = 0.0f;              // x = 255 is outside every lag's sum: P[.][255] = 0
                    sabs = __fadd_rn(sabs, __fadd_rn(__fadd_rn(fabsf(d0), fabsf(e0)), __fadd_rn(fabsf(d1), fabsf(e1))));
                    split2_pair(d0, d1, dh[q], dm[q]);
                    split2_pair(e0, e1, eh[q], em[q]);
                }
                // column n = 16 p + w: row group 2 p + (w >> 3), row w & 7; the 8 symbols are one 16-byte row
                unsigned char *dst = B + c * SU_B_LBO + (w >> 3) * 128 + (w & 7) * 16;
                *reinterpret_cast<uint4 *>(dst + 0 * 256) = make_uint4(dh[0], dh[1], dh[2], dh[3]);
                *reinterpret_cast<uint4 *>(dst + 1 * 256) = make_uint4(dm[0], dm[1], dm[2], dm[3]);
                *reinterpret_cast<uint4 *>(dst + 2 * 256) = make_uint4(eh[0], eh[1], eh[2], eh[3]);
                *reinterpret_cast<uint4 *>(dst + 3 * 256) = make_uint4(em[0], em[1], em[2], em[3]);
            }
            sabs = __fadd_rn(sabs, __shfl_xor_sync(0xffffffffu, sabs, 16));
            if (lane < 16) sh.part_abs[n & 3][sw][w] = sabs;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&sh.b_full[j]);
                mbar_arrive(&sh.raw_empty[st]);
            }
        }
    } else if (warp < 2 * SU_EPI_WARPS) {
        // ================= epilogue: |re|^2 + |im|^2 per lag and window, maxima, bound, candidate lists.  Two sets of
        // four warps, one per accumulator buffer: a batch's epilogue is ~600 dependent instructions on one warp per scheduler
        const int set = warp >> 2, qw = warp & 3, lag = tid & 127;
        for (long n = set; n < my_batches; n += 2) {
            const int j = set, slot = (int) (n % SU_LISTS);
            const long b = blockIdx.x + n * gridDim.x;
            { SU_T0 mbar_wait(&sh.mma_done[j], (uint32_t) (n >> 1) & 1u); SU_T1(w0) }
            tc_fence_after();
            const uint32_t taddr = tmem + ((uint32_t) (32 * qw) << 16) + SU_TMEM_D + (uint32_t) j * SU_N;
            float val[SU_WIN];
#ifdef SU_PROFILE
            const long long te0 = clock64();
#endif
            {
                float h[16], m[16];
                tmem_ld16(taddr + 0, h);
                tmem_ld16(taddr + 16, m);
                tmem_ld_wait();
#pragma unroll
                for (int w = 0; w < SU_WIN; w++) val[w] = __fadd_rn(h[w], m[w]);          // re
                tmem_ld16(taddr + 32, h);
                tmem_ld16(taddr + 48, m);
                tmem_ld_wait();
#pragma unroll
                for (int w = 0; w < SU_WIN; w++) {
                    const float im = __fadd_rn(h[w], m[w]);
                    val[w] = __fadd_rn(__fmul_rn(val[w], val[w]), __fmul_rn(im, im));
                }
            }
            tc_fence_before();
#ifdef SU_PROFILE
            w2 += clock64() - te0;
            const long long te1 = clock64();
#endif
#ifndef SU_PROFILE
            if (dbg_approx != nullptr) {
#pragma unroll
                for (int w = 0; w < SU_WIN; w++)
                    if (b * SU_WIN + w < n_streams) dbg_approx[(b * SU_WIN + w) * PRE + lag] = val[w];
            }
#endif
            // warp maxima (non-negative floats order like their bit patterns; a NaN sorts above everything, gives a NaN
            // threshold, no candidate, and ends in the verify warp's fallback)
            uint32_t wm = 0;
#pragma unroll
            for (int w = 0; w < SU_WIN; w++) {
                const uint32_t m = __reduce_max_sync(0xffffffffu, __float_as_uint(val[w]));
                if (lane == w) wm = m;
            }
            float s_abs = 0.0f;
            if (lane < SU_WIN) {
                sh.warp_max[set][qw][lane] = wm;
#pragma unroll
                for (int q = 0; q < SU_STG_WARPS; q++) s_abs = __fadd_rn(s_abs, sh.part_abs[n & 3][q][lane]);
            }
#ifdef SU_PROFILE
            w1 += clock64() - te1;
            const long long te2 = clock64();
#endif
            epi_bar(set);
#ifdef SU_PROFILE
            w3 += clock64() - te2;
            const long long te3 = clock64();
#endif
            if (lag == 0) mbar_arrive(&sh.d_empty[j]);                  // every warp has its accumulators in registers
            float thr = 0.0f;                                           // lane w < 16 of every warp: window w's threshold
            if (lane < SU_WIN) {
                const uint32_t m = max(max(sh.warp_max[set][0][lane], sh.warp_max[set][1][lane]),
                                       max(sh.warp_max[set][2][lane], sh.warp_max[set][3][lane]));
                // |approx - reference| per component <= delta = 2^-13 sum(|d| + |e|): truncation of the split 2^-16 per
                // piece pair, tensor-core accumulation of 128 non-zero terms in fp32 (<= 2^-15 even with truncating
                // adders), the reference's own rounding 127 * 2^-24 -- the bound of sc_search_mma.cuh
                thr = su_candidate_threshold(__uint_as_float(m), __fmul_rn(s_abs, 0x1.004p-13f));
            }
#ifdef SU_PROFILE
            w4 += clock64() - te3;
            const long long te4 = clock64();
#endif
            mbar_wait(&sh.cand_empty[slot], ((uint32_t) (n / SU_LISTS) & 1u) ^ 1u);
#ifdef SU_PROFILE
            w5 += clock64() - te4;
#endif
#pragma unroll
            for (int w = 0; w < SU_WIN; w++) {
                const float t = __shfl_sync(0xffffffffu, thr, w);
                const bool is = val[w] >= t;
                const unsigned m = __ballot_sync(0xffffffffu, is);
                const int pos = __popc(m & ((1u << lane) - 1u));
                if (is && pos < SM_MAX_CAND) sh.cand[slot][qw][w][pos] = (unsigned char) lag;
                if (lane == 0) sh.n_cand[slot][qw][w] = __popc(m);
            }
            epi_bar(set);
            if (lag == 0) mbar_arrive(&sh.cand_full[slot]);
        }
    } else if (warp >= SU_VER_WARP0) {
        // ================= verify: the reference's exact sums for the candidates; lane = (window, component)
        const int vw = warp - SU_VER_WARP0;
        const int w = lane >> 1, comp = lane & 1;
        float2 *scr = reinterpret_cast<float2 *>(sScr + vw * SU_SCR_BYTES) + w * SU_SCR_STRIDE;
        // a candidate's symbols from the even lag below it (65 16-byte chunks, every 32-byte sector once), L2 -> shared
        // memory, all in flight at once; the lane pair shares the work
        auto gather = [&](const float2 *win, int L) {
            const float2 *src = win + (L & ~1);
#pragma unroll
            for (int i = 0; i < 33; i++) {
                const int ch = 2 * i + comp;
                if (ch < 65) cp_async16(scr + 2 * ch, src + 2 * ch);
            }
        };
        auto exact_value = [&](int L) {                             // cnormf of the pair's two sums, qpsk.c:75-80
            const float part = su_exact_sum(scr + (L & 1), comp);
            const float sq = __fmul_rn(part, part);
            return __fadd_rn(sq, __shfl_xor_sync(0x3u << (lane & ~1), sq, 1));
        };
        // One round (an L2 round trip + 128 dependent adds) serves 16 candidates.  4 % of noise-only windows have a
        // second candidate -- every other batch has one -- so second candidates are not verified with their batch but
        // queued, and a full queue is verified in one round of its own; the window's result is then corrected in place
        // (only this warp ever touches it).  Third and later candidates (0.1 %) are verified at once.
        int qn = 0;                                                 // queued second candidates, the same in every lane
        uint32_t gphase = 0;                                        // parity of the next gather's mbarrier phase
        int2 *queue = sh.defer[vw];
        auto flush = [&]() {
            su_flush_queue(queue, qn, scr, symbols, symbol_stride, max_index, max_value, lane);
            qn = 0;
        };
        for (long n = vw; n < my_batches; n += SU_SLOTS) {
            const long b = blockIdx.x + n * gridDim.x;
            const int slot = (int) (n % SU_LISTS);
            { SU_T0 mbar_wait(&sh.cand_full[slot], (uint32_t) (n / SU_LISTS) & 1u); SU_T1(w0) }
#ifdef SU_PROFILE
            const long long tv0 = clock64();
#endif
            const bool exists = b * SU_WIN + w < n_streams;
            int cnt[SU_EPI_WARPS], nc = 0;
#pragma unroll
            for (int q = 0; q < SU_EPI_WARPS; q++) {
                cnt[q] = sh.n_cand[slot][q][w];
                nc += cnt[q];
            }
            const bool direct = exists && nc >= 1 && nc <= SM_MAX_CAND;
            const float2 *W = symbols + (b * SU_WIN + (exists ? w : 0)) * symbol_stride;
            auto kth = [&](int k) {                                     // the k-th candidate of this lane's window, in lag order
                int q = 0, p = k;
#pragma unroll
                for (int qq = 0; qq < SU_EPI_WARPS - 1; qq++)
                    if (q == qq && p >= cnt[qq]) {
                        p -= cnt[qq];
                        q = qq + 1;
                    }
                return (int) sh.cand[slot][q][w][p];
            };
            // largest exact value, smallest lag among equals == the reference's strict '>' scanning the lags upwards
            float ev = -1.0f;
            int ei = 1 << 20;
            // rounds for candidate ranks 0, 2, 3, ... (rank 1 goes to the queue)
            int rounds = direct ? (nc >= 3 ? nc : 1) : 0;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) rounds = max(rounds, __shfl_xor_sync(0xffffffffu, rounds, off));
#ifdef SU_PROFILE
            w3 += clock64() - tv0;
#endif
#pragma unroll 1
            for (int k = 0; k < rounds; k++) {
                if (k == 1) continue;
                const bool have = direct && k < nc;
                const int L = have ? kth(k) : 0;
                {
                    // the candidates' symbols from the even lag below each (130 symbols = 1,040 bytes, 16-byte aligned):
                    // one TMA bulk copy per candidate, L2 -> shared memory
                    SU_T0
                    const unsigned hm = __ballot_sync(0xffffffffu, have && comp == 0);
                    if (lane == 0) mbar_expect_tx(&sh.gathered[vw], (uint32_t) __popc(hm) * (uint32_t) (SU_SCR_STRIDE * 8));
                    __syncwarp();
                    if (have && comp == 0) tma_bulk_g2s(scr, W + (L & ~1), SU_SCR_STRIDE * 8, &sh.gathered[vw]);
                    mbar_wait(&sh.gathered[vw], gphase);
                    gphase ^= 1u;
                    SU_T1(w1)
                }
                {
                    SU_T0
                    if (have) {
                        const float v = exact_value(L);
                        if (v > ev || (v == ev && L < ei)) {
                            ev = v;
                            ei = L;
                        }
                    }
                    __syncwarp();
                    SU_T1(w2)
                }
            }
#ifdef SU_PROFILE
            const long long tv1 = clock64();
#endif
            const bool second = direct && nc >= 2;
            const int L1 = second ? kth(1) : 0;
            __syncwarp();
            if (lane == 0) mbar_arrive(&sh.cand_empty[slot]);           // the lists have been read: the slot is free
            if (!(ev > 0.0f)) ei = 0, ev = fmaxf(ev, 0.0f);
            // no candidate (NaNs) or too many (silence, ties over many lags): the full exact search, the warp per window
            unsigned fb = __ballot_sync(0xffffffffu, exists && !direct && comp == 0);
            while (fb) {
                const int l2 = __ffs(fb) - 1;
                fb &= fb - 1;
                int bi;
                float bv;
                su_search_warp(symbols + (b * SU_WIN + (l2 >> 1)) * symbol_stride, lane, bi, bv);
                if (w == (l2 >> 1)) {
                    ei = bi;
                    ev = bv;
                }
            }
            if (exists && comp == 0) {
                max_index[b * SU_WIN + w] = ei;
                max_value[b * SU_WIN + w] = ev;
            }
            // queue the second candidates (after the window's result is stored: the queue's round corrects it)
            const unsigned m2 = __ballot_sync(0xffffffffu, second && comp == 0);
            const int add = __popc(m2);
            if (qn + add > SU_WIN) flush();
            if (second && comp == 0) queue[qn + __popc(m2 & ((1u << lane) - 1u))] = make_int2((int) (b * SU_WIN + w), L1);
            qn += add;
            __syncwarp();
#ifdef SU_PROFILE
            w4 += clock64() - tv1;
#endif
        }
        if (qn > 0) flush();
    }

#ifdef SU_PROFILE
    if (blockIdx.x == 0 && lane == 0 && dbg_approx != nullptr) {       // [warp][total, wait0, wait1, wait2] of CTA 0
        float *o = dbg_approx + 8 * warp;
        o[0] = (float) (clock64() - t_begin);
        o[1] = (float) w0;
        o[2] = (float) w1;
        o[3] = (float) w2;
        o[4] = (float) w3;
        o[5] = (float) w4;
        o[6] = (float) w5;
    }
#endif
    (void) w0, (void) w1, (void) w2, (void) w3, (void) w4, (void) w5, (void) t_begin;
    tc_fence_before();
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t) SU_TMEM_COLS) : "memory");
}


bool search_umma_eligible(const float2 *symbols, long symbol_stride) {
    // TMA bulk copies want 16-byte aligned sources of 2,048 bytes: 256 symbols per window (the 256th is never used)
    return symbol_stride >= 2 * PRE && (symbol_stride & 1) == 0 && (((uintptr_t) symbols) & 15) == 0;
}

cudaError_t launch_search_umma_batch(long n_streams, const float2 *symbols, long symbol_stride,
                                     int *max_index, float *max_value, float *dbg_approx, cudaStream_t st) {
    static std::atomic<unsigned long long> configured{0};              // bit per device
    int dev = 0, sms = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev >= 64 || !((configured.load() >> dev) & 1ull)) {
        e = cudaFuncSetAttribute(search_umma_batch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SU_SMEM);
        if (e != cudaSuccess) return e;
        if (dev < 64) configured.fetch_or(1ull << dev);
    }
    e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return e;
    const long n_batches = (n_streams + SU_WIN - 1) / SU_WIN;
    const int grid = (int) std::min<long>(n_batches, sms);
    search_umma_batch_kernel<<<grid, SU_THREADS, SU_SMEM, st>>>(symbols, symbol_stride, max_index,
                                                                max_value, n_streams, dbg_approx);
    g_launch_count++;
    return cudaGetLastError();
}

}  // namespace sc
