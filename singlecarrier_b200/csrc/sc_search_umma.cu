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
//   * one persistent CTA per SM works on batches of 16 windows; the windows arrive by cp.async.bulk (TMA, one 2 KB
//     copy per window, four batches in flight), so no thread ever waits for HBM;
//   * the 128 threads turn a batch into the B operand, N = 64 columns (4 pieces x 16 windows) x K = 256, bf16, in the
//     K-major no-swizzle core-matrix layout the tensor core reads from shared memory;
//   * one thread issues 16 tcgen05.mma (M = 128 lags, N = 64, K = 16 each): D[128 x 64] fp32 in tensor memory,
//     completion signalled on an mbarrier by tcgen05.commit.  While they run the CTA finishes the previous batch;
//   * the A operand is NOT the 64 KB matrix: slice s of P (128 lags x 16 symbols) is the same 368 x 16 "master"
//     matrix M[r][k] = pre[k - r] read from row 240 - 16 s on, and a row offset that is a multiple of 8 is just
//     another start address in the shared-memory descriptor -- 11.5 KB of shared memory serve all 16 slices;
//   * epilogue: thread L (lag) reads its 64 accumulators with tcgen05.ld, forms |re|^2 + |im|^2 for the 16 windows,
//     redux.sync gives the warp maxima, the bound gives the threshold, the (few) candidates are listed and verified
//     exactly from the fp32 samples still in shared memory.
//
// The shared-memory (LSU) data pipe, which bounds the mma.sync kernel at 85 %, carries one 2 KB store and one 2 KB
// load per window here; fragments never pass through registers.
#include "sc_common.cuh"
#include "sc_tables.cuh"
#include "sc_search_mma.cuh"
#include "sc_kernels.h"

namespace sc {

constexpr int SU_WIN = 16;                           // windows per batch
constexpr int SU_N = 4 * SU_WIN;                     // MMA N: column p * 16 + w, p = d_hi, d_mid, e_hi, e_mid
constexpr int SU_KSTEPS = 2 * PRE / 16;              // 16 MMAs of K = 16
constexpr int SU_THREADS = 128;                      // thread = lag in the epilogue (4 warps = the 4 TMEM lane quarters)
constexpr int SU_STAGES = 4;                         // raw-sample buffers (TMA in flight)
constexpr int SU_WIN_BYTES = 2 * PRE * 8;            // 256 complex floats
constexpr int SU_RAW_STRIDE = SU_WIN_BYTES + 16;     // per window in a raw buffer: 8 windows apart = 8 different 16-byte bank groups
constexpr int SU_RAW_BYTES = SU_WIN * SU_RAW_STRIDE;
constexpr int SU_A_GROUPS = (SU_A_ROWS + 7) / 8;     // 46 groups of 8 master rows
constexpr int SU_A_LBO = SU_A_GROUPS * 128;          // bytes between the two 8-symbol halves of the master
constexpr int SU_A_BYTES = 2 * SU_A_LBO;             // 11,776
constexpr int SU_B_LBO = (SU_N / 8) * 128;           // 1,024: bytes between 8-symbol chunks of B
constexpr int SU_B_BYTES = (2 * PRE / 8) * SU_B_LBO; // 32 KB per batch
constexpr int SU_SBO = 128;                          // 8 rows x 16 bytes: one core matrix
constexpr int SU_TMEM_COLS = 2 * SU_N;               // two accumulator buffers
static_assert(SU_A_BYTES == 16 * SU_A_WORDS4, "master size");

struct SuShared {                                    // behind the big buffers
    unsigned long long raw_full[SU_STAGES];          // mbarriers: a raw buffer has landed
    unsigned long long mma_done[2];                  // mbarriers: an accumulator buffer is complete
    uint32_t tmem_base;
    uint32_t warp_max[4][SU_WIN];
    float part_abs[2][4][SU_WIN];                    // sum(|d| + |e|) per warp and window, by batch parity
    float thr[SU_WIN];
    int n_cand[SU_WIN];
    int cand[SU_WIN][SM_MAX_CAND];
};
constexpr int SU_SMEM = SU_A_BYTES + 2 * SU_B_BYTES + SU_STAGES * SU_RAW_BYTES + (int) sizeof(SuShared) + 128;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// Spin on a phase of an mbarrier.  A barrier that never completes is a bug of this file, not a state to wait out:
// after ~2 s the kernel traps instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    for (uint32_t spin = 0;; spin++) {
        uint32_t done;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.b32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (done) return;
        if (spin > (1u << 24)) asm volatile("trap;");
    }
}
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, uint32_t bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// Shared-memory matrix descriptor, K-major, no swizzle (cute::UMMA::SmemDescriptor): start address, the byte offset
// between the two 8-element K halves of an MMA (leading), between groups of 8 rows (stride), all in 16-byte units;
// bits 46-47 = 1 (sm_100 descriptor version).
__device__ __forceinline__ uint64_t umma_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t) ((addr >> 4) & 0x3fffu) | ((uint64_t) (lbo >> 4) << 16) | ((uint64_t) (sbo >> 4) << 32) | (1ull << 46);
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): D fp32, A and B bf16, both K-major, N = 64, M = 128
constexpr uint32_t SU_IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t) (SU_N >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(a_desc), "l"(b_desc), "r"(SU_IDESC), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(unsigned long long *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// this thread's TMEM lane (row of D), 16 consecutive columns
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; i++) v[i] = __uint_as_float(r[i]);
}

// The reference's sum for lag L of one component, from the raw symbols (d = s.r - s.i, e = s.i + s.r are formed with
// the same two operations the all-exact search uses; the lane pair of a candidate reads the same addresses)
__device__ __forceinline__ float su_exact_sum(const float2 *__restrict__ W, int L, int comp) {
    const float2 *p = W + L;
    float a = 0.0f;
#pragma unroll
    for (int i = 0; i < PRE; i++) {
        const float2 s = p[i];
        const float x = comp ? __fadd_rn(s.y, s.x) : __fsub_rn(s.x, s.y);
        a = pre_neg(i) ? __fsub_rn(a, x) : __fadd_rn(a, x);
    }
    return a;
}

// search_verify16() of sc_search_mma.cuh on raw symbols
__device__ __forceinline__ void su_verify16(const float2 *__restrict__ W, const int *__restrict__ cand, int nc, int lane,
                                            int &ei, float &ev) {
    const int vk = (lane & 15) >> 1, vc = lane & 1;
    const bool have = vk < nc;
    const int L = have ? cand[vk] : 0;
    const float part = su_exact_sum(W, L, vc);
    const float sq = __fmul_rn(part, part);
    ev = __fadd_rn(sq, __shfl_xor_sync(0xffffffffu, sq, 1));           // cnormf, qpsk.c:75-80
    ei = L;
    if (!have) {
        ev = -1.0f;
        ei = 1 << 20;
    }
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

// search_warp_unpadded() of sc_search_mma.cuh on raw symbols: the full exact search by one warp (rare fallback)
__device__ __forceinline__ void su_search_warp(const float2 *__restrict__ W, int lane, int &best_idx, float &best_val) {
    const int comp = lane >> 4, g = lane & 15;
    const float2 *p = W + 8 * g;
    float a[8];
#pragma unroll
    for (int q = 0; q < 8; q++) a[q] = 0.0f;
#pragma unroll 1
    for (int j0 = 0; j0 < PRE + 8; j0 += 8) {
#pragma unroll
        for (int jj = 0; jj < 8; jj++) {
            const int j = j0 + jj;
            if (j >= PRE + 7) break;
            const float2 s = p[j];
            const float v = comp ? __fadd_rn(s.y, s.x) : __fsub_rn(s.x, s.y);
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

__global__ void __launch_bounds__(SU_THREADS, 1)
search_umma_batch_kernel(const float2 *__restrict__ symbols, long symbol_stride, const uint4 *__restrict__ a_master,
                         int *__restrict__ max_index, float *__restrict__ max_value, long n_streams,
                         float *__restrict__ dbg_approx) {
    extern __shared__ __align__(128) unsigned char su_smem[];
    unsigned char *sA = su_smem;
    unsigned char *sB = sA + SU_A_BYTES;
    unsigned char *sRaw = sB + 2 * SU_B_BYTES;
    SuShared &sh = *reinterpret_cast<SuShared *>(sRaw + SU_STAGES * SU_RAW_BYTES);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long n_batches = (n_streams + SU_WIN - 1) / SU_WIN;
    const long my_batches = blockIdx.x < n_batches ? (n_batches - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

    // ---- set-up: barriers, tensor memory, the master of the A operand
    if (tid == 0) {
        for (int i = 0; i < SU_STAGES; i++) mbar_init(&sh.raw_full[i], 1);
        mbar_init(&sh.mma_done[0], 1);
        mbar_init(&sh.mma_done[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sh.tmem_base)),
                     "r"((uint32_t) SU_TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = tid; i < SU_A_WORDS4; i += SU_THREADS) reinterpret_cast<uint4 *>(sA)[i] = __ldg(a_master + i);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // the master is read by the tensor core
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = sh.tmem_base;

    auto issue_tma = [&](long it) {                                     // thread 0: batch `it` of this CTA -> raw[it % STAGES]
        const long b = blockIdx.x + it * gridDim.x;
        const int nw = (int) min((long) SU_WIN, n_streams - b * SU_WIN);
        unsigned long long *bar = &sh.raw_full[it % SU_STAGES];
        mbar_expect_tx(bar, (uint32_t) nw * SU_WIN_BYTES);
        unsigned char *dst = sRaw + (it % SU_STAGES) * SU_RAW_BYTES;
        for (int w = 0; w < nw; w++)
            tma_bulk_g2s(dst + w * SU_RAW_STRIDE, symbols + (b * SU_WIN + w) * symbol_stride, SU_WIN_BYTES, bar);
    };
    if (tid == 0)
        for (long it = 0; it < min((long) SU_STAGES - 1, my_batches); it++) issue_tma(it);

    for (long it = 0; it <= my_batches; it++) {
        if (it < my_batches) {
            // ---- batch `it`: raw symbols -> B operand (bf16 pieces, core-matrix layout) -> 16 MMAs
            const unsigned char *raw = sRaw + (it % SU_STAGES) * SU_RAW_BYTES;
            unsigned char *B = sB + (it & 1) * SU_B_BYTES;
            mbar_wait(&sh.raw_full[it % SU_STAGES], (uint32_t) (it / SU_STAGES) & 1u);
            const int w = lane & 15;
            float sabs = 0.0f;
#pragma unroll 1
            for (int k = 0; k < 4; k++) {
                const int c = 8 * k + 2 * warp + (lane >> 4);           // 8-symbol chunk of the window
                const uint4 *src = reinterpret_cast<const uint4 *>(raw + w * SU_RAW_STRIDE + c * 64);
                uint32_t dh[8], dm[8], eh[8], em[8];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const uint4 v = src[j];                             // symbols 8c + 2j, 8c + 2j + 1
                    float d0 = __fsub_rn(__uint_as_float(v.x), __uint_as_float(v.y));      // qpsk.c:88-96, pre = v(1+i)
                    float e0 = __fadd_rn(__uint_as_float(v.y), __uint_as_float(v.x));
                    float d1 = __fsub_rn(__uint_as_float(v.z), __uint_as_float(v.w));
                    float e1 = __fadd_rn(__uint_as_float(v.w), __uint_as_float(v.z));
                    if (c == 31 && j == 3) d1 = e1 = 0.0f;              // x = 255 is outside every lag's sum: P[.][255] = 0
                    sabs = __fadd_rn(sabs, __fadd_rn(__fadd_rn(fabsf(d0), fabsf(e0)), __fadd_rn(fabsf(d1), fabsf(e1))));
                    split2(d0, dh[2 * j], dm[2 * j]);
                    split2(d1, dh[2 * j + 1], dm[2 * j + 1]);
                    split2(e0, eh[2 * j], em[2 * j]);
                    split2(e1, eh[2 * j + 1], em[2 * j + 1]);
                }
                // column n = 16 p + w: row group 2 p + (w >> 3), row w & 7; the 8 symbols are one 16-byte row
                unsigned char *dst = B + c * SU_B_LBO + (w >> 3) * 128 + (w & 7) * 16;
#define SU_PACK(a) make_uint4(__byte_perm(a[0], a[1], 0x7632), __byte_perm(a[2], a[3], 0x7632), \
                              __byte_perm(a[4], a[5], 0x7632), __byte_perm(a[6], a[7], 0x7632))
                *reinterpret_cast<uint4 *>(dst + 0 * 256) = SU_PACK(dh);
                *reinterpret_cast<uint4 *>(dst + 1 * 256) = SU_PACK(dm);
                *reinterpret_cast<uint4 *>(dst + 2 * 256) = SU_PACK(eh);
                *reinterpret_cast<uint4 *>(dst + 3 * 256) = SU_PACK(em);
#undef SU_PACK
            }
            sabs = __fadd_rn(sabs, __shfl_xor_sync(0xffffffffu, sabs, 16));
            if (lane < 16) sh.part_abs[it & 1][warp][w] = sabs;        // read by the epilogue of this batch, next iteration
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            tc_fence_before();
            __syncthreads();
            if (tid == 0) {
                tc_fence_after();
                const uint32_t a0 = smem_u32(sA), b0 = smem_u32(B);
#pragma unroll 1
                for (int s = 0; s < SU_KSTEPS; s++)
                    umma_bf16(tmem + (uint32_t) (it & 1) * SU_N, umma_desc(a0 + (30 - 2 * s) * 128, SU_A_LBO, SU_SBO),
                              umma_desc(b0 + s * 2 * SU_B_LBO, SU_B_LBO, SU_SBO), s > 0);
                umma_commit(&sh.mma_done[it & 1]);
            }
        }
        if (it >= 1) {
            // ---- epilogue of batch it - 1 (its MMAs ran while the block above staged batch `it`)
            const long pit = it - 1;
            const long b = blockIdx.x + pit * gridDim.x;
            const unsigned char *raw = sRaw + (pit % SU_STAGES) * SU_RAW_BYTES;
            mbar_wait(&sh.mma_done[pit & 1], (uint32_t) (pit >> 1) & 1u);
            tc_fence_after();
            const uint32_t taddr = tmem + ((uint32_t) (32 * warp) << 16) + (uint32_t) (pit & 1) * SU_N;
            float val[SU_WIN];
            {
                float dh[16], dm[16];
                tmem_ld16(taddr + 0, dh);
                tmem_ld16(taddr + 16, dm);
#pragma unroll
                for (int w = 0; w < SU_WIN; w++) val[w] = __fadd_rn(dh[w], dm[w]);        // re
                tmem_ld16(taddr + 32, dh);
                tmem_ld16(taddr + 48, dm);
#pragma unroll
                for (int w = 0; w < SU_WIN; w++) {
                    const float im = __fadd_rn(dh[w], dm[w]);
                    val[w] = __fadd_rn(__fmul_rn(val[w], val[w]), __fmul_rn(im, im));
                }
            }
            tc_fence_before();
            if (dbg_approx != nullptr) {
#pragma unroll
                for (int w = 0; w < SU_WIN; w++)
                    if (b * SU_WIN + w < n_streams) dbg_approx[(b * SU_WIN + w) * PRE + tid] = val[w];
            }
            // warp maxima (non-negative floats order like their bit patterns; a NaN sorts above everything and ends
            // in the fallback below)
            uint32_t wm = 0;
#pragma unroll
            for (int w = 0; w < SU_WIN; w++) {
                const uint32_t m = __reduce_max_sync(0xffffffffu, __float_as_uint(val[w]));
                if (lane == w) wm = m;
            }
            if (lane < SU_WIN) {
                sh.warp_max[warp][lane] = wm;
                if (warp == 0) sh.n_cand[lane] = 0;
            }
            __syncthreads();
            if (tid < SU_WIN) {
                const uint32_t m = max(max(sh.warp_max[0][tid], sh.warp_max[1][tid]), max(sh.warp_max[2][tid], sh.warp_max[3][tid]));
                const float(&pa)[4][SU_WIN] = sh.part_abs[pit & 1];
                const float s_abs = __fadd_rn(__fadd_rn(pa[0][tid], pa[1][tid]), __fadd_rn(pa[2][tid], pa[3][tid]));
                // |approx - reference| per component <= delta: truncation of the split 2^-14, tensor-core accumulation
                // (fp32, 128 non-zero terms, truncating adders allowed for) 2^-14, the reference's own rounding
                // 127 * 2^-24, all relative to sum(|d| + |e|) -- twice the margin of the mma.sync kernel
                sh.thr[tid] = search_candidate_threshold(__uint_as_float(m), __fmul_rn(s_abs, 0x1.004p-12f));
            }
            __syncthreads();
#pragma unroll
            for (int w = 0; w < SU_WIN; w++) {
                if (val[w] >= sh.thr[w]) {
                    const int pos = atomicAdd(&sh.n_cand[w], 1);
                    if (pos < SM_MAX_CAND) sh.cand[w][pos] = tid;
                }
            }
            __syncthreads();
            // ---- exact verification: half-warp h takes windows h and h + 8
#pragma unroll 1
            for (int r = 0; r < 2; r++) {
                const int w = 2 * warp + (lane >> 4) + 8 * r;
                const bool exists = b * SU_WIN + w < n_streams;
                const int nc = exists ? sh.n_cand[w] : 1;
                const bool direct = nc >= 1 && nc <= SM_MAX_CAND;
                const float2 *W = reinterpret_cast<const float2 *>(raw + w * SU_RAW_STRIDE);
                int ei;
                float ev;
                su_verify16(W, sh.cand[w], direct ? nc : 0, lane, ei, ev);
                // no candidate (NaNs) or too many (silence, ties over many lags): the full exact search, a warp per window
#pragma unroll 1
                for (int half = 0; half < 2; half++) {
                    const int fb = __shfl_sync(0xffffffffu, (int) (exists && !direct), 16 * half);
                    if (fb) {
                        const int w2 = 2 * warp + half + 8 * r;
                        int bi;
                        float bv;
                        su_search_warp(reinterpret_cast<const float2 *>(raw + w2 * SU_RAW_STRIDE), lane, bi, bv);
                        if ((lane >> 4) == half) {
                            ei = bi;
                            ev = bv;
                        }
                    }
                }
                if (exists && (lane & 15) == 0) {
                    max_index[b * SU_WIN + w] = ei;
                    max_value[b * SU_WIN + w] = ev;
                }
            }
        }
        __syncthreads();
        // the raw buffer of batch it - 1 is free again: batch it + STAGES - 1 goes there
        if (tid == 0 && it + SU_STAGES - 1 < my_batches) issue_tma(it + SU_STAGES - 1);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t) SU_TMEM_COLS) : "memory");
}

// The master of the A operand, host side: M[r][k] = pre[k - r] for r = -240 .. 127 (row index r + 240), k < 16, as bf16
// in the core-matrix layout: byte (k / 8) * LBO + (row / 8) * 128 + (row % 8) * 16 + (k % 8) * 2.
void search_umma_make_master(uint16_t *table /* [SU_A_WORDS4 * 8] */) {
    for (int i = 0; i < SU_A_WORDS4 * 8; i++) table[i] = 0;
    for (int row = 0; row < SU_A_ROWS; row++) {
        for (int k = 0; k < 16; k++) {
            const int i = k - (row - 240);
            if (i < 0 || i >= PRE) continue;
            const int byte = (k / 8) * SU_A_LBO + (row / 8) * 128 + (row % 8) * 16 + (k % 8) * 2;
            table[byte / 2] = pre_neg(i) ? 0xBF80 : 0x3F80;             // -1.0 / +1.0
        }
    }
}

bool search_umma_eligible(const float2 *symbols, long symbol_stride) {
    // TMA bulk copies want 16-byte aligned sources of 2,048 bytes: 256 symbols per window (the 256th is never used)
    return symbol_stride >= 2 * PRE && (symbol_stride & 1) == 0 && (((uintptr_t) symbols) & 15) == 0;
}

cudaError_t launch_search_umma_batch(long n_streams, const float2 *symbols, long symbol_stride, const void *a_master,
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
    search_umma_batch_kernel<<<grid, SU_THREADS, SU_SMEM, st>>>(symbols, symbol_stride, (const uint4 *) a_master, max_index,
                                                                max_value, n_streams, dbg_approx);
    g_launch_count++;
    return cudaGetLastError();
}

}  // namespace sc
