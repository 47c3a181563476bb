// sc_umma.cuh -- the PTX of the 5th-generation tensor-core path shared by sc_search_umma.cu (the stage search) and
// sc_frontend_umma.cu (the fused front-end): mbarriers, TMA bulk copies, shared-memory matrix descriptors, tcgen05
// commit / fences, and the exact pieces of the proposer / verifier scheme of sc_search_mma.cuh that both kernels run
// on raw symbols (the two-piece bf16 split, the candidate threshold, the reference's sequential sums).
#pragma once
#include "sc_common.cuh"
#include "sc_tables.cuh"
#include "sc_search_mma.cuh"

namespace sc {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
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

// the same for a role with slack: backs off between polls so that the spin does not take issue slots from busy warps
__device__ __forceinline__ void mbar_wait_relaxed(unsigned long long *bar, uint32_t parity) {
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
        __nanosleep(100);
        if (spin > (1u << 22)) asm volatile("trap;");
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

__device__ __forceinline__ void umma_commit(unsigned long long *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// The reference's sum of one component over the 128 symbols X[0..127] of a candidate lag, gathered into shared memory
// beforehand.  comp 1: e = s.i + s.r; comp 0: d = s.r - s.i = s.r + (-s.i) -- one IEEE addition either way (addition
// commutes and negation is exact, so these are the bits the all-exact search forms), the sign of s.i chosen by a mask.
// Loads run one block of 16 ahead of the dependent chain of adds.
__device__ __forceinline__ float su_exact_sum(const float2 *__restrict__ X, int comp) {
    const uint32_t flip = comp ? 0u : 0x80000000u;
    float a = 0.0f;
    float2 cur[16], nxt[16];
#pragma unroll
    for (int i = 0; i < 16; i++) cur[i] = X[i];
#pragma unroll
    for (int blk = 0; blk < PRE; blk += 16) {
        if (blk + 16 < PRE) {
#pragma unroll
            for (int i = 0; i < 16; i++) nxt[i] = X[blk + 16 + i];
        }
        float x[16];                                                // d or e of the block first: the chain below is adds only
#pragma unroll
        for (int i = 0; i < 16; i++) x[i] = __fadd_rn(cur[i].x, __uint_as_float(__float_as_uint(cur[i].y) ^ flip));
#pragma unroll
        for (int i = 0; i < 16; i++) a = pre_neg(blk + i) ? __fsub_rn(a, x[i]) : __fadd_rn(a, x[i]);
#pragma unroll
        for (int i = 0; i < 16; i++) cur[i] = nxt[i];
    }
    return a;
}
// split2() of sc_search_mma.cuh for two values at once: hi = the values truncated to bf16 (cvt.rz packs both), mid =
// the truncated remainders; same pieces bit for bit (truncation of a float to its top 16 bits is round-toward-zero)
__device__ __forceinline__ void split2_pair(float a, float b, uint32_t &hi, uint32_t &mid) {
    asm("cvt.rz.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(b), "f"(a));             // upper half <- b, lower half <- a
    const float ra = __fsub_rn(a, __uint_as_float(hi << 16));                      // exact: the low 16 significand bits
    const float rb = __fsub_rn(b, __uint_as_float(hi & 0xffff0000u));
    asm("cvt.rz.bf16x2.f32 %0, %1, %2;" : "=r"(mid) : "f"(rb), "f"(ra));
}

// search_candidate_threshold() of sc_search_mma.cuh with the hardware square root (2 ulp; m carries a 1e-4 margin)
__device__ __forceinline__ float su_candidate_threshold(float vmax, float delta) {
    float r;
    asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(__fmul_rn(2.0f, vmax)));
    const float m = __fmul_rn(r, 1.0001f);
    const float mu = __fadd_rn(__fmul_rn(__fmul_rn(2.0f, delta), __fadd_rn(m, delta)), __fmul_rn(vmax, 0x1p-20f));
    return __fsub_rn(vmax, __fmul_rn(mu, 2.002f));
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

}  // namespace sc
