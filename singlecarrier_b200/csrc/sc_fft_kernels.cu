// sc_fft_kernels.cu -- batched mixed-radix FFT behind the reference's fft.h interface.
//
// The reference ships a kiss-fft derivative (src/fft.c) that none of its modem code calls
// (SURVEY F1); fft.h is nevertheless one of the drop-in headers, so this kernel reproduces that
// transform's data flow -- the same factorisation (radix 4, 2, 3, 5, then generic), the same
// decimation-in-time recursion flattened into stages, the same butterfly arithmetic in the same
// order with the same (host-libm) twiddles -- so results match src/fft.c to the last bit for
// finite data (every butterfly is an independent, deterministic expression; only the order in
// which independent butterflies run differs).  One CTA per transform; the working array lives in
// shared memory (n <= 4096) or in a global scratch slab (larger n).
//   kf_work / kf_factor   src/fft.c:388-459      kf_bfly2/4/3/5/generic  src/fft.c:190-386
//   encode_fftr/fftri     src/fft.c:139-186
#include "sc_common.cuh"
#include "sc_fft256.cuh"
#include "sc_kernels.h"

namespace sc {


// one butterfly of radix p at (base, u) -- mirrors kf_bfly2/3/4/5 of src/fft.c
__device__ __forceinline__ void bfly2(c32 *F, const c32 *tw, int base, int u, int m, int fs) {
    const c32 t = cmul(F[base + m], tw[fs * u]);
    const c32 f0 = F[base];
    F[base + m] = csub(f0, t);
    F[base] = cadd(f0, t);
}

__device__ __forceinline__ void bfly4(c32 *F, const c32 *tw, int base, int u, int m, int fs, int inverse) {
    const c32 s0 = cmul(F[base + m], tw[fs * u]);
    const c32 s1 = cmul(F[base + 2 * m], tw[2 * fs * u]);
    const c32 s2 = cmul(F[base + 3 * m], tw[3 * fs * u]);
    c32 f0 = F[base];
    const c32 s5 = csub(f0, s1);
    f0 = cadd(f0, s1);
    const c32 s3 = cadd(s0, s2);
    const c32 s4 = csub(s0, s2);
    F[base + 2 * m] = csub(f0, s3);
    F[base] = cadd(f0, s3);
    if (inverse) {
        F[base + m] = mk(__fsub_rn(s5.r, s4.i), __fadd_rn(s5.i, s4.r));
        F[base + 3 * m] = mk(__fadd_rn(s5.r, s4.i), __fsub_rn(s5.i, s4.r));
    } else {
        F[base + m] = mk(__fadd_rn(s5.r, s4.i), __fsub_rn(s5.i, s4.r));
        F[base + 3 * m] = mk(__fsub_rn(s5.r, s4.i), __fadd_rn(s5.i, s4.r));
    }
}

__device__ __forceinline__ void bfly3(c32 *F, const c32 *tw, int base, int u, int m, int fs) {
    const c32 epi3 = tw[fs * m];
    const c32 s1 = cmul(F[base + m], tw[fs * u]);
    const c32 s2 = cmul(F[base + 2 * m], tw[2 * fs * u]);
    const c32 s3 = cadd(s1, s2);
    c32 s0 = csub(s1, s2);
    const c32 f0 = F[base];
    c32 f1 = mk(__fsub_rn(f0.r, __fmul_rn(s3.r, 0.5f)), __fsub_rn(f0.i, __fmul_rn(s3.i, 0.5f)));
    s0 = cscale(s0, epi3.i);
    F[base] = cadd(f0, s3);
    F[base + 2 * m] = mk(__fadd_rn(f1.r, s0.i), __fsub_rn(f1.i, s0.r));
    F[base + m] = mk(__fsub_rn(f1.r, s0.i), __fadd_rn(f1.i, s0.r));
}

__device__ __forceinline__ void bfly5(c32 *F, const c32 *tw, int base, int u, int m, int fs) {
    const c32 ya = tw[fs * m], yb = tw[fs * 2 * m];
    const c32 s0 = F[base];
    const c32 s1 = cmul(F[base + m], tw[fs * u]);
    const c32 s2 = cmul(F[base + 2 * m], tw[2 * fs * u]);
    const c32 s3 = cmul(F[base + 3 * m], tw[3 * fs * u]);
    const c32 s4 = cmul(F[base + 4 * m], tw[4 * fs * u]);
    const c32 s7 = cadd(s1, s4), s10 = csub(s1, s4), s8 = cadd(s2, s3), s9 = csub(s2, s3);
    F[base] = mk(__fadd_rn(s0.r, __fadd_rn(s7.r, s8.r)), __fadd_rn(s0.i, __fadd_rn(s7.i, s8.i)));
    const c32 s5 = mk(__fadd_rn(__fadd_rn(s0.r, __fmul_rn(s7.r, ya.r)), __fmul_rn(s8.r, yb.r)),
                      __fadd_rn(__fadd_rn(s0.i, __fmul_rn(s7.i, ya.r)), __fmul_rn(s8.i, yb.r)));
    const c32 s6 = mk(-__fadd_rn(__fmul_rn(s10.i, ya.i), __fmul_rn(s9.i, yb.i)),
                      -__fsub_rn(__fmul_rn(s10.r, ya.i), __fmul_rn(s9.r, yb.i)));
    F[base + m] = csub(s5, s6);
    F[base + 4 * m] = cadd(s5, s6);
    const c32 s11 = mk(__fadd_rn(__fadd_rn(s0.r, __fmul_rn(s7.r, yb.r)), __fmul_rn(s8.r, ya.r)),
                       __fadd_rn(__fadd_rn(s0.i, __fmul_rn(s7.i, yb.r)), __fmul_rn(s8.i, ya.r)));
    const c32 s12 = mk(-__fadd_rn(__fmul_rn(s10.i, yb.i), __fmul_rn(s9.i, ya.i)),
                       __fsub_rn(__fmul_rn(s10.r, yb.i), __fmul_rn(s9.r, ya.i)));
    F[base + 2 * m] = cadd(s11, s12);
    F[base + 3 * m] = csub(s11, s12);
}

// mode 0: complex transform; 1: encode_fftr (real in -> n/2+1 complex out); 2: encode_fftri
__global__ void __launch_bounds__(256)
fft_kernel(FftPlan plan, const c32 *__restrict__ tw, const c32 *__restrict__ super_tw,
           const int *__restrict__ perm, int mode, const void *__restrict__ in, void *__restrict__ out,
           c32 *__restrict__ scratch, int use_smem, long n_batches) {
    extern __shared__ __align__(16) unsigned char fft_smem[];
    const int n = plan.n;                                  // complex length of the core transform
    c32 *F, *G;
    if (use_smem) {
        F = reinterpret_cast<c32 *>(fft_smem);
        G = F + n;
    } else {
        F = scratch + (long) blockIdx.x * 2 * n;
        G = F + n;
    }
    const int t = threadIdx.x, nt = blockDim.x;

    for (long b = blockIdx.x; b < n_batches; b += gridDim.x) {
        // ---- input (with encode_fftri's pre-processing, src/fft.c:166-183) into G, natural order
        if (mode == 2) {
            const c32 *f = reinterpret_cast<const c32 *>(in) + b * (long) (n + 1);
            if (t == 0) G[0] = mk(__fadd_rn(f[0].r, f[n].r), __fsub_rn(f[0].r, f[n].r));
            for (int k = 1 + t; k <= n / 2; k += nt) {
                const c32 fk = f[k], fnkc = cconj(f[n - k]);
                const c32 fek = cadd(fk, fnkc);
                const c32 fok = cmul(csub(fk, fnkc), super_tw[k - 1]);
                G[k] = cadd(fek, fok);
                G[n - k] = cconj(csub(fek, fok));
            }
        } else {
            const c32 *x = reinterpret_cast<const c32 *>(in) + b * (long) n;   // mode 1: n/2.. real pairs
            for (int i = t; i < n; i += nt) G[i] = x[i];
        }
        __syncthreads();
        // ---- leaves of kf_work (src/fft.c:399-404): digit-reversed gather through the host-made table
        for (int i = t; i < n; i += nt) F[i] = G[__ldg(perm + i)];
        __syncthreads();
        // ---- butterflies, innermost factor first (the recursion unwinds this way, src/fft.c:412-430)
        int fs = n;
        for (int l = plan.n_stages - 1; l >= 0; l--) {
            const int p = plan.p[l], m = plan.m[l];
            fs /= p;                                       // product of the radices before stage l
            if (p == 2 || p == 3 || p == 4 || p == 5) {
                for (int q = t; q < n / p; q += nt) {
                    const int blk = q / m, u = q - blk * m, base = blk * p * m + u;
                    if (p == 4) bfly4(F, tw, base, u, m, fs, plan.inverse);
                    else if (p == 2) bfly2(F, tw, base, u, m, fs);
                    else if (p == 3) bfly3(F, tw, base, u, m, fs);
                    else bfly5(F, tw, base, u, m, fs);
                }
            } else {                                       // kf_bfly_generic, src/fft.c:346-386
                for (int i = t; i < n; i += nt) G[i] = F[i];
                __syncthreads();
                for (int i = t; i < n; i += nt) {
                    const int blk = i / (p * m), k = i - blk * p * m, u = k % m;   // k = u + q1*m inside the block
                    const c32 *sc0 = G + blk * p * m + u;
                    c32 acc = sc0[0];
                    int twidx = 0;
                    for (int q = 1; q < p; q++) {
                        twidx += fs * k;
                        if (twidx >= n) twidx -= n;
                        acc = cadd(acc, cmul(sc0[q * m], tw[twidx]));
                    }
                    F[i] = acc;
                }
            }
            __syncthreads();
        }
        // ---- output (with encode_fftr's post-processing, src/fft.c:139-162)
        if (mode == 1) {
            c32 *f = reinterpret_cast<c32 *>(out) + b * (long) (n + 1);
            if (t == 0) {
                const c32 tdc = F[0];
                f[0] = mk(__fadd_rn(tdc.r, tdc.i), 0.0f);
                f[n] = mk(__fsub_rn(tdc.r, tdc.i), 0.0f);
            }
            for (int k = 1 + t; k <= n / 2; k += nt) {
                const c32 fpk = F[k], fpnk = cconj(F[n - k]);
                const c32 f1k = cadd(fpk, fpnk), f2k = csub(fpk, fpnk);
                const c32 w = cmul(f2k, super_tw[k - 1]);
                f[k] = mk(__fmul_rn(__fadd_rn(f1k.r, w.r), 0.5f), __fmul_rn(__fadd_rn(f1k.i, w.i), 0.5f));
                f[n - k] = mk(__fmul_rn(__fsub_rn(f1k.r, w.r), 0.5f), __fmul_rn(__fsub_rn(w.i, f1k.i), 0.5f));
            }
        } else {
            c32 *y = reinterpret_cast<c32 *>(out) + b * (long) n;
            for (int i = t; i < n; i += nt) y[i] = F[i];
        }
        __syncthreads();
    }
}

constexpr int FFT256_WARPS = 4;

template <bool INVERSE>
__global__ void __launch_bounds__(FFT256_WARPS * 32)
fft256_warp_kernel(const c32 *__restrict__ tw, const float2 *__restrict__ in, float2 *__restrict__ out, long n_batches) {
    const int lane = threadIdx.x & 31;
    const long warp = (long) blockIdx.x * FFT256_WARPS + (threadIdx.x >> 5);
    const long n_warps = (long) gridDim.x * FFT256_WARPS;
    Fft256Twiddles T;                                        // fetched once per warp
    T.load(tw, lane);
    c32 v[8], nx[8];
    if (warp < n_batches) {
        const float2 *x = in + warp * 256 + lane;
#pragma unroll
        for (int r = 0; r < 8; r++) nx[r] = from2(__ldg(x + 32 * r));
    }
    for (long b = warp; b < n_batches; b += n_warps) {
#pragma unroll
        for (int r = 0; r < 8; r++) v[r] = nx[r];
        if (b + n_warps < n_batches) {                       // next transform's loads fly during this one's math
            const float2 *x = in + (b + n_warps) * 256 + lane;
#pragma unroll
            for (int r = 0; r < 8; r++) nx[r] = from2(__ldg(x + 32 * r));
        }
        fft256_regs<INVERSE>(v, lane, T);
        // for a fixed register the 32 lanes cover 32 consecutive outputs: one 256-byte segment per store
        float2 *y = out + b * 256;
#pragma unroll
        for (int r = 0; r < 8; r++) y[fft256_out_index(lane, r)] = to2(v[r]);
    }
}

cudaError_t launch_fft(const FftPlan &plan, const float2 *tw, const float2 *super_tw, const int *perm, int mode,
                       const void *in, void *out, float2 *scratch, long n_batches, cudaStream_t st) {
    const int n = plan.n;
    const bool radix4x4 = plan.n_stages == 4 && plan.p[0] == 4 && plan.p[1] == 4 && plan.p[2] == 4 && plan.p[3] == 4;
    if (mode == 0 && n == 256 && radix4x4 && in != out) {                  // 4 x 4 x 4 x 4: register/warp-shuffle specialisation
        const long want = (n_batches + FFT256_WARPS - 1) / FFT256_WARPS;
        const int grid = (int) std::min<long>(want, 148L * 12);
        if (plan.inverse)
            fft256_warp_kernel<true><<<grid, FFT256_WARPS * 32, 0, st>>>(reinterpret_cast<const c32 *>(tw), (const float2 *) in,
                                                                          (float2 *) out, n_batches);
        else
            fft256_warp_kernel<false><<<grid, FFT256_WARPS * 32, 0, st>>>(reinterpret_cast<const c32 *>(tw), (const float2 *) in,
                                                                           (float2 *) out, n_batches);
        g_launch_count++;
        return cudaGetLastError();
    }
    const size_t smem = (size_t) 2 * n * sizeof(float2);
    const int use_smem = !fft_needs_scratch(n);
    int grid = (int) std::min<long>(n_batches, 148L * 4);
    if (!use_smem) grid = (int) std::min<long>(grid, FFT_SCRATCH_CTAS);
    if (use_smem && smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(fft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FFT_SMEM_LIMIT);
        if (e != cudaSuccess) return e;
    }
    const int threads = n >= 1024 ? 256 : (n >= 256 ? 128 : 64);
    fft_kernel<<<grid, threads, use_smem ? smem : 0, st>>>(plan, reinterpret_cast<const c32 *>(tw),
                                                           reinterpret_cast<const c32 *>(super_tw), perm, mode, in, out,
                                                           reinterpret_cast<c32 *>(scratch), use_smem, n_batches);
    g_launch_count++;
    return cudaGetLastError();
}

}  // namespace sc
