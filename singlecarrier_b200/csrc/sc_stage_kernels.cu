// sc_stage_kernels.cu -- batched forms of the reference's L1 primitives as stand-alone kernels
// (the stage entry points of include/singlecarrier_b200.h and the back end of the drop-in symbols).
//   fir_batch10_kernel   fir()                            src/fir.c:22-44
//   search_batch_kernel  correlate() + argmax             src/qpsk.c:88-96, 172-183
//   track_window_kernel  kalman_reset .. data_eq loop     src/qpsk.c:186-238
#include "sc_common.cuh"
#include "sc_tables.cuh"
#include "sc_tracker.cuh"
#include "sc_search.cuh"
#include "sc_search_mma.cuh"
#include "sc_fft256.cuh"
#include "sc_track_core.cuh"
#include "sc_kernels.h"

#include <math.h>
#include <stdlib.h>

namespace sc {

// ------------------------------------------------------------------------------------------------
// fir(): 49-tap real-coefficient FIR on complex samples, in place, caller-owned delay line holding
// the RAW past inputs, output scaled by GAIN.  One WARP per stream walks the samples in tiles with the
// last 49 raw inputs carried in shared memory between tiles, so each sample is read and written once
// and only __syncwarp() is needed.
// ext[0..48] = memory[0..48] (oldest first), ext[49 + j] = sample[j];
// out[j] = GAIN * sum_k ext[j + 1 + k] * coeff[k], k ascending (src/fir.c:36-42).
//
// FAST = true is the explicitly named tolerance mode (SC_FIR_FAST): the multiply-add is contracted
// into one FFMA2, which halves the FP32 work, at the price of last-bit differences from the
// reference (never used on the parity path).
//
// Tile pipeline of one warp (both modes):
//   registers  <- global   the NEXT tile's raw samples, issued before this tile's arithmetic so the DRAM latency
//                          hides behind it (cp.async into rotating buffers was measured 8 % slower);
//   ext[49..]  <- registers, __syncwarp, the window loads and the packed multiply-adds;
//   obuf       <- results  (obuf is ext + 49: every lane has finished reading by then), carry -> ext[0..48];
//   global     <- obuf     coalesced stores (a lane's own outputs are 40 / 80 bytes apart: stored directly they
//                          would cost several partial-sector stores per 32-byte sector).
//
// 10 outputs per lane (tiles of 320) and 128-bit shared-memory loads: each staged sample is read by 5.8 lanes
// (10.6 with 5 outputs per lane, the round-1 form), which is what the contracted mode needs -- with half the
// arithmetic per sample it was bound by the shared-memory data pipe (ncu: 90 %), not by the FP32 pipe.
// buf[1 + i] holds ext[i], so that every lane's window (10 lane + 2) and its 10 outputs (50 + 10 lane) start on a
// 16-byte boundary; lane stride 80 bytes = 5 x 16 is odd in 16-byte units, hence conflict-free for LDS.128 / STS.128.
// ------------------------------------------------------------------------------------------------
constexpr int FIR10_WARPS = 4;
constexpr int FIR10_R = 10;
constexpr int FIR10_TILE = 32 * FIR10_R;                // 320
constexpr int FIR10_BUF = 2 + NTAPS + FIR10_TILE + 9;   // 380 slots

// VEC: every stream starts on a 16-byte boundary (checked by the launcher), so global memory is moved two
// complex samples (128 bits) per lane and instruction; otherwise one sample (64 bits) at a time.
template <bool WIDE, bool FAST, bool VEC>
__global__ void __launch_bounds__(FIR10_WARPS * 32, 6)
fir_batch10_kernel(float2 *__restrict__ memory, float2 *__restrict__ sample, long sample_stride, int length,
                   long n_streams) {
    __shared__ __align__(16) float2 buf_all[FIR10_WARPS][FIR10_BUF];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float2 *ext = buf_all[warp] + 1;                        // ext[0..48] carry, ext[49..368] tile
    float2 *obuf = ext + NTAPS;
    constexpr int NV = VEC ? FIR10_R / 2 : FIR10_R;         // loads per lane and tile
    for (long s = (long) blockIdx.x * FIR10_WARPS + warp; s < n_streams; s += (long) gridDim.x * FIR10_WARPS) {
        float2 *mem = memory + s * NTAPS;
        float2 *x = sample + s * sample_stride;
        __syncwarp();
        for (int i = lane; i < NTAPS; i += 32) ext[i] = mem[i];
        float4 nxt[NV];                                     // VEC: two samples; else .x/.y only
        auto fetch = [&](int t0) {
#pragma unroll
            for (int k = 0; k < NV; k++) {
                nxt[k] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (VEC) {
                    const int j = t0 + 2 * (lane + 32 * k);
                    if (j + 1 < length) nxt[k] = *reinterpret_cast<const float4 *>(x + j);
                    else if (j < length) {
                        const float2 v = x[j];
                        nxt[k] = make_float4(v.x, v.y, 0.f, 0.f);
                    }
                } else {
                    const int j = t0 + lane + 32 * k;
                    if (j < length) {
                        const float2 v = x[j];
                        nxt[k] = make_float4(v.x, v.y, 0.f, 0.f);
                    }
                }
            }
        };
        fetch(0);
        for (int t0 = 0; t0 < length; t0 += FIR10_TILE) {
            const int n = min(FIR10_TILE, length - t0);
#pragma unroll
            for (int k = 0; k < NV; k++) {
                if (VEC) *reinterpret_cast<float4 *>(obuf + 2 * (lane + 32 * k)) = nxt[k];
                else obuf[lane + 32 * k] = make_float2(nxt[k].x, nxt[k].y);
            }
            __syncwarp();
            fetch(t0 + FIR10_TILE);                             // the next tile's loads fly during the math
            u64 acc[FIR10_R];
#pragma unroll
            for (int r = 0; r < FIR10_R; r++) acc[r] = 0ull;
            {
                // output r of the lane is sum_k ext[10 lane + r + 1 + k] * c[k]: window ext[10 lane + 1 ..] = buf[10 lane + 2 ..]
                const ulonglong2 *wp = reinterpret_cast<const ulonglong2 *>(buf_all[warp] + FIR10_R * lane + 2);
#pragma unroll
                for (int jj = 0; jj < (NTAPS + FIR10_R - 1) / 2; jj++) {           // 29 x 128-bit = 58 samples
                    const ulonglong2 v2 = wp[jj];
#pragma unroll
                    for (int half = 0; half < 2; half++) {
                        const int j = 2 * jj + half;
                        const u64 v = half ? v2.y : v2.x;
#pragma unroll
                        for (int r = 0; r < FIR10_R; r++) {
                            const int k = j - r;
                            if (k >= 0 && k < NTAPS) {
                                // tolerance mode: the output gain is folded into the taps (one rounding fewer)
                                if (FAST) acc[r] = pk_fma_bcast(v, tap<WIDE>(k) * FIR_GAIN, acc[r]);
                                else acc[r] = pk_add(acc[r], pk_mul_bcast_pz(v, tap<WIDE>(k)));
                            }
                        }
                    }
                }
            }
            float2 c0 = ext[n + lane], c1 = make_float2(0.f, 0.f);
            if (lane + 32 < NTAPS) c1 = ext[n + lane + 32];
            __syncwarp();
#pragma unroll
            for (int r = 0; r < FIR10_R; r += 2) {
                float ar, ai, br, bi;
                unpk(acc[r], ar, ai);
                unpk(acc[r + 1], br, bi);
                if (FAST)
                    *reinterpret_cast<float4 *>(obuf + FIR10_R * lane + r) = make_float4(ar, ai, br, bi);
                else
                    *reinterpret_cast<float4 *>(obuf + FIR10_R * lane + r) =
                        make_float4(__fmul_rn(ar, FIR_GAIN), __fmul_rn(ai, FIR_GAIN), __fmul_rn(br, FIR_GAIN), __fmul_rn(bi, FIR_GAIN));
            }
            ext[lane] = c0;
            if (lane + 32 < NTAPS) ext[lane + 32] = c1;
            __syncwarp();
#pragma unroll
            for (int k = 0; k < NV; k++) {
                if (VEC) {
                    const int j = 2 * (lane + 32 * k);
                    if (j + 1 < n) *reinterpret_cast<float4 *>(x + t0 + j) = *reinterpret_cast<const float4 *>(obuf + j);
                    else if (j < n) x[t0 + j] = obuf[j];
                } else {
                    const int j = lane + 32 * k;
                    if (j < n) x[t0 + j] = obuf[j];
                }
            }
        }
        __syncwarp();
        for (int i = lane; i < NTAPS; i += 32) mem[i] = ext[i];
    }
}

cudaError_t launch_fir_batch(bool wide, long n_streams, float2 *memory, float2 *sample, long sample_stride,
                             int length, cudaStream_t st, bool fast) {
    const int grid = (int) std::min<long>((n_streams + FIR10_WARPS - 1) / FIR10_WARPS, 148L * 6);
    const int thr = FIR10_WARPS * 32;
    const bool vec = (((uintptr_t) sample) & 15) == 0 && (sample_stride & 1) == 0;
#define SC_FIR_LAUNCH(W, F, V) fir_batch10_kernel<W, F, V><<<grid, thr, 0, st>>>(memory, sample, sample_stride, length, n_streams)
    if (wide) {
        if (fast) { if (vec) SC_FIR_LAUNCH(true, true, true); else SC_FIR_LAUNCH(true, true, false); }
        else { if (vec) SC_FIR_LAUNCH(true, false, true); else SC_FIR_LAUNCH(true, false, false); }
    } else {
        if (fast) { if (vec) SC_FIR_LAUNCH(false, true, true); else SC_FIR_LAUNCH(false, true, false); }
        else { if (vec) SC_FIR_LAUNCH(false, false, true); else SC_FIR_LAUNCH(false, false, false); }
    }
#undef SC_FIR_LAUNCH
    g_launch_count++;
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// correlate() over lags 0..127 + the reference's argmax, one warp per stream.
// ------------------------------------------------------------------------------------------------
constexpr int SB_WARPS = 8;

__global__ void __launch_bounds__(SB_WARPS * 32)
search_batch_kernel(const float2 *__restrict__ symbols, long symbol_stride, int *__restrict__ max_index,
                    float *__restrict__ max_value, long n_streams) {
    __shared__ __align__(16) float de[SB_WARPS][SEARCH_WORDS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (long s = (long) blockIdx.x * SB_WARPS + warp; s < n_streams; s += (long) gridDim.x * SB_WARPS) {
        const float2 *x = symbols + s * symbol_stride;
        __syncwarp();
        for (int i = lane; i < SEARCH_SYMS; i += 32) de_store(de[warp], i, x[i]);
        __syncwarp();
        int bi;
        float bv;
        search_warp(de[warp], lane, bi, bv);
        if (lane == 0) {
            max_index[s] = bi;
            max_value[s] = bv;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// The same search with the tensor-core proposer (sc_search_mma.cuh): one warp per PAIR of windows.
// ------------------------------------------------------------------------------------------------
constexpr int SMM_WARPS = 4;

__global__ void __launch_bounds__(SMM_WARPS * 32, 5)
search_mma_batch_kernel(const float2 *__restrict__ symbols, long symbol_stride, const uint4 *__restrict__ a_table,
                        int *__restrict__ max_index, float *__restrict__ max_value, long n_streams) {
    __shared__ __align__(16) SearchMmaSmem sm_all[SMM_WARPS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    SearchMmaSmem &sm = sm_all[warp];
    const long n_pairs = (n_streams + 1) / 2;
    for (long pr = (long) blockIdx.x * SMM_WARPS + warp; pr < n_pairs; pr += (long) gridDim.x * SMM_WARPS) {
        __syncwarp();
        // all 16 loads of the pair are in flight before anything is staged
        float2 v0[2][4], v1[2][4];
#pragma unroll
        for (int w = 0; w < 2; w++) {
            const long s = 2 * pr + w;
            const float2 *x = symbols + s * symbol_stride;
#pragma unroll
            for (int k = 0; k < 4; k++) {                      // symbol pairs (xx, xx + 1), xx = 2 lane + 64 k
                const int xx = 2 * lane + 64 * k;
                v0[w][k] = v1[w][k] = make_float2(0.0f, 0.0f);
                if (s < n_streams) {
                    v0[w][k] = x[xx];
                    if (xx + 1 < SEARCH_SYMS) v1[w][k] = x[xx + 1];
                }
            }
        }
        float part[2] = {0.0f, 0.0f};
#pragma unroll
        for (int w = 0; w < 2; w++)
#pragma unroll
            for (int k = 0; k < 4; k++)
                part[w] = __fadd_rn(part[w], search_mma_stage_pair(sm, w, 2 * lane + 64 * k, v0[w][k], v1[w][k]));
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            part[0] = __fadd_rn(part[0], __shfl_xor_sync(0xffffffffu, part[0], off));
            part[1] = __fadd_rn(part[1], __shfl_xor_sync(0xffffffffu, part[1], off));
        }
        if (lane == 0) {
            sm.d.s_abs[0] = part[0];
            sm.d.s_abs[1] = part[1];
        }
        __syncwarp();
        int bi[2];
        float bv[2];
        search_mma_pair<true>(sm.b, sm.d, a_table, lane, bi, bv);    // B fragments in registers: the kernel is LSU-bound
        if (lane < 2 && 2 * pr + lane < n_streams) {
            max_index[2 * pr + lane] = lane ? bi[1] : bi[0];
            max_value[2 * pr + lane] = lane ? bv[1] : bv[0];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// The same search with an FFT proposer (BASELINE.json north star: "preamble correlation ... as a hand-written
// batched FFT with warp-shuffle butterflies ... plus a warp-reduced argmax"): one warp per window, the 256-point
// transforms stay in registers (sc_fft256.cuh).
//   c[L] = sum_i pre[i] s[L+i]  <=>  C[k] = S[k] * Pt[k],  Pt[k] = sum_i pre[i] e^{+2 pi j k i / 256}
//   S = FFT(s) (255 symbols, zero padded), C = S .* Pt / 256, c = IFFT(C) = conj(FFT(conj C))
// The two transforms use different register layouts (natural order in, digit-permuted out), so conj(C) goes
// through shared memory once.  Lags 0..127 are the registers {0, 1, 4, 5} of every lane.  Error of the float32
// transforms is bounded by delta = 2^-12 * sum(|d| + |e|) per component (8 radix-4 stages, |Pt| <= 182); the
// candidates are verified exactly as in the tensor-core variant, so the result is the reference's, bit for bit.
// ------------------------------------------------------------------------------------------------
constexpr int SFFT_WARPS = 4;

struct SearchFftSmem {
    float de[2][SM_DE_FLOATS];          // d[256], e[256] (+8 words: 8 banks apart)
    float2 xpose[256];
    int cand[SM_MAX_CAND];
    int n_cand;
};

__global__ void __launch_bounds__(SFFT_WARPS * 32)
search_fft_batch_kernel(const float2 *__restrict__ symbols, long symbol_stride, const c32 *__restrict__ tw,
                        const float2 *__restrict__ ptab, int *__restrict__ max_index, float *__restrict__ max_value,
                        long n_streams) {
    __shared__ __align__(16) SearchFftSmem sm_all[SFFT_WARPS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    SearchFftSmem &sm = sm_all[warp];
    Fft256Twiddles T;
    T.load(tw, lane);
    c32 pt[8];                                              // Pt / 256 in the transform's output layout
#pragma unroll
    for (int r = 0; r < 8; r++) pt[r] = from2(__ldg(ptab + r * 32 + lane));
    for (long s = (long) blockIdx.x * SFFT_WARPS + warp; s < n_streams; s += (long) gridDim.x * SFFT_WARPS) {
        const float2 *x = symbols + s * symbol_stride;
        c32 v[8];
        float part = 0.0f;
        __syncwarp();
#pragma unroll
        for (int r = 0; r < 8; r++) {
            const int xx = lane + 32 * r;
            const float2 sv = xx < SEARCH_SYMS ? x[xx] : make_float2(0.0f, 0.0f);
            v[r] = from2(sv);
            const float d = __fsub_rn(sv.x, sv.y), e = __fadd_rn(sv.y, sv.x);      // qpsk.c:88-96 with pre = v(1+i)
            sm.de[0][xx] = d;
            sm.de[1][xx] = e;
            part = __fadd_rn(part, __fadd_rn(fabsf(d), fabsf(e)));
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) part = __fadd_rn(part, __shfl_xor_sync(0xffffffffu, part, off));
        // ---- propose
        fft256_regs<false>(v, lane, T);
#pragma unroll
        for (int r = 0; r < 8; r++) sm.xpose[fft256_out_index(lane, r)] = to2(cconj(cmul(v[r], pt[r])));
        __syncwarp();
#pragma unroll
        for (int r = 0; r < 8; r++) v[r] = from2(sm.xpose[lane + 32 * r]);
        fft256_regs<false>(v, lane, T);
        float val[4], vmax = 0.0f;
        int imax = 0;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int r = (q & 1) + 4 * (q >> 1);                                   // registers 0, 1, 4, 5: lags < 128
            val[q] = __fadd_rn(__fmul_rn(v[r].r, v[r].r), __fmul_rn(v[r].i, v[r].i));
            if (val[q] > vmax) {
                vmax = val[q];
                imax = fft256_out_index(lane, r);
            }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {                                    // warp-reduced argmax
            const float ov = __shfl_xor_sync(0xffffffffu, vmax, off);
            const int oi = __shfl_xor_sync(0xffffffffu, imax, off);
            if (ov > vmax || (ov == vmax && oi < imax)) {
                vmax = ov;
                imax = oi;
            }
        }
        const float thr = search_candidate_threshold(vmax, __fmul_rn(part, 0x1.004p-12f));
        if (lane == 0) sm.n_cand = 0;
        __syncwarp();
#pragma unroll
        for (int q = 0; q < 4; q++) {
            if (val[q] >= thr) {
                const int pos = atomicAdd(&sm.n_cand, 1);
                if (pos < SM_MAX_CAND) sm.cand[pos] = fft256_out_index(lane, (q & 1) + 4 * (q >> 1));
            }
        }
        __syncwarp();
        // ---- verify (both halves of the warp do the same work: only one window here)
        int bi;
        float bv;
        search_verify16(sm.de[0], sm.de[1], sm.cand, sm.n_cand, lane, bi, bv);
        if (sm.n_cand > SM_MAX_CAND || sm.n_cand == 0) search_warp_unpadded(sm.de[0], sm.de[1], lane, bi, bv);   // 0: NaN / Inf samples
        if (lane == 0) {
            max_index[s] = bi;
            max_value[s] = bv;
        }
    }
}

// Pt[k] / 256 in the output layout of fft256_regs (host side, double precision)
void search_fft_make_table(float *table /* [8][32][2] */) {
    for (int r = 0; r < 8; r++)
        for (int lane = 0; lane < 32; lane++) {
            const int k = fft256_out_index(lane, r);
            double re = 0.0, im = 0.0;
            for (int i = 0; i < PRE; i++) {
                const double ph = 2.0 * M_PI * (double) ((k * i) & 255) / 256.0, v = (double) preamble_value(i);
                // pre[i] = v (1 + j):  v (1 + j)(cos + j sin) = v (cos - sin) + j v (cos + sin)
                re += v * (cos(ph) - sin(ph));
                im += v * (cos(ph) + sin(ph));
            }
            table[(r * 32 + lane) * 2] = (float) (re / 256.0);
            table[(r * 32 + lane) * 2 + 1] = (float) (im / 256.0);
        }
}

cudaError_t launch_search_fft_batch(long n_streams, const float2 *symbols, long symbol_stride, const float2 *tw,
                                    const void *ptab, int *max_index, float *max_value, cudaStream_t st) {
    const int grid = (int) std::min<long>((n_streams + SFFT_WARPS - 1) / SFFT_WARPS, 148L * 4);
    search_fft_batch_kernel<<<grid, SFFT_WARPS * 32, 0, st>>>(symbols, symbol_stride, reinterpret_cast<const c32 *>(tw),
                                                              (const float2 *) ptab, max_index, max_value, n_streams);
    g_launch_count++;
    return cudaGetLastError();
}

// A fragments of the 9 Toeplitz tiles (sc_search_mma.cuh), host side
void search_mma_make_table(uint32_t *table /* [9][32][4] */) {
    auto val = [](int t, int r, int c) -> uint32_t {
        const int i = 16 * t + c - r;
        if (i < 0 || i >= PRE) return 0u;
        return preamble_value(i) < 0 ? 0xBF80u : 0x3F80u;                // bf16 -1.0 / +1.0
    };
    for (int t = 0; t < SM_TILES; t++)
        for (int lane = 0; lane < 32; lane++) {
            const int g = lane >> 2, tid = lane & 3;
            uint32_t *o = table + ((size_t) t * 32 + lane) * 4;
            o[0] = val(t, g, 2 * tid) | (val(t, g, 2 * tid + 1) << 16);
            o[1] = val(t, g + 8, 2 * tid) | (val(t, g + 8, 2 * tid + 1) << 16);
            o[2] = val(t, g, 2 * tid + 8) | (val(t, g, 2 * tid + 9) << 16);
            o[3] = val(t, g + 8, 2 * tid + 8) | (val(t, g + 8, 2 * tid + 9) << 16);
        }
}

cudaError_t launch_search_mma_batch(long n_streams, const float2 *symbols, long symbol_stride, const void *a_table,
                                    int *max_index, float *max_value, cudaStream_t st) {
    const long n_pairs = (n_streams + 1) / 2;
    const int grid = (int) std::min<long>((n_pairs + SMM_WARPS - 1) / SMM_WARPS, 148L * 5);     // one resident wave
    search_mma_batch_kernel<<<grid, SMM_WARPS * 32, 0, st>>>(symbols, symbol_stride, (const uint4 *) a_table, max_index,
                                                             max_value, n_streams);
    g_launch_count++;
    return cudaGetLastError();
}

cudaError_t launch_search_batch(long n_streams, const float2 *symbols, long symbol_stride, int *max_index,
                                float *max_value, cudaStream_t st) {
    const int grid = (int) std::min<long>((n_streams + SB_WARPS - 1) / SB_WARPS, 148L * 8);
    search_batch_kernel<<<grid, SB_WARPS * 32, 0, st>>>(symbols, symbol_stride, max_index, max_value, n_streams);
    g_launch_count++;
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// The decision half of qpsk_rx_frame() on explicit symbol windows dec[0..289], one thread per stream.
// ------------------------------------------------------------------------------------------------
struct RawLoader {
    const float2 *dec;
    int mi, ti;
    __device__ __forceinline__ c32 x(int r) const { return from2(dec[min(mi + r, WIN - 1)]); }
    __device__ __forceinline__ c32 y(int r) const { return from2(dec[min(max(ti + r, 0), WIN - 1)]); }
};

__global__ void __launch_bounds__(128)
track_window_kernel(const float2 *__restrict__ symbols, long symbol_stride, const int *__restrict__ max_index,
                    const float *__restrict__ max_value, int *__restrict__ rx_timing, uint32_t call_index,
                    unsigned long long keystream, sc_frame_result *__restrict__ results, float *__restrict__ eq_dbg,
                    long n_streams) {
    const long s = (long) blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_streams) return;
    RawLoader ld;
    ld.dec = symbols + s * symbol_stride;
    ld.mi = min(max(max_index[s], 0), PRE - 1);
    ld.ti = rx_timing[s];
    TrackOut o;
    track_core(ld, o);
    const int t_out = o.valid ? ld.mi + PRE : ld.ti;
    rx_timing[s] = t_out;
    store_result(results + s, o, keystream, max_value ? max_value[s] : 0.0f, ld.mi, t_out, call_index);
    if (eq_dbg != nullptr) {
#pragma unroll
        for (int i = 0; i < EQ; i++) {
            eq_dbg[s * 10 + 2 * i] = o.tk.C[i].r;
            eq_dbg[s * 10 + 2 * i + 1] = o.tk.C[i].i;
        }
    }
}

cudaError_t launch_track_window_batch(long n_streams, const float2 *symbols, long symbol_stride,
                                      const int *max_index, const float *max_value, int *rx_timing,
                                      uint32_t call_index, unsigned long long keystream,
                                      sc_frame_result *results, float *eq_dbg, cudaStream_t st) {
    const int grid = (int) ((n_streams + 127) / 128);
    track_window_kernel<<<grid, 128, 0, st>>>(symbols, symbol_stride, max_index, max_value, rx_timing, call_index,
                                              keystream, results, eq_dbg, n_streams);
    g_launch_count++;
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Self-test: rcp_rn_normal() (sc_exact.cuh) against __frcp_rn over a range of float bit patterns.
// ------------------------------------------------------------------------------------------------
__global__ void selftest_rcp_kernel(unsigned lo, unsigned hi, unsigned long long *mismatches) {
    unsigned long long bad = 0;
    for (unsigned long long b = (unsigned long long) lo + (unsigned long long) blockIdx.x * blockDim.x + threadIdx.x;
         b <= hi; b += (unsigned long long) gridDim.x * blockDim.x) {
        const float x = __uint_as_float((unsigned) b);
        if (__float_as_uint(rcp_rn_normal(x)) != __float_as_uint(__frcp_rn(x))) bad++;
    }
    if (bad) atomicAdd(mismatches, bad);
}

cudaError_t launch_selftest_rcp(unsigned lo, unsigned hi, unsigned long long *mismatches, cudaStream_t st) {
    selftest_rcp_kernel<<<148 * 16, 256, 0, st>>>(lo, hi, mismatches);
    g_launch_count++;
    return cudaGetLastError();
}

}  // namespace sc
