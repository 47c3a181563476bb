// sc_stage_kernels.cu -- batched forms of the reference's L1 primitives as stand-alone kernels
// (the stage entry points of include/singlecarrier_b200.h and the back end of the drop-in symbols).
//   fir_batch_kernel     fir()                            src/fir.c:22-44
//   search_batch_kernel  correlate() + argmax             src/qpsk.c:88-96, 172-183
//   track_window_kernel  kalman_reset .. data_eq loop     src/qpsk.c:186-238
#include "sc_common.cuh"
#include "sc_tables.cuh"
#include "sc_tracker.cuh"
#include "sc_search.cuh"
#include "sc_search_mma.cuh"
#include "sc_track_core.cuh"
#include "sc_kernels.h"

namespace sc {

// ------------------------------------------------------------------------------------------------
// fir(): 49-tap real-coefficient FIR on complex samples, in place, caller-owned delay line holding
// the RAW past inputs, output scaled by GAIN.  One WARP per stream walks the samples in tiles of
// 160 (32 lanes x 5 consecutive outputs; lane stride 5 slots => conflict-free LDS.64) with the last
// 49 raw inputs carried in shared memory between tiles, so each sample is read and written once and
// only __syncwarp() is needed; 32 warps per SM hide the global-memory latency of each other's tiles.
// ext[0..48] = memory[0..48] (oldest first), ext[49 + j] = sample[j];
// out[j] = GAIN * sum_k ext[j + 1 + k] * coeff[k], k ascending (src/fir.c:36-42).
//
// FAST = true is the explicitly named tolerance mode (SC_FIR_FAST): the multiply-add is contracted
// into one FFMA2, which halves the FP32 work, at the price of last-bit differences from the
// reference (never used on the parity path).
// ------------------------------------------------------------------------------------------------
constexpr int FIR_WARPS = 8;
constexpr int FIR_R = 5;
constexpr int FIR_TILE = 32 * FIR_R;                   // 160

template <bool WIDE, bool FAST>
__global__ void __launch_bounds__(FIR_WARPS * 32, 4)
fir_batch_kernel(float2 *__restrict__ memory, float2 *__restrict__ sample, long sample_stride, int length,
                 long n_streams) {
    __shared__ __align__(16) float2 ext_all[FIR_WARPS][NTAPS + FIR_TILE + 7];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float2 *ext = ext_all[warp];
    for (long s = (long) blockIdx.x * FIR_WARPS + warp; s < n_streams; s += (long) gridDim.x * FIR_WARPS) {
        float2 *mem = memory + s * NTAPS;
        float2 *x = sample + s * sample_stride;
        __syncwarp();
        for (int i = lane; i < NTAPS; i += 32) ext[i] = mem[i];
        for (int t0 = 0; t0 < length; t0 += FIR_TILE) {
            const int n = min(FIR_TILE, length - t0);
            for (int j = lane; j < n; j += 32) ext[NTAPS + j] = x[t0 + j];
            __syncwarp();
            u64 acc[FIR_R];
#pragma unroll
            for (int r = 0; r < FIR_R; r++) acc[r] = 0ull;
            if (FIR_R * lane < n) {
                const u64 *ep = reinterpret_cast<const u64 *>(ext) + FIR_R * lane + 1;
#pragma unroll
                for (int j = 0; j < NTAPS + FIR_R - 1; j++) {
                    const u64 v = ep[j];
#pragma unroll
                    for (int r = 0; r < FIR_R; r++) {
                        const int k = j - r;
                        if (k >= 0 && k < NTAPS) {
                            if (FAST) acc[r] = pk_fma_bcast(v, tap<WIDE>(k), acc[r]);
                            else acc[r] = pk_add(acc[r], pk_mul_bcast_pz(v, tap<WIDE>(k)));
                        }
                    }
                }
            }
            // the last 49 raw inputs become the head of the next tile
            float2 c0 = make_float2(0.f, 0.f), c1 = c0;
            c0 = ext[n + lane];
            if (lane + 32 < NTAPS) c1 = ext[n + lane + 32];
            __syncwarp();
#pragma unroll
            for (int r = 0; r < FIR_R; r++) {
                const int j = FIR_R * lane + r;
                if (j < n) {
                    float yr, yi;
                    unpk(acc[r], yr, yi);
                    x[t0 + j] = make_float2(__fmul_rn(yr, FIR_GAIN), __fmul_rn(yi, FIR_GAIN));
                }
            }
            ext[lane] = c0;
            if (lane + 32 < NTAPS) ext[lane + 32] = c1;
            __syncwarp();
        }
        for (int i = lane; i < NTAPS; i += 32) mem[i] = ext[i];
    }
}

cudaError_t launch_fir_batch(bool wide, long n_streams, float2 *memory, float2 *sample, long sample_stride,
                             int length, cudaStream_t st, bool fast) {
    const int grid = (int) std::min<long>((n_streams + FIR_WARPS - 1) / FIR_WARPS, 148L * 4);
    const int thr = FIR_WARPS * 32;
    if (fast) {
        if (wide) fir_batch_kernel<true, true><<<grid, thr, 0, st>>>(memory, sample, sample_stride, length, n_streams);
        else fir_batch_kernel<false, true><<<grid, thr, 0, st>>>(memory, sample, sample_stride, length, n_streams);
    } else if (wide) fir_batch_kernel<true, false><<<grid, thr, 0, st>>>(memory, sample, sample_stride, length, n_streams);
    else fir_batch_kernel<false, false><<<grid, thr, 0, st>>>(memory, sample, sample_stride, length, n_streams);
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

__global__ void __launch_bounds__(SMM_WARPS * 32, 6)
search_mma_batch_kernel(const float2 *__restrict__ symbols, long symbol_stride, const uint4 *__restrict__ a_table,
                        int *__restrict__ max_index, float *__restrict__ max_value, long n_streams) {
    __shared__ __align__(16) SearchMmaSmem sm_all[SMM_WARPS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    SearchMmaSmem &sm = sm_all[warp];
    const long n_pairs = (n_streams + 1) / 2;
    for (long pr = (long) blockIdx.x * SMM_WARPS + warp; pr < n_pairs; pr += (long) gridDim.x * SMM_WARPS) {
        float s_abs[2];
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
        s_abs[0] = part[0];
        s_abs[1] = part[1];
        __syncwarp();
        int bi[2];
        float bv[2];
        search_mma_pair(sm, a_table, lane, s_abs, bi, bv);
        if (lane < 2 && 2 * pr + lane < n_streams) {
            max_index[2 * pr + lane] = lane ? bi[1] : bi[0];
            max_value[2 * pr + lane] = lane ? bv[1] : bv[0];
        }
    }
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
    const int grid = (int) std::min<long>((n_pairs + SMM_WARPS - 1) / SMM_WARPS, 148L * 6);     // one resident wave
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
