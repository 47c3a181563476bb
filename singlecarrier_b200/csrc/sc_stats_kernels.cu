// sc_stats_kernels.cu -- lock / bit statistics over a batch of results (the
// reference only has the DEBUG2 printf and the preamble_frames_detected counter, qpsk.c:70,193,198).
// Counters are plain sums so that ranks can combine them with one ncclAllReduce(sum).
#include "sc_common.cuh"
#include "sc_kernels.h"

namespace sc {

__global__ void __launch_bounds__(256)
lock_stats_kernel(const sc_frame_result *__restrict__ results, long n_streams, long result_stride, int n_frames,
                  unsigned long long *__restrict__ counters) {
    __shared__ unsigned long long sh[SC_N_COUNTERS];
    if (threadIdx.x < SC_N_COUNTERS) sh[threadIdx.x] = 0ull;
    __syncthreads();
    unsigned long long c[SC_N_COUNTERS];
#pragma unroll
    for (int i = 0; i < SC_N_COUNTERS; i++) c[i] = 0ull;
    const long total = n_streams * (long) n_frames;
    for (long k = (long) blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (long) gridDim.x * blockDim.x) {
        const long s = k / n_frames, j = k - s * n_frames;
        const uint4 *p = reinterpret_cast<const uint4 *>(results + s * result_stride + j);
        const uint4 a = __ldg(p), b = __ldg(p + 1);
        const unsigned long long bits = ((unsigned long long) a.y << 32) | a.x;
        const int matches = (int) (short) (b.x >> 16);
        const int max_index = (int) (short) (b.x & 0xffffu);
        const int rx_timing = (int) (short) (b.y & 0xffffu);
        const bool valid = ((b.y >> 16) & 0xffu) != 0;
        c[0] += 1;
        c[2] += (unsigned long long) matches;
        c[7] += (unsigned long long) rx_timing;
        if (valid) {
            c[1] += 1;
            c[3] += (unsigned long long) matches;
            c[4] += (unsigned long long) max_index;
            c[5] += (unsigned long long) __popcll(bits);
            c[6] += (bits & 0xffffffffull) + (bits >> 32);
        }
        int bin = matches >> 4;
        bin = bin > 7 ? 7 : bin;
#pragma unroll
        for (int h = 0; h < 8; h++) c[8 + h] += (bin == h) ? 1ull : 0ull;
    }
#pragma unroll
    for (int i = 0; i < SC_N_COUNTERS; i++) {
        unsigned long long v = c[i];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        if ((threadIdx.x & 31) == 0 && v) atomicAdd(&sh[i], v);
    }
    __syncthreads();
    if (threadIdx.x < SC_N_COUNTERS && sh[threadIdx.x]) atomicAdd(&counters[threadIdx.x], sh[threadIdx.x]);
}

cudaError_t launch_lock_stats(const sc_frame_result *results, long n_streams, long result_stride, int n_frames,
                              unsigned long long *counters, cudaStream_t st) {
    const long total = n_streams * (long) n_frames;
    int grid = (int) std::min<long>((total + 255) / 256, 148 * 8);
    if (grid < 1) grid = 1;
    lock_stats_kernel<<<grid, 256, 0, st>>>(results, n_streams, result_stride, n_frames, counters);
    g_launch_count++;
    return cudaGetLastError();
}

}  // namespace sc
