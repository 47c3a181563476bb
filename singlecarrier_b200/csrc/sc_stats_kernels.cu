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

// ------------------------------------------------------------------------------------------------
// Bit errors against the transmitted bits (sc_ber_stats_dev).  One thread per (stream, call).  The
// alignment rule is the one documented in include/singlecarrier_b200.h; the keystream words of calls
// 0..n_frames-1 are generated on the device by keystream_table_kernel (integer LFSR, src/scramble.c:57-69).
// ------------------------------------------------------------------------------------------------
__global__ void keystream_table_kernel(unsigned long long *__restrict__ ks, int n_frames) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    unsigned m = 0x4A80u;                                            // SEED, scramble.h:16
    for (int n = 0; n < n_frames; n++) {
        unsigned long long w = 0ull;
        for (int j = 0; j < SC_BITS_PER_CALL; j++) {
            const unsigned out = ((m >> 1) ^ m) & 1u;
            m = (m >> 1) | (out << 14);
            w |= (unsigned long long) out << j;
        }
        ks[n] = w;
    }
}

constexpr int BER_GROUP_DELAY = 48;                                  // TX RRC (24) + RX RRC (24) samples
constexpr int BER_ALIGN_TOL = 10;

__global__ void __launch_bounds__(256)
ber_stats_kernel(const sc_frame_result *__restrict__ results, long n_streams, long result_stride, int n_frames,
                 const uint8_t *__restrict__ tx_bits, int n_packets, const int *__restrict__ lead_in, int gap,
                 const int *__restrict__ group, int n_groups, const unsigned long long *__restrict__ ks,
                 unsigned long long *__restrict__ counters) {
    const long total = n_streams * (long) n_frames;
    const int period = FRAME + gap;
    for (long k = (long) blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (long) gridDim.x * blockDim.x) {
        const long s = k / n_frames;
        const int n = (int) (k - s * n_frames);
        if (n < 2) continue;
        const sc_frame_result *rec = results + s * result_stride;
        const uint4 *p = reinterpret_cast<const uint4 *>(rec + n);
        const uint4 a = __ldg(p), b = __ldg(p + 1);
        const bool valid = ((b.y >> 16) & 0xffu) != 0;
        int g = group ? group[s] : 0;
        if (g < 0 || g >= n_groups) continue;
        unsigned long long *c = counters + (long) g * SC_N_BER_COUNTERS;
        atomicAdd(&c[0], 1ull);
        if (!valid) continue;
        atomicAdd(&c[1], 1ull);
        const int max_index = (int) (short) (b.x & 0xffffu);
        // rx_timing in force when frame n-2 was decimated = its value after call n-2
        const int t_prev = (int) rec[n - 2].rx_timing;
        const long pos = (long) (n - 2) * FRAME + 5 * max_index + t_prev - BER_GROUP_DELAY - (lead_in ? lead_in[s] : 0);
        long pkt = pos + period / 2;
        pkt = pkt >= 0 ? pkt / period : -((-pkt + period - 1) / period);          // floor division
        if (pkt < 0 || pkt >= n_packets) continue;
        const long off = pos - pkt * period;
        if (off > BER_ALIGN_TOL || off < -BER_ALIGN_TOL) continue;
        const uint8_t *tb = tx_bits + ((s * n_packets + pkt) * 8) * (long) SC_BITS_PER_CALL;   // first data frame
        unsigned long long txw = 0ull;
        for (int j = 0; j < SC_BITS_PER_CALL; j++) txw |= (unsigned long long) (tb[j] & 1u) << j;
        const unsigned long long bits = ((unsigned long long) a.y << 32) | a.x;
        const unsigned long long diff = (bits ^ ks[n] ^ txw) & ((1ull << SC_BITS_PER_CALL) - 1ull);
        atomicAdd(&c[2], 1ull);
        atomicAdd(&c[3], (unsigned long long) SC_BITS_PER_CALL);
        atomicAdd(&c[4], (unsigned long long) __popcll(diff));
    }
}

cudaError_t launch_ber_stats(const sc_frame_result *results, long n_streams, long result_stride, int n_frames,
                             const uint8_t *tx_bits, int n_packets, const int *lead_in, int gap, const int *group,
                             int n_groups, unsigned long long *counters, cudaStream_t st) {
    unsigned long long *ks = nullptr;
    cudaError_t e = cudaMallocAsync((void **) &ks, (size_t) n_frames * sizeof(unsigned long long), st);
    if (e != cudaSuccess) return e;
    keystream_table_kernel<<<1, 32, 0, st>>>(ks, n_frames);
    g_launch_count++;
    const long total = n_streams * (long) n_frames;
    int grid = (int) std::min<long>((total + 255) / 256, 148 * 8);
    if (grid < 1) grid = 1;
    ber_stats_kernel<<<grid, 256, 0, st>>>(results, n_streams, result_stride, n_frames, tx_bits, n_packets, lead_in, gap,
                                           group, n_groups, ks, counters);
    g_launch_count++;
    e = cudaGetLastError();
    cudaFreeAsync(ks, st);
    return e;
}

}  // namespace sc
