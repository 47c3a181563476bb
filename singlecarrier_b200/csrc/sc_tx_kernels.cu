// sc_tx_kernels.cu -- TX synthesis (and the loop-back channel) as sm_100a kernels.
//
// Replaces qpsk_tx_frame()/preamble_modulate()/qpsk_modulate() and the transmit loop of main()
// (src/qpsk.c:278-342, 380-413) for a bank of streams.  One thread per output sample:
//   zero-stuffed symbols -> 49-tap RRC (src/fir.c) -> x tx phasor (+1100 Hz) -> real part ->
//   int16 by C truncation, preamble at half amplitude.
// Exactness: only every fifth filter input is non-zero and it is +-1(+-1i), so every non-zero
// product mem*c is exactly +-c; the zero products the reference also adds are +-0 and cannot
// change a running sum that started at +0.  The <= 10 non-zero terms are added in the reference's
// (ascending tap) order, so the result is bit-identical to fir() on the zero-stuffed signal.
// tx_filter persists across frames and packets and the dead air between packets bypasses the
// filter and the NCO (qpsk.c:410-412), so "filter time" u runs over packet samples only.
#include "sc_common.cuh"
#include "sc_tables.cuh"
#include "sc_kernels.h"

#include <algorithm>
#include <mutex>

namespace sc {

__constant__ float c_taps[2][NTAPS];
__constant__ uint32_t c_tx_pre_neg[4] = {pre_neg_word(0), pre_neg_word(1), pre_neg_word(2), pre_neg_word(3)};

// the tap table is uploaded once per device; banks may be created from several threads at once
static std::mutex g_taps_mu;
static bool g_taps_uploaded[64] = {};

static cudaError_t upload_taps() {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lock(g_taps_mu);
    if (dev >= 0 && dev < 64 && g_taps_uploaded[dev]) return cudaSuccess;
    float h[2][NTAPS];
    for (int k = 0; k < NTAPS; k++) {
        h[0][k] = tap35(k);
        h[1][k] = tap50(k);
    }
    e = cudaMemcpyToSymbol(c_taps, h, sizeof h);
    if (e == cudaSuccess && dev >= 0 && dev < 64) g_taps_uploaded[dev] = true;
    return e;
}

__device__ __forceinline__ unsigned long long mix64(unsigned long long z) {   // splitmix64 finaliser
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

// ---- data bits: pack the caller's bits, or draw them, one 62-bit word per (stream, packet, frame)
__global__ void tx_bits_kernel(const uint8_t *__restrict__ bits, uint8_t *__restrict__ bits_out,
                               unsigned long long *__restrict__ words, unsigned long long seed, long n_words,
                               bool scramble, PacketKey key) {
    const long w = (long) blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_words) return;
    unsigned long long v = 0;
    if (bits != nullptr) {
        const uint8_t *b = bits + w * SC_BITS_PER_CALL;
        for (int j = 0; j < SC_BITS_PER_CALL; j++) v |= (unsigned long long) (b[j] == 1) << j;   // qpsk.c:252-253
    } else {
        v = mix64(seed ^ mix64((unsigned long long) w)) & ((1ull << SC_BITS_PER_CALL) - 1ull);
    }
    // packet mode (extension): scramble(&sdata, tx) with the register seeded after each preamble, qpsk.c:386,397;
    // w % 8 is the data frame inside the packet.  bits_out stays the payload.
    words[w] = scramble ? v ^ key.k[w & 7] : v;
    if (bits_out != nullptr) {
        uint8_t *b = bits_out + w * SC_BITS_PER_CALL;
        for (int j = 0; j < SC_BITS_PER_CALL; j++) b[j] = (uint8_t) ((v >> j) & 1ull);
    }
}

struct TxGeom {
    int n_packets, gap, period;        // period = 1880 + gap
    int wide;
};

// symbol Q of the stream's continuous symbol sequence (376 per packet); Q < 0 = cold filter memory
__device__ __forceinline__ c32 tx_symbol(const unsigned long long *__restrict__ words, int Q) {
    if (Q < 0) return mk(0.0f, 0.0f);
    const int P = Q / SC_PACKET_SYMBOLS, q = Q - P * SC_PACKET_SYMBOLS;
    if (q < PRE) {                                          // preambletable, qpsk.c:361-365
        const float v = ((c_tx_pre_neg[q >> 5] >> (q & 31)) & 1u) ? -1.0f : 1.0f;
        return mk(v, v);
    }
    const int f = (q - PRE) / NDATA, i = (q - PRE) - f * NDATA;
    const unsigned long long w = words[P * 8 + f];
    const float vi = ((w >> (2 * i + 1)) & 1ull) ? -1.0f : 1.0f;    // qpsk_mod(), qpsk.c:251-256
    const float vq = ((w >> (2 * i)) & 1ull) ? -1.0f : 1.0f;
    return mk(vi, vq);
}

// analytic TX sample (before taking the real part), scaled to int16 units, at filter time u
__device__ __forceinline__ c32 tx_analytic(const unsigned long long *__restrict__ words,
                                           const float2 *__restrict__ tx_table, int wide, int u) {
    // non-zero filter inputs sit at v = u-48+k with v % 5 == 0
    const int v0 = u - (NTAPS - 1);
    int k = ((-v0) % CYC + CYC) % CYC;                     // smallest k >= 0 with (v0 + k) % 5 == 0
    float yr = 0.0f, yi = 0.0f;
    for (; k < NTAPS; k += CYC) {
        const int v = v0 + k;
        if (v < 0) continue;
        const c32 sy = tx_symbol(words, v / CYC);
        const float c = c_taps[wide][k];
        yr = __fadd_rn(yr, __fmul_rn(sy.r, c));            // src/fir.c:38-40
        yi = __fadd_rn(yi, __fmul_rn(sy.i, c));
    }
    const c32 sig = mk(__fmul_rn(yr, FIR_GAIN), __fmul_rn(yi, FIR_GAIN));      // src/fir.c:42
    const c32 ph = from2(__ldg(tx_table + u));
    const c32 z = cmul(sig, ph);                                               // qpsk.c:303
    const int r = u % FRAME;
    const float scale = r < PRE * CYC ? 8192.0f : 16384.0f;                    // qpsk.c:313-319
    return mk(__fmul_rn(z.r, scale), __fmul_rn(z.i, scale));
}

// output position p of a stream -> filter time u, or -1 for lead-in / dead air / tail
__device__ __forceinline__ int tx_filter_time(const TxGeom &g, long p, int lead) {
    const long pp = p - lead;
    if (pp < 0) return -1;
    const long P = pp / g.period;
    const int r = (int) (pp - P * g.period);
    if (P >= g.n_packets || r >= FRAME) return -1;
    return (int) P * FRAME + r;
}

template <bool CHANNEL>
__global__ void __launch_bounds__(256)
tx_kernel(TxGeom g, const unsigned long long *__restrict__ words, const float2 *__restrict__ tx_table,
          const int *__restrict__ lead_in, sc_channel ch, unsigned long long seed, long s_base,
          int16_t *__restrict__ out, long stream_stride, long samples_per_stream) {
    const long s = blockIdx.y;
    const unsigned long long *w = words + s * (long) g.n_packets * 8;
    const int lead = lead_in ? lead_in[s] : 0;
    int16_t *o = out + s * stream_stride;

    for (long p = (long) blockIdx.x * blockDim.x + threadIdx.x; p < samples_per_stream;
         p += (long) gridDim.x * blockDim.x) {
        const int u = tx_filter_time(g, p, lead);
        if (!CHANNEL) {
            int16_t v = 0;
            if (u >= 0) v = (int16_t) __float2int_rz(tx_analytic(w, tx_table, g.wide, u).r);   // (int16_t) cast
            o[p] = v;
        } else {
            c32 z = mk(0.0f, 0.0f);
            if (u >= 0) z = tx_analytic(w, tx_table, g.wide, u);
            const float a = ch.echo_amp ? ch.echo_amp[s] : 0.0f;
            if (a != 0.0f) {                                 // h = [1, a*exp(j*theta)] at delay d
                const int d = ch.echo_delay ? ch.echo_delay[s] : 1;
                const int u2 = tx_filter_time(g, p - d, lead);
                if (u2 >= 0) {
                    const c32 z2 = tx_analytic(w, tx_table, g.wide, u2);
                    float sn, cs;
                    sincosf(ch.echo_theta ? ch.echo_theta[s] : 0.0f, &sn, &cs);
                    z.r += a * (z2.r * cs - z2.i * sn);
                    z.i += a * (z2.r * sn + z2.i * cs);
                }
            }
            // rotation by exp(j(2*pi*(df + drift*t/2)*t + phi)), phase kept in cycles in double
            const double t = (double) p / 8000.0;
            const double df = ch.df_hz ? (double) ch.df_hz[s] : 0.0;
            const double dr = ch.drift_hz_s ? (double) ch.drift_hz_s[s] : 0.0;
            double cyc = (df + 0.5 * dr * t) * t + (ch.phi_rad ? (double) ch.phi_rad[s] * 0.15915494309189535 : 0.0);
            cyc -= floor(cyc);
            float sn, cs;
            sincospif((float) (2.0 * cyc), &sn, &cs);
            float re = z.r * cs - z.i * sn;
            const float sigma = ch.sigma_lsb ? ch.sigma_lsb[s] : 0.0f;
            if (sigma != 0.0f) {                             // Box-Muller on a counter-based draw
                const unsigned long long h = mix64(seed ^ mix64(((unsigned long long) (s_base + s) << 34) ^ (unsigned long long) p ^ 0xA5A5A5A5ull));
                const float u1 = ((float) (unsigned) (h >> 40) + 0.5f) * (1.0f / 16777216.0f);
                const float u2f = (float) (unsigned) ((h >> 8) & 0xffffffu) * (1.0f / 16777216.0f);
                re += sigma * sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2f);
            }
            re = fminf(fmaxf(re, -32767.0f), 32767.0f);
            o[p] = (int16_t) __float2int_rn(re);
        }
    }
}

cudaError_t launch_tx(const TxArgs &a, cudaStream_t st) {
    cudaError_t e = upload_taps();
    if (e != cudaSuccess) return e;
    const long n_words = a.n_streams * (long) a.n_packets * 8;
    unsigned long long *words = nullptr;
    e = cudaMallocAsync((void **) &words, (size_t) std::max<long>(n_words, 1) * sizeof(unsigned long long), st);
    if (e != cudaSuccess) return e;
    if (n_words > 0) {
        tx_bits_kernel<<<(unsigned) ((n_words + 255) / 256), 256, 0, st>>>(a.bits, a.bits_out, words, a.seed, n_words,
                                                                            a.scramble, a.key);
        g_launch_count++;
    }
    TxGeom g;
    g.n_packets = a.n_packets;
    g.gap = a.gap_samples;
    g.period = FRAME + a.gap_samples;
    g.wide = a.wide ? 1 : 0;
    // grid.y is limited to 65535: loop over slices of streams
    const long per_stream_blocks = std::min<long>((a.samples_per_stream + 255) / 256, 1024);
    for (long s0 = 0; s0 < a.n_streams; s0 += 65535) {
        const long ns = std::min<long>(65535, a.n_streams - s0);
        dim3 grid((unsigned) std::max<long>(per_stream_blocks, 1), (unsigned) ns);
        sc_channel ch = a.ch;
        if (a.use_channel) {
            // per-stream parameter arrays are indexed by the slice-local stream: advance them
            if (ch.df_hz) ch.df_hz += s0;
            if (ch.phi_rad) ch.phi_rad += s0;
            if (ch.drift_hz_s) ch.drift_hz_s += s0;
            if (ch.sigma_lsb) ch.sigma_lsb += s0;
            if (ch.echo_amp) ch.echo_amp += s0;
            if (ch.echo_theta) ch.echo_theta += s0;
            if (ch.echo_delay) ch.echo_delay += s0;
            tx_kernel<true><<<grid, 256, 0, st>>>(g, words + s0 * (long) a.n_packets * 8, a.tx_table,
                                                  a.lead_in ? a.lead_in + s0 : nullptr, ch, a.seed, s0,
                                                  a.out + s0 * a.stream_stride, a.stream_stride, a.samples_per_stream);
        } else {
            tx_kernel<false><<<grid, 256, 0, st>>>(g, words + s0 * (long) a.n_packets * 8, a.tx_table,
                                                   a.lead_in ? a.lead_in + s0 : nullptr, ch, a.seed, s0,
                                                   a.out + s0 * a.stream_stride, a.stream_stride, a.samples_per_stream);
        }
        g_launch_count++;
    }
    e = cudaGetLastError();
    cudaFreeAsync(words, st);
    return e;
}

}  // namespace sc
