// sc_legacy_dev.cu -- CUDA side of the drop-in single-stream symbols (scl_* bridge of
// sc_legacy_internal.h).  Each call marshals its arguments to the device, runs the same device
// code the batched kernels use on ONE stream, and copies the result back.  This is a
// compatibility path (one launch per call); the batched API is the performance path.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <mutex>
#include <vector>

#include "sc_common.cuh"
#include "sc_tables.cuh"
#include "sc_tracker.cuh"
#include "sc_kernels.h"
#include "sc_legacy_internal.h"

namespace sc {
int api_fail(int code, const char *msg);
uint16_t modem_lfsr_rx(const sc_modem *m);
void modem_set_lfsr_rx(sc_modem *m, uint16_t v);
void modem_set_state_dbg(sc_modem *m, float *p);

// the reference's process-global equalizer / scrambler state (src/kalman.c:19-35, src/scramble.c:41-42)
struct LegacyState {
    float2 C[5], G[5], U[10];
    float D[5];
    float KY;
    float2 x[5];        // in[index .. index+4] of the current call; also scratch for the misc ops
    float ret;
    int dibit;
    unsigned lfsr_tx, lfsr_rx;
};

__device__ __forceinline__ unsigned lfsr_scramble2(unsigned v, unsigned &m) {     // src/scramble.c:57-69
#pragma unroll
    for (int i = 0; i < 2; i++) {
        const unsigned out = ((m >> 1) ^ m) & 1u;
        v ^= out << i;
        m = (m >> 1) | (out << 14);
    }
    return v & 3u;
}

// op: 0 kalman_reset, 1 kalman_calculate, 2 train_eq, 3 data_eq, 4 kalman_init,
//     5 scramble_init(arg), 6 scramble(arg = value | sr << 8), 7 cnormf, 8 qpsk_mod, 9 qpsk_demod
__global__ void legacy_op_kernel(LegacyState *st, int op, float ref, int arg) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    Tracker tk;
#pragma unroll
    for (int i = 0; i < EQ; i++) {
        tk.C[i] = from2(st->C[i]);
        tk.G[i] = from2(st->G[i]);
        tk.D[i] = st->D[i];
    }
#pragma unroll
    for (int i = 0; i < 10; i++) tk.U[i] = from2(st->U[i]);
    tk.KY = st->KY;
    c32 x[EQ];
#pragma unroll
    for (int i = 0; i < EQ; i++) x[i] = from2(st->x[i]);

    if (op == 0 || op == 4) {
        const float ky = tk.KY;              // kalman_reset() leaves kalman_y alone (src/kalman.c:42-55)
        tk.reset();
        tk.KY = ky;
    } else if (op == 1) {
        tk.kalman(x);
    } else if (op == 2) {
        st->ret = tk.train(x, ref);
    } else if (op == 3) {
        int bI, bQ;
        st->ret = tk.data(x, bI, bQ);
        unsigned m = st->lfsr_rx;
        st->dibit = (int) lfsr_scramble2((unsigned) ((bI << 1) | bQ), m);     // equalizer.c:83-87
        st->lfsr_rx = m;
    } else if (op == 5) {                    // scramble_init(), src/scramble.c:46-55
        if (arg == 0 || arg == 2) st->lfsr_tx = 0x4A80u;
        if (arg == 1 || arg == 2) st->lfsr_rx = 0x4A80u;
    } else if (op == 6) {
        unsigned m = (arg >> 8) == 0 ? st->lfsr_tx : st->lfsr_rx;
        // scramble_internal() only rewrites bits 0 and 1 of *input
        st->dibit = (int) (((unsigned) arg & 0xfcu) | lfsr_scramble2((unsigned) arg & 3u, m));
        if ((arg >> 8) == 0) st->lfsr_tx = m; else st->lfsr_rx = m;
    } else if (op == 7) {                    // cnormf(), src/qpsk.c:75-80
        st->ret = __fadd_rn(__fmul_rn(x[0].r, x[0].r), __fmul_rn(x[0].i, x[0].i));
    } else if (op == 8) {                    // qpsk_mod(), src/qpsk.c:251-256: x[0] = (bitI, bitQ) as 0/1
        st->x[0] = make_float2(x[0].r == 1.0f ? -1.0f : 1.0f, x[0].i == 1.0f ? -1.0f : 1.0f);
    } else if (op == 9) {                    // qpsk_demod(), src/qpsk.c:268-271: bits[0] = Q, bits[1] = I
        st->x[0] = make_float2(x[0].i < 0.0f ? 1.0f : 0.0f, x[0].r < 0.0f ? 1.0f : 0.0f);
    }
    if (op <= 4) {
#pragma unroll
        for (int i = 0; i < EQ; i++) {
            st->C[i] = to2(tk.C[i]);
            st->G[i] = to2(tk.G[i]);
            st->D[i] = tk.D[i];
        }
#pragma unroll
        for (int i = 0; i < 10; i++) st->U[i] = to2(tk.U[i]);
        st->KY = tk.KY;
    }
}

// qpsk_tx_frame() pieces (src/qpsk.c:285-291, 301-319): zero-stuffing and the mix/convert stage
__global__ void legacy_tx_stuff_kernel(const float2 *__restrict__ sym, int length, float2 *__restrict__ sig) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < length * CYC) sig[i] = (i % CYC == 0) ? sym[i / CYC] : make_float2(0.0f, 0.0f);
}
__global__ void legacy_tx_mix_kernel(const float2 *__restrict__ sig, const float2 *__restrict__ ph, int n, float scale,
                                     int16_t *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const c32 z = cmul(from2(sig[i]), from2(ph[i]));                       // signal[i] *= fbb_tx_phase
        out[i] = (int16_t) __float2int_rz(__fmul_rn(z.r, scale));              // (int16_t)(crealf() * scale)
    }
}
}  // namespace sc

using namespace sc;

namespace {
struct Legacy {
    bool ready = false;
    int device = 0;
    LegacyState *d_state = nullptr;
    sc_modem *bank = nullptr;
    float *d_track_state = nullptr;          // the bank's tracker state after its last call (48 floats)
    float2 *d_txfilter = nullptr, *d_txphase = nullptr;
    float2 tx_rect;
    std::mutex mu;
} g;

// device scratch that is released on every return path
template <typename T>
struct DevBuf {
    T *p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t n) { return cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)); }
    operator T *() const { return p; }
};

#define LCU(call)                                                                                  \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            char buf[256];                                                                         \
            snprintf(buf, sizeof buf, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
            return api_fail(SC_ECUDA, buf);                                                        \
        }                                                                                          \
    } while (0)

int ensure() {
    if (g.ready) {
        LCU(cudaSetDevice(g.device));
        return SC_OK;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return api_fail(SC_ECUDA, "no CUDA device (this library has no CPU path)");
    }
    const char *e = getenv("SC_DEVICE");
    g.device = e ? atoi(e) : 0;
    LCU(cudaSetDevice(g.device));
    LCU(cudaMalloc(&g.d_state, sizeof(LegacyState)));
    LegacyState h;
    memset(&h, 0, sizeof h);                         // C statics start as zero; kalman_init() sets d[] = 1
    h.lfsr_tx = h.lfsr_rx = 0;
    LCU(cudaMemcpy(g.d_state, &h, sizeof h, cudaMemcpyHostToDevice));
    LCU(cudaMalloc(&g.d_txfilter, NTAPS * sizeof(float2)));
    LCU(cudaMemset(g.d_txfilter, 0, NTAPS * sizeof(float2)));
    LCU(cudaMalloc(&g.d_txphase, sizeof(float2)));
    const float2 one = make_float2(1.0f, 0.0f);      // fbb_tx_phase = cmplx(0.0f), qpsk.c:375
    LCU(cudaMemcpy(g.d_txphase, &one, sizeof one, cudaMemcpyHostToDevice));
    const double tau = 2.0f * M_PI;
    const float xr = (float) (tau * 1100.0f / 8000.0f);                       // cmplx(TAU * CENTER / FS), qpsk.c:376
    g.tx_rect = make_float2(cosf(xr), sinf(xr));
    g.ready = true;
    return SC_OK;
}
}  // namespace

extern "C" int scl_fir(float *memory, int wide, float *sample, int length) {
    std::lock_guard<std::mutex> lock(g.mu);
    int rc = ensure();
    if (rc != SC_OK) return rc;
    DevBuf<float2> d_mem, d_x;
    LCU(d_mem.alloc(NTAPS));
    LCU(d_x.alloc((size_t) length));
    LCU(cudaMemcpy(d_mem, memory, NTAPS * sizeof(float2), cudaMemcpyHostToDevice));
    LCU(cudaMemcpy(d_x, sample, (size_t) length * sizeof(float2), cudaMemcpyHostToDevice));
    LCU(launch_fir_batch(wide != 0, 1, d_mem, d_x, length, length, 0));
    LCU(cudaMemcpy(memory, d_mem, NTAPS * sizeof(float2), cudaMemcpyDeviceToHost));
    LCU(cudaMemcpy(sample, d_x, (size_t) length * sizeof(float2), cudaMemcpyDeviceToHost));
    return SC_OK;
}

extern "C" int scl_eq_op(int op, const float *x5, float ref, float *eq_coeff, float *gain, float *ky, float *ret,
                         int *dibit) {
    std::lock_guard<std::mutex> lock(g.mu);
    int rc = ensure();
    if (rc != SC_OK) return rc;
    // the reference's globals are caller-visible (and caller-writable): push them, run, pull them
    LegacyState h;
    LCU(cudaMemcpy(&h, g.d_state, sizeof h, cudaMemcpyDeviceToHost));
    memcpy(h.C, eq_coeff, sizeof h.C);
    memcpy(h.G, gain, sizeof h.G);
    h.KY = *ky;
    memcpy(h.x, x5, sizeof h.x);
    LCU(cudaMemcpy(g.d_state, &h, sizeof h, cudaMemcpyHostToDevice));
    legacy_op_kernel<<<1, 32>>>(g.d_state, op, ref, 0);
    g_launch_count++;
    LCU(cudaGetLastError());
    LCU(cudaMemcpy(&h, g.d_state, sizeof h, cudaMemcpyDeviceToHost));
    memcpy(eq_coeff, h.C, sizeof h.C);
    memcpy(gain, h.G, sizeof h.G);
    *ky = h.KY;
    if (ret) *ret = h.ret;
    if (dibit) *dibit = h.dibit;
    return SC_OK;
}

static int simple_op(int op, float a, float b, int arg, LegacyState *h) {
    int rc = ensure();
    if (rc != SC_OK) return rc;
    const float2 v = make_float2(a, b);
    LCU(cudaMemcpy(&g.d_state->x[0], &v, sizeof v, cudaMemcpyHostToDevice));
    legacy_op_kernel<<<1, 32>>>(g.d_state, op, 0.0f, arg);
    g_launch_count++;
    LCU(cudaGetLastError());
    LCU(cudaMemcpy(h, g.d_state, sizeof *h, cudaMemcpyDeviceToHost));
    return SC_OK;
}

extern "C" int scl_scramble_init(int sr) {
    std::lock_guard<std::mutex> lock(g.mu);
    LegacyState h;
    return simple_op(5, 0.f, 0.f, sr, &h);
}

extern "C" int scl_scramble(uint8_t *v, int sr) {
    std::lock_guard<std::mutex> lock(g.mu);
    LegacyState h;
    int rc = simple_op(6, 0.f, 0.f, (int) *v | (sr << 8), &h);
    if (rc == SC_OK) *v = (uint8_t) h.dibit;
    return rc;
}

extern "C" int scl_misc(int op, float a, float b, float *out) {
    std::lock_guard<std::mutex> lock(g.mu);
    LegacyState h;
    int rc = simple_op(7 + op, a, b, 0, &h);
    if (rc != SC_OK) return rc;
    if (op == 0) {
        out[0] = h.ret;
    } else {
        out[0] = h.x[0].x;
        out[1] = h.x[0].y;
    }
    return SC_OK;
}

// qpsk_rx_frame(), src/qpsk.c:133-239, on a one-stream bank.  The reference's globals stay coherent with it:
// the descrambler register RXMemory is the one scramble(&x, rx) / scramble_init(rx) / data_eq() use
// (src/scramble.c:42, src/equalizer.c:87), and eq_coeff, kalman_gain, kalman_y and the internal u / d are left
// as the frame's last data_eq() left them (src/kalman.c:19-35), so a caller may continue with train_eq().
extern "C" int scl_rx_frame(const int16_t *in, uint8_t *bits, float *eq_coeff, float *gain, float *ky) {
    std::lock_guard<std::mutex> lock(g.mu);
    int rc = ensure();
    if (rc != SC_OK) return rc;
    if (!g.bank) {
        // firwide = false, FOFFSET = 0 as compiled into the reference (qpsk.c:60,67)
        rc = sc_create(&g.bank, g.device, 1, SC_FLAG_DEBUG_EQ, 0.0f);
        if (rc != SC_OK) return rc;
        LCU(cudaMalloc(&g.d_track_state, TRACK_STATE_FLOATS * sizeof(float)));
        modem_set_state_dbg(g.bank, g.d_track_state);
    }
    LegacyState h;
    LCU(cudaMemcpy(&h, g.d_state, sizeof h, cudaMemcpyDeviceToHost));
    modem_set_lfsr_rx(g.bank, (uint16_t) h.lfsr_rx);
    sc_frame_result r;
    float eq[10];
    rc = sc_rx_frames_host(g.bank, in, SC_FRAME_SIZE, 1, &r, 1, eq);
    if (rc != SC_OK) return rc;
    float ts[TRACK_STATE_FLOATS];
    LCU(cudaMemcpy(ts, g.d_track_state, sizeof ts, cudaMemcpyDeviceToHost));
    memcpy(h.C, ts, sizeof h.C);
    memcpy(h.G, ts + 10, sizeof h.G);
    memcpy(h.U, ts + 20, sizeof h.U);
    memcpy(h.D, ts + 40, sizeof h.D);
    h.KY = ts[45];
    h.lfsr_rx = modem_lfsr_rx(g.bank);
    LCU(cudaMemcpy(g.d_state, &h, sizeof h, cudaMemcpyHostToDevice));
    memcpy(eq_coeff, h.C, sizeof h.C);
    memcpy(gain, h.G, sizeof h.G);
    *ky = h.KY;
    if (r.valid) sc_unpack_bits(&r, 1, bits);
    return r.valid ? 1 : 0;
}

extern "C" int scl_tx_frame(int16_t *samples, const float *symbols, int length, int preamble) {
    std::lock_guard<std::mutex> lock(g.mu);
    int rc = ensure();
    if (rc != SC_OK) return rc;
    const int n = length * CYC;
    DevBuf<float2> d_sym, d_sig, d_ph;
    DevBuf<int16_t> d_out;
    LCU(d_sym.alloc((size_t) length));
    LCU(d_sig.alloc((size_t) n));
    LCU(d_ph.alloc((size_t) n));
    LCU(d_out.alloc((size_t) n));
    LCU(cudaMemcpy(d_sym, symbols, (size_t) length * sizeof(float2), cudaMemcpyHostToDevice));
    legacy_tx_stuff_kernel<<<(n + 255) / 256, 256>>>(d_sym, length, d_sig);
    g_launch_count++;
    LCU(launch_fir_batch(false, 1, g.d_txfilter, d_sig, n, n, 0));             // fir(tx_filter, firwide, ...), qpsk.c:296
    LCU(launch_nco_table(g.d_txphase, g.tx_rect, NCO_SINGLE, n, 1, 1.0f, d_ph, 0));   // phasor recurrence + renorm, :301-306
    legacy_tx_mix_kernel<<<(n + 255) / 256, 256>>>(d_sig, d_ph, n, preamble ? 8192.0f : 16384.0f, d_out);
    g_launch_count++;
    LCU(cudaGetLastError());
    LCU(cudaMemcpy(samples, d_out, (size_t) n * sizeof(int16_t), cudaMemcpyDeviceToHost));
    return SC_OK;
}

static int plan_from_factors(int n, int inverse, const int *factors, FftPlan *plan) {
    plan->n = n;
    plan->inverse = inverse ? 1 : 0;
    int k = 0, rem = n;
    while (rem > 1 && k < 32) {
        plan->p[k] = factors[2 * k];
        plan->m[k] = factors[2 * k + 1];
        if (plan->p[k] < 2 || plan->p[k] * plan->m[k] != rem) return api_fail(SC_EINVAL, "fft: corrupt configuration");
        rem = plan->m[k];
        k++;
    }
    if (n == 1) {
        plan->p[0] = 1;
        plan->m[0] = 1;
        k = 1;
    }
    plan->n_stages = k;
    return SC_OK;
}

static int fft_host(int n, int inverse, int mode, const int *factors, const float *twiddles, const float *super_tw,
                    const float *in, float *out) {
    int rc = ensure();
    if (rc != SC_OK) return rc;
    FftPlan plan;
    if ((rc = plan_from_factors(n, inverse, factors, &plan)) != SC_OK) return rc;
    const size_t in_c = mode == 2 ? (size_t) n + 1 : (size_t) n, out_c = mode == 1 ? (size_t) n + 1 : (size_t) n;
    DevBuf<float2> d_tw, d_st, d_in, d_out, d_scratch;
    DevBuf<int> d_perm;
    std::vector<int> perm((size_t) n);
    fft_make_permutation(plan, perm.data());
    LCU(d_perm.alloc((size_t) n));
    LCU(cudaMemcpy(d_perm, perm.data(), (size_t) n * sizeof(int), cudaMemcpyHostToDevice));
    LCU(d_tw.alloc((size_t) n));
    LCU(d_in.alloc(in_c));
    LCU(d_out.alloc(out_c));
    LCU(cudaMemcpy(d_tw, twiddles, (size_t) n * sizeof(float2), cudaMemcpyHostToDevice));
    LCU(cudaMemcpy(d_in, in, in_c * sizeof(float2), cudaMemcpyHostToDevice));
    if (mode != 0) {
        LCU(d_st.alloc((size_t) std::max(n / 2, 1)));
        LCU(cudaMemcpy(d_st, super_tw, (size_t) (n / 2) * sizeof(float2), cudaMemcpyHostToDevice));
    }
    if (fft_needs_scratch(n)) LCU(d_scratch.alloc((size_t) 2 * n));
    LCU(launch_fft(plan, d_tw, d_st, d_perm, mode, d_in, d_out, d_scratch, 1, 0));
    LCU(cudaMemcpy(out, d_out, out_c * sizeof(float2), cudaMemcpyDeviceToHost));
    return SC_OK;
}

extern "C" int scl_fft(int nfft, int inverse, const int *factors, const float *twiddles, const float *in, float *out) {
    std::lock_guard<std::mutex> lock(g.mu);
    return fft_host(nfft, inverse, 0, factors, twiddles, nullptr, in, out);
}

extern "C" int scl_fftr(int ncfft, int inverse, int mode, const int *factors, const float *twiddles,
                        const float *super_twiddles, const float *in, float *out) {
    std::lock_guard<std::mutex> lock(g.mu);
    return fft_host(ncfft, inverse, mode, factors, twiddles, super_twiddles, in, out);
}

extern "C" void scl_kf_factor(int n, int *facbuf) {
    FftPlan plan;
    fft_make_plan(n, 0, &plan);
    for (int k = 0; k < plan.n_stages; k++) {
        facbuf[2 * k] = plan.p[k];
        facbuf[2 * k + 1] = plan.m[k];
    }
}
extern "C" void scl_twiddles(int n, int inverse, float *tw) { fft_make_twiddles(n, inverse, (float2 *) tw); }
extern "C" void scl_super_twiddles(int ncfft, int inverse, float *tw) {
    fft_make_super_twiddles(ncfft, inverse, (float2 *) tw);
}
