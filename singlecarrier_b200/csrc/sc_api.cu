// sc_api.cu -- host side of libsinglecarrier_b200.so: the modem bank handle and the C ABI.
//
// Host logic only: state ownership, the per-call launch sequence, slab/stream pipelining and
// host<->device staging.  All modem arithmetic runs in the sm_100a kernels; the two scalars the
// reference takes from glibc (cosf/sinf of the NCO step, qpsk.c:376,428) are evaluated here with
// the same libm, everything else (the phasor recurrences included) is done on the device.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <map>
#include <tuple>
#include <mutex>
#include <new>
#include <vector>

#include "sc_kernels.h"
#include "sc_common.cuh"

namespace sc {
std::atomic<unsigned long long> g_launch_count{0};
}

using namespace sc;

static thread_local char g_err[512] = "";

static int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}

#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(e_ == cudaErrorMemoryAllocation ? SC_ENOMEM : SC_ECUDA, "%s:%d %s: %s",    \
                        __FILE__, __LINE__, #call, cudaGetErrorString(e_));                        \
    } while (0)

static const int N_PIPE = 3;        // internal CUDA streams (slabs in flight)
static const int H2D_FIRST = 80, H2D_COUNT = 1624;   // sample columns of a frame the RX chain can read (host path)

struct sc_modem {
    int device = 0;
    int64_t n = 0;                  // streams
    int64_t n_pad = 0;              // rounded up to 128 for the tiled window layout
    uint32_t flags = 0;
    float foffset = 0.f;
    bool wide = false, debug_eq = false;
    uint32_t call = 0;              // qpsk_rx_frame() calls made so far
    uint16_t lfsr_rx = 0x4A80;      // RXMemory, src/scramble.c:42, seeded at qpsk.c:434
    float2 rx_rect, tx_rect;

    // per-stream device state
    int *timing[2] = {nullptr, nullptr};    // rx_timing at entry of call n in timing[n & 1]
    int *timing_cold = nullptr;             // FINE_TIMING_OFFSET for every stream (source of resets)
    float2 *win = nullptr;                  // tracker windows, tiles of 32 streams x WIN_ROWS
    int *max_index = nullptr;
    float *max_value = nullptr;
    int16_t *hist = nullptr;                // last frame of the previous batch, [n][1880]

    // shared tables
    float2 *rx_phase = nullptr;             // fbb_rx_phase (device scalar)
    float2 *mix_table = nullptr;            // slot 0 = frame of call-1, slots 1.. = this batch
    int mix_cap = 0;                        // frames
    int *seg_len = nullptr;
    int seg_cap = 0;

    cudaStream_t pipe[N_PIPE] = {};
    cudaEvent_t ev_start = nullptr, ev_done[N_PIPE] = {};

    // staging for the host entry point
    int16_t *stage_in[N_PIPE] = {};
    sc_frame_result *stage_res[N_PIPE] = {};
    float *stage_eq[N_PIPE] = {};
    size_t stage_in_cap = 0, stage_res_cap = 0, stage_eq_cap = 0;

    // options / per-kernel profiling (bench.py's roofline leg)
    int slab_parts = 0;                     // 0 = default (2 x N_PIPE slabs)
    bool profile = false;
    std::vector<cudaEvent_t> ev_fe, ev_tk;  // (start, stop) pairs
};

// ---- scrambler keystream (integer LFSR, src/scramble.c:57-69) ----------------------------------
static inline unsigned lfsr_step(uint16_t &m) {
    unsigned out = ((m >> 1) ^ m) & 1u;
    m = (uint16_t) ((m >> 1) | (out << 14));
    return out;
}
static uint64_t keystream_next_word(uint16_t &m) {
    uint64_t w = 0;
    for (int j = 0; j < SC_BITS_PER_CALL; j++) w |= (uint64_t) lfsr_step(m) << j;
    return w;
}

extern "C" uint64_t sc_keystream_word(uint32_t call_index) {
    uint16_t m = 0x4A80;
    uint64_t w = 0;
    for (uint32_t n = 0; n <= call_index; n++) w = keystream_next_word(m);
    return w;
}

// cmplx(TAU * f / FS), headers/qpsk_internal.h:60,67: TAU is a double because glibc's M_PI is
static float2 nco_rect(float freq_hz) {
    double tau = 2.0f * M_PI;
    float x = (float) (tau * freq_hz / 8000.0f);
    return make_float2(cosf(x), sinf(x));
}

extern "C" const char *sc_last_error(void) { return g_err; }
extern "C" const char *sc_version(void) { return "singlecarrier_b200 0.1 (sm_100a)"; }
extern "C" uint64_t sc_launch_count(void) { return g_launch_count; }
extern "C" int sc_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}
extern "C" int64_t sc_n_streams(const sc_modem *m) { return m ? m->n : 0; }
extern "C" uint32_t sc_call_index(const sc_modem *m) { return m ? m->call : 0; }

static int modem_cold(sc_modem *m) {
    CU(cudaSetDevice(m->device));
    const int64_t np = m->n_pad;
    // rx_timing = FINE_TIMING_OFFSET (qpsk.c:53); windows of calls 0 and 1 are the zero-initialised
    // decimated_frame (qpsk.c:42), whose search gives max_index 0 / max_value 0.0
    // (hist needs no reset: it is only read by a batch that continues a previous one, which saved it)
    CU(cudaMemcpyAsync(m->timing[0], m->timing_cold, np * sizeof(int), cudaMemcpyDeviceToDevice, m->pipe[0]));
    CU(cudaMemcpyAsync(m->timing[1], m->timing_cold, np * sizeof(int), cudaMemcpyDeviceToDevice, m->pipe[0]));
    CU(cudaMemsetAsync(m->win, 0, (size_t) (np / 32) * WIN_ROWS * 32 * sizeof(float2), m->pipe[0]));
    CU(cudaMemsetAsync(m->max_index, 0, np * sizeof(int), m->pipe[0]));
    CU(cudaMemsetAsync(m->max_value, 0, np * sizeof(float), m->pipe[0]));
    const float2 one = make_float2(1.0f, 0.0f);                         // cmplx(0.0f), qpsk.c:427
    CU(cudaMemcpyAsync(m->rx_phase, &one, sizeof one, cudaMemcpyHostToDevice, m->pipe[0]));
    CU(cudaStreamSynchronize(m->pipe[0]));
    m->call = 0;
    m->lfsr_rx = 0x4A80;
    return SC_OK;
}

extern "C" int sc_create(sc_modem **out, int device, int64_t n_streams, uint32_t flags, float foffset_hz) {
    if (!out || n_streams <= 0 || n_streams > (int64_t) 1 << 30) return fail(SC_EINVAL, "sc_create: bad arguments");
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(SC_ECUDA, "sc_create: no CUDA device (this library has no CPU path)");
    }
    if (device < 0 || device >= ndev) return fail(SC_EINVAL, "sc_create: device %d of %d", device, ndev);
    sc_modem *m = new (std::nothrow) sc_modem();
    if (!m) return fail(SC_ENOMEM, "sc_create: host allocation");
    m->device = device;
    m->n = n_streams;
    m->n_pad = (n_streams + 127) / 128 * 128;
    m->flags = flags;
    m->wide = (flags & SC_FLAG_WIDE) != 0;
    m->debug_eq = (flags & SC_FLAG_DEBUG_EQ) != 0;
    m->foffset = foffset_hz;
    if (const char *e = getenv("SC_SLAB_PARTS")) m->slab_parts = atoi(e);
    m->rx_rect = nco_rect(-1100.0f + foffset_hz);                       // qpsk.c:428
    m->tx_rect = nco_rect(1100.0f);                                     // qpsk.c:376
    int rc = SC_OK;
    auto body = [&]() -> int {
        CU(cudaSetDevice(device));
        const int64_t np = m->n_pad;
        CU(cudaMalloc(&m->timing[0], np * sizeof(int)));
        CU(cudaMalloc(&m->timing[1], np * sizeof(int)));
        CU(cudaMalloc(&m->timing_cold, np * sizeof(int)));
        {
            std::vector<int> t(np, 3);
            CU(cudaMemcpy(m->timing_cold, t.data(), np * sizeof(int), cudaMemcpyHostToDevice));
        }
        CU(cudaMalloc(&m->win, (size_t) (np / 32) * WIN_ROWS * 32 * sizeof(float2)));
        CU(cudaMalloc(&m->max_index, np * sizeof(int)));
        CU(cudaMalloc(&m->max_value, np * sizeof(float)));
        CU(cudaMalloc(&m->hist, (size_t) m->n * FRAME * sizeof(int16_t)));
        CU(cudaMemset(m->hist, 0, (size_t) m->n * FRAME * sizeof(int16_t)));
        CU(cudaMalloc(&m->rx_phase, sizeof(float2)));
        for (int i = 0; i < N_PIPE; i++) {
            CU(cudaStreamCreateWithFlags(&m->pipe[i], cudaStreamNonBlocking));
            CU(cudaEventCreateWithFlags(&m->ev_done[i], cudaEventDisableTiming));
        }
        CU(cudaEventCreateWithFlags(&m->ev_start, cudaEventDisableTiming));
        return modem_cold(m);
    };
    rc = body();
    if (rc != SC_OK) {
        sc_destroy(m);
        return rc;
    }
    *out = m;
    return SC_OK;
}

extern "C" void sc_destroy(sc_modem *m) {
    if (!m) return;
    cudaSetDevice(m->device);
    cudaDeviceSynchronize();
    cudaFree(m->timing[0]);
    cudaFree(m->timing[1]);
    cudaFree(m->timing_cold);
    for (cudaEvent_t e : m->ev_fe) cudaEventDestroy(e);
    for (cudaEvent_t e : m->ev_tk) cudaEventDestroy(e);
    cudaFree(m->win);
    cudaFree(m->max_index);
    cudaFree(m->max_value);
    cudaFree(m->hist);
    cudaFree(m->rx_phase);
    cudaFree(m->mix_table);
    cudaFree(m->seg_len);
    for (int i = 0; i < N_PIPE; i++) {
        cudaFree(m->stage_in[i]);
        cudaFree(m->stage_res[i]);
        cudaFree(m->stage_eq[i]);
        if (m->pipe[i]) cudaStreamDestroy(m->pipe[i]);
        if (m->ev_done[i]) cudaEventDestroy(m->ev_done[i]);
    }
    if (m->ev_start) cudaEventDestroy(m->ev_start);
    cudaGetLastError();
    delete m;
}

extern "C" int sc_reset(sc_modem *m) {
    if (!m) return fail(SC_EINVAL, "sc_reset: null handle");
    CU(cudaSetDevice(m->device));
    for (int i = 0; i < N_PIPE; i++) CU(cudaStreamSynchronize(m->pipe[i]));
    return modem_cold(m);
}

// Make sure the mix table can hold slot 0 + n_frames frames and generate this batch's phasors
// (calls call .. call+n_frames-1 into slots 1..n_frames) on stream st.
static int prepare_tables(sc_modem *m, int n_frames, cudaStream_t st) {
    if (m->mix_cap < n_frames + 1) {
        float2 *nt = nullptr;
        CU(cudaMalloc(&nt, (size_t) (n_frames + 1) * FRAME * sizeof(float2)));
        if (m->mix_table) {
            CU(cudaMemcpyAsync(nt, m->mix_table, (size_t) FRAME * sizeof(float2), cudaMemcpyDeviceToDevice, st));
            CU(cudaStreamSynchronize(st));
            CU(cudaFree(m->mix_table));
        } else {
            CU(cudaMemsetAsync(nt, 0, (size_t) FRAME * sizeof(float2), st));
        }
        m->mix_table = nt;
        m->mix_cap = n_frames + 1;
    }
    if (m->seg_cap < n_frames) {
        if (m->seg_len) CU(cudaFree(m->seg_len));
        CU(cudaMalloc(&m->seg_len, (size_t) n_frames * sizeof(int)));
        std::vector<int> seg(n_frames, FRAME);
        CU(cudaMemcpy(m->seg_len, seg.data(), (size_t) n_frames * sizeof(int), cudaMemcpyHostToDevice));
        m->seg_cap = n_frames;
    }
    // stored pre-multiplied by 1/16384: phasor*((float)in/16384) == (phasor/16384)*(float)in exactly
    CU(launch_nco_table(m->rx_phase, m->rx_rect, m->seg_len, n_frames, 1.0f / 16384.0f,
                        m->mix_table + FRAME, st));
    return SC_OK;
}

static cudaError_t prof_mark(std::vector<cudaEvent_t> &v, cudaStream_t st) {
    cudaEvent_t e;
    cudaError_t rc = cudaEventCreate(&e);
    if (rc != cudaSuccess) return rc;
    v.push_back(e);
    return cudaEventRecord(e, st);
}

// One slab of streams [s0, s0+ns): the call loop.  `in` points at stream s0's first sample of the
// first frame of this block, `results`/`eq_dbg` at stream s0's record of the block's first call.
// call0 = index of the block's first call; kw = its keystream words; mix0 = table slot of frame
// call0-1 (the frame filtered by call call0).
static int run_slab(sc_modem *m, cudaStream_t st, int64_t s0, int ns, const int16_t *in, int64_t stride,
                    int n_frames, uint32_t call0, const uint64_t *kw, const float2 *mix0,
                    sc_frame_result *results, int64_t result_stride, float *eq_dbg, bool save_hist) {
    float2 *win = m->win + (size_t) (s0 / 32) * WIN_ROWS * 32;
    for (int j = 0; j < n_frames; j++) {
        const uint32_t n = call0 + (uint32_t) j;
        int *tc = m->timing[n & 1] + s0, *tn = m->timing[(n + 1) & 1] + s0;
        if (m->profile) CU(prof_mark(m->ev_tk, st));
        CU(launch_track(m->debug_eq && eq_dbg != nullptr, win, m->max_index + s0, m->max_value + s0, tc, tn,
                        results + j, result_stride, eq_dbg ? eq_dbg + (size_t) j * 10 : nullptr, n, kw[j], ns, st));
        if (m->profile) CU(prof_mark(m->ev_tk, st));
        if (n >= 1) {
            const int16_t *frame;
            int64_t fstride;
            if (j >= 1) {
                frame = in + (size_t) (j - 1) * FRAME;
                fstride = stride;
            } else {
                frame = m->hist + (size_t) s0 * FRAME;
                fstride = FRAME;
            }
            if (m->profile) CU(prof_mark(m->ev_fe, st));
            CU(launch_frontend(m->wide, frame, fstride, mix0 + (size_t) j * FRAME, tc, tn, win, m->max_index + s0,
                               m->max_value + s0, ns, st));
            if (m->profile) CU(prof_mark(m->ev_fe, st));
        }
    }
    if (save_hist && n_frames > 0) {
        CU(cudaMemcpy2DAsync(m->hist + (size_t) s0 * FRAME, FRAME * sizeof(int16_t),
                             in + (size_t) (n_frames - 1) * FRAME, (size_t) stride * sizeof(int16_t),
                             FRAME * sizeof(int16_t), (size_t) ns, cudaMemcpyDeviceToDevice, st));
    }
    return SC_OK;
}

static int64_t pick_slab(int64_t n, int64_t lo, int64_t parts) {
    int64_t s = (n + parts - 1) / parts;
    s = std::max<int64_t>(s, lo);
    s = (s + 127) / 128 * 128;
    return s;
}

extern "C" int sc_rx_frames_dev(sc_modem *m, const int16_t *in, int64_t stream_stride, int n_frames,
                                sc_frame_result *results, int64_t result_stride, float *eq_dbg, void *stream) {
    if (!m || !in || !results || n_frames < 0) return fail(SC_EINVAL, "sc_rx_frames_dev: bad arguments");
    if (n_frames == 0) return SC_OK;
    if (stream_stride < (int64_t) n_frames * FRAME || result_stride < n_frames)
        return fail(SC_EINVAL, "sc_rx_frames_dev: stride smaller than the batch");
    if (eq_dbg && !m->debug_eq) return fail(SC_EINVAL, "sc_rx_frames_dev: eq_dbg needs SC_FLAG_DEBUG_EQ");
    CU(cudaSetDevice(m->device));
    cudaStream_t user = (cudaStream_t) stream;

    std::vector<uint64_t> kw(n_frames);
    for (int j = 0; j < n_frames; j++) kw[j] = keystream_next_word(m->lfsr_rx);

    // order after the caller's stream, build the tables once, fan out over the pipe streams
    CU(cudaEventRecord(m->ev_start, user));
    CU(cudaStreamWaitEvent(m->pipe[0], m->ev_start, 0));
    int rc = prepare_tables(m, n_frames, m->pipe[0]);
    if (rc != SC_OK) return rc;
    CU(cudaEventRecord(m->ev_start, m->pipe[0]));
    for (int i = 1; i < N_PIPE; i++) CU(cudaStreamWaitEvent(m->pipe[i], m->ev_start, 0));

    // default: two slabs, each on its own stream, so one slab's tail waves overlap the other's kernels
    const int64_t slab = pick_slab(m->n, m->slab_parts > 0 ? 128 : 8192, m->slab_parts > 0 ? m->slab_parts : 2);
    int k = 0;
    for (int64_t s0 = 0; s0 < m->n; s0 += slab, k++) {
        const int ns = (int) std::min<int64_t>(slab, m->n - s0);
        rc = run_slab(m, m->pipe[k % N_PIPE], s0, ns, in + (size_t) s0 * stream_stride, stream_stride, n_frames,
                      m->call, kw.data(), m->mix_table, results + (size_t) s0 * result_stride, result_stride,
                      eq_dbg ? eq_dbg + (size_t) s0 * result_stride * 10 : nullptr, true);
        if (rc != SC_OK) return rc;
    }
    // the last frame's phasors become slot 0 of the next batch
    for (int i = 0; i < N_PIPE; i++) {
        CU(cudaEventRecord(m->ev_done[i], m->pipe[i]));
        CU(cudaStreamWaitEvent(user, m->ev_done[i], 0));
    }
    CU(cudaMemcpyAsync(m->mix_table, m->mix_table + (size_t) n_frames * FRAME, FRAME * sizeof(float2),
                       cudaMemcpyDeviceToDevice, user));
    CU(cudaEventRecord(m->ev_start, user));
    for (int i = 0; i < N_PIPE; i++) CU(cudaStreamWaitEvent(m->pipe[i], m->ev_start, 0));
    m->call += (uint32_t) n_frames;
    return SC_OK;
}

template <typename T>
static int grow(T **p, size_t *cap, size_t want, int count) {
    if (*cap >= want) return SC_OK;
    for (int i = 0; i < count; i++) {
        if (p[i]) CU(cudaFree(p[i]));
        p[i] = nullptr;
        CU(cudaMalloc(&p[i], want));
    }
    *cap = want;
    return SC_OK;
}

extern "C" int sc_rx_frames_host(sc_modem *m, const int16_t *in, int64_t stream_stride, int n_frames,
                                 sc_frame_result *results, int64_t result_stride, float *eq_dbg) {
    if (!m || !in || !results || n_frames < 0) return fail(SC_EINVAL, "sc_rx_frames_host: bad arguments");
    if (n_frames == 0) return SC_OK;
    if (stream_stride < (int64_t) n_frames * FRAME || result_stride < n_frames)
        return fail(SC_EINVAL, "sc_rx_frames_host: stride smaller than the batch");
    if (eq_dbg && !m->debug_eq) return fail(SC_EINVAL, "sc_rx_frames_host: eq_dbg needs SC_FLAG_DEBUG_EQ");
    CU(cudaSetDevice(m->device));

    std::vector<uint64_t> kw(n_frames);
    for (int j = 0; j < n_frames; j++) kw[j] = keystream_next_word(m->lfsr_rx);

    // slabs of streams x blocks of frames; slab k always runs on pipe k % N_PIPE, so its blocks
    // stay in order and the staging buffer of that pipe is reused safely (stream order)
    const int64_t slab = pick_slab(m->n, m->slab_parts > 0 ? 128 : 4096, m->slab_parts > 0 ? m->slab_parts : 2 * N_PIPE);
    const size_t frame_bytes = FRAME * sizeof(int16_t);
    int fblk = (int) std::max<int64_t>(1, std::min<int64_t>(n_frames, ((int64_t) 768 << 20) / (slab * (int64_t) frame_bytes)));
    int rc;
    if ((rc = grow(m->stage_in, &m->stage_in_cap, (size_t) slab * fblk * frame_bytes, N_PIPE)) != SC_OK) return rc;
    if ((rc = grow(m->stage_res, &m->stage_res_cap, (size_t) slab * fblk * sizeof(sc_frame_result), N_PIPE)) != SC_OK) return rc;
    if (eq_dbg && (rc = grow(m->stage_eq, &m->stage_eq_cap, (size_t) slab * fblk * 10 * sizeof(float), N_PIPE)) != SC_OK) return rc;

    if ((rc = prepare_tables(m, n_frames, m->pipe[0])) != SC_OK) return rc;
    CU(cudaEventRecord(m->ev_start, m->pipe[0]));
    for (int i = 1; i < N_PIPE; i++) CU(cudaStreamWaitEvent(m->pipe[i], m->ev_start, 0));

    for (int f0 = 0; f0 < n_frames; f0 += fblk) {
        const int nf = std::min(fblk, n_frames - f0);
        int k = 0;
        for (int64_t s0 = 0; s0 < m->n; s0 += slab, k++) {
            const int p = k % N_PIPE;
            cudaStream_t st = m->pipe[p];
            const int ns = (int) std::min<int64_t>(slab, m->n - s0);
            const int64_t dstride = (int64_t) nf * FRAME;
            // Only samples [80, 1704) of a frame can influence any output: rx_timing is 128..255 after
            // the first call, so the front-end reads samples rx_timing-48 .. rx_timing+1445 (+ pair
            // slack) and nothing else.  Copying just those columns saves 14 % of the
            // PCIe traffic; the other staging columns are never read.
            for (int j = 0; j < nf; j++)
                CU(cudaMemcpy2DAsync(m->stage_in[p] + (size_t) j * FRAME + H2D_FIRST, (size_t) dstride * sizeof(int16_t),
                                     in + (size_t) s0 * stream_stride + (size_t) (f0 + j) * FRAME + H2D_FIRST,
                                     (size_t) stream_stride * sizeof(int16_t), (size_t) H2D_COUNT * sizeof(int16_t),
                                     (size_t) ns, cudaMemcpyHostToDevice, st));
            rc = run_slab(m, st, s0, ns, m->stage_in[p], dstride, nf, m->call + (uint32_t) f0, kw.data() + f0,
                          m->mix_table + (size_t) f0 * FRAME, m->stage_res[p], nf, eq_dbg ? m->stage_eq[p] : nullptr,
                          true);
            if (rc != SC_OK) return rc;
            CU(cudaMemcpy2DAsync(results + (size_t) s0 * result_stride + f0, (size_t) result_stride * sizeof(sc_frame_result),
                                 m->stage_res[p], (size_t) nf * sizeof(sc_frame_result),
                                 (size_t) nf * sizeof(sc_frame_result), (size_t) ns, cudaMemcpyDeviceToHost, st));
            if (eq_dbg)
                CU(cudaMemcpy2DAsync(eq_dbg + ((size_t) s0 * result_stride + f0) * 10, (size_t) result_stride * 10 * sizeof(float),
                                     m->stage_eq[p], (size_t) nf * 10 * sizeof(float), (size_t) nf * 10 * sizeof(float),
                                     (size_t) ns, cudaMemcpyDeviceToHost, st));
        }
    }
    for (int i = 0; i < N_PIPE; i++) CU(cudaStreamSynchronize(m->pipe[i]));
    CU(cudaMemcpyAsync(m->mix_table, m->mix_table + (size_t) n_frames * FRAME, FRAME * sizeof(float2),
                       cudaMemcpyDeviceToDevice, m->pipe[0]));
    CU(cudaStreamSynchronize(m->pipe[0]));
    m->call += (uint32_t) n_frames;
    return SC_OK;
}

extern "C" void sc_unpack_bits(const sc_frame_result *results, int64_t n_records, uint8_t *rows) {
    for (int64_t k = 0; k < n_records; k++) {
        if (!results[k].valid) continue;
        const uint64_t w = results[k].bits;
        uint8_t *row = rows + k * SC_BITS_PER_CALL;
        for (int j = 0; j < SC_BITS_PER_CALL; j++) row[j] = (uint8_t) ((w >> j) & 1u);
    }
}

extern "C" int sc_nco_table_host(sc_modem *m, int tx, uint32_t first_call, int n_calls, float *out) {
    if (!m || !out || n_calls <= 0) return fail(SC_EINVAL, "sc_nco_table_host: bad arguments");
    CU(cudaSetDevice(m->device));
    // regenerate from the cold phasor: first_call + n_calls segments, return the last n_calls
    const int total = (int) first_call + n_calls;
    std::vector<int> seg;
    if (tx) {
        for (int c = 0; c < total; c++) {          // one "call" = one packet: 640 + 8 x 155
            seg.push_back(640);
            for (int j = 0; j < 8; j++) seg.push_back(155);
        }
    } else {
        seg.assign(total, FRAME);
    }
    int *d_seg = nullptr;
    float2 *d_ph = nullptr, *d_out = nullptr;
    const float2 one = make_float2(1.0f, 0.0f);
    CU(cudaMalloc(&d_seg, seg.size() * sizeof(int)));
    CU(cudaMalloc(&d_ph, sizeof(float2)));
    CU(cudaMalloc(&d_out, (size_t) total * FRAME * sizeof(float2)));
    CU(cudaMemcpy(d_seg, seg.data(), seg.size() * sizeof(int), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(d_ph, &one, sizeof one, cudaMemcpyHostToDevice));
    CU(launch_nco_table(d_ph, tx ? m->tx_rect : m->rx_rect, d_seg, (int) seg.size(), 1.0f, d_out, 0));
    CU(cudaMemcpy(out, d_out + (size_t) first_call * FRAME, (size_t) n_calls * FRAME * sizeof(float2),
                  cudaMemcpyDeviceToHost));
    cudaFree(d_seg);
    cudaFree(d_ph);
    cudaFree(d_out);
    return SC_OK;
}

// ---- accessors used by the TX / stage translation units ----------------------------------------
namespace sc {
int modem_device(const sc_modem *m) { return m->device; }
int64_t modem_n(const sc_modem *m) { return m->n; }
bool modem_wide(const sc_modem *m) { return m->wide; }
float2 modem_tx_rect(const sc_modem *m) { return m->tx_rect; }
int api_fail(int code, const char *msg) { return fail(code, "%s", msg); }
}  // namespace sc

// ---- TX ---------------------------------------------------------------------------------------
static int tx_common(sc_modem *m, const uint8_t *bits, uint8_t *bits_out, uint64_t seed, int n_packets,
                     int gap_samples, const int32_t *lead_in, const sc_channel *ch, int16_t *out,
                     int64_t stream_stride, int64_t samples_per_stream, void *stream) {
    if (!m || !out || n_packets < 0 || gap_samples < 0 || samples_per_stream < 0 || stream_stride < samples_per_stream)
        return fail(SC_EINVAL, "sc_tx: bad arguments");
    CU(cudaSetDevice(m->device));
    cudaStream_t st = (cudaStream_t) stream;
    // TX phasor table from the cold phasor: renormalised after every qpsk_tx_frame() call, i.e.
    // after the 640-sample preamble and after each 155-sample data frame (qpsk.c:306)
    std::vector<int> seg;
    for (int p = 0; p < n_packets; p++) {
        seg.push_back(SC_PREAMBLE_LENGTH * CYC);
        for (int j = 0; j < 8; j++) seg.push_back(SC_DATA_SYMBOLS * CYC);
    }
    int *d_seg = nullptr;
    float2 *d_ph = nullptr, *d_tab = nullptr;
    const float2 one = make_float2(1.0f, 0.0f);                        // cmplx(0.0f), qpsk.c:375
    CU(cudaMallocAsync((void **) &d_seg, std::max<size_t>(seg.size(), 1) * sizeof(int), st));
    CU(cudaMallocAsync((void **) &d_ph, sizeof(float2), st));
    CU(cudaMallocAsync((void **) &d_tab, std::max<size_t>((size_t) n_packets * FRAME, 1) * sizeof(float2), st));
    CU(cudaMemcpyAsync(d_seg, seg.data(), seg.size() * sizeof(int), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_ph, &one, sizeof one, cudaMemcpyHostToDevice, st));
    CU(cudaStreamSynchronize(st));                                     // seg/one are host temporaries
    if (n_packets > 0) CU(launch_nco_table(d_ph, m->tx_rect, d_seg, (int) seg.size(), 1.0f, d_tab, st));
    TxArgs a;
    a.bits = bits;
    a.bits_out = bits_out;
    a.seed = seed;
    a.n_packets = n_packets;
    a.gap_samples = gap_samples;
    a.lead_in = lead_in;
    a.tx_table = d_tab;
    a.out = out;
    a.stream_stride = stream_stride;
    a.samples_per_stream = samples_per_stream;
    a.n_streams = m->n;
    a.wide = m->wide;
    a.use_channel = ch != nullptr;
    memset(&a.ch, 0, sizeof a.ch);
    if (ch) a.ch = *ch;
    CU(launch_tx(a, st));
    CU(cudaFreeAsync(d_seg, st));
    CU(cudaFreeAsync(d_ph, st));
    CU(cudaFreeAsync(d_tab, st));
    return SC_OK;
}

extern "C" int sc_tx_packets_dev(sc_modem *m, const uint8_t *bits, uint8_t *bits_out, uint64_t seed, int n_packets,
                                 int gap_samples, const int32_t *lead_in, int16_t *out, int64_t stream_stride,
                                 int64_t samples_per_stream, void *stream) {
    return tx_common(m, bits, bits_out, seed, n_packets, gap_samples, lead_in, nullptr, out, stream_stride,
                     samples_per_stream, stream);
}

extern "C" int sc_tx_channel_dev(sc_modem *m, const uint8_t *bits, uint8_t *bits_out, uint64_t seed, int n_packets,
                                 int gap_samples, const int32_t *lead_in, const sc_channel *ch, int16_t *out,
                                 int64_t stream_stride, int64_t samples_per_stream, void *stream) {
    if (!ch) return fail(SC_EINVAL, "sc_tx_channel_dev: null channel");
    return tx_common(m, bits, bits_out, seed, n_packets, gap_samples, lead_in, ch, out, stream_stride,
                     samples_per_stream, stream);
}

// ---- stage entry points -----------------------------------------------------------------------
static int stage_device(int device) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(SC_ECUDA, "no CUDA device (this library has no CPU path)");
    }
    if (device < 0 || device >= ndev) return fail(SC_EINVAL, "device %d of %d", device, ndev);
    CU(cudaSetDevice(device));
    return SC_OK;
}

extern "C" int sc_fir_batch_dev(int device, int64_t n_streams, int wide, float *memory, float *sample,
                                int64_t sample_stride, int length, void *stream) {
    if (n_streams < 0 || !memory || !sample || length < 0 || sample_stride < length)
        return fail(SC_EINVAL, "sc_fir_batch_dev: bad arguments");
    int rc = stage_device(device);
    if (rc != SC_OK) return rc;
    if (n_streams == 0 || length == 0) return SC_OK;
    CU(launch_fir_batch((wide & 1) != 0, n_streams, (float2 *) memory, (float2 *) sample, sample_stride, length,
                        (cudaStream_t) stream, (wide & SC_FIR_FAST) != 0));
    return SC_OK;
}

extern "C" int sc_preamble_search_batch_dev(int device, int64_t n_streams, const float *symbols, int64_t symbol_stride,
                                            int32_t *max_index, float *max_value, void *stream) {
    if (n_streams < 0 || !symbols || !max_index || !max_value || symbol_stride < 255)
        return fail(SC_EINVAL, "sc_preamble_search_batch_dev: bad arguments");
    int rc = stage_device(device);
    if (rc != SC_OK) return rc;
    if (n_streams == 0) return SC_OK;
    CU(launch_search_batch(n_streams, (const float2 *) symbols, symbol_stride, max_index, max_value,
                           (cudaStream_t) stream));
    return SC_OK;
}

extern "C" int sc_track_decide_batch_dev(int device, int64_t n_streams, const float *symbols, int64_t symbol_stride,
                                         const int32_t *max_index, const float *max_value, int32_t *rx_timing,
                                         uint32_t call_index, sc_frame_result *results, float *eq_dbg, void *stream) {
    if (n_streams < 0 || !symbols || !max_index || !rx_timing || !results || symbol_stride < WIN)
        return fail(SC_EINVAL, "sc_track_decide_batch_dev: bad arguments");
    int rc = stage_device(device);
    if (rc != SC_OK) return rc;
    if (n_streams == 0) return SC_OK;
    CU(launch_track_window_batch(n_streams, (const float2 *) symbols, symbol_stride, max_index, max_value, rx_timing,
                                 call_index, sc_keystream_word(call_index), results, eq_dbg, (cudaStream_t) stream));
    return SC_OK;
}

// ---- FFT (fft.h) --------------------------------------------------------------------------------
// kf_factor(), src/fft.c:433-459: radix 4 first, then 2, 3, 5, 7, ...; p > floor(sqrt(n)) => p = n
namespace sc {
void fft_make_plan(int n, int inverse, FftPlan *plan) {
    plan->n = n;
    plan->inverse = inverse ? 1 : 0;
    int p = 4, k = 0;
    const float floor_sqrt = floorf(sqrtf((float) n));
    do {
        while (n % p) {
            switch (p) {
                case 4: p = 2; break;
                case 2: p = 3; break;
                default: p += 2;
            }
            if (p > floor_sqrt) p = n;
        }
        n /= p;
        plan->p[k] = p;
        plan->m[k] = n;
        k++;
    } while (n > 1);
    plan->n_stages = k;
}

// twiddles exactly as fft_alloc() computes them (src/fft.c:70-77), same libm as the reference
void fft_make_twiddles(int n, int inverse, float2 *tw) {
    const double tau = 2.0f * M_PI;
    for (int i = 0; i < n; i++) {
        float phase = (float) (-tau * (float) i / (float) n);
        if (inverse) phase *= -1.0f;
        tw[i] = make_float2(cosf(phase), sinf(phase));
    }
}

// super twiddles of fftr_alloc() (src/fft.c:118-126); ncfft = nfft/2
void fft_make_super_twiddles(int ncfft, int inverse, float2 *tw) {
    for (int i = 0; i < ncfft / 2; i++) {
        float phase = (float) (-M_PI * ((float) (i + 1) / (float) ncfft + .5f));
        if (inverse) phase *= -1.0f;
        tw[i] = make_float2(cosf(phase), sinf(phase));
    }
}
}  // namespace sc

struct FftDeviceTables {
    float2 *tw = nullptr, *super_tw = nullptr, *scratch = nullptr;
};
static std::mutex g_fft_mu;
static std::map<std::tuple<int, int, int, int>, FftDeviceTables> g_fft_tables;   // (device, n, inverse, real)

static int fft_tables(int device, int n, int inverse, bool real, FftDeviceTables *out) {
    std::lock_guard<std::mutex> lock(g_fft_mu);
    auto key = std::make_tuple(device, n, inverse ? 1 : 0, real ? 1 : 0);
    auto it = g_fft_tables.find(key);
    if (it != g_fft_tables.end()) {
        *out = it->second;
        return SC_OK;
    }
    FftDeviceTables t;
    std::vector<float2> h(n);
    fft_make_twiddles(n, inverse, h.data());
    CU(cudaMalloc(&t.tw, (size_t) n * sizeof(float2)));
    CU(cudaMemcpy(t.tw, h.data(), (size_t) n * sizeof(float2), cudaMemcpyHostToDevice));
    if (real) {
        std::vector<float2> hs(std::max(n / 2, 1));
        fft_make_super_twiddles(n, inverse, hs.data());
        CU(cudaMalloc(&t.super_tw, hs.size() * sizeof(float2)));
        CU(cudaMemcpy(t.super_tw, hs.data(), hs.size() * sizeof(float2), cudaMemcpyHostToDevice));
    }
    if ((size_t) 2 * n * sizeof(float2) > 64 * 1024)
        CU(cudaMalloc(&t.scratch, (size_t) FFT_SCRATCH_CTAS * 2 * n * sizeof(float2)));
    g_fft_tables[key] = t;
    *out = t;
    return SC_OK;
}

static int fft_run(int device, int64_t n_batches, int n_core, int inverse, int mode, const void *in, void *out,
                   void *stream) {
    int rc = stage_device(device);
    if (rc != SC_OK) return rc;
    if (n_batches == 0) return SC_OK;
    FftDeviceTables t;
    if ((rc = fft_tables(device, n_core, inverse, mode != 0, &t)) != SC_OK) return rc;
    FftPlan plan;
    fft_make_plan(n_core, inverse, &plan);
    CU(launch_fft(plan, t.tw, t.super_tw, mode, in, out, t.scratch, n_batches, (cudaStream_t) stream));
    return SC_OK;
}

extern "C" int sc_fft_batch_dev(int device, int64_t n_batches, int nfft, int inverse, const float *in, float *out,
                                void *stream) {
    if (n_batches < 0 || !in || !out || nfft < 1 || nfft > (1 << 20))
        return fail(SC_EINVAL, "sc_fft_batch_dev: bad arguments");
    return fft_run(device, n_batches, nfft, inverse, 0, in, out, stream);
}

extern "C" int sc_fftr_batch_dev(int device, int64_t n_batches, int nfft, const float *in, float *out, void *stream) {
    if (n_batches < 0 || !in || !out || nfft < 2 || (nfft & 1) || nfft > (1 << 21))
        return fail(SC_EINVAL, "sc_fftr_batch_dev: nfft must be even");
    return fft_run(device, n_batches, nfft / 2, 0, 1, in, out, stream);
}

extern "C" int sc_fftri_batch_dev(int device, int64_t n_batches, int nfft, const float *in, float *out, void *stream) {
    if (n_batches < 0 || !in || !out || nfft < 2 || (nfft & 1) || nfft > (1 << 21))
        return fail(SC_EINVAL, "sc_fftri_batch_dev: nfft must be even");
    return fft_run(device, n_batches, nfft / 2, 1, 2, in, out, stream);
}

extern "C" int sc_lock_stats_dev(int device, const sc_frame_result *results, int64_t n_streams, int64_t result_stride,
                                 int n_frames, uint64_t *counters, void *stream) {
    if (!results || !counters || n_streams < 0 || n_frames < 0 || result_stride < n_frames)
        return fail(SC_EINVAL, "sc_lock_stats_dev: bad arguments");
    int rc = stage_device(device);
    if (rc != SC_OK) return rc;
    if (n_streams == 0 || n_frames == 0) return SC_OK;
    CU(launch_lock_stats(results, n_streams, result_stride, n_frames, (unsigned long long *) counters,
                         (cudaStream_t) stream));
    return SC_OK;
}

// ---- options and per-kernel profiling ----------------------------------------------------------
extern "C" int sc_set_option(sc_modem *m, int option, int64_t value) {
    if (!m) return fail(SC_EINVAL, "sc_set_option: null handle");
    switch (option) {
        case SC_OPT_SLAB_PARTS:
            if (value < 0 || value > 1024) return fail(SC_EINVAL, "sc_set_option: slab parts out of range");
            m->slab_parts = (int) value;
            return SC_OK;
        case SC_OPT_PROFILE:
            m->profile = value != 0;
            return SC_OK;
        default:
            return fail(SC_EINVAL, "sc_set_option: unknown option %d", option);
    }
}

static int prof_sum(std::vector<cudaEvent_t> &v, double *ms, double *count) {
    *ms = 0.0;
    *count = 0.0;
    for (size_t i = 0; i + 1 < v.size(); i += 2) {
        float t = 0.f;
        CU(cudaEventElapsedTime(&t, v[i], v[i + 1]));
        *ms += t;
        *count += 1.0;
    }
    for (cudaEvent_t e : v) cudaEventDestroy(e);
    v.clear();
    return SC_OK;
}

extern "C" int sc_profile_read(sc_modem *m, double out[4]) {
    if (!m || !out) return fail(SC_EINVAL, "sc_profile_read: bad arguments");
    CU(cudaSetDevice(m->device));
    CU(cudaDeviceSynchronize());
    int rc = prof_sum(m->ev_fe, &out[0], &out[1]);
    if (rc != SC_OK) return rc;
    return prof_sum(m->ev_tk, &out[2], &out[3]);
}

// Self-test hook (tests/test_stage_gpu.py): counts float bit patterns in [lo_bits, hi_bits] for which
// the tracker's branch-free reciprocal differs from __frcp_rn.  *mismatches is a device uint64.
extern "C" int sc_selftest_rcp_dev(int device, uint32_t lo_bits, uint32_t hi_bits, uint64_t *mismatches, void *stream) {
    if (!mismatches || hi_bits < lo_bits) return fail(SC_EINVAL, "sc_selftest_rcp_dev: bad arguments");
    int rc = stage_device(device);
    if (rc != SC_OK) return rc;
    CU(launch_selftest_rcp(lo_bits, hi_bits, (unsigned long long *) mismatches, (cudaStream_t) stream));
    return SC_OK;
}
