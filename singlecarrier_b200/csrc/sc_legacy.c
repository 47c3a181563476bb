/*
 * sc_legacy.c -- the reference's single-stream symbols (include/sc_compat/singlecarrier_compat.h),
 * as a thin C99 shim over the CUDA back end.
 *
 * This file holds no modem arithmetic: it converts between the reference's calling conventions
 * (C99 complex arrays, process-global state, void returns) and the scl_* bridge in
 * sc_legacy_dev.cu, which runs the same device code the batched kernels use on a one-stream
 * batch.  A failing CUDA call cannot be reported through the reference's void signatures, so it
 * aborts with a message (no silent CPU fallback exists).
 */
#include "../../include/sc_compat/singlecarrier_compat.h"
#include "../../include/singlecarrier_b200.h"
#include "sc_legacy_internal.h"

/* ---- data symbols ------------------------------------------------------------------------------ */
complex float eq_coeff[EQ_LENGTH];                 /* src/kalman.c:19 */
complex float kalman_gain[EQ_LENGTH];              /* src/kalman.c:20 */
float kalman_y;                                    /* src/kalman.c:21 */
int preamble_frames_detected = 0;                  /* src/qpsk.c:70   */

const complex float constellation[4] = { 1.0f + 0.0f * I, 0.0f + 1.0f * I, 0.0f - 1.0f * I, -1.0f + 0.0f * I };   /* src/constants.c:11-16 */
#define SC_TABLE_PREAMBLE const int8_t preamblevalues[PREAMBLE_LENGTH]
#define SC_TABLE_RRC35    const float alpha35_root[NTAPS]
#define SC_TABLE_RRC50    const float alpha50_root[NTAPS]
#include "../../include/sc_tables.inc"

static void must(int rc, const char *what) {
    if (rc != SC_OK) {
        fprintf(stderr, "singlecarrier_b200: %s failed (%d): %s\n", what, rc, sc_last_error());
        abort();
    }
}

/* ---- fir.h ------------------------------------------------------------------------------------- */
void fir(complex float memory[], bool choice, complex float sample[], int length) {
    if (length <= 0) return;
    must(scl_fir((float *) memory, choice ? 1 : 0, (float *) sample, length), "fir");
}

/* ---- kalman.h / equalizer.h ---------------------------------------------------------------------- */
static void eq_call(int op, complex float in[], int index, float ref, float *ret, int *dibit) {
    float x[2 * EQ_LENGTH] = { 0 };
    if (in != NULL) memcpy(x, &in[index], sizeof x);
    must(scl_eq_op(op, x, ref, (float *) eq_coeff, (float *) kalman_gain, &kalman_y, ret, dibit), "equalizer");
}

void kalman_init(void) { eq_call(4, NULL, 0, 0.0f, NULL, NULL); }
void kalman_reset(void) { eq_call(0, NULL, 0, 0.0f, NULL, NULL); }
void kalman_calculate(complex float x[], int index) { eq_call(1, x, index, 0.0f, NULL, NULL); }

float train_eq(complex float in[], int index, float ref) {
    float ret = 0.0f;
    eq_call(2, in, index, ref, &ret, NULL);
    return ret;
}

float data_eq(uint8_t *bits, complex float in[], int index) {
    float ret = 0.0f;
    int dibit = 0;
    eq_call(3, in, index, 0.0f, &ret, &dibit);      /* slices, updates, descrambles with the rx register */
    *bits = (uint8_t) dibit;
    return ret;
}

/* ---- scramble.h ---------------------------------------------------------------------------------- */
void scramble_init(SRegister sr) { must(scl_scramble_init((int) sr), "scramble_init"); }

int scramble(uint8_t *bits, SRegister sr) {
    if (sr == both) return -1;                       /* src/scramble.c:79-81 */
    must(scl_scramble(bits, (int) sr), "scramble");
    return 0;
}

/* ---- qpsk_internal.h ----------------------------------------------------------------------------- */
float cnormf(complex float val) {
    float out[2];
    must(scl_misc(0, crealf(val), cimagf(val), out), "cnormf");
    return out[0];
}

complex float qpsk_mod(uint8_t bits[], int index) {
    float out[2];
    must(scl_misc(1, (float) (bits[index + 1] == 1), (float) (bits[index] == 1), out), "qpsk_mod");
    return out[0] + out[1] * I;
}

void qpsk_demod(uint8_t bits[], complex float symbol) {
    float out[2];
    must(scl_misc(2, crealf(symbol), cimagf(symbol), out), "qpsk_demod");
    bits[0] = (uint8_t) out[0];
    bits[1] = (uint8_t) out[1];
}

int qpsk_rx_frame(int16_t in[], uint8_t bits[]) {
    int valid = scl_rx_frame(in, bits, (float *) eq_coeff, (float *) kalman_gain, &kalman_y);
    if (valid < 0) must(valid, "qpsk_rx_frame");
    preamble_frames_detected++;
    return valid;
}

int qpsk_tx_frame(int16_t samples[], complex float symbol[], int length, bool preamble) {
    if (length <= 0) return 0;
    must(scl_tx_frame(samples, (const float *) symbol, length, preamble ? 1 : 0), "qpsk_tx_frame");
    return length * CYCLES;
}

/* ---- fft.h: configuration objects keep the reference's public layout ------------------------------ */

/* bytes of a struct fft_state for n points: the struct ends in twiddles[1], so n - 1 more follow it */
static size_t fft_state_bytes(int n) { return sizeof (struct fft_state) + sizeof (complex float) * (size_t) (n - 1); }

/* The reference's placement protocol (src/fft.c:57-64, 98-106): lenmem == NULL -> malloc; otherwise use the
 * caller's block if it is large enough, and always report the size needed through *lenmem. */
static void *place(void *mem, size_t *lenmem, size_t need) {
    if (lenmem == NULL) return malloc(need);
    void *p = (mem != NULL && *lenmem >= need) ? mem : NULL;
    *lenmem = need;
    return p;
}

static void fft_state_fill(fft_cfg st, int nfft, int inverse_fft) {
    st->nfft = nfft;
    st->inverse = inverse_fft;
    scl_twiddles(nfft, inverse_fft, (float *) st->twiddles);     /* same libm expressions as src/fft.c:70-77 */
    scl_kf_factor(nfft, st->factors);                            /* kf_factor(), src/fft.c:433-459 */
}

fft_cfg fft_alloc(int nfft, int inverse_fft, void *mem, size_t *lenmem) {
    fft_cfg st = (fft_cfg) place(mem, lenmem, fft_state_bytes(nfft));
    if (st != NULL) fft_state_fill(st, nfft, inverse_fft);
    return st;
}

void fft(fft_cfg cfg, const complex float *fin, complex float *fout) {
    must(scl_fft(cfg->nfft, cfg->inverse, cfg->factors, (const float *) cfg->twiddles, (const float *) fin,
                 (float *) fout), "fft");
}

/* One block: [struct fftr_state][fft_state for nfft/2 points][tmpbuf: nfft/2][super_twiddles: nfft/4], the layout
 * callers of the reference may rely on (headers/fft.h:33-41, src/fft.c:85-131). */
fftr_cfg fftr_alloc(int nfft, int inverse_fft, void *mem, size_t *lenmem) {
    if (nfft & 1) return NULL;                       /* src/fft.c:89-91 */
    const int half = nfft / 2;
    const size_t sub = fft_state_bytes(half);
    const size_t need = sizeof (struct fftr_state) + sub + sizeof (complex float) * (size_t) (half * 3 / 2);
    fftr_cfg st = (fftr_cfg) place(mem, lenmem, need);
    if (st == NULL) return NULL;
    char *base = (char *) (st + 1);
    st->substate = (fft_cfg) base;
    st->tmpbuf = (complex float *) (base + sub);
    st->super_twiddles = st->tmpbuf + half;
    fft_state_fill(st->substate, half, inverse_fft);
    scl_super_twiddles(half, inverse_fft, (float *) st->super_twiddles);
    return st;
}

void encode_fftr(fftr_cfg st, const float *timedata, complex float *freqdata) {
    must(scl_fftr(st->substate->nfft, st->substate->inverse, 1, st->substate->factors,
                  (const float *) st->substate->twiddles, (const float *) st->super_twiddles, timedata,
                  (float *) freqdata), "fftr");
}

void encode_fftri(fftr_cfg st, const complex float *freqdata, float *timedata) {
    must(scl_fftr(st->substate->nfft, st->substate->inverse, 2, st->substate->factors,
                  (const float *) st->substate->twiddles, (const float *) st->super_twiddles,
                  (const float *) freqdata, timedata), "fftri");
}

/* fft.h:51-52 declares fftr/fftri, src/fft.c defines encode_fftr/encode_fftri: export both */
void fftr(fftr_cfg st, const float *timedata, complex float *freqdata) { encode_fftr(st, timedata, freqdata); }
void fftri(fftr_cfg st, const complex float *freqdata, float *timedata) { encode_fftri(st, freqdata, timedata); }
