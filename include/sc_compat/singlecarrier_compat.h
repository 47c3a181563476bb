/*
 * singlecarrier_compat.h -- the reference's single-stream C interface, served by
 * libsinglecarrier_b200.so.
 *
 * A program written against srsampson/SingleCarrier's headers (fir.h, equalizer.h, kalman.h,
 * scramble.h, fft.h, qpsk_internal.h) compiles against the same-named headers in this directory
 * (each just includes this file) and links with -lsinglecarrier_b200 instead of the reference's
 * objects.  Every symbol below keeps the reference's name, argument meaning and return
 * convention; each cites the reference declaration it stands in for.  Behind them the arithmetic
 * runs in the library's CUDA kernels with a one-stream batch (there is no CPU path; without a
 * CUDA device the calls abort with a message).  For throughput use the batched API in
 * ../singlecarrier_b200.h.
 *
 * Like the reference, these entry points share process-global state and are not thread safe.
 */
#ifndef SINGLECARRIER_COMPAT_H
#define SINGLECARRIER_COMPAT_H

#include <complex.h>
#include <math.h>
#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- modem constants: reference headers/qpsk_internal.h:23-60 ----------------------------- */
#define FINE_TIMING_OFFSET 3
#define TX_FILENAME        "/tmp/spectrum-filtered.raw"
#define RX_FILENAME        "/tmp/databits.txt"
#define EOF_COST_VALUE     5.0f
#define EQ_LENGTH          5
#define FS                 8000.0f
#define RS                 1600.0f
#define TS                 (1.0f / RS)
#define CYCLES             (int) (FS / RS)
#define CYCLESF            5
#define CENTER             1100.0f
#define NS                 8
#define DATA_SYMBOLS       31
#define FRAME_SYMBOLS      (DATA_SYMBOLS * NS)
#define DATA_SAMPLES       (DATA_SYMBOLS * CYCLES * NS)
#define DATA_SIZE          1240
#define FRAME_SIZE         1880
#define BITS_PER_FRAME     496
#define PREAMBLE_LENGTH    128
#define PREAMBLE_SIZE      (PREAMBLE_LENGTH * CYCLESF)
#ifndef M_PI
#define M_PI               3.14159265358979323846f
#endif
#define TAU                (2.0f * M_PI)
#define ROT45              (M_PI / 4.0f)
#define cmplx(float_value)     (cosf(float_value) + sinf(float_value) * I)
#define cmplxconj(float_value) (cosf(float_value) + sinf(float_value) * -I)

typedef enum { hunt, process } RXState;                    /* qpsk_internal.h:71-74 */

/* ---- fir.h:16-19 --------------------------------------------------------------------------- */
#define NTAPS 49
#define GAIN  2.2f
void fir(complex float memory[], bool choice, complex float sample[], int length);

/* ---- kalman.h:26-32 and the data symbols of src/kalman.c:19-21 ------------------------------ */
void kalman_init(void);
void kalman_reset(void);
void kalman_calculate(complex float x[], int index);
extern complex float eq_coeff[EQ_LENGTH];
extern complex float kalman_gain[EQ_LENGTH];
extern float kalman_y;

/* ---- equalizer.h:17-18 ---------------------------------------------------------------------- */
float train_eq(complex float in[], int index, float ref);
float data_eq(uint8_t *bits, complex float in[], int index);

/* ---- scramble.h:16-30 ----------------------------------------------------------------------- */
#define SEED 0x4A80
#define BITS 2
typedef enum { tx, rx, both } SRegister;
void scramble_init(SRegister sr);
int scramble(uint8_t *bits, SRegister sr);                 /* -1 for sr == both */

/* ---- qpsk_internal.h:79-84 ------------------------------------------------------------------ */
float cnormf(complex float val);
complex float qpsk_mod(uint8_t bits[], int index);
void qpsk_demod(uint8_t bits[], complex float symbol);
int qpsk_rx_frame(int16_t in[], uint8_t bits[]);           /* 1 = valid frame, bits[0..61] written */
int qpsk_tx_frame(int16_t samples[], complex float symbol[], int length, bool preamble);
extern int preamble_frames_detected;                       /* src/qpsk.c:70 (DEBUG2) */

/* ---- tables of src/constants.c -------------------------------------------------------------- */
extern const complex float constellation[4];
extern const int8_t preamblevalues[PREAMBLE_LENGTH];
extern const float alpha50_root[NTAPS];
extern const float alpha35_root[NTAPS];

/* ---- fft.h:24-52 ---------------------------------------------------------------------------- */
struct fft_state {
    int nfft;
    int inverse;
    int factors[64];
    complex float twiddles[1];
};
typedef struct fft_state *fft_cfg;

struct fftr_state {
    fft_cfg substate;
    complex float *tmpbuf;
    complex float *super_twiddles;
};
typedef struct fftr_state *fftr_cfg;

fft_cfg fft_alloc(int nfft, int inverse_fft, void *mem, size_t *lenmem);
void fft(fft_cfg cfg, const complex float *fin, complex float *fout);
fftr_cfg fftr_alloc(int nfft, int inverse_fft, void *mem, size_t *lenmem);
void fftr(fftr_cfg cfg, const float *timedata, complex float *freqdata);      /* declared in fft.h:51 */
void fftri(fftr_cfg cfg, const complex float *freqdata, float *timedata);     /* declared in fft.h:52 */
void encode_fftr(fftr_cfg cfg, const float *timedata, complex float *freqdata);   /* defined in src/fft.c:139 */
void encode_fftri(fftr_cfg cfg, const complex float *freqdata, float *timedata);  /* defined in src/fft.c:166 */

#ifdef __cplusplus
}
#endif
#endif
