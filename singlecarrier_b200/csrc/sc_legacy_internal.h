/* sc_legacy_internal.h -- private bridge between the C99 drop-in shim (sc_legacy.c) and the CUDA
 * side (sc_legacy_dev.cu).  Complex values cross as float pairs.  All return SC_OK or SC_E*. */
#ifndef SC_LEGACY_INTERNAL_H
#define SC_LEGACY_INTERNAL_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
int scl_fir(float *memory, int wide, float *sample, int length);
/* op: 0 kalman_reset, 1 kalman_calculate, 2 train_eq, 3 data_eq, 4 kalman_init */
int scl_eq_op(int op, const float *x5, float ref, float *eq_coeff, float *gain, float *ky, float *ret, int *dibit);
int scl_scramble_init(int sr);
int scl_scramble(uint8_t *v, int sr);
/* op 0: cnormf(a + bi) -> out[0]; 1: qpsk_mod(bitI=a, bitQ=b) -> out[0..1]; 2: qpsk_demod(a + bi) -> out[0]=Q bit, out[1]=I bit */
int scl_misc(int op, float a, float b, float *out);
int scl_rx_frame(const int16_t *in, uint8_t *bits, float *eq_coeff, float *gain, float *ky);
int scl_tx_frame(int16_t *samples, const float *symbols, int length, int preamble);
int scl_fft(int nfft, int inverse, const int *factors, const float *twiddles, const float *in, float *out);
int scl_fftr(int ncfft, int inverse, int mode, const int *factors, const float *twiddles, const float *super_twiddles,
             const float *in, float *out);
void scl_kf_factor(int n, int *facbuf);
void scl_twiddles(int n, int inverse, float *tw);
void scl_super_twiddles(int ncfft, int inverse, float *tw);
#ifdef __cplusplus
}
#endif
#endif
