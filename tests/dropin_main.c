/*
 * dropin_main.c -- a caller written against the REFERENCE's headers (fir.h, equalizer.h, kalman.h,
 * scramble.h, fft.h, qpsk_internal.h), compiled unchanged against include/sc_compat/ and linked
 * with libsinglecarrier_b200.so.  It does what the reference's main() does (src/qpsk.c:346-464):
 * synthesises packets with qpsk_tx_frame(), then feeds a sample file through qpsk_rx_frame(),
 * printing one line per call, plus a few L1 calls.  tests/test_dropin_gpu.py compares the output
 * with the committed reference vectors.
 */
#include "qpsk_internal.h"
#include "equalizer.h"
#include "kalman.h"
#include "scramble.h"
#include "fir.h"
#include "fft.h"

extern const int8_t preamblevalues[];

int main(int argc, char **argv) {
    if (argc < 2) return 2;
    complex float preambletable[PREAMBLE_LENGTH];
    for (int i = 0; i < PREAMBLE_LENGTH; i++) {
        float v = (float) preamblevalues[i];
        preambletable[i] = v + v * I;
    }
    kalman_init();
    scramble_init(both);

    /* TX: one preamble + one data frame of alternating bits, like preamble_modulate()/qpsk_modulate() */
    int16_t pre[PREAMBLE_SIZE], frame[DATA_SYMBOLS * 5];
    int n = qpsk_tx_frame(pre, preambletable, PREAMBLE_LENGTH, true);
    printf("TXP %d", n);
    for (int i = 0; i < 24; i++) printf(" %d", pre[i]);
    printf("\n");
    uint8_t obits[DATA_SYMBOLS * 2];
    complex float sym[DATA_SYMBOLS];
    for (int i = 0; i < DATA_SYMBOLS * 2; i++) obits[i] = (uint8_t) ((i * 7 + 3) % 5 < 2);
    for (int i = 0; i < DATA_SYMBOLS; i++) sym[i] = qpsk_mod(obits, 2 * i);
    n = qpsk_tx_frame(frame, sym, DATA_SYMBOLS, false);
    long acc = 0;
    for (int i = 0; i < n; i++) acc = acc * 31 + frame[i];
    printf("TXD %d %ld\n", n, acc);

    /* RX: the reference's read loop */
    FILE *fin = fopen(argv[1], "rb");
    if (!fin) return 3;
    int16_t in[FRAME_SIZE];
    uint8_t ibits[BITS_PER_FRAME];
    scramble_init(rx);
    int call = 0;
    while (fread(in, sizeof (int16_t), FRAME_SIZE, fin) == FRAME_SIZE) {
        int valid = qpsk_rx_frame(in, ibits);
        printf("RX %d %d ", call++, valid);
        if (valid) for (int i = 0; i < 62; i++) printf("%d", ibits[i]);
        printf(" %.9g %.9g\n", crealf(eq_coeff[0]), cimagf(eq_coeff[4]));
    }
    fclose(fin);

    /* L1: scrambler keystream, cnormf, demod, fir */
    scramble_init(rx);
    printf("KS ");
    for (int i = 0; i < 31; i++) {
        uint8_t d = 0;
        scramble(&d, rx);
        printf("%d%d", d & 1, d >> 1);
    }
    uint8_t bad = 0;
    printf(" %d\n", scramble(&bad, both));
    uint8_t db[2];
    qpsk_demod(db, -0.5f + 2.0f * I);
    printf("MISC %.9g %d%d\n", cnormf(3.0f - 4.0f * I), db[0], db[1]);
    complex float mem[NTAPS] = { 0 }, x[8] = { 1, 0, 0, 0, 0, I, 0, 0 };
    fir(mem, false, x, 8);
    printf("FIR %.9g %.9g %.9g\n", crealf(x[0]), crealf(x[4]), cimagf(x[7]));
    fft_cfg cfg = fft_alloc(16, 0, NULL, NULL);
    complex float fi[16], fo[16];
    for (int i = 0; i < 16; i++) fi[i] = (float) (i % 3) - 0.5f * I * (float) (i % 5);
    fft(cfg, fi, fo);
    printf("FFT %d %.9g %.9g\n", cfg->nfft, crealf(fo[1]), cimagf(fo[7]));
    free(cfg);
    return 0;
}
