/*
 * batched_demo.c -- the batched C ABI from plain C: demodulate K copies of a raw int16 sample file as
 * one bank of K streams (stream k is the file delayed by k samples), print every valid call.
 *
 *   gcc -std=gnu11 -O2 -I include examples/batched_demo.c -L singlecarrier_b200 -lsinglecarrier_b200 \
 *       -Wl,-rpath,$PWD/singlecarrier_b200 -o batched_demo && ./batched_demo tests/golden/preamble_qpsk_8k.raw 4
 *
 * This is what replaces the reference's while(1){ fread; qpsk_rx_frame(); fwrite } loop
 * (src/qpsk.c:436-458) when there are many streams.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "singlecarrier_b200.h"

int main(int argc, char **argv) {
    if (argc < 2) {
        fprintf(stderr, "usage: %s samples.raw [n_streams]\n", argv[0]);
        return 2;
    }
    const int n_streams = argc > 2 ? atoi(argv[2]) : 4;
    FILE *f = fopen(argv[1], "rb");
    if (!f) return 3;
    fseek(f, 0, SEEK_END);
    const long n_samples = ftell(f) / 2;
    fseek(f, 0, SEEK_SET);
    int16_t *file = malloc((size_t) n_samples * 2);
    if (fread(file, 2, (size_t) n_samples, f) != (size_t) n_samples) return 3;
    fclose(f);

    const int n_frames = (int) ((n_samples + n_streams) / SC_FRAME_SIZE) + 2;      /* + 2 calls to flush the pipeline */
    const int64_t stride = (int64_t) n_frames * SC_FRAME_SIZE;
    int16_t *in = calloc((size_t) (n_streams * stride), 2);
    for (int k = 0; k < n_streams; k++) memcpy(in + k * stride + k, file, (size_t) n_samples * 2);

    sc_modem *m = NULL;
    if (sc_create(&m, 0, n_streams, 0, 0.0f) != SC_OK) {
        fprintf(stderr, "sc_create: %s\n", sc_last_error());
        return 1;
    }
    sc_frame_result *res = calloc((size_t) n_streams * n_frames, sizeof *res);
    if (sc_rx_frames_host(m, in, stride, n_frames, res, n_frames, NULL) != SC_OK) {
        fprintf(stderr, "sc_rx_frames_host: %s\n", sc_last_error());
        return 1;
    }
    uint8_t *rows = malloc((size_t) n_streams * n_frames * SC_BITS_PER_CALL);
    memset(rows, 0xff, (size_t) n_streams * n_frames * SC_BITS_PER_CALL);
    sc_unpack_bits(res, (int64_t) n_streams * n_frames, rows);
    for (int k = 0; k < n_streams; k++)
        for (int n = 0; n < n_frames; n++) {
            const sc_frame_result *r = &res[k * n_frames + n];
            if (!r->valid) continue;
            printf("stream %d call %d matches %d max_index %d max_value %.2f bits ", k, n, r->matches, r->max_index,
                   r->max_value);
            for (int j = 0; j < SC_BITS_PER_CALL; j++) putchar('0' + rows[((size_t) k * n_frames + n) * SC_BITS_PER_CALL + j]);
            putchar('\n');
        }
    printf("%d streams x %d calls, %llu kernel launches\n", n_streams, n_frames, (unsigned long long) sc_launch_count());
    sc_destroy(m);
    free(rows);
    free(res);
    free(in);
    free(file);
    return 0;
}
