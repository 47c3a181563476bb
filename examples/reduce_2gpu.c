/*
 * reduce_2gpu.c -- the multi-GPU form of the path from plain C (no CUDA headers, no Python): the streams of one
 * job are sharded over the GPUs of the box (contiguous blocks, no data-path traffic between them), every GPU
 * demodulates its shard and counts lock / bit statistics on the device, and ONE collective -- an all-reduce of the
 * 16 counters over NCCL -- gives every rank the totals.
 *
 *   gcc -std=gnu11 -O2 -I include examples/reduce_2gpu.c -L singlecarrier_b200 -lsinglecarrier_b200 \
 *       -Wl,-rpath,$PWD/singlecarrier_b200 -o reduce_2gpu && ./reduce_2gpu tests/golden/preamble_qpsk_8k.raw 8
 *
 * Uses min(2, visible GPUs) devices; with one GPU the communicator has a single rank and the reduce is the identity.
 * One process drives both ranks here (sc_comm_init_all + group start/end); a process per GPU would use
 * sc_comm_unique_id / sc_comm_init_rank instead, as bench.py does.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "singlecarrier_b200.h"

#define CHECK(call)                                                                 \
    do {                                                                            \
        int rc_ = (call);                                                           \
        if (rc_ != SC_OK) {                                                         \
            fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, sc_last_error());   \
            return 1;                                                               \
        }                                                                           \
    } while (0)

int main(int argc, char **argv) {
    if (argc < 2) {
        fprintf(stderr, "usage: %s samples.raw [n_streams]\n", argv[0]);
        return 2;
    }
    const int n_streams = argc > 2 ? atoi(argv[2]) : 8;
    FILE *f = fopen(argv[1], "rb");
    if (!f) return 3;
    fseek(f, 0, SEEK_END);
    const long n_samples = ftell(f) / 2;
    fseek(f, 0, SEEK_SET);
    int16_t *file = malloc((size_t) n_samples * 2);
    if (fread(file, 2, (size_t) n_samples, f) != (size_t) n_samples) return 3;
    fclose(f);

    /* stream k = the file delayed by 3k samples */
    const int n_frames = (int) ((n_samples + 3 * n_streams) / SC_FRAME_SIZE) + 2;
    const int64_t stride = (int64_t) n_frames * SC_FRAME_SIZE;
    int16_t *in = calloc((size_t) (n_streams * stride), 2);
    for (int k = 0; k < n_streams; k++) memcpy(in + k * stride + 3 * k, file, (size_t) n_samples * 2);

    int n_dev = sc_device_count();
    if (n_dev < 1) {
        fprintf(stderr, "no CUDA device (this library has no CPU path)\n");
        return 1;
    }
    if (n_dev > 2) n_dev = 2;
    if (n_dev > n_streams) n_dev = n_streams;

    void *comms[2] = { NULL, NULL };
    const int devs[2] = { 0, 1 };
    CHECK(sc_comm_init_all(comms, n_dev, devs));

    sc_modem *bank[2] = { NULL, NULL };
    void *d_in[2], *d_res[2], *d_cnt[2];
    int lo[3];
    for (int r = 0; r <= n_dev; r++) lo[r] = (int) ((long) n_streams * r / n_dev);      /* contiguous shards */
    for (int r = 0; r < n_dev; r++) {
        const int ns = lo[r + 1] - lo[r];
        CHECK(sc_create(&bank[r], devs[r], ns, 0, 0.0f));
        CHECK(sc_device_malloc(devs[r], (size_t) ns * stride * 2, &d_in[r]));
        CHECK(sc_device_malloc(devs[r], (size_t) ns * n_frames * sizeof(sc_frame_result), &d_res[r]));
        CHECK(sc_device_malloc(devs[r], SC_N_COUNTERS * sizeof(uint64_t), &d_cnt[r]));
        CHECK(sc_device_copy(devs[r], d_in[r], in + (size_t) lo[r] * stride, (size_t) ns * stride * 2, SC_COPY_H2D));
        CHECK(sc_rx_frames_dev(bank[r], d_in[r], stride, n_frames, d_res[r], n_frames, NULL, NULL));
        CHECK(sc_lock_stats_dev(devs[r], d_res[r], ns, n_frames, n_frames, d_cnt[r], NULL));
    }
    /* the path's only collective */
    CHECK(sc_comm_group_start());
    for (int r = 0; r < n_dev; r++) CHECK(sc_reduce_stats(d_cnt[r], SC_N_COUNTERS, comms[r], NULL));
    CHECK(sc_comm_group_end());

    for (int r = 0; r < n_dev; r++) {
        uint64_t c[SC_N_COUNTERS];
        CHECK(sc_device_synchronize(devs[r]));
        CHECK(sc_device_copy(devs[r], c, d_cnt[r], sizeof c, SC_COPY_D2H));
        printf("rank %d of %d (streams %d..%d): calls %llu valid %llu sum_matches %llu bit_popcount %llu\n", r, n_dev, lo[r],
               lo[r + 1] - 1, (unsigned long long) c[0], (unsigned long long) c[1], (unsigned long long) c[2],
               (unsigned long long) c[5]);
    }
    for (int r = 0; r < n_dev; r++) {
        sc_comm_destroy(comms[r]);
        sc_device_free(devs[r], d_in[r]);
        sc_device_free(devs[r], d_res[r]);
        sc_device_free(devs[r], d_cnt[r]);
        sc_destroy(bank[r]);
    }
    free(in);
    free(file);
    return 0;
}
