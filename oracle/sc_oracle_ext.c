/*
 * sc_oracle_ext.c -- TEST INFRASTRUCTURE ONLY.  Specification (as code) of the PACKET-MODE EXTENSION.
 *
 * PARITY UNPINNED: this part has NO counterpart in srsampson/SingleCarrier that could serve as an oracle.
 * The reference stops short of a full receiver in three places, each marked TODO or commented out:
 *   /root/reference/src/qpsk.c:206-215  only DATA_SYMBOLS = 31 of the packet's NS x 31 = 248 data symbols are
 *                                       decoded (headers/qpsk_internal.h:38-40),
 *   /root/reference/src/qpsk.c:386,397  "//scramble_init(tx);" and "// scramble(&sdata, tx); TODO": the
 *                                       transmitter never scrambles although the receiver descrambles,
 *   /root/reference/src/qpsk.c:161      symbols 325.. of a decimated window are taken from the UNFILTERED half of
 *                                       input_frame (5 i + rx_timing runs past 1879), which is harmless only
 *                                       because nothing beyond index 289 is ever read.
 * The extension finishes exactly these, re-using the reference's own arithmetic (train_eq, data_eq, the Kalman
 * recursion, fir, the mixer -- all through the pinned restatement in sc_oracle.c) and changes NOTHING of what the
 * reference already computes: every field of the ordinary per-call result stays as sc_oracle.c produces it.
 *
 * Definition.  Let filt[] be the stream's matched-filter output as one continuous sequence (the reference's
 * rx_filter memory persists across calls, so it is one), frame f occupying filt[1880 f .. 1880 f + 1879].
 * Call n searches the window decimated from frame n-2 with the rx_timing T that was in force when that window
 * was taken (the value at entry of call n-1).  Whenever call n is VALID (matches > 98, qpsk.c:196) and n >= 2:
 *   S[j] = filt[1880 (n-2) + 5 (max_index + j) + T],  j = 0 .. 128 + 248 + 3      (one continuous symbol grid:
 *          the samples beyond frame n-2 come from frame n-1, which call n has already filtered)
 *   kalman_reset(); 128 x train_eq(S, i, preamble[i])                              (as qpsk.c:186-188)
 *   RX register := SEED                                                             (scramble_init per packet, the
 *                                                                                    mirror of qpsk.c:386)
 *   for k = 0 .. 247: data_eq(&dibit, S, 128 + k)                                  (equalizer state carried through
 *                                                                                    all 8 data frames)
 *   bits[2k] = dibit & 1 (Q), bits[2k+1] = dibit >> 1 (I)                          (qpsk.c:211-212)
 * Transmitter: the TX register is re-seeded after every preamble and every data dibit goes through
 * scramble(&sdata, tx) before qpsk_mod (qpsk.c:386,397 un-commented).
 */
#include <stdlib.h>
#include <string.h>

#include "sc_oracle.h"

#define EXT_NS 8
#define EXT_DATA (EXT_NS * SCO_DATA_SYMBOLS)          /* 248 */
#define EXT_SYMS (SCO_PREAMBLE_LENGTH + EXT_DATA + SCO_EQ_LENGTH - 1)   /* 380 */
#define EXT_SEED 0x4A80

typedef struct {
    int32_t call;          /* n */
    int32_t max_index;
    int32_t matches;       /* of the re-run training: equals the ordinary result's */
    int32_t timing;        /* T used for the symbol grid */
    float cost;            /* sum of the 248 data_eq() returns */
    uint8_t bits[2 * EXT_DATA];
} sco_ext_packet;

/*
 * Runs the ordinary receiver over n_frames frames of one stream (identical calls to sco_rx_frame) and, beside
 * it, the packet-mode decode.  stats/bits as sco_run_stream(); packets[] receives up to max_packets entries.
 * Returns the number of packets decoded.
 */
int sco_ext_run_stream(const int16_t in[], int n_frames, int wide, float foffset_hz, uint8_t bits[],
                       sco_frame_stats stats[], sco_ext_packet packets[], int max_packets) {
    sco_state *s = malloc(sizeof *s);
    sco_c32 *filt = calloc((size_t) (n_frames + 1) * SCO_FRAME_SIZE, sizeof *filt);
    int *t_entry = calloc((size_t) n_frames + 1, sizeof *t_entry);
    int n_packets = 0;

    sco_init(s, wide, foffset_hz);
    for (int n = 0; n < n_frames; n++) {
        t_entry[n] = s->rx_timing;                     /* rx_timing at entry of call n */
        sco_frame_stats st;
        uint8_t row[SCO_BITS_PER_CALL];
        memset(row, 255, sizeof row);
        int valid = sco_rx_frame(s, in + (size_t) n * SCO_FRAME_SIZE, row, &st);
        if (bits) memcpy(bits + (size_t) n * SCO_BITS_PER_CALL, row, sizeof row);
        if (stats) stats[n] = st;
        /* call n has just filtered frame n-1 in place (older half of input_frame, qpsk.c:152) */
        if (n >= 1) memcpy(filt + (size_t) (n - 1) * SCO_FRAME_SIZE, s->input_frame, SCO_FRAME_SIZE * sizeof *filt);

        if (valid && n >= 2 && n_packets < max_packets) {
            sco_ext_packet *p = &packets[n_packets++];
            sco_c32 S[EXT_SYMS];
            const int T = t_entry[n - 1];
            for (int j = 0; j < EXT_SYMS; j++)
                S[j] = filt[(size_t) (n - 2) * SCO_FRAME_SIZE + 5 * (st.max_index + j) + T];
            sco_kalman k;
            memset(&k, 0, sizeof k);
            sco_kalman_init(&k);                       /* E, q, and kalman_reset() */
            int matches = 0;
            for (int i = 0; i < SCO_PREAMBLE_LENGTH; i++) {
                float ref = (float) sco_preamblevalues[i];
                if (sco_train_eq(&k, S, i, ref) * ref > 0.0f) matches++;
            }
            uint16_t lfsr = EXT_SEED;
            float cost = 0.0f;
            for (int i = 0; i < EXT_DATA; i++) {
                uint8_t dibit;
                cost = cost + sco_data_eq(&k, &lfsr, &dibit, S, SCO_PREAMBLE_LENGTH + i);
                p->bits[2 * i + 1] = dibit >> 1;
                p->bits[2 * i] = dibit & 0x1;
            }
            p->call = n;
            p->max_index = st.max_index;
            p->matches = matches;
            p->timing = T;
            p->cost = cost;
        }
    }
    free(t_entry);
    free(filt);
    free(s);
    return n_packets;
}

/* Transmitter with the two commented-out scrambler lines enabled: one packet = preamble + 8 scrambled frames. */
int sco_ext_tx_packet(sco_state *s, int16_t samples[SCO_FRAME_SIZE], const uint8_t bits[2 * EXT_DATA]) {
    int n = sco_tx_preamble(s, samples);
    uint16_t lfsr = EXT_SEED;                          /* scramble_init(tx), qpsk.c:386 */
    for (int f = 0; f < EXT_NS; f++) {
        uint8_t obits[SCO_BITS_PER_CALL];
        for (int i = 0; i < SCO_DATA_SYMBOLS; i++) {
            const uint8_t *b = bits + (size_t) f * SCO_BITS_PER_CALL + 2 * i;
            uint8_t sdata = (uint8_t) ((b[1] << 1) | b[0]);
            sco_scramble2(&sdata, &lfsr);              /* scramble(&sdata, tx), qpsk.c:397 */
            obits[2 * i + 1] = (sdata >> 1) & 0x1;
            obits[2 * i] = sdata & 0x1;
        }
        n += sco_tx_data(s, samples + n, obits, SCO_DATA_SYMBOLS);
    }
    return n;
}

unsigned long sco_ext_sizeof_packet(void) { return sizeof(sco_ext_packet); }
