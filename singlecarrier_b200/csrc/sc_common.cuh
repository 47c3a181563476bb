// sc_common.cuh -- constants, tables and layouts shared by the sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/singlecarrier_b200.h"
#include "sc_exact.cuh"

namespace sc {

constexpr int FRAME = SC_FRAME_SIZE;          // 1880 samples per call
constexpr int CYC = 5;                        // samples per symbol (CYCLES)
constexpr int NTAPS = SC_NTAPS;               // 49
constexpr int PRE = SC_PREAMBLE_LENGTH;       // 128
constexpr int NDATA = SC_DATA_SYMBOLS;        // 31
constexpr int EQ = SC_EQ_LENGTH;              // 5
constexpr int WIN = 290;                      // symbols of the older half that a call can read:
                                              // lag<=127 + 128 preamble + 31 data + 4 taps
constexpr float FIR_GAIN = 2.2f;              // fir.h:17
constexpr int MATCH_THRESHOLD = PRE - 30;     // qpsk.c:196: matches > 98

// Tracker input window handed from the front-end kernel to the tracking kernel, per stream:
//   rows 0..162   X[r] = W[max_index + r]   (128 training steps + 31 data steps + 4 taps)
//   rows 163..197 Y[r] = W[rx_timing + r]   (the 31 data steps of an invalid call + 4 taps)
// stored in tiles of 32 streams: win[(tile*WIN_ROWS + row)*32 + (s & 31)], so a warp of the
// tracking kernel (32 consecutive streams) reads one contiguous 256-byte row per step.
constexpr int X_ROWS = PRE + NDATA + EQ - 1;  // 163
constexpr int Y_ROWS = NDATA + EQ - 1;        // 35
constexpr int WIN_ROWS = X_ROWS + Y_ROWS;     // 198
constexpr int WIN_ROWS_OV = X_ROWS + (PRE - 1) + Y_ROWS;   // 325: overlapped chains keep W[128..289] instead of 35 chosen rows
constexpr int TRK_STATE_WORDS = 64;           // floats per stream handed from the training to the data kernel

}  // namespace sc
