// sc_fft256.cuh -- a 256-point complex FFT held in the registers of one warp (8 points per lane), radix-4
// butterflies in kiss-fft's order, lane <-> register index exchanges by __shfl_xor_sync, no shared memory.
// Used by fft256_warp_kernel (fft.h, sc_fft_kernels.cu) and by the FFT-proposed preamble search
// (sc_stage_kernels.cu).
#pragma once
#include "sc_common.cuh"

namespace sc {

__device__ __forceinline__ c32 csub(c32 a, c32 b) { return mk(__fsub_rn(a.r, b.r), __fsub_rn(a.i, b.i)); }

// ------------------------------------------------------------------------------------------------
// n = 256 = 4 x 4 x 4 x 4 (the size a 128-lag x 128-tap correlation needs): one WARP per
// transform, 8 points per lane in registers, no shared memory.  kf_factor gives (4,64)(4,16)(4,4)(4,1), so
// working-array index i = 64 d0 + 16 d1 + 4 d2 + d3 holds input element d0 + 4 d1 + 16 d2 + 64 d3 and the
// stages run over d3, d2, d1, d0 in that order (the recursion unwinds innermost first).  A radix-4 butterfly
// needs its digit in the register index; between stages one or two index bits are exchanged between lane
// and register position with __shfl_xor_sync (5 exchanges of 4 complex values per lane in all):
//
//   position     L0    L1    L2    L3    L4  | R0    R1    R2        (L = lane bit, R = register-index bit)
//   load         d0lo  d0hi  d1lo  d1hi  d2lo| d2hi  d3lo  d3hi      coalesced: element t + 32 r
//   stage d3                                                         legs r = b + 2k          (b = R0)
//   L4<->R2      d0lo  d0hi  d1lo  d1hi  d3hi| d2hi  d3lo  d2lo
//   stage d2                                                         legs r = (k>>1) + 2b + 4(k&1)   (b = R1)
//   L2<->R2, L3<->R0
//                d0lo  d0hi  d2lo  d2hi  d3hi| d1hi  d3lo  d1lo
//   stage d1                                                         same leg mapping
//   L0<->R2, L1<->R1
//                d1lo  d3lo  d2lo  d2hi  d3hi| d1hi  d0hi  d0lo
//   stage d0                                                         legs r = b + 2(k>>1) + 4(k&1)   (b = R0)
//   store        i = 64 d0 + 16 d1 + 4 d2 + d3: for a fixed register the 32 lanes cover 32 consecutive
//                outputs (in permuted lane order), so every store instruction writes one 256-byte segment.
//
// Every butterfly is kf_bfly4's expression sequence (src/fft.c:218-266) including the multiplications by
// twiddle 0, with the host-made twiddle table, so the output is bit-identical to src/fft.c.
// ------------------------------------------------------------------------------------------------
template <bool INVERSE>
__device__ __forceinline__ void bfly4_reg(c32 &a0, c32 &a1, c32 &a2, c32 &a3, c32 t1, c32 t2, c32 t3) {
    const c32 s0 = cmul(a1, t1), s1 = cmul(a2, t2), s2 = cmul(a3, t3);
    const c32 s5 = csub(a0, s1);
    const c32 f0 = cadd(a0, s1);
    const c32 s3 = cadd(s0, s2), s4 = csub(s0, s2);
    a2 = csub(f0, s3);
    a0 = cadd(f0, s3);
    if (INVERSE) {
        a1 = mk(__fsub_rn(s5.r, s4.i), __fadd_rn(s5.i, s4.r));
        a3 = mk(__fadd_rn(s5.r, s4.i), __fsub_rn(s5.i, s4.r));
    } else {
        a1 = mk(__fadd_rn(s5.r, s4.i), __fsub_rn(s5.i, s4.r));
        a3 = mk(__fsub_rn(s5.r, s4.i), __fadd_rn(s5.i, s4.r));
    }
}

// exchange lane bit LANE_MASK with register-index bit REG_BIT of an 8-value-per-lane array
template <int LANE_MASK, int REG_BIT>
__device__ __forceinline__ void xchg_bit(c32 (&v)[8], int lane) {
    const bool up = (lane & LANE_MASK) != 0;
#pragma unroll
    for (int r = 0; r < 8; r++) {
        if (r & (1 << REG_BIT)) continue;
        const int r1 = r | (1 << REG_BIT);
        const float sr = up ? v[r].r : v[r1].r, si = up ? v[r].i : v[r1].i;
        const float gr = __shfl_xor_sync(0xffffffffu, sr, LANE_MASK), gi = __shfl_xor_sync(0xffffffffu, si, LANE_MASK);
        if (up) v[r] = mk(gr, gi);
        else v[r1] = mk(gr, gi);
    }
}

// the twiddles one lane needs (they depend on the lane only):
// stage d2: u = d3 = R1 + 2 L4, tw[16 u k]; stage d1: u = 4 d2 + d3 = 4(L2 + 2 L3) + R1 + 2 L4, tw[4 u k];
// stage d0: u = 16 d1 + 4 d2 + d3 = 16(L0 + 2 R0) + 4(L2 + 2 L3) + L1 + 2 L4, tw[u k]
struct Fft256Twiddles {
    c32 tw0, t2[2][3], t1[2][3], t0[2][3];
    __device__ __forceinline__ void load(const c32 *__restrict__ tw, int lane) {
        const int L0 = lane & 1, L1 = (lane >> 1) & 1, L2 = (lane >> 2) & 1, L3 = (lane >> 3) & 1, L4 = (lane >> 4) & 1;
        tw0 = tw[0];
#pragma unroll
        for (int b = 0; b < 2; b++) {
            const int u2 = b + 2 * L4, u1 = 4 * (L2 + 2 * L3) + b + 2 * L4, u0 = 16 * (L0 + 2 * b) + 4 * (L2 + 2 * L3) + L1 + 2 * L4;
#pragma unroll
            for (int k = 1; k < 4; k++) {
                t2[b][k - 1] = tw[16 * u2 * k];
                t1[b][k - 1] = tw[4 * u1 * k];
                t0[b][k - 1] = tw[u0 * k];
            }
        }
    }
};

// in: v[r] = x[lane + 32 r]; out: v[r] = X[fft256_out_index(lane, r)]
template <bool INVERSE>
__device__ __forceinline__ void fft256_regs(c32 (&v)[8], int lane, const Fft256Twiddles &T) {
    // stage d3 (m = 1, every twiddle is tw[0])
#pragma unroll
    for (int q = 0; q < 2; q++) bfly4_reg<INVERSE>(v[q], v[q + 2], v[q + 4], v[q + 6], T.tw0, T.tw0, T.tw0);
    xchg_bit<16, 2>(v, lane);
    // stage d2 (m = 4)
#pragma unroll
    for (int q = 0; q < 2; q++) bfly4_reg<INVERSE>(v[2 * q], v[2 * q + 4], v[2 * q + 1], v[2 * q + 5], T.t2[q][0], T.t2[q][1], T.t2[q][2]);
    xchg_bit<4, 2>(v, lane);
    xchg_bit<8, 0>(v, lane);
    // stage d1 (m = 16)
#pragma unroll
    for (int q = 0; q < 2; q++) bfly4_reg<INVERSE>(v[2 * q], v[2 * q + 4], v[2 * q + 1], v[2 * q + 5], T.t1[q][0], T.t1[q][1], T.t1[q][2]);
    xchg_bit<1, 2>(v, lane);
    xchg_bit<2, 1>(v, lane);
    // stage d0 (m = 64)
#pragma unroll
    for (int q = 0; q < 2; q++) bfly4_reg<INVERSE>(v[q], v[q + 4], v[q + 2], v[q + 6], T.t0[q][0], T.t0[q][1], T.t0[q][2]);
}

// natural-order index of register r of lane after fft256_regs():
// i = 64 (R2 + 2 R1) + 32 R0 + 16 L0 + 4 L2 + 8 L3 + L1 + 2 L4
__host__ __device__ __forceinline__ int fft256_out_index(int lane, int r) {
    const int L0 = lane & 1, L1 = (lane >> 1) & 1, L2 = (lane >> 2) & 1, L3 = (lane >> 3) & 1, L4 = (lane >> 4) & 1;
    return 64 * ((r >> 2) + 2 * ((r >> 1) & 1)) + 32 * (r & 1) + 16 * L0 + L1 + 4 * L2 + 8 * L3 + 2 * L4;
}

}  // namespace sc
