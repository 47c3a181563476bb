// sc_tracker_coop.cuh -- the per-stream sequential loop of sc_tracker.cuh spread over 16 lanes.
//
// One thread per stream (sc_tracker.cuh) is the right shape when there are enough streams to fill the GPU: the loop
// is a 63,000-operation dependent chain per call and a lone warp per scheduler runs it at ~0.45 instructions per
// clock (940 clocks per step).  Small banks (fewer streams than the GPU has lanes) are bound by exactly that latency,
// so here one stream's step is cut along the data flow of src/kalman.c:85-141 instead:
//
//   lanes 0..9   own one element U(i,j), i < j, of the unit upper triangle (index j(j-1)/2 + i, as kalman.c:25 is
//                flattened in sc_tracker.cuh) and a copy of d[j];
//   lanes 10..14 own column j = lane - 10: d[j], eq_coeff[j], and the final kalman_gain[j];
//   lane 15      holds -0 in every register the others gather from: x + (-0) == x for every x (including -0) in
//                round-to-nearest, so a sum with fewer terms than the longest one reads its missing terms from
//                here and no lane needs a conditional add.
//
// Every value is produced by the same operations in the same order as in the one-thread version (and therefore as in
// the reference): the products of a sum are formed where their operands live, gathered with shuffles, and added in the
// reference's order by every lane that needs the sum.  What changes is only which lane executes an operation, so the
// results are bit-identical by construction; tests/test_round2_gpu.py compares the two kernels on whole banks.
//
//   6.2   F[j] = conj(x[j]) + sum_{i<j} U(i,j) conj(x[i])        product on lane (i,j), summed by every lane of column j
//   6.4   G[j] = F[j] d[j]                                        every lane of column j
//   6.5-6 A[j] = A[j-1] + Re(G[j] conj(F[j]))                     terms gathered from lanes 10..14, prefix on every lane
//   6.7.. ht, the denominators and their reciprocals              every lane, for its own column and the one before
//   6.13  d[j] *= hq (A[j-1] + ht) / (A[j] + ht)                  every lane of column j
//   6.15  U(i,j) += H[j] conj(G[i] as it stands before column j)  lane (i,j); G[i]'s running value = G[i] + sum of the
//   6.16  G[i] += G[j] conj(U(i,j))                               products P(i,j') = G[j'] conj(U(i,j')) of the lanes
//                                                                  (i,j'), i < j' < j, gathered and added in order
//   update_eq: eq_coeff[i] += (err kalman_y) conj(G[i])           lanes 10..14 (G[i] final = all of row i's products)
//   train_eq / data_eq: sum_i x[i] eq_coeff[i]                    products on lanes 10..14, summed by every lane
//
// 33 shuffles and ~130 arithmetic instructions per lane and step instead of ~400 in one thread.
#pragma once
#include "sc_common.cuh"

namespace sc {

constexpr int TC_LANES = 16;

__device__ __forceinline__ c32 shfl_c(c32 v, int src) {
    return mk(__shfl_sync(0xffffffffu, v.r, src), __shfl_sync(0xffffffffu, v.i, src));
}

__device__ __forceinline__ float bits_sel(float v, uint32_t keep, uint32_t other) {   // (v & keep) | other
    return __uint_as_float((__float_as_uint(v) & keep) | other);
}

struct CoopTracker {
    // role of this lane
    int base;            // first lane of the 16-lane group
    int myi, myj;        // element (myi, myj); column lanes have myi == myj
    int srcF[4], srcP[4], srcG;
    uint32_t keep, neg0; // lane 15: keep = 0, neg0 = sign bit; every other lane: keep = ~0, neg0 = 0
    uint32_t mA[EQ];     // ~0 where j == myj
    uint32_t mP[EQ];     // ~0 where j == myj - 1; mP[4] selects E instead (myj == 0: 6.20 uses E + ht)
    // state
    c32 U;               // element lanes: U(myi, myj)
    c32 C;               // column lanes: eq_coeff[myj]
    float D;             // d[myj]

    __device__ __forceinline__ void init(int lane) {
        const int g = lane & (TC_LANES - 1);
        base = lane & ~(TC_LANES - 1);
        const bool elem = g < 10;
        if (elem) {
            myj = g >= 6 ? 4 : g >= 3 ? 3 : g >= 1 ? 2 : 1;
            myi = g - myj * (myj - 1) / 2;
        } else {
            myj = g == 15 ? 0 : g - 10;
            myi = myj;
        }
        keep = g == 15 ? 0u : 0xffffffffu;
        neg0 = g == 15 ? 0x80000000u : 0u;
        const int zero_lane = base + 15;
        const int jlim = elem ? myj : EQ;
#pragma unroll
        for (int r = 0; r < 4; r++) {
            srcF[r] = r < myj ? base + myj * (myj - 1) / 2 + r : zero_lane;   // product U(r, myj) conj(x[r])
            const int jp = myi + 1 + r;                        // P(myi, jp) is added before column jlim is reached
            srcP[r] = jp < jlim ? base + jp * (jp - 1) / 2 + myi : zero_lane;
        }
        srcG = base + 10 + myi;
#pragma unroll
        for (int j = 0; j < EQ; j++) mA[j] = myj == j ? 0xffffffffu : 0u;
#pragma unroll
        for (int j = 0; j < EQ - 1; j++) mP[j] = myj - 1 == j ? 0xffffffffu : 0u;
        mP[EQ - 1] = myj == 0 ? 0xffffffffu : 0u;
    }

    __device__ __forceinline__ void reset() {                  // kalman_reset(), src/kalman.c:42-55
        U = mk(0.0f, 0.0f);
        C = mk(0.0f, 0.0f);
        D = 1.0f;
    }

    // One train_eq() (DATA = false, src/equalizer.c:45-58) or data_eq() (DATA = true, :64-85) step on the window
    // x[0..4]; this lane is handed xi = x[myi] and xj = x[myj].  Returns crealf(error) on every lane.
    template <bool DATA>
    __device__ __forceinline__ float step(c32 xi, c32 xj, float ref, int &bI, int &bQ) {
        const float E = 0.1f, q = 0.08f;                       // kalman_init(), src/kalman.c:61-62

        // ---- equalizer output from the taps as they stand ----
        const c32 prod = DATA ? cmulc(xj, C) : cmul(xj, C);
        c32 v = mk(0.0f, 0.0f);
#pragma unroll
        for (int i = 0; i < EQ; i++) v = cadd(v, shfl_c(prod, base + 10 + i));
        c32 err;
        if (DATA) {
            bI = v.r < 0.0f;
            bQ = v.i < 0.0f;
            const float ci = bI ? -1.0f : 1.0f;
            const float cq = bQ ? -1.0f : 1.0f;
            err = mk(__fmul_rn(__fsub_rn(ci, v.r), 0.1f), __fmul_rn(__fsub_rn(cq, v.i), 0.1f));
        } else {
            err = mk(__fsub_rn(ref, v.r), v.i);                // conjf(ref - val)
        }

        // ---- kalman_calculate(), src/kalman.c:85-141 ----
        c32 p = cmul(U, cconj(xi));                            // 6.2, this lane's term (-0 on lane 15)
        p = mk(bits_sel(p.r, keep, neg0), bits_sel(p.i, keep, neg0));
        c32 F = cadd(shfl_c(p, srcF[0]), cconj(xj));
#pragma unroll
        for (int r = 1; r < 4; r++) F = cadd(F, shfl_c(p, srcF[r]));
        const c32 G = cscale(F, D);                            // 6.4
        const float t = __fsub_rn(__fmul_rn(G.r, F.r), __fmul_rn(G.i, -F.i));   // 6.5, 6.6

        float A[EQ];
        A[0] = __fadd_rn(E, __shfl_sync(0xffffffffu, t, base + 10));
#pragma unroll
        for (int j = 1; j < EQ; j++) A[j] = __fadd_rn(A[j - 1], __shfl_sync(0xffffffffu, t, base + 10 + j));
        uint32_t am = 0u, ap = __float_as_uint(E) & mP[EQ - 1];
#pragma unroll
        for (int j = 0; j < EQ; j++) am |= __float_as_uint(A[j]) & mA[j];
#pragma unroll
        for (int j = 0; j < EQ - 1; j++) ap |= __float_as_uint(A[j]) & mP[j];
        const float hq = __fadd_rn(1.0f, q);                   // 6.7
        const float ht = __fmul_rn(A[EQ - 1], q);
        const float den_lo = __fadd_rn(A[0], ht), den_hi = __fadd_rn(A[EQ - 1], ht);
        const float den_m = __fadd_rn(__uint_as_float(am), ht);       // a[myj] + ht
        const float den_p = __fadd_rn(__uint_as_float(ap), ht);       // a[myj-1] + ht (6.21), or E + ht for column 0 (6.20)
        float rc_m, rc_p, rc_last;                             // kalman_y after columns myj, myj - 1 and 4
        if (den_lo >= 0x1p-120f && den_hi <= 0x1p120f) {       // a[] is non-decreasing: these two bound the rest,
            rc_m = rcp_rn_normal(den_m);                       // and E + ht lies between ht and den_lo
            rc_p = rcp_rn_normal(den_p);
            rc_last = rcp_rn_normal(den_hi);
        } else {
            rc_m = __frcp_rn(den_m);
            rc_p = __frcp_rn(den_p);
            rc_last = __frcp_rn(den_hi);
        }
        D = __fmul_rn(D, __fmul_rn(__fmul_rn(hq, den_p), rc_m));               // 6.20 / 6.13
        const c32 H = mk(__fmul_rn(-F.r, rc_p), __fmul_rn(-F.i, rc_p));        // 6.11 (unused in column 0)

        // kalman_gain[myi] as it stands when column myj is reached (element lanes) / at the end (column lanes)
        c32 Pm = cmulc(G, U);                                  // 6.16, this lane's term (-0 on lane 15)
        Pm = mk(bits_sel(Pm.r, keep, neg0), bits_sel(Pm.i, keep, neg0));
        c32 acc = shfl_c(G, srcG);
#pragma unroll
        for (int r = 0; r < 4; r++) acc = cadd(acc, shfl_c(Pm, srcP[r]));
        U = cadd(U, cmulc(H, acc));                            // 6.15

        // ---- update_eq(), src/equalizer.c:25-40 ----
        const c32 e2 = cscale(err, rc_last);
        C = cadd(C, cmulc(e2, acc));
        return err.r;
    }
};

}  // namespace sc
