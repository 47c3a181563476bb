"""CPU (numpy): the error bound that makes the tensor-core PROPOSED preamble search exact.

Three kernels (search_mma_batch_kernel, search_umma_batch_kernel, frontend_umma_kernel) compute the 128 correlations of
a window approximately -- d = s.r - s.i and e = s.i + s.r split by truncation into two bf16 pieces, +-1 weights, fp32
accumulation on the tensor cores -- and then evaluate exactly only the lags whose approximate value reaches a threshold
derived from a bound on |approx - reference| (csrc/sc_search_mma.cuh, csrc/sc_umma.cuh: su_candidate_threshold).
The GPU parity tests show the result is the reference's on the inputs they try; this file checks the ARGUMENT on
adversarial inputs with a model of the arithmetic that is pessimistic about the hardware:

  * per component, |approx - reference| <= delta = 1.004 * 2^-13 * sum(|d| + |e|), with the approximate sum accumulated
    in fp32 by round-to-nearest adds, by TRUNCATING adds, and in blocks of 16 (the MMA's K) -- the tensor core's internal
    order and rounding are not specified, all three must fit;
  * the lag the reference picks (strict '>', first maximum, qpsk.c:172-183) is always among the candidates
    { L : v_approx[L] >= threshold(v_max, delta) }.

Reference arithmetic: src/qpsk.c:88-96 (correlate), 75-80 (cnormf)."""
import os
import re

import numpy as np
import pytest

PRE = 128
F32 = np.float32


def preamble_signs():
    """+-1 of the reference's preamble (include/sc_tables.inc, the table the kernels are built from)."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = open(os.path.join(root, "include", "sc_tables.inc")).read()
    body = re.search(r"SC_TABLE_PREAMBLE = \{(.*?)\};", text, re.S).group(1)
    pre = np.array([int(t) for t in re.findall(r"-?\d+", body)], dtype=np.float32)
    assert pre.size == PRE and set(np.unique(pre)) == {-1.0, 1.0}
    return pre


def trunc_bf16(v):
    return (v.view(np.uint32) & np.uint32(0xFFFF0000)).view(np.float32)


def split2(v):
    hi = trunc_bf16(v)
    r = (v - hi).astype(F32)                                 # exact: the low 16 significand bits
    return hi, trunc_bf16(r)


def trunc_f32(x64):
    """float64 -> float32 rounding toward zero (a truncating adder's result)."""
    y = x64.astype(F32)
    over = np.abs(y.astype(np.float64)) > np.abs(x64)
    y[over] = np.nextafter(y[over], F32(0.0))
    return y


def reference_sums(x, sign):
    """The reference's 128 sequential fp32 adds for every lag: out[L] = (((0 + s0 x[L]) + s1 x[L+1]) + ...)."""
    a = np.zeros(PRE, dtype=F32)
    lags = np.arange(PRE)
    for i in range(PRE):
        a = (a + sign[i] * x[lags + i]).astype(F32)
    return a


def approx_sums(x, sign, mode):
    hi, mid = split2(x)
    lags = np.arange(PRE)
    out = []
    for piece in (hi, mid):
        if mode == "rn":
            a = np.zeros(PRE, dtype=F32)
            for i in range(PRE):
                a = (a + sign[i] * piece[lags + i]).astype(F32)
        elif mode == "trunc":
            a = np.zeros(PRE, dtype=F32)
            for i in range(PRE):
                a = trunc_f32(a.astype(np.float64) + (sign[i] * piece[lags + i]).astype(np.float64))
        else:                                                # blocks of 16 exact, the accumulator add truncated
            a = np.zeros(PRE, dtype=F32)
            for b in range(0, PRE, 16):
                blk = np.zeros(PRE, dtype=np.float64)
                for i in range(b, b + 16):
                    blk += (sign[i] * piece[lags + i]).astype(np.float64)
                a = trunc_f32(a.astype(np.float64) + blk)
        out.append(a)
    return (out[0] + out[1]).astype(F32)                     # the epilogue's hi + mid


def threshold(vmax, delta):
    """su_candidate_threshold() (csrc/sc_umma.cuh), with an exact sqrt (the kernel's approximation carries a 1e-4 margin)."""
    m = F32(np.sqrt(np.float64(F32(2.0) * vmax))) * F32(1.0001)
    mu = F32(F32(F32(2.0) * delta) * F32(m + delta)) + F32(vmax * F32(2.0 ** -20))
    return F32(vmax - F32(mu * F32(2.002)))


def windows(rng):
    n = 2 * PRE - 1
    yield "gaussian", rng.normal(0, 3000, (n, 2))
    yield "tiny + huge", rng.normal(0, 1, (n, 2)) * 10.0 ** rng.integers(-6, 7, (n, 1))
    yield "cancelling", np.repeat(rng.normal(0, 1e4, (1, 2)), n, axis=0) + rng.normal(0, 1e-2, (n, 2))
    yield "one spike", np.where(np.arange(n)[:, None] == 77, 1e7, rng.normal(0, 1, (n, 2)))
    yield "mantissa full", (rng.integers(1 << 23, 1 << 24, (n, 2)) * rng.choice([-1, 1], (n, 2))).astype(np.float64)
    yield "near silence", rng.integers(-3, 4, (n, 2)).astype(np.float64) * 2.2e-3


@pytest.mark.parametrize("mode", ["rn", "trunc", "block16"])
def test_bound_holds_and_the_reference_maximum_is_a_candidate(mode, oracle):
    sign = preamble_signs()
    rng = np.random.default_rng(20261019)
    worst = 0.0
    for rep in range(12):
        for name, w in windows(rng):
            s = w.astype(F32)
            d = (s[:, 0] - s[:, 1]).astype(F32)              # qpsk.c:88-96 with pre = v(1 + i)
            e = (s[:, 1] + s[:, 0]).astype(F32)
            s_abs = F32(np.sum(np.abs(d.astype(np.float64)) + np.abs(e.astype(np.float64))))
            delta = F32(s_abs * F32(float.fromhex("0x1.004p-13")))
            ref_d, ref_e = reference_sums(d, sign), reference_sums(e, sign)
            app_d, app_e = approx_sums(d, sign, mode), approx_sums(e, sign, mode)
            err = max(np.max(np.abs(app_d.astype(np.float64) - ref_d)), np.max(np.abs(app_e.astype(np.float64) - ref_e)))
            assert err <= float(delta), (name, rep, err, float(delta))
            if delta > 0:
                worst = max(worst, err / float(delta))
            v_ref = (ref_d * ref_d).astype(F32) + (ref_e * ref_e).astype(F32)        # cnormf, qpsk.c:75-80
            v_app = (app_d * app_d).astype(F32) + (app_e * app_e).astype(F32)
            if not np.all(np.isfinite(v_app)):
                continue                                     # overflow: no candidate, the kernels fall back to the exact search
            best = int(np.argmax(v_ref)) if v_ref.max() > 0 else 0                    # first maximum, strict '>'
            if mode == "rn":                                 # the model of the reference above IS the oracle's search
                oi, ov = oracle.search((s[:, 0] + 1j * s[:, 1]).astype(np.complex64))
                assert oi == best and F32(ov).view(np.uint32) == v_ref[best].view(np.uint32), (name, rep, oi, best)
            thr = threshold(v_app.max(), delta)
            assert v_app[best] >= thr, (name, rep, best, float(v_app[best]), float(thr))
    assert worst < 0.6, worst                                # the margin the kernels' comments claim (about a factor of two)
