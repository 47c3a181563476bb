"""GPU parity of the stage entry points (batched forms of fir / correlate+argmax / the decision loop /
fft) against the oracle and the committed reference vectors.  Bit-exact unless stated."""
import ctypes as C

import numpy as np
import pytest

from oracle import pyoracle as po

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sc():
    import singlecarrier_b200 as m
    assert m.lib.sc_device_count() > 0
    return m


def dev(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def c64_to_f32(t):
    return t.cpu().numpy().view(np.complex64)


@pytest.mark.parametrize("wide", [0, 1])
def test_fir_batch_golden_and_oracle(sc, oracle, gold, wide):
    import torch
    g = gold("stage_golden.npz")
    mem = dev(g[f"fir_mem_in_{wide}"].view(np.float32))
    x = dev(g[f"fir_x_{wide}"].view(np.float32))
    sc._lib.check(sc.lib.sc_fir_batch_dev(0, 1, wide, mem.data_ptr(), x.data_ptr(), 700, 700, 0))
    torch.cuda.synchronize()
    assert np.array_equal(c64_to_f32(x).view(np.uint32), g[f"fir_y_{wide}"].view(np.uint32))
    assert np.array_equal(c64_to_f32(mem).view(np.uint32), g[f"fir_mem_out_{wide}"].view(np.uint32))
    # many streams, ragged lengths (shorter than the delay line, tile edges), split calls == one call
    rng = np.random.default_rng(3 + wide)
    for length in (1, 7, 48, 49, 50, 639, 640, 641, 1880):
        ns = 19
        mem0 = (rng.normal(size=(ns, 49)) + 1j * rng.normal(size=(ns, 49))).astype(np.complex64)
        x0 = (rng.normal(size=(ns, length + 5)) + 1j * rng.normal(size=(ns, length + 5))).astype(np.complex64)
        dm, dx = dev(mem0.view(np.float32)), dev(x0.view(np.float32))
        sc._lib.check(sc.lib.sc_fir_batch_dev(0, ns, wide, dm.data_ptr(), dx.data_ptr(), length + 5, length, 0))
        torch.cuda.synchronize()
        ym, yx = c64_to_f32(dm), c64_to_f32(dx)
        for s in range(ns):
            m, v = mem0[s].copy(), x0[s, :length].copy()
            oracle.fir(m, bool(wide), v)
            assert np.array_equal(yx[s, :length].view(np.uint32), v.view(np.uint32)), (length, s)
            assert np.array_equal(ym[s].view(np.uint32), m.view(np.uint32)), (length, s)
            assert np.array_equal(yx[s, length:], x0[s, length:])            # beyond `length` untouched


def test_preamble_search_batch(sc, oracle):
    import torch
    rng = np.random.default_rng(4)
    ns = 300
    sym = (rng.normal(size=(ns, 260)) + 1j * rng.normal(size=(ns, 260))).astype(np.complex64)
    pre = np.array(po.Oracle().lib and [1.0] * 0)                           # placeholder to keep flake quiet
    del pre
    # plant preambles at known lags in some streams, all-zero and tie cases in others
    pv = np.frombuffer((C.c_int8 * 128).in_dll(oracle.lib, "sco_preamblevalues"), np.int8).astype(np.float32)
    for s in range(0, 100):
        lag = int(rng.integers(0, 128))
        sym[s, lag:lag + 128] += (6.0 * pv * (1 + 1j)).astype(np.complex64)
    sym[100] = 0                                                            # silence -> (0, 0.0)
    sym[101] = 0
    sym[101, 200] = 1 + 1j                                                  # several lags tie -> first wins
    d = dev(sym.view(np.float32))
    idx = torch.empty(ns, dtype=torch.int32, device="cuda")
    val = torch.empty(ns, dtype=torch.float32, device="cuda")
    sc._lib.check(sc.lib.sc_preamble_search_batch_dev(0, ns, d.data_ptr(), 260, idx.data_ptr(), val.data_ptr(), 0))
    torch.cuda.synchronize()
    gi, gv = idx.cpu().numpy(), val.cpu().numpy()
    for s in range(ns):
        oi, ov = oracle.search(sym[s])
        assert gi[s] == oi and gv[s].view(np.uint32) == np.float32(ov).view(np.uint32), s
    assert gi[100] == 0 and gv[100] == 0.0


def test_track_decide_batch_vs_oracle_frames(sc, oracle):
    """Feed the oracle's own decimated windows (all 290 symbols) through the stage API."""
    import torch
    from helpers import synth_streams
    rng = np.random.default_rng(12)
    nf, ns = 8, 24
    samples = synth_streams(oracle, rng, ns, nf)
    wins, mi, mv, tin, exp = [], [], [], [], []
    for s in range(ns):
        st = oracle.new_state()
        for n in range(nf):
            timing_before = int(np.frombuffer(st, np.uint8)[:0].size)         # (unused)
            del timing_before
            bits, stats = oracle.rx_frame(st, samples[s, n * 1880:(n + 1) * 1880])
            # the oracle state keeps dec[] after the call: bytes [input_frame][dec]...
            dec = np.frombuffer(st, np.float32, count=2 * 752, offset=3760 * 8).view(np.complex64)[:290].copy()
            if n >= 2:
                wins.append(dec)
                mi.append(stats["max_index"])
                mv.append(stats["max_value"])
                exp.append((n, bits.copy(), stats.copy()))
    # rx_timing at entry of each call = rx_timing after the previous call
    k = 0
    tin = []
    for s in range(ns):
        prev = None
        st = oracle.new_state()
        for n in range(nf):
            _, stats = oracle.rx_frame(st, samples[s, n * 1880:(n + 1) * 1880])
            if n >= 2:
                tin.append(prev)
            prev = int(stats["rx_timing"])
    nrec = len(wins)
    d_sym = dev(np.stack(wins).view(np.float32))
    d_mi = dev(np.array(mi, np.int32))
    d_mv = dev(np.array(mv, np.float32))
    calls = sorted(set(e[0] for e in exp))
    res_all = np.zeros(nrec, sc.RESULT_DTYPE)
    eq_all = np.zeros((nrec, 10), np.float32)
    tout = np.zeros(nrec, np.int32)
    for n in calls:                                                          # one launch per call index (keystream)
        sel = np.array([i for i, e in enumerate(exp) if e[0] == n])
        ds = dev(np.stack(wins)[sel].view(np.float32))
        dmi, dmv = dev(np.array(mi, np.int32)[sel]), dev(np.array(mv, np.float32)[sel])
        dt = dev(np.array(tin, np.int32)[sel])
        dres = torch.zeros((len(sel), 32), dtype=torch.uint8, device="cuda")
        deq = torch.zeros((len(sel), 10), dtype=torch.float32, device="cuda")
        sc._lib.check(sc.lib.sc_track_decide_batch_dev(0, len(sel), ds.data_ptr(), 290, dmi.data_ptr(), dmv.data_ptr(),
                                                       dt.data_ptr(), n, dres.data_ptr(), deq.data_ptr(), 0))
        torch.cuda.synchronize()
        res_all[sel] = dres.cpu().numpy().view(sc.RESULT_DTYPE)[:, 0]
        eq_all[sel] = deq.cpu().numpy()
        tout[sel] = dt.cpu().numpy()
    del d_sym, d_mi, d_mv, k
    for i, (n, bits, stats) in enumerate(exp):
        r = res_all[i]
        assert r["valid"] == stats["valid"] and r["matches"] == stats["matches"], i
        assert r["rx_timing"] == stats["rx_timing"] == tout[i]
        assert np.float32(r["cost"]).view(np.uint32) == np.float32(stats["cost"]).view(np.uint32)
        assert np.array_equal(eq_all[i].view(np.uint32), stats["eq_coeff"].view(np.uint32))
        if stats["valid"]:
            assert np.array_equal(sc.unpack_bits(res_all[i:i + 1])[0], bits)


FFT_SIZES = [2, 3, 4, 5, 6, 7, 8, 11, 16, 30, 49, 60, 64, 100, 125, 128, 243, 256, 512, 1024, 2048]


@pytest.mark.parametrize("n", FFT_SIZES)
def test_fft_batch_matches_reference_fft(sc, gold, n):
    """fft.h: same factorisation/butterflies/twiddles as src/fft.c => bit-identical, for
    radix 4, 2, 3, 5 and generic (7, 11) stages.  The reference's radix-5 butterfly (src/fft.c:324-334) does
    not compute a DFT (relative error ~1 against the definition for every length with a factor 5); it is
    mirrored as written, so the DFT-definition check below only runs for lengths without that factor."""
    import torch
    g = gold("fft_golden.npz")
    x = g[f"c{n}_in"]
    batch = np.stack([x, x[::-1].copy(), x * np.complex64(0.5)])
    for inv in (0, 1):
        d_in = dev(batch.view(np.float32))
        d_out = torch.zeros_like(d_in)
        sc._lib.check(sc.lib.sc_fft_batch_dev(0, 3, n, inv, d_in.data_ptr(), d_out.data_ptr(), 0))
        torch.cuda.synchronize()
        y = c64_to_f32(d_out)
        ref = g[f"c{n}_{inv}"]
        assert np.array_equal(y[0].view(np.uint32), ref.view(np.uint32)), (n, inv, np.abs(y[0] - ref).max())
        # independent check against the DFT definition in float64
        if n % 5 != 0:
            want = np.fft.ifft(batch.astype(np.complex128), axis=1) * n if inv else np.fft.fft(batch.astype(np.complex128), axis=1)
            assert np.abs(y - want).max() <= 1e-5 * np.abs(want).max() * max(1.0, np.log2(n))


@pytest.mark.parametrize("n", [6, 20, 64, 250, 256])
def test_fftr_fftri_match_reference(sc, gold, n):
    import torch
    g = gold("fft_golden.npz")
    xr = g[f"r{n}_in"]
    d_in = dev(xr)
    d_spec = torch.zeros((n // 2 + 1) * 2, dtype=torch.float32, device="cuda")
    sc._lib.check(sc.lib.sc_fftr_batch_dev(0, 1, n, d_in.data_ptr(), d_spec.data_ptr(), 0))
    d_back = torch.zeros(n, dtype=torch.float32, device="cuda")
    sc._lib.check(sc.lib.sc_fftri_batch_dev(0, 1, n, d_spec.data_ptr(), d_back.data_ptr(), 0))
    torch.cuda.synchronize()
    assert np.array_equal(c64_to_f32(d_spec).view(np.uint32), g[f"r{n}_spec"].view(np.uint32))
    assert np.array_equal(d_back.cpu().numpy().view(np.uint32), g[f"r{n}_back"].view(np.uint32))
    if n % 5 != 0:
        assert np.abs(d_back.cpu().numpy() / n - xr).max() < 1e-5


def test_large_fft_uses_global_scratch(sc):
    import torch
    rng = np.random.default_rng(1)
    n = 16384
    x = (rng.normal(size=(2, n)) + 1j * rng.normal(size=(2, n))).astype(np.complex64)
    d_in = dev(x.view(np.float32))
    d_out = torch.zeros_like(d_in)
    sc._lib.check(sc.lib.sc_fft_batch_dev(0, 2, n, 0, d_in.data_ptr(), d_out.data_ptr(), 0))
    torch.cuda.synchronize()
    want = np.fft.fft(x.astype(np.complex128), axis=1)
    assert np.abs(c64_to_f32(d_out) - want).max() <= 2e-5 * np.abs(want).max() * np.log2(n)


def test_lock_stats_counters(sc, oracle):
    import torch
    from helpers import synth_streams
    rng = np.random.default_rng(21)
    ns, nf = 50, 9
    samples = synth_streams(oracle, rng, ns, nf)
    bank = sc.ModemBank(ns)
    d_in = dev(samples)
    d_res = torch.zeros((ns, nf * 32), dtype=torch.uint8, device="cuda")
    bank.rx_frames_dev(d_in, nf, d_res)
    cnt = torch.zeros(16, dtype=torch.int64, device="cuda")
    bank.lock_stats(d_res, nf, cnt)
    torch.cuda.synchronize()
    c = cnt.cpu().numpy()
    r = d_res.cpu().numpy().view(sc.RESULT_DTYPE)
    v = r["valid"].astype(bool)
    assert c[0] == ns * nf and c[1] == v.sum() and c[2] == r["matches"].sum() and c[3] == r["matches"][v].sum()
    assert c[4] == r["max_index"][v].sum() and c[7] == r["rx_timing"].sum()
    assert c[5] == sum(bin(int(b)).count("1") for b in r["bits"][v])
    assert c[8:].sum() == ns * nf
    bank.close()


def test_branch_free_reciprocal_is_correctly_rounded(sc):
    """Every float in [2^-120, 2^120] (2.0e9 bit patterns): rcp_rn_normal == __frcp_rn, bit for bit."""
    import struct
    import torch
    lo = struct.unpack("<I", struct.pack("<f", 2.0 ** -120))[0]
    hi = struct.unpack("<I", struct.pack("<f", 2.0 ** 120))[0]
    bad = torch.zeros(1, dtype=torch.int64, device="cuda")
    sc._lib.check(sc.lib.sc_selftest_rcp_dev(0, lo, hi, bad.data_ptr(), 0))
    torch.cuda.synchronize()
    assert hi - lo > 2_000_000_000 and int(bad.item()) == 0


def test_fir_fast_mode_is_tolerance_parity_only(sc, oracle):
    """SC_FIR_FAST (contracted multiply-adds): agrees with the reference to 1e-5 relative, not bit for bit."""
    import torch
    rng = np.random.default_rng(31)
    ns, length = 7, 3000
    mem0 = (rng.normal(size=(ns, 49)) + 1j * rng.normal(size=(ns, 49))).astype(np.complex64)
    x0 = (rng.normal(size=(ns, length)) + 1j * rng.normal(size=(ns, length))).astype(np.complex64)
    dm, dx = dev(mem0.view(np.float32)), dev(x0.view(np.float32))
    sc._lib.check(sc.lib.sc_fir_batch_dev(0, ns, 2, dm.data_ptr(), dx.data_ptr(), length, length, 0))
    torch.cuda.synchronize()
    y = c64_to_f32(dx)
    exact = x0.copy()
    for s in range(ns):
        m = mem0[s].copy()
        oracle.fir(m, False, exact[s])
        assert np.array_equal(c64_to_f32(dm)[s].view(np.uint32), m.view(np.uint32))       # the delay line holds raw inputs
    scale = np.abs(exact).max()
    assert np.abs(y - exact).max() <= 1e-5 * scale                                         # tolerance stated: 1e-5 relative
    assert not np.array_equal(y.view(np.uint32), exact.view(np.uint32))                    # and it really is a different rounding
