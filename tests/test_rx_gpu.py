"""GPU parity: the CUDA RX chain behind the C ABI vs the CPU oracle, bit for bit.

Covers SURVEY section 8 rows a-1..a-8 end to end (mixer, RRC, decimator, preamble search, training,
decision loop, Kalman gain, descrambler) on the reference's shipped file (config 1), the committed
golden streams, seeded synthetic loop-back streams (config 2 style) and the silence/noise edge cases
(SURVEY F6).  Tolerance: none -- every integer AND every float field must be bit-identical.
"""
import numpy as np
import pytest

from helpers import compare_results, oracle_results, synth_streams

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sc():
    import singlecarrier_b200 as m
    assert m.lib.sc_device_count() > 0, "no CUDA device"
    return m


def run_gpu(sc, samples, n_frames, **kw):
    bank = sc.ModemBank(samples.shape[0], debug_eq=True, **kw)
    res, eq = bank.rx_frames_host(samples, n_frames)
    bank.close()
    return res, eq


def test_shipped_file_config1(sc, oracle, gold):
    x = gold("preamble_qpsk_8k.raw")
    nf = x.size // 1880
    assert nf == 14
    samples = x[: nf * 1880].reshape(1, -1).copy()
    res, eq = run_gpu(sc, samples, nf)
    g = gold("rx_shipped.npz")
    assert res["valid"][0].tolist() == g["valid"].tolist() == [1, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 0]
    assert res["max_index"][0].tolist() == g["max_index"].tolist()
    assert res["rx_timing"][0].tolist() == g["rx_timing"].tolist()
    assert np.array_equal(res["max_value"][0].view(np.uint32), g["max_value"].view(np.uint32))
    assert np.array_equal(eq[0].view(np.uint32), g["eq_coeff"].view(np.uint32))
    rows = sc.unpack_bits(res)[0]
    v = g["valid"].astype(bool)
    assert np.array_equal(rows[v], g["bits"][v])
    assert res["matches"][0][12] == 110 and res["max_index"][0][12] == 120
    obits, ostats = oracle_results(oracle, samples, nf)
    assert compare_results(res, eq, obits, ostats) == []


def test_golden_synth_streams(sc, oracle, gold):
    g = gold("rx_synth.npz")
    samples = g["samples"]
    nf = samples.shape[1] // 1880
    res, eq = run_gpu(sc, samples, nf)
    assert np.array_equal(res["valid"].astype(np.int32), g["valid"])
    assert np.array_equal(res["max_index"].astype(np.int32), g["max_index"])
    assert np.array_equal(res["max_value"].view(np.uint32), g["max_value"].view(np.uint32))
    assert np.array_equal(eq.view(np.uint32), g["eq_coeff"].view(np.uint32))
    v = g["valid"].astype(bool)
    assert np.array_equal(sc.unpack_bits(res)[v], g["bits"][v])
    assert v.sum() > 20


@pytest.mark.parametrize("n_streams,n_frames,seed", [(1, 6, 1), (37, 9, 2), (300, 12, 3)])
def test_synthetic_loopback_vs_oracle(sc, oracle, n_streams, n_frames, seed):
    rng = np.random.default_rng(seed)
    samples = synth_streams(oracle, rng, n_streams, n_frames)
    res, eq = run_gpu(sc, samples, n_frames)
    obits, ostats = oracle_results(oracle, samples, n_frames)
    assert compare_results(res, eq, obits, ostats) == []


def test_wide_filter(sc, oracle):
    rng = np.random.default_rng(7)
    samples = synth_streams(oracle, rng, 40, 8)
    res, eq = run_gpu(sc, samples, 8, wide=True)
    obits, ostats = oracle_results(oracle, samples, 8, wide=True)
    assert compare_results(res, eq, obits, ostats) == []


def test_foffset(sc, oracle):
    rng = np.random.default_rng(8)
    samples = synth_streams(oracle, rng, 33, 8)
    res, eq = run_gpu(sc, samples, 8, foffset_hz=7.5)
    obits, ostats = oracle_results(oracle, samples, 8, foffset=7.5)
    assert compare_results(res, eq, obits, ostats) == []


def test_silence_and_noise_edge_cases(sc, oracle):
    rng = np.random.default_rng(9)
    nf = 8
    rows = [np.zeros(nf * 1880, np.int16)]                                   # silence: every call "valid" (F6)
    for amp in (1, 3, 100, 3000, 32767):
        rows.append(rng.integers(-amp, amp + 1, nf * 1880).astype(np.int16))
    rows.append(np.where(rng.integers(0, 2, nf * 1880) > 0, 32767, -32768).astype(np.int16))
    samples = np.stack(rows)
    res, eq = run_gpu(sc, samples, nf)
    obits, ostats = oracle_results(oracle, samples, nf)
    assert compare_results(res, eq, obits, ostats) == []
    assert res["valid"][0].all() and (res["max_index"][0] == 0).all()
    ks = [sc.keystream_word(n) for n in range(nf)]
    assert res["bits"][0].tolist() == ks                                     # silence decodes the keystream


def test_streaming_equals_one_shot(sc, oracle):
    """Feeding the same frames in several API calls (state carried in the handle) changes nothing."""
    rng = np.random.default_rng(10)
    nf = 11
    samples = synth_streams(oracle, rng, 70, nf)
    one, eq1 = run_gpu(sc, samples, nf)
    bank = sc.ModemBank(70, debug_eq=True)
    parts, eqs = [], []
    for a, b in ((0, 1), (1, 4), (4, 5), (5, 11)):
        r, e = bank.rx_frames_host(np.ascontiguousarray(samples[:, a * 1880:b * 1880]), b - a)
        parts.append(r)
        eqs.append(e)
    bank.close()
    cat = np.concatenate(parts, axis=1)
    assert cat.tobytes() == one.tobytes()
    assert np.concatenate(eqs, axis=1).tobytes() == eq1.tobytes()


def test_device_api_and_strides(sc, oracle):
    import torch
    rng = np.random.default_rng(11)
    nf, ns = 7, 130
    samples = synth_streams(oracle, rng, ns, nf)
    pad = 64
    d_in = torch.zeros((ns, nf * 1880 + pad), dtype=torch.int16, device="cuda")
    d_in[:, : nf * 1880] = torch.from_numpy(samples).cuda()
    d_res = torch.zeros((ns, (nf + 1) * 32), dtype=torch.uint8, device="cuda")
    bank = sc.ModemBank(ns)
    bank.rx_frames_dev(d_in, nf, d_res)
    torch.cuda.synchronize()
    res = d_res.cpu().numpy().view(sc.RESULT_DTYPE)[:, :nf]
    bank.close()
    obits, ostats = oracle_results(oracle, samples, nf)
    assert compare_results(res, None, obits, ostats) == []
    assert (res["call_index"] == np.arange(nf)[None, :]).all()


def test_odd_strides_use_the_generic_load_path(sc, oracle):
    """Streams whose first sample is not 4-byte aligned (odd stream stride) take the scalar-load path of
    the front-end; results must not change."""
    import torch
    rng = np.random.default_rng(13)
    nf, ns = 6, 21
    samples = synth_streams(oracle, rng, ns, nf)
    stride = nf * 1880 + 3                                              # odd: every other stream starts at an odd sample
    flat = torch.zeros(ns * stride + 8, dtype=torch.int16, device="cuda")
    d_in = flat[1:1 + ns * stride].view(ns, stride)                     # and the base pointer itself is 2-byte aligned only
    d_in[:, : nf * 1880] = torch.from_numpy(samples).cuda()
    assert d_in.data_ptr() % 4 == 2 and d_in.stride(0) % 2 == 1
    d_res = torch.zeros((ns, nf * 32), dtype=torch.uint8, device="cuda")
    bank = sc.ModemBank(ns)
    bank.rx_frames_dev(d_in, nf, d_res)
    torch.cuda.synchronize()
    res = d_res.cpu().numpy().view(sc.RESULT_DTYPE)
    bank.close()
    obits, ostats = oracle_results(oracle, samples, nf)
    assert compare_results(res, None, obits, ostats) == []


def test_argument_validation(sc):
    import ctypes as C
    L = sc.lib
    h = C.c_void_p()
    assert L.sc_create(C.byref(h), 0, 0, 0, 0.0) == -1                   # n_streams must be positive
    assert L.sc_create(C.byref(h), 99, 4, 0, 0.0) == -1                  # no such device
    assert b"device" in L.sc_last_error()
    bank = sc.ModemBank(4)
    x = np.zeros((4, 1880), np.int16)
    r = np.zeros((4, 1), sc.RESULT_DTYPE)
    assert L.sc_rx_frames_host(bank._h, x.ctypes.data, 100, 1, r.ctypes.data, 1, None) == -1      # stride < batch
    assert L.sc_rx_frames_host(bank._h, x.ctypes.data, 1880, 1, r.ctypes.data, 1, r.ctypes.data) == -1  # eq_dbg w/o flag
    assert L.sc_rx_frames_host(bank._h, x.ctypes.data, 1880, 0, r.ctypes.data, 1, None) == 0      # empty batch is a no-op
    assert bank.call_index == 0
    assert L.sc_fft_batch_dev(0, 1, 0, 0, x.ctypes.data, x.ctypes.data, None) == -1
    bank.close()


def test_long_run_many_small_batches(sc, oracle):
    """120 calls (> 1 minute of audio) fed in uneven batches: the NCO phasor table, its per-frame
    renormalisation and the scrambler position must carry across API calls exactly."""
    rng = np.random.default_rng(14)
    nf, ns = 120, 5
    samples = synth_streams(oracle, rng, ns, nf, noise_levels=(0.0, 200.0, 2000.0))
    bank = sc.ModemBank(ns, debug_eq=True)
    parts, eqs, a = [], [], 0
    sizes = [1, 2, 3, 7, 1, 13, 29, 5, 11, 48]
    assert sum(sizes) == nf
    for n in sizes:
        r, e = bank.rx_frames_host(np.ascontiguousarray(samples[:, a * 1880:(a + n) * 1880]), n)
        parts.append(r)
        eqs.append(e)
        a += n
    assert bank.call_index == nf
    bank.close()
    res, eq = np.concatenate(parts, axis=1), np.concatenate(eqs, axis=1)
    obits, ostats = oracle_results(oracle, samples, nf)
    assert compare_results(res, eq, obits, ostats) == []
    assert (res["call_index"] == np.arange(nf)[None, :]).all()
    assert ostats["valid"][:, 2:].sum() > 20


def test_nco_table_matches_reference_recurrence(sc, oracle):
    """The device-generated RX/TX phasor tables against a float32 re-run of the reference's recurrence
    (cmul per sample, cabsf renormalisation per frame), bit for bit over 60 frames."""
    import ctypes as C
    libm = C.CDLL("libm.so.6")
    libm.hypotf.restype = C.c_float
    libm.hypotf.argtypes = [C.c_float, C.c_float]
    bank = sc.ModemBank(1)
    tab = bank.nco_table(0, 60)
    bank.close()
    r = oracle.nco_rect(-1100.0)
    rr, ri = np.float32(r.real), np.float32(r.imag)
    pr, pi = np.float32(1), np.float32(0)
    want = np.zeros((60, 1880), np.complex64)
    for f in range(60):
        for i in range(1880):
            pr, pi = np.float32(np.float32(pr * rr) - np.float32(pi * ri)), np.float32(np.float32(pr * ri) + np.float32(pi * rr))
            want[f, i] = pr + 1j * pi
        m = np.float32(libm.hypotf(float(pr), float(pi)))
        pr, pi = np.float32(pr / m), np.float32(pi / m)
    assert np.array_equal(tab.view(np.uint32), want.view(np.uint32))


def test_create_destroy_does_not_leak_device_memory(sc):
    import torch
    torch.cuda.synchronize()
    x = np.zeros((4096, 3 * 1880), np.int16)
    for k in range(3):                                      # warm the allocator / library state
        b = sc.ModemBank(4096, debug_eq=True)
        b.rx_frames_host(x, 3)
        b.close()
    free0, _ = torch.cuda.mem_get_info()
    for k in range(25):
        b = sc.ModemBank(4096, debug_eq=(k % 2 == 0))
        b.rx_frames_host(x, 3)
        b.close()
    free1, _ = torch.cuda.mem_get_info()
    assert free0 - free1 < 64 << 20


def test_banks_are_independent_and_thread_safe_per_handle(sc, oracle):
    """State lives in the handle (the reference keeps it in process globals): two banks driven from two host
    threads at the same time, with different filters and interleaved batches, do not disturb each other."""
    import threading
    rng = np.random.default_rng(15)
    nf = 9
    sa = synth_streams(oracle, rng, 33, nf)
    sb = synth_streams(oracle, rng, 45, nf)
    out = {}

    def drive(key, samples, wide):
        bank = sc.ModemBank(samples.shape[0], wide=wide, debug_eq=True)
        parts, eqs = [], []
        for a, b in ((0, 2), (2, 3), (3, 9)):
            r, e = bank.rx_frames_host(np.ascontiguousarray(samples[:, a * 1880:b * 1880]), b - a)
            parts.append(r)
            eqs.append(e)
        bank.close()
        out[key] = (np.concatenate(parts, axis=1), np.concatenate(eqs, axis=1))

    ts = [threading.Thread(target=drive, args=("a", sa, False)), threading.Thread(target=drive, args=("b", sb, True))]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    for key, samples, wide in (("a", sa, False), ("b", sb, True)):
        obits, ostats = oracle_results(oracle, samples, nf, wide=wide)
        assert compare_results(out[key][0], out[key][1], obits, ostats) == []
