"""GPU: round-2 additions -- parity against the reference's OWN objects on the GPU box, the host-path copy
modes, checkpoint/resume, the warp-shuffle FFT, stream-ordered FFT scratch, the device counters and the
NCCL reduction in the C ABI."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

from helpers import synth_streams

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def sc():
    import singlecarrier_b200 as m
    assert m.lib.sc_device_count() > 0, "no CUDA device"
    return m


def test_cuda_path_vs_reference_objects_directly(sc, ref, oracle, gold):
    """CUDA <-> oracle/_ref/libsc_ref.so (the reference's own .c files compiled by oracle/Makefile) with no
    restatement in between: the shipped file plus 64 noisy loop-back streams."""
    x = gold("preamble_qpsk_8k.raw")
    shipped = np.zeros((1, 14 * 1880), np.int16)
    shipped[0] = x[: 14 * 1880]
    rng = np.random.default_rng(2024)
    noisy = synth_streams(oracle, rng, 64, 14)                 # oracle TX is only the signal source here
    samples = np.concatenate([shipped, noisy])
    bank = sc.ModemBank(samples.shape[0], debug_eq=True)
    res, eq = bank.rx_frames_host(samples, 14)
    bank.close()
    rows = sc.unpack_bits(res)
    n_valid = 0
    for s in range(samples.shape[0]):
        rbits, rst = ref.run_stream(samples[s])
        assert res["valid"][s].astype(np.int32).tolist() == rst["valid"].tolist(), s
        assert res["max_index"][s].astype(np.int32).tolist() == rst["max_index"].tolist(), s
        assert res["rx_timing"][s].astype(np.int32).tolist() == rst["rx_timing"].tolist(), s
        assert np.array_equal(res["max_value"][s].view(np.uint32), rst["max_value"].view(np.uint32)), s
        assert np.array_equal(eq[s].view(np.uint32), np.ascontiguousarray(rst["eq_coeff"]).view(np.uint32)), s
        v = rst["valid"].astype(bool)
        assert np.array_equal(rows[s][v], rbits[v]), s
        # matches / Mean only exist in the reference's DEBUG2 printf of valid frames (qpsk.c:198)
        assert res["matches"][s][v].astype(np.int32).tolist() == rst["matches"][v].tolist(), s
        assert np.array_equal(res["cost"][s][v].view(np.uint32), rst["mean"][v].view(np.uint32)), s
        n_valid += int(v.sum())
    assert n_valid > 100


@pytest.mark.parametrize("extra_stride", [0, 1880, 246])
def test_h2d_modes_give_identical_results(sc, oracle, extra_stride):
    """SC_H2D_COLUMNS / ROWS / FULL / COLUMNS_3D move different bytes but never a different result; the byte
    counters report what was queued."""
    from singlecarrier_b200.modem import H2D_COLUMNS, H2D_COLUMNS_3D, H2D_FULL, H2D_ROWS, OPT_H2D_MODE, OPT_SLAB_PARTS
    rng = np.random.default_rng(5)
    ns, nf = 300, 7
    dense = synth_streams(oracle, rng, ns, nf)
    wide = np.zeros((ns, nf * 1880 + extra_stride), np.int16)
    wide[:, : nf * 1880] = dense
    got = {}
    for mode in (H2D_COLUMNS, H2D_ROWS, H2D_FULL, H2D_COLUMNS_3D):
        for parts in (0, 3):
            bank = sc.ModemBank(ns)
            bank.set_option(OPT_H2D_MODE, mode)
            bank.set_option(OPT_SLAB_PARTS, parts)
            res, _ = bank.rx_frames_host(wide, nf)
            h2d, d2h = bank.transfer_bytes()
            bank.close()
            got[(mode, parts)] = res.tobytes()
            assert d2h == ns * nf * 32
            if mode == H2D_FULL:
                assert h2d == ns * nf * 3760
            elif mode in (H2D_COLUMNS, H2D_COLUMNS_3D):
                assert h2d == ns * nf * 1624 * 2
            else:
                assert ns * nf * 1624 * 2 < h2d < ns * nf * 3760
    first = got[(H2D_COLUMNS, 0)]
    assert all(v == first for v in got.values())


def test_pinned_buffer_and_probe(sc):
    buf = sc.PinnedBuffer(4 * 1880 * 2 * 10, device=0)
    a = buf.array(np.int16, (10, 4 * 1880))
    a[:] = 0
    bank = sc.ModemBank(10)
    res, _ = bank.rx_frames_host(a, 4)
    bank.close()
    assert res["valid"][:, :2].all()                             # silence is "valid" for the first two calls (F6)
    buf.close()
    gbs = sc.h2d_probe(0, 64 << 20, min_seconds=0.05)
    gbs2d = sc.h2d_probe(0, 64 << 20, row_bytes=3248, src_pitch_bytes=157920, min_seconds=0.05)
    assert gbs > 1.0 and gbs2d > 1.0


def test_state_export_import_resumes_bit_exactly(sc, oracle):
    """Checkpoint after 5 calls, continue in a NEW bank: identical to the uninterrupted run."""
    rng = np.random.default_rng(77)
    ns, nf = 130, 12
    samples = synth_streams(oracle, rng, ns, nf)
    a = sc.ModemBank(ns, debug_eq=True)
    whole, eq_whole = a.rx_frames_host(samples, nf)
    a.close()
    b = sc.ModemBank(ns, debug_eq=True)
    first, _ = b.rx_frames_host(samples[:, : 5 * 1880].copy(), 5)
    image = b.state_export()
    b.close()
    assert first.tobytes() == whole[:, :5].tobytes()
    c = sc.ModemBank(ns, debug_eq=True)
    c.state_import(image)
    assert c.call_index == 5
    rest, eq_rest = c.rx_frames_host(samples[:, 5 * 1880:].copy(), nf - 5)
    c.close()
    assert rest.tobytes() == np.ascontiguousarray(whole[:, 5:]).tobytes()
    assert np.array_equal(eq_rest.view(np.uint32), np.ascontiguousarray(eq_whole[:, 5:]).view(np.uint32))
    d = sc.ModemBank(ns + 1)
    with pytest.raises(sc.SingleCarrierError):
        d.state_import(image)                                     # another bank size: refused
    d.close()


def test_fft256_warp_kernel_large_batch(sc):
    """The register / warp-shuffle n=256 kernel (out of place) against the generic shared-memory kernel (taken
    for in-place calls), bit for bit, over a batch larger than the persistent grid; both directions."""
    import torch
    rng = np.random.default_rng(3)
    nb = 148 * 12 * 4 * 2 + 37
    x = (rng.normal(size=(nb, 256)) + 1j * rng.normal(size=(nb, 256))).astype(np.complex64)
    x[5] = 0
    x[6, 1:] = 0
    for inv in (0, 1):
        d_in = torch.from_numpy(x.view(np.float32)).cuda()
        d_out = torch.zeros_like(d_in)
        sc._lib.check(sc.lib.sc_fft_batch_dev(0, nb, 256, inv, d_in.data_ptr(), d_out.data_ptr(), 0))
        d_ip = d_in.clone()
        sc._lib.check(sc.lib.sc_fft_batch_dev(0, nb, 256, inv, d_ip.data_ptr(), d_ip.data_ptr(), 0))
        torch.cuda.synchronize()
        assert torch.equal(d_out.view(torch.int32), d_ip.view(torch.int32))
        y = d_out.cpu().numpy().view(np.complex64)
        want = np.fft.ifft(x[:64].astype(np.complex128), axis=1) * 256 if inv else np.fft.fft(x[:64].astype(np.complex128), axis=1)
        assert np.abs(y[:64] - want).max() <= 1e-5 * np.abs(want).max() * 8


def test_fft_two_streams_do_not_share_scratch(sc):
    """Transforms too long for shared memory use stream-ordered scratch per call: two n=8192 batches queued on two
    CUDA streams at once give the results of running them one after the other."""
    import torch
    rng = np.random.default_rng(8)
    n, nb = 8192, 96
    xa = torch.from_numpy((rng.normal(size=(nb, n, 2))).astype(np.float32)).cuda()
    xb = torch.from_numpy((rng.normal(size=(nb, n, 2))).astype(np.float32)).cuda()
    ya, yb = torch.zeros_like(xa), torch.zeros_like(xb)
    sc._lib.check(sc.lib.sc_fft_batch_dev(0, nb, n, 0, xa.data_ptr(), ya.data_ptr(), 0))
    sc._lib.check(sc.lib.sc_fft_batch_dev(0, nb, n, 0, xb.data_ptr(), yb.data_ptr(), 0))
    torch.cuda.synchronize()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    for _ in range(3):
        za, zb = torch.zeros_like(xa), torch.zeros_like(xb)
        torch.cuda.synchronize()
        sc._lib.check(sc.lib.sc_fft_batch_dev(0, nb, n, 0, xa.data_ptr(), za.data_ptr(), s1.cuda_stream))
        sc._lib.check(sc.lib.sc_fft_batch_dev(0, nb, n, 0, xb.data_ptr(), zb.data_ptr(), s2.cuda_stream))
        torch.cuda.synchronize()
        assert torch.equal(za, ya) and torch.equal(zb, yb)
    sc._lib.check(sc.lib.sc_release_caches())
    zc = torch.zeros_like(xa)
    sc._lib.check(sc.lib.sc_fft_batch_dev(0, nb, n, 0, xa.data_ptr(), zc.data_ptr(), 0))   # tables are rebuilt
    torch.cuda.synchronize()
    assert torch.equal(zc, ya)


def test_lock_stats_field_decoding_stride_and_bin_clamp(sc):
    """lock_stats_kernel against a numpy recount on hand-made records: result_stride > n_frames, matches up to
    128 (bin clamp at 7), all 16 counters."""
    import torch
    from test_multiprocess import counters_from_results
    rng = np.random.default_rng(12)
    ns, nf, stride = 517, 7, 11
    rec = np.zeros((ns, stride), sc.RESULT_DTYPE)
    rec["valid"] = rng.integers(0, 2, (ns, stride))
    rec["matches"] = rng.integers(0, 129, (ns, stride))
    rec["matches"][0, :nf] = 128
    rec["max_index"] = rng.integers(0, 128, (ns, stride))
    rec["rx_timing"] = rng.integers(3, 256, (ns, stride))
    rec["bits"] = rng.integers(0, 2 ** 62, (ns, stride), dtype=np.uint64)
    d = torch.from_numpy(rec.view(np.uint8).reshape(ns, stride * 32)).cuda()
    cnt = torch.zeros(16, dtype=torch.int64, device="cuda")
    sc._lib.check(sc.lib.sc_lock_stats_dev(0, d.data_ptr(), ns, stride, nf, cnt.data_ptr(), 0))
    torch.cuda.synchronize()
    r = rec[:, :nf]
    want = counters_from_results(r["valid"], r["matches"].astype(np.int64), r["max_index"].astype(np.int64),
                                 r["rx_timing"].astype(np.int64), r["bits"])
    assert cnt.cpu().numpy().tolist() == want.tolist()


def test_reduce_stats_through_the_c_abi(sc, oracle):
    """sc_reduce_stats on communicators made by the library's NCCL bridge: one rank (identity) always; with two
    or more GPUs, two banks on devices 0 and 1 reduce to the single-bank counters."""
    import torch
    rng = np.random.default_rng(31)
    ns, nf = 64, 6
    samples = synth_streams(oracle, rng, ns, nf)

    def counters(dev, lo, hi):
        torch.cuda.set_device(dev)
        bank = sc.ModemBank(hi - lo, device=dev)
        d_in = torch.from_numpy(samples[lo:hi]).cuda(dev)
        d_res = torch.zeros((hi - lo, nf * 32), dtype=torch.uint8, device=f"cuda:{dev}")
        bank.rx_frames_dev(d_in, nf, d_res)
        cnt = torch.zeros(16, dtype=torch.int64, device=f"cuda:{dev}")
        bank.lock_stats(d_res, nf, cnt)
        torch.cuda.synchronize(dev)
        bank.close()
        return cnt

    whole = counters(0, 0, ns)
    comm = sc.NcclComm.init_all([0])[0]
    one = whole.clone()
    comm.all_reduce_counters(one)
    torch.cuda.synchronize()
    comm.close()
    assert torch.equal(one, whole) and int(whole[0]) == ns * nf
    if torch.cuda.device_count() >= 2:
        parts = [counters(0, 0, ns // 2), counters(1, ns // 2, ns)]
        comms = sc.NcclComm.init_all([0, 1])
        # one process drives both ranks: the two collective calls must be grouped
        nccl = C.CDLL("libnccl.so.2")
        nccl.ncclGroupStart()
        for dev, (cm, p) in enumerate(zip(comms, parts)):
            torch.cuda.set_device(dev)
            cm.all_reduce_counters(p)
        nccl.ncclGroupEnd()
        for dev in (0, 1):
            torch.cuda.synchronize(dev)
        assert parts[0].cpu().tolist() == whole.cpu().tolist() == parts[1].cpu().tolist()
        for cm in comms:
            cm.close()
        torch.cuda.set_device(0)


def test_tensor_core_proposed_search_equals_direct_search(sc, oracle):
    """sc_preamble_search_batch_dev (tensor-core proposer + exact verifier: tcgen05 + tensor memory + TMA where the
    window layout allows it, mma.sync otherwise; both also by name) against the all-exact kernel, bit for
    bit, on 200k windows built to stress the candidate logic: noise only, planted preambles at every lag, several
    equal maxima (ties), silence, one non-zero symbol, tiny and huge amplitudes, a window repeated with a one-ulp
    change, NaN / Inf samples, and odd window counts / strides."""
    import torch
    pv = np.frombuffer((C.c_int8 * 128).in_dll(oracle.lib, "sco_preamblevalues"), np.int8).astype(np.float32)
    rng = np.random.default_rng(42)
    for ns, stride in ((200001, 255), (777, 260), (200003, 256)):
        sym = (rng.normal(size=(ns, stride)) + 1j * rng.normal(size=(ns, stride))).astype(np.complex64)
        k = min(ns - 500, 4000)
        for s in range(0, k):                                   # planted preambles, every lag, various SNR
            lag = s % 128
            sym[s, lag:lag + 128] += (np.float32(0.2 + (s % 7)) * pv * (1 + 1j)).astype(np.complex64)
        sym[k:k + 50] = 0                                       # silence -> (0, 0.0)
        for s in range(k + 50, k + 100):                        # a single non-zero symbol: many lags tie
            sym[s] = 0
            sym[s, 100 + (s % 100)] = 1 + 1j
        for s in range(k + 100, k + 150):                       # two planted preambles of equal strength, no noise
            sym[s] = 0
            a, b = (s * 3) % 60, 64 + (s * 5) % 60
            sym[s, a:a + 128] += (pv * (1 + 1j)).astype(np.complex64)
            sym[s, b:b + 128] += (pv * (1 + 1j)).astype(np.complex64)
        sym[k + 150:k + 250] *= np.float32(1e-6)                # tiny
        sym[k + 250:k + 350] *= np.float32(3e4)                 # int16-scale and beyond
        sym[k + 350:k + 400] = sym[k + 400:k + 450]             # near duplicates ...
        sym[k + 350:k + 400, 77] = np.nextafter(sym[k + 400:k + 450, 77].real, np.float32(9)) + 1j * sym[k + 400:k + 450, 77].imag
        sym[k + 450, 50] = np.nan                               # a NaN poisons the tensor-core proposal of its whole window
        sym[k + 451, 200] = np.inf + 0j                         # (0 x NaN): those windows must come out of the exact fallback
        sym[k + 452] = np.nan
        sym[k + 453, 255 % stride] = np.nan                     # symbol 255 is read (when it exists) but never used
        d = torch.from_numpy(sym.view(np.float32)).cuda()
        out = {}
        names = ["sc_preamble_search_batch_dev", "sc_preamble_search_fft_batch_dev", "sc_preamble_search_direct_batch_dev",
                 "sc_preamble_search_mma_batch_dev"]
        if stride >= 256 and stride % 2 == 0:                   # the layout the tcgen05 / TMA kernel takes
            names.append("sc_preamble_search_tcgen05_batch_dev")
        for name in names:
            idx = torch.full((ns,), -7, dtype=torch.int32, device="cuda")
            val = torch.full((ns,), -7.0, dtype=torch.float32, device="cuda")
            args = (0, ns, d.data_ptr(), stride, idx.data_ptr(), val.data_ptr())
            sc._lib.check(getattr(sc.lib, name)(*args, None, 0) if "tcgen05" in name else getattr(sc.lib, name)(*args, 0))
            torch.cuda.synchronize()
            out[name] = (idx.cpu().numpy(), val.cpu().numpy())
        for name in names:
            xi, xv = out[name]
            yi, yv = out["sc_preamble_search_direct_batch_dev"]
            bad = np.nonzero((xi != yi) | (xv.view(np.uint32) != yv.view(np.uint32)))[0]
            assert bad.size == 0, (name, bad[:10], xi[bad[:10]], yi[bad[:10]], xv[bad[:10]], yv[bad[:10]])
        (ai, av), (bi, bv) = out["sc_preamble_search_batch_dev"], out["sc_preamble_search_direct_batch_dev"]
        bad = np.nonzero((ai != bi) | (av.view(np.uint32) != bv.view(np.uint32)))[0]
        assert bad.size == 0, (bad[:10], ai[bad[:10]], bi[bad[:10]], av[bad[:10]], bv[bad[:10]])
        fi, fv = out["sc_preamble_search_fft_batch_dev"]
        bad = np.nonzero((fi != bi) | (fv.view(np.uint32) != bv.view(np.uint32)))[0]
        assert bad.size == 0, ("fft proposer", bad[:10], fi[bad[:10]], bi[bad[:10]], fv[bad[:10]], bv[bad[:10]])
        assert (ai[:k] == np.arange(k) % 128)[np.arange(k) % 7 >= 2].mean() > 0.99       # strong preambles are found
        for s in list(range(0, 40)) + list(range(k, k + 150, 7)):                          # and both equal the oracle
            oi, ov = oracle.search(sym[s])
            assert ai[s] == oi and av[s].view(np.uint32) == np.float32(ov).view(np.uint32), s


@pytest.mark.parametrize("mode_name", ["FE_SEARCH_MMA", "FE_SEARCH_TCGEN05"])
def test_frontend_tensor_core_search_mode_is_bit_identical(sc, oracle, mode_name):
    """SC_OPT_FE_SEARCH = SC_FE_SEARCH_MMA / SC_FE_SEARCH_TCGEN05: the fused front-end proposes the preamble search on
    the tensor cores (mma.sync per warp pair; tcgen05.mma / tensor memory per 8-warp CTA) and verifies the candidates
    exactly; every field of every call equals the default (all-exact) mode and the oracle, on noisy loop-back streams,
    silence, dead air with exact zeros (many equal correlations) and an odd bank size."""
    from helpers import compare_results, oracle_results
    from singlecarrier_b200 import modem
    from singlecarrier_b200.modem import OPT_FE_SEARCH
    mode = getattr(modem, mode_name)
    rng = np.random.default_rng(1234)
    ns, nf = 301, 12
    samples = synth_streams(oracle, rng, ns, nf)
    samples[7] = 0                                               # silence: every lag ties at zero
    samples[8, 5000:] = 0
    samples[9] = (rng.integers(-3, 4, samples.shape[1])).astype(np.int16)       # near silence (SURVEY pin 7)
    out = {}
    for md in (0, mode):
        bank = sc.ModemBank(ns, debug_eq=True)
        bank.set_option(OPT_FE_SEARCH, md)
        out[md] = bank.rx_frames_host(samples, nf)
        bank.close()
    assert out[0][0].tobytes() == out[mode][0].tobytes()
    assert out[0][1].tobytes() == out[mode][1].tobytes()
    # the proposer's table is a process-wide cache: releasing it between two batches of a live handle is harmless
    bank = sc.ModemBank(ns, debug_eq=True)
    bank.set_option(OPT_FE_SEARCH, mode)
    a = bank.rx_frames_host(np.ascontiguousarray(samples[:, : 6 * 1880]), 6)
    sc._lib.check(sc.lib.sc_release_caches())
    b = bank.rx_frames_host(np.ascontiguousarray(samples[:, 6 * 1880:]), nf - 6)
    bank.close()
    assert np.concatenate([a[0], b[0]], axis=1).tobytes() == out[0][0].tobytes()
    obits, ostats = oracle_results(oracle, samples, nf)
    assert compare_results(out[mode][0], out[mode][1], obits, ostats) == []


def test_lane_cooperative_tracker_is_bit_identical(sc, oracle):
    """SC_OPT_TRACKER: the 16-lanes-per-stream tracker (the default for small banks) executes the one-thread tracker's
    operations in the same order on other lanes; every field of every call is the same, and equals the oracle's, on
    noisy loop-back streams, silence (exact zeros: the -0 / +0 cases of the gathered sums), near silence, a stream that
    goes dead mid-way, full-scale noise, and a bank size that leaves the last warp half empty."""
    from helpers import compare_results, oracle_results
    from singlecarrier_b200.modem import OPT_TRACKER, TRACKER_COOP, TRACKER_THREAD
    rng = np.random.default_rng(4321)
    ns, nf = 301, 12
    samples = synth_streams(oracle, rng, ns, nf)
    samples[7] = 0
    samples[8, 5000:] = 0
    samples[9] = (rng.integers(-3, 4, samples.shape[1])).astype(np.int16)
    samples[10] = (rng.integers(-32768, 32768, samples.shape[1])).astype(np.int16)
    out = {}
    for mode in (TRACKER_THREAD, TRACKER_COOP):
        bank = sc.ModemBank(ns)
        bank.set_option(OPT_TRACKER, mode)
        out[mode] = bank.rx_frames_host(samples, nf)
        bank.close()
    assert out[TRACKER_THREAD][0].tobytes() == out[TRACKER_COOP][0].tobytes()
    obits, ostats = oracle_results(oracle, samples, nf)
    assert compare_results(out[TRACKER_COOP][0], None, obits, ostats) == []
    assert int(out[TRACKER_COOP][0]["valid"].sum()) > ns          # the locked branch of the data loop ran too


def test_overlapped_chains_are_bit_identical(sc, oracle):
    """SC_OPT_OVERLAP: the tracker cut at qpsk.c:196 and the even / odd calls queued as two chains on two CUDA streams
    (the default for small banks) give every field of every call as the serial call-by-call chain does and as the
    oracle does -- with both trackers, from the host and the device entry point, for batches of 1, 2, 3 and many calls
    (batches of one call cannot overlap and take the serial path, so the two modes also alternate on one handle), on
    noisy loop-back streams, silence, a stream that goes dead, full-scale noise and a ragged last warp."""
    import torch
    from helpers import compare_results, oracle_results
    from singlecarrier_b200.modem import (OPT_OVERLAP, OPT_TRACKER, OVERLAP_OFF, OVERLAP_ON, TRACKER_COOP,
                                          TRACKER_THREAD)
    rng = np.random.default_rng(97531)
    ns, nf = 301, 14
    samples = synth_streams(oracle, rng, ns, nf)
    samples[7] = 0
    samples[8, 5000:] = 0
    samples[9] = (rng.integers(-3, 4, samples.shape[1])).astype(np.int16)
    samples[10] = (rng.integers(-32768, 32768, samples.shape[1])).astype(np.int16)
    obits, ostats = oracle_results(oracle, samples, nf)

    bank = sc.ModemBank(ns)
    bank.set_option(OPT_OVERLAP, OVERLAP_OFF)
    serial = bank.rx_frames_host(samples, nf)
    bank.close()
    assert compare_results(serial[0], None, obits, ostats) == []
    assert int(serial[0]["valid"].sum()) > ns

    dev = torch.from_numpy(samples).cuda()
    for tracker in (TRACKER_THREAD, TRACKER_COOP):
        for splits in ([nf], [1, 2, 3, 1, 7], [2, 2, 2, 2, 2, 2, 2], [5, 9]):
            bank = sc.ModemBank(ns)
            bank.set_option(OPT_OVERLAP, OVERLAP_ON)
            bank.set_option(OPT_TRACKER, tracker)
            parts, f0 = [], 0
            for k in splits:
                parts.append(bank.rx_frames_host(samples[:, f0 * 1880:(f0 + k) * 1880], k)[0])
                f0 += k
            got = np.concatenate(parts, axis=1)
            assert got.tobytes() == serial[0].tobytes(), (tracker, splits)
            bank.close()
        # device entry point, one batch and two
        bank = sc.ModemBank(ns)
        bank.set_option(OPT_OVERLAP, OVERLAP_ON)
        bank.set_option(OPT_TRACKER, tracker)
        res = torch.zeros((ns, nf * 32), dtype=torch.uint8, device="cuda")
        bank.rx_frames_dev(dev, nf, res)
        torch.cuda.synchronize()
        assert res.cpu().numpy().tobytes() == serial[0].tobytes(), tracker
        bank.reset()
        res.zero_()
        bank.rx_frames_dev(dev[:, :6 * 1880], 6, res[:, :6 * 32])
        bank.rx_frames_dev(dev[:, 6 * 1880:], nf - 6, res[:, 6 * 32:])
        torch.cuda.synchronize()
        assert res.cpu().numpy().tobytes() == serial[0].tobytes(), tracker
        bank.close()


def test_overlapped_chains_many_slabs_and_blocks(sc, oracle):
    """The overlapped chains with several slabs per pipe (host entry point: slabs of 4,096 streams round-robin on three
    pipes, each with its own partner and data streams and its own region of the hand-over rings) and with two slabs on
    the device entry point: byte-identical to the serial chain."""
    import torch
    from singlecarrier_b200.modem import OPT_OVERLAP, OVERLAP_OFF, OVERLAP_ON
    rng = np.random.default_rng(2468)
    nf = 7
    base = synth_streams(oracle, rng, 150, nf)
    ns = 17000
    samples = np.empty((ns, base.shape[1]), np.int16)
    for k in range(ns):                                            # shifted, scaled copies: cheap, all different
        samples[k] = np.roll(base[k % 150], 37 * (k // 150)) // (1 + (k // 150) % 3)
    out = {}
    for mode in (OVERLAP_OFF, OVERLAP_ON):
        bank = sc.ModemBank(ns)
        bank.set_option(OPT_OVERLAP, mode)
        a = bank.rx_frames_host(samples[:, :3 * 1880], 3)[0]
        b = bank.rx_frames_host(samples[:, 3 * 1880:], nf - 3)[0]
        bank.reset()
        dev = torch.from_numpy(samples).cuda()
        res = torch.zeros((ns, nf * 32), dtype=torch.uint8, device="cuda")
        bank.rx_frames_dev(dev, nf, res)
        torch.cuda.synchronize()
        out[mode] = (np.concatenate([a, b], axis=1).tobytes(), res.cpu().numpy().tobytes())
        bank.close()
    assert out[OVERLAP_ON][0] == out[OVERLAP_OFF][0]
    assert out[OVERLAP_ON][1] == out[OVERLAP_OFF][1]
    assert out[OVERLAP_ON][0] == out[OVERLAP_ON][1]
