"""GPU: PACKET MODE, an extension with no counterpart in the reference (full-packet decode of all
8 x 31 data symbols and TX scrambling, both TODOs at /root/reference/src/qpsk.c:206-215,386,397).  The written
specification is oracle/sc_oracle_ext.c, itself built from the pinned primitives of the restatement; the CUDA path
must equal it bit for bit, and must leave everything the reference DOES compute untouched."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sc():
    import singlecarrier_b200 as m
    assert m.lib.sc_device_count() > 0
    return m


def gpu_packet_streams(sc, ns, n_packets, seed, sigma_scale=300.0, gap=903):
    """Loop-back streams from the library's own packet-mode TX (scrambled) + channel."""
    import torch
    bank = sc.ModemBank(ns, packet=True)
    g = torch.Generator(device="cuda").manual_seed(seed)
    lead = torch.randint(0, 2783, (ns,), generator=g, device="cuda", dtype=torch.int32)
    df = (torch.rand(ns, generator=g, device="cuda") * 10 - 5).float()
    sigma = (torch.arange(ns, device="cuda") % 4).float() * sigma_scale
    total = ((n_packets * (1880 + gap) + 2783 + 2 * 1880) // 1880) * 1880
    d_in = torch.empty((ns, total), dtype=torch.int16, device="cuda")
    bits = torch.empty((ns, n_packets, 8, 62), dtype=torch.uint8, device="cuda")
    bank.tx_packets_dev(d_in, n_packets, gap_samples=gap, seed=seed, bits_out=bits, lead_in=lead,
                        channel={"df_hz": df, "sigma_lsb": sigma})
    torch.cuda.synchronize()
    bank.close()
    return d_in.cpu().numpy(), bits.cpu().numpy(), lead.cpu().numpy()


def test_packet_tx_scrambles_like_the_spec(sc, oracle):
    """sc_tx_packets_dev of a SC_FLAG_PACKET bank == the spec's transmitter (scramble_init(tx) per packet and
    scramble(&sdata, tx) per dibit enabled), sample for sample."""
    import torch
    from oracle import pyoracle as po
    rng = np.random.default_rng(5)
    ns, npk = 3, 4
    bits = rng.integers(0, 2, (ns, npk, 8, 62)).astype(np.uint8)
    bank = sc.ModemBank(ns, packet=True)
    out = torch.zeros((ns, npk * 2783), dtype=torch.int16, device="cuda")
    bank.tx_packets_dev(out, npk, gap_samples=903, bits=torch.from_numpy(bits).cuda())
    torch.cuda.synchronize()
    bank.close()
    got = out.cpu().numpy()
    for s in range(ns):
        st = oracle.new_state()
        for p in range(npk):
            want = po.ext_tx_packet(oracle, st, bits[s, p].reshape(-1))
            assert np.array_equal(got[s, p * 2783:p * 2783 + 1880], want), (s, p)
            assert not got[s, p * 2783 + 1880:(p + 1) * 2783].any()
    # and it differs from the unscrambled transmitter
    plain = sc.ModemBank(ns)
    out2 = torch.zeros_like(out)
    plain.tx_packets_dev(out2, npk, gap_samples=903, bits=torch.from_numpy(bits).cuda())
    torch.cuda.synchronize()
    plain.close()
    assert not np.array_equal(out2.cpu().numpy(), got)
    assert np.array_equal(out2.cpu().numpy()[:, :640], got[:, :640])            # the preamble is sent unscrambled


@pytest.mark.parametrize("ns,npk,seed", [(1, 3, 1), (70, 5, 2), (300, 4, 3)])
def test_packet_rx_equals_the_spec(sc, oracle, ns, npk, seed):
    from oracle import pyoracle as po
    samples, txbits, lead = gpu_packet_streams(sc, ns, npk, seed)
    nf = samples.shape[1] // 1880
    bank = sc.ModemBank(ns, packet=True)
    res, pk, n = bank.rx_packets_host(samples, nf)
    bank.close()
    plain = sc.ModemBank(ns)
    res0, _ = plain.rx_frames_host(samples, nf)
    plain.close()
    assert res.tobytes() == res0.tobytes()                                      # nothing of the ordinary result moves
    assert n == len(pk) == int(res["valid"][:, 2:].sum())
    rows = sc.unpack_packet_bits(pk)
    k = 0
    errs = bits = 0
    for s in range(ns):
        obits, ostats, opk = po.ext_run_stream(oracle, samples[s])
        assert np.array_equal(ostats["valid"], res["valid"][s].astype(np.int32))
        for q in opk:
            p = pk[k]
            assert (p["stream"], p["call_index"], p["max_index"], p["matches"]) == (s, q["call"], q["max_index"], q["matches"]), (s, q["call"])
            assert np.float32(p["cost"]).view(np.uint32) == np.float32(q["cost"]).view(np.uint32), (s, q["call"])
            assert np.array_equal(rows[k].reshape(-1), q["bits"]), (s, q["call"])
            assert p["matches"] == res["matches"][s, q["call"]]                 # the packet's training pass is the call's
            # bit errors of packets found where one was sent (48 = the two RRC group delays)
            pos = (q["call"] - 2) * 1880 + 5 * q["max_index"] + q["timing"] - 48 - lead[s]
            j = int(round(pos / 2783))
            if 0 <= j < npk and abs(pos - j * 2783) <= 10:
                errs += int((rows[k] != txbits[s, j]).sum())
                bits += 496
            k += 1
    assert k == len(pk)
    if ns >= 70:
        assert bits > 496 * ns // 2
        print(f"packet mode, {ns} streams: {bits // 496} aligned packets, BER {errs / bits:.3f}")


def test_packet_mode_streaming_equals_one_shot(sc, oracle):
    """Frames delivered 1, 2, 3, 5 ... at a time (packets whose frames straddle API calls use the two frames and the
    rx_timing snapshots kept by the handle) == one call with all frames."""
    ns, npk = 97, 5
    samples, _, _ = gpu_packet_streams(sc, ns, npk, 11)
    nf = samples.shape[1] // 1880
    a = sc.ModemBank(ns, packet=True)
    res, pk, n = a.rx_packets_host(samples, nf)
    a.close()
    b = sc.ModemBank(ns, packet=True)
    got_r, got_p = [], []
    f = 0
    for step in [1, 2, 1, 3, 1, 1, 5, 2, 100]:
        k = min(step, nf - f)
        if k <= 0:
            break
        r, p, _ = b.rx_packets_host(np.ascontiguousarray(samples[:, f * 1880:(f + k) * 1880]), k)
        got_r.append(r)
        got_p.append(p)
        f += k
    b.close()
    assert np.concatenate(got_r, axis=1).tobytes() == res.tobytes()
    allp = np.concatenate(got_p)
    allp = allp[np.lexsort((allp["call_index"], allp["stream"]))]
    assert allp.tobytes() == pk.tobytes() and n == len(pk) > ns


def test_packet_mode_device_entry_and_capacity(sc):
    import torch
    ns, npk = 4000, 3
    samples, _, _ = gpu_packet_streams(sc, ns, npk, 21)
    nf = samples.shape[1] // 1880
    host = sc.ModemBank(ns, packet=True)
    res, pk, n = host.rx_packets_host(samples, nf)
    host.close()
    bank = sc.ModemBank(ns, packet=True)
    d_in = torch.from_numpy(samples).cuda()
    d_res = torch.zeros((ns, nf * 32), dtype=torch.uint8, device="cuda")
    cap = n // 2                                                                 # too small on purpose
    d_pk = torch.zeros(cap * 96, dtype=torch.uint8, device="cuda")
    d_n = torch.zeros(1, dtype=torch.int64, device="cuda")
    bank.rx_packets_dev(d_in, nf, d_res, d_pk, d_n)
    torch.cuda.synchronize()
    bank.close()
    assert int(d_n.item()) == n                                                  # all are counted, the surplus is dropped
    assert d_res.cpu().numpy().tobytes() == res.tobytes()
    got = d_pk.cpu().numpy().view(sc.PACKET_DTYPE)
    key = {(int(p["stream"]), int(p["call_index"])): p.tobytes() for p in pk}
    assert all(key[(int(p["stream"]), int(p["call_index"]))] == p.tobytes() for p in got)
    no = sc.ModemBank(8)
    with pytest.raises(sc.SingleCarrierError):
        no.rx_packets_host(np.zeros((8, 1880), np.int16), 1)                     # needs SC_FLAG_PACKET
    no.close()
