"""GPU parity for the TX path and the loop-back generator (row f-1)."""
import numpy as np
import pytest

from helpers import compare_results, oracle_results

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sc():
    import singlecarrier_b200 as m
    assert m.lib.sc_device_count() > 0
    return m


def gpu_tx(sc, bits, n_packets, gap, lead=None, total=None, wide=False):
    import torch
    ns = bits.shape[0]
    total = total or n_packets * (1880 + gap) + (int(max(lead)) if lead is not None else 0)
    bank = sc.ModemBank(ns, wide=wide)
    out = torch.full((ns, total), -7, dtype=torch.int16, device="cuda")
    d_bits = torch.from_numpy(bits).cuda()
    d_lead = torch.from_numpy(np.asarray(lead, np.int32)).cuda() if lead is not None else None
    bank.tx_packets_dev(out, n_packets, gap_samples=gap, bits=d_bits, lead_in=d_lead)
    torch.cuda.synchronize()
    bank.close()
    return out.cpu().numpy()


def oracle_tx(oracle, bits, gap, lead, total, wide=False):
    st = oracle.new_state(wide=wide)
    parts = [np.zeros(lead, np.int16)]
    for p in range(bits.shape[0]):
        parts.append(oracle.tx_preamble(st))
        for j in range(8):
            parts.append(oracle.tx_data(st, bits[p, j]))
        parts.append(np.zeros(gap, np.int16))
    x = np.concatenate(parts)
    out = np.zeros(total, np.int16)
    n = min(total, x.size)
    out[:n] = x[:n]
    return out


def test_tx_golden_reference_vectors(sc, gold):
    g = gold("tx_golden.npz")
    bits = g["bits"][None]                                    # [1, 3, 8, 62]
    y = gpu_tx(sc, bits, 3, 0)
    assert np.array_equal(y[0], g["samples"])


def test_tx_shipped_file_preamble(sc, gold):
    """SURVEY section 4 pin 1: the first 640 samples of the shipped file are the cold-start preamble."""
    x = gold("preamble_qpsk_8k.raw")
    y = gpu_tx(sc, np.zeros((1, 1, 8, 62), np.uint8), 1, 903)
    assert np.array_equal(y[0, :640], x[:640])
    assert y[0, :24].tolist() == [0, -45, -72, -20, -6, -74, -112, -13, -22, -158, -111, -3, -251, -462, -160, 7,
                                  -628, -972, -99, 397, -885, -1602, 1044, 4959]
    assert (y[0, 1880:1880 + 903] == 0).all()


@pytest.mark.parametrize("wide", [False, True])
def test_tx_vs_oracle_lead_gap_tail(sc, oracle, wide):
    rng = np.random.default_rng(5)
    ns, npk, gap = 9, 4, 903
    bits = rng.integers(0, 2, (ns, npk, 8, 62)).astype(np.uint8)
    lead = rng.integers(0, 3000, ns)
    total = 12000
    y = gpu_tx(sc, bits, npk, gap, lead=lead, total=total, wide=wide)
    for s in range(ns):
        assert np.array_equal(y[s], oracle_tx(oracle, bits[s], gap, int(lead[s]), total, wide=wide)), s


def test_generated_bits_roundtrip_and_loopback_config2(sc, oracle):
    """Config 2 in miniature: device-generated packets + frequency/phase offsets, no noise; GPU RX of the
    generated int16 must equal the oracle's RX of the same int16, field for field."""
    import torch
    ns, nf = 96, 11
    total = nf * 1880
    bank = sc.ModemBank(ns, debug_eq=True)
    g = torch.Generator(device="cuda").manual_seed(3)
    lead = (80 + 5 * torch.randint(0, 101, (ns,), generator=g, device="cuda")).int()
    df = (torch.rand(ns, generator=g, device="cuda") * 40 - 20).float()
    phi = (torch.rand(ns, generator=g, device="cuda") * 6.2831853).float()
    out = torch.empty((ns, total), dtype=torch.int16, device="cuda")
    bits_out = torch.empty((ns, 8, 8, 62), dtype=torch.uint8, device="cuda")
    bank.tx_packets_dev(out, 8, gap_samples=0, seed=1234, bits_out=bits_out, lead_in=lead,
                        channel={"df_hz": df, "phi_rad": phi})
    torch.cuda.synchronize()
    samples = out.cpu().numpy()
    b = bits_out.cpu().numpy()
    assert set(np.unique(b)) == {0, 1} and 0.45 < b.mean() < 0.55
    assert np.abs(samples).max() > 4000
    res, eq = bank.rx_frames_host(samples, nf)
    bank.close()
    obits, ostats = oracle_results(oracle, samples, nf)
    assert compare_results(res, eq, obits, ostats) == []
    assert res["valid"][:, 2:].sum() > ns            # preambles are found


def test_channel_noise_statistics(sc):
    import torch
    ns, total = 4, 200000
    bank = sc.ModemBank(ns)
    sigma = torch.tensor([0.0, 100.0, 1000.0, 3000.0], device="cuda")
    out = torch.empty((ns, total), dtype=torch.int16, device="cuda")
    bank.tx_packets_dev(out, 0, seed=9, channel={"sigma_lsb": sigma})          # no packets: noise only
    torch.cuda.synchronize()
    x = out.float().cpu().numpy()
    bank.close()
    assert (x[0] == 0).all()
    for k in (1, 2, 3):
        assert abs(x[k].std() / float(sigma[k]) - 1.0) < 0.03
        assert abs(x[k].mean()) < 0.02 * float(sigma[k])
