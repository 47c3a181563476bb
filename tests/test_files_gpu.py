"""GPU: the reference's file formats either side of the RX path."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_demodulate_files_matches_reference_output_format(tmp_path, oracle, gold):
    from singlecarrier_b200 import files
    raw = os.path.join(ROOT, "tests", "golden", "preamble_qpsk_8k.raw")
    x = gold("preamble_qpsk_8k.raw")
    short = tmp_path / "short.raw"                       # a shorter file in the same bank: 7 calls + 100 stray samples
    x[: 7 * 1880 + 100].tofile(short)
    noisy = tmp_path / "noisy.raw"
    rng = np.random.default_rng(2)
    y = np.clip(x.astype(np.float64) + rng.normal(0, 200, x.size), -32767, 32767).round().astype("<i2")
    y.tofile(noisy)
    outs = [str(tmp_path / f"bits{k}.bin") for k in range(3)]
    res = files.demodulate_files([raw, str(short), str(noisy)], outs)
    assert [r.shape[0] for r in res] == [14, 7, 14]
    for k, samples in enumerate((x, x[: 7 * 1880], y)):
        obits, ost = oracle.run_stream(samples)
        assert np.array_equal(res[k]["valid"].astype(np.int32), ost["valid"])
        want = bytearray()
        for n in range(ost.shape[0]):
            if ost["valid"][n]:
                want += obits[n].tobytes() + bytes(496 - 62)
        assert open(outs[k], "rb").read() == bytes(want)
    g = gold("rx_shipped.npz")
    assert os.path.getsize(outs[0]) == 496 * int(g["valid"].sum()) == 496 * 3
