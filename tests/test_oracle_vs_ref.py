"""CPU: the restatement against the REFERENCE's own object code on fresh random inputs.  Only runs where
oracle/_ref/libsc_ref.so exists (built from /root/reference by oracle/Makefile; it travels to the GPU
box as a prebuilt file)."""
import numpy as np

from helpers import synth_streams


def same(a, b):
    return np.array_equal(np.ascontiguousarray(a).view(np.uint32), np.ascontiguousarray(b).view(np.uint32))


def test_rx_random_streams(oracle, ref):
    rng = np.random.default_rng(99)
    samples = synth_streams(oracle, rng, 24, 10)
    noise = [rng.integers(-a, a + 1, 1880 * 8).astype(np.int16) for a in (3, 100, 3000, 32767)]
    for x in list(samples) + noise:
        b1, s1 = oracle.run_stream(x)
        b2, s2 = ref.run_stream(x)
        assert np.array_equal(s1["valid"], s2["valid"]) and np.array_equal(s1["max_index"], s2["max_index"])
        assert np.array_equal(s1["rx_timing"], s2["rx_timing"])
        assert same(s1["max_value"], s2["max_value"]) and same(s1["eq_coeff"], s2["eq_coeff"])
        v = s2["valid"].astype(bool)
        assert np.array_equal(b1[v], b2[v]) and np.array_equal(s1["matches"][v], s2["matches"][v])
        assert same(s1["cost"][v], s2["mean"][v])


def test_wide_filter(oracle, ref):
    rng = np.random.default_rng(5)
    x = synth_streams(oracle, rng, 3, 8)
    for s in range(3):
        b1, s1 = oracle.run_stream(x[s], wide=True)
        b2, s2 = ref.run_stream(x[s], wide=True)
        assert np.array_equal(s1["valid"], s2["valid"]) and same(s1["eq_coeff"], s2["eq_coeff"])


def test_stage_taps_filtered_and_decimated(oracle, ref):
    """Filtered samples and decimated symbols after a call: bit-identical."""
    rng = np.random.default_rng(6)
    x = synth_streams(oracle, rng, 1, 6)[0]
    ref.reset()
    st = oracle.new_state()
    for n in range(6):
        fr = x[n * 1880:(n + 1) * 1880]
        ref.rx_frame(fr)
        oracle.rx_frame(st, fr)
        mine = np.frombuffer(st, np.float32, count=2 * 3760).view(np.complex64)
        dec = np.frombuffer(st, np.float32, count=2 * 752, offset=3760 * 8).view(np.complex64)
        assert same(mine[:1880].view(np.float32), ref.filtered().view(np.float32))
        assert same(dec.view(np.float32), ref.decimated().view(np.float32))


def test_tx_random(oracle, ref):
    rng = np.random.default_rng(7)
    st = oracle.new_state()
    ref.reset()
    for _ in range(5):
        assert np.array_equal(oracle.tx_preamble(st), ref.tx_preamble())
        for _ in range(8):
            b = rng.integers(0, 2, 62).astype(np.uint8)
            assert np.array_equal(oracle.tx_data(st, b), ref.tx_data(b))
