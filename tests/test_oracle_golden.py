"""CPU: the oracle restatement pinned against the committed reference vectors (tests/golden/, made by
tools/make_golden.py from the reference's own object code) and against the known answers of
SURVEY.md section 4.  Runs anywhere gcc is; no GPU, no /root/reference."""
import ctypes as C
import hashlib

import numpy as np

from oracle import pyoracle as po


def u32(a):
    return np.ascontiguousarray(a).view(np.uint32)


def test_shipped_file_structure(gold):
    """Pin 2: 10 x (640 preamble + 1240 data + 903 zeros), md5 known."""
    x = gold("preamble_qpsk_8k.raw")
    assert x.size == 27830
    assert hashlib.md5(x.tobytes()).hexdigest() == "1175fea4332f8e742524d49641e32c62"
    for k in range(10):
        z = x[1880 + 2783 * k: 1880 + 2783 * k + 903]
        assert (z == 0).all() and x[1880 + 2783 * k - 1] != 0
    assert np.abs(x[:640]).max() == 7837


def test_rx_shipped_file_golden(oracle, gold):
    """Pin 4: qpsk_rx_frame over the shipped file, all 14 calls, every observable."""
    x = gold("preamble_qpsk_8k.raw")
    g = gold("rx_shipped.npz")
    bits, st = oracle.run_stream(x)
    assert st["valid"].tolist() == [1, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 0] == g["valid"].tolist()
    assert np.array_equal(st["max_index"], g["max_index"])
    assert np.array_equal(st["rx_timing"], g["rx_timing"])
    assert np.array_equal(u32(st["max_value"]), u32(g["max_value"]))
    assert np.array_equal(u32(st["eq_coeff"]), u32(g["eq_coeff"]))
    v = g["valid"].astype(bool)
    assert np.array_equal(bits[v], g["bits"][v])
    assert np.array_equal(st["matches"][v], g["matches"][v])
    assert np.array_equal(u32(st["cost"][v]), u32(g["mean"][v]))
    assert st["matches"][12] == 110 and st["max_index"][12] == 120 and st["rx_timing"][12] == 248
    assert f"{st['max_value'][12]:.2f}" == "3472.29" and f"{st['cost'][12]:.2f}" == "15.33"
    assert "".join(map(str, bits[12])) == "00111100000011111010110000011010101001101001101111000000101100"
    assert "".join(map(str, bits[1])) == "11110010010110100010110111011100111011001100101001101010101111"


def test_rx_synth_golden(oracle, gold):
    g = gold("rx_synth.npz")
    for s in range(g["samples"].shape[0]):
        bits, st = oracle.run_stream(g["samples"][s])
        assert np.array_equal(st["valid"], g["valid"][s])
        assert np.array_equal(st["max_index"], g["max_index"][s])
        assert np.array_equal(st["rx_timing"], g["rx_timing"][s])
        assert np.array_equal(u32(st["max_value"]), u32(g["max_value"][s]))
        assert np.array_equal(u32(st["eq_coeff"]), u32(g["eq_coeff"][s]))
        v = g["valid"][s].astype(bool)
        assert np.array_equal(bits[v], g["bits"][s][v])
        assert np.array_equal(st["matches"][v], g["matches"][s][v])


def test_tx_golden(oracle, gold):
    """Pin 1 + reference TX vectors: preamble and data frames, filter memory carried across frames."""
    g = gold("tx_golden.npz")
    st = oracle.new_state()
    parts = []
    for p in range(g["bits"].shape[0]):
        parts.append(oracle.tx_preamble(st))
        for j in range(8):
            parts.append(oracle.tx_data(st, g["bits"][p, j]))
    y = np.concatenate(parts)
    assert np.array_equal(y, g["samples"])
    x = gold("preamble_qpsk_8k.raw")
    assert np.array_equal(y[:640], x[:640])
    assert y[:24].tolist() == [0, -45, -72, -20, -6, -74, -112, -13, -22, -158, -111, -3, -251, -462, -160, 7,
                               -628, -972, -99, 397, -885, -1602, 1044, 4959]


def test_stage_golden(oracle, gold):
    g = gold("stage_golden.npz")
    for wide in (0, 1):
        m, x = g[f"fir_mem_in_{wide}"].copy(), g[f"fir_x_{wide}"].copy()
        oracle.fir(m, bool(wide), x)
        assert np.array_equal(u32(x.view(np.float32)), u32(g[f"fir_y_{wide}"].view(np.float32)))
        assert np.array_equal(u32(m.view(np.float32)), u32(g[f"fir_mem_out_{wide}"].view(np.float32)))
    L = oracle.lib
    k = np.zeros(1024, np.uint8)                                     # sco_kalman
    L.sco_kalman_init(k.ctypes.data)
    sym = g["eq_sym"]
    rets = []
    for i in range(128):
        rets.append(L.sco_train_eq(k.ctypes.data, sym.ctypes.data, i, float(g["eq_ref"][i])))
        c = np.frombuffer(k, np.float32, count=10)
        assert np.array_equal(u32(c), u32(g["eq_traj"][i].view(np.float32))), i
    lfsr = C.c_uint16(0x4A80)
    dibits = []
    for i in range(31):
        d = C.c_uint8(0)
        rets.append(L.sco_data_eq(k.ctypes.data, C.byref(lfsr), C.byref(d), sym.ctypes.data, 128 + i))
        dibits.append(d.value)
    assert np.array_equal(u32(np.array(rets, np.float32)), u32(g["eq_ret"]))
    assert dibits == g["eq_dibits"].tolist()


def test_scrambler_dvb_prbs(oracle, gold):
    """Pin 3: the keystream is the published DVB energy-dispersal PRBS for init 100101010000000."""
    lfsr = C.c_uint16(0x4A80)
    ks = []
    for _ in range(124):
        d = C.c_uint8(0)
        oracle.lib.sco_scramble2(C.byref(d), C.byref(lfsr))
        ks += [d.value & 1, d.value >> 1]
    s = "".join(map(str, ks))
    assert s.startswith("0000001111110110000010000011010000110000101110001010001110010")
    assert ks == gold("stage_golden.npz")["keystream"].tolist()
    # independent re-derivation of 1 + x^14 + x^15
    reg = [1, 0, 0, 1, 0, 1, 0, 1, 0, 0, 0, 0, 0, 0, 0]
    ref = []
    for _ in range(248):
        o = reg[13] ^ reg[14]
        ref.append(o)
        reg = [o] + reg[:-1]
    assert ks == ref


def test_preamble_and_taps_properties(oracle):
    """Pin 9: PN sum +2, autocorrelation peak 128 / max sidelobe 38; taps symmetric, known sums."""
    L = oracle.lib
    pv = np.frombuffer((C.c_int8 * 128).in_dll(L, "sco_preamblevalues"), np.int8).astype(int)
    assert pv.sum() == 2 and set(pv.tolist()) == {-1, 1}
    ac = np.correlate(pv, pv, "full")
    assert ac[127] == 128 and np.abs(np.delete(ac, 127)).max() == 38
    a35 = np.frombuffer((C.c_float * 49).in_dll(L, "sco_alpha35_root"), np.float32)
    a50 = np.frombuffer((C.c_float * 49).in_dll(L, "sco_alpha50_root"), np.float32)
    for t, total in ((a35, 0.9930), (a50, 1.0034)):
        assert np.array_equal(t, t[::-1]) and abs(float(t.astype(np.float64).sum()) - total) < 1e-4


def test_silence_is_detected_as_valid(oracle):
    """F6: on silence every call is 'valid', decodes the raw keystream and sets rx_timing = 128."""
    bits, st = oracle.run_stream(np.zeros(1880 * 5, np.int16))
    assert st["valid"].all() and (st["matches"] == 128).all() and (st["rx_timing"] == 128).all()
    assert "".join(map(str, bits[0])).startswith("00000011111101100000100000110100")


def test_nco_rect_and_hypot_model(oracle):
    """The device renormalisation uses (float)sqrt((double)x*x + (double)y*y); check it equals glibc's
    cabsf on the phasors the recurrence visits (SURVEY section 4 pin 6) and pin the NCO step."""
    r = oracle.nco_rect(-1100.0)
    x = np.float32(2 * np.pi * -1100.0 / 8000.0)
    assert f"{x:.9f}" == "-0.863937974"                      # the NCO step, SURVEY section 8c
    assert abs(float(r.real) - np.cos(np.float64(x))) < 6e-8 and abs(float(r.imag) - np.sin(np.float64(x))) < 6e-8
    ph = np.complex64(1)
    libm = C.CDLL("libm.so.6")
    libm.cabsf.restype = C.c_float

    class Cf(C.Structure):
        _fields_ = [("r", C.c_float), ("i", C.c_float)]
    libm.hypotf.restype = C.c_float
    libm.hypotf.argtypes = [C.c_float, C.c_float]
    rr, ri = np.float32(r.real), np.float32(r.imag)
    pr, pi = np.float32(1), np.float32(0)
    for n in range(30000):
        pr, pi = np.float32(np.float32(pr * rr) - np.float32(pi * ri)), np.float32(np.float32(pr * ri) + np.float32(pi * rr))
        if n % 1880 == 1879 or n % 97 == 0:
            model = np.float32(np.sqrt(np.float64(pr) * np.float64(pr) + np.float64(pi) * np.float64(pi)))
            assert model == np.float32(libm.hypotf(float(pr), float(pi)))
        if n % 1880 == 1879:
            m = np.float32(libm.hypotf(float(pr), float(pi)))
            pr, pi = np.float32(pr / m), np.float32(pi / m)
    del ph
