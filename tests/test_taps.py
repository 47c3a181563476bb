"""CPU: the root-Nyquist tap generator (port of the reference's Octave tool) reproduces both C tables."""
import ctypes as C

import numpy as np


def test_generator_reproduces_reference_tables(oracle):
    from singlecarrier_b200.taps import gen_rn_coeffs, modem_taps
    for name, alpha in (("sco_alpha35_root", 0.35), ("sco_alpha50_root", 0.5)):
        table = np.frombuffer((C.c_float * 49).in_dll(oracle.lib, name), np.float32)
        h = gen_rn_coeffs(alpha, 1.0 / 8000.0, 1600.0, 10, 5)
        assert h.size == 50
        assert np.abs(h[1:50] - table.astype(np.float64)).max() < 1e-8        # the tables carry 8 decimals
        assert np.abs(modem_taps(alpha) - table).max() < 1e-8
    h31 = modem_taps(0.31)                                                     # octave/test_filter.m uses alpha = 0.31
    assert np.allclose(h31, h31[::-1], atol=1e-7) and abs(float(h31.sum()) - 1.0) < 0.02
