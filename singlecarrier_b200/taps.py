"""Root-Nyquist tap generator.

Port of the reference's offline Octave tool ``octave/gen_rn_coeffs.m`` (David Rowe, 2012): raised-cosine
pulse -> 4096-point FFT -> square root of the magnitude (stop band pushed down) -> inverse FFT.  The
reference's C tables (``src/constants.c:49-156``) are ``gen_rn_coeffs(alpha, 1/8000, 1600, 10, 5)``
with the first of the 50 returned values dropped; ``tests/test_taps.py`` checks that this port
reproduces both tables to their printed precision.  Host-side tooling, like the original: the
kernels take their taps at compile time from ``include/sc_tables.inc``.
"""
from __future__ import annotations

import numpy as np


def gen_rn_coeffs(alpha: float, T: float, Rs: float, Nsym: int, M: int) -> np.ndarray:
    """gen_rn_coeffs.m:7-40.  Returns Nsym*M coefficients (float64)."""
    Ts = 1.0 / Rs
    n = np.arange(-Nsym * M // 2, Nsym * M // 2 + 1) * T          # -Nsym*Ts/2 : T : Nsym*Ts/2
    nfilter = Nsym * M
    sinc_den = np.pi * n / Ts
    with np.errstate(divide="ignore", invalid="ignore"):
        sinc_op = np.sin(np.pi * n / Ts) / sinc_den
        cos_den = 1.0 - (2.0 * alpha * n / Ts) ** 2
        cos_op = np.cos(alpha * np.pi * n / Ts) / cos_den
    sinc_op[np.abs(sinc_den) < 1e-10] = 1.0
    cos_op[np.abs(cos_den) < 1e-10] = np.pi / 4.0
    gt = sinc_op * cos_op
    nfft = 4096
    gf = np.fft.fft(gt, nfft) / M
    gf = np.where(np.abs(gf) < 0.02, gf * 0.001, gf)             # "pushes the stop band down again"
    root = np.sqrt(np.abs(gf)) * np.exp(1j * np.angle(gf))
    return np.real(np.fft.ifft(root)[:nfilter])


def modem_taps(alpha: float) -> np.ndarray:
    """The 49 taps the modem uses for a given roll-off (alpha=0.35 and 0.5 give the reference's tables)."""
    return gen_rn_coeffs(alpha, 1.0 / 8000.0, 1600.0, 10, 5)[1:50].astype(np.float32)
