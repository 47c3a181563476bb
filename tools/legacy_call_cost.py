#!/usr/bin/env python
"""Cost of the drop-in single-stream symbols (one kernel launch + small copies per call): microseconds per call
on this box, for INTEGRATION.md.  Usage: python tools/legacy_call_cost.py"""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import singlecarrier_b200 as sc  # noqa: E402

L = sc.lib
L.cnormf.restype = C.c_float
L.train_eq.restype = C.c_float
L.train_eq.argtypes = [C.c_void_p, C.c_int, C.c_float]
L.data_eq.restype = C.c_float
L.data_eq.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
L.scramble.argtypes = [C.c_void_p, C.c_int]
L.fir.argtypes = [C.c_void_p, C.c_bool, C.c_void_p, C.c_int]
L.qpsk_rx_frame.argtypes = [C.c_void_p, C.c_void_p]
L.fft_alloc.restype = C.c_void_p
L.fft_alloc.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p]
L.fft.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]

rng = np.random.default_rng(0)
sym = (rng.normal(size=64) + 1j * rng.normal(size=64)).astype(np.complex64)
mem = np.zeros(49, np.complex64)
x = (rng.normal(size=1880) + 1j * rng.normal(size=1880)).astype(np.complex64)
frame = rng.integers(-3000, 3000, 1880).astype(np.int16)
bits = np.zeros(496, np.uint8)
d = C.c_uint8(0)
cfg = L.fft_alloc(256, 0, None, None)
fi = sym.repeat(4).copy()
fo = np.zeros(256, np.complex64)
L.kalman_init()
L.scramble_init(2)


def cost(name, fn, n=300):
    for _ in range(20):
        fn()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    us = (time.perf_counter() - t0) / n * 1e6
    print(f"| `{name}` | {us:.0f} |")


print("| drop-in call | microseconds per call |\n|---|---|")
cost("train_eq(in, i, ref)", lambda: L.train_eq(sym.ctypes.data, 3, 1.0))
cost("data_eq(&bits, in, i)", lambda: L.data_eq(C.byref(d), sym.ctypes.data, 3))
cost("scramble(&bits, rx)", lambda: L.scramble(C.byref(d), 1))
cost("fir(mem, false, x, 1880)", lambda: L.fir(mem.ctypes.data, False, x.ctypes.data, 1880))
cost("qpsk_rx_frame(in, bits)", lambda: L.qpsk_rx_frame(frame.ctypes.data, bits.ctypes.data))
cost("fft(cfg256, in, out)", lambda: L.fft(cfg, fi.ctypes.data, fo.ctypes.data))
