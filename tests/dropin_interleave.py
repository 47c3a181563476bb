"""Driver for tests/test_dropin_gpu.py::test_interleaved_globals_match_reference (test infrastructure).

Runs ONE fixed sequence of the reference's public symbols -- qpsk_rx_frame() interleaved with
scramble(&x, rx), scramble_init(rx) and train_eq()/data_eq() that continue from the state the frame left
behind -- against a shared library given on the command line, and prints one JSON line per step.  The test
runs it once on oracle/_ref/libsc_ref.so (the reference's own objects) and once on
libsinglecarrier_b200.so (a fresh process each: both keep their state in process globals) and diffs.

usage: python dropin_interleave.py LIB.so SAMPLES.raw {ref|ours}
"""
import ctypes as C
import json
import sys

import numpy as np

lib = C.CDLL(sys.argv[1])
x = np.fromfile(sys.argv[2], dtype="<i2")
kind = sys.argv[3]

lib.train_eq.restype = C.c_float
lib.train_eq.argtypes = [C.c_void_p, C.c_int, C.c_float]
lib.data_eq.restype = C.c_float
lib.data_eq.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
lib.scramble_init.argtypes = [C.c_int]
lib.scramble.argtypes = [C.c_void_p, C.c_int]
lib.qpsk_rx_frame.argtypes = [C.c_void_p, C.c_void_p]
lib.qpsk_rx_frame.restype = C.c_int

if kind == "ref":
    lib.ref_reset.argtypes = [C.c_int]
    lib.ref_reset(0)                  # main()'s start-up block (statics of qpsk.c are not reachable otherwise)
else:
    lib.kalman_init()                 # the same start-up through the public symbols
    lib.scramble_init(2)
    lib.scramble_init(1)

eq = np.frombuffer((C.c_float * 10).in_dll(lib, "eq_coeff"), np.float32)
gain = np.frombuffer((C.c_float * 10).in_dll(lib, "kalman_gain"), np.float32)
ky = C.c_float.in_dll(lib, "kalman_y")


def snap(tag, **kw):
    kw.update(tag=tag, eq=eq.view(np.uint32).tolist(), gain=gain.view(np.uint32).tolist(),
              ky=int(np.float32(ky.value).view(np.uint32)))
    print(json.dumps(kw))


rng = np.random.default_rng(99)
sym = (rng.normal(size=64) + 1j * rng.normal(size=64)).astype(np.complex64)
nf = x.size // 1880
bits = np.zeros(496, np.uint8)
for n in range(nf):
    frame = np.ascontiguousarray(x[n * 1880:(n + 1) * 1880])
    bits[:] = 255
    v = lib.qpsk_rx_frame(frame.ctypes.data, bits.ctypes.data)
    snap("rx", call=n, valid=int(v), bits=bits[:62].tolist() if v else [])
    if n == 1:                        # advance RXMemory by hand: later frames must see the shifted keystream
        out = []
        for _ in range(3):
            d = C.c_uint8(0)
            lib.scramble(C.byref(d), 1)
            out.append(d.value)
        snap("scramble", dibits=out)
    if n == 5:
        lib.scramble_init(1)          # ... and a re-seed
        snap("scramble_init")
    if n in (3, 12):                  # continue from the equalizer state the frame left (no kalman_reset)
        r1 = lib.train_eq(sym.ctypes.data, 0, 1.0)
        d = C.c_uint8(0)
        r2 = lib.data_eq(C.byref(d), sym.ctypes.data, 7)
        snap("eq_continue", r=[int(np.float32(r1).view(np.uint32)), int(np.float32(r2).view(np.uint32))], dibit=d.value)
