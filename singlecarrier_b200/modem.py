"""Host mirror of the reference's modem interface, batched over a bank of streams.

``ModemBank.rx_frames`` is N lock-step copies of the reference's ``while(1) { fread;
qpsk_rx_frame(); fwrite }`` loop (/root/reference/src/qpsk.c:436-458); ``ModemBank.tx_packets`` is
N copies of its transmit loop (qpsk.c:380-413).  Everything is computed by the CUDA kernels behind
``libsinglecarrier_b200.so``; this file only moves pointers.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from ._lib import SC_FLAG_DEBUG_EQ, SC_FLAG_WIDE, Channel, check, lib

FRAME_SIZE = 1880
SYMBOLS_PER_FRAME = 376
BITS_PER_CALL = 62
PACKET_SAMPLES = 1880
OPT_SLAB_PARTS, OPT_PROFILE = 1, 2
REFERENCE_GAP = 903         # dead air between packets in the reference's main(), qpsk.c:410-412

# sc_frame_result, include/singlecarrier_b200.h
RESULT_DTYPE = np.dtype([("bits", "<u8"), ("max_value", "<f4"), ("cost", "<f4"), ("max_index", "<i2"),
                         ("matches", "<i2"), ("rx_timing", "<i2"), ("valid", "u1"), ("reserved0", "u1"),
                         ("call_index", "<u4"), ("reserved1", "<u4")])
assert RESULT_DTYPE.itemsize == 32


def keystream_word(call_index: int) -> int:
    return int(lib.sc_keystream_word(call_index))


def launch_count() -> int:
    return int(lib.sc_launch_count())


def unpack_bits(results: np.ndarray, rows: Optional[np.ndarray] = None) -> np.ndarray:
    """Results -> the reference's bit-per-byte rows (62 per VALID call); invalid rows stay 255."""
    r = np.ascontiguousarray(results)
    assert r.dtype == RESULT_DTYPE
    if rows is None:
        rows = np.full(r.shape + (BITS_PER_CALL,), 255, np.uint8)
    lib.sc_unpack_bits(r.ctypes.data, r.size, rows.ctypes.data)
    return rows


def _ptr(t) -> int:
    """Device/host pointer of a torch tensor or numpy array (0 for None)."""
    if t is None:
        return 0
    if isinstance(t, np.ndarray):
        return t.ctypes.data
    return t.data_ptr()


class ModemBank:
    """A bank of ``n_streams`` synchronized modems on one GPU (sc_modem handle)."""

    def __init__(self, n_streams: int, device: int = 0, wide: bool = False, foffset_hz: float = 0.0,
                 debug_eq: bool = False):
        self._h = C.c_void_p()
        flags = (SC_FLAG_WIDE if wide else 0) | (SC_FLAG_DEBUG_EQ if debug_eq else 0)
        check(lib.sc_create(C.byref(self._h), device, n_streams, flags, foffset_hz))
        self.n_streams = n_streams
        self.device = device
        self.debug_eq = debug_eq

    @staticmethod
    def result_dtype() -> np.dtype:
        return RESULT_DTYPE

    def close(self) -> None:
        if self._h:
            lib.sc_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self) -> None:
        check(lib.sc_reset(self._h))

    @property
    def call_index(self) -> int:
        return int(lib.sc_call_index(self._h))

    def set_option(self, option: int, value: int) -> None:
        check(lib.sc_set_option(self._h, option, value))

    def profile_read(self) -> dict:
        out = (C.c_double * 4)()
        check(lib.sc_profile_read(self._h, out))
        return {"frontend_ms": out[0], "frontend_launches": int(out[1]), "track_ms": out[2], "track_launches": int(out[3])}

    def lock_stats(self, results, n_frames: int, counters, stream: int = 0) -> None:
        """Accumulate the 16 lock/bit counters of results (CUDA uint8 [n, >= n_frames*32]) into counters (CUDA int64[16])."""
        check(lib.sc_lock_stats_dev(self.device, results.data_ptr(), self.n_streams, results.stride(0) // 32, n_frames,
                                    counters.data_ptr(), stream))

    # ---- RX ------------------------------------------------------------------------------------
    def rx_frames_host(self, samples: np.ndarray, n_frames: Optional[int] = None, results: Optional[np.ndarray] = None):
        """samples: host int16 [n_streams, >= n_frames*1880] (numpy, or a pinned torch tensor's
        numpy view).  Returns (results[n_streams, n_frames], eq_coeff or None)."""
        assert samples.dtype == np.int16 and samples.ndim == 2 and samples.shape[0] == self.n_streams
        assert samples.strides[1] == 2
        stride = samples.strides[0] // 2
        if n_frames is None:
            n_frames = samples.shape[1] // FRAME_SIZE
        if results is None:
            results = np.zeros((self.n_streams, n_frames), RESULT_DTYPE)
        eq = np.zeros((self.n_streams, n_frames, 10), np.float32) if self.debug_eq else None
        check(lib.sc_rx_frames_host(self._h, samples.ctypes.data, stride, n_frames, results.ctypes.data,
                                    results.strides[0] // 32, _ptr(eq)))
        return results, eq

    def rx_frames_dev(self, samples, n_frames: int, results, eq_dbg=None, stream: int = 0) -> None:
        """samples: CUDA int16 tensor [n_streams, stride]; results: CUDA uint8 tensor
        [n_streams, n_frames*32] (view as RESULT_DTYPE after copying back)."""
        check(lib.sc_rx_frames_dev(self._h, samples.data_ptr(), samples.stride(0), n_frames,
                                   results.data_ptr(), results.stride(0) // 32, _ptr(eq_dbg), stream))

    # ---- TX ------------------------------------------------------------------------------------
    def tx_packets_dev(self, out, n_packets: int, gap_samples: int = REFERENCE_GAP, bits=None, bits_out=None,
                       seed: int = 0, lead_in=None, channel: Optional[dict] = None, stream: int = 0) -> None:
        """out: CUDA int16 tensor [n_streams, samples_per_stream]."""
        if channel is None:
            check(lib.sc_tx_packets_dev(self._h, _ptr(bits), _ptr(bits_out), seed, n_packets, gap_samples,
                                        _ptr(lead_in), out.data_ptr(), out.stride(0), out.shape[1], stream))
        else:
            ch = Channel(*[_ptr(channel.get(k)) for k in ("df_hz", "phi_rad", "drift_hz_s", "sigma_lsb",
                                                           "echo_amp", "echo_theta", "echo_delay")])
            check(lib.sc_tx_channel_dev(self._h, _ptr(bits), _ptr(bits_out), seed, n_packets, gap_samples,
                                        _ptr(lead_in), C.byref(ch), out.data_ptr(), out.stride(0), out.shape[1],
                                        stream))

    def nco_table(self, first_call: int, n_calls: int, tx: bool = False) -> np.ndarray:
        out = np.zeros((n_calls, FRAME_SIZE), np.complex64)
        check(lib.sc_nco_table_host(self._h, int(tx), first_call, n_calls, out.ctypes.data))
        return out
