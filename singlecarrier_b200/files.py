"""The reference's on-disk formats around the RX path.

Input: raw little-endian int16 mono 8 kHz sample files, read 1880 samples at a time until a short read
(``src/qpsk.c:420,436-445``).  Output: for every VALID call 496 bytes are appended to the bit file
(``fwrite(ibits, 1, BITS_PER_FRAME, fout)``, ``src/qpsk.c:455-457``) of which only the first 62 are
written by the modem (one bit per byte, bits[2i] = Q, bits[2i+1] = I) -- the reference leaves
the other 434 bytes uninitialised, here they are zero.

``demodulate_files`` runs any number of files as one bank: file k is stream k, shorter files are padded
with silence and their surplus calls discarded (later calls never influence earlier ones).
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np

from .modem import BITS_PER_CALL, FRAME_SIZE, ModemBank, unpack_bits

BITS_PER_FRAME_ON_DISK = 496            # headers/qpsk_internal.h:50


def read_raw(path: str) -> np.ndarray:
    return np.fromfile(path, dtype="<i2")


def bits_file_bytes(results_row: np.ndarray) -> bytes:
    """What the reference would have appended to RX_FILENAME for one stream's calls."""
    rows = unpack_bits(results_row)
    out = bytearray()
    for n in range(results_row.shape[0]):
        if results_row["valid"][n]:
            rec = np.zeros(BITS_PER_FRAME_ON_DISK, np.uint8)
            rec[:BITS_PER_CALL] = rows[n]
            out += rec.tobytes()
    return bytes(out)


def demodulate_files(paths: Sequence[str], out_paths: Optional[Sequence[str]] = None, device: int = 0,
                     wide: bool = False, foffset_hz: float = 0.0) -> List[np.ndarray]:
    """Demodulate raw sample files as one bank; returns per file its results (one record per call) and,
    if out_paths is given, writes the reference-format bit files."""
    data = [read_raw(p) for p in paths]
    frames = [d.size // FRAME_SIZE for d in data]
    nf = max(frames) if frames else 0
    if nf == 0:
        return [np.zeros(0, dtype=ModemBank.result_dtype()) for _ in paths]
    batch = np.zeros((len(data), nf * FRAME_SIZE), np.int16)
    for k, d in enumerate(data):
        batch[k, : frames[k] * FRAME_SIZE] = d[: frames[k] * FRAME_SIZE]
    bank = ModemBank(len(data), device=device, wide=wide, foffset_hz=foffset_hz)
    try:
        res, _ = bank.rx_frames_host(batch, nf)
    finally:
        bank.close()
    out = [res[k, : frames[k]].copy() for k in range(len(data))]
    if out_paths is not None:
        for k, p in enumerate(out_paths):
            with open(p, "wb") as f:
                f.write(bits_file_bytes(out[k]))
    return out
