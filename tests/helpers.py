"""Shared helpers for the parity tests (test infrastructure)."""
import numpy as np

from oracle import pyoracle as po


synth_streams = po.synth_streams


def oracle_results(oracle, samples, n_frames, wide=False, foffset=0.0):
    """Run every stream through the oracle; returns (bits[ns, nf, 62], stats[ns, nf]).  The restatement is
    re-entrant (one state per call) and ctypes drops the GIL, so the streams are spread over the host cores."""
    import os
    from concurrent.futures import ThreadPoolExecutor
    ns = samples.shape[0]
    bits = np.zeros((ns, n_frames, 62), np.uint8)
    stats = np.zeros((ns, n_frames), po.STATS_DTYPE)

    def one(s):
        bits[s], stats[s] = oracle.run_stream(samples[s, : n_frames * po.FRAME_SIZE], wide=wide, foffset=foffset)

    try:
        workers = len(os.sched_getaffinity(0))
    except AttributeError:
        workers = os.cpu_count() or 1
    if ns < 8 or workers == 1:
        for s in range(ns):
            one(s)
    else:
        with ThreadPoolExecutor(max_workers=workers) as ex:
            list(ex.map(one, range(ns)))
    return bits, stats


def compare_results(res, eq, obits, ostats, check_invalid_bits=False):
    """GPU results (RESULT_DTYPE) vs oracle stats, bit-exact.  Returns a list of mismatch strings."""
    import singlecarrier_b200 as sc
    bad = []

    def chk(name, a, b):
        if not np.array_equal(a, b):
            idx = np.argwhere(a != b)
            bad.append(f"{name}: {len(idx)} mismatches, first at {idx[0].tolist()}: gpu={a[tuple(idx[0])]} oracle={b[tuple(idx[0])]}")

    chk("valid", res["valid"].astype(np.int32), ostats["valid"])
    chk("max_index", res["max_index"].astype(np.int32), ostats["max_index"])
    chk("matches", res["matches"].astype(np.int32), ostats["matches"])
    chk("rx_timing", res["rx_timing"].astype(np.int32), ostats["rx_timing"])
    chk("max_value(bits)", res["max_value"].view(np.uint32), ostats["max_value"].view(np.uint32))
    chk("cost(bits)", res["cost"].view(np.uint32), ostats["cost"].view(np.uint32))
    rows = sc.unpack_bits(res)
    v = ostats["valid"].astype(bool)
    chk("bits(valid rows)", rows[v], obits[v])
    if eq is not None:
        chk("eq_coeff(bits)", np.ascontiguousarray(eq).view(np.uint32), np.ascontiguousarray(ostats["eq_coeff"]).view(np.uint32))
    return bad
