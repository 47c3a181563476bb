"""Shared helpers for the parity tests (test infrastructure)."""
import numpy as np

from oracle import pyoracle as po


def synth_streams(oracle, rng, n_streams, n_frames, noise_levels=(0.0, 30.0, 300.0, 1500.0, 4000.0),
                  max_lead=2000, gaps=(903, 0, 500, 1880)):
    """Oracle-TX loop-back streams with random lead-in, dead air and additive noise -> int16[n, n_frames*1880]."""
    total = n_frames * po.FRAME_SIZE
    out = np.zeros((n_streams, total), np.int16)
    for s in range(n_streams):
        st = oracle.new_state()
        lead = int(rng.integers(0, max_lead))
        gap = int(gaps[s % len(gaps)])
        parts = [np.zeros(lead, np.int16)]
        n = lead
        while n < total:
            parts.append(oracle.tx_preamble(st))
            for _ in range(8):
                parts.append(oracle.tx_data(st, rng.integers(0, 2, 62).astype(np.uint8)))
            parts.append(np.zeros(gap, np.int16))
            n += 1880 + gap
        x = np.concatenate(parts)[:total].astype(np.float64)
        noise = noise_levels[s % len(noise_levels)]
        if noise > 0:
            x = x + rng.normal(0, noise, total)
        out[s] = np.clip(x, -32767, 32767).round().astype(np.int16)
    return out


def oracle_results(oracle, samples, n_frames, wide=False, foffset=0.0):
    """Run every stream through the oracle; returns dict of arrays [n_streams, n_frames(, ...)]."""
    ns = samples.shape[0]
    bits = np.zeros((ns, n_frames, 62), np.uint8)
    stats = np.zeros((ns, n_frames), po.STATS_DTYPE)
    for s in range(ns):
        b, st = oracle.run_stream(samples[s, : n_frames * po.FRAME_SIZE], wide=wide, foffset=foffset)
        bits[s], stats[s] = b, st
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
