#!/usr/bin/env python
"""Time the reference's CPU implementation of the RX path on this box's host cores.

TEST/BENCH INFRASTRUCTURE: used only by bench.py (the ``cpu_baseline`` leg and ``--impl reference``).
The reference keeps all state in file-scope statics (not re-entrant), so the
multi-core baseline is one PROCESS per core, each looping qpsk_rx_frame() over its own slice of the
streams, exactly as N copies of the reference binary would.  ``kind`` is "reference" when
oracle/_ref/libsc_ref.so (the reference's own objects) is present, else "port" (the restatement).

usage: python -m oracle.cpu_bench SAMPLES.npy N_FRAMES [N_PROCS] [REPEATS]
SAMPLES.npy holds int16 [n_streams, >= N_FRAMES*1880].  Prints one JSON line.
"""
from __future__ import annotations

import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))


def _worker(rank, n_procs, path, n_frames, kind, repeats, barrier, q):
    try:
        cores = sorted(os.sched_getaffinity(0))
        os.sched_setaffinity(0, {cores[rank % len(cores)]})
    except (AttributeError, OSError):
        pass
    from oracle import pyoracle as po
    x = np.load(path, mmap_mode="r")
    ns = x.shape[0]
    lo, hi = ns * rank // n_procs, ns * (rank + 1) // n_procs
    mine = np.ascontiguousarray(x[lo:hi])
    eng = po.Reference() if kind == "reference" else po.Oracle()
    eng.run_streams(mine[: max(1, min(2, hi - lo))], n_frames)          # warm the caches / page in
    times, valid = [], 0
    for _ in range(repeats):
        barrier.wait()
        t0 = time.perf_counter()
        _, v = eng.run_streams(mine, n_frames)
        t1 = time.perf_counter()
        times.append((t0, t1))
        valid = int(v.sum())
    q.put((rank, hi - lo, times, valid))


def run(path: str, n_frames: int, n_procs: int | None = None, repeats: int = 1) -> dict:
    from oracle import pyoracle as po
    po.build()
    kind = "reference" if po.have_ref() else "port"
    if n_procs is None:
        try:
            n_procs = len(os.sched_getaffinity(0))
        except AttributeError:
            n_procs = os.cpu_count() or 1
    ns = np.load(path, mmap_mode="r").shape[0]
    n_procs = max(1, min(n_procs, ns))
    ctx = mp.get_context("spawn")
    barrier, q = ctx.Barrier(n_procs), ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, n_procs, path, n_frames, kind, repeats, barrier, q))
             for r in range(n_procs)]
    for p in procs:
        p.start()
    got = [q.get() for _ in procs]
    for p in procs:
        p.join()
    walls = []
    for k in range(repeats):
        t0 = min(g[2][k][0] for g in got)
        t1 = max(g[2][k][1] for g in got)
        walls.append(t1 - t0)
    streams = sum(g[1] for g in got)
    syms = streams * n_frames * 376
    return {"kind": kind, "cores": n_procs, "streams": streams, "n_frames": n_frames,
            "wall_s": walls, "msym_per_s": [syms / w / 1e6 for w in walls],
            "valid_frames": sum(g[3] for g in got)}


if __name__ == "__main__":
    a = sys.argv[1:]
    out = run(a[0], int(a[1]), int(a[2]) if len(a) > 2 and int(a[2]) > 0 else None, int(a[3]) if len(a) > 3 else 1)
    print(json.dumps(out))
