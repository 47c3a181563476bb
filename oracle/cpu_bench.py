#!/usr/bin/env python
"""Time the reference's CPU implementation of the RX path on this box's host cores.

TEST/BENCH INFRASTRUCTURE: used only by bench.py (the ``cpu_baseline`` leg and ``--impl reference``).
The reference keeps all state in file-scope statics (not re-entrant), so the
multi-core baseline is one PROCESS per core, each looping qpsk_rx_frame() over its own slice of the
streams, exactly as N copies of the reference binary would.

Engines (``kind`` / ``build`` in the result):
  reference / parity    oracle/_ref/libsc_ref.so          the reference's own objects, gcc -O2 -ffp-contract=off
                                                          (the flags every parity claim is pinned to)
  reference / fast_v3   oracle/_ref/libsc_ref_fast_v3.so  same sources, -O3 -march=x86-64-v3 (AVX2 + FMA contraction)
  reference / fast_v4   oracle/_ref/libsc_ref_fast_v4.so  same sources, -O3 -march=x86-64-v4 (AVX-512)
  port / parity         oracle/libsc_oracle.so            the restatement, when no reference build is present

Input: an .npy of int16 [n_streams, >= N_FRAMES*1880], or ``synth:SEED:N_STREAMS`` -- each worker then makes
its own streams on the CPU with pyoracle.synth_bench_stream (bench.py's workload; nothing of the product is
loaded).

usage: python -m oracle.cpu_bench INPUT N_FRAMES [N_PROCS] [REPEATS] [BUILD]
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

BUILDS = {"parity": "libsc_ref.so", "fast_v3": "libsc_ref_fast_v3.so", "fast_v4": "libsc_ref_fast_v4.so"}
FLAGS = {"parity": "gcc -std=gnu11 -O2 -ffp-contract=off", "fast_v3": "gcc -std=gnu11 -O3 -march=x86-64-v3",
         "fast_v4": "gcc -std=gnu11 -O3 -march=x86-64-v4", "port": "gcc -std=gnu11 -O2 -ffp-contract=off (restatement)"}


def cpu_flags() -> set:
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("flags"):
                return set(line.split(":", 1)[1].split())
    except OSError:
        pass
    return set()


def available_builds() -> list:
    """Reference builds that exist here AND can run on this CPU, parity build first."""
    f = cpu_flags()
    ok = []
    for b, so in BUILDS.items():
        if not os.path.exists(os.path.join(HERE, "_ref", so)):
            continue
        if b == "fast_v3" and not {"avx2", "fma", "bmi2"} <= f:
            continue
        if b == "fast_v4" and not {"avx512f", "avx512bw", "avx512cd", "avx512dq", "avx512vl"} <= f:
            continue
        ok.append(b)
    return ok


def _engine(build):
    from oracle import pyoracle as po
    if build == "port":
        return po.Oracle()
    return po.Reference(os.path.join(HERE, "_ref", BUILDS[build]))


def _worker(rank, n_procs, source, n_frames, build, repeats, barrier, q):
    try:
        cores = sorted(os.sched_getaffinity(0))
        os.sched_setaffinity(0, {cores[rank % len(cores)]})
    except (AttributeError, OSError):
        pass
    from oracle import pyoracle as po
    if source.startswith("synth:"):
        _, seed, ns = source.split(":")
        seed, ns = int(seed), int(ns)
        lo, hi = ns * rank // n_procs, ns * (rank + 1) // n_procs
        o = po.Oracle()
        mine = np.stack([po.synth_bench_stream(o, seed, s, n_frames * po.FRAME_SIZE) for s in range(lo, hi)])
    else:
        x = np.load(source, mmap_mode="r")
        ns = x.shape[0]
        lo, hi = ns * rank // n_procs, ns * (rank + 1) // n_procs
        mine = np.ascontiguousarray(x[lo:hi])
    eng = _engine(build)
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


def run(source: str, n_frames: int, n_procs: int | None = None, repeats: int = 1, build: str | None = None) -> dict:
    from oracle import pyoracle as po
    po.build()
    if build is None:
        build = "parity" if po.have_ref() else "port"
    if build != "port" and build not in available_builds():
        raise RuntimeError(f"reference build {build} is not available on this box")
    if n_procs is None:
        try:
            n_procs = len(os.sched_getaffinity(0))
        except AttributeError:
            n_procs = os.cpu_count() or 1
    ns = int(source.split(":")[2]) if source.startswith("synth:") else np.load(source, mmap_mode="r").shape[0]
    n_procs = max(1, min(n_procs, ns))
    ctx = mp.get_context("spawn")
    barrier, q = ctx.Barrier(n_procs), ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, n_procs, source, n_frames, build, repeats, barrier, q))
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
    return {"kind": "port" if build == "port" else "reference", "build": build, "flags": FLAGS[build], "cores": n_procs,
            "streams": streams, "n_frames": n_frames, "wall_s": walls, "msym_per_s": [syms / w / 1e6 for w in walls],
            "valid_frames": sum(g[3] for g in got)}


if __name__ == "__main__":
    a = sys.argv[1:]
    out = run(a[0], int(a[1]), int(a[2]) if len(a) > 2 and int(a[2]) > 0 else None, int(a[3]) if len(a) > 3 else 1,
              a[4] if len(a) > 4 else None)
    print(json.dumps(out))
