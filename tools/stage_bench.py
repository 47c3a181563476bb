#!/usr/bin/env python
"""Throughput of the stage kernels (the batched forms of the reference's L1 functions) on one B200.
CUDA events, 3 warm-ups, inputs far larger than L2.  Prints a markdown table (-> profiles/)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import singlecarrier_b200 as sc  # noqa: E402
from singlecarrier_b200._lib import check  # noqa: E402

L = sc.lib
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0


def timeit(fn, iters=5):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


rows = []
# fir(): 16384 streams x 18800 complex samples = 2.46 GB in place (read + write = 16 B / sample)
ns, n = 16384, 18800
x = torch.randn((ns, n, 2), device="cuda")
mem = torch.zeros((ns, 49, 2), device="cuda")
ms = timeit(lambda: check(L.sc_fir_batch_dev(0, ns, 0, mem.data_ptr(), x.data_ptr(), n, n, None)))
gbs = ns * n * 16 / ms / 1e6
ops = ns * n * 198 / ms / 1e9
rows.append(("fir_batch10_kernel (fir.h, exact)", f"{ns} streams x {n} samples", ms, gbs, gbs / peak, f"{ops:.1f} Tops/s of 37.2"))
ms = timeit(lambda: check(L.sc_fir_batch_dev(0, ns, 2, mem.data_ptr(), x.data_ptr(), n, n, None)))
gbs = ns * n * 16 / ms / 1e6
rows.append(("fir_batch10_kernel (SC_FIR_FAST: FMA, tolerance parity)", f"{ns} streams x {n} samples", ms, gbs, gbs / peak, f"{ns * n * 100 / ms / 1e9:.1f} Tfma-slots/s of 37.2"))
del x, mem
# correlate+argmax: 2^20 windows of 255 symbols (2,048 B each)
ns = 1 << 20
s = torch.randn((ns, 256, 2), device="cuda")
idx = torch.empty(ns, dtype=torch.int32, device="cuda")
val = torch.empty(ns, dtype=torch.float32, device="cuda")
ms = timeit(lambda: check(L.sc_preamble_search_tcgen05_batch_dev(0, ns, s.data_ptr(), 256, idx.data_ptr(), val.data_ptr(), None, None)))
gbs = ns * 2048 / ms / 1e6
rows.append(("search_umma_batch_kernel (tcgen05 + TMEM + TMA proposer, exact verifier), noise-only windows", f"{ns} windows", ms, gbs, gbs / peak, "-"))
# the same with a preamble in every window (one clear maximum: a single candidate per window)
import ctypes as C
import numpy as np
pre = torch.from_numpy(np.frombuffer((C.c_int8 * 128).in_dll(L, "preamblevalues"), np.int8).astype(np.float32)).cuda()
s2 = s.clone()
lag = (torch.arange(ns, device="cuda") * 37) % 128
cols = lag[:, None] + torch.arange(128, device="cuda")[None, :]
s2.scatter_add_(1, cols[:, :, None].expand(-1, -1, 2), (3.0 * pre)[None, :, None].expand(ns, -1, 2).contiguous())
ms = timeit(lambda: check(L.sc_preamble_search_tcgen05_batch_dev(0, ns, s2.data_ptr(), 256, idx.data_ptr(), val.data_ptr(), None, None)))
gbs = ns * 2048 / ms / 1e6
rows.append(("search_umma_batch_kernel, a preamble in every window", f"{ns} windows", ms, gbs, gbs / peak, "-"))
assert bool((idx.long() == lag).float().mean() > 0.999)
del s2, cols
ms = timeit(lambda: check(L.sc_preamble_search_mma_batch_dev(0, ns, s.data_ptr(), 256, idx.data_ptr(), val.data_ptr(), None)))
gbs = ns * 2048 / ms / 1e6
rows.append(("search_mma_batch_kernel (mma.sync proposer + exact verifier)", f"{ns} windows", ms, gbs, gbs / peak, "-"))
ms = timeit(lambda: check(L.sc_preamble_search_fft_batch_dev(0, ns, s.data_ptr(), 256, idx.data_ptr(), val.data_ptr(), None)))
gbs = ns * 2048 / ms / 1e6
rows.append(("search_fft_batch_kernel (warp-shuffle FFT proposer + exact verifier)", f"{ns} windows", ms, gbs, gbs / peak, "-"))
ms = timeit(lambda: check(L.sc_preamble_search_direct_batch_dev(0, ns, s.data_ptr(), 256, idx.data_ptr(), val.data_ptr(), None)))
gbs = ns * 2048 / ms / 1e6
rows.append(("search_batch_kernel (every lag exact)", f"{ns} windows", ms, gbs, gbs / peak, f"{ns * 33152 / ms / 1e9:.1f} Tops/s of 37.2"))
del s
# decision loop on explicit windows: 2^18 windows of 290 symbols
ns = 1 << 18
s = torch.randn((ns, 290, 2), device="cuda") * 0.5
mi = torch.randint(0, 128, (ns,), dtype=torch.int32, device="cuda")
mv = torch.zeros(ns, device="cuda")
tm = torch.full((ns,), 128, dtype=torch.int32, device="cuda")
res = torch.empty((ns, 32), dtype=torch.uint8, device="cuda")
ms = timeit(lambda: check(L.sc_track_decide_batch_dev(0, ns, s.data_ptr(), 290, mi.data_ptr(), mv.data_ptr(), tm.data_ptr(), 5,
                                                      res.data_ptr(), None, None)))
gbs = ns * 1374 / ms / 1e6
rows.append(("track_window_kernel (decision loop, exact)", f"{ns} windows", ms, gbs, gbs / peak, f"{ns * 63251 / ms / 1e9:.1f} Tops/s of 37.2"))
# FFT 256 (the size a 128x128 correlation would need), 2^18 transforms
nb = 1 << 18
a = torch.randn((nb, 256, 2), device="cuda")
b = torch.empty_like(a)
ms = timeit(lambda: check(L.sc_fft_batch_dev(0, nb, 256, 0, a.data_ptr(), b.data_ptr(), None)))
gbs = nb * 256 * 16 / ms / 1e6
rows.append(("fft_kernel n=256 (kiss order, exact)", f"{nb} transforms", ms, gbs, gbs / peak, "-"))

print("| kernel | workload | ms | algorithmic GB/s | of measured HBM peak | FP32 |\n|---|---|---|---|---|---|")
for r in rows:
    print(f"| `{r[0]}` | {r[1]} | {r[2]:.3f} | {r[3]:.0f} | {100 * r[4]:.1f}% | {r[5]} |")
