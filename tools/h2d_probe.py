#!/usr/bin/env python
"""Host -> device copy ceiling of this box, all ranks at once.

Plain cudaMemcpyAsync / cudaMemcpy2DAsync from pinned host memory through the library's own probe
(sc_h2d_probe; NUMA-aware pinned allocation, CUDA events).  Run alone for one GPU or under torchrun for N:

    python tools/h2d_probe.py
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/h2d_probe.py

Patterns: contiguous; rows of 3,248 bytes every 157,920 bytes (sc_rx_frames_host's per-frame column copy of a
42-frame stream); rows of 33,328 bytes (SC_H2D_ROWS blocks of 9 frames); device -> host contiguous.  Prints
one markdown table (per-GPU min / max over ranks and the aggregate) from rank 0.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import singlecarrier_b200 as sc  # noqa: E402

rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def gather(x):
    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    if world == 1:
        return [x]
    out = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    return [float(o.item()) for o in out]


PATTERNS = [("contiguous 512 MiB", dict()),
            ("2-D, 3,248-byte rows / 157,920-byte pitch", dict(row_bytes=3248, src_pitch_bytes=157920)),
            ("2-D, 33,328-byte rows / 157,920-byte pitch", dict(row_bytes=33328, src_pitch_bytes=157920)),
            ("device -> host contiguous", dict(d2h=True))]
rows = []
for name, kw in PATTERNS:
    barrier()
    g = sc.h2d_probe(local, 512 << 20, min_seconds=float(os.environ.get("PROBE_SECONDS", "1.0")), **kw)
    rows.append((name, gather(g)))
if rank == 0:
    print(f"| pattern ({world} GPU{'s' if world > 1 else ''} at once) | per-GPU min GB/s | per-GPU max GB/s | aggregate GB/s |\n|---|---|---|---|")
    for name, v in rows:
        print(f"| {name} | {min(v):.1f} | {max(v):.1f} | {sum(v):.1f} |")
if world > 1:
    dist.destroy_process_group()
