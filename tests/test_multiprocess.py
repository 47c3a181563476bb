"""CPU, world_size 2 over gloo: the N>1 host logic of bench.py -- contiguous sharding of the streams by
rank (no data-path collective), the all-reduce of the 16 lock/bit counters, and max-over-ranks timing."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def shard(n_streams, world, rank):
    per = n_streams // world
    return rank * per, (rank + 1) * per


def counters_from_results(valid, matches, max_index, rx_timing, bits):
    c = np.zeros(16, np.int64)
    v = valid.astype(bool)
    c[0], c[1], c[2], c[3] = valid.size, v.sum(), matches.sum(), matches[v].sum()
    c[4], c[7] = max_index[v].sum(), rx_timing.sum()
    c[5] = sum(bin(int(b)).count("1") for b in bits[v])
    c[6] = sum((int(b) & 0xffffffff) + (int(b) >> 32) for b in bits[v])
    for h in range(8):
        c[8 + h] = (np.minimum(matches >> 4, 7) == h).sum()
    return c


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(1234)                       # same table on every rank = the whole job's results
    n, nf = 64, 5
    valid = rng.integers(0, 2, (n, nf))
    matches = rng.integers(60, 129, (n, nf))
    max_index = rng.integers(0, 128, (n, nf))
    rx_timing = rng.integers(128, 256, (n, nf))
    bits = rng.integers(0, 2 ** 62, (n, nf), dtype=np.uint64)
    lo, hi = shard(n, world, rank)
    mine = counters_from_results(valid[lo:hi], matches[lo:hi], max_index[lo:hi], rx_timing[lo:hi], bits[lo:hi])
    t = torch.from_numpy(mine.copy())
    dist.all_reduce(t)                                      # the only collective of the path
    whole = counters_from_results(valid, matches, max_index, rx_timing, bits)
    ms = torch.tensor([10.0 + 5.0 * rank], dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    q.put((rank, t.numpy().tolist() == whole.tolist(), float(ms.item()), (lo, hi)))
    dist.barrier()
    dist.destroy_process_group()


def test_sharding_and_counter_allreduce_world2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [g[1] for g in got] == [True, True]
    assert [g[2] for g in got] == [15.0, 15.0]              # max over ranks
    assert [g[3] for g in got] == [(0, 32), (32, 64)]        # contiguous, disjoint, covering


def test_reference_arm_json_contract(tmp_path):
    """bench.py --impl reference prints one JSON line with the contract keys (tiny sample, CPU only)."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--ref-streams", "16", "--seconds", "1"], cwd=root, capture_output=True, text=True,
                         timeout=600)
    assert out.returncode == 0, out.stderr
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "demodulated_msym_per_s" and line["unit"] == "Msym/s"
    assert line["value"] > 0 and line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["higher_is_better"] is True
