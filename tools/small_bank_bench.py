"""Small banks: the RX chain with the input resident in HBM -- serial call chain or overlapped even / odd chains
(SC_OPT_OVERLAP), one-thread or lane-cooperative tracker (SC_OPT_TRACKER).  All four give identical results.
usage: python tools/small_bank_bench.py [n_streams ...]"""
import sys, torch
sys.path.insert(0, '/root/repo')
import singlecarrier_b200 as sc
from singlecarrier_b200 import harness
from singlecarrier_b200.modem import (OPT_OVERLAP, OPT_TRACKER, OVERLAP_OFF, OVERLAP_ON, TRACKER_COOP, TRACKER_THREAD)

sizes = [int(a) for a in sys.argv[1:]] or [1024, 4096, 8192, 16384, 32768]
nf = 42
print("| streams | chain | tracker | ms / 42 calls | us / call | Gsym/s |\n|---|---|---|---|---|---|")
for ns in sizes:
    outs = []
    for overlap in (OVERLAP_OFF, OVERLAP_ON):
        for tracker in (TRACKER_THREAD, TRACKER_COOP):
            if tracker == TRACKER_COOP and ns > 16384:
                continue
            bank = sc.ModemBank(ns)
            bank.set_option(OPT_TRACKER, tracker)
            bank.set_option(OPT_OVERLAP, overlap)
            wl = harness.synthesize(bank, nf * 1880 + 1040, seed=7, config=4)
            res = torch.zeros((ns, nf * 32), dtype=torch.uint8, device="cuda")

            def step():
                bank.reset()
                bank.rx_frames_dev(wl.samples, nf, res)

            for _ in range(3):
                step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                step()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            outs.append(res.cpu().numpy().tobytes())
            print(f"| {ns} | {'serial' if overlap == OVERLAP_OFF else 'overlapped'} | "
                  f"{'thread' if tracker == TRACKER_THREAD else 'coop'} | {ms:.3f} | {ms / nf * 1e3:.1f} | "
                  f"{ns * nf * 376 / ms / 1e6:.2f} |", flush=True)
            bank.close()
    print(f"<!-- {ns} streams: results identical: {all(o == outs[0] for o in outs)} -->", flush=True)
