"""Small banks: the call-by-call RX chain, input resident in HBM, with the one-thread and the lane-cooperative tracker.
usage: python tools/small_bank_bench.py [n_streams ...]"""
import sys, torch
sys.path.insert(0, '/root/repo')
import singlecarrier_b200 as sc
from singlecarrier_b200 import harness

sizes = [int(a) for a in sys.argv[1:]] or [1024, 4096, 8192, 16384, 32768]
nf = 42
for ns in sizes:
    outs = {}
    for tracker in (sc.modem.TRACKER_THREAD, sc.modem.TRACKER_COOP):
        bank = sc.ModemBank(ns)
        bank.set_option(sc.modem.OPT_TRACKER, tracker)
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
        outs[tracker] = res.cpu().numpy().tobytes()
        name = {1: "thread", 2: "coop"}[tracker]
        print(f"{ns} streams x {nf} calls, tracker {name:6s}: {ms:.3f} ms  {ms / nf * 1e3:.1f} us/call  "
              f"{ns * nf * 376 / ms / 1e6:.2f} Gsym/s", flush=True)
        bank.close()
    print(f"{ns} streams: results identical: {outs[1] == outs[2]}", flush=True)
