"""diagnostic: host enqueue time vs device time of one batch, serial vs overlapped (small bank)"""
import sys, time, torch
sys.path.insert(0, '/root/repo')
import singlecarrier_b200 as sc
from singlecarrier_b200 import harness
from singlecarrier_b200.modem import OPT_OVERLAP, OPT_TRACKER, OVERLAP_OFF, OVERLAP_ON, TRACKER_COOP, TRACKER_THREAD
ns = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
nf = 42
for overlap in (OVERLAP_OFF, OVERLAP_ON):
    for tracker in (TRACKER_THREAD, TRACKER_COOP):
        bank = sc.ModemBank(ns)
        bank.set_option(OPT_TRACKER, tracker)
        bank.set_option(OPT_OVERLAP, overlap)
        wl = harness.synthesize(bank, nf * 1880 + 1040, seed=7, config=4)
        res = torch.zeros((ns, nf * 32), dtype=torch.uint8, device="cuda")
        for _ in range(3):
            bank.reset(); bank.rx_frames_dev(wl.samples, nf, res)
        torch.cuda.synchronize()
        host, dev = [], []
        for _ in range(5):
            bank.reset()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            t0 = time.perf_counter()
            bank.rx_frames_dev(wl.samples, nf, res)
            t1 = time.perf_counter()
            e1.record()
            torch.cuda.synchronize()
            host.append((t1 - t0) * 1e3); dev.append(e0.elapsed_time(e1))
        print(f"{ns} overlap={overlap} tracker={tracker}: host enqueue {min(host):.3f} ms, device {min(dev):.3f} ms", flush=True)
        bank.close()
