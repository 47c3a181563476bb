"""development aid: time stamps of the overlapped chains (SC_OV_TRACE=1 python tools/ov_trace.py n_streams tracker)"""
import sys, torch
sys.path.insert(0, '/root/repo')
import singlecarrier_b200 as sc
from singlecarrier_b200 import harness
from singlecarrier_b200.modem import OPT_OVERLAP, OPT_TRACKER, OVERLAP_ON
ns, tracker = int(sys.argv[1]), int(sys.argv[2])
nf = 42
bank = sc.ModemBank(ns)
bank.set_option(OPT_TRACKER, tracker)
bank.set_option(OPT_OVERLAP, OVERLAP_ON)
wl = harness.synthesize(bank, nf * 1880 + 1040, seed=7, config=4)
res = torch.zeros((ns, nf * 32), dtype=torch.uint8, device="cuda")
for _ in range(4):
    bank.reset(); bank.rx_frames_dev(wl.samples, nf, res)
    torch.cuda.synchronize()
    print("=== batch done", file=sys.stderr, flush=True)
