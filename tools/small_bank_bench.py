import sys, torch
sys.path.insert(0, '/root/repo')
import singlecarrier_b200 as sc
from singlecarrier_b200 import harness
for ns in (1024, 8192):
    nf = 42
    bank = sc.ModemBank(ns)
    wl = harness.synthesize(bank, nf * 1880 + 1040, seed=7, config=4)
    res = torch.zeros((ns, nf * 32), dtype=torch.uint8, device="cuda")
    def step():
        bank.reset(); bank.rx_frames_dev(wl.samples, nf, res)
    for _ in range(3): step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"{ns} streams x {nf} calls: {ms:.3f} ms  {ns*nf*376/ms/1e6:.2f} Gsym/s")
    bank.close()
