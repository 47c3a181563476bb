#!/usr/bin/env python
"""Packet mode (extension) at the bench workload: cost of the full-packet pass on top of the ordinary chain.
131,072 streams x 10 s, packet-mode TX (scrambled) + channel; CUDA events; prints one markdown table."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import singlecarrier_b200 as sc  # noqa: E402
from singlecarrier_b200 import harness  # noqa: E402

ns, nf = int(os.environ.get("STREAMS", 131072)), 42
bank = sc.ModemBank(ns, packet=True)
wl = harness.synthesize(bank, 80000, seed=0x5C0DE5, config=4)
d_res = torch.zeros((ns, nf * 32), dtype=torch.uint8, device="cuda")
cap = ns * 8
d_pk = torch.zeros(cap * 96, dtype=torch.uint8, device="cuda")
d_n = torch.zeros(1, dtype=torch.int64, device="cuda")


def timeit(fn, iters=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def plain():
    bank.reset()
    bank.rx_frames_dev(wl.samples, nf, d_res)


def packets():
    bank.reset()
    d_n.zero_()
    bank.rx_packets_dev(wl.samples, nf, d_res, d_pk, d_n)


t0 = timeit(plain)
t1 = timeit(packets)
n = int(d_n.item())
pk = d_pk.cpu().numpy().view(sc.PACKET_DTYPE)[: min(n, cap)]
sym = ns * nf * 376
# bit errors of the packets that sit on a transmitted packet
r = d_res.cpu().numpy().view(sc.RESULT_DTYPE)
lead = wl.lead.cpu().numpy()
tx = wl.tx_bits.cpu().numpy()
s_, c_ = pk["stream"], pk["call_index"].astype(np.int64)
t_prev = r["rx_timing"][s_, c_ - 2].astype(np.int64)
pos = (c_ - 2) * 1880 + 5 * pk["max_index"].astype(np.int64) + t_prev - 48 - lead[s_]
j = np.rint(pos / 2783).astype(np.int64)
ok = (j >= 0) & (j < tx.shape[1]) & (np.abs(pos - j * 2783) <= 10)
bits = sc.unpack_packet_bits(pk[ok])
err = (bits != tx[s_[ok], j[ok]]).reshape(-1, 8, 62).sum(2)
print(f"| | ms per step ({ns} streams x {nf} calls) | Gsym/s |\n|---|---|---|")
print(f"| ordinary chain on a SC_FLAG_PACKET bank (`sc_rx_frames_dev`) | {t0:.2f} | {sym / t0 / 1e6:.1f} |")
print(f"| + full-packet pass (`sc_rx_packets_dev`) | {t1:.2f} | {sym / t1 / 1e6:.1f} |")
print(f"\n{n} packets decoded ({100.0 * n / (ns * nf):.1f} % of the calls), {int(ok.sum())} of them on a transmitted packet; "
      f"bit error rate per data frame 1..8: {np.round(err.mean(0) / 62, 3).tolist()}, overall {err.sum() / (ok.sum() * 496):.3f}")
