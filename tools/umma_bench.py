"""tcgen05 preamble search: noise-only windows and windows with a preamble each, against the other search kernels"""
import ctypes as C, sys
import numpy as np, torch
sys.path.insert(0, '/root/repo')
import singlecarrier_b200 as sc
L = sc.lib
ns = 1 << 20
torch.manual_seed(1)
s = torch.randn((ns, 256, 2), device="cuda")
pre = torch.from_numpy(np.frombuffer((C.c_int8 * 128).in_dll(L, "preamblevalues"), np.int8).astype(np.float32)).cuda()
s2 = s.clone()
lag = (torch.arange(ns, device="cuda") * 37) % 128
cols = lag[:, None] + torch.arange(128, device="cuda")[None, :]
s2.scatter_add_(1, cols[:, :, None].expand(-1, -1, 2), (3.0 * pre)[None, :, None].expand(ns, -1, 2).contiguous())
idx = torch.empty(ns, dtype=torch.int32, device="cuda"); val = torch.empty(ns, dtype=torch.float32, device="cuda")
i2 = torch.empty(ns, dtype=torch.int32, device="cuda"); v2 = torch.empty(ns, dtype=torch.float32, device="cuda")
def timeit(fn):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 10
for name, x in (("noise", s), ("preamble", s2)):
    sc._lib.check(L.sc_preamble_search_direct_batch_dev(0, ns, x.data_ptr(), 256, i2.data_ptr(), v2.data_ptr(), None))
    for kn, fn in (("tcgen05", lambda: L.sc_preamble_search_tcgen05_batch_dev(0, ns, x.data_ptr(), 256, idx.data_ptr(), val.data_ptr(), None, None)),
                   ("mma.sync", lambda: L.sc_preamble_search_mma_batch_dev(0, ns, x.data_ptr(), 256, idx.data_ptr(), val.data_ptr(), None))):
        ms = timeit(fn)
        torch.cuda.synchronize()
        same = bool((idx == i2).all()) and bool((val.view(torch.int32) == v2.view(torch.int32)).all())
        print(f"{name:9s} {kn:9s} {ms:.4f} ms  {ns * 2048 / ms / 1e6:7.1f} GB/s  identical to the all-exact kernel: {same}", flush=True)
