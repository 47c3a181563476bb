"""development check of the tcgen05 search: proposed values against numpy, results against the all-exact kernel, timing"""
import ctypes as C, sys
import numpy as np, torch
sys.path.insert(0, '/root/repo')
import singlecarrier_b200 as sc
from oracle import pyoracle as po
oracle = po.Oracle()
pv = np.frombuffer((C.c_int8 * 128).in_dll(oracle.lib, "sco_preamblevalues"), np.int8).astype(np.float32)
rng = np.random.default_rng(1)
ns, stride = int(sys.argv[1]) if len(sys.argv) > 1 else 100, 256
sym = (rng.normal(size=(ns, stride)) + 1j * rng.normal(size=(ns, stride))).astype(np.complex64)
for s in range(0, ns, 2):
    lag = (s // 2) % 128
    sym[s, lag:lag + 128] += (np.float32(0.5 + (s % 5)) * pv * (1 + 1j)).astype(np.complex64)
if ns > 40:
    sym[33] = 0
    sym[35] = 0
    sym[35, 120] = 1 + 1j
d = torch.from_numpy(sym.view(np.float32)).cuda()
idx = torch.full((ns,), -7, dtype=torch.int32, device="cuda")
val = torch.full((ns,), -7.0, dtype=torch.float32, device="cuda")
approx = torch.zeros((ns, 128), dtype=torch.float32, device="cuda")
sc._lib.check(sc.lib.sc_preamble_search_tcgen05_batch_dev(0, ns, d.data_ptr(), stride, idx.data_ptr(), val.data_ptr(), approx.data_ptr(), 0))
torch.cuda.synchronize()
a = approx.cpu().numpy()
if ns > 5000:
    sym_small, a = sym[:2000], a[:2000]
else:
    sym_small = sym
# numpy reference of the correlation values (float64)
dd = (sym_small.real.astype(np.float64) - sym_small.imag), (sym_small.imag.astype(np.float64) + sym_small.real)
ref = np.zeros((len(sym_small), 128))
for L in range(128):
    re = (dd[0][:, L:L + 128] * pv).sum(axis=1)
    im = (dd[1][:, L:L + 128] * pv).sum(axis=1)
    ref[:, L] = re * re + im * im
err = np.abs(a - ref) / (ref.max(axis=1, keepdims=True) + 1e-30)
print("approx: max rel err vs float64", err.max(), "at", np.unravel_index(err.argmax(), err.shape))
if err.max() > 1e-3:
    print("window 0 approx[:8]", a[0, :8], "ref", ref[0, :8])
    print("window 1 approx[:8]", a[1, :8], "ref", ref[1, :8])
i2 = torch.full((ns,), -7, dtype=torch.int32, device="cuda")
v2 = torch.full((ns,), -7.0, dtype=torch.float32, device="cuda")
sc._lib.check(sc.lib.sc_preamble_search_direct_batch_dev(0, ns, d.data_ptr(), stride, i2.data_ptr(), v2.data_ptr(), 0))
torch.cuda.synchronize()
ai, av, bi, bv = idx.cpu().numpy(), val.cpu().numpy(), i2.cpu().numpy(), v2.cpu().numpy()
bad = np.nonzero((ai != bi) | (av.view(np.uint32) != bv.view(np.uint32)))[0]
print("results: mismatches", bad.size, "of", ns, bad[:10], ai[bad[:10]], bi[bad[:10]])
for name, fn in (("tcgen05", lambda: sc.lib.sc_preamble_search_tcgen05_batch_dev(0, ns, d.data_ptr(), stride, idx.data_ptr(), val.data_ptr(), None, 0)),
                 ("mma.sync", lambda: sc.lib.sc_preamble_search_mma_batch_dev(0, ns, d.data_ptr(), stride, idx.data_ptr(), val.data_ptr(), 0)),
                 ("direct", lambda: sc.lib.sc_preamble_search_direct_batch_dev(0, ns, d.data_ptr(), stride, idx.data_ptr(), val.data_ptr(), 0))):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"{name}: {ms:.4f} ms for {ns} windows = {ns * 2048 / ms / 1e6:.1f} GB/s")
