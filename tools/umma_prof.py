"""development aid: where the roles of search_umma_batch_kernel wait (library built with make SU_DEFS=-DSU_PROFILE)"""
import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
import singlecarrier_b200 as sc
ns, stride = 1048576, 256
rng = np.random.default_rng(1)
sym = (rng.normal(size=(ns, stride)) + 1j * rng.normal(size=(ns, stride))).astype(np.complex64)
d = torch.from_numpy(sym.view(np.float32)).cuda()
idx = torch.zeros((ns,), dtype=torch.int32, device="cuda"); val = torch.zeros((ns,), dtype=torch.float32, device="cuda")
dbg = torch.zeros((ns, 128), dtype=torch.float32, device="cuda")
for _ in range(2):
    sc._lib.check(sc.lib.sc_preamble_search_tcgen05_batch_dev(0, ns, d.data_ptr(), stride, idx.data_ptr(), val.data_ptr(), dbg.data_ptr(), 0))
torch.cuda.synchronize()
o = dbg.view(-1)[:8 * 22].cpu().numpy().reshape(-1, 8)
names = [f"epi{i}" for i in range(8)] + ["mma0", "mma1", "tma"] + [f"stg{i}" for i in range(8)] + [f"ver{i}" for i in range(3)]
for n, r in zip(names, o):
    print(f"{n:5s} total {r[0]:10.0f} " + " ".join(f"w{k} {100*r[k+1]/r[0]:5.1f}%" for k in range(6)))
