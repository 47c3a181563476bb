import sys, torch
sys.path.insert(0, '/root/repo')
import singlecarrier_b200 as sc
from singlecarrier_b200._lib import check
L = sc.lib
ns = 1 << 19
s = torch.randn((ns, 256, 2), device="cuda")
idx = torch.empty(ns, dtype=torch.int32, device="cuda")
val = torch.empty(ns, dtype=torch.float32, device="cuda")
for _ in range(2):
    check(L.sc_preamble_search_batch_dev(0, ns, s.data_ptr(), 256, idx.data_ptr(), val.data_ptr(), None))
    check(L.sc_preamble_search_direct_batch_dev(0, ns, s.data_ptr(), 256, idx.data_ptr(), val.data_ptr(), None))
torch.cuda.synchronize()
