import sys, torch
sys.path.insert(0, '/root/repo')
import singlecarrier_b200 as sc
from singlecarrier_b200._lib import check
L = sc.lib
ns, n = 16384, 18800
x = torch.randn((ns, n, 2), device="cuda")
mem = torch.zeros((ns, 49, 2), device="cuda")
for _ in range(2):
    check(L.sc_fir_batch_dev(0, ns, 2, mem.data_ptr(), x.data_ptr(), n, n, None))
    check(L.sc_fir_batch_dev(0, ns, 0, mem.data_ptr(), x.data_ptr(), n, n, None))
torch.cuda.synchronize()
