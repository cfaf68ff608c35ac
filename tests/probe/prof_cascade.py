"""Profiling aid: the cascade loss / multiplier kernels at the cfg2 size (ncu: -k regex:'cascade|multiplier' -s 3 -c 3)."""
import sys, torch
sys.path.insert(0, ".")
from lshm_b200._lib import lib
dev = torch.device("cuda:0"); st = torch.cuda.current_stream().cuda_stream
N, C, P = 1024, 8, 128
n = N * C * P * P
t = [torch.randn(n, device=dev) for _ in range(7)]
g = [torch.empty(n, device=dev) for _ in range(4)]
sums = torch.zeros(8, dtype=torch.float64, device=dev)
db = [torch.zeros(C, device=dev) for _ in range(3)]
L = lib()
for _ in range(2):
    L.cascade_losses(*(x.data_ptr() for x in t), 1.0, N, C, P, 1.0 / n, sums.data_ptr(), g[0].data_ptr(), g[1].data_ptr(),
                     g[2].data_ptr(), db[1].data_ptr(), db[2].data_ptr(), st)
    L.cascade_combine(g[0].data_ptr(), g[1].data_ptr(), g[2].data_ptr(), g[3].data_ptr(), N, C, P, db[0].data_ptr(), st)
    L.multiplier_update(t[0].data_ptr(), t[1].data_ptr(), t[2].data_ptr(), t[3].data_ptr(), 1.0, t[4].data_ptr(),
                        t[5].data_ptr(), t[6].data_ptr(), N, C, P, st)
torch.cuda.synchronize(); print("ok")
