"""ncu probe: the loss pass with plane outputs (the dominant kernel of the step) and the residual-split writer at N=1024."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from lshm_b200._lib import lib
from lshm_b200.engine import planes_buffer
dev = torch.device("cuda:0"); L = lib(); st = torch.cuda.current_stream().cuda_stream
N, C, P = 1024, 8, 128
x, x1, x2, x3f = (torch.randn(N, C, P, P, device=dev) for _ in range(4))
ys = [torch.randn(N * C * P * P, device=dev) for _ in range(3)]
g1 = torch.empty_like(x); p2 = planes_buffer(1, N, C, 1, 4096, dev); p3 = planes_buffer(1, N, C, 1, 4096, dev)
pT = planes_buffer(1, N, C, 1, 4096, dev); pF = planes_buffer(1, N, C, 1, 4096, dev)
sums = torch.zeros(8, dtype=torch.float64, device=dev); db2 = torch.empty(C, device=dev); db3 = torch.empty(C, device=dev)
d = lambda t: t.data_ptr()
for _ in range(3):
    L.cascade_losses_planes(d(x), d(x1), d(x2), d(x3f), d(ys[0]), d(ys[1]), d(ys[2]), 1.0, 1, N, C, P, 1.0 / x.numel(), d(sums),
                            d(g1), d(p2), d(p3), d(db2), d(db3), st)
    L.residual_split_planes(d(x), d(x1), d(pT), d(pF), N, C, P, st)
torch.cuda.synchronize(); print("ok")
