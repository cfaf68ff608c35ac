"""Profiling aid: K-harmonic forward+backward at the HBM-side corner of the sweep (K=10, L=64)."""
import sys, torch
sys.path.insert(0, ".")
from lshm_b200._lib import lib
N, K, L = 4_000_000, 10, 64
dev = torch.device("cuda:0"); st = torch.cuda.current_stream().cuda_stream
X = torch.randn(N, L, device=dev); M = torch.rand(K, L, device=dev)
acc = torch.zeros(1, dtype=torch.float64, device=dev)
gX = torch.empty_like(X); gM = torch.zeros(K, L, device=dev)
for _ in range(4):
    lib().khm_fwd_bwd(X.data_ptr(), L, M.data_ptr(), N, K, L, 4.0, 1e-3, acc.data_ptr(), gX.data_ptr(), L, 0, gM.data_ptr(), st)
torch.cuda.synchronize(); print("ok")
