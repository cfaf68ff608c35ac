"""Profiling aid: K-harmonic forward at L = 32 (rows fetched as consecutive words, transposed per warp)."""
import sys, torch
sys.path.insert(0, ".")
from lshm_b200._lib import lib
N, K, L = 8_000_000, 10, 32
dev = torch.device("cuda:0"); st = torch.cuda.current_stream().cuda_stream
X = torch.randn(N, L, device=dev); M = torch.rand(K, L, device=dev)
acc = torch.zeros(1, dtype=torch.float64, device=dev)
for _ in range(4):
    lib().khm_fwd(X.data_ptr(), L, M.data_ptr(), N, K, L, 4.0, acc.data_ptr(), None, st)
torch.cuda.synchronize(); print("ok")
