"""Profiling aid: FFT features and the K-harmonic forward (ncu: -k regex:'fft2_kernel|khm_pass1' -s 4 -c 2)."""
import sys, torch
sys.path.insert(0, ".")
from lshm_b200._lib import lib
dev = torch.device("cuda:0"); st = torch.cuda.current_stream().cuda_stream
N, K, L = 4_000_000, 10, 64
X = torch.randn(N, L, device=dev); M = torch.rand(K, L, device=dev)
acc = torch.zeros(1, dtype=torch.float64, device=dev)
P, C = 2048, 8
x = torch.randn(P, C, 128, 128, device=dev); xhat = torch.randn(P, C, 128, 128, device=dev)
out = torch.empty(P, 2 * C, 128, 128, device=dev)
for _ in range(3):
    lib().fft2_reim_shift_clamp(x.data_ptr(), xhat.data_ptr(), out.data_ptr(), P, C, 1e3, st)
    lib().khm_fwd(X.data_ptr(), L, M.data_ptr(), N, K, L, 4.0, acc.data_ptr(), None, st)
torch.cuda.synchronize(); print("ok")
