"""Probe: device time of the dense-head products (back-to-back launches, CUDA events)."""
import sys, torch
sys.path.insert(0, ".")
from lshm_b200._lib import lib
dev = torch.device("cuda:0"); st = torch.cuda.current_stream().cuda_stream
L = lib(); N = 1024
def t(fn, n=200):
    for _ in range(10): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
for K, J in ((784, 16), (784, 32), (16, 16), (32, 32), (32, 768)):
    x = torch.randn(N, K, device=dev); w = torch.randn(J, K, device=dev); b = torch.randn(J, device=dev)
    y = torch.empty(N, J, device=dev); dz = torch.randn(N, J, device=dev); dx = torch.empty(N, K, device=dev)
    f = t(lambda: L.linear_fwd(x.data_ptr(), K, w.data_ptr(), b.data_ptr(), y.data_ptr(), J, N, K, J, 1, st))
    d = t(lambda: L.linear_bwd_data(dz.data_ptr(), J, w.data_ptr(), None, 0, None, 0, dx.data_ptr(), K, N, K, J, st))
    print(f"K={K} J={J}: fwd {f:.1f} us  bwd_data {d:.1f} us")
