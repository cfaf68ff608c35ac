"""Profiling aid: conv0-shaped 1-D down kernel."""
import sys, torch
sys.path.insert(0, ".")
from lshm_b200._lib import lib
from lshm_b200.engine import conv_image
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
dev = torch.device("cuda:0")
st = torch.cuda.current_stream().cuda_stream
A, Bc, l = 8, 8, 4096
big = torch.randn(N, Bc, 4 * l, device=dev)
small = torch.empty(N, A, l, device=dev)
w = torch.randn(A, Bc, 4, device=dev) * 0.1
bias = torch.randn(A, device=dev)
wd = conv_image(w, 1, 0, st)
for _ in range(5):
    lib().down1d(big.data_ptr(), Bc * 4 * l, wd.data_ptr(), bias.data_ptr(), None, 0, small.data_ptr(), A * l, N, A, Bc, l, 1, 1, st)
torch.cuda.synchronize()
print("ok")
