"""Profiling aid: runs one conv shape through the tensor-core kernels a few times (for ncu)."""
import sys, torch
sys.path.insert(0, ".")
from lshm_b200._lib import lib
from lshm_b200.engine import conv_image
which = sys.argv[1] if len(sys.argv) > 1 else "down2d"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
dev = torch.device("cuda:0")
st = torch.cuda.current_stream().cuda_stream
A, Bc, s = 8, 8, 64
big = torch.randn(N, Bc, 2 * s, 2 * s, device=dev)
small = torch.randn(N, A, s, s, device=dev)
w = torch.randn(A, Bc, 4, 4, device=dev) * 0.1
bias = torch.randn(A, device=dev)
wd, wu = conv_image(w, 2, 0, st), conv_image(w, 2, 1, st)
dw = torch.empty_like(w)
for _ in range(5):
    if which == "down2d":
        lib().down2d(big.data_ptr(), Bc * 4 * s * s, wd.data_ptr(), bias.data_ptr(), None, 0, small.data_ptr(), A * s * s, N, A, Bc, s, s, 1, st)
    elif which == "up2d":
        lib().up2d(small.data_ptr(), A * s * s, wu.data_ptr(), bias.data_ptr(), None, 0, big.data_ptr(), Bc * 4 * s * s, N, A, Bc, s, s, 1, st)
    elif which == "wgrad2d":
        lib().wgrad2d(small.data_ptr(), A * s * s, big.data_ptr(), Bc * 4 * s * s, dw.data_ptr(), N, A, Bc, s, s, st)
torch.cuda.synchronize()
print("ok")
