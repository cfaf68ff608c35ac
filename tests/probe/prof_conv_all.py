"""Profiling aid: the six conv kernels at the first-layer shape, two rounds (ncu: -k regex:igemm -s 6 -c 6)."""
import sys, torch
sys.path.insert(0, ".")
from lshm_b200._lib import lib
from lshm_b200.engine import conv_image
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
dev = torch.device("cuda:0")
st = torch.cuda.current_stream().cuda_stream
A, Bc, s, l = 8, 8, 64, 4096
big2 = torch.randn(N, Bc, 2 * s, 2 * s, device=dev); small2 = torch.randn(N, A, s, s, device=dev)
big1 = torch.randn(N, Bc, 4 * l, device=dev); small1 = torch.randn(N, A, l, device=dev)
w2 = torch.randn(A, Bc, 4, 4, device=dev) * 0.1; w1 = torch.randn(A, Bc, 4, device=dev) * 0.1
bias = torch.randn(A, device=dev)
wd2, wu2 = conv_image(w2, 2, 0, st), conv_image(w2, 2, 1, st)
wd1, wu1 = conv_image(w1, 1, 0, st), conv_image(w1, 1, 1, st)
dw2, dw1 = torch.empty_like(w2), torch.empty_like(w1)
L = lib()
for _ in range(2):
    L.down2d(big2.data_ptr(), Bc * 4 * s * s, wd2.data_ptr(), bias.data_ptr(), None, 0, small2.data_ptr(), A * s * s, N, A, Bc, s, s, 1, st)
    L.up2d(small2.data_ptr(), A * s * s, wu2.data_ptr(), bias.data_ptr(), None, 0, big2.data_ptr(), Bc * 4 * s * s, N, A, Bc, s, s, 1, st)
    L.wgrad2d(small2.data_ptr(), A * s * s, big2.data_ptr(), Bc * 4 * s * s, dw2.data_ptr(), N, A, Bc, s, s, st)
    L.down1d(big1.data_ptr(), Bc * 4 * l, wd1.data_ptr(), bias.data_ptr(), None, 0, small1.data_ptr(), A * l, N, A, Bc, l, 1, 1, st)
    L.up1d(small1.data_ptr(), A * l, wu1.data_ptr(), bias.data_ptr(), None, 0, big1.data_ptr(), Bc * 4 * l, N, A, Bc, l, 0, 1, st)
    L.wgrad1d(small1.data_ptr(), A * l, big1.data_ptr(), Bc * 4 * l, dw1.data_ptr(), N, A, Bc, l, 1, st)
torch.cuda.synchronize()
print("ok")
