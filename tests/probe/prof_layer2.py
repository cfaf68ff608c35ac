"""ncu probe: the six conv kernels at the second-layer shape (8 <-> 12 channels, 32x32 <-> 64x64 / 1024 <-> 4096, N=1024),
plus the gradient-combine plane writer."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from lshm_b200._lib import lib
from lshm_b200.engine import conv_image, planes_buffer

dev = torch.device("cuda:0")
L = lib(); st = torch.cuda.current_stream().cuda_stream
N, A, Bc, s, l = 1024, 12, 8, 32, 1024
big2 = torch.randn(N, Bc, 64, 64, device=dev); small2 = torch.randn(N, A, s, s, device=dev)
w2 = torch.randn(A, Bc, 4, 4, device=dev) * 0.1; w1 = torch.randn(A, Bc, 4, device=dev) * 0.1
ba, bb = torch.randn(A, device=dev), torch.randn(Bc, device=dev)
dw2, dw1 = torch.empty_like(w2), torch.empty_like(w1)
d2, u2 = conv_image(w2, 2, 0, st), conv_image(w2, 2, 1, st)
d1, u1 = conv_image(w1, 1, 0, st), conv_image(w1, 1, 1, st)
out_s, out_b = torch.empty_like(small2), torch.empty_like(big2)
d = lambda t: t.data_ptr()
g1p, gT, gF = (torch.randn(N, 8, 128, 128, device=dev) for _ in range(3))
pl = planes_buffer(2, N, 8, 64, 64, dev); db = torch.empty(8, device=dev)
for _ in range(2):
    L.down2d(d(big2), Bc * 4096, d(d2), d(ba), None, 0, d(out_s), A * s * s, N, A, Bc, s, s, 1, st)
    L.up2d(d(small2), A * s * s, d(u2), d(bb), None, 0, d(out_b), Bc * 4096, N, A, Bc, s, s, 1, st)
    L.wgrad2d(d(small2), A * s * s, d(big2), Bc * 4096, d(dw2), N, A, Bc, s, s, st)
    L.down1d(d(big2), Bc * 4096, d(d1), d(ba), None, 0, d(out_s), A * l, N, A, Bc, l, 1, 1, st)
    L.up1d(d(small2), A * l, d(u1), d(bb), None, 0, d(out_b), Bc * 4096, N, A, Bc, l, 0, 1, st)
    L.wgrad1d(d(small2), A * l, d(big2), Bc * 4096, d(dw1), N, A, Bc, l, 1, st)
    L.cascade_combine_planes(d(g1p), d(gT), d(gF), d(pl), N, 8, 128, d(db), st)
torch.cuda.synchronize()
print("ok")
