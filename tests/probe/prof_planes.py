"""ncu probe: one launch of each first-layer down / wgrad instance (fp32-input and operand-plane) at N=1024."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from lshm_b200._lib import lib
from lshm_b200.engine import conv_image

dev = torch.device("cuda:0")
L = lib(); st = torch.cuda.current_stream().cuda_stream
N, A, Bc, s, l = 1024, 8, 8, 64, 4096
big = torch.randn(N, Bc, 128, 128, device=dev); small = torch.randn(N, A, s, s, device=dev)
out_s = torch.empty(N, A, s, s, device=dev)
w2 = torch.randn(A, Bc, 4, 4, device=dev) * 0.1; w1 = torch.randn(A, Bc, 4, device=dev) * 0.1
bias = torch.randn(A, device=dev); dw2, dw1 = torch.empty_like(w2), torch.empty_like(w1)
i2, i1 = conv_image(w2, 2, 0, st), conv_image(w1, 1, 0, st)
nb = ctypes.c_int64()
L.cdll.lshm_planes_bytes(2, N, Bc, s, s, ctypes.byref(nb)); p2 = torch.zeros(nb.value, dtype=torch.uint8, device=dev)
L.cdll.lshm_planes_bytes(1, N, Bc, 1, l, ctypes.byref(nb)); p1 = torch.zeros(nb.value, dtype=torch.uint8, device=dev)
d = lambda t: t.data_ptr()
L.stage_planes2d(d(big), Bc * 16384, d(p2), N, Bc, s, s, st)
L.stage_planes1d(d(big), Bc * 16384, d(p1), N, Bc, l, 1, st)
for _ in range(2):
    L.down1d_planes(d(p1), d(i1), d(bias), None, 0, d(out_s), A * l, N, A, Bc, l, 1, st)
    L.down1d(d(big), Bc * 16384, d(i1), d(bias), None, 0, d(out_s), A * l, N, A, Bc, l, 1, 1, st)
    L.down2d_planes(d(p2), d(i2), d(bias), None, 0, d(out_s), A * s * s, N, A, Bc, s, s, 1, st)
    L.wgrad2d_planes(d(small), A * s * s, d(p2), d(dw2), N, A, Bc, s, s, st)
    L.wgrad1d_planes(d(small), A * l, d(p1), d(dw1), N, A, Bc, l, st)
torch.cuda.synchronize()
print("ok")
