"""First-layer 1-D up kernel: transposed-conv forward (pad 0) against the conv data gradient (pad 1), no activation."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from lshm_b200._lib import lib
from lshm_b200.engine import conv_image

def timeit(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3

dev = torch.device("cuda:0"); L = lib(); st = torch.cuda.current_stream().cuda_stream
N, A, Bc, l = 1024, 8, 8, 4096
d = lambda t: t.data_ptr()
small = torch.randn(N, A, l, device=dev); big = torch.empty(N, Bc, 4 * l, device=dev)
w = torch.randn(A, Bc, 4, device=dev) * 0.1
img = conv_image(w, 1, 1, st)
for pad in (0, 1):
    t = timeit(lambda: L.up1d(d(small), A * l, d(img), None, None, 0, d(big), Bc * 4 * l, N, A, Bc, l, pad, 0, st))
    print(f"up1d first layer pad={pad} epilogue none: {t:.1f} us ({(small.numel() + big.numel()) * 4 / t / 1e6:.2f} TB/s)")
