"""Last transposed conv of a 1-D net at cfg2 size: separate weight / data gradient kernels against the fused one."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from lshm_b200._lib import lib
from lshm_b200.engine import conv_image, planes_buffer

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
big = torch.randn(N, Bc, 4 * l, device=dev); act = torch.nn.functional.elu(torch.randn(N, A, l, device=dev))
w = torch.randn(A, Bc, 4, device=dev) * 0.1
img = conv_image(w, 1, 0, st)
pl = planes_buffer(1, N, Bc, 1, l, dev)
L.stage_planes1d(d(big), Bc * 4 * l, d(pl), N, Bc, l, 0, st)
dz, dw = torch.empty(N, A, l, device=dev), torch.empty(A, Bc, 4, device=dev)
t_d = timeit(lambda: L.down1d_planes(d(pl), d(img), None, d(act), A * l, d(dz), A * l, N, A, Bc, l, 2, st))
t_w = timeit(lambda: L.wgrad1d_planes(d(act), A * l, d(pl), d(dw), N, A, Bc, l, st))
t_f = timeit(lambda: L.tconv_bwd1d_planes(d(act), A * l, d(pl), d(img), d(dz), A * l, d(dw), N, A, Bc, l, st))
mb = (pl.numel() + 2 * act.numel() * 4) / 1e6
print(f"down1d_planes {t_d:.1f} us, wgrad1d_planes {t_w:.1f} us, fused {t_f:.1f} us ({mb:.0f} MB algorithmic -> {mb / t_f:.2f} TB/s)")
