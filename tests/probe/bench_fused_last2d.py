"""Last transposed conv of the 2-D net at cfg2 size: separate weight / data gradient kernels against the fused one."""
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
N, A, Bc, s = 1024, 8, 8, 64
d = lambda t: t.data_ptr()
big = torch.randn(N, Bc, 2 * s, 2 * s, device=dev); act = torch.nn.functional.elu(torch.randn(N, A, s, s, device=dev))
w = torch.randn(A, Bc, 4, 4, device=dev) * 0.1
img = conv_image(w, 2, 0, st)
pl = planes_buffer(2, N, Bc, s, s, dev)
L.stage_planes2d(d(big), Bc * 4 * s * s, d(pl), N, Bc, s, s, st)
dz, dw = torch.empty(N, A, s, s, device=dev), torch.empty(A, Bc, 4, 4, device=dev)
ns = A * s * s
t_d = timeit(lambda: L.down2d_planes(d(pl), d(img), None, d(act), ns, d(dz), ns, N, A, Bc, s, s, 2, st))
t_w = timeit(lambda: L.wgrad2d_planes(d(act), ns, d(pl), d(dw), N, A, Bc, s, s, st))
t_f = timeit(lambda: L.tconv_bwd2d_planes(d(act), ns, d(pl), d(img), d(dz), ns, d(dw), N, A, Bc, s, s, st))
mb = (pl.numel() + 2 * act.numel() * 4) / 1e6
print(f"down2d_planes {t_d:.1f} us, wgrad2d_planes {t_w:.1f} us, fused {t_f:.1f} us ({mb:.0f} MB algorithmic -> {mb / t_f:.2f} TB/s)")
