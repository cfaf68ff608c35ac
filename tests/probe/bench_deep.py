"""Deep conv layers at N = 1024: fp32-input 'down' instances against stage_planes + operand-plane instances."""
import ctypes, os, sys
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

def main():
    dev = torch.device("cuda:0"); L = lib(); st = torch.cuda.current_stream().cuda_stream
    N = 1024
    d = lambda t: t.data_ptr()
    print("dim,A,Bc,small,down_us,stage_us,down_planes_us,max_abs_diff")
    for dim in (2, 1):
        for (A, Bc, lvl) in ((12, 8, 2), (24, 12, 3), (48, 24, 4), (96, 48, 5), (192, 96, 6)):
            if dim == 2:
                s = 128 >> lvl; big = torch.randn(N, Bc, 2 * s, 2 * s, device=dev); small = torch.empty(N, A, s, s, device=dev)
                w = torch.randn(A, Bc, 4, 4, device=dev) * 0.1
            else:
                s = 16384 >> (2 * lvl); big = torch.randn(N, Bc, 4 * s, device=dev); small = torch.empty(N, A, s, device=dev)
                w = torch.randn(A, Bc, 4, device=dev) * 0.1
            small2 = torch.empty_like(small)
            bias = torch.randn(A, device=dev)
            img = conv_image(w, dim, 0, st)
            nb = ctypes.c_int64()
            L.cdll.lshm_planes_bytes(dim, N, Bc, s if dim == 2 else 1, s, ctypes.byref(nb))
            pl = torch.zeros(nb.value, dtype=torch.uint8, device=dev)
            bns, sns = big[0].numel(), small[0].numel()
            if dim == 2:
                f0 = lambda: L.down2d(d(big), bns, d(img), d(bias), None, 0, d(small), sns, N, A, Bc, s, s, 1, st)
                f1 = lambda: L.stage_planes2d(d(big), bns, d(pl), N, Bc, s, s, st)
                f2 = lambda: L.down2d_planes(d(pl), d(img), d(bias), None, 0, d(small2), sns, N, A, Bc, s, s, 1, st)
            else:
                f0 = lambda: L.down1d(d(big), bns, d(img), d(bias), None, 0, d(small), sns, N, A, Bc, s, 0, 1, st)
                f1 = lambda: L.stage_planes1d(d(big), bns, d(pl), N, Bc, s, 0, st)
                f2 = lambda: L.down1d_planes(d(pl), d(img), d(bias), None, 0, d(small2), sns, N, A, Bc, s, 1, st)
            t0, t1, t2 = timeit(f0), timeit(f1), timeit(f2)
            print(f"{dim},{A},{Bc},{s},{t0:.1f},{t1:.1f},{t2:.1f},{float((small - small2).abs().max()):.2e}", flush=True)

if __name__ == "__main__":
    main()
