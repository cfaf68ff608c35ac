"""First / second / third layer conv kernels at N = 1024 (fp32-input instances), back to back, CUDA events."""
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

def main():
    dev = torch.device("cuda:0"); L = lib(); st = torch.cuda.current_stream().cuda_stream
    N = 1024
    d = lambda t: t.data_ptr()
    print("dim,A(small ch),Bc(big ch),small,up_us,down_us,wgrad_us,MB")
    for dim in (1, 2):
        for (A, Bc, lvl) in ((8, 8, 1), (12, 8, 2), (24, 12, 3)):
            if dim == 2:
                s = 128 >> lvl; big = torch.randn(N, Bc, 2 * s, 2 * s, device=dev); small = torch.randn(N, A, s, s, device=dev)
                w = torch.randn(A, Bc, 4, 4, device=dev) * 0.1
            else:
                s = 16384 >> (2 * lvl); big = torch.randn(N, Bc, 4 * s, device=dev); small = torch.randn(N, A, s, device=dev)
                w = torch.randn(A, Bc, 4, device=dev) * 0.1
            bias = torch.randn(Bc, device=dev); bias_a = torch.randn(A, device=dev)
            iu, idn = conv_image(w, dim, 1, st), conv_image(w, dim, 0, st)
            dw = torch.empty_like(w)
            bns, sns = big[0].numel(), small[0].numel()
            ob, os_ = torch.empty_like(big), torch.empty_like(small)
            if dim == 2:
                fu = lambda: L.up2d(d(small), sns, d(iu), d(bias), None, 0, d(ob), bns, N, A, Bc, s, s, 1, st)
                fd = lambda: L.down2d(d(big), bns, d(idn), d(bias_a), None, 0, d(os_), sns, N, A, Bc, s, s, 1, st)
                fw = lambda: L.wgrad2d(d(small), sns, d(big), bns, d(dw), N, A, Bc, s, s, st)
            else:
                fu = lambda: L.up1d(d(small), sns, d(iu), d(bias), None, 0, d(ob), bns, N, A, Bc, s, 0, 1, st)
                fd = lambda: L.down1d(d(big), bns, d(idn), d(bias_a), None, 0, d(os_), sns, N, A, Bc, s, 0, 1, st)
                fw = lambda: L.wgrad1d(d(small), sns, d(big), bns, d(dw), N, A, Bc, s, 0, st)
            print(f"{dim},{A},{Bc},{s},{timeit(fu):.1f},{timeit(fd):.1f},{timeit(fw):.1f},{(big.numel() + small.numel()) * 4 / 1e6:.0f}", flush=True)

if __name__ == "__main__":
    main()
