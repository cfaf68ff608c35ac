#!/usr/bin/env python
"""bench_conv.py - first-layer conv kernels at the cfg2 shape (N = 1024, 8 channels, 128x128 / 16384): the fp32-input
instances (producer warps gather + split) against the operand-plane instances (tensor-TMA box loads), plus the plane
writers.  CUDA events, L2 flushed by the working set (each tensor is 537 MB).  One CSV line per kernel."""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import torch

from lshm_b200._lib import lib
from lshm_b200.engine import conv_image


def main():
    dev = torch.device("cuda:0")
    L = lib()
    st = torch.cuda.current_stream().cuda_stream
    N, A, Bc, s, l = 1024, 8, 8, 64, 4096
    peak = 6551.4
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    big = torch.randn(N, Bc, 128, 128, device=dev)
    small = torch.randn(N, A, s, s, device=dev)
    out_s = torch.empty(N, A, s, s, device=dev)
    out_b = torch.empty_like(big)
    w2 = torch.randn(A, Bc, 4, 4, device=dev) * 0.1
    w1 = torch.randn(A, Bc, 4, device=dev) * 0.1
    bias = torch.randn(A, device=dev)
    dw2, dw1 = torch.empty_like(w2), torch.empty_like(w1)
    i2, i1 = conv_image(w2, 2, 0, st), conv_image(w1, 1, 0, st)
    nb = ctypes.c_int64()
    L.cdll.lshm_planes_bytes(2, N, Bc, s, s, ctypes.byref(nb)); p2 = torch.zeros(nb.value, dtype=torch.uint8, device=dev)
    L.cdll.lshm_planes_bytes(1, N, Bc, 1, l, ctypes.byref(nb)); p1 = torch.zeros(nb.value, dtype=torch.uint8, device=dev)
    p1b = torch.zeros_like(p1)
    x1 = torch.randn_like(big); g3 = torch.randn_like(big)
    db = torch.empty(Bc, device=dev)
    f32 = 4.0 * big.numel()
    sm = 4.0 * small.numel()
    d = lambda t: t.data_ptr()
    cases = [
        ("stage_planes2d", lambda: L.stage_planes2d(d(big), Bc * 16384, d(p2), N, Bc, s, s, st), f32 + p2.numel()),
        ("stage_planes1d[pad=1]", lambda: L.stage_planes1d(d(big), Bc * 16384, d(p1), N, Bc, l, 1, st), f32 + p1.numel()),
        ("down2d fp32-in", lambda: L.down2d(d(big), Bc * 16384, d(i2), d(bias), None, 0, d(out_s), A * s * s, N, A, Bc, s, s, 1, st), f32 + sm),
        ("down2d planes", lambda: L.down2d_planes(d(p2), d(i2), d(bias), None, 0, d(out_s), A * s * s, N, A, Bc, s, s, 1, st), p2.numel() + sm),
        ("down1d fp32-in[pad=1]", lambda: L.down1d(d(big), Bc * 16384, d(i1), d(bias), None, 0, d(out_s), A * l, N, A, Bc, l, 1, 1, st), f32 + sm),
        ("down1d planes", lambda: L.down1d_planes(d(p1), d(i1), d(bias), None, 0, d(out_s), A * l, N, A, Bc, l, 1, st), p1.numel() + sm),
        ("wgrad2d fp32-in", lambda: L.wgrad2d(d(small), A * s * s, d(big), Bc * 16384, d(dw2), N, A, Bc, s, s, st), f32 + sm),
        ("wgrad2d planes", lambda: L.wgrad2d_planes(d(small), A * s * s, d(p2), d(dw2), N, A, Bc, s, s, st), p2.numel() + sm),
        ("wgrad1d fp32-in[pad=1]", lambda: L.wgrad1d(d(small), A * l, d(big), Bc * 16384, d(dw1), N, A, Bc, l, 1, st), f32 + sm),
        ("wgrad1d planes", lambda: L.wgrad1d_planes(d(small), A * l, d(p1), d(dw1), N, A, Bc, l, st), p1.numel() + sm),
        ("residual_split", lambda: L.residual_split(d(big), d(x1), d(out_b), d(g3), N, Bc, 128, st), 4 * f32),
        ("residual_split_planes", lambda: L.residual_split_planes(d(big), d(x1), d(p1), d(p1b), N, Bc, 128, st), 2 * f32 + 2 * p1.numel()),
        ("cascade_combine", lambda: L.cascade_combine(d(big), d(x1), d(g3), d(out_b), N, Bc, 128, d(db), st), 4 * f32),
        ("cascade_combine_planes", lambda: L.cascade_combine_planes(d(big), d(x1), d(g3), d(p2), N, Bc, 128, d(db), st), 3 * f32 + p2.numel()),
    ]
    print("kernel,ms,GB/s(algorithmic),frac_of_hbm_peak")
    for name, fn, byts in cases:
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        gbs = byts / (ms * 1e-3) / 1e9
        print(f"{name},{ms:.4f},{gbs:.1f},{gbs / peak:.3f}", flush=True)


if __name__ == "__main__":
    main()
