#!/usr/bin/env python
"""Roofline numbers for the loader and Fourier-feature kernels (one GPU, CUDA events).
Inputs larger than L2: 1024 patches x 8 channels x 128 x 128 fp32 = 537 MB."""
import json, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.abspath(__file__)); sys.path.insert(0, ROOT)
from lshm_b200 import lofar_tools as T
from lshm_b200 import synthetic as S

def timeit(fn, reps=5, warm=3):
    for _ in range(warm): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

def main():
    pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    dev = torch.device("cuda:0")
    out = {}
    N, C = 1024, 8
    x = torch.randn(N, C, 128, 128, device=dev); xh = torch.randn(N, C, 128, 128, device=dev)
    ms = timeit(lambda: T.fft_features(x))
    b = N * C * 196608.0
    out["fft2_reim_shift_clamp"] = dict(ms=ms, gbs=b / ms / 1e6, frac=b / ms / 1e6 / pk, planes_per_s=N * C / ms * 1e3)
    ms = timeit(lambda: T.fft_features(x, xh))
    b = N * C * (196608.0 + 65536.0)
    out["fft2_reim_shift_clamp(x-xhat)"] = dict(ms=ms, gbs=b / ms / 1e6, frac=b / ms / 1e6 / pk)
    meas = S.make_measurement(256, 192, 192, seed=0)["measurement"]["saps"]["0"]
    vis = torch.from_numpy(meas["visibilities"]).to(dev); sc = torch.from_numpy(meas["visibility_scale_factors"]).to(dev)
    sel = torch.arange(256, dtype=torch.int32, device=dev)
    ms = timeit(lambda: T.patchify_device(vis, sc, sel, 128, 8, 1e3, False))
    b = 1024 * 8 * 16384 * 4.25
    out["patchify_scale_i8"] = dict(ms=ms, gbs=b / ms / 1e6, frac=b / ms / 1e6 / pk)
    ms = timeit(lambda: T.patchify_device(vis, sc, sel, 128, 8, 1e3, True))
    b = 1024 * 8 * 16384 * (4.25 + 8)
    out["patchify+normalise"] = dict(ms=ms, gbs=b / ms / 1e6, frac=b / ms / 1e6 / pk)
    print(json.dumps(out, indent=1))

if __name__ == "__main__":
    main()
