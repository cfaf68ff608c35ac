#!/usr/bin/env python
"""BASELINE config 4: K-harmonic kernel sweep (HBM / fp32-ALU roofline study), one GPU.

X ~ N(0,1) fp32 [N,L], M ~ U(0,1) [K,L], p=4.  Reports, per (N,K,L) and per
kernel (fwd loss, fused fwd+bwd, assignment), the CUDA-event time, algorithmic GB/s against the
measured HBM peak and fp32 TFLOP/s against 74.4 TFLOP/s (148 SM x 128 lanes x 2 x 1.965 GHz); the
bound is HBM for K <~ 16 and fp32 ALU beyond (direct-difference form).
Writes CSV to --out.  Inputs are larger than L2 (126 MB) for every N >= 1M, L >= 32.
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
from lshm_b200._lib import lib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/khm_sweep.csv")
    ap.add_argument("--big", action="store_true", help="also run N=100M at L=32")
    args = ap.parse_args()
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
    hbm, alu = peaks["hbm_gbs"], 74.4
    dev = torch.device("cuda:0")
    st = torch.cuda.current_stream().cuda_stream
    L_ = lib()
    rows = ["N,K,L,kernel,ms,GBps,frac_hbm,TFLOPs,frac_fp32,bound"]
    cases = [(n, k, l) for n in (1_000_000, 10_000_000) for k in (10, 64, 256, 1024) for l in (32, 64, 128, 256)
             if n * k * l <= 10_000_000 * 1024 * 64]
    if args.big:
        cases += [(100_000_000, 10, 32), (100_000_000, 64, 32)]
    for N, K, L in cases:
        g = torch.Generator(device=dev).manual_seed(0)
        X = torch.randn(N, L, device=dev, generator=g)
        M = torch.rand(K, L, device=dev, generator=g)
        acc = torch.zeros(1, dtype=torch.float64, device=dev)
        gX = torch.empty_like(X)
        gM = torch.zeros(K, L, device=dev)
        ids = torch.empty(N, dtype=torch.int32, device=dev)
        ops = {
            "fwd": (lambda: L_.khm_fwd(X.data_ptr(), L, M.data_ptr(), N, K, L, 4.0, acc.data_ptr(), None, st), 4.0 * N * L, 1),
            "fwd_bwd": (lambda: L_.khm_fwd_bwd(X.data_ptr(), L, M.data_ptr(), N, K, L, 4.0, 1e-3, acc.data_ptr(), gX.data_ptr(), L, 0, gM.data_ptr(), st), 8.0 * N * L, 3),
            "assign": (lambda: L_.khm_assign(X.data_ptr(), L, M.data_ptr(), N, K, L, ids.data_ptr(), st), 4.0 * N * L + 4.0 * N, 1),
        }
        for name, (fn, byts, passes) in ops.items():
            for _ in range(3):
                fn()
            reps = 5
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            gbs = byts / (ms * 1e-3) / 1e9
            tfl = (3.0 * L + 6) * K * N * passes / (ms * 1e-3) / 1e12
            bound = "hbm" if K <= 16 else "fp32"
            rows.append(f"{N},{K},{L},{name},{ms:.4f},{gbs:.1f},{gbs / hbm:.3f},{tfl:.2f},{tfl / alu:.3f},{bound}")
            print(rows[-1], flush=True)
        del X, gX, ids
        torch.cuda.empty_cache()
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    open(args.out, "w").write("\n".join(rows) + "\n")


if __name__ == "__main__":
    main()
