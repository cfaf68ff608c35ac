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
    ap.add_argument("--big", action="store_true", help="also run N=100M (every L that fits)")
    ap.add_argument("--quick", action="store_true", help="K in {10,64} only")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        import datetime
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local),
                                             timeout=datetime.timedelta(seconds=300))
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
    hbm, alu = peaks["hbm_gbs"], 74.4
    dev = torch.device("cuda", local)
    st = torch.cuda.current_stream().cuda_stream
    L_ = lib()
    # N = GLOBAL number of latents, sharded evenly over the ranks (no data-path collective; the K x L gradient /
    # loss partial sums would ride in the step's one all-reduce); GB/s and TFLOP/s are whole-job aggregates
    # against world x the single-GPU peaks.
    rows = ["n_gpus,N,K,L,kernel,ms,GBps,frac_hbm,TFLOPs,frac_fp32,bound"]
    ks = (10, 64) if args.quick else (10, 64, 256, 1024)
    cases = [(n, k, l) for n in (1_000_000, 10_000_000) for k in ks for l in (32, 64, 128, 256)
             if n // world * k * l <= 10_000_000 * 1024 * 64]
    if args.big:
        cases += [(100_000_000, k, l) for k in (10, 64) for l in (32, 64, 128, 256)
                  if 100_000_000 // world * l * 8 <= 120e9 and 100_000_000 // world * k * l <= 10_000_000 * 1024 * 64]
    Nglob = None
    for Nglob, K, L in cases:
        N = Nglob // world
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
            if world > 1:
                torch.distributed.barrier()
            torch.cuda.synchronize()
            e0.record()
            for _ in range(reps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            if world > 1:
                t = torch.tensor([ms], device=dev)
                torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
                ms = float(t)
            gbs = byts * world / (ms * 1e-3) / 1e9
            tfl = (3.0 * L + 6) * K * N * world * passes / (ms * 1e-3) / 1e12
            bound = "hbm" if K <= 16 else "fp32"
            rows.append(f"{world},{Nglob},{K},{L},{name},{ms:.4f},{gbs:.1f},{gbs / (hbm * world):.3f},{tfl:.2f},{tfl / (alu * world):.3f},{bound}")
            if rank == 0:
                print(rows[-1], flush=True)
        del X, gX, ids
        torch.cuda.empty_cache()
    if rank == 0:
        os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
        open(args.out, "w").write("\n".join(rows) + "\n")
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
