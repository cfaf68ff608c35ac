#!/usr/bin/env python
"""Inference throughput of the clustering path (BASELINE.json configs[2]): cascade encode of 8-channel
128x128 patches + K-harmonic distances / assignment with K=64 centres, patches sharded over the GPUs
(no data-path collective: every rank clusters its own baselines, src/evaluate_clustering.py:75-119).

    python bench_infer.py [--chunk 1024] [--chunks 8]                 # one GPU
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 bench_infer.py

One "chunk" is `--chunk` patches = chunk/4 baselines of 2x2 patches, loaded from pinned host int8 through the
patchify kernels, encoded by the three autoencoders and assigned.  10M patches do not fit in HBM at once
(5.2 TB as fp32), so the job streams chunks: with `--total P` (default 10 000 000) every rank streams
ceil(P / world / chunk) chunks and the line reports the MEASURED wall time of the whole job (barrier to barrier,
max over ranks, including the read-back of every baseline's cluster id) - not an extrapolation.  `--chunks n`
times n chunks per GPU instead (quick runs).  Model = the cfg model (4 harmonic scales, src/kharmonic_lofar.py:57).
Prints one JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
SCALES = [1e-4, 1e-3, 1e-2, 1e-1]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chunk", type=int, default=1024, help="patches per chunk and GPU (multiple of 4)")
    ap.add_argument("--chunks", type=int, default=0, help="timed chunks per GPU (0: derive from --total)")
    ap.add_argument("--total", type=int, default=10_000_000, help="patches of the whole job (all GPUs)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--K", type=int, default=64)
    args = ap.parse_args()
    from lshm_b200 import lofar_tools as T
    from lshm_b200 import synthetic as S
    from lshm_b200._lib import lib
    from lshm_b200.evaluate_clustering import encode_assign
    from lshm_b200.lofar_models import AutoEncoder1DCNN, AutoEncoderCNN2, Kmeans

    import time
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if args.chunks <= 0:
        args.chunks = -(-args.total // (world * args.chunk))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import datetime
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))
    torch.manual_seed(0)
    C, L, Lt = 8, 32, 16
    hs = torch.tensor(SCALES).to(dev)
    net = AutoEncoderCNN2(L, C, hs, True).to(dev)
    netT = AutoEncoder1DCNN(Lt, C, hs, True).to(dev)
    netF = AutoEncoder1DCNN(Lt, C, hs, True).to(dev)
    mod = Kmeans(L + 2 * Lt, args.K, 4.0).to(dev)
    nb = args.chunk // 4
    meas = S.make_measurement(nb, 192, 192, seed=200 + rank)
    sap = meas["measurement"]["saps"]["0"]
    vis_h = torch.from_numpy(sap["visibilities"]).pin_memory()
    sc_h = torch.from_numpy(sap["visibility_scale_factors"]).pin_memory()
    uv_h = torch.from_numpy(S.make_uv(args.chunk, seed=rank, per_group=4)).pin_memory()
    sel = torch.arange(nb, dtype=torch.int32, device=dev)
    L_ = lib()

    # double-buffered ingest: the next chunk's H2D copy + loader kernels run on a side stream while this
    # chunk is encoded (preallocated buffers, nothing is allocated in the loop)
    sets = [(torch.empty_like(vis_h, device=dev), torch.empty_like(sc_h, device=dev), torch.empty_like(uv_h, device=dev),
             torch.empty(args.chunk, C, 128, 128, device=dev), torch.zeros(2, dtype=torch.float64, device=dev))
            for _ in range(2)]

    def load(k):
        vis, sc, uv, y, stats = sets[k % 2]
        vis.copy_(vis_h, non_blocking=True); sc.copy_(sc_h, non_blocking=True); uv.copy_(uv_h, non_blocking=True)
        px, py, x = T.patchify_device(vis, sc, sel, 128, C, 1e3, True, out=y, stats=stats)
        return px * py, x, uv

    pf = T.DevicePrefetcher(dev, record_streams=False)

    def run(nchunks, out=None):
        gid = None
        pf.submit(lambda: load(0))
        for k in range(nchunks):
            bpb, x, uv = pf.get()
            if k + 1 < nchunks:
                pf.submit(lambda k=k: load(k + 1))
            dist_, gid, ids, Mu = encode_assign(net, netT, netF, mod, x, uv, bpb)
            if out is not None:
                out[k * nb:(k + 1) * nb].copy_(gid)           # every baseline's cluster id stays on the device ...
        return gid

    run(max(args.warmup, 3))
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    l0 = L_.launches
    all_ids = torch.empty(args.chunks * nb, dtype=torch.int32, device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    run(args.chunks, all_ids)
    host_ids = all_ids.cpu()                   # ... and is read back at the end (what a caller gets)
    e1.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    sec = e0.elapsed_time(e1) * 1e-3
    launches = L_.launches - l0
    if world > 1:
        t = torch.tensor([sec, wall], device=dev)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        sec, wall = float(t[0]), float(t[1])
    if rank == 0:
        total = args.chunk * args.chunks * world
        print(json.dumps({
            "metric": "inference patches/sec (encode + K-harmonic assignment)", "value": total / sec,
            "unit": "patches/s", "n_gpus": world, "chunks": args.chunks, "ms_per_chunk": sec / args.chunks * 1e3,
            "higher_is_better": True, "scaling": "strong (a fixed job of --total patches)", "dtype": "f32", "data": "synthetic",
            "config": {"workload": "cfg3 evaluate_clustering: cascade encode + assignment", "K": args.K,
                       "chunk_patches_per_gpu": args.chunk, "channels": C, "L": L, "Lt": Lt,
                       "inputs": "pinned host int8 -> patchify kernels every chunk (larger than L2), next chunk staged on a side stream"},
            "gpu_launches": int(launches), "patches_processed": int(total), "job_seconds_device": sec,
            "job_seconds_wall": wall, "baselines_assigned": int(host_ids.numel()) * world,
            "cluster_histogram_rank0": torch.bincount(host_ids.long(), minlength=args.K)[:8].tolist(),
            "baseline_ids_head": host_ids[:4].tolist(),
        }))
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
