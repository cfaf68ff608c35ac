#!/usr/bin/env python
"""bench_lbfgs.py - BASELINE cfg5: full cascaded training with the LBFGSNew closure (multiple fwd / bwd per step) on
a 62-baseline synthetic observation (62 x 2x2 patches = 248 patches), data-parallel over N GPUs (STRONG scaling: the
248 patches are sharded by baseline group; one all-reduce of [gradients | loss scalars] per gradient closure, of
the 16 loss scalars per line-search closure).

    python bench_lbfgs.py                       # 1 GPU
    torchrun --nproc-per-node 8 bench_lbfgs.py  # 8 GPUs

One "iteration" = optimizer.step(closure) + multiplier update (src/kharmonic_lofar.py:131-202) with
LBFGSNew(history_size=7, max_iter=4, line_search_fn=True, batch_mode=True) over all parameters (:93).
Arms: `flat` = lshm_b200.lbfgsnew.LBFGSNew on the flat buffer (vector ops as views, cached f_old probes, CUDA-graph
replays), `reference_optimizer` = the unmodified reference LBFGSNew (baseline/_ref) driving the same fused closure,
`cpu_reference` (rank 0, 1 GPU only) = reference modules + reference optimiser on the host cores.
Prints one JSON line per arm."""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch

SCALES = [1e-4, 1e-3, 1e-2, 1e-1]
NB, BPB, C, L, LT, K, P = 62, 4, 8, 32, 16, 10, 4
LBFGS_KW = dict(history_size=7, max_iter=4, line_search_fn=True, batch_mode=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--cpu", default="on", choices=["on", "off"])
    args = ap.parse_args()
    from lshm_b200 import parallel
    from lshm_b200 import synthetic as S
    from lshm_b200._lib import lib
    from lshm_b200.kharmonic_lofar import DeepKHarmonicStep
    from lshm_b200.lbfgsnew import LBFGSNew
    from lshm_b200.lofar_models import AutoEncoder1DCNN, AutoEncoderCNN2, Kmeans
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import datetime
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))
    Np = NB * BPB
    x_all = torch.from_numpy(S.make_patches(Np, C, seed=5))
    uv_all = torch.from_numpy(S.make_uv(Np, seed=5, per_group=BPB))
    r0, r1 = parallel.shard_rows(Np, BPB, rank, world)

    def build():
        torch.manual_seed(0)
        hs = torch.tensor(SCALES).to(dev)
        mods = [AutoEncoderCNN2(L, C, hs, True).to(dev), AutoEncoder1DCNN(LT, C, hs, True).to(dev),
                AutoEncoder1DCNN(LT, C, hs, True).to(dev), Kmeans(L + 2 * LT, K, P).to(dev)]
        step = DeepKHarmonicStep(*mods, distributed=world > 1)
        step.set_batch(x_all[r0:r1].to(dev), uv_all[r0:r1].to(dev), BPB, global_patches=Np)
        return step

    def run(step, opt, label, graphs):
        if graphs:
            step.enable_graphs()
        calls = [0, 0]

        def closure():
            calls[0 if torch.is_grad_enabled() else 1] += 1
            return step.closure()

        def one():
            loss = opt.step(closure)
            step.update_multipliers()
            return loss
        for _ in range(args.warmup):
            one()
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()
        calls[:] = [0, 0]
        l0 = lib().launches
        t0 = time.perf_counter()          # the optimiser reads every loss on the host: wall clock = device time + syncs
        for _ in range(args.iters):
            loss = one()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], device=dev)
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            dt = float(t)
        per = dt / args.iters
        nclos = (calls[0] + calls[1]) / args.iters
        if rank == 0:
            print(json.dumps(dict(
                bench="cfg5_lbfgs", arm=label, n_gpus=world, global_patches=Np, patches_per_gpu=r1 - r0,
                ms_per_iteration=per * 1e3, iterations_per_s=1.0 / per, grad_closures_per_iteration=calls[0] / args.iters,
                linesearch_closures_per_iteration=calls[1] / args.iters,
                closure_patches_per_s=Np * nclos / per, ms_per_closure=per * 1e3 / nclos,
                launches_per_iteration=(lib().launches - l0) / args.iters, cuda_graphs=bool(graphs),
                last_loss=float(loss), scaling="strong", optimizer="LBFGSNew(history_size=7,max_iter=4,line_search_fn=True,batch_mode=True)")),
                flush=True)

    s = build()
    run(s, LBFGSNew(s.flat, **LBFGS_KW), "flat", True)
    del s
    torch.cuda.empty_cache()
    from oracle import reference_loop as RL       # the reference optimiser (unmodified) as a client of the closure
    if RL.reference_dir() is not None:
        Ref = RL.load_reference_module("lbfgsnew").LBFGSNew
        s = build()
        run(s, Ref(s.flat.params, **LBFGS_KW), "reference_optimizer", False)
        del s
        if world == 1 and rank == 0 and args.cpu == "on":
            torch.set_num_threads(os.cpu_count() or 1)
            R = RL.ReferenceLoop(L=L, Lt=LT, C=C, K=K, Khp=P, optimizer="lbfgs")
            R.set_batch(x_all, uv_all, BPB)
            R.admm_iteration()
            c0 = R.closures
            t0 = time.perf_counter()
            R.admm_iteration()
            dt = time.perf_counter() - t0
            print(json.dumps(dict(bench="cfg5_lbfgs", arm="cpu_reference", cores=os.cpu_count(), global_patches=Np,
                                  ms_per_iteration=dt * 1e3, closures_per_iteration=R.closures - c0,
                                  closure_patches_per_s=Np * (R.closures - c0) / dt,
                                  note="unmodified reference modules + LBFGSNew, restated script loop, 1 warm-up + 1 timed iteration")),
                  flush=True)
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
