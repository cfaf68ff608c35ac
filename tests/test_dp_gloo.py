"""Data-parallel host logic on CPU with gloo, world_size 2: shard by baseline group, compute each
shard's partial closure with the GLOBAL divisors, ONE all-reduce of [gradients | loss scalars],
compare with the unsharded closure.  (The GPU path uses the same ShardPlan / exchange with NCCL.)"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from common import SCALES, closure_case, oracle_closure
from lshm_b200 import parallel
from oracle import lofar_oracle as O


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    try:
        case = closure_case(N=8, bpb=2)
        r0, r1 = parallel.shard_rows(case["N"], case["bpb"], rank, world)
        npix = case["C"] * 16384
        hs = torch.tensor(SCALES)
        params = []
        ps = []
        for key in ("pn", "pT", "pF"):
            d = {k: v.clone().requires_grad_() for k, v in case[key].items()}
            ps.append(d)
            params += list(d.values())
        M = case["M"].clone().requires_grad_()
        params.append(M)
        ys = [y[r0 * npix:r1 * npix] for y in case["ys"]]
        total, terms = O.closure_losses(ps[0], ps[1], ps[2], M, case["x"][r0:r1], case["uv"][r0:r1], hs, *ys,
                                        batch_per_bline=case["bpb"], batch_size=(r1 - r0) // case["bpb"],
                                        shard=(case["N"], world))
        total.backward()
        flat = torch.cat([p.grad.reshape(-1) for p in params] + [total.detach().reshape(1)])
        parallel.exchange(flat)   # the single collective of the path
        # loader statistics helper
        xs = case["x"][r0:r1]
        mean, std = parallel.global_mean_std(xs.double().sum(), (xs.double() ** 2).sum(), case["x"].numel())
        if rank == 0:
            out.put((flat.numpy(), float(mean), float(std)))
    finally:
        dist.destroy_process_group()


def test_sharded_closure_equals_unsharded():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    flat, mean, std = q.get(timeout=600)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    case = closure_case(N=8, bpb=2)
    ref = oracle_closure(case)
    ref_flat = torch.cat([g.reshape(-1) for g in ref["grads"].values()] + [torch.tensor([ref["total"]])]).numpy()
    assert abs(flat[-1] - ref["total"]) < 1e-5 * abs(ref["total"])
    err = np.linalg.norm(flat[:-1] - ref_flat[:-1]) / np.linalg.norm(ref_flat[:-1])
    assert err < 1e-4, err
    assert abs(mean - case["x"].double().mean().item()) < 1e-9
    assert abs(std - case["x"].double().std().item()) < 1e-9


def test_shard_groups_cover_everything_once():
    for n, w in ((62, 8), (7, 2), (3, 4), (256, 8)):
        spans = [parallel.shard_groups(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1
    assert parallel.shard_rows(24, 4, 1, 2) == (12, 24)
    with pytest.raises(ValueError):
        parallel.shard_rows(10, 4, 0, 2)


def test_shard_plan_constants():
    plan = parallel.ShardPlan(n_local=512, n_global=1024, world=2, bpb=4, channels=8, K=10, Ltot=64)
    assert plan.numel_global == 1024 * 8 * 16384
    assert plan.khm_scale(0.01) == 0.01 / (1024 * 10 * 64)
    assert plan.aug_scale(0.01) == 0.01 / (4 * 256 * 4)
    assert plan.sim_scale(0.01) == 0.005 and plan.rica_scale(0.01) == 0.005
