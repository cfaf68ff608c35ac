"""Multi-GPU parity check (run under torchrun on >= 2 GPUs; tests/test_gpu_dp.py launches it from pytest):
each rank takes its shard of whole baseline groups, runs the fused closure with the NCCL exchange,
and rank 0 compares the all-reduced loss and gradients with the unsharded CPU oracle.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29533 tests/dp_check.py
"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from common import SCALES, closure_case, oracle_closure, rel_err  # noqa: E402
from lshm_b200 import parallel  # noqa: E402
from lshm_b200.kharmonic_lofar import DeepKHarmonicStep  # noqa: E402
from lshm_b200.lofar_models import AutoEncoder1DCNN, AutoEncoderCNN2, Kmeans  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    case = closure_case(N=int(os.environ.get("DP_CHECK_N", "32")), bpb=2)
    hs = torch.tensor(SCALES).to(dev)
    net = AutoEncoderCNN2(case["L"], case["C"], hs, True)
    netT = AutoEncoder1DCNN(case["Lt"], case["C"], hs, True)
    netF = AutoEncoder1DCNN(case["Lt"], case["C"], hs, True)
    mod = Kmeans(case["L"] + 2 * case["Lt"], case["K"], 4)
    if rank == 0:   # only rank 0 gets the real parameters: the constructor must broadcast them
        net.load_state_dict(case["pn"]); netT.load_state_dict(case["pT"]); netF.load_state_dict(case["pF"])
        mod.load_state_dict({"M": case["M"]})
    step = DeepKHarmonicStep(net.to(dev), netT.to(dev), netF.to(dev), mod.to(dev), distributed=True, centre_sums=True)
    r0, r1 = parallel.shard_rows(case["N"], case["bpb"], rank, world)
    npix = case["C"] * 16384
    step.set_batch(case["x"][r0:r1].to(dev), case["uv"][r0:r1].to(dev), case["bpb"], global_patches=case["N"])
    for dst, src in zip((step.y1, step.y2, step.y3), case["ys"]):
        dst.copy_(src[r0 * npix:r1 * npix].to(dev))
    loss = float(step.closure())
    with torch.no_grad():
        loss_fwd = float(step.closure())
    losses = [None] * world
    dist.all_gather_object(losses, loss)
    if rank == 0:
        ref = oracle_closure(case)
        worst = max(rel_err(p.grad, ref["grads"][nm]) for nm, p in zip(step.flat.names, step.flat.params))
        ok = (abs(loss - ref["total"]) < 5e-5 * abs(ref["total"]) and abs(loss_fwd - loss) < 1e-6 * abs(loss)
              and worst < 2e-4 and len(set(losses)) == 1)
        print(f"dp_check world={world}: loss {loss:.7f} oracle {ref['total']:.7f} worst grad rel err {worst:.2e} "
              f"identical loss on all ranks: {len(set(losses)) == 1} -> {'PASS' if ok else 'FAIL'}")
        if not ok:
            sys.exit(1)
        # a12: the K x L numerator / K denominator of Kmeans.offline_update travelled in the same all-reduce;
        # the reduced sums must equal the oracle's on the UNSHARDED latents
        from oracle import lofar_oracle as O
        Mn, num_ref, den_ref = O.offline_update(ref["Mu"], case["M"], 4)
        num, den = step.centre_sums_view()
        e_num, e_den = rel_err(num, num_ref), rel_err(den, den_ref)
        ok_c = e_num < 1e-3 and e_den < 1e-3
        print(f"dp_check world={world}: centre sums in the exchange buffer: num rel err {e_num:.2e} den rel err {e_den:.2e} "
              f"-> {'PASS' if ok_c else 'FAIL'}")
        if not ok_c:
            sys.exit(1)
    # graphed data-parallel steps (CUDA graphs around the eager all-reduce, forward reuse + deferred multiplier
    # update under the tracking FlatAdam) against the plain eager loop
    from lshm_b200.kharmonic_lofar import FlatAdam, GraphedStep
    state = [t.clone() for t in (step.flat.flat, step.y1, step.y2, step.y3)]
    step.reuse = False
    opt = FlatAdam(step.flat, lr=1e-3)
    eager = []
    for _ in range(3):
        eager.append(float(opt.step(step.closure)))
        step.update_multipliers()
    p_eager = step.flat.flat.clone()
    y_eager = step.y1.clone()
    for dst, src in zip((step.flat.flat, step.y1, step.y2, step.y3), state):
        dst.copy_(src)
    step.invalidate()
    step.reuse = True
    opt2 = FlatAdam(step.flat, lr=1e-3)
    gs = GraphedStep(step, opt2)
    graphed = [float(gs.replay()) for _ in range(3)]
    err = rel_err(step.flat.flat, p_eager)
    err_y = rel_err(step.y1, y_eager)
    ok2 = all(abs(a - b) <= 2e-4 * abs(b) for a, b in zip(graphed, eager)) and err < 2e-4 and err_y < 2e-3
    flags = [None] * world
    dist.all_gather_object(flags, ok2)
    if rank == 0:
        print(f"dp_check world={world}: graphed losses {graphed} eager {eager} param rel err {err:.2e} y1 rel err {err_y:.2e} "
              f"-> {'PASS' if all(flags) else 'FAIL'}")
        if not all(flags):
            sys.exit(1)
    # loader z-score with GLOBAL statistics (src/lofar_tools.py:190-193 over the whole minibatch): each rank patchifies
    # its own baselines, (sum, sum of squares) are all-reduced, every rank normalises with the global mean / std
    from lshm_b200 import lofar_tools as T
    from lshm_b200 import synthetic as S
    nb = 4 * world
    meas = S.make_measurement(nb, 192, 192, seed=9)["measurement"]["saps"]["0"]
    vis = torch.from_numpy(meas["visibilities"]).to(dev)
    sc = torch.from_numpy(meas["visibility_scale_factors"]).to(dev)
    sel_all = torch.arange(nb, dtype=torch.int32, device=dev)
    _, _, y_all = T.patchify_device(vis, sc, sel_all, 128, 8, 1e3, True)
    g0, g1 = parallel.shard_groups(nb, rank, world)
    _, _, y_loc = T.patchify_device(vis, sc, sel_all[g0:g1].contiguous(), 128, 8, 1e3, True, group=dist.group.WORLD,
                                    n_global=y_all.numel())
    # rows are patch-major (n = patch * nb + k): compare baseline by baseline
    ya = y_all.view(4, nb, -1)[:, g0:g1]
    yl = y_loc.view(4, g1 - g0, -1)
    ok3 = rel_err(yl, ya) < 1e-5
    dist.all_gather_object(flags, ok3)
    if rank == 0:
        print(f"dp_check world={world}: sharded loader with all-reduced z-score statistics -> {'PASS' if all(flags) else 'FAIL'}")
        if not all(flags):
            sys.exit(1)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
