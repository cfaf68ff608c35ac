"""LBFGSNew for real on the GPU path (BASELINE cfg5).

The fused closure is driven (i) by the UNMODIFIED reference optimiser (/root/reference/src/lbfgsnew.py, from
baseline/_ref) over the leaf Parameters and (ii) by lshm_b200.lbfgsnew.LBFGSNew over the flat buffer (vector ops as
views, cached f_old probes, CUDA-graph replays), with the reference's settings
LBFGSNew(history_size=7, max_iter=4, line_search_fn=True, batch_mode=True) (src/kharmonic_lofar.py:93), and both
trajectories are compared with the reference modules + reference optimiser running the restated script loop on
the CPU (oracle/reference_loop.py) from the same parameters on the same patches."""
import numpy as np
import pytest
import torch

from common import SCALES, closure_case
from lshm_b200._lib import lib
from oracle import reference_loop as RL

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(RL.reference_dir() is None, reason="reference sources not available (baseline/_ref)")]
LBFGS_KW = dict(history_size=7, max_iter=4, line_search_fn=True, batch_mode=True)


def build_step(case, cuda, **kw):
    from lshm_b200.kharmonic_lofar import DeepKHarmonicStep
    from lshm_b200.lofar_models import AutoEncoder1DCNN, AutoEncoderCNN2, Kmeans
    hs = torch.tensor(SCALES).to(cuda)
    net = AutoEncoderCNN2(case["L"], case["C"], hs, True)
    netT = AutoEncoder1DCNN(case["Lt"], case["C"], hs, True)
    netF = AutoEncoder1DCNN(case["Lt"], case["C"], hs, True)
    mod = Kmeans(case["L"] + 2 * case["Lt"], case["K"], 4)
    net.load_state_dict(case["pn"]); netT.load_state_dict(case["pT"]); netF.load_state_dict(case["pF"])
    mod.load_state_dict({"M": case["M"]})
    step = DeepKHarmonicStep(net.to(cuda), netT.to(cuda), netF.to(cuda), mod.to(cuda), **kw)
    step.set_batch(case["x"].to(cuda), case["uv"].to(cuda), case["bpb"])
    return step


def cpu_trajectory(case, n_admm):
    R = RL.ReferenceLoop(L=case["L"], Lt=case["Lt"], C=case["C"], K=case["K"], Khp=4, optimizer="lbfgs",
                         state=(case["pn"], case["pT"], case["pF"], case["M"]))
    R.set_batch(case["x"], case["uv"], case["bpb"])
    losses, calls = [], []
    for _ in range(n_admm):
        c0 = R.closures
        losses.append(float(R.admm_iteration()))
        calls.append(R.closures - c0)
    with torch.no_grad():
        final = float(R.closure())
    return losses, calls, final, R


def gpu_trajectory(step, opt, n_admm):
    losses, calls = [], []
    count = [0]

    def closure():
        count[0] += 1
        return step.closure()
    for _ in range(n_admm):
        c0 = count[0]
        losses.append(float(opt.step(closure)))
        calls.append(count[0] - c0)
        step.update_multipliers()
    with torch.no_grad():
        final = float(step.closure())
    return losses, calls, final


@pytest.mark.parametrize("N,n_admm", [(16, 3), (248, 2)], ids=["N16", "cfg5_62_baselines"])
def test_reference_lbfgsnew_and_flat_lbfgsnew_follow_the_cpu_reference(cuda, N, n_admm):
    from lshm_b200.lbfgsnew import LBFGSNew
    Ref = RL.load_reference_module("lbfgsnew").LBFGSNew
    case = closure_case(N=N, bpb=4, seed=3)
    l_cpu, c_cpu, f_cpu, R = cpu_trajectory(case, n_admm)
    assert min(c_cpu) >= 12          # >= 4 gradient + >= 8 line-search closures per step
    # (i) the unmodified reference optimiser on the drop-in Parameters
    s1 = build_step(case, cuda)
    l_ref, c_ref, f_ref = gpu_trajectory(s1, Ref(s1.flat.params, **LBFGS_KW), n_admm)
    # (ii) the flat optimiser, launch sequences replayed from CUDA graphs
    s2 = build_step(case, cuda)
    s2.enable_graphs()
    l0 = lib().launches
    l_new, c_new, f_new = gpu_trajectory(s2, LBFGSNew(s2.flat, **LBFGS_KW), n_admm)
    launches_new = lib().launches - l0
    print(f"\nN={N}: cpu {l_cpu} calls {c_cpu} final {f_cpu}\n   ref-opt/gpu {l_ref} calls {c_ref} final {f_ref}\n"
          f"   flat-opt/gpu {l_new} calls {c_new} final {f_new} launches {launches_new}")
    for got in (l_ref, l_new):
        assert np.allclose(got, l_cpu, rtol=5e-3), (got, l_cpu)
    assert abs(f_ref - f_cpu) <= 2e-2 * abs(f_cpu) and abs(f_new - f_cpu) <= 2e-2 * abs(f_cpu)
    # (the total loss GROWS over the ADMM iterations of one minibatch - the multipliers y_i grow - exactly as on the CPU)
    # same line-search decisions as the CPU reference in the first step (later steps may flip on last-bit ties)
    assert c_ref[0] == c_cpu[0] and c_new[0] == c_cpu[0], (c_ref, c_new, c_cpu)
    # parameters after the run: both GPU runs stay close to the CPU reference
    sd = R.net.state_dict()
    for s in (s1, s2):
        num = sum(float((p.detach().cpu() - sd[nm.split(".", 1)[1]]).pow(2).sum())
                  for nm, p in zip(s.flat.names, s.flat.params) if nm.startswith("0."))
        den = sum(float(v.pow(2).sum()) for v in sd.values())
        assert (num / den) ** 0.5 < 2e-2


def test_flat_lbfgsnew_skips_repeated_work(cuda):
    """f2: vector ops on the flat buffer; the f_old probe of every line search (src/lbfgsnew.py:140) is answered from
    the loss scalars of the gradient closure just evaluated at the same parameters: no kernel is launched for it."""
    from lshm_b200.lbfgsnew import LBFGSNew
    case = closure_case(N=8, bpb=4, seed=4)
    step = build_step(case, cuda)
    opt = LBFGSNew(step.flat, **LBFGS_KW)
    free, total = [0], [0]

    def closure():
        n0 = lib().launches
        out = step.closure()
        total[0] += 1
        free[0] += int(lib().launches == n0)
        return out
    v0 = step.flat.version
    opt.step(closure)
    assert total[0] >= 12 and free[0] >= 4              # one cached probe per inner iteration
    assert step.flat.version > v0
    # the optimiser's flat gradient IS the gradient buffer (no gather copy), and its step IS one axpy
    assert opt._gather_flat_grad().data_ptr() == step.flat.grad.data_ptr()
    before = step.flat.flat.clone()
    d = torch.randn_like(before)
    opt._add_grad(0.5, d)
    assert torch.allclose(step.flat.flat, before + 0.5 * d)
    saved = opt._copy_params_out()
    opt._add_grad(1.0, d)
    opt._copy_params_in(saved)
    assert torch.equal(step.flat.flat, saved)
    # NaN costs are backtracked, not raised (src/lbfgsnew.py:153)
    step.x[0, 0, 0, 0] = float("nan")
    step.invalidate()
    opt2 = LBFGSNew(step.flat, **LBFGS_KW)
    out = opt2.step(step.closure)
    assert np.isnan(float(out))
