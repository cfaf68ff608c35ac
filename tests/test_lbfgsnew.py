"""lshm_b200.lbfgsnew.LBFGSNew against the reference optimiser (/root/reference/src/lbfgsnew.py, unmodified, from
baseline/_ref or /root/reference/src) on CPU problems: same closure protocol, same trajectories in both modes
(stochastic: backtracking line search + running gradient statistics; full batch: cubic line search) and without
a line search.  The GPU path (flat vector ops on the fused closure) is in tests/test_gpu_lbfgs.py."""
import copy

import pytest
import torch

from oracle import reference_loop as RL

pytestmark = pytest.mark.skipif(RL.reference_dir() is None, reason="reference sources not available (baseline/_ref)")


def make_problem(seed=0, n=64):
    torch.manual_seed(seed)
    model = torch.nn.Sequential(torch.nn.Linear(6, 16), torch.nn.Tanh(), torch.nn.Linear(16, 3))
    X = torch.randn(4, n, 6)
    W = torch.randn(6, 3)
    Y = torch.tanh(X @ W) + 0.05 * torch.randn(4, n, 3)
    return model, X, Y


def run(opt_cls, model, X, Y, steps, **kw):
    opt = opt_cls(model.parameters(), **kw)
    losses, evals = [], []
    for it in range(steps):
        xb, yb = X[it % X.shape[0]], Y[it % X.shape[0]]
        calls = [0, 0]

        def closure():
            if torch.is_grad_enabled():
                opt.zero_grad()
            loss = torch.nn.functional.mse_loss(model(xb), yb)
            if loss.requires_grad:
                loss.backward()
                calls[0] += 1
            else:
                calls[1] += 1
            return loss
        losses.append(float(opt.step(closure)))
        evals.append(tuple(calls))
    flat = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
    return losses, evals, flat, opt


@pytest.mark.parametrize("kw", [
    dict(history_size=7, max_iter=4, line_search_fn=True, batch_mode=True),     # src/kharmonic_lofar.py:93
    dict(history_size=7, max_iter=6, line_search_fn=True, batch_mode=False),    # full batch, cubic line search
    dict(history_size=3, max_iter=5, line_search_fn=False, batch_mode=False, lr=0.1),
    dict(history_size=5, max_iter=3, line_search_fn=False, batch_mode=True, lr=0.05),
], ids=["batch_backtrack", "fullbatch_cubic", "fixed_step", "batch_fixed_step"])
def test_trajectory_matches_reference_optimiser(kw):
    from lshm_b200.lbfgsnew import LBFGSNew
    Ref = RL.load_reference_module("lbfgsnew").LBFGSNew
    model, X, Y = make_problem()
    if not kw.get("batch_mode"):
        X, Y = X[:1], Y[:1]
    m_ref, m_new = copy.deepcopy(model), copy.deepcopy(model)
    l_ref, e_ref, p_ref, o_ref = run(Ref, m_ref, X, Y, 8, **kw)
    l_new, e_new, p_new, o_new = run(LBFGSNew, m_new, X, Y, 8, **kw)
    assert e_new == e_ref, (e_new, e_ref)          # same number of gradient / line-search closures every step
    assert torch.allclose(torch.tensor(l_new), torch.tensor(l_ref), rtol=2e-4, atol=1e-7), (l_new, l_ref)
    assert (p_new - p_ref).norm() <= 2e-3 * p_ref.norm()
    s_ref, s_new = o_ref.state[o_ref._params[0]], o_new.state[o_new._params[0]]
    assert s_new["n_iter"] == s_ref["n_iter"] and s_new["func_evals"] == s_ref["func_evals"]
    assert len(s_new["old_dirs"]) == len(s_ref["old_dirs"])
    assert l_new[-1] < l_new[0]


def test_constructor_contract():
    from lshm_b200.lbfgsnew import LBFGSNew
    model, _, _ = make_problem()
    opt = LBFGSNew(model.parameters(), max_iter=8)
    g = opt.param_groups[0]
    assert g["max_eval"] == 10 and g["history_size"] == 7 and g["lr"] == 1 and g["tolerance_grad"] == 1e-5
    assert opt._numel() == sum(p.numel() for p in model.parameters())
    with pytest.raises(ValueError):
        LBFGSNew([{"params": [next(model.parameters())]}, {"params": list(model.parameters())[1:]}])


def test_nan_loss_is_backtracked_not_raised():
    """src/lbfgsnew.py:153: a NaN cost in the line search halves the step instead of failing."""
    from lshm_b200.lbfgsnew import LBFGSNew
    p = torch.nn.Parameter(torch.tensor([3.0]))
    opt = LBFGSNew([p], max_iter=2, line_search_fn=True, batch_mode=True)

    def closure():
        if torch.is_grad_enabled():
            opt.zero_grad()
        loss = (torch.sqrt(p) - 1.0).pow(2).sum()     # NaN for p < 0: the first full step overshoots
        if loss.requires_grad:
            loss.backward()
        return loss
    l0 = float(opt.step(closure))
    assert torch.isfinite(p).all() and float(closure()) <= l0
