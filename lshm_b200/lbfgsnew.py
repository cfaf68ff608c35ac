"""LBFGSNew with GPU-resident vector operations on the flat parameter / gradient storage.

Same optimiser as /root/reference/src/lbfgsnew.py (class LBFGSNew, :9-759): stochastic / full-batch L-BFGS with a
backtracking line search in batch mode (:115-190) or Fletcher's cubic-interpolation line search (:192-495), the
inter-batch running mean / variance of the gradient that bounds the step in batch mode (:592-607), the same
constructor arguments, defaults, `state` keys and closure protocol (grad-enabled closures must leave `.grad`,
line-search closures run under `torch.set_grad_enabled(False)`, :686-693).  The unmodified reference optimiser
also works on the drop-in modules; this one exists for SURVEY.md 8(f2): when the parameters live in a
`kharmonic_lofar.FlatParams` buffer

  * `_gather_flat_grad` (:84-94, a `torch.cat` of ~110 tensors per call) is a VIEW of the flat gradient buffer,
  * `_add_grad` (:96-103, ~110 `p.data.add_` launches) is ONE axpy on the flat parameter buffer,
  * `_copy_params_out` / `_copy_params_in` (:106-112, ~110 clones / copies) are one clone / one copy,
  * the two-loop recursion (:637-651) takes its step sizes as device scalars (no `alpha=<tensor>` host syncs),

and every parameter change bumps `FlatParams.version`, so the fused closure knows when the activations and loss
it holds are still valid (the `f_old` probe of every line search, :140, costs nothing).  With a plain parameter
list it falls back to gather / scatter like the reference (that path also runs on CPU tensors: it is how the CPU
test-suite checks this implementation against the reference optimiser, tests/test_lbfgsnew.py).

The padding floats between the 256-byte aligned tensors of a FlatParams buffer have zero gradient, so they stay
zero in every vector of the recursion and do not change any dot product, norm or sum.
"""
from __future__ import annotations

import math
from typing import List, Optional

import numpy as np
import torch
from torch.optim.optimizer import Optimizer

be_verbose = False
_f32 = np.float32


class _ListBackend:
    """Parameters as a list of tensors: gather / scatter (the reference's own scheme, src/lbfgsnew.py:84-112)."""

    def __init__(self, params: List[torch.nn.Parameter]):
        self.params = params
        self.numel = sum(p.numel() for p in params)

    def grad(self) -> torch.Tensor:
        views = []
        for p in self.params:
            if p.grad is None:
                views.append(p.data.new_zeros(p.data.numel()))
            elif p.grad.data.is_sparse:
                views.append(p.grad.data.to_dense().contiguous().view(-1))
            else:
                views.append(p.grad.data.contiguous().view(-1))
        return torch.cat(views, 0)

    def axpy(self, alpha: float, vec: torch.Tensor):
        off = 0
        for p in self.params:
            n = p.numel()
            p.data.add_(vec[off:off + n].view_as(p.data), alpha=alpha)
            off += n
        assert off == self.numel

    def save(self):
        return [p.clone(memory_format=torch.contiguous_format) for p in self.params]

    def restore(self, saved):
        for p, q in zip(self.params, saved):
            p.copy_(q)


class _FlatBackend:
    """Parameters are views of one FlatParams buffer: every vector operation is a single launch on it."""

    def __init__(self, flat):
        self.flat = flat
        self.numel = flat.numel
        flat.tracked = True

    def grad(self) -> torch.Tensor:
        self.flat.attach_grads()
        return self.flat.grad[:self.numel]          # a view: callers copy what must outlive the next closure

    def axpy(self, alpha: float, vec: torch.Tensor):
        self.flat.flat.add_(vec, alpha=alpha)
        self.flat.bump()

    def save(self):
        return self.flat.flat.clone()

    def restore(self, saved):
        self.flat.flat.copy_(saved)
        self.flat.bump()


class LBFGSNew(Optimizer):
    """L-BFGS, constructor and behaviour of src/lbfgsnew.py:62-79.

    params: an iterable of Parameters, or a `kharmonic_lofar.FlatParams` (then `flat.params` are optimised through
    the flat buffer).  Only one parameter group (like the reference, :72-74).
    """

    def __init__(self, params, lr=1, max_iter=10, max_eval=None, tolerance_grad=1e-5, tolerance_change=1e-9,
                 history_size=7, line_search_fn=False, batch_mode=False, cost_use_gradient=False):
        flat = params if hasattr(params, "flat") and hasattr(params, "grad_views") else None
        plist = list(flat.params) if flat is not None else list(params)
        if max_eval is None:
            max_eval = max_iter * 5 // 4
        defaults = dict(lr=lr, max_iter=max_iter, max_eval=max_eval, tolerance_grad=tolerance_grad,
                        tolerance_change=tolerance_change, history_size=history_size, line_search_fn=line_search_fn,
                        batch_mode=batch_mode, cost_use_gradient=cost_use_gradient)
        super().__init__(plist, defaults)
        if len(self.param_groups) != 1:
            raise ValueError("LBFGS doesn't support per-parameter options (parameter groups)")
        self._params = self.param_groups[0]["params"]
        self._vec = _FlatBackend(flat) if flat is not None else _ListBackend(self._params)

    # ---- the four vector primitives of the reference (:84-112), kept under their names
    def _numel(self):
        return self._vec.numel

    def _gather_flat_grad(self):
        return self._vec.grad()

    def _add_grad(self, step_size, update):
        self._vec.axpy(float(step_size), update)

    def _copy_params_out(self):
        return self._vec.save()

    def _copy_params_in(self, new_params):
        self._vec.restore(new_params)

    def _glob(self):
        return self.state[self._params[0]]

    # ------------------------------------------------------------------ line searches
    def _linesearch_backtrack(self, closure, pk, gk, alphabar):
        """Backtracking (Armijo) search with a negative-step probe, src/lbfgsnew.py:115-190.  The comparisons the
        reference makes between Python floats and 0-dim fp32 tensors are made in fp32 here as well."""
        c1, citer = 1e-4, 35
        alphak = float(alphabar)
        state = self._glob()
        xk = self._copy_params_out()
        f_old = float(closure())
        self._add_grad(alphak, pk)
        f_new = float(closure())
        prodterm = _f32(c1) * _f32(float(gk.dot(pk)))

        def too_high(f, a):       # f > f_old + a * prodterm in fp32, or f is NaN
            return math.isnan(f) or _f32(f) > _f32(f_old) + _f32(a) * prodterm

        ci = 0
        if be_verbose:
            print("LN %d alpha=%f fnew=%f fold=%f prod=%f" % (ci, alphak, f_new, f_old, prodterm))
        while ci < citer and too_high(f_new, alphak):
            alphak = 0.5 * alphak
            self._copy_params_in(xk)
            self._add_grad(alphak, pk)
            f_new = float(closure())
            ci += 1
        if _f32(f_old - f_new) < abs(prodterm):
            # the cost did not decrease enough: also try the opposite direction (:163-177)
            alphak1 = -float(alphabar)
            self._copy_params_in(xk)
            self._add_grad(alphak1, pk)
            f_new1 = float(closure())
            while ci < citer and too_high(f_new1, alphak1):
                alphak1 = 0.5 * alphak1
                self._copy_params_in(xk)
                self._add_grad(alphak1, pk)
                f_new1 = float(closure())
                ci += 1
            if f_new1 < f_new:
                alphak = alphak1
        self._copy_params_in(xk)
        state["func_evals"] += ci
        return alphak

    def _phi(self, closure, delta, pk):
        """Move along pk by `delta` from wherever the parameters are and evaluate the cost."""
        self._add_grad(delta, pk)
        return float(closure())

    def _dphi(self, closure, pk, step, pre=0.0):
        """Central difference of the cost along pk at (current point + pre); the parameters end at that point - step."""
        up = self._phi(closure, pre + step, pk)
        dn = self._phi(closure, -2.0 * step, pk)
        return (up - dn) / (2.0 * step)

    def _cubic_interpolate(self, closure, xk, pk, a, b, step):
        """Minimiser of the cubic through (a, b) with finite-difference slopes, src/lbfgsnew.py:319-408."""
        state = self._glob()
        self._copy_params_in(xk)
        f0 = self._phi(closure, a, pk)
        f0d = self._dphi(closure, pk, step)                 # parameters now at a - step
        f1 = self._phi(closure, -a + step + b, pk)
        f1d = self._dphi(closure, pk, step)                 # parameters now at b - step
        evals = 6
        aa = 3.0 * (f0 - f1) / (b - a) + f1d - f0d
        disc = aa * aa - f0d * f1d
        if disc > 0.0:
            cc = math.sqrt(disc)
            if (f1d - f0d + 2.0 * cc) == 0.0:
                return (a + b) * 0.5                        # (the reference returns before counting the evals, :379)
            z0 = b - (f1d + cc - aa) * (b - a) / (f1d - f0d + 2.0 * cc)
            hi, lo = max(a, b), min(a, b)
            if z0 > hi or z0 < lo:
                fz0 = f0 + f1
            else:
                fz0 = self._phi(closure, -b + step + a + z0 * (b - a), pk)   # the offset of :387, as written
                evals += 1
            state["func_evals"] += evals
            if f0 < f1 and f0 < fz0:
                return a
            if f1 < fz0:
                return b
            return z0
        state["func_evals"] += evals
        return a if f0 < f1 else b

    def _linesearch_zoom(self, closure, xk, pk, a, b, phi_0, gphi_0, sigma, rho, t1, t2, t3, step):
        """Sectioning phase of Fletcher's line search, src/lbfgsnew.py:412-495 (at most 4 rounds)."""
        state = self._glob()
        evals = 0
        aj, bj = a, b
        alphaj = None
        for _ in range(4):
            alphaj = self._cubic_interpolate(closure, xk, pk, aj + t2 * (bj - aj), bj - t3 * (bj - aj), step)
            self._copy_params_in(xk)
            phi_j = self._phi(closure, alphaj, pk)
            phi_aj = self._phi(closure, -alphaj + aj, pk)
            evals += 2
            if (phi_j > phi_0 + rho * alphaj * gphi_0) or phi_j >= phi_aj:
                bj = alphaj
                continue
            gphi_j = self._dphi(closure, pk, step, pre=-aj + alphaj)     # back to alphaj (+ step) in one move, :461
            evals += 2
            if (aj - alphaj) * gphi_j <= step or abs(gphi_j) <= -sigma * gphi_0:
                break
            if gphi_j * (bj - aj) >= 0.0:
                bj = aj
            aj = alphaj
        state["func_evals"] += evals
        return alphaj

    def _linesearch_cubic(self, closure, pk, step):
        """Bracketing phase (strong Wolfe, Fletcher), src/lbfgsnew.py:192-315."""
        lr = self.param_groups[0]["lr"]
        alpha1, sigma, rho, t1, t2, t3 = 10 * lr, 0.1, 0.01, 9, 0.1, 0.5
        alphak = lr
        state = self._glob()
        xk = self._copy_params_out()
        phi_0 = float(closure())
        tol = min(phi_0 * 0.01, 1e-6)
        gphi_0 = self._dphi(closure, pk, step)
        if abs(gphi_0) < 1e-12:
            return 1.0                                      # (parameters are left at -step, as in the reference :240-241)
        mu = (tol - phi_0) / (rho * gphi_0)
        if math.isnan(mu):
            return 1.0
        evals = 3
        ci, alphai, alphai1, phi_alphai1 = 1, alpha1, 0.0, phi_0
        while ci < 4:
            self._copy_params_in(xk)
            phi_alphai = self._phi(closure, alphai, pk)
            if phi_alphai < tol:
                alphak = alphai
                break
            if (phi_alphai > phi_0 + alphai * gphi_0) or (ci > 1 and phi_alphai >= phi_alphai1):
                alphak = self._linesearch_zoom(closure, xk, pk, alphai1, alphai, phi_0, gphi_0, sigma, rho, t1, t2, t3, step)
                break
            gphi_i = self._dphi(closure, pk, step)
            if abs(gphi_i) <= -sigma * gphi_0:
                alphak = alphai
                break
            if gphi_i >= 0.0:
                alphak = self._linesearch_zoom(closure, xk, pk, alphai, alphai1, phi_0, gphi_0, sigma, rho, t1, t2, t3, step)
                break
            if mu <= 2.0 * alphai - alphai1:
                alphai1, alphai = alphai, mu
            else:
                lo = 2.0 * alphai - alphai1
                hi = min(mu, alphai + t1 * (alphai - alphai1))
                # (the reference overwrites alphai here and leaves alphai1 unchanged, :299-303)
                alphai = self._cubic_interpolate(closure, xk, pk, lo, hi, step)
            phi_alphai1 = phi_alphai
            evals += 3
            ci += 1
        self._copy_params_in(xk)
        state["func_evals"] += evals
        return alphak

    # ------------------------------------------------------------------ direction
    def _two_loop(self, flat_grad, old_dirs, old_stps, H_diag):
        """L-BFGS two-loop recursion, src/lbfgsnew.py:628-651: returns H * (-g).

        Flat backend: the step sizes stay on the device (`addcmul` with 0-dim tensors; on CUDA the multiply-add is
        contracted to the same fused multiply-add as `add_(x, alpha=a)`), so the recursion issues 4m+1 launches and no
        host synchronisation.  List backend: the reference's own `add_(..., alpha=tensor)` form, bit for bit (the
        finite-difference cubic line search amplifies last-bit differences of the direction)."""
        num_old = len(old_dirs)
        ro = [1.0 / old_dirs[i].dot(old_stps[i]) for i in range(num_old)]
        al: List[Optional[torch.Tensor]] = [None] * num_old
        on_device = isinstance(self._vec, _FlatBackend)
        q = flat_grad.neg()
        for i in range(num_old - 1, -1, -1):
            al[i] = old_stps[i].dot(q) * ro[i]
            if on_device:
                q.addcmul_(old_dirs[i], al[i], value=-1)
            else:
                q.add_(old_dirs[i], alpha=-al[i])
        r = torch.mul(q, H_diag)
        for i in range(num_old):
            be_i = old_dirs[i].dot(r) * ro[i]
            if on_device:
                r.addcmul_(old_stps[i], al[i] - be_i)
            else:
                r.add_(old_stps[i], alpha=al[i] - be_i)
        return r, ro, al

    # ------------------------------------------------------------------ step
    def step(self, closure):
        """One optimisation step, src/lbfgsnew.py:498-759."""
        if isinstance(self._vec, _FlatBackend):
            with self._vec.flat.owning():       # every closure of this step sees parameters only we change
                return self._step(closure)
        return self._step(closure)

    def _step(self, closure):
        assert len(self.param_groups) == 1
        group = self.param_groups[0]
        lr, max_iter, max_eval = group["lr"], group["max_iter"], group["max_eval"]
        tolerance_grad, tolerance_change = group["tolerance_grad"], group["tolerance_change"]
        line_search_fn, history_size = group["line_search_fn"], group["history_size"]
        batch_mode, cost_use_gradient = group["batch_mode"], group["cost_use_gradient"]
        state = self._glob()
        state.setdefault("func_evals", 0)
        state.setdefault("n_iter", 0)

        orig_loss = closure()
        loss = float(orig_loss)
        current_evals = 1
        state["func_evals"] += 1
        def gather():
            g = self._gather_flat_grad()
            # a view of the live gradient buffer is only safe while line-search closures leave gradients alone
            return g.clone() if (cost_use_gradient and isinstance(self._vec, _FlatBackend)) else g

        flat_grad = gather()
        abs_grad_sum = float(flat_grad.abs().sum())
        if abs_grad_sum <= tolerance_grad:
            return orig_loss

        d, t = state.get("d"), state.get("t")
        old_dirs, old_stps = state.get("old_dirs"), state.get("old_stps")
        H_diag = state.get("H_diag")
        prev_flat_grad, prev_loss = state.get("prev_flat_grad"), state.get("prev_loss")
        running_avg = running_avg_sq = None
        n_iter = 0
        alphabar, lm0 = lr, 1e-6
        grad_nrm = flat_grad.norm().item()          # (not refreshed inside the loop, as in the reference :567)
        while n_iter < max_iter and not math.isnan(grad_nrm):
            n_iter += 1
            state["n_iter"] += 1
            # ---------------------------------------------------------------- direction
            if state["n_iter"] == 1:
                d = flat_grad.neg()
                old_dirs, old_stps, H_diag = [], [], 1
                if batch_mode:
                    running_avg = torch.zeros_like(flat_grad)
                    running_avg_sq = torch.zeros_like(flat_grad)
            else:
                if batch_mode:
                    running_avg, running_avg_sq = state.get("running_avg"), state.get("running_avg_sq")
                    if running_avg is None:
                        running_avg = torch.zeros_like(flat_grad)
                        running_avg_sq = torch.zeros_like(flat_grad)
                y = flat_grad.sub(prev_flat_grad)
                s = d.mul(t)
                if batch_mode:
                    y.add_(s, alpha=lm0)            # trust-region term (:586-587)
                ys_t, yy_t, sn_t = y.dot(s), y.dot(y), s.norm()
                ys, yy, sn = torch.stack((ys_t, yy_t, sn_t)).tolist()      # one host read for the three scalars
                batch_changed = batch_mode and (n_iter == 1 and state["n_iter"] > 1)
                if batch_changed:
                    # inter-batch running mean / second moment of the gradient -> maximum step (:592-607)
                    g_old = flat_grad - running_avg
                    running_avg.add_(g_old, alpha=1.0 / state["n_iter"])
                    g_new = flat_grad - running_avg
                    running_avg_sq.addcmul_(g_new, g_old, value=1)
                    alphabar = float(1 / (1 + running_avg_sq.sum() / ((state["n_iter"] - 1) * grad_nrm)))
                if _f32(ys) > _f32(1e-10 * sn * sn) and not batch_changed:
                    if len(old_dirs) == history_size:
                        old_dirs.pop(0)
                        old_stps.pop(0)
                    old_dirs.append(y)
                    old_stps.append(s)
                    H_diag = ys_t / yy_t
                    if yy == 0.0 or math.isnan(ys / yy):
                        print("Warning H_diag nan")
                d, ro, al = self._two_loop(flat_grad, old_dirs, old_stps, H_diag)
                state["ro"], state["al"] = ro, al
            if prev_flat_grad is None:
                prev_flat_grad = flat_grad.clone(memory_format=torch.contiguous_format)
            else:
                prev_flat_grad.copy_(flat_grad)
            prev_loss = loss
            # ---------------------------------------------------------------- step length
            t = min(1.0, 1.0 / abs_grad_sum) * lr if state["n_iter"] == 1 else lr
            gtd = float(flat_grad.dot(d))
            if math.isnan(gtd):
                print("Warning grad norm infinite")
                print("iter %d" % state["n_iter"])
                print("||grad||=%f" % grad_nrm)
                print("||d||=%f" % d.norm().item())
            ls_func_evals = 0
            if line_search_fn:
                # the line search evaluates the cost only: no gradients (:686-693)
                if not cost_use_gradient:
                    torch.set_grad_enabled(False)
                try:
                    if not batch_mode:
                        t = self._linesearch_cubic(closure, d, 1e-6)
                    else:
                        t = self._linesearch_backtrack(closure, d, flat_grad, alphabar)
                finally:
                    if not cost_use_gradient:
                        torch.set_grad_enabled(True)
                if math.isnan(t):
                    print("Warning: stepsize nan")
                    t = lr
            self._add_grad(t, d)
            if be_verbose:
                print("step size=%f" % (t))
            if n_iter != max_iter:
                # re-evaluate unless this was the last iteration (stochastic setting, :706-711)
                loss = float(closure())
                flat_grad = gather()
                abs_grad_sum = float(flat_grad.abs().sum())
                if math.isnan(abs_grad_sum):
                    print("Warning: gradient nan")
                    break
                ls_func_evals = 1
            current_evals += ls_func_evals
            state["func_evals"] += ls_func_evals
            # ---------------------------------------------------------------- termination (:725-741)
            if n_iter == max_iter or current_evals >= max_eval or abs_grad_sum <= tolerance_grad:
                break
            if gtd > -tolerance_change:
                break
            if float(d.mul(t).abs_().sum()) <= tolerance_change:
                break
            if abs(loss - prev_loss) < tolerance_change:
                break

        state["d"], state["t"] = d, t
        state["old_dirs"], state["old_stps"], state["H_diag"] = old_dirs, old_stps, H_diag
        state["prev_flat_grad"], state["prev_loss"] = prev_flat_grad, prev_loss
        if batch_mode:
            if running_avg is None:
                running_avg = state.get("running_avg")
                running_avg_sq = state.get("running_avg_sq")
            if running_avg is None:
                running_avg = torch.zeros_like(flat_grad)
                running_avg_sq = torch.zeros_like(flat_grad)
            state["running_avg"], state["running_avg_sq"] = running_avg, running_avg_sq
        return orig_loss
