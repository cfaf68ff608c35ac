"""The reference's OWN modules driven by a restatement of its training script.  TEST / BASELINE INFRASTRUCTURE ONLY.

`/root/reference/src/kharmonic_lofar.py` is module-level script code with hard-coded data paths:
it cannot be imported.  What can be imported, unmodified, are the files it is made of -
`lofar_models.py` (AutoEncoderCNN2, AutoEncoder1DCNN, Kmeans) and `lbfgsnew.py` (LBFGSNew).  They are looked up in

  1. `baseline/_ref/`   (verbatim copies made by `__graft_entry__.build()` when /root/reference is present; the
                         directory is git-ignored and travels to the GPU box with the snapshot), then
  2. `/root/reference/src` (the authoring container).

`ReferenceLoop` restates the script around them line by line: module construction (:59-65), optimiser (:84-93),
`augmented_loss` (:97-110), the closure (:132-182) and the multiplier update (:187-202).  Used by
`bench.py --impl reference` / `cpu_baseline` (kind "reference") as the timed CPU arm and by the tests as the
trajectory oracle for LBFGSNew.  Nothing under `lshm_b200/` imports this file.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import time
from typing import Optional

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CANDIDATES = (os.path.join(ROOT, "baseline", "_ref"), "/root/reference/src")


def reference_dir() -> Optional[str]:
    for d in CANDIDATES:
        if os.path.exists(os.path.join(d, "lofar_models.py")) and os.path.exists(os.path.join(d, "lbfgsnew.py")):
            return d
    return None


def load_reference_module(name: str):
    """Import `name`.py of the reference under the private module name `_lshm_reference_<name>`."""
    d = reference_dir()
    if d is None:
        raise ImportError("the reference sources are neither under baseline/_ref nor under /root/reference/src")
    key = f"_lshm_reference_{name}"
    if key in sys.modules:
        return sys.modules[key]
    spec = importlib.util.spec_from_file_location(key, os.path.join(d, name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[key] = mod
    spec.loader.exec_module(mod)
    return mod


def augmented_loss(mu, batch_per_bline, batch_size):
    """src/kharmonic_lofar.py:97-110, restated (same loops, same arithmetic order)."""
    loss = torch.zeros(1)
    for ck in range(batch_size):
        Z = mu[ck * batch_per_bline:(ck + 1) * batch_per_bline, :]
        prod = torch.zeros(1)
        for ci in range(batch_per_bline):
            zi = Z[ci, :] / (torch.norm(Z[ci, :]) + 1e-6)
            for cj in range(ci + 1, batch_per_bline):
                zj = Z[cj, :] / (torch.norm(Z[cj, :]) + 1e-6)
                prod = prod + torch.exp(-torch.dot(zi, zj))
        loss = loss + prod / batch_per_bline
    return loss / (batch_size * batch_per_bline)


class ReferenceLoop:
    """One minibatch of the reference training loop on the reference's modules (CPU).

    optimizer: "adam" (torch.optim.Adam lr 1e-4, :92) or "lbfgs"
    (LBFGSNew(history_size=7, max_iter=4, line_search_fn=True, batch_mode=True), :93), over `param_modules`
    (indices into [net, netT, netF, mod]; the script as shipped uses (0,), BASELINE's configs all four).
    """

    def __init__(self, L=32, Lt=16, C=8, K=10, Khp=4, alpha=0.01, beta=0.01, gamma=0.01, rho=1.0, use_rica=True,
                 rica_lambda=0.01, scales=(1e-4, 1e-3, 1e-2, 1e-1), optimizer="adam", param_modules=(0, 1, 2, 3),
                 state=None, seed=0, lr=None):
        RM = load_reference_module("lofar_models")
        torch.manual_seed(seed)
        hs = torch.tensor(list(scales))
        self.net = RM.AutoEncoderCNN2(latent_dim=L, channels=C, harmonic_scales=hs, rica=use_rica)
        self.netT = RM.AutoEncoder1DCNN(latent_dim=Lt, channels=C, harmonic_scales=hs, rica=use_rica)
        self.netF = RM.AutoEncoder1DCNN(latent_dim=Lt, channels=C, harmonic_scales=hs, rica=use_rica)
        self.mod = RM.Kmeans(latent_dim=(L + Lt + Lt), K=K, p=Khp)
        self.modules = [self.net, self.netT, self.netF, self.mod]
        if state is not None:       # (pn, pT, pF, M) dicts / tensor keyed like the reference state_dict
            self.net.load_state_dict(state[0]); self.netT.load_state_dict(state[1]); self.netF.load_state_dict(state[2])
            self.mod.load_state_dict({"M": state[3]})
        self.alpha, self.beta, self.gamma, self.rho = alpha, beta, gamma, rho
        self.use_rica, self.rica_lambda = use_rica, rica_lambda
        self.criterion = torch.nn.MSELoss(reduction="sum")
        params = []
        for mi in param_modules:
            params.extend(list(self.modules[mi].parameters()))
        if optimizer == "adam":
            self.optimizer = torch.optim.Adam(params, lr=1e-4 if lr is None else lr)
        elif optimizer == "lbfgs":
            LB = load_reference_module("lbfgsnew")
            self.optimizer = LB.LBFGSNew(params, history_size=7, max_iter=4, line_search_fn=True, batch_mode=True)
        else:
            raise ValueError(optimizer)
        self.khm_seconds = 0.0      # time inside Kmeans.clustering_error (the Python N x K loop), forward only
        self.closures = 0
        self.last_terms = None

    def set_batch(self, x, uv, batch_per_bline):
        """:118-130 after the loader: x [N,C,128,128], uv [N,2]; multipliers restart."""
        self.x, self.uv, self.bpb = x, uv, batch_per_bline
        self.default_batch = x.shape[0] // batch_per_bline
        n = x.numel()
        self.y1, self.y2, self.y3 = torch.zeros(n), torch.zeros(n), torch.zeros(n)

    def closure(self):
        """:132-182 (the print at :179 is replaced by storing the same columns)."""
        x, uv, rho, criterion = self.x, self.uv, self.rho, self.criterion
        if torch.is_grad_enabled():
            self.optimizer.zero_grad()
        x1, mu = self.net(x, uv)
        x11 = (x - x1) / 2
        iy1 = torch.flatten(x11, start_dim=2, end_dim=3)
        yyT, yyTmu = self.netT(iy1, uv)
        x2 = yyT.view_as(x11)
        iy2 = torch.flatten(torch.transpose(x11, 2, 3), start_dim=2, end_dim=3)
        yyF, yyFmu = self.netF(iy2, uv)
        x3 = torch.transpose(yyF.view_as(x11), 2, 3)
        xrecon = x1 + x2 + x3
        loss0 = (criterion(xrecon, x)) / (x.numel())
        loss1 = (torch.dot(self.y1, (x - x1).view(-1)) + rho / 2 * criterion(x, x1)) / (x.numel())
        loss2 = (torch.dot(self.y2, (x11 - x2).view(-1)) + rho / 2 * criterion(x11, x2)) / (x.numel())
        loss3 = (torch.dot(self.y3, (x11 - x3).reshape(-1)) + rho / 2 * criterion(x11, x3)) / (x.numel())
        Mu = torch.cat((mu, yyTmu, yyFmu), 1)
        t0 = time.perf_counter()
        kdist = self.alpha * self.mod.clustering_error(Mu)
        self.khm_seconds += time.perf_counter() - t0
        clus_sim = self.beta * self.mod.cluster_similarity()
        augmentation_loss = self.gamma * augmented_loss(Mu, self.bpb, self.default_batch)
        loss = loss0 + loss1 + loss2 + loss3 + kdist + augmentation_loss + clus_sim
        rica_loss = torch.zeros(())
        if self.use_rica:
            rica_loss = self.rica_lambda * (torch.sum(torch.log(torch.cosh(mu))) / mu.numel()
                                            + torch.sum(torch.log(torch.cosh(yyTmu))) / yyTmu.numel()
                                            + torch.sum(torch.log(torch.cosh(yyFmu))) / yyFmu.numel())
            loss += rica_loss
        if loss.requires_grad:
            loss.backward(retain_graph=True)
        self.closures += 1
        f = lambda t: float(t.detach())
        self.last_terms = dict(total=f(loss), loss0=f(loss0), loss1=f(loss1), loss2=f(loss2), loss3=f(loss3),
                               kdist=f(kdist), aug=f(augmentation_loss), sim=f(clus_sim), rica=f(rica_loss))
        return loss

    def update_multipliers(self):
        """:187-202."""
        x, uv, rho = self.x, self.uv, self.rho
        with torch.no_grad():
            x1, _ = self.net(x, uv)
            x11 = (x - x1) / 2
            iy1 = torch.flatten(x11, start_dim=2, end_dim=3)
            yyT, _ = self.netT(iy1, uv)
            x2 = yyT.view_as(x11)
            iy2 = torch.flatten(torch.transpose(x11, 2, 3), start_dim=2, end_dim=3)
            yyF, _ = self.netF(iy2, uv)
            x3 = torch.transpose(yyF.view_as(x11), 2, 3)
            self.y1 = self.y1 + rho * (x - x1).view(-1)
            self.y2 = self.y2 + rho * (x11 - x2).view(-1)
            self.y3 = self.y3 + rho * (x11 - x3).reshape(-1)

    def admm_iteration(self):
        """One pass of the `for admm` body (:131-202): optimizer.step(closure) + multiplier update."""
        loss = self.optimizer.step(self.closure)
        self.update_multipliers()
        return loss

    def state(self):
        return (self.net.state_dict(), self.netT.state_dict(), self.netF.state_dict(), self.mod.M.detach().clone())
