"""Shared helpers for the parity tests (seeded inputs, golden fixtures, comparisons)."""
import os

import numpy as np
import torch

from lshm_b200 import synthetic as S
from oracle import lofar_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SCALES = [1e-4, 1e-3, 1e-2, 1e-1]  # /root/reference/src/kharmonic_lofar.py:57
REFERENCE_SRC = "/root/reference/src"


def golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def rel_err(a, b):
    a = torch.as_tensor(a).double().cpu()
    b = torch.as_tensor(b).double().cpu()
    den = b.norm().item()
    return (a - b).norm().item() / (den if den > 0 else 1.0)


def max_abs(a, b):
    return (torch.as_tensor(a).double().cpu() - torch.as_tensor(b).double().cpu()).abs().max().item()


def closure_case(C=8, L=32, Lt=16, K=10, N=8, bpb=4, seed=0, ymag=0.1):
    """Inputs of one training closure (cfg1 shapes by default).  seed=0 reproduces the case of
    oracle/gen_golden.py (closure_cfg1.npz)."""
    pn = O.make_ae_params(L, C, ndim=2, seed=seed + 1)
    pT = O.make_ae_params(Lt, C, ndim=1, seed=seed + 2)
    pF = O.make_ae_params(Lt, C, ndim=1, seed=seed + 3)
    M = O.make_centres(K, L + 2 * Lt, seed=seed + 4)
    x = torch.from_numpy(S.make_patches(N, C, seed=seed + 5))
    uv = torch.from_numpy(S.make_uv(N, seed=seed + 5, per_group=bpb))
    g = torch.Generator().manual_seed(seed + 1)
    ys = [ymag * torch.randn(x.numel(), generator=g) for _ in range(3)]
    return dict(pn=pn, pT=pT, pF=pF, M=M, x=x, uv=uv, ys=ys, C=C, L=L, Lt=Lt, K=K, N=N, bpb=bpb)


def oracle_closure(case, Khp=4, alpha=0.01, beta=0.01, gamma=0.01, rho=1.0, lam=0.01, grads=True):
    """Loss terms and (optionally) every gradient from the CPU oracle via autograd."""
    hs = torch.tensor(SCALES)
    pn = {k: v.clone().requires_grad_(grads) for k, v in case["pn"].items()}
    pT = {k: v.clone().requires_grad_(grads) for k, v in case["pT"].items()}
    pF = {k: v.clone().requires_grad_(grads) for k, v in case["pF"].items()}
    M = case["M"].clone().requires_grad_(grads)
    total, terms = O.closure_losses(pn, pT, pF, M, case["x"], case["uv"], hs, *case["ys"],
                                    batch_per_bline=case["bpb"], batch_size=case["N"] // case["bpb"],
                                    Khp=Khp, alpha=alpha, beta=beta, gamma=gamma, rho=rho, rica_lambda=lam)
    out = {k: float(v.detach()) for k, v in terms.items() if k != "Mu"}
    out["total"] = float(total.detach())
    out["Mu"] = terms["Mu"].detach()
    if grads:
        total.backward()
        g = {}
        for tag, p in (("0", pn), ("1", pT), ("2", pF)):
            for k, v in p.items():
                g[f"{tag}.{k}"] = v.grad
        g["3.M"] = M.grad
        out["grads"] = g
    return out
