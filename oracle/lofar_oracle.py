"""CPU (torch fp32) restatement of the reference hot path.  TEST INFRASTRUCTURE ONLY.

Every function cites the reference lines it follows (paths relative to
/root/reference).  The conv / linear / FFT arithmetic is PyTorch's own (the reference
has no arithmetic of its own below torch, SURVEY.md §8c "Third-party arithmetic");
what is restated here is the *composition*: layer order, activations, the cascade,
the loss terms, the K-harmonic formulas, the loader's index map.

Everything is functional over plain dicts of tensors keyed exactly like the
reference ``state_dict`` (``conv0.weight`` ... ``tconv5.bias``; ``M``), so the same
parameter dict can be loaded into the reference modules, this oracle and the CUDA
modules.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Params = Dict[str, torch.Tensor]

CONV_CHANNELS = (8, 12, 24, 48, 96, 192)  # src/lofar_models.py:31-41, :115-125
EPS = 1e-9  # src/lofar_models.py:195


# --------------------------------------------------------------------------------------
# parameter construction (deterministic, numpy-seeded so it is platform independent)
# --------------------------------------------------------------------------------------
def ae_param_shapes(latent_dim: int, channels: int, harmonic_dim: int, rica: bool, ndim: int):
    """Shapes of every tensor in an autoencoder ``state_dict``.

    ndim=2 -> AutoEncoderCNN2 (src/lofar_models.py:31-57), ndim=1 -> AutoEncoder1DCNN
    (src/lofar_models.py:115-142).  Conv weights are [Cout,Cin,4(,4)], transposed-conv
    weights are [Cin,Cout,4(,4)].
    """
    k = (4, 4) if ndim == 2 else (4,)
    chans = (channels,) + CONV_CHANNELS
    shapes = {}
    for i in range(6):
        shapes[f"conv{i}.weight"] = (chans[i + 1], chans[i]) + k
        shapes[f"conv{i}.bias"] = (chans[i + 1],)
    shapes["fcuv1.weight"] = (harmonic_dim, harmonic_dim)
    shapes["fcuv1.bias"] = (harmonic_dim,)
    shapes["fcuv3.weight"] = (harmonic_dim, harmonic_dim)
    shapes["fcuv3.bias"] = (harmonic_dim,)
    shapes["fc1.weight"] = (latent_dim, 768 + harmonic_dim)
    shapes["fc1.bias"] = (latent_dim,)
    if rica:
        shapes["fc2in.weight"] = (latent_dim, latent_dim)
        shapes["fc2in.bias"] = (latent_dim,)
        shapes["fc2out.weight"] = (latent_dim, latent_dim)
        shapes["fc2out.bias"] = (latent_dim,)
    shapes["fc3.weight"] = (768, latent_dim + harmonic_dim)
    shapes["fc3.bias"] = (768,)
    rch = chans[::-1]
    for i in range(6):
        shapes[f"tconv{i}.weight"] = (rch[i], rch[i + 1]) + k
        shapes[f"tconv{i}.bias"] = (rch[i + 1],)
    return shapes


def make_ae_params(latent_dim, channels, harmonic_dim=16, rica=True, ndim=2, seed=0) -> Params:
    """Deterministic parameters: U(-b, b), b = 1/sqrt(fan_in) (torch's default scale)."""
    rng = np.random.default_rng(seed)
    out = {}
    for name, shp in ae_param_shapes(latent_dim, channels, harmonic_dim, rica, ndim).items():
        if name.endswith(".weight"):
            if name.startswith("tconv"):
                fan_in = shp[1] * int(np.prod(shp[2:]))  # torch: weight.size(1)*receptive field
            else:
                fan_in = int(np.prod(shp[1:]))
            last_fan_in = fan_in
        else:
            fan_in = last_fan_in
        b = 1.0 / math.sqrt(fan_in)
        out[name] = torch.from_numpy(rng.uniform(-b, b, size=shp).astype(np.float32))
    return out


def make_centres(K, latent_dim, seed=0) -> torch.Tensor:
    """Centres M ~ U(0,1) like ``torch.rand`` in src/lofar_models.py:197."""
    rng = np.random.default_rng(seed)
    return torch.from_numpy(rng.uniform(0.0, 1.0, size=(K, latent_dim)).astype(np.float32))


# --------------------------------------------------------------------------------------
# autoencoders
# --------------------------------------------------------------------------------------
def uv_harmonics(uv: torch.Tensor, scales: torch.Tensor) -> torch.Tensor:
    """[B,2] -> [B,4H]: columns [sin(s0 u), sin(s0 v), sin(s1 u), ..., cos(s0 u), ...].

    src/lofar_models.py:60-62 (kron of a 1-D scale vector with uv, then cat(sin,cos)).
    """
    z = (scales.reshape(1, -1, 1) * uv.reshape(uv.shape[0], 1, 2)).reshape(uv.shape[0], -1)
    return torch.cat((torch.sin(z), torch.cos(z)), dim=1)


def _conv(ndim):
    return (F.conv2d, F.conv_transpose2d) if ndim == 2 else (F.conv1d, F.conv_transpose1d)


def ae_encode(p: Params, x, uvh, ndim: int):
    """src/lofar_models.py:71-84 (2D: k4 s2 p1) / :156-169 (1D: k4 s4 p1)."""
    conv, _ = _conv(ndim)
    stride = 2 if ndim == 2 else 4
    for i in range(6):
        x = F.elu(conv(x, p[f"conv{i}.weight"], p[f"conv{i}.bias"], stride=stride, padding=1))
    x = torch.flatten(x, start_dim=1)
    u = F.elu(F.linear(uvh, p["fcuv1.weight"], p["fcuv1.bias"]))
    x = torch.cat((x, u), dim=1)
    return F.elu(F.linear(x, p["fc1.weight"], p["fc1.bias"]))


def ae_decode(p: Params, z, uvh, ndim: int):
    """src/lofar_models.py:86-99 (2D: k4 s2 p1) / :171-184 (1D: k4 s4 p0)."""
    _, tconv = _conv(ndim)
    stride, pad = (2, 1) if ndim == 2 else (4, 0)
    u = F.elu(F.linear(uvh, p["fcuv3.weight"], p["fcuv3.bias"]))
    x = F.linear(torch.cat((z, u), dim=1), p["fc3.weight"], p["fc3.bias"])
    x = x.reshape((-1, 192, 2, 2) if ndim == 2 else (-1, 192, 4))
    for i in range(5):
        x = F.elu(tconv(x, p[f"tconv{i}.weight"], p[f"tconv{i}.bias"], stride=stride, padding=pad))
    return tconv(x, p["tconv5.weight"], p["tconv5.bias"], stride=stride, padding=pad)


def ae_forward(p: Params, x, uv, scales, ndim: int, rica: bool = True):
    """forward(x,uv) -> (xhat, mu).  src/lofar_models.py:59-69 / :144-154.

    With rica the returned latent is the post-fc2in one and the decoder eats the
    post-fc2out one.  (rica=False on the 1-D net raises in the reference,
    src/lofar_models.py:150; here it simply decodes mu.)
    """
    uvh = uv_harmonics(uv, scales)
    mu = ae_encode(p, x, uvh, ndim)
    if not rica:
        return ae_decode(p, mu, uvh, ndim), mu
    mu = F.elu(F.linear(mu, p["fc2in.weight"], p["fc2in.bias"]))
    mup = F.elu(F.linear(mu, p["fc2out.weight"], p["fc2out.bias"]))
    return ae_decode(p, mup, uvh, ndim), mu


# --------------------------------------------------------------------------------------
# K-harmonic means
# --------------------------------------------------------------------------------------
def khm_loss_loops(X, M, p):
    """Literal N x K double loop of src/lofar_models.py:199-209 (small cases only)."""
    n, L = X.shape
    K = M.shape[0]
    loss = 0
    for i in range(n):
        ek = 0
        for k in range(K):
            ek = ek + 1.0 / (torch.pow(torch.linalg.norm(M[k, :] - X[i, :], 2), p) + EPS)
        loss = loss + K / (ek + EPS)
    return loss / (n * K * L)


def khm_loss(X, M, p):
    """Vectorised form of src/lofar_models.py:199-209 (same arithmetic per element)."""
    n, L = X.shape
    K = M.shape[0]
    d = torch.linalg.norm(M.unsqueeze(0) - X.unsqueeze(1), dim=2)  # [n,K]
    e = (1.0 / (torch.pow(d, p) + EPS)).sum(dim=1)
    return (K / (e + EPS)).sum() / (n * K * L)


def khm_grads_analytic(X, M, p):
    """Closed-form gradient of :func:`khm_loss`, float64 internally."""
    X64, M64 = X.double(), M.double()
    n, L = X.shape
    K = M.shape[0]
    diff = X64.unsqueeze(1) - M64.unsqueeze(0)  # x_i - m_k
    d2 = (diff * diff).sum(dim=2)
    d = d2.sqrt()
    t = d.pow(p) + EPS
    e = (1.0 / t).sum(dim=1, keepdim=True)
    w = K / (e + EPS) ** 2 * p * torch.where(d > 0, d.pow(p - 2), torch.zeros_like(d)) / t**2
    w = torch.where(d > 0, w, torch.zeros_like(w))
    scale = 1.0 / (n * K * L)
    gx = scale * (w.unsqueeze(2) * diff).sum(dim=1)
    gm = -scale * (w.unsqueeze(2) * diff).sum(dim=0)
    return gx.float(), gm.float()


def cluster_similarity(M):
    """src/lofar_models.py:214-229 (contrastive penalty between centres)."""
    K, L = M.shape
    nrm = torch.linalg.norm(M, dim=1)
    G = M @ M.t()
    S = torch.exp(G / (nrm.unsqueeze(1) * nrm.unsqueeze(0) + EPS))
    numer = S.sum(dim=1) - torch.diagonal(S)
    denom = torch.exp(torch.diagonal(G) / (nrm * nrm + EPS))
    return (numer / (denom + EPS)).sum() / (K * L)


def cluster_similarity_loops(M):
    """Literal K x K loop of src/lofar_models.py:214-229."""
    K, L = M.shape
    loss = 0
    for i in range(K):
        ni = torch.linalg.norm(M[i], 2)
        den = torch.exp(torch.dot(M[i], M[i]) / (ni * ni + EPS))
        num = 0
        for j in range(K):
            if j != i:
                num = num + torch.exp(torch.dot(M[i], M[j]) / (ni * torch.linalg.norm(M[j], 2) + EPS))
        loss = loss + num / (den + EPS)
    return loss / (K * L)


def offline_update(X, M, p):
    """Centre update, Zhang GKHM eq. 7.1-7.5 as *intended* by src/lofar_models.py:231-261.

    The reference body cannot run (``torch.linlag`` typo at :248, in-place write into a
    leaf Parameter at :258); this follows its comments at :241,:246,:249,:252,:256.
    Returns (M_new, numerator [K,L], denominator [K]).
    """
    X64, M64 = X.double(), M.double()
    d = torch.linalg.norm(M64.unsqueeze(0) - X64.unsqueeze(1), dim=2)
    e = (1.0 / (d.pow(p) + EPS)).sum(dim=1)
    alpha = 1.0 / (e**2 + EPS)
    Q = alpha.unsqueeze(1) / (d.pow(p + 2) + EPS)
    num = Q.t() @ X64
    den = Q.sum(dim=0)
    return (num / den.unsqueeze(1)).float(), num.float(), den.float()


def augmented_loss(mu, batch_per_bline, batch_size):
    """src/kharmonic_lofar.py:97-110: rows [ck*bpb,(ck+1)*bpb) form one baseline group."""
    loss = torch.zeros(1, dtype=mu.dtype)
    for ck in range(batch_size):
        Z = mu[ck * batch_per_bline:(ck + 1) * batch_per_bline]
        Zn = Z / (torch.linalg.norm(Z, dim=1, keepdim=True) + 1e-6)
        E = torch.exp(-(Zn @ Zn.t()))
        prod = torch.triu(E, diagonal=1).sum()
        loss = loss + prod / batch_per_bline
    return loss / (batch_size * batch_per_bline)


def eval_distances(Mu, M, p):
    """src/evaluate_clustering.py:110-119: dist_k = mean_n ||Mu_n - M_k||^p, argmin_k.

    Returns (dist [K] fp32, baseline cluster id, per-patch argmin_k d_nk).
    """
    d = torch.linalg.norm(Mu.unsqueeze(1) - M.unsqueeze(0), dim=2)  # [n,K]
    dist = torch.pow(d, p).sum(dim=0) / Mu.shape[0]
    _, idx = torch.min(dist.view(-1, 1), 0)
    return dist, int(idx[0]), torch.argmin(d, dim=1)


# --------------------------------------------------------------------------------------
# cascade closure (training step) and multiplier update
# --------------------------------------------------------------------------------------
def cascade_forward(pn, pT, pF, x, uv, scales, rica=True):
    """src/kharmonic_lofar.py:135-150 (also :188-198 and src/evaluate_clustering.py:81-91)."""
    x1, mu = ae_forward(pn, x, uv, scales, 2, rica)
    x11 = (x - x1) / 2
    yT, muT = ae_forward(pT, torch.flatten(x11, 2, 3), uv, scales, 1, rica)
    x2 = yT.view_as(x11)
    yF, muF = ae_forward(pF, torch.flatten(torch.transpose(x11, 2, 3), 2, 3), uv, scales, 1, rica)
    x3 = torch.transpose(yF.view_as(x11), 2, 3)
    return x1, x11, x2, x3, mu, muT, muF


def closure_losses(pn, pT, pF, M, x, uv, scales, y1, y2, y3, *, batch_per_bline, batch_size,
                   Khp=4, alpha=0.01, beta=0.01, gamma=0.01, rho=1.0, rica=True, rica_lambda=0.01,
                   shard=None):
    """Loss terms of src/kharmonic_lofar.py:132-172, in the order of its print at :179.

    Returns (total, dict of terms).  Differentiable w.r.t. every tensor in pn/pT/pF and M.

    shard=(global_patches, world): the rows given are one data-parallel shard (whole baseline
    groups) of a global batch; every per-patch sum is divided by the GLOBAL constant and the
    replicated centre penalty is divided by `world`, so that the SUM over shards of the returned
    totals (and of their gradients) equals the unsharded value.  shard=None is the reference.
    """
    n_local = x.shape[0]
    n_global, world = (n_local, 1) if shard is None else shard
    x1, x11, x2, x3, mu, muT, muF = cascade_forward(pn, pT, pF, x, uv, scales, rica)
    xrecon = x1 + x2 + x3
    n = x.numel() // n_local * n_global
    sse = lambda a, b: ((a - b) ** 2).sum()
    loss0 = sse(xrecon, x) / n
    loss1 = (torch.dot(y1, (x - x1).reshape(-1)) + rho / 2 * sse(x, x1)) / n
    loss2 = (torch.dot(y2, (x11 - x2).reshape(-1)) + rho / 2 * sse(x11, x2)) / n
    loss3 = (torch.dot(y3, (x11 - x3).reshape(-1)) + rho / 2 * sse(x11, x3)) / n
    Mu = torch.cat((mu, muT, muF), 1)
    frac = n_local / n_global
    kdist = alpha * khm_loss(Mu, M, Khp) * frac
    sim = beta * cluster_similarity(M) / world
    aug = gamma * augmented_loss(Mu, batch_per_bline, batch_size).reshape(()) * frac
    total = loss0 + loss1 + loss2 + loss3 + kdist + aug + sim
    terms = dict(loss0=loss0, loss1=loss1, loss2=loss2, loss3=loss3, kdist=kdist, aug=aug, sim=sim)
    if rica:
        logcosh = lambda z: torch.log(torch.cosh(z)).sum() / z.numel()
        rl = rica_lambda * (logcosh(mu) + logcosh(muT) + logcosh(muF)) * frac
        total = total + rl
        terms["rica"] = rl
    terms["Mu"] = Mu
    return total, terms


def multiplier_update(pn, pT, pF, x, uv, scales, y1, y2, y3, rho=1.0, rica=True):
    """src/kharmonic_lofar.py:187-202: y_i += rho * r_i after a no-grad forward."""
    with torch.no_grad():
        x1, x11, x2, x3, *_ = cascade_forward(pn, pT, pF, x, uv, scales, rica)
        return (y1 + rho * (x - x1).reshape(-1),
                y2 + rho * (x11 - x2).reshape(-1),
                y3 + rho * (x11 - x3).reshape(-1))


# --------------------------------------------------------------------------------------
# loader: scale, patchify, clamp, normalise, uv  (numpy index-map restatement)
# --------------------------------------------------------------------------------------
def assemble_channels(vis: np.ndarray, scale: np.ndarray, baselines: Sequence[int],
                      num_channels: int, patch_size: int) -> np.ndarray:
    """int8 [nbase,T,F,4,2] x fp32 [nbase,F,4] -> fp32 [nb,C,max(T,P),max(F,P)].

    src/lofar_tools.py:86,113-141: channel 2*ci+ri for C=8; C=4 takes pols 0 and 3.
    """
    nbase, T, Fq, npol, _ = vis.shape
    pols = (0, 1, 2, 3) if num_channels == 8 else (0, 3)
    x = np.zeros((len(baselines), num_channels, max(T, patch_size), max(Fq, patch_size)), np.float32)
    for k, b in enumerate(baselines):
        for j, ci in enumerate(pols):
            sf = scale[b, :, ci].astype(np.float32)[None, :]
            for ri in range(2):
                x[k, 2 * j + ri, :T, :Fq] = vis[b, :, :, ci, ri].astype(np.float32) * sf
    return x


def patchify(x: np.ndarray, patch_size: int) -> Tuple[int, int, np.ndarray]:
    """Half-overlapping windows, rows ordered patch-major: n = (ci*py+cj)*nb + k.

    src/lofar_tools.py:157-173 (and :303-319).
    """
    nb, C, T, Fq = x.shape
    s = patch_size // 2
    px = (T - patch_size) // s + 1
    py = (Fq - patch_size) // s + 1
    y = np.zeros((nb * px * py, C, patch_size, patch_size), np.float32)
    ck = 0
    for ci in range(px):
        for cj in range(py):
            y[ck * nb:(ck + 1) * nb] = x[:, :, ci * s:ci * s + patch_size, cj * s:cj * s + patch_size]
            ck += 1
    return px, py, y


def clamp_normalise(y: np.ndarray, clamp: float, normalise: bool) -> np.ndarray:
    """src/lofar_tools.py:187-193 / :333-338: clamp, then (y-mean)/std with UNBIASED std."""
    t = torch.from_numpy(y).clone()
    t.clamp_(-clamp, clamp)
    if normalise:
        t.sub_(t.mean()).div_(t.std())
    return t.numpy()


def uv_coordinates(xyz: np.ndarray, baselines: np.ndarray, sel: Sequence[int], start_time_h: float,
                   freq0: float, reps: int) -> np.ndarray:
    """src/lofar_tools.py:90-110,143-151,175-178: rotate (xx,yy) by theta, in wavelengths;
    each baseline's (u,v) repeated ``reps`` = px*py times, BASELINE-major."""
    c = 2.99792458e8
    theta = start_time_h / 24.0 * (2 * math.pi)
    inv_lambda = freq0 / c
    r00 = math.cos(theta) * inv_lambda
    r01 = math.sin(theta) * inv_lambda
    uv = np.zeros((len(sel), 2), np.float32)
    for k, b in enumerate(sel):
        s1, s2 = baselines[b]
        xx = xyz[s1][0] - xyz[s2][0]
        yy = xyz[s1][1] - xyz[s2][1]
        uv[k, 0] = xx * r00 + yy * r01
        uv[k, 1] = -xx * r01 + yy * r00
    return np.repeat(uv, reps, axis=0)


def load_minibatch(vis, scale, baselines, *, patch_size=128, num_channels=8, normalise=True,
                   clamp=1e3):
    """get_data_minibatch body (src/lofar_tools.py:113-193) for a given baseline draw."""
    x = assemble_channels(vis, scale, baselines, num_channels, patch_size)
    px, py, y = patchify(x, patch_size)
    return px, py, clamp_normalise(y, clamp, normalise)


# --------------------------------------------------------------------------------------
# Fourier features (notebook path)
# --------------------------------------------------------------------------------------
def fft_features(x, xhat: Optional[torch.Tensor] = None, clamp: float = 10.0, mode: str = "reim"):
    """Demo.ipynb:169-174 with src/lofar_tools.py:24-30: ortho 2-D FFT of (x - xhat),
    roll by size//2 on dims 2,3, cat(Re, Im) on the channel axis, clamp to +-10.
    mode="magphase" (north_star's wording; the reference feature is Re / Im): cat(|F|, angle F) of the same
    shifted spectrum, the magnitude clamped."""
    r = x if xhat is None else x - xhat
    f = torch.fft.fftn(r, dim=(2, 3), norm="ortho")
    re, im = f.real, f.imag
    for dim in (2, 3):
        re = torch.roll(re, dims=dim, shifts=re.size(dim) // 2)
        im = torch.roll(im, dims=dim, shifts=im.size(dim) // 2)
    if mode == "magphase":
        return torch.cat((torch.sqrt(re * re + im * im).clamp_(max=clamp), torch.atan2(im, re)), 1)
    y = torch.cat((re, im), 1)
    return y.clamp_(min=-clamp, max=clamp)
