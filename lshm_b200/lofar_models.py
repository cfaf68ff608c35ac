"""Drop-in replacements for the reference modules of /root/reference/src/lofar_models.py.

Same class names, constructor signatures, attribute names, sub-module names and
``state_dict`` keys/shapes (``conv0..5``, ``fcuv1``, ``fcuv3``, ``fc1``, ``fc2in``,
``fc2out``, ``fc3``, ``tconv0..5`` ``.weight/.bias``; ``M``), so checkpoints written by
the reference (src/kharmonic_lofar.py:210-222) load unchanged and the reference loops
(src/kharmonic_lofar.py:132-202, src/evaluate_clustering.py:75-119) and the
``lbfgsnew.LBFGSNew`` closure protocol work as before.  The sub-modules are only
parameter containers: the arithmetic runs in liblshm_sm100 (hand-written sm_100a
kernels) with an analytic backward; CUDA tensors only, no CPU fallback.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from ._lib import lib
from .engine import AEEngine, param_names

__all__ = ["AutoEncoderCNN2", "AutoEncoder1DCNN", "Kmeans", "augmented_loss"]


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _require_cuda(t: torch.Tensor, what: str):
    if not t.is_cuda:
        raise RuntimeError(f"lshm_b200: {what} must be a CUDA tensor (this build has no CPU path)")
    if t.dtype != torch.float32:
        raise RuntimeError(f"lshm_b200: {what} must be float32, got {t.dtype}")


class _AEFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, engine, scales, names, track, x, uv, *params):
        N = x.shape[0]
        xc = x.contiguous()
        uvc = uv.contiguous()
        need_dx = track and x.requires_grad
        ws = engine.workspace(N, x.device, with_grad=track, need_dx=need_dx)
        p = dict(zip(names, params))
        xhat, mu = engine.forward(xc.view(N, -1), uvc, scales, p, ws, _stream())
        if track:
            ctx.engine, ctx.names, ctx.ws, ctx.need_dx = engine, names, ws, need_dx
            ctx.xc = xc
            ctx.params = params
        return xhat.view(x.shape), mu

    @staticmethod
    def backward(ctx, g_xhat, g_mu):
        engine, ws, names = ctx.engine, ctx.ws, ctx.names
        N = ws.N
        p = dict(zip(names, ctx.params))
        g = {nm: torch.empty_like(t) for nm, t in p.items()}
        gx = None if g_xhat is None else g_xhat.contiguous().view(N, -1)
        gm = None if g_mu is None else g_mu.contiguous()
        dx = engine.backward(ctx.xc.view(N, -1), p, g, ws, _stream(), gx, gm, ws.mu, ctx.need_dx)
        if dx is not None:
            dx = dx.view(ctx.xc.shape)
        return (None, None, None, None, dx, None) + tuple(g[nm] for nm in names)


ENC_NAMES = tuple(f"conv{i}.{w}" for i in range(6) for w in ("weight", "bias")) + ("fcuv1.weight", "fcuv1.bias", "fc1.weight", "fc1.bias")
DEC_NAMES = ("fcuv3.weight", "fcuv3.bias", "fc3.weight", "fc3.bias") + tuple(f"tconv{i}.{w}" for i in range(6) for w in ("weight", "bias"))


class _EncodeFunction(torch.autograd.Function):
    """encode(x, uvh) with its own backward (the reference's encode is an ordinary differentiable method)."""

    @staticmethod
    def forward(ctx, engine, names, track, x, uvh, *params):
        N = x.shape[0]
        xc, uc = x.contiguous(), uvh.contiguous()
        need_dx = track and x.requires_grad
        ws = engine.workspace(N, x.device, with_grad=track, need_dx=need_dx)
        p = dict(zip(names, params))
        engine.prepare_images(p, _stream(), track)
        out = engine.encode(xc.view(N, -1), uc, p, ws, _stream())
        if track:
            ctx.engine, ctx.names, ctx.ws, ctx.need_dx, ctx.xc, ctx.uc, ctx.params = engine, names, ws, need_dx, xc, uc, params
        return out.clone()

    @staticmethod
    def backward(ctx, g_out):
        engine, ws, names = ctx.engine, ctx.ws, ctx.names
        p = dict(zip(names, ctx.params))
        g = {nm: torch.empty_like(p[nm]) for nm in ENC_NAMES}
        dx, g_uvh = engine.backward_encode(ctx.xc.view(ws.N, -1), ctx.uc, p, g, ws, _stream(), g_out.contiguous(), ctx.need_dx)
        if dx is not None:
            dx = dx.view(ctx.xc.shape)
        return (None, None, None, dx, g_uvh) + tuple(g.get(nm) for nm in names)


class _DecodeFunction(torch.autograd.Function):
    """decode(z, uvh) with its own backward."""

    @staticmethod
    def forward(ctx, engine, names, track, shape, z, uvh, *params):
        N = z.shape[0]
        uc = uvh.contiguous()
        ws = engine.workspace(N, z.device, with_grad=track)
        ws.zcat[:, :engine.L].copy_(z)
        p = dict(zip(names, params))
        engine.prepare_images(p, _stream(), track)
        xhat = engine.decode(uc, p, ws, _stream())
        if track:
            ctx.engine, ctx.names, ctx.ws, ctx.uc, ctx.params = engine, names, ws, uc, params
        return xhat.view(shape).clone()

    @staticmethod
    def backward(ctx, g_xhat):
        engine, ws, names = ctx.engine, ctx.ws, ctx.names
        p = dict(zip(names, ctx.params))
        g = {nm: torch.empty_like(p[nm]) for nm in DEC_NAMES}
        g_z, g_uvh = engine.backward_decode(ctx.uc, p, g, ws, _stream(), g_xhat.contiguous().view(ws.N, -1))
        return (None, None, None, None, g_z, g_uvh) + tuple(g.get(nm) for nm in names)


class _AutoEncoderBase(nn.Module):
    _ndim = 2

    def __init__(self, latent_dim=128, channels=3, harmonic_scales=None, rica=False):
        super().__init__()
        self.rica = rica
        self.latent_dim = latent_dim
        # plain attribute like the reference (src/lofar_models.py:27): callers move it to the device
        self.harmonic_scales = harmonic_scales
        self.harmonic_dim = (self.harmonic_scales.size()[0]) * 2 * 2
        conv, tconv = (nn.Conv2d, nn.ConvTranspose2d) if self._ndim == 2 else (nn.Conv1d, nn.ConvTranspose1d)
        stride, tpad = (2, 1) if self._ndim == 2 else (4, 0)
        ch = (channels, 8, 12, 24, 48, 96, 192)
        for i in range(6):
            setattr(self, f"conv{i}", conv(ch[i], ch[i + 1], 4, stride=stride, padding=1))
        self.fcuv1 = nn.Linear(self.harmonic_dim, self.harmonic_dim)
        self.fcuv3 = nn.Linear(self.harmonic_dim, self.harmonic_dim)
        self.fc1 = nn.Linear(768 + self.harmonic_dim, self.latent_dim)
        if self.rica:
            self.fc2in = nn.Linear(self.latent_dim, self.latent_dim)
            self.fc2out = nn.Linear(self.latent_dim, self.latent_dim)
        self.fc3 = nn.Linear(self.latent_dim + self.harmonic_dim, 768)
        for i in range(6):
            setattr(self, f"tconv{i}", tconv(ch[6 - i], ch[5 - i], 4, stride=stride, padding=tpad))
        self._channels = channels
        self._names = param_names(rica)
        self._engine = None

    # engine is created lazily so the module can be constructed / state-dict-loaded without a GPU
    def engine(self) -> AEEngine:
        if self._engine is None:
            self._engine = AEEngine(self._ndim, self._channels, self.latent_dim, self.harmonic_dim, self.rica)
        return self._engine

    def named_param_dict(self):
        d = dict(self.named_parameters())
        return {nm: d[nm] for nm in self._names}

    def _check_input(self, x, uv):
        _require_cuda(x, "x")
        _require_cuda(uv, "uv")
        expect = (128, 128) if self._ndim == 2 else (16384,)
        if tuple(x.shape[1:]) != (self._channels,) + expect:
            raise RuntimeError(f"lshm_b200: expected input [N,{self._channels},{','.join(map(str, expect))}], "
                               f"got {tuple(x.shape)}")
        if uv.shape != (x.shape[0], 2):
            raise RuntimeError(f"lshm_b200: uv must be [N,2], got {tuple(uv.shape)}")

    def forward(self, x, uv):
        """(xhat, mu) = forward(x, uv); src/lofar_models.py:59-69 / :144-154."""
        self._check_input(x, uv)
        scales = self.harmonic_scales
        if scales.device != x.device:
            scales = scales.to(x.device)
        p = self.named_param_dict()
        track = torch.is_grad_enabled() and (x.requires_grad or any(t.requires_grad for t in p.values()))
        return _AEFunction.apply(self.engine(), scales.contiguous().float(), self._names, track,
                                 x, uv, *p.values())

    # encode / decode keep the reference signatures: `uv` here is the [N,4H] harmonic vector
    # (src/lofar_models.py:71,86), and - like the reference's - they are differentiable.
    def encode(self, x, uv):
        _require_cuda(x, "x")
        _require_cuda(uv, "uv")
        p = self.named_param_dict()
        track = torch.is_grad_enabled() and (x.requires_grad or uv.requires_grad or any(t.requires_grad for t in p.values()))
        return _EncodeFunction.apply(self.engine(), self._names, track, x, uv, *p.values())

    def decode(self, z, uv):
        _require_cuda(z, "z")
        _require_cuda(uv, "uv")
        p = self.named_param_dict()
        track = torch.is_grad_enabled() and (z.requires_grad or uv.requires_grad or any(t.requires_grad for t in p.values()))
        N = z.shape[0]
        shape = (N, self._channels, 128, 128) if self._ndim == 2 else (N, self._channels, 16384)
        return _DecodeFunction.apply(self.engine(), self._names, track, shape, z, uv, *p.values())


class AutoEncoderCNN2(_AutoEncoderBase):
    """2-D CNN autoencoder, 128x128 patches (src/lofar_models.py:12-99)."""
    _ndim = 2


class AutoEncoder1DCNN(_AutoEncoderBase):
    """1-D CNN autoencoder on vectorised patches (src/lofar_models.py:103-184).

    The reference's rica=False forward raises (decode called without uv, :150); here it
    decodes the fc1 latent like the 2-D class.
    """
    _ndim = 1


# ----------------------------------------------------------------------------------------
# K-harmonic means
# ----------------------------------------------------------------------------------------
class _KhmLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, X, M, p):
        N, L = X.shape
        K = M.shape[0]
        Xc, Mc = X.contiguous(), M.contiguous()
        acc = torch.zeros(1, dtype=torch.float64, device=X.device)
        lib().khm_fwd(Xc.data_ptr(), L, Mc.data_ptr(), N, K, L, float(p), acc.data_ptr(), None, _stream())
        ctx.save_for_backward(Xc, Mc)
        ctx.p = float(p)
        return (acc / float(N * K * L)).to(torch.float32).reshape(())

    @staticmethod
    def backward(ctx, gout):
        Xc, Mc = ctx.saved_tensors
        N, L = Xc.shape
        K = Mc.shape[0]
        gX = torch.empty_like(Xc)
        gM = torch.zeros_like(Mc)
        lib().khm_bwd(Xc.data_ptr(), L, Mc.data_ptr(), N, K, L, ctx.p, 1.0 / float(N * K * L),
                      gX.data_ptr(), L, 0, gM.data_ptr(), _stream())
        return gX * gout, gM * gout, None


class _Similarity(torch.autograd.Function):
    @staticmethod
    def forward(ctx, M):
        K, L = M.shape
        Mc = M.contiguous()
        acc = torch.zeros(1, dtype=torch.float64, device=M.device)
        gM = torch.zeros_like(Mc)
        work = torch.empty(2 * K * K, dtype=torch.float32, device=M.device)
        lib().similarity(Mc.data_ptr(), K, L, 1.0, acc.data_ptr(), gM.data_ptr(), work.data_ptr(), _stream())
        ctx.save_for_backward(gM)
        return acc.to(torch.float32).reshape(())

    @staticmethod
    def backward(ctx, gout):
        (gM,) = ctx.saved_tensors
        return gM * gout


class _Augment(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mu, bpb, scale):
        N, L = mu.shape
        muc = mu.contiguous()
        acc = torch.zeros(1, dtype=torch.float64, device=mu.device)
        g = torch.zeros_like(muc)
        lib().augment(muc.data_ptr(), L, N, L, int(bpb), float(scale), acc.data_ptr(), g.data_ptr(), L, _stream())
        ctx.save_for_backward(g)
        return acc.to(torch.float32)

    @staticmethod
    def backward(ctx, gout):
        (g,) = ctx.saved_tensors
        return g * gout, None, None


def augmented_loss(mu, batch_per_bline, batch_size):
    """src/kharmonic_lofar.py:97-110.  Rows [ck*bpb,(ck+1)*bpb) of mu form one group;
    returns a 1-element tensor like the reference."""
    _require_cuda(mu, "mu")
    if mu.shape[0] != batch_per_bline * batch_size:
        raise RuntimeError("lshm_b200: augmented_loss expects batch_per_bline*batch_size rows")
    return _Augment.apply(mu, batch_per_bline, 1.0 / float(batch_per_bline * batch_size * batch_per_bline))


class Kmeans(nn.Module):
    """K-harmonic means module (src/lofar_models.py:189-261)."""

    def __init__(self, latent_dim=128, K=10, p=2):
        super().__init__()
        self.latent_dim = latent_dim
        self.K = K
        self.p = p
        self.EPS = 1e-9
        self.M = torch.nn.Parameter(torch.rand(self.K, self.latent_dim), requires_grad=True)

    def forward(self, X):
        """Harmonic-mean clustering error, src/lofar_models.py:199-209."""
        _require_cuda(X, "X")
        if X.shape[1] != self.latent_dim:
            raise RuntimeError(f"lshm_b200: X must be [N,{self.latent_dim}]")
        return _KhmLoss.apply(X, self.M, self.p)

    def clustering_error(self, X):
        return self.forward(X)

    def cluster_similarity(self):
        """Contrastive penalty between centres, src/lofar_models.py:214-229."""
        _require_cuda(self.M, "M")
        return _Similarity.apply(self.M)

    @torch.no_grad()
    def offline_update(self, X):
        """Centre update, Zhang GKHM eq. 7.1-7.5 (intent of src/lofar_models.py:231-261; the
        reference body raises AttributeError at :248 and is never called).  Returns the
        (numerator [K,L], denominator [K]) sums so a data-parallel caller can all-reduce them
        and re-apply."""
        _require_cuda(X, "X")
        N, L = X.shape
        Xc = X.contiguous()
        num = torch.zeros(self.K, L, dtype=torch.float32, device=X.device)
        den = torch.zeros(self.K, dtype=torch.float32, device=X.device)
        lib().khm_center_sums(Xc.data_ptr(), L, self.M.data_ptr(), N, self.K, L, float(self.p),
                              num.data_ptr(), den.data_ptr(), _stream())
        lib().khm_center_apply(num.data_ptr(), den.data_ptr(), self.M.data_ptr(), self.K, L, _stream())
        return num, den

    @torch.no_grad()
    def assign(self, X):
        """Per-patch nearest centre (argmin_k ||x_n - m_k||, first index on ties)."""
        _require_cuda(X, "X")
        Xc = X.contiguous()
        ids = torch.empty(X.shape[0], dtype=torch.int32, device=X.device)
        lib().khm_assign(Xc.data_ptr(), X.shape[1], self.M.data_ptr(), X.shape[0], self.K, X.shape[1],
                         ids.data_ptr(), _stream())
        return ids

    @torch.no_grad()
    def group_distances(self, X, group, power=None):
        """src/evaluate_clustering.py:110-119 for every group of `group` consecutive rows:
        dist[g,k] = mean_n ||x_n - m_k||^power and its argmin.  power defaults to the module's p;
        power=1 gives the node labels of src/train_graph.py:152-157 (mean Euclidean distance)."""
        _require_cuda(X, "X")
        Xc = X.contiguous()
        G = X.shape[0] // group
        dist = torch.empty(G, self.K, dtype=torch.float32, device=X.device)
        gid = torch.empty(G, dtype=torch.int32, device=X.device)
        lib().khm_group_dist(Xc.data_ptr(), X.shape[1], self.M.data_ptr(), X.shape[0], self.K, X.shape[1],
                             float(self.p if power is None else power), int(group), dist.data_ptr(), gid.data_ptr(),
                             _stream())
        return dist, gid
