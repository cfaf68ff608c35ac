"""Host-side driver of the autoencoder kernels: forward and hand-written backward of
AutoEncoderCNN2 / AutoEncoder1DCNN as sequences of liblshm_sm100 calls.

Follows /root/reference/src/lofar_models.py:59-99 (2-D) and :144-184 (1-D): six strided
convs + ELU, uv-harmonic MLP, fc1, optional RICA pair fc2in/fc2out, fc3, six transposed
convs.  The backward is analytic (no autograd graph): every stored gradient tensor is the
gradient w.r.t. a *pre-activation* (the producing kernel multiplies by ELU'(a), which only
needs the stored post-ELU activation a), concatenations are realised by writing straight
into column slices of one buffer, and weight/bias gradients are written into caller-provided
tensors (views of the flat gradient buffer).

torch is used for device memory, streams and views only.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from ._lib import lib

import os
EPI_NONE, EPI_ELU, EPI_DELU = 0, 1, 2
FUSE_LAST = os.environ.get("LSHM_NO_FUSED_LAST") is None   # (switch for A/B measurements)
CONV_CHANNELS = (8, 12, 24, 48, 96, 192)   # src/lofar_models.py:31-41
FLAT = 768                                  # 192*2*2 (2-D) = 192*4 (1-D)
INPUT_ELEMS = 16384                         # 128*128 pixels or 16384 samples per channel


def _p(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def param_names(rica: bool):
    names = []
    for i in range(6):
        names += [f"conv{i}.weight", f"conv{i}.bias"]
    names += ["fcuv1.weight", "fcuv1.bias", "fcuv3.weight", "fcuv3.bias", "fc1.weight", "fc1.bias"]
    if rica:
        names += ["fc2in.weight", "fc2in.bias", "fc2out.weight", "fc2out.bias"]
    names += ["fc3.weight", "fc3.bias"]
    for i in range(6):
        names += [f"tconv{i}.weight", f"tconv{i}.bias"]
    return names


class Workspace:
    """Activation (+ optionally gradient) buffers of one autoencoder for N samples."""

    def __init__(self, N: int, C: int, L: int, H4: int, device, with_grad: bool, need_dx: bool):
        f = dict(device=device, dtype=torch.float32)
        ch = (C,) + CONV_CHANNELS
        sz = [ch[i] * (INPUT_ELEMS >> (2 * i)) for i in range(7)]  # per-sample elements per level
        self.N, self.sizes = N, sz
        self.uvh = torch.empty(N, H4, **f)
        self.enc = [None] + [torch.empty(N, sz[i], **f) for i in range(1, 6)]  # enc[6] lives in cat1
        self.cat1 = torch.empty(N, FLAT + H4, **f)
        self.mu0 = torch.empty(N, L, **f)
        self.mu = torch.empty(N, L, **f)
        self.zcat = torch.empty(N, L + H4, **f)
        self.dec = [torch.empty(N, sz[6 - i], **f) for i in range(6)]  # dec[0]=fc3 out .. dec[5]
        self.xhat = torch.empty(N, sz[0], **f)
        if with_grad:
            self.g_enc = [None] + [torch.empty(N, sz[i], **f) for i in range(1, 6)]
            self.g_cat1 = torch.empty(N, FLAT + H4, **f)
            self.g_mu0 = torch.empty(N, L, **f)
            self.g_mu = torch.empty(N, L, **f)
            self.g_zcat = torch.empty(N, L + H4, **f)
            self.g_dec = [torch.empty(N, sz[6 - i], **f) for i in range(6)]
            self.dx = torch.empty(N, sz[0], **f) if need_dx else None

    def rows(self, n0: int, n1: int) -> "Workspace":
        """The same buffers restricted to samples [n0, n1): every buffer is [N, per-sample] row-major, so a micro-batch
        is a contiguous slice of each.  The step runs its micro-batches on such views (one stream each) while the
        whole-batch kernels (latent-space terms, multiplier update) keep using the full buffers."""
        v = Workspace.__new__(Workspace)
        v.N, v.sizes = n1 - n0, self.sizes
        for k, t in self.__dict__.items():
            if k in ("N", "sizes"):
                continue
            if isinstance(t, torch.Tensor):
                setattr(v, k, t[n0:n1])
            elif isinstance(t, list):
                setattr(v, k, [None if u is None else u[n0:n1] for u in t])
            else:
                setattr(v, k, t)
        return v


def planes_buffer(dim: int, N: int, Bc: int, h: int, w: int, device) -> torch.Tensor:
    """Zero-filled operand-plane buffer (include/lshm.h "operand planes") for a big map [N,Bc,2h,2w] (dim 2) or
    [N,Bc,4w] (dim 1).  Zero-filled once: the <= 31 padding positions per chunk are never written."""
    import ctypes
    n = ctypes.c_int64()
    if lib().cdll.lshm_planes_bytes(dim, N, Bc, h, w, ctypes.byref(n)) != 0:
        raise RuntimeError("lshm_planes_bytes failed")
    return torch.zeros(n.value, dtype=torch.uint8, device=device)


def conv_image(w: torch.Tensor, dim: int, which: int, st: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """bf16 hi/lo operand image of a conv / transposed-conv weight W[A,Bc,4(,4)] for the tensor-core
    kernels (which: 0 = lshm_down*, 1 = lshm_up*); lshm_conv_prep."""
    A, Bc = w.shape[0], w.shape[1]
    if out is None:
        out = torch.empty(max(16, lib().conv_image_bytes(dim, A, Bc, which)), dtype=torch.uint8, device=w.device)
    lib().conv_prep(w.data_ptr(), dim, A, Bc, out.data_ptr() if which == 0 else None,
                    out.data_ptr() if which == 1 else None, st)
    return out


class AEEngine:
    """Forward / backward of one autoencoder (ndim=2: AutoEncoderCNN2, ndim=1: AutoEncoder1DCNN)."""

    def __init__(self, ndim: int, channels: int, latent_dim: int, harmonic_dim: int, rica: bool):
        assert ndim in (1, 2)
        self.ndim, self.C, self.L, self.H4, self.rica = ndim, channels, latent_dim, harmonic_dim, rica
        self.ch = (channels,) + CONV_CHANNELS
        self.lib = lib()
        self.img = {}   # (layer name, which) -> weight image buffer (re-filled every closure)
        self.need_input_grad = ndim == 1   # the 1-D nets sit behind the 2-D net: dx is always wanted

    def _build_tables(self, p: Dict[str, torch.Tensor]):
        """Allocate the weight images and the two device tables of preparation records
        (forward-only, forward+backward).  Addresses are stable (parameters live in the flat
        buffer), so this runs once; it re-runs if a parameter is re-homed."""
        import numpy as np
        dev = p["conv0.weight"].device
        fwd, bwd = [], []
        for i in range(6):
            cn, tn = f"conv{i}.weight", f"tconv{i}.weight"
            fwd += [(cn, 0), (tn, 1)]
            bwd += [(tn, 0)]
            if i > 0 or self.need_input_grad:
                bwd += [(cn, 1)]
        self.img = {}
        recs = np.zeros((len(fwd) + len(bwd), 8), dtype=np.int64)
        for r, (nm, which) in enumerate(fwd + bwd):
            w = p[nm]
            A, Bc = w.shape[0], w.shape[1]
            img = torch.empty(max(16, self.lib.conv_image_bytes(self.ndim, A, Bc, which)), dtype=torch.uint8, device=dev)
            self.img[(nm, which)] = img
            self.lib.conv_prep_record(w.data_ptr(), self.ndim, A, Bc, which, img.data_ptr(), recs[r].ctypes.data)
        self._table = torch.from_numpy(recs).to(dev)
        self._n_fwd, self._n_all = len(fwd), len(fwd) + len(bwd)
        self._table_key = tuple(p[f"{k}{i}.weight"].data_ptr() for k in ("conv", "tconv") for i in range(6))

    def prepare_images(self, p: Dict[str, torch.Tensor], st: int, grads: bool):
        """Re-make (ONE launch) the weight images the coming forward (and backward) will read: conv
        layers use their 'down' image forward and their 'up' image in dgrad, transposed convs the
        other way round."""
        key = tuple(p[f"{k}{i}.weight"].data_ptr() for k in ("conv", "tconv") for i in range(6))
        if getattr(self, "_table_key", None) != key:
            self._build_tables(p)
        self.lib.conv_prep_batch(self._table.data_ptr(), self._n_all if grads else self._n_fwd, st)

    def workspace(self, N, device, with_grad, need_dx=False) -> Workspace:
        return Workspace(N, self.C, self.L, self.H4, device, with_grad, need_dx)

    # ------------------------------------------------------------------ conv dispatch
    def _down(self, big, big_ns, w, bias, aux, aux_ns, small, small_ns, N, A, Bc, lvl, pad, epi, st):
        # lvl = index of the small map (1..6): small side / length at that level
        if self.ndim == 2:
            s = 128 >> lvl
            self.lib.down2d(big, big_ns, w, bias, aux, aux_ns, small, small_ns, N, A, Bc, s, s, epi, st)
        else:
            self.lib.down1d(big, big_ns, w, bias, aux, aux_ns, small, small_ns, N, A, Bc,
                            INPUT_ELEMS >> (2 * lvl), pad, epi, st)

    def _down_planes(self, planes, w, bias, aux, aux_ns, small, small_ns, N, A, Bc, lvl, epi, st):
        if self.ndim == 2:
            s = 128 >> lvl
            self.lib.down2d_planes(planes, w, bias, aux, aux_ns, small, small_ns, N, A, Bc, s, s, epi, st)
        else:
            self.lib.down1d_planes(planes, w, bias, aux, aux_ns, small, small_ns, N, A, Bc, INPUT_ELEMS >> (2 * lvl), epi, st)

    def _wgrad_planes(self, small, small_ns, planes, dw, N, A, Bc, lvl, st):
        if self.ndim == 2:
            s = 128 >> lvl
            self.lib.wgrad2d_planes(small, small_ns, planes, dw, N, A, Bc, s, s, st)
        else:
            self.lib.wgrad1d_planes(small, small_ns, planes, dw, N, A, Bc, INPUT_ELEMS >> (2 * lvl), st)

    def _up(self, small, small_ns, w, bias, aux, aux_ns, big, big_ns, N, A, Bc, lvl, pad, epi, st):
        if self.ndim == 2:
            s = 128 >> lvl
            self.lib.up2d(small, small_ns, w, bias, aux, aux_ns, big, big_ns, N, A, Bc, s, s, epi, st)
        else:
            self.lib.up1d(small, small_ns, w, bias, aux, aux_ns, big, big_ns, N, A, Bc,
                          INPUT_ELEMS >> (2 * lvl), pad, epi, st)

    def _wgrad(self, small, small_ns, big, big_ns, dw, N, A, Bc, lvl, pad, st):
        if self.ndim == 2:
            s = 128 >> lvl
            self.lib.wgrad2d(small, small_ns, big, big_ns, dw, N, A, Bc, s, s, st)
        else:
            self.lib.wgrad1d(small, small_ns, big, big_ns, dw, N, A, Bc, INPUT_ELEMS >> (2 * lvl), pad, st)

    # ------------------------------------------------------------------ forward
    def encode(self, x, uvh, p, ws: Workspace, st: int, out=None, x_planes=None):
        """ELU(fc1(cat(flatten(conv stack(x)), ELU(fcuv1(uvh))))) -> `out` (default ws.mu0).

        src/lofar_models.py:71-84 / :156-169.  uvh is the [N,4H] harmonic vector.
        x_planes: the input as operand planes (then `x` is not read: the first conv fetches its tiles by TMA).
        """
        lb, N, L, H4, ch, sz = self.lib, ws.N, self.L, self.H4, self.ch, ws.sizes
        ld1 = FLAT + H4
        src, src_ns = x, sz[0]
        for i in range(6):  # conv_i + ELU ; Conv1d uses pad=1 (src/lofar_models.py:115)
            if i < 5:
                dst, dst_ns = ws.enc[i + 1], sz[i + 1]
            else:
                dst, dst_ns = ws.cat1, ld1
            if i == 0 and x_planes is not None:
                self._down_planes(_p(x_planes), _p(self.img[("conv0.weight", 0)]), _p(p["conv0.bias"]), None, 0,
                                  _p(dst), dst_ns, N, ch[1], ch[0], 1, EPI_ELU, st)
            else:
                self._down(_p(src), src_ns, _p(self.img[(f"conv{i}.weight", 0)]), _p(p[f"conv{i}.bias"]), None, 0,
                           _p(dst), dst_ns, N, ch[i + 1], ch[i], i + 1, 1, EPI_ELU, st)
            src, src_ns = dst, dst_ns
        lb.linear_fwd(_p(uvh), uvh.stride(0), _p(p["fcuv1.weight"]), _p(p["fcuv1.bias"]),
                      ws.cat1.data_ptr() + 4 * FLAT, ld1, N, H4, H4, EPI_ELU, st)
        out = ws.mu0 if out is None else out
        lb.linear_fwd(_p(ws.cat1), ld1, _p(p["fc1.weight"]), _p(p["fc1.bias"]), _p(out), out.stride(0), N, ld1, L, EPI_ELU, st)
        return out

    def decode(self, uvh, p, ws: Workspace, st: int):
        """Decoder on the latent already stored in ws.zcat[:, :L]; src/lofar_models.py:86-99 /
        :171-184 (ConvTranspose1d uses pad=0)."""
        lb, N, L, H4, ch, sz = self.lib, ws.N, self.L, self.H4, self.ch, ws.sizes
        zc_ld = L + H4
        lb.linear_fwd(_p(uvh), uvh.stride(0), _p(p["fcuv3.weight"]), _p(p["fcuv3.bias"]),
                      ws.zcat.data_ptr() + 4 * L, zc_ld, N, H4, H4, EPI_ELU, st)
        lb.linear_fwd(_p(ws.zcat), zc_ld, _p(p["fc3.weight"]), _p(p["fc3.bias"]), _p(ws.dec[0]), FLAT, N, zc_ld, FLAT, EPI_NONE, st)
        rch = ch[::-1]
        src = ws.dec[0]
        for i in range(6):
            dst = ws.dec[i + 1] if i < 5 else ws.xhat
            self._up(_p(src), sz[6 - i], _p(self.img[(f"tconv{i}.weight", 1)]), _p(p[f"tconv{i}.bias"]), None, 0,
                     _p(dst), sz[5 - i], N, rch[i], rch[i + 1], 6 - i, 0, EPI_ELU if i < 5 else EPI_NONE, st)
            src = dst
        return ws.xhat

    def forward(self, x: torch.Tensor, uv: torch.Tensor, scales: torch.Tensor,
                p: Dict[str, torch.Tensor], ws: Workspace, st: int,
                mu_out: Optional[torch.Tensor] = None, decode: bool = True, x_planes: Optional[torch.Tensor] = None,
                prepare: bool = True):
        """Runs the network; returns (xhat [N,C*16384] in ws, mu view); decode=False stops at the
        latent (the 1-D nets of the clustering path, src/evaluate_clustering.py:86-89) and returns (None, mu).

        mu_out: optional [N,L] view (any row stride) that receives the returned latent
        (lets the three nets write straight into the concatenated Mu buffer).
        """
        lb, N, L, H4 = self.lib, ws.N, self.L, self.H4
        if prepare:     # (False: the caller made the weight images once for several micro-batches)
            self.prepare_images(p, st, hasattr(ws, "g_enc"))
        lb.uv_harmonics(_p(uv), _p(scales), N, scales.numel(), _p(ws.uvh), st)
        mu_final = mu_out if mu_out is not None else ws.mu
        mu_ld, zc_ld = mu_final.stride(0), L + H4
        if self.rica:
            self.encode(x, ws.uvh, p, ws, st, x_planes=x_planes)
            lb.linear_fwd(_p(ws.mu0), L, _p(p["fc2in.weight"]), _p(p["fc2in.bias"]), _p(mu_final), mu_ld, N, L, L, EPI_ELU, st)
            lb.linear_fwd(_p(mu_final), mu_ld, _p(p["fc2out.weight"]), _p(p["fc2out.bias"]), _p(ws.zcat), zc_ld, N, L, L, EPI_ELU, st)
        else:
            self.encode(x, ws.uvh, p, ws, st, out=ws.zcat[:, :L], x_planes=x_planes)
            mu_final.copy_(ws.zcat[:, :L])
        if not decode:
            return None, mu_final
        return self.decode(ws.uvh, p, ws, st), mu_final

    # ------------------------------------------------------------------ backward
    def backward(self, x: torch.Tensor, p: Dict[str, torch.Tensor], g: Dict[str, torch.Tensor],
                 ws: Workspace, st: int, g_xhat: Optional[torch.Tensor],
                 g_mu: Optional[torch.Tensor], mu: torch.Tensor, need_dx: bool,
                 wstream: Optional[torch.cuda.Stream] = None, out_bias_done: bool = False,
                 x_planes: Optional[torch.Tensor] = None, g_xhat_planes: Optional[torch.Tensor] = None):
        """Writes every parameter gradient into g[name] (overwrite) and returns dx or None.

        wstream: optional second stream for the leaf work of the conv layers (weight and bias gradients).
        out_bias_done: g["tconv5.bias"] (the per-channel sum of g_xhat) was already written by the kernel
        that produced g_xhat (lshm_cascade_losses / lshm_cascade_combine).
        The data-gradient chain stays on `st` (the current stream); each layer's wgrad / bias sum is forked
        to `wstream` once its output gradient exists and everything is joined before returning, so the
        latency-bound deep layers overlap instead of queueing behind each other.

        x_planes / g_xhat_planes: the network input / the reconstruction gradient as operand planes (then the fp32
        tensors `x` / `g_xhat` are not read; g_xhat_planes needs out_bias_done, the bias sum has no fp32 source);
        g_xhat: gradient w.r.t. the reconstruction ([N, C*16384] contiguous) or None;
        g_mu:   gradient w.r.t. the returned latent ([N,L] view, any row stride) or None;
        mu:     the latent view returned by forward (post-fc2in activation when rica).
        """
        lb, N, L, H4, ch, sz = self.lib, ws.N, self.L, self.H4, self.ch, ws.sizes
        ld1, zc_ld = FLAT + H4, L + H4
        rch = ch[::-1]
        main = torch.cuda.current_stream(x.device)
        wst = st if wstream is None else wstream.cuda_stream

        def fork():      # leaf work may start once everything queued so far on the main stream is done
            if wstream is not None:
                ev = torch.cuda.Event()
                ev.record(main)
                wstream.wait_event(ev)
        if g_xhat_planes is not None and not out_bias_done:
            raise RuntimeError("lshm_b200: g_xhat_planes needs the output-bias gradient from the producing kernel")
        if g_xhat is None and g_xhat_planes is None:
            # no reconstruction gradient: decoder parameters get zero gradient
            for i in range(6):
                g[f"tconv{i}.weight"].zero_(); g[f"tconv{i}.bias"].zero_()
            for nm in ("fc3.weight", "fc3.bias", "fcuv3.weight", "fcuv3.bias"):
                g[nm].zero_()
            if self.rica:
                g["fc2out.weight"].zero_(); g["fc2out.bias"].zero_()
            ws.g_zcat.zero_()
        else:
            # decoder: dz_i = gradient w.r.t. tconv_i pre-activation (big map of level 5-i)
            dz = g_xhat
            for i in range(5, -1, -1):
                inp = ws.dec[i]                     # input of tconv_i (small map, level 6-i)
                A, Bc, lvl = rch[i], rch[i + 1], 6 - i
                fork()
                nxt = ws.g_dec[i]
                if i == 5 and g_xhat_planes is not None:
                    if self.ndim == 1 and A <= 16 and Bc <= 8 and FUSE_LAST:
                        # both gradients of the last transposed conv from ONE read of the reconstruction gradient
                        lb.tconv_bwd1d_planes(_p(inp), sz[lvl], _p(g_xhat_planes), _p(self.img[("tconv5.weight", 0)]),
                                              _p(nxt), sz[lvl], _p(g["tconv5.weight"]), N, A, Bc, INPUT_ELEMS >> (2 * lvl), st)
                    elif self.ndim == 2 and A <= 8 and Bc <= 8 and FUSE_LAST:
                        s2 = 128 >> lvl
                        lb.tconv_bwd2d_planes(_p(inp), sz[lvl], _p(g_xhat_planes), _p(self.img[("tconv5.weight", 0)]),
                                              _p(nxt), sz[lvl], _p(g["tconv5.weight"]), N, A, Bc, s2, s2, st)
                    else:
                        self._wgrad_planes(_p(inp), sz[lvl], _p(g_xhat_planes), _p(g["tconv5.weight"]), N, A, Bc, lvl, wst)
                        self._down_planes(_p(g_xhat_planes), _p(self.img[("tconv5.weight", 0)]), None, _p(inp), sz[lvl],
                                          _p(nxt), sz[lvl], N, A, Bc, lvl, EPI_DELU, st)
                    dz = nxt
                    continue
                fused = FUSE_LAST and i > 0 and 8 < A <= 16 and Bc == 8     # the 12 -> 8 channel layer
                if not fused:
                    self._wgrad(_p(inp), sz[lvl], _p(dz), sz[lvl - 1], _p(g[f"tconv{i}.weight"]), N, A, Bc, lvl, 0, wst)
                if not (i == 5 and out_bias_done):
                    lb.channel_sum(_p(dz), sz[lvl - 1], _p(g[f"tconv{i}.bias"]), N, Bc, sz[lvl - 1] // Bc, wst)
                if fused:
                    # weight and data gradient from one gather of the output gradient (lshm_tconv_bwd*)
                    if self.ndim == 2:
                        s2 = 128 >> lvl
                        lb.tconv_bwd2d(_p(inp), sz[lvl], _p(dz), sz[lvl - 1], _p(self.img[(f"tconv{i}.weight", 0)]),
                                       _p(nxt), sz[lvl], _p(g[f"tconv{i}.weight"]), N, A, Bc, s2, s2, st)
                    else:
                        lb.tconv_bwd1d(_p(inp), sz[lvl], _p(dz), sz[lvl - 1], _p(self.img[(f"tconv{i}.weight", 0)]),
                                       _p(nxt), sz[lvl], _p(g[f"tconv{i}.weight"]), N, A, Bc, INPUT_ELEMS >> (2 * lvl), st)
                    dz = nxt
                    continue
                # dgrad of the transposed conv = "down"; ELU' of the producing layer unless it is fc3
                self._down(_p(dz), sz[lvl - 1], _p(self.img[(f"tconv{i}.weight", 0)]), None,
                           _p(inp) if i > 0 else None, sz[lvl], _p(nxt), sz[lvl], N, A, Bc, lvl, 0,
                           EPI_DELU if i > 0 else EPI_NONE, st)
                dz = nxt
            # fc3 (no activation) and the decoder uv branch
            lb.linear_bwd_weight(_p(ws.zcat), zc_ld, _p(dz), FLAT, _p(g["fc3.weight"]), _p(g["fc3.bias"]), N, zc_ld, FLAT, st)
            lb.linear_bwd_data(_p(dz), FLAT, _p(p["fc3.weight"]), None, 0, _p(ws.zcat), zc_ld, _p(ws.g_zcat), zc_ld, N, zc_ld, FLAT, st)
            lb.linear_bwd_weight(_p(ws.uvh), H4, ws.g_zcat.data_ptr() + 4 * L, zc_ld, _p(g["fcuv3.weight"]), _p(g["fcuv3.bias"]), N, H4, H4, st)
        mu_ld = mu.stride(0)
        gmu_p, gmu_ld = (_p(g_mu), g_mu.stride(0)) if g_mu is not None else (None, 0)
        if self.rica:
            if g_xhat is not None or g_xhat_planes is not None:
                lb.linear_bwd_weight(_p(mu), mu_ld, _p(ws.g_zcat), zc_ld, _p(g["fc2out.weight"]), _p(g["fc2out.bias"]), N, L, L, st)
            # dz(fc2in) = (dz(fc2out) W2out + g_mu) * ELU'(mu)
            lb.linear_bwd_data(_p(ws.g_zcat), zc_ld, _p(p["fc2out.weight"]), gmu_p, gmu_ld, _p(mu), mu_ld, _p(ws.g_mu), L, N, L, L, st)
            lb.linear_bwd_weight(_p(ws.mu0), L, _p(ws.g_mu), L, _p(g["fc2in.weight"]), _p(g["fc2in.bias"]), N, L, L, st)
            lb.linear_bwd_data(_p(ws.g_mu), L, _p(p["fc2in.weight"]), None, 0, _p(ws.mu0), L, _p(ws.g_mu0), L, N, L, L, st)
            cat_in = ws.g_mu0
        else:
            tot = ws.g_zcat[:, :L] if g_mu is None else ws.g_zcat[:, :L] + g_mu
            tot = tot.contiguous()
            lb.delu(_p(tot), L, _p(ws.zcat), zc_ld, _p(ws.g_mu0), L, N, L, st)
            cat_in = ws.g_mu0
        lb.linear_bwd_weight(_p(ws.cat1), ld1, _p(cat_in), L, _p(g["fc1.weight"]), _p(g["fc1.bias"]), N, ld1, L, st)
        lb.linear_bwd_data(_p(cat_in), L, _p(p["fc1.weight"]), None, 0, _p(ws.cat1), ld1, _p(ws.g_cat1), ld1, N, ld1, L, st)
        lb.linear_bwd_weight(_p(ws.uvh), H4, ws.g_cat1.data_ptr() + 4 * FLAT, ld1, _p(g["fcuv1.weight"]), _p(g["fcuv1.bias"]), N, H4, H4, st)
        # encoder: dz_i = gradient w.r.t. conv_i pre-activation (small map of level i+1)
        dz, dz_ns = ws.g_cat1, ld1
        for i in range(5, -1, -1):
            inp, inp_ns = (x, sz[0]) if i == 0 else (ws.enc[i], sz[i])
            A, Bc, lvl = ch[i + 1], ch[i], i + 1
            fork()
            if i == 0 and x_planes is not None:
                self._wgrad_planes(_p(dz), dz_ns, _p(x_planes), _p(g["conv0.weight"]), N, A, Bc, lvl, wst)
            else:
                self._wgrad(_p(dz), dz_ns, _p(inp), inp_ns, _p(g[f"conv{i}.weight"]), N, A, Bc, lvl, 1, wst)
            lb.channel_sum(_p(dz), dz_ns, _p(g[f"conv{i}.bias"]), N, A, sz[lvl] // A, wst)
            if i > 0:
                nxt = ws.g_enc[i]
                self._up(_p(dz), dz_ns, _p(self.img[(f"conv{i}.weight", 1)]), None, _p(inp), inp_ns, _p(nxt), sz[i], N, A, Bc, lvl, 1, EPI_DELU, st)
                dz, dz_ns = nxt, sz[i]
            elif need_dx:
                if ("conv0.weight", 1) not in self.img:   # 2-D net asked for dx (not on the training path)
                    self.img[("conv0.weight", 1)] = conv_image(p["conv0.weight"], self.ndim, 1, st)
                elif not self.need_input_grad:
                    conv_image(p["conv0.weight"], self.ndim, 1, st, self.img[("conv0.weight", 1)])
                self._up(_p(dz), dz_ns, _p(self.img[("conv0.weight", 1)]), None, None, 0, _p(ws.dx), sz[0], N, A, Bc, lvl, 1, EPI_NONE, st)
        if wstream is not None:
            ev = torch.cuda.Event()
            ev.record(wstream)
            main.wait_event(ev)
        return ws.dx if need_dx else None

    # ------------------------------------------------------------------ encode() / decode() on their own
    def backward_encode(self, x: torch.Tensor, uvh: torch.Tensor, p, g, ws: Workspace, st: int, g_out: torch.Tensor,
                        need_dx: bool):
        """Backward of `encode(x, uvh)` alone (src/lofar_models.py:71-84): g_out = gradient w.r.t. the returned
        ELU(fc1(.)) [N,L].  Writes g[...] for conv0..5, fcuv1, fc1 and returns (dx or None, g_uvh [N,4H])."""
        lb, N, L, H4, ch, sz = self.lib, ws.N, self.L, self.H4, self.ch, ws.sizes
        ld1 = FLAT + H4
        lb.delu(_p(g_out), g_out.stride(0), _p(ws.mu0), L, _p(ws.g_mu0), L, N, L, st)
        lb.linear_bwd_weight(_p(ws.cat1), ld1, _p(ws.g_mu0), L, _p(g["fc1.weight"]), _p(g["fc1.bias"]), N, ld1, L, st)
        lb.linear_bwd_data(_p(ws.g_mu0), L, _p(p["fc1.weight"]), None, 0, _p(ws.cat1), ld1, _p(ws.g_cat1), ld1, N, ld1, L, st)
        lb.linear_bwd_weight(_p(uvh), uvh.stride(0), ws.g_cat1.data_ptr() + 4 * FLAT, ld1, _p(g["fcuv1.weight"]), _p(g["fcuv1.bias"]), N, H4, H4, st)
        g_uvh = torch.empty(N, H4, dtype=torch.float32, device=x.device)
        lb.linear_bwd_data(ws.g_cat1.data_ptr() + 4 * FLAT, ld1, _p(p["fcuv1.weight"]), None, 0, None, 0, _p(g_uvh), H4, N, H4, H4, st)
        dz, dz_ns = ws.g_cat1, ld1
        for i in range(5, -1, -1):
            inp, inp_ns = (x, sz[0]) if i == 0 else (ws.enc[i], sz[i])
            A, Bc, lvl = ch[i + 1], ch[i], i + 1
            self._wgrad(_p(dz), dz_ns, _p(inp), inp_ns, _p(g[f"conv{i}.weight"]), N, A, Bc, lvl, 1, st)
            lb.channel_sum(_p(dz), dz_ns, _p(g[f"conv{i}.bias"]), N, A, sz[lvl] // A, st)
            if i > 0:
                nxt = ws.g_enc[i]
                self._up(_p(dz), dz_ns, _p(self.img[(f"conv{i}.weight", 1)]), None, _p(inp), inp_ns, _p(nxt), sz[i], N, A, Bc, lvl, 1, EPI_DELU, st)
                dz, dz_ns = nxt, sz[i]
            elif need_dx:
                if ("conv0.weight", 1) not in self.img:
                    self.img[("conv0.weight", 1)] = conv_image(p["conv0.weight"], self.ndim, 1, st)
                elif not self.need_input_grad:
                    conv_image(p["conv0.weight"], self.ndim, 1, st, self.img[("conv0.weight", 1)])
                self._up(_p(dz), dz_ns, _p(self.img[("conv0.weight", 1)]), None, None, 0, _p(ws.dx), sz[0], N, A, Bc, lvl, 1, EPI_NONE, st)
        return (ws.dx if need_dx else None), g_uvh

    def backward_decode(self, uvh: torch.Tensor, p, g, ws: Workspace, st: int, g_xhat: torch.Tensor):
        """Backward of `decode(z, uvh)` alone (src/lofar_models.py:86-99): writes g[...] for tconv0..5, fc3, fcuv3 and
        returns (g_z [N,L], g_uvh [N,4H]).  z is an input here: no ELU' on its gradient."""
        lb, N, L, H4, ch, sz = self.lib, ws.N, self.L, self.H4, self.ch, ws.sizes
        zc_ld = L + H4
        rch = ch[::-1]
        dz = g_xhat
        for i in range(5, -1, -1):
            inp = ws.dec[i]
            A, Bc, lvl = rch[i], rch[i + 1], 6 - i
            self._wgrad(_p(inp), sz[lvl], _p(dz), sz[lvl - 1], _p(g[f"tconv{i}.weight"]), N, A, Bc, lvl, 0, st)
            lb.channel_sum(_p(dz), sz[lvl - 1], _p(g[f"tconv{i}.bias"]), N, Bc, sz[lvl - 1] // Bc, st)
            nxt = ws.g_dec[i]
            self._down(_p(dz), sz[lvl - 1], _p(self.img[(f"tconv{i}.weight", 0)]), None,
                       _p(inp) if i > 0 else None, sz[lvl], _p(nxt), sz[lvl], N, A, Bc, lvl, 0,
                       EPI_DELU if i > 0 else EPI_NONE, st)
            dz = nxt
        lb.linear_bwd_weight(_p(ws.zcat), zc_ld, _p(dz), FLAT, _p(g["fc3.weight"]), _p(g["fc3.bias"]), N, zc_ld, FLAT, st)
        lb.linear_bwd_data(_p(dz), FLAT, _p(p["fc3.weight"]), None, 0, None, 0, _p(ws.g_zcat), zc_ld, N, zc_ld, FLAT, st)
        g_z = ws.g_zcat[:, :L].clone()
        guv = torch.empty(N, H4, dtype=torch.float32, device=uvh.device)
        lb.delu(ws.g_zcat.data_ptr() + 4 * L, zc_ld, ws.zcat.data_ptr() + 4 * L, zc_ld, _p(guv), H4, N, H4, st)   # through ELU(fcuv3(uvh))
        lb.linear_bwd_weight(_p(uvh), uvh.stride(0), _p(guv), H4, _p(g["fcuv3.weight"]), _p(g["fcuv3.bias"]), N, H4, H4, st)
        g_uvh = torch.empty(N, H4, dtype=torch.float32, device=uvh.device)
        lb.linear_bwd_data(_p(guv), H4, _p(p["fcuv3.weight"]), None, 0, None, 0, _p(g_uvh), H4, N, H4, H4, st)
        return g_z, g_uvh

