"""Deep K-harmonic training step: the closure of /root/reference/src/kharmonic_lofar.py:132-182,
the multiplier update of :187-202 and the loop of :115-208, on liblshm_sm100 kernels.

Two ways to use it:

* the drop-in modules of :mod:`lshm_b200.lofar_models` work in the reference loop *unchanged*
  (autograd sees one node per module);
* :class:`DeepKHarmonicStep` is the fused path used by ``bench.py``: one flat parameter buffer
  and one flat gradient buffer for ``net``/``netT``/``netF``/``mod``, the whole closure as a
  fixed sequence of kernel launches with analytic gradients (no autograd graph, no host sync),
  and - under data parallelism - ONE all-reduce per closure evaluation over
  ``[all gradients | 16 loss scalars]``.

The closure honours the ``lbfgsnew.LBFGSNew`` contract (src/lbfgsnew.py:498-759): with grad
enabled it leaves ``.grad`` on every leaf Parameter; under ``torch.set_grad_enabled(False)``
(:686-693) it is forward-only; it returns a 0-dim tensor whose ``float()`` is the loss and is
identical on every rank.
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.distributed as dist

from ._lib import lib
from .lofar_models import AutoEncoder1DCNN, AutoEncoderCNN2, Kmeans
from .parallel import ShardPlan, exchange

LOSS_TAIL = 16  # floats appended to the flat gradient buffer: total + the 8 printed terms
TERM_NAMES = ("loss0", "loss1", "loss2", "loss3", "kdist", "aug", "sim", "rica")


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


class FlatParams:
    """Re-homes every Parameter of `modules` into one flat fp32 buffer (and `.grad` into a
    second one).  Parameters stay ordinary dense leaf tensors (views), so ``torch.optim.Adam``,
    ``LBFGSNew`` (``p.data.add_``, ``p.copy_``, ``p.grad.data`` - src/lbfgsnew.py:84-112) and
    ``state_dict`` keep working; the flat layout makes the data-parallel exchange one call."""

    ALIGN = 64  # floats: every tensor starts 256-byte aligned

    def __init__(self, modules, device):
        self.params: List[torch.nn.Parameter] = []
        self.names: List[str] = []
        for mi, m in enumerate(modules):
            for nm, p in m.named_parameters():
                self.params.append(p)
                self.names.append(f"{mi}.{nm}")
        offs, off = [], 0
        for p in self.params:
            offs.append(off)
            off += (p.numel() + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        self.offsets, self.numel = offs, off
        self.flat = torch.zeros(off, dtype=torch.float32, device=device)
        self.grad = torch.zeros(off + LOSS_TAIL, dtype=torch.float32, device=device)
        self.grad_views = []
        with torch.no_grad():
            for p, o in zip(self.params, offs):
                v = self.flat[o:o + p.numel()].view(p.shape)
                v.copy_(p.data)
                p.data = v
                gv = self.grad[o:o + p.numel()].view(p.shape)
                p.grad = gv
                self.grad_views.append(gv)
        self.loss_tail = self.grad[off:off + LOSS_TAIL]

    def attach_grads(self):
        for p, gv in zip(self.params, self.grad_views):
            if p.grad is not gv:
                p.grad = gv


class FlatAdam:
    """torch.optim.Adam semantics (src/kharmonic_lofar.py:92) as ONE kernel over the flat
    buffer.  The step count lives in device memory (lshm_adam_step_dev), so a step can be
    captured in a CUDA graph (GraphedStep)."""

    def __init__(self, flat: FlatParams, lr=1e-4, betas=(0.9, 0.999), eps=1e-8):
        self.flat, self.lr, self.betas, self.eps = flat, lr, betas, eps
        self.m = torch.zeros_like(flat.flat)
        self.v = torch.zeros_like(flat.flat)
        self.t_dev = torch.zeros(1, dtype=torch.int32, device=flat.flat.device)

    @property
    def t(self) -> int:
        return int(self.t_dev.item())

    def zero_grad(self):
        pass  # the fused closure overwrites every gradient

    def step(self, closure):
        with torch.enable_grad():
            loss = closure()
        self.apply()
        return loss

    def apply(self):
        """The parameter update alone, from the gradients already in the flat buffer."""
        lib().adam_step_dev(self.flat.flat.data_ptr(), self.flat.grad.data_ptr(), self.m.data_ptr(),
                            self.v.data_ptr(), self.flat.numel, self.lr, self.betas[0], self.betas[1], self.eps,
                            self.t_dev.data_ptr(), _stream())


class GraphedStep:
    """`optimizer.step(step.closure); step.update_multipliers()` captured once as a CUDA graph and
    replayed: ~270 kernel launches become one, which removes the launch gaps between the small kernels
    of the deep layers and the host's launch lead after every loss read-back.

    The minibatch lives in the static tensors `.x` [N,C,128,128] and `.uv` [N,2] (the ones given to
    `step.set_batch` before construction): write the next minibatch into them (`load(x, uv)` copies, or
    let `lofar_tools.patchify_device(..., out=graphed.x)` produce it in place), call `new_batch()` when
    the multipliers must restart, then `replay()`.  Needs a FlatAdam optimiser (device-side step count).
    """

    def __init__(self, step: "DeepKHarmonicStep", optimizer: FlatAdam, warmup: int = 2):
        if not isinstance(optimizer, FlatAdam):
            raise RuntimeError("lshm_b200: GraphedStep needs a FlatAdam optimiser (device-side step count)")
        if step.N == 0:
            raise RuntimeError("lshm_b200: call step.set_batch(...) before capturing the step")
        self.step, self.opt = step, optimizer
        self.x, self.uv = step.x, step.uv
        dev = step.device
        # warm-up and capture must not change the training state: snapshot, run, restore
        keep = [t.clone() for t in (step.flat.flat, optimizer.m, optimizer.v, optimizer.t_dev, step.y1, step.y2, step.y3)]
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        # (thread-local capture mode: the NCCL watchdog thread polls events while we capture.)
        # Data parallel: the NCCL all-reduce stays outside (collectives inside a capture are fragile), so the
        # step is two graphs with the eager exchange between them: [closure] -> all-reduce -> [Adam, multipliers]
        self.graph = torch.cuda.CUDAGraph()
        self.graph2 = torch.cuda.CUDAGraph() if step.distributed else None
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                self._one()
            side.synchronize()
            l0 = lib().launches
            if self.graph2 is None:
                with torch.cuda.graph(self.graph, stream=side, capture_error_mode="thread_local"):
                    self.loss = self._one()
            else:
                step._defer_exchange = True
                try:
                    with torch.cuda.graph(self.graph, stream=side, capture_error_mode="thread_local"):
                        with torch.enable_grad():
                            self.loss = step.closure()
                    with torch.cuda.graph(self.graph2, stream=side, pool=self.graph.pool(), capture_error_mode="thread_local"):
                        optimizer.apply()
                        step.update_multipliers()
                finally:
                    step._defer_exchange = False
            self.launches_per_replay = lib().launches - l0
        torch.cuda.current_stream(dev).wait_stream(side)
        for dst, src in zip((step.flat.flat, optimizer.m, optimizer.v, optimizer.t_dev, step.y1, step.y2, step.y3), keep):
            dst.copy_(src)

    def _one(self):
        loss = self.opt.step(self.step.closure)
        self.step.update_multipliers()
        return loss

    def load(self, x: torch.Tensor, uv: torch.Tensor, reset_multipliers: bool = True):
        self.x.copy_(x.view_as(self.x), non_blocking=True)
        self.uv.copy_(uv.view_as(self.uv), non_blocking=True)
        if reset_multipliers:
            self.new_batch()

    def new_batch(self):
        """src/kharmonic_lofar.py:128-130: the multipliers restart with every minibatch."""
        self.step.y1.zero_(); self.step.y2.zero_(); self.step.y3.zero_()

    def replay(self) -> torch.Tensor:
        """One optimiser step + multiplier update; returns the (static) total-loss tensor."""
        self.graph.replay()
        if self.graph2 is not None:
            exchange(self.step.flat.grad, self.step.group)
            self.graph2.replay()
        lib().launches += self.launches_per_replay     # kernels launched by the replay
        return self.loss


class DeepKHarmonicStep:
    """Fused closure + multiplier update for one minibatch (see module docstring).

    Hyper-parameter names and defaults follow src/kharmonic_lofar.py:37-48.
    `group` is an optional torch.distributed process group: each rank then holds a shard of
    whole baseline groups (rows [g*bpb,(g+1)*bpb)) and `global_patches` is the global N.
    """

    def __init__(self, net: AutoEncoderCNN2, netT: AutoEncoder1DCNN, netF: AutoEncoder1DCNN, mod: Kmeans, *,
                 alpha=0.01, beta=0.01, gamma=0.01, rho=1.0, use_rica=True, rica_lambda=0.01,
                 group: Optional[dist.ProcessGroup] = None, distributed: bool = False):
        self.net, self.netT, self.netF, self.mod = net, netT, netF, mod
        self.alpha, self.beta, self.gamma, self.rho = alpha, beta, gamma, rho
        self.use_rica, self.rica_lambda = use_rica, rica_lambda
        self.distributed, self.group = distributed, group
        self.world = dist.get_world_size(group) if distributed else 1
        dev = next(net.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("lshm_b200: DeepKHarmonicStep needs CUDA modules (no CPU path)")
        self.device = dev
        self.flat = FlatParams([net, netT, netF, mod], dev)
        self.L, self.Lt = net.latent_dim, netT.latent_dim
        self.Ltot = self.L + 2 * self.Lt
        if mod.latent_dim != self.Ltot:
            raise RuntimeError("lshm_b200: Kmeans.latent_dim must equal L + 2*Lt")
        self._pd = [m.named_param_dict() for m in (net, netT, netF)]
        self._gd = []
        views = dict(zip(self.flat.names, self.flat.grad_views))
        for mi, m in enumerate((net, netT, netF)):
            self._gd.append({nm: views[f"{mi}.{nm}"] for nm in m._names})
        self._gM = views["3.M"]
        self.N = 0
        self._defer_exchange = False     # GraphedStep runs the all-reduce itself, between its two graphs
        self._side = None                # second stream for the frequency-axis net
        self.overlap_streams = True      # False: everything on the current stream (per-kernel profiling)
        self._wst = [None, None, None]   # per-net streams for the weight / bias gradients
        self.launches = 0
        if distributed:
            self.broadcast_parameters()

    # ------------------------------------------------------------------ data parallel
    def broadcast_parameters(self, src: int = 0):
        dist.broadcast(self.flat.flat, src=src, group=self.group)

    # ------------------------------------------------------------------ batch
    def set_batch(self, x: torch.Tensor, uv: torch.Tensor, batch_per_bline: int,
                  global_patches: Optional[int] = None):
        """x [N,C,128,128], uv [N,2] (this rank's shard); resets y1..y3 (src/kharmonic_lofar.py:128-130)."""
        N, C = x.shape[0], x.shape[1]
        if N % batch_per_bline:
            raise RuntimeError("lshm_b200: shard must hold whole baseline groups")
        self.x = x.contiguous()
        self.uv = uv.contiguous()
        self.bpb = batch_per_bline
        self.Nglobal = int(global_patches) if global_patches is not None else N * self.world
        if N != self.N or getattr(self, "C", None) != C:
            dev, f = self.device, dict(device=self.device, dtype=torch.float32)
            self.N, self.C = N, C
            e = self.net.engine(), self.netT.engine(), self.netF.engine()
            self.ws = [e[0].workspace(N, dev, True, False), e[1].workspace(N, dev, True, True),
                       e[2].workspace(N, dev, True, True)]
            n = N * C * 16384
            self.iyT, self.iyF = torch.empty(n, **f), torch.empty(n, **f)
            self.g1p, self.g2, self.g3f, self.gx1 = (torch.empty(n, **f) for _ in range(4))
            self.y1, self.y2, self.y3 = (torch.empty(n, **f) for _ in range(3))
            self.Mu = torch.empty(N, self.Ltot, **f)
            self.gMu = torch.empty(N, self.Ltot, **f)
            self.terms = torch.zeros(16, dtype=torch.float64, device=dev)
            K = self.mod.K
            self.simwork = torch.empty(2 * K * K, **f)
            self.scales = self.net.harmonic_scales.to(dev).float().contiguous()
        self.y1.zero_(); self.y2.zero_(); self.y3.zero_()

    # ------------------------------------------------------------------ forward pieces
    def _forward(self, st):
        L, Lt, N, C = self.L, self.Lt, self.N, self.C
        e = self.net.engine(), self.netT.engine(), self.netF.engine()
        xf = self.x.view(N, -1)
        x1, _ = e[0].forward(xf, self.uv, self.scales, self._pd[0], self.ws[0], st, mu_out=self.Mu[:, :L])
        lib().residual_split(self.x.data_ptr(), x1.data_ptr(), self.iyT.data_ptr(), self.iyF.data_ptr(), N, C, 128, st)
        # The time-axis and frequency-axis nets are independent: they run on two streams (fork / join by
        # events, also inside a graph capture), so the latency-bound deep layers of one overlap the other's.
        side = self._fork()
        with torch.cuda.stream(side):
            x3f, _ = e[2].forward(self.iyF.view(N, -1), self.uv, self.scales, self._pd[2], self.ws[2],
                                  side.cuda_stream, mu_out=self.Mu[:, L + Lt:])
        x2, _ = e[1].forward(self.iyT.view(N, -1), self.uv, self.scales, self._pd[1], self.ws[1], st,
                             mu_out=self.Mu[:, L:L + Lt])
        self._join(side)
        return x1, x2, x3f

    def _wstream(self, i: int) -> Optional[torch.cuda.Stream]:
        """Stream for the weight / bias gradients of net i (leaf work beside the data-gradient chain)."""
        if not self.overlap_streams:
            return None
        if self._wst[i] is None:
            self._wst[i] = torch.cuda.Stream(self.device)
        return self._wst[i]

    def _fork(self) -> torch.cuda.Stream:
        """Side stream that starts after everything queued so far on the current stream."""
        if not self.overlap_streams:
            return torch.cuda.current_stream(self.device)
        if self._side is None:
            self._side = torch.cuda.Stream(self.device)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self._side.wait_event(ev)
        return self._side

    def _join(self, side: torch.cuda.Stream):
        if not self.overlap_streams:
            return
        ev = torch.cuda.Event()
        ev.record(side)
        torch.cuda.current_stream(self.device).wait_event(ev)

    def closure(self) -> torch.Tensor:
        """src/kharmonic_lofar.py:132-182.  Returns the total loss (0-dim device tensor)."""
        lb, st = lib(), _stream()
        grads = torch.is_grad_enabled()
        start = lb.launches
        N, C, L, Lt, Ltot, K = self.N, self.C, self.L, self.Lt, self.Ltot, self.mod.K
        M = self.mod.M
        plan = ShardPlan(N, self.Nglobal, self.world, self.bpb, C, K, Ltot)
        numel_g = plan.numel_global
        self.terms.zero_()
        tp = self.terms.data_ptr()
        x1, x2, x3f = self._forward(st)
        g1p, g2, g3f = (self.g1p.data_ptr(), self.g2.data_ptr(), self.g3f.data_ptr()) if grads else (None, None, None)
        # the bias gradients of the three last transposed convs (= channel sums of the reconstruction
        # gradients) come out of the kernels that write those gradients
        fuse_db = grads and C <= 64
        if grads:
            self.flat.attach_grads()
        db2, db3 = ((self._gd[1]["tconv5.bias"].data_ptr(), self._gd[2]["tconv5.bias"].data_ptr()) if fuse_db
                    else (None, None))
        # the latent-space terms (a dozen small, latency-bound launches) run beside the HBM-bound cascade
        # losses: forked here, joined before the backward passes
        lside = self._fork()
        lb.cascade_losses(self.x.data_ptr(), x1.data_ptr(), x2.data_ptr(), x3f.data_ptr(),
                          self.y1.data_ptr(), self.y2.data_ptr(), self.y3.data_ptr(), self.rho,
                          N, C, 128, 1.0 / numel_g, tp, g1p, g2, g3f, db2, db3, st)
        khm_scale = plan.khm_scale(self.alpha)
        p = float(self.mod.p)
        aug_scale = plan.aug_scale(self.gamma)
        sim_scale = plan.sim_scale(self.beta)        # M is replicated: count its penalty once
        rica_scale = plan.rica_scale(self.rica_lambda)  # the kernel divides by the LOCAL numel
        Mu, gMu = self.Mu, self.gMu
        with torch.cuda.stream(lside):
            sl = lside.cuda_stream
            if grads:
                self._gM.zero_()
                lb.khm_fwd_bwd(Mu.data_ptr(), Ltot, M.data_ptr(), N, K, Ltot, p, khm_scale, tp + 8 * 8,
                               gMu.data_ptr(), Ltot, 0, self._gM.data_ptr(), sl)
            else:
                lb.khm_fwd(Mu.data_ptr(), Ltot, M.data_ptr(), N, K, Ltot, p, tp + 8 * 8, None, sl)
            lb.similarity(M.data_ptr(), K, Ltot, sim_scale, tp + 9 * 8, self._gM.data_ptr() if grads else None,
                          self.simwork.data_ptr(), sl)
            lb.augment(Mu.data_ptr(), Ltot, N, Ltot, self.bpb, aug_scale, tp + 10 * 8,
                       gMu.data_ptr() if grads else None, Ltot, sl)
            if self.use_rica:
                for off, width in ((0, L), (L, Lt), (L + Lt, Lt)):
                    lb.logcosh(Mu.data_ptr() + 4 * off, Ltot, N, width, rica_scale, tp + 11 * 8,
                               gMu.data_ptr() + 4 * off if grads else None, Ltot, sl)
        self._join(lside)
        if grads:
            e = self.net.engine(), self.netT.engine(), self.netF.engine()
            side = self._fork()
            with torch.cuda.stream(side):
                dF = e[2].backward(self.iyF.view(N, -1), self._pd[2], self._gd[2], self.ws[2], side.cuda_stream,
                                   self.g3f.view(N, -1), gMu[:, L + Lt:], Mu[:, L + Lt:], True, self._wstream(2), fuse_db)
            dT = e[1].backward(self.iyT.view(N, -1), self._pd[1], self._gd[1], self.ws[1], st, self.g2.view(N, -1),
                               gMu[:, L:L + Lt], Mu[:, L:L + Lt], True, self._wstream(1), fuse_db)
            self._join(side)
            lb.cascade_combine(self.g1p.data_ptr(), dT.data_ptr(), dF.data_ptr(), self.gx1.data_ptr(), N, C, 128,
                               self._gd[0]["tconv5.bias"].data_ptr() if fuse_db else None, st)
            e[0].backward(self.x.view(N, -1), self._pd[0], self._gd[0], self.ws[0], st, self.gx1.view(N, -1),
                          gMu[:, :L], Mu[:, :L], False, self._wstream(0), fuse_db)
        tail = self.flat.loss_tail
        lb.closure_total(tp, self.rho, numel_g, khm_scale, tail.data_ptr(), st)
        if self.distributed and not self._defer_exchange:
            # ONE exchange per closure evaluation: gradients + loss scalars (forward-only: scalars)
            buf = self.flat.grad if grads else tail
            exchange(buf, self.group)
        self.launches = lb.launches - start
        return tail[0]

    def loss_terms(self) -> dict:
        """The columns printed at src/kharmonic_lofar.py:179 (one device->host copy)."""
        v = self.flat.loss_tail[:9].tolist()
        d = dict(total=v[0])
        d.update(zip(TERM_NAMES, v[1:]))
        return d

    def update_multipliers(self):
        """src/kharmonic_lofar.py:187-202: no-grad forward of the cascade, then y_i += rho*r_i."""
        st = _stream()
        with torch.no_grad():
            x1, x2, x3f = self._forward(st)
            lib().multiplier_update(self.x.data_ptr(), x1.data_ptr(), x2.data_ptr(), x3f.data_ptr(), self.rho,
                                    self.y1.data_ptr(), self.y2.data_ptr(), self.y3.data_ptr(),
                                    self.N, self.C, 128, st)

    def latents(self) -> torch.Tensor:
        return self.Mu


def train(step: DeepKHarmonicStep, optimizer, batches, Nadmm=10, log=None):
    """The loop of src/kharmonic_lofar.py:115-208 over an iterable of
    (patchx, patchy, x, uv) minibatches (e.g. lofar_tools.get_data_minibatch(..., uvdist=True))."""
    for i, (patchx, patchy, x, uv) in enumerate(batches):
        step.set_batch(x, uv, patchx * patchy)
        for admm in range(Nadmm):
            optimizer.step(step.closure)
            if log is not None:
                t = step.loss_terms()
                log("%d %d %f %f %f %f %f %f %f %f" % ((i, admm) + tuple(t[k] for k in TERM_NAMES)))
            step.update_multipliers()
