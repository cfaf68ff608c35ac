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

import contextlib
from typing import List, Optional, Tuple

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
    ``state_dict`` keep working; the flat layout makes the data-parallel exchange one call.

    `extra_tail` floats follow the 16 loss scalars in the gradient buffer (the K x L centre numerator and
    K denominator sums of Kmeans.offline_update ride in the same all-reduce).

    `version` / `tracked` / `owning()`: an optimiser that owns every parameter update (FlatAdam,
    lbfgsnew.LBFGSNew on a FlatParams) sets `tracked`, calls `bump()` after each change and evaluates its
    closures inside `with flat.owning():`; the step then knows when the activations it already holds belong
    to the current parameters and reuses them.  A closure called from anywhere else (a hand-written client,
    torch.optim.*, the unmodified reference LBFGSNew with its `p.data.add_`) is always computed in full."""

    ALIGN = 64  # floats: every tensor starts 256-byte aligned

    def __init__(self, modules, device, extra_tail: int = 0):
        self.params: List[torch.nn.Parameter] = []
        self.names: List[str] = []
        self.module_of: List[int] = []
        for mi, m in enumerate(modules):
            for nm, p in m.named_parameters():
                self.params.append(p)
                self.names.append(f"{mi}.{nm}")
                self.module_of.append(mi)
        offs, off = [], 0
        for p in self.params:
            offs.append(off)
            off += (p.numel() + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        self.offsets, self.numel = offs, off
        self.flat = torch.zeros(off, dtype=torch.float32, device=device)
        self.grad = torch.zeros(off + LOSS_TAIL + extra_tail, dtype=torch.float32, device=device)
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
        self.extra_tail = self.grad[off + LOSS_TAIL:]
        self.version = 0
        self.tracked = False
        self._owner_depth = 0

    @property
    def in_owner_step(self) -> bool:
        return self._owner_depth > 0

    @contextlib.contextmanager
    def owning(self):
        """Entered by the tracking optimiser around the closure calls of its step()."""
        self._owner_depth += 1
        try:
            yield self
        finally:
            self._owner_depth -= 1

    def bump(self):
        """The parameters changed (called by the optimiser that owns the updates)."""
        self.version += 1

    def attach_grads(self):
        for p, gv in zip(self.params, self.grad_views):
            if p.grad is not gv:
                p.grad = gv

    def range_of(self, modules) -> Tuple[int, int]:
        """[start, stop) of the flat buffer that holds the parameters of the given module indices
        (they must be adjacent in the construction order)."""
        ids = sorted(set(int(m) for m in modules))
        if ids != list(range(ids[0], ids[-1] + 1)):
            raise ValueError("FlatParams.range_of: module indices must be adjacent")
        sel = [i for i, m in enumerate(self.module_of) if m in ids]
        if not sel:
            raise ValueError("FlatParams.range_of: no parameters selected")
        last = sel[-1]
        stop = self.offsets[last + 1] if last + 1 < len(self.offsets) else self.numel
        return self.offsets[sel[0]], stop


class FlatAdam:
    """torch.optim.Adam semantics as ONE kernel over (a range of) the flat buffer.  The step count lives
    in device memory (lshm_adam_step_dev), so a step can be replayed from a CUDA graph.

    `modules`: indices into the step's module list (0 net, 1 netT, 2 netF, 3 mod) whose parameters are
    updated; None = all four.  The reference script as shipped optimises `net.parameters()` only
    (src/kharmonic_lofar.py:84-92, the other three `params.extend` lines are commented out):
    that is ``modules=(0,)``; BASELINE's configs train all four."""

    def __init__(self, flat: FlatParams, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, modules=None):
        self.flat, self.lr, self.betas, self.eps = flat, lr, betas, eps
        self.start, self.stop = (0, flat.numel) if modules is None else flat.range_of(modules)
        self.m = torch.zeros_like(flat.flat)
        self.v = torch.zeros_like(flat.flat)
        self.t_dev = torch.zeros(1, dtype=torch.int32, device=flat.flat.device)
        flat.tracked = True

    @property
    def t(self) -> int:
        return int(self.t_dev.item())

    def zero_grad(self):
        pass  # the fused closure overwrites every gradient

    def step(self, closure):
        with torch.enable_grad(), self.flat.owning():
            loss = closure()
        self.apply()
        return loss

    def apply(self):
        """The parameter update alone, from the gradients already in the flat buffer."""
        o = 4 * self.start
        lib().adam_step_dev(self.flat.flat.data_ptr() + o, self.flat.grad.data_ptr() + o, self.m.data_ptr() + o,
                            self.v.data_ptr() + o, self.stop - self.start, self.lr, self.betas[0], self.betas[1],
                            self.eps, self.t_dev.data_ptr(), _stream())
        self.flat.bump()


class GraphedStep:
    """`optimizer.step(step.closure); step.update_multipliers()` with every launch sequence replayed from CUDA
    graphs (`DeepKHarmonicStep.enable_graphs`): the ~270 kernel launches of an ADMM iteration become three graph
    launches, which removes the launch gaps between the small kernels of the deep layers.

    The minibatch lives in the static tensors `.x` [N,C,128,128] and `.uv` [N,2] (the ones given to
    `step.set_batch` before construction): write the next minibatch into them (`load(x, uv)` copies, or
    let `lofar_tools.patchify_device(..., out=graphed.x)` produce it in place), call `new_batch()` when
    the multipliers must restart, then `replay()`.  Any optimiser works (the graphs belong to the closure);
    with FlatAdam the parameter update is a fixed launch too.
    """

    def __init__(self, step: "DeepKHarmonicStep", optimizer, warmup: int = 0):
        if step.N == 0:
            raise RuntimeError("lshm_b200: call step.set_batch(...) before capturing the step")
        self.step, self.opt = step, optimizer
        self.x, self.uv = step.x, step.uv
        step.enable_graphs()

    @property
    def launches_per_replay(self) -> int:
        return self.step.launches_last_iteration

    def load(self, x: torch.Tensor, uv: torch.Tensor, reset_multipliers: bool = True):
        self.x.copy_(x.view_as(self.x), non_blocking=True)
        self.uv.copy_(uv.view_as(self.uv), non_blocking=True)
        self.step.invalidate()
        if reset_multipliers:
            self.new_batch()

    def new_batch(self):
        """src/kharmonic_lofar.py:128-130: the multipliers restart with every minibatch."""
        self.step.reset_multipliers()

    def replay(self) -> torch.Tensor:
        """One optimiser step + multiplier update; returns the total loss of the closure."""
        l0 = lib().launches
        loss = self.opt.step(self.step.closure)
        self.step.update_multipliers()
        self.step.launches_last_iteration = lib().launches - l0
        return loss


class _MicroBatch:
    """Rows [n0, n1) of the minibatch as one launch chain of DeepKHarmonicStep: views of the workspaces, its own operand
    planes, gradient views (chain 0 writes the flat gradient buffer, the others a buffer of their own) and streams."""

    def __init__(self, index: int, n0: int, n1: int):
        self.index, self.n0, self.n1, self.N = index, n0, n1, n1 - n0
        self.ws = self.gd = self.gbuf = None
        self.xp = self.gx1p = self.pT = self.pF = self.p2 = self.p3 = None
        self.stream = self.side = None
        self.wst = [None, None, None]


class DeepKHarmonicStep:
    """Fused closure + multiplier update for one minibatch (see module docstring).

    Hyper-parameter names and defaults follow src/kharmonic_lofar.py:37-48.
    `group` is an optional torch.distributed process group: each rank then holds a shard of
    whole baseline groups (rows [g*bpb,(g+1)*bpb)) and `global_patches` is the global N.

    Work that the reference repeats is done once (only when the optimiser tracks its updates, see
    FlatParams): the no-grad forward of the multiplier update (:187-198) IS the forward of the next
    closure on the same minibatch (same parameters, same x; only y1..y3 differ), so `update_multipliers`
    runs the forward, leaves the update y_i += rho r_i pending, and the next closure applies it inside the
    loss pass that reads x, x1, x2, x3 anyway (lshm_cascade_losses_upd) and goes straight to the backward.
    `y1`, `y2`, `y3` read through properties that apply a pending update first, so observable state always
    equals the reference's.  A closure evaluated again at unchanged parameters and multipliers (the f_old
    probe of the LBFGSNew line search, src/lbfgsnew.py:140) returns the loss already computed.

    `centre_sums=True` adds the K x L numerator / K denominator of Kmeans.offline_update
    (src/lofar_models.py:231-261) to every gradient closure; they travel in the same all-reduce as the
    gradients and `apply_centre_update()` sets M = num / den from the reduced sums.
    """

    def __init__(self, net: AutoEncoderCNN2, netT: AutoEncoder1DCNN, netF: AutoEncoder1DCNN, mod: Kmeans, *,
                 alpha=0.01, beta=0.01, gamma=0.01, rho=1.0, use_rica=True, rica_lambda=0.01,
                 group: Optional[dist.ProcessGroup] = None, distributed: bool = False, centre_sums: bool = False,
                 use_planes: bool = True, micro_batches: Optional[int] = None):
        self.net, self.netT, self.netF, self.mod = net, netT, netF, mod
        self.alpha, self.beta, self.gamma, self.rho = alpha, beta, gamma, rho
        self.use_rica, self.rica_lambda = use_rica, rica_lambda
        self.distributed, self.group = distributed, group
        self.world = dist.get_world_size(group) if distributed else 1
        dev = next(net.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("lshm_b200: DeepKHarmonicStep needs CUDA modules (no CPU path)")
        self.device = dev
        self.L, self.Lt = net.latent_dim, netT.latent_dim
        self.Ltot = self.L + 2 * self.Lt
        if mod.latent_dim != self.Ltot:
            raise RuntimeError("lshm_b200: Kmeans.latent_dim must equal L + 2*Lt")
        self.centre_sums = bool(centre_sums)
        # input-sized tensors that feed the first conv layers (x, x11, x11^T, the gradient of x1) live as operand
        # planes (include/lshm.h): those kernels fetch their tiles by tensor-TMA instead of gathering fp32
        self.use_planes = bool(use_planes)
        K = mod.K
        self.flat = FlatParams([net, netT, netF, mod], dev, extra_tail=(K * self.Ltot + K) if centre_sums else 0)
        self._pd = [m.named_param_dict() for m in (net, netT, netF)]
        self._gd = []
        views = dict(zip(self.flat.names, self.flat.grad_views))
        for mi, m in enumerate((net, netT, netF)):
            self._gd.append({nm: views[f"{mi}.{nm}"] for nm in m._names})
        self._gM = views["3.M"]
        # Micro-batches: the minibatch is cut into H row ranges that run the cascade (forward, loss pass, backward) as
        # independent launch chains on their own streams; the latency-bound deep layers of one chain then overlap the
        # HBM-bound first layers of another.  Each chain writes its parameter gradients into its own flat buffer, the
        # buffers are summed (lshm_vec_add) before the loss scalars are totalled.  Default 1: on B200 at cfg2 the chains do not
        # overlap enough (persistent kernels that own every SM), 8.76 / 9.19 / 10.07 ms per step for H = 1 / 2 / 4.
        self.micro = micro_batches
        self.N = 0
        self._side = None                # stream of the latent-space terms
        self.overlap_streams = True      # False: everything on the current stream (per-kernel profiling)
        self._wst = [None, None, None]   # per-net streams for the weight / bias gradients
        self.launches = 0
        self.launches_last_iteration = 0
        self.reuse = True                # reuse activations / losses when the optimiser tracks its updates
        self._graphs_on = False
        self._graphs = {}                # launch-sequence key -> (CUDAGraph, launches)
        self._fwd_key = None             # (flat.version, batch id) the held activations belong to
        self._loss_key = None            # same for the loss scalars in the tail (+ multiplier state)
        self._batch_id = 0
        self._pending = False            # multiplier update deferred into the next loss pass
        self._y_zero = True              # y1..y3 are identically zero and not materialised (new minibatch)
        if distributed:
            self.broadcast_parameters()

    # ------------------------------------------------------------------ data parallel
    def broadcast_parameters(self, src: int = 0):
        dist.broadcast(self.flat.flat, src=src, group=self.group)
        self.invalidate()

    # ------------------------------------------------------------------ batch
    def set_batch(self, x: torch.Tensor, uv: torch.Tensor, batch_per_bline: int,
                  global_patches: Optional[int] = None):
        """x [N,C,128,128], uv [N,2] (this rank's shard); resets y1..y3 (src/kharmonic_lofar.py:128-130)."""
        N, C = x.shape[0], x.shape[1]
        if N % batch_per_bline:
            raise RuntimeError("lshm_b200: shard must hold whole baseline groups")
        same_shape = N == self.N and getattr(self, "C", None) == C
        if self._graphs_on and same_shape:
            # the captured graphs read the static buffers: copy instead of re-binding
            if x.data_ptr() != self.x.data_ptr():
                self.x.copy_(x.view_as(self.x), non_blocking=True)
            if uv.data_ptr() != self.uv.data_ptr():
                self.uv.copy_(uv.view_as(self.uv), non_blocking=True)
        else:
            self.x = x.contiguous()
            self.uv = uv.contiguous()
        if self._graphs_on and (not same_shape or batch_per_bline != self.bpb):
            self._graphs = {}
        self.bpb = batch_per_bline
        Ng = int(global_patches) if global_patches is not None else N * self.world
        if self._graphs_on and getattr(self, "Nglobal", Ng) != Ng:
            self._graphs = {}
        self.Nglobal = Ng
        if not same_shape:
            dev, f = self.device, dict(device=self.device, dtype=torch.float32)
            self.N, self.C = N, C
            e = self.net.engine(), self.netT.engine(), self.netF.engine()
            self.ws = [e[0].workspace(N, dev, True, False), e[1].workspace(N, dev, True, True),
                       e[2].workspace(N, dev, True, True)]
            n = N * C * 16384
            H = self.micro if self.micro is not None else 1
            if H < 1 or N % H:
                raise RuntimeError(f"lshm_b200: {N} patches do not split into {H} micro-batches")
            old = getattr(self, "mb", [])
            self.mb = []
            for h in range(H):
                m = _MicroBatch(h, h * (N // H), (h + 1) * (N // H))
                m.ws = self.ws if H == 1 else [w.rows(m.n0, m.n1) for w in self.ws]
                if h == 0:
                    m.gd = self._gd
                else:
                    m.gbuf = torch.zeros(self.flat.numel, **f)
                    views = {nm: m.gbuf[o:o + p_.numel()].view(p_.shape)
                             for nm, o, p_ in zip(self.flat.names, self.flat.offsets, self.flat.params)}
                    m.gd = [{nm: views[f"{mi}.{nm}"] for nm in mod_._names} for mi, mod_ in enumerate((self.net, self.netT, self.netF))]
                if h < len(old):                     # keep the streams (captured graphs were dropped above)
                    m.stream, m.side, m.wst = old[h].stream, old[h].side, old[h].wst
                if self.use_planes:
                    from .engine import planes_buffer
                    m.xp, m.gx1p = planes_buffer(2, m.N, C, 64, 64, dev), planes_buffer(2, m.N, C, 64, 64, dev)
                    m.pT, m.pF = planes_buffer(1, m.N, C, 1, 4096, dev), planes_buffer(1, m.N, C, 1, 4096, dev)
                    m.p2, m.p3 = planes_buffer(1, m.N, C, 1, 4096, dev), planes_buffer(1, m.N, C, 1, 4096, dev)
                self.mb.append(m)
            if self.use_planes:
                self.iyT = self.iyF = self.gx1 = self.g2 = self.g3f = None
            else:
                self.iyT, self.iyF, self.gx1 = torch.empty(n, **f), torch.empty(n, **f), torch.empty(n, **f)
                self.g2, self.g3f = torch.empty(n, **f), torch.empty(n, **f)
            self.g1p = torch.empty(n, **f)
            self._y = [torch.empty(n, **f) for _ in range(3)]
            self.Mu = torch.empty(N, self.Ltot, **f)
            self.gMu = torch.empty(N, self.Ltot, **f)
            self.terms = torch.zeros(16, dtype=torch.float64, device=dev)
            K = self.mod.K
            self.simwork = torch.empty(2 * K * K, **f)
            self.scales = self.net.harmonic_scales.to(dev).float().contiguous()
        self.invalidate()
        self.reset_multipliers()

    def reset_multipliers(self):
        """y1 = y2 = y3 = 0 (a new minibatch); a deferred update of the previous minibatch is dropped, as the
        reference drops the multipliers themselves (src/kharmonic_lofar.py:128-130)."""
        self._y_zero = True              # nothing is written: the next kernels treat the multipliers as zero
        self._pending = False
        self._loss_key = None

    def invalidate(self):
        """The minibatch or the parameters were changed behind the step's back: recompute everything."""
        self._batch_id += 1
        self._fwd_key = None
        self._loss_key = None
        self.stage_input()

    def stage_input(self):
        """x -> operand planes for the 2-D net's first conv (forward and weight gradient); once per minibatch."""
        if self.use_planes and self.N:
            for m in self.mb:
                lib().stage_planes2d(self.x[m.n0:m.n1].data_ptr(), self.C * 16384, m.xp.data_ptr(), m.N, self.C, 64, 64, _stream())

    # multipliers: reading them applies a deferred update first, so they always hold the reference's values
    @property
    def y1(self) -> torch.Tensor:
        self._materialise_multipliers()
        return self._y[0]

    @property
    def y2(self) -> torch.Tensor:
        self._materialise_multipliers()
        return self._y[1]

    @property
    def y3(self) -> torch.Tensor:
        self._materialise_multipliers()
        return self._y[2]

    def _materialise_multipliers(self):
        self.flush_multipliers()
        if self._y_zero:
            for y in self._y:
                y.zero_()
            self._y_zero = False
        self._loss_key = None            # the caller may write into the returned tensors

    # ------------------------------------------------------------------ launch sequences
    def _state_key(self):
        return (self.flat.version, self._batch_id)

    def _trusted(self) -> bool:
        """The caller is the optimiser that tracks every parameter change (see FlatParams)."""
        return self.reuse and self.flat.tracked and self.flat.in_owner_step

    def _forward_is_current(self) -> bool:
        return self.reuse and self.flat.tracked and self._fwd_key == self._state_key()

    def _outputs(self):
        return self.ws[0].xhat, self.ws[1].xhat, self.ws[2].xhat

    def _forward(self, st):
        """The cascade forward of every micro-batch (each on its own stream when there are several)."""
        e = self.net.engine(), self.netT.engine(), self.netF.engine()
        many = len(self.mb) > 1
        if many:
            for i in range(3):                      # the weight images are shared: made once, before the chains fork
                e[i].prepare_images(self._pd[i], st, True)
        for m in self.mb:
            s = self._fork_to(m, "stream") if many else torch.cuda.current_stream(self.device)
            with torch.cuda.stream(s):
                self._forward_mb(m, s.cuda_stream, not many)
        if many:
            for m in self.mb:
                self._join(m.stream)

    def _forward_mb(self, m: "_MicroBatch", st: int, prepare: bool):
        L, Lt, C = self.L, self.Lt, self.C
        e = self.net.engine(), self.netT.engine(), self.netF.engine()
        x = self.x[m.n0:m.n1]
        xf, uv, Mu = x.view(m.N, -1), self.uv[m.n0:m.n1], self.Mu[m.n0:m.n1]
        x1, _ = e[0].forward(xf, uv, self.scales, self._pd[0], m.ws[0], st, mu_out=Mu[:, :L], x_planes=m.xp, prepare=prepare)
        if self.use_planes:
            lib().residual_split_planes(x.data_ptr(), x1.data_ptr(), m.pT.data_ptr(), m.pF.data_ptr(), m.N, C, 128, st)
            inT = inF = xf          # not read: the first conv of the 1-D nets takes the planes
        else:
            iyT, iyF = self.iyT.view(self.N, -1)[m.n0:m.n1], self.iyF.view(self.N, -1)[m.n0:m.n1]
            lib().residual_split(x.data_ptr(), x1.data_ptr(), iyT.data_ptr(), iyF.data_ptr(), m.N, C, 128, st)
            inT, inF = iyT, iyF
        # The time-axis and frequency-axis nets are independent: they run on two streams (fork / join by
        # events, also inside a graph capture), so the latency-bound deep layers of one overlap the other's.
        side = self._fork_to(m, "side")
        with torch.cuda.stream(side):
            e[2].forward(inF, uv, self.scales, self._pd[2], m.ws[2], side.cuda_stream, mu_out=Mu[:, L + Lt:],
                         x_planes=m.pF, prepare=prepare)
        e[1].forward(inT, uv, self.scales, self._pd[1], m.ws[1], st, mu_out=Mu[:, L:L + Lt], x_planes=m.pT, prepare=prepare)
        self._join(side)

    def _wstream(self, m: "_MicroBatch", i: int) -> Optional[torch.cuda.Stream]:
        """Stream for the weight / bias gradients of net i of a micro-batch (leaf work beside the data-gradient chain)."""
        if not self.overlap_streams:
            return None
        if m.wst[i] is None:
            m.wst[i] = torch.cuda.Stream(self.device)
        return m.wst[i]

    def _fork_to(self, owner, name: str) -> torch.cuda.Stream:
        """The stream `owner.<name>` (created on first use), made to start after everything queued so far on the
        current stream."""
        if not self.overlap_streams:
            return torch.cuda.current_stream(self.device)
        st = getattr(owner, name)
        if st is None:
            st = torch.cuda.Stream(self.device)
            setattr(owner, name, st)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        st.wait_event(ev)
        return st

    def _fork(self) -> torch.cuda.Stream:
        """Side stream (latent-space terms) that starts after everything queued so far on the current stream."""
        return self._fork_to(self, "_side")

    def _join(self, side: torch.cuda.Stream):
        if not self.overlap_streams:
            return
        ev = torch.cuda.Event()
        ev.record(side)
        torch.cuda.current_stream(self.device).wait_event(ev)

    def _seq_closure(self, grads: bool, forward: bool, upd: bool, yzero: bool = False):
        """The launch sequence of one closure evaluation (no host decisions inside: it can be captured)."""
        lb, st = lib(), _stream()
        N, C, L, Lt, Ltot, K = self.N, self.C, self.L, self.Lt, self.Ltot, self.mod.K
        M = self.mod.M
        plan = ShardPlan(N, self.Nglobal, self.world, self.bpb, C, K, Ltot)
        numel_g = plan.numel_global
        self.terms.zero_()
        tp = self.terms.data_ptr()
        if forward:
            self._forward(st)
        # the latent-space terms (a dozen small, latency-bound launches, whole minibatch) run beside the HBM-bound
        # cascade losses: forked here (before the loss passes are queued, which go first so that they own the SMs),
        # joined by every micro-batch chain before its backward passes
        lside = self._fork()
        many = len(self.mb) > 1
        chains = []
        for m in self.mb:
            s = self._fork_to(m, "stream") if many else torch.cuda.current_stream(self.device)
            chains.append(s)
            with torch.cuda.stream(s):
                self._loss_mb(m, s.cuda_stream, grads, upd, yzero, numel_g)
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
                if self.centre_sums:
                    ex = self.flat.extra_tail
                    ex.zero_()
                    lb.khm_center_sums(Mu.data_ptr(), Ltot, M.data_ptr(), N, K, Ltot, p, ex.data_ptr(),
                                       ex.data_ptr() + 4 * K * Ltot, sl)
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
            latent_done = torch.cuda.Event()
            latent_done.record(lside)
        for m, s in zip(self.mb, chains):
            if self.overlap_streams:
                s.wait_event(latent_done)
            if grads:
                with torch.cuda.stream(s):
                    self._backward_mb(m, s.cuda_stream)
        if many:
            for m in self.mb:
                self._join(m.stream)
            if grads:
                for m in self.mb[1:]:
                    lb.vec_add(self.flat.grad.data_ptr(), m.gbuf.data_ptr(), self.flat.numel, st)
        lb.closure_total(tp, self.rho, numel_g, khm_scale, self.flat.loss_tail.data_ptr(), st)

    def _loss_mb(self, m: "_MicroBatch", st: int, grads: bool, upd: bool, yzero: bool, numel_g: float):
        """Loss pass of one micro-batch (applies a deferred multiplier update, writes the reconstruction gradients)."""
        lb, N, C = lib(), self.N, self.C
        tp = self.terms.data_ptr()
        x = self.x[m.n0:m.n1]
        x1, x2, x3f = m.ws[0].xhat, m.ws[1].xhat, m.ws[2].xhat
        rows = lambda t: t.view(N, -1)[m.n0:m.n1]
        # the bias gradients of the three last transposed convs (= channel sums of the reconstruction
        # gradients) come out of the kernels that write those gradients
        fuse_db = grads and C <= 64
        planes = self.use_planes and fuse_db
        db2, db3 = ((m.gd[1]["tconv5.bias"].data_ptr(), m.gd[2]["tconv5.bias"].data_ptr()) if fuse_db else (None, None))
        y = [rows(t) for t in self._y]
        ymode = (1 if upd else 0) | (2 if yzero else 0)    # include/lshm.h lshm_cascade_losses_upd
        if planes:
            # gradient closure: d/dx2 and d/dx3 leave the loss pass as the operand planes of the 1-D nets' last layers
            lb.cascade_losses_planes(x.data_ptr(), x1.data_ptr(), x2.data_ptr(), x3f.data_ptr(),
                                     y[0].data_ptr(), y[1].data_ptr(), y[2].data_ptr(), self.rho, ymode,
                                     m.N, C, 128, 1.0 / numel_g, tp, rows(self.g1p).data_ptr(), m.p2.data_ptr(),
                                     m.p3.data_ptr(), db2, db3, st)
        else:
            g = [rows(t).data_ptr() for t in (self.g1p, self.g2, self.g3f)] if grads else [None, None, None]
            lb.cascade_losses_upd(x.data_ptr(), x1.data_ptr(), x2.data_ptr(), x3f.data_ptr(),
                                  y[0].data_ptr(), y[1].data_ptr(), y[2].data_ptr(), self.rho, ymode,
                                  m.N, C, 128, 1.0 / numel_g, tp, g[0], g[1], g[2], db2, db3, st)

    def _backward_mb(self, m: "_MicroBatch", st: int):
        """The three backward passes of one micro-batch, on the current stream (+ its side / leaf streams)."""
        lb, N, C, L, Lt = lib(), self.N, self.C, self.L, self.Lt
        rows = lambda t: t.view(N, -1)[m.n0:m.n1]
        fuse_db = C <= 64
        planes = self.use_planes and fuse_db
        e = self.net.engine(), self.netT.engine(), self.netF.engine()
        Mu, gMu = self.Mu[m.n0:m.n1], self.gMu[m.n0:m.n1]
        xf = self.x[m.n0:m.n1].view(m.N, -1)
        inT, inF = (xf, xf) if self.use_planes else (rows(self.iyT), rows(self.iyF))
        g1p = rows(self.g1p)
        g2, g3f = (None, None) if planes else (rows(self.g2), rows(self.g3f))
        side = self._fork_to(m, "side")
        with torch.cuda.stream(side):
            dF = e[2].backward(inF, self._pd[2], m.gd[2], m.ws[2], side.cuda_stream,
                               g3f, gMu[:, L + Lt:], Mu[:, L + Lt:], True, self._wstream(m, 2), fuse_db,
                               x_planes=m.pF, g_xhat_planes=m.p3 if planes else None)
        dT = e[1].backward(inT, self._pd[1], m.gd[1], m.ws[1], st, g2,
                           gMu[:, L:L + Lt], Mu[:, L:L + Lt], True, self._wstream(m, 1), fuse_db, x_planes=m.pT,
                           g_xhat_planes=m.p2 if planes else None)
        self._join(side)
        db1 = m.gd[0]["tconv5.bias"].data_ptr() if fuse_db else None
        if planes:
            lb.cascade_combine_planes(g1p.data_ptr(), dT.data_ptr(), dF.data_ptr(), m.gx1p.data_ptr(), m.N, C, 128, db1, st)
            e[0].backward(xf, self._pd[0], m.gd[0], m.ws[0], st, None, gMu[:, :L], Mu[:, :L], False,
                          self._wstream(m, 0), True, x_planes=m.xp, g_xhat_planes=m.gx1p)
        else:
            if self.gx1 is None:
                self.gx1 = torch.empty(N * C * 16384, dtype=torch.float32, device=self.device)
            gx1 = rows(self.gx1)
            lb.cascade_combine(g1p.data_ptr(), dT.data_ptr(), dF.data_ptr(), gx1.data_ptr(), m.N, C, 128, db1, st)
            e[0].backward(xf, self._pd[0], m.gd[0], m.ws[0], st, gx1,
                          gMu[:, :L], Mu[:, :L], False, self._wstream(m, 0), fuse_db, x_planes=m.xp)

    def _seq_forward(self):
        self._forward(_stream())

    def _seq_flush(self, yzero: bool = False):
        x1, x2, x3f = self._outputs()
        y = self._y
        lib().multiplier_update_z(self.x.data_ptr(), x1.data_ptr(), x2.data_ptr(), x3f.data_ptr(), self.rho,
                                  y[0].data_ptr(), y[1].data_ptr(), y[2].data_ptr(), 1 if yzero else 0,
                                  self.N, self.C, 128, _stream())

    # ------------------------------------------------------------------ CUDA graphs
    def enable_graphs(self, on: bool = True):
        """Replay every launch sequence (closure variants, forward, multiplier update) from a CUDA graph.
        A sequence is launched kernel by kernel the first time it is needed and captured right after
        (capturing executes nothing), so no state has to be saved and restored.  The minibatch then lives
        in the static tensors `x` / `uv` given to the first `set_batch` (later calls copy into them)."""
        if self.N == 0 and on:
            raise RuntimeError("lshm_b200: call step.set_batch(...) before enabling graphs")
        self._graphs_on = bool(on)
        if not on:
            self._graphs = {}

    def _run(self, key, fn):
        lb = lib()
        if not self._graphs_on:
            fn()
            return
        hit = self._graphs.get(key)
        if hit is not None:
            hit[0].replay()
            lb.launches += hit[1]
            return
        fn()                                    # this call's real execution (also warms up lazy state)
        cur = torch.cuda.current_stream(self.device)
        side = torch.cuda.Stream(self.device)
        side.wait_stream(cur)
        g = torch.cuda.CUDAGraph()
        l0 = lb.launches
        # (thread-local capture mode: the NCCL watchdog thread polls events while we capture.)
        with torch.cuda.stream(side):
            with torch.cuda.graph(g, stream=side, capture_error_mode="thread_local"):
                fn()
        n = lb.launches - l0
        lb.launches = l0                        # the capture launched nothing
        cur.wait_stream(side)
        self._graphs[key] = (g, n)

    # ------------------------------------------------------------------ public: closure
    def closure(self) -> torch.Tensor:
        """src/kharmonic_lofar.py:132-182.  Returns the total loss (0-dim device tensor)."""
        lb = lib()
        grads = torch.is_grad_enabled()
        start = lb.launches
        tail = self.flat.loss_tail
        tracked = self._trusted()
        # nothing changed since these loss scalars were computed (LBFGSNew's f_old probe): no launch at all
        if tracked and not grads and not self._pending and self._loss_key == self._state_key():
            self.launches = 0
            return tail[0].clone()
        fresh = tracked and self._forward_is_current()
        if self._pending and not fresh:
            self.flush_multipliers()            # the deferred update belongs to the activations still held
        upd, yz = self._pending, self._y_zero
        if grads:
            self.flat.attach_grads()
        self._run(("closure", grads, not fresh, upd, yz), lambda: self._seq_closure(grads, not fresh, upd, yz))
        self._pending = False
        if upd:
            self._y_zero = False            # the loss pass wrote the multipliers
        self._fwd_key = self._state_key() if tracked else None
        self._loss_key = self._state_key() if tracked else None
        if self.distributed:
            # ONE exchange per closure evaluation: gradients + loss scalars (+ centre sums); forward-only: scalars
            exchange(self.flat.grad if grads else tail, self.group)
        self.launches = lb.launches - start
        return tail[0].clone()

    def loss_terms(self) -> dict:
        """The columns printed at src/kharmonic_lofar.py:179 (one device->host copy)."""
        v = self.flat.loss_tail[:9].tolist()
        d = dict(total=v[0])
        d.update(zip(TERM_NAMES, v[1:]))
        return d

    # ------------------------------------------------------------------ public: multipliers
    def update_multipliers(self):
        """src/kharmonic_lofar.py:187-202: no-grad forward of the cascade, then y_i += rho*r_i.
        With a tracking optimiser the update itself is deferred into the next loss pass (class docstring)."""
        with torch.no_grad():
            if self._pending:
                self.flush_multipliers()
            if not self._forward_is_current():
                self._run(("forward",), self._seq_forward)
                if self.reuse and self.flat.tracked:
                    self._fwd_key = self._state_key()
            self._pending = True
            self._loss_key = None
            if not (self.reuse and self.flat.tracked):
                self.flush_multipliers()

    def flush_multipliers(self):
        """Apply a deferred multiplier update now (stand-alone pass over x, x1, x2, x3, y1..y3)."""
        if self._pending:
            self._pending = False
            self._loss_key = None
            yz = self._y_zero
            with torch.no_grad():
                self._run(("flush", yz), lambda: self._seq_flush(yz))
            self._y_zero = False

    # ------------------------------------------------------------------ public: centre update
    def centre_sums_view(self):
        """(num [K,Ltot], den [K]) views of the exchange buffer (valid after a gradient closure; already
        summed over ranks under data parallelism)."""
        if not self.centre_sums:
            raise RuntimeError("lshm_b200: construct the step with centre_sums=True")
        K, Lt = self.mod.K, self.Ltot
        ex = self.flat.extra_tail
        return ex[:K * Lt].view(K, Lt), ex[K * Lt:K * Lt + K]

    def apply_centre_update(self):
        """Kmeans.offline_update (src/lofar_models.py:231-261, Zhang GKHM 7.1-7.5) from the sums the last
        gradient closure left in the exchange buffer: M_k = num_k / den_k."""
        num, den = self.centre_sums_view()
        lib().khm_center_apply(num.data_ptr(), den.data_ptr(), self.mod.M.data_ptr(), self.mod.K, self.Ltot, _stream())
        self.flat.bump()
        self._fwd_key = None      # centres do not enter the activations, but keep the bookkeeping simple
        self._loss_key = None

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
