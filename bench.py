#!/usr/bin/env python
"""bench.py - headline metric of BASELINE.json: deep-K-harmonic TRAINING patches/sec
(fwd + bwd + K-harmonic + Adam + multiplier update) on N B200s, with roofline and CPU baseline.

    python bench.py --gpus N --steps K --warmup W            # ours (torchrun for N>1)
    python bench.py --impl reference --gpus N --steps K ...  # reference CPU path on host cores

Workload (config.workload = "cfg2"): 256 baselines x 2x2 half-overlapping 128x128 patches =
1024 patches per GPU per step, 8 channels, L=32, Lt=16 (Ltot=64), K=10, p=4, RICA on, Adam over
all four modules, rho=1.  A "step" is one ADMM iteration of
/root/reference/src/kharmonic_lofar.py:131-202: optimizer.step(closure) [cascade forward, losses,
analytic backward, Adam] + the no-grad multiplier-update forward.  Weak scaling: every rank
holds its own 1024 patches; one all-reduce of [gradients | loss scalars] per closure.

`value` = patches/s with the batch resident in HBM: consecutive ADMM iterations on the same minibatch, as in
the reference's `for admm` loop; the closure's forward is the multiplier-update forward of the previous
iteration (same parameters, same x - computed once) and the multiplier update is applied inside the next loss
pass (kharmonic_lofar.DeepKHarmonicStep).  `e2e` = the same iteration driven through the loader API from HOST
memory: pinned int8 visibilities of the step's baselines -> H2D -> scale/patchify/z-score kernels -> full
iteration (forward, backward, Adam, multiplier-update forward, multiplier update) -> D2H of the loss columns,
a FRESH minibatch every step (nothing can be reused across steps there).

The first closure of the run is checked against the CPU oracle on the same 1024 patches (`parity_check`).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np
import torch

CFG = dict(workload="cfg2", baselines_per_gpu=256, patches_per_baseline=4, patches_per_gpu=1024, channels=8,
           patch=128, L=32, Lt=16, K=10, p=4, rica=True, optimizer="Adam(lr=1e-4) over net+netT+netF+mod",
           admm_iterations_per_step=1, l2="inputs larger than L2 (x alone is 537 MB per GPU)",
           reuse="value: closure forward = preceding multiplier-update forward (same params, same x), multiplier "
                 "update fused into the next loss pass; e2e: fresh minibatch every step, nothing reused")
SCALES = [1e-4, 1e-3, 1e-2, 1e-1]


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return dict(hbm_gbs=d["hbm_gbs"], bf16_tflops=d["bf16_tflops"], bf16_sustained=d["bf16_tflops_sustained"],
                    source="measured")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_sustained=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (the sampler is
    started early because nvidia-smi needs ~1 s to produce its first row; rows are time-stamped
    on arrival and only those inside [mark_start, mark_stop] are summarised)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc, self.t0, self.t1 = index, [], None, None, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def mark_start(self):
        self.t0 = time.perf_counter()

    def mark_stop(self):
        self.t1 = time.perf_counter()

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        inside = [r for t, r in self.rows if self.t0 is not None and self.t0 <= t <= (self.t1 or t)]
        window = "timed region"
        if not inside:   # region shorter than the sampling period: fall back to the rows under load around it
            inside = [r for t, r in self.rows if self.t0 is not None and t >= self.t0 - 1.0]
            window = "timed region +-1s"
        sm, mx, reasons = [], [], set()
        for r in inside:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    samples=len(sm), window=window, reasons=sorted(reasons))


# ------------------------------------------------------------------------------------------
# algorithmic bytes / flops of one call (DESIGN.md "Kernels"); used for the per-kernel roofline
# ------------------------------------------------------------------------------------------
def call_cost(name, a):
    """(key, algorithmic HBM bytes, flops) for one library call with ctypes args `a`."""
    if name in ("down2d", "up2d", "wgrad2d", "down1d", "up1d", "wgrad1d"):
        two_d = name.endswith("2d")
        if name.startswith("wgrad"):
            N, A, B = a[5], a[6], a[7]
            small_px = a[8] * a[9] if two_d else a[8]
        else:
            N, A, B = a[8], a[9], a[10]
            small_px = a[11] * a[12] if two_d else a[11]
        taps = 16 if two_d else 4
        small, big, w = N * A * small_px, N * B * small_px * 4, A * B * taps
        flops = 2.0 * N * small_px * A * B * taps
        epi = None if name.startswith("wgrad") else a[-2]
        byts = 4.0 * (small + big + w) + (4.0 * (small if name.startswith("down") else big) if epi == 2 else 0.0)
        return f"{name}[A={A},B={B},px={small_px}]", byts, flops
    if name in ("cascade_losses",):
        n = a[8] * a[9] * a[10] * a[10]
        return name, 4.0 * n * (7 + (3 if a[13] else 0)), 30.0 * n
    if name == "cascade_losses_upd":
        # reads x, x1, x2, x3, y1..y3; writes 3 gradients (grad closure) and 3 multipliers (deferred update)
        n = a[9] * a[10] * a[11] * a[11]
        tag = ("+grads" if a[14] else "") + ("+mult_update" if a[8] else "")
        return f"cascade_losses{tag}", 4.0 * n * (7 + (3 if a[14] else 0) + (3 if a[8] else 0)), (30.0 + (6 if a[8] else 0)) * n
    if name in ("down2d_planes", "down1d_planes"):
        two_d = name.startswith("down2d")
        N, A, B = a[7], a[8], a[9]
        small_px = a[10] * a[11] if two_d else a[10]
        epi = a[-2]
        small, big = N * A * small_px, N * B * small_px * 4
        return f"{name}[A={A},B={B},px={small_px}]", 4.0 * (small + big) + (4.0 * small if epi == 2 else 0.0), 2.0 * N * small_px * A * B * (16 if two_d else 4)
    if name in ("wgrad2d_planes", "wgrad1d_planes"):
        two_d = name.startswith("wgrad2d")
        N, A, B = a[4], a[5], a[6]
        small_px = a[7] * a[8] if two_d else a[7]
        return f"{name}[A={A},B={B},px={small_px}]", 4.0 * (N * A * small_px + N * B * small_px * 4), 2.0 * N * small_px * A * B * (16 if two_d else 4)
    if name in ("tconv_bwd1d_planes", "tconv_bwd2d_planes"):
        # fused weight + data gradient of the last transposed conv: planes once, small map read (operand + ELU') and written
        two_d = name.endswith("2d_planes")
        N, A, B = a[7], a[8], a[9]
        small_px = a[10] * a[11] if two_d else a[10]
        small, big = N * A * small_px, N * B * small_px * 4
        return f"{name}[A={A},B={B},px={small_px}]", 4.0 * (2 * small + big), 4.0 * N * small_px * A * B * (16 if two_d else 4)
    if name == "stage_planes2d":
        n = a[3] * a[4] * a[5] * a[6] * 4
        return name, 8.0 * n, 3.0 * n
    if name == "residual_split_planes":
        n = a[4] * a[5] * a[6] * a[6]
        return name, 4.0 * n * 4, 9.0 * n
    if name == "cascade_combine_planes":
        n = a[4] * a[5] * a[6] * a[6]
        return name, 4.0 * n * 4, 6.0 * n
    if name == "cascade_losses_planes":
        # reads x, x1, x2, x3, y1..y3; writes g1p + two gradient planes (and 3 multipliers with the deferred update)
        n = a[9] * a[10] * a[11] * a[11]
        tag = "+grads(planes)" + ("+mult_update" if a[8] else "")
        return f"cascade_losses{tag}", 4.0 * n * (10 + (3 if a[8] else 0)), (36.0 + (6 if a[8] else 0)) * n
    if name in ("residual_split", "cascade_combine"):
        n = a[4] * a[5] * a[6] * a[6]
        return name, 4.0 * n * 4, 3.0 * n
    if name == "multiplier_update":
        n = a[8] * a[9] * a[10] * a[10]
        return name, 4.0 * n * 10, 8.0 * n
    if name in ("khm_fwd", "khm_fwd_bwd", "khm_bwd"):
        N, K, L = a[3], a[4], a[5]
        bwd = name != "khm_fwd"
        return f"{name}[K={K},L={L}]", 4.0 * N * L * (2 if bwd else 1), (3.0 * L + 6) * K * N * (3 if bwd else 1)
    if name == "channel_sum":
        N, Cn, ln = a[3], a[4], a[5]
        return f"channel_sum[C={Cn},len={ln}]", 4.0 * N * Cn * ln, 1.0 * N * Cn * ln
    if name == "linear_fwd":
        N, K, J = a[6], a[7], a[8]
        return f"linear_fwd[K={K},J={J}]", 4.0 * (N * K + K * J + N * J), 2.0 * N * K * J
    if name == "linear_bwd_data":
        N, K, J = a[9], a[10], a[11]
        return f"linear_bwd_data[K={K},J={J}]", 4.0 * (N * J + K * J + N * K), 2.0 * N * K * J
    if name == "linear_bwd_weight":
        N, K, J = a[6], a[7], a[8]
        return f"linear_bwd_weight[K={K},J={J}]", 4.0 * (N * K + N * J + K * J), 2.0 * N * K * J
    return name, 0.0, 0.0


def ncu_traffic(kernel):
    """DRAM bytes per launch of `kernel` from the committed ncu capture (profiles/), or None."""
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles")
    for fn in ("r2_ncu_dram_traffic.json", "r1_ncu_dram_traffic.json"):
        try:
            with open(os.path.join(here, fn)) as fh:
                v = json.load(fh).get(kernel)
            if v is not None:
                return v
        except Exception:
            pass
    return None


class KernelProfiler:
    """CUDA-event timing of every library call (events on the launching stream)."""

    def __init__(self, L):
        self.L, self.records, self.saved = L, [], {}

    def __enter__(self):
        for name in self.L.protos:
            short = name[len("lshm_"):]
            fn = getattr(self.L, short, None)
            if fn is None or short in ("version", "device_info"):
                continue
            self.saved[short] = fn
            setattr(self.L, short, self._wrap(short, fn))
        return self

    def _wrap(self, short, fn):
        def call(*args):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn(*args)
            e1.record()
            self.records.append((short, args, e0, e1))
        return call

    def __exit__(self, *exc):
        for short, fn in self.saved.items():
            setattr(self.L, short, fn)

    def summary(self, steps):
        torch.cuda.synchronize()
        agg = {}
        for short, args, e0, e1 in self.records:
            key, byts, flops = call_cost(short, args)
            d = agg.setdefault(key, dict(ms=0.0, calls=0, bytes=0.0, flops=0.0))
            d["ms"] += e0.elapsed_time(e1); d["calls"] += 1; d["bytes"] += byts; d["flops"] += flops
        for d in agg.values():
            d["ms_per_step"] = d["ms"] / steps
        return agg


def build_step(dev, rank, world, distributed):
    from lshm_b200.kharmonic_lofar import DeepKHarmonicStep, FlatAdam
    from lshm_b200.lofar_models import AutoEncoder1DCNN, AutoEncoderCNN2, Kmeans
    torch.manual_seed(0)  # identical random init on every rank
    hs = torch.tensor(SCALES).to(dev)
    C, L, Lt, K = CFG["channels"], CFG["L"], CFG["Lt"], CFG["K"]
    net = AutoEncoderCNN2(L, C, hs, True).to(dev)
    netT = AutoEncoder1DCNN(Lt, C, hs, True).to(dev)
    netF = AutoEncoder1DCNN(Lt, C, hs, True).to(dev)
    mod = Kmeans(L + 2 * Lt, K, CFG["p"]).to(dev)
    step = DeepKHarmonicStep(net, netT, netF, mod, distributed=distributed)
    opt = FlatAdam(step.flat, lr=1e-4)
    return step, opt


def run_ours(args):
    from lshm_b200 import lofar_tools as T
    from lshm_b200 import synthetic as S
    from lshm_b200._lib import lib
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    distributed = world > 1
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if distributed:
        import torch.distributed as dist
        import datetime
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))
    sampler = ClockSampler(local)
    sampler.start()
    step, opt = build_step(dev, rank, world, distributed)
    nb, bpb = CFG["baselines_per_gpu"], CFG["patches_per_baseline"]
    Np = nb * bpb
    # synthetic observation of this rank: 192x192 -> 2x2 patches per baseline
    meas = S.make_measurement(nb, 192, 192, seed=100 + rank)
    sap = meas["measurement"]["saps"]["0"]
    vis_h = torch.from_numpy(sap["visibilities"]).pin_memory()
    sc_h = torch.from_numpy(sap["visibility_scale_factors"]).pin_memory()
    uv_h = torch.from_numpy(S.make_uv(Np, seed=rank, per_group=bpb)).pin_memory()
    sel = torch.arange(nb, dtype=torch.int32, device=dev)

    def load_from_host(bufs=None):
        if bufs is None:
            vis = vis_h.to(dev, non_blocking=True)
            sc = sc_h.to(dev, non_blocking=True)
            uv = uv_h.to(dev, non_blocking=True)
            px, py, x = T.patchify_device(vis, sc, sel, 128, CFG["channels"], 1e3, True)
        else:   # staging loop: preallocated device buffers, nothing is allocated
            vis, sc, uv, y, stats = bufs
            vis.copy_(vis_h, non_blocking=True); sc.copy_(sc_h, non_blocking=True); uv.copy_(uv_h, non_blocking=True)
            px, py, x = T.patchify_device(vis, sc, sel, 128, CFG["channels"], 1e3, True, out=y, stats=stats)
        # the reference orders uv rows baseline-major while patches are patch-major (SURVEY.md bug 3);
        # reproduced as is
        return px * py, x, uv

    bpb_, x, uv = load_from_host()
    step.set_batch(x, uv, bpb_, global_patches=Np * world)

    # ---- parity of the batch this run times: first closure (y = 0) against the CPU oracle on the SAME 1024 patches
    parity = None
    if rank == 0 and world == 1 and args.parity == "on":
        parity = parity_check(step, x, uv, bpb_)

    def one_step():
        loss = opt.step(step.closure)
        step.update_multipliers()
        return loss

    # every launch sequence (loss pass + backward, forward, ...) is replayed from a CUDA graph
    graph_note = None
    if args.graph == "on":
        try:
            step.enable_graphs()
            one_step(); one_step()          # first use of each sequence: launched kernel by kernel, then captured
            torch.cuda.synchronize()
        except Exception as exc:      # keep the benchmark alive: same kernels, launched one by one
            graph_note = f"graph capture failed ({type(exc).__name__}: {exc}); eager launches"
            print(graph_note, file=sys.stderr)
            step.enable_graphs(False)
            torch.cuda.synchronize()
    graphed = step._graphs_on

    def barrier():
        if distributed:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        one_step()
    barrier()
    L = lib()
    sampler.mark_start()
    l0 = L.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        one_step()
    e1.record()
    barrier()
    sampler.mark_stop()
    ms = e0.elapsed_time(e1)
    launches = L.launches - l0
    clocks = sampler.stop()
    if distributed:
        t = torch.tensor([ms], device=dev)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms = float(t)
    ms_per_step = ms / args.steps
    value = Np * world / (ms_per_step * 1e-3)

    # ---- end to end through the loader API from host memory
    # the next minibatch (H2D + loader kernels) is staged on a side stream while this step computes;
    # every step's copy and read-back are inside the timed region.  A fresh minibatch every step: the
    # closure runs its own forward, and the multiplier update of the iteration is applied (flush) although
    # the next minibatch resets y1..y3 - the reference does that work too.
    def new_batch(x2, uv2, b):
        step.set_batch(x2, uv2, b, global_patches=Np * world)

    def e2e_step():
        one_step()
        step.flush_multipliers()

    pf = T.DevicePrefetcher(dev, record_streams=False)
    sets = [(torch.empty_like(vis_h, device=dev), torch.empty_like(sc_h, device=dev), torch.empty_like(uv_h, device=dev),
             torch.empty(Np, CFG["channels"], 128, 128, device=dev), torch.zeros(2, dtype=torch.float64, device=dev))
            for _ in range(2)]
    for k in range(3):      # warm-up through the same path
        pf.submit(lambda: load_from_host(sets[k % 2]))
        b, x2, uv2 = pf.get(); new_batch(x2, uv2, b); e2e_step(); step.loss_terms()
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(2, min(args.steps, 10))
    pf.submit(lambda: load_from_host(sets[0]))
    for i in range(e2e_steps):
        b, x2, uv2 = pf.get()
        if i + 1 < e2e_steps:
            nxt = sets[(i + 1) % 2]      # last used by step i-1, which has completed (loss read-back)
            pf.submit(lambda: load_from_host(nxt))
        new_batch(x2, uv2, b)
        e2e_step()
        terms = step.loss_terms()   # D2H of the 9 loss columns (synchronises)
    barrier()
    e2e_s = time.perf_counter() - t0
    if distributed:
        t = torch.tensor([e2e_s], device=dev)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        e2e_s = float(t)
    e2e = dict(value=Np * world / (e2e_s / e2e_steps), unit="patches/s",
               h2d_bytes_per_step=int(vis_h.numel() + sc_h.numel() * 4 + uv_h.numel() * 4), d2h_bytes_per_step=9 * 4,
               steps=e2e_steps, note="fresh minibatch from pinned host int8 every step (staged on a side stream): full "
               "forward + backward + Adam + multiplier-update forward + multiplier update, then loss read-back")

    # ---- per-kernel device times (CUDA events) over two extra steps -> dominant kernel roofline.
    #      Every rank runs the steps (the closure contains the all-reduce); rank 0 records.
    #      Same steady-state iteration as `value` (resident minibatch, reused forward), launched kernel by kernel.
    prof_steps = 2
    agg = None
    step.enable_graphs(False)
    step.overlap_streams = False     # one stream: the event pair around a call then times that kernel alone
    one_step()                       # settle into the steady state (forward held, multiplier update pending)
    if rank == 0:
        with KernelProfiler(L) as kp:
            for _ in range(prof_steps):
                one_step()
            agg = kp.summary(prof_steps)
    else:
        for _ in range(prof_steps):
            one_step()
    barrier()
    out = None
    if rank == 0:
        pk = peaks()
        total_ms = sum(d["ms_per_step"] for d in agg.values())
        top = sorted(agg.items(), key=lambda kv: -kv[1]["ms"])
        name, d = next((kv for kv in top if kv[1]["bytes"] > 0), top[0])
        sec = d["ms"] * 1e-3
        hbm_frac = (d["bytes"] / sec / 1e9) / pk["hbm_gbs"] if sec > 0 else 0.0
        intensity = d["flops"] / d["bytes"] if d["bytes"] else 0.0
        roof = dict(kernel=name, bound="hbm", achieved=d["bytes"] / sec / 1e9, peak=pk["hbm_gbs"], unit="GB/s",
                    frac=hbm_frac, traffic=ncu_traffic(name), peak_source=pk["source"], ms_per_launch=d["ms"] / d["calls"],
                    share_of_step=d["ms_per_step"] / total_ms if total_ms else None, flop_per_byte=intensity,
                    achieved_tflops=d["flops"] / sec / 1e12,
                    top5=[dict(kernel=k, ms_per_step=round(v["ms_per_step"], 4), calls_per_step=v["calls"] // prof_steps,
                               gbs=round(v["bytes"] / (v["ms"] * 1e-3) / 1e9, 1) if v["ms"] else 0,
                               tflops=round(v["flops"] / (v["ms"] * 1e-3) / 1e12, 2) if v["ms"] else 0) for k, v in top[:5]])
        if args.kernel_table:
            with open(args.kernel_table, "w") as fh:
                fh.write("kernel,calls_per_step,ms_per_step,share,GB/s(algorithmic),TFLOP/s\n")
                for k, v in top:
                    s_ = v["ms"] * 1e-3
                    fh.write(f"{k},{v['calls'] // prof_steps},{v['ms_per_step']:.4f},{v['ms_per_step'] / total_ms:.4f},"
                             f"{v['bytes'] / s_ / 1e9 if s_ else 0:.1f},{v['flops'] / s_ / 1e12 if s_ else 0:.2f}\n")
                fh.write(f"TOTAL,,{total_ms:.4f},1.0,,\n")
        # the CPU arm is timed on rank 0 at N=1 only (contract); other world sizes report null
        cpu = cpu_baseline(sample_patches=args.cpu_patches, steps=1) if (world == 1 and args.cpu == "on") else None
        out = {
            "metric": "train patches/sec (fwd+bwd+K-harmonic)", "value": value, "unit": "patches/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(CFG, global_patches=Np * world, parallelism=f"dp{world}", cuda_graph=bool(graphed), **({"graph_note": graph_note} if graph_note else {})),
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cpu,
            "parity_check": parity, "loss_terms_last": terms,
        }
    if distributed:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()
    return json.dumps(out) if out is not None else None


# ------------------------------------------------------------------------------------------
# parity of the timed batch (oracle = checker)
# ------------------------------------------------------------------------------------------
def parity_check(step, x, uv, bpb):
    """First closure of the run (y1..y3 = 0) on the GPU vs the CPU oracle on the SAME patches and parameters:
    the nine loss columns and the argmin assignment of every latent.  Forward only (the gradients of this
    batch size are compared in tests/test_gpu_sizes.py::test_fused_closure_at_benchmark_sizes)."""
    from oracle import lofar_oracle as O
    t0 = time.perf_counter()
    with torch.no_grad():
        step.closure()
    got = step.loss_terms()
    Mu_gpu = step.latents().detach().cpu()
    ids_gpu = step.mod.assign(step.latents()).cpu().long()
    cpu = lambda sd: {k: v.detach().cpu() for k, v in sd.items()}
    pn, pT, pF = cpu(step.net.state_dict()), cpu(step.netT.state_dict()), cpu(step.netF.state_dict())
    M = step.mod.M.detach().cpu()
    xc, uvc = x.detach().cpu(), uv.detach().cpu()
    N = xc.shape[0]
    zeros = torch.zeros(xc.numel())
    torch.set_num_threads(os.cpu_count() or 1)
    with torch.no_grad():
        total, terms = O.closure_losses(pn, pT, pF, M, xc, uvc, torch.tensor(SCALES), zeros, zeros, zeros,
                                        batch_per_bline=bpb, batch_size=N // bpb, Khp=CFG["p"])
    ref = {k: float(v) for k, v in terms.items() if k != "Mu"}
    ref["total"] = float(total)
    errs = {k: abs(got[k] - ref[k]) / max(abs(ref[k]), 1e-30) for k in ref}
    ids_ref = torch.cdist(terms["Mu"].double(), M.double()).argmin(dim=1)
    agree = float((ids_gpu == ids_ref).double().mean())
    mu_err = float((Mu_gpu.double() - terms["Mu"].double()).norm() / terms["Mu"].double().norm())
    worst = max(errs.values())
    step.invalidate()
    return dict(patches=int(N), against="oracle/lofar_oracle.closure_losses (CPU fp32), same patches and parameters",
                loss_terms_max_rel_err=worst, loss_terms_rel_err={k: float(f"{v:.3e}") for k, v in errs.items()},
                latents_rel_err=mu_err, assignment_agreement=agree, tolerance=dict(loss=1e-3, assignment=0.999),
                ok=bool(worst < 1e-3 and agree >= 0.999), seconds=round(time.perf_counter() - t0, 2))


# ------------------------------------------------------------------------------------------
# CPU arm: the reference's own modules (baseline/_ref) on the host cores; oracle port only if they are absent
# ------------------------------------------------------------------------------------------
def cpu_step_factory(n_patches, prefer_reference=True):
    """Returns (one_step, kind, info): one ADMM iteration (closure + Adam + multiplier update) on `n_patches`
    of the cfg2 batch.  kind "reference": the UNMODIFIED reference modules (lofar_models.py from baseline/_ref)
    driven by the restated script loop (oracle/reference_loop.py); kind "port": the oracle port."""
    from lshm_b200 import synthetic as S
    C, L, Lt, K, bpb = CFG["channels"], CFG["L"], CFG["Lt"], CFG["K"], CFG["patches_per_baseline"]
    x = torch.from_numpy(S.make_patches(n_patches, C, seed=5))
    uv = torch.from_numpy(S.make_uv(n_patches, seed=5, per_group=bpb))
    if prefer_reference:
        from oracle import reference_loop as RL
        if RL.reference_dir() is not None:
            R = RL.ReferenceLoop(L=L, Lt=Lt, C=C, K=K, Khp=CFG["p"], optimizer="adam", param_modules=(0, 1, 2, 3))
            R.set_batch(x, uv, bpb)
            return R.admm_iteration, "reference", R
    from oracle import lofar_oracle as O
    hs = torch.tensor(SCALES)
    pn = O.make_ae_params(L, C, ndim=2, seed=1); pT = O.make_ae_params(Lt, C, ndim=1, seed=2)
    pF = O.make_ae_params(Lt, C, ndim=1, seed=3); M = O.make_centres(K, L + 2 * Lt, seed=4).requires_grad_()
    params = []
    for p in (pn, pT, pF):
        for v in p.values():
            v.requires_grad_(); params.append(v)
    params.append(M)
    opt = torch.optim.Adam(params, lr=1e-4)
    ys = [torch.zeros(x.numel()) for _ in range(3)]

    def closure():
        opt.zero_grad()
        total, _ = O.closure_losses(pn, pT, pF, M, x, uv, hs, *ys, batch_per_bline=bpb, batch_size=n_patches // bpb,
                                    Khp=CFG["p"])
        total.backward()
        return total

    def one_step():
        opt.step(closure)
        ys[:] = O.multiplier_update(pn, pT, pF, x, uv, hs, *ys)
    return one_step, "port", None


REF_NOTE = ("unmodified reference modules (lofar_models.py: AutoEncoderCNN2 / AutoEncoder1DCNN / Kmeans with its Python "
            "N x K loop) from baseline/_ref + the script loop of src/kharmonic_lofar.py:84-202 restated in "
            "oracle/reference_loop.py; torch CPU fp32, Adam over all four modules")
PORT_NOTE = ("oracle port of /root/reference/src/kharmonic_lofar.py:131-202 (baseline/_ref absent: run "
             "__graft_entry__.build() where /root/reference exists); vectorised K-harmonic")


def timed_cpu_run(n_patches, warmup, steps, budget_s):
    """W warm-up + K timed ADMM iterations on n_patches; if the first iteration shows that the whole run would not
    fit in budget_s, the batch is halved (time is ~linear in patches) until it does - said in `sample`."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    n = n_patches
    while True:
        one_step, kind, R = cpu_step_factory(n)
        t0 = time.perf_counter()
        one_step()
        first = time.perf_counter() - t0
        if first * (warmup + steps) <= budget_s or n <= 64:
            break
        n = max(64, n // 2)
    for _ in range(max(0, warmup - 1)):
        one_step()
    if R is not None:
        R.khm_seconds, R.closures = 0.0, 0
    t0 = time.perf_counter()
    for _ in range(steps):
        one_step()
    dt = (time.perf_counter() - t0) / steps
    extra = {}
    if R is not None:
        extra = dict(khm_python_loop_forward_ms_per_step=1e3 * R.khm_seconds / steps,
                     khm_python_loop_share_of_step=R.khm_seconds / steps / dt, closures_per_step=R.closures / steps)
    sample = (f"{warmup} warm-up + {steps} timed ADMM iteration(s) (closure + Adam + multiplier update) on {n} of the "
              f"{n_patches} patches of the cfg2 batch" + ("" if n == n_patches else f" (batch reduced so that the run fits in {budget_s:.0f} s)")
              + "; " + (REF_NOTE if kind == "reference" else PORT_NOTE))
    return dict(value=n / dt, unit="patches/s", cores=cores, kind=kind, sample=sample, sample_patches=n,
                ms_per_step=dt * 1e3, **extra)


def cpu_baseline(sample_patches=1024, steps=1):
    """Rank-0 CPU baseline of the `ours` arm: one warm-up + `steps` timed iterations, bounded to ~40 s."""
    return timed_cpu_run(sample_patches, 1, steps, budget_s=40.0)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return None
    warmup, steps = max(args.warmup, 0), max(args.steps, 1)
    r = timed_cpu_run(args.cpu_patches, warmup, steps, budget_s=args.cpu_budget)
    n = r["sample_patches"]
    return json.dumps({
        "impl": "reference", "metric": "train patches/sec (fwd+bwd+K-harmonic)", "value": r["value"], "unit": "patches/s",
        "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": steps, "warmup": warmup,
        "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": dict(CFG, sample_patches=n, reuse="none (reference loop as written)"),
        "cpu_baseline": {k: v for k, v in r.items() if k != "ms_per_step"},
        "e2e": dict(value=r["value"], unit="patches/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0)})


class StdoutToStderr:
    """Everything written to fd 1 while active goes to stderr (NCCL prints its version banner on
    stdout); the contract is ONE JSON line on stdout."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-patches", type=int, default=CFG["patches_per_gpu"],
                    help="patches per CPU step (default: the full cfg2 batch; reduced automatically if too slow)")
    ap.add_argument("--cpu-budget", type=float, default=270.0, help="wall-time budget (s) of the --impl reference run")
    ap.add_argument("--cpu", default="on", choices=["on", "off"], help="time the CPU baseline in the `ours` arm")
    ap.add_argument("--parity", default="on", choices=["on", "off"], help="check the first closure against the CPU oracle")
    ap.add_argument("--graph", default="on", choices=["on", "off"],
                    help="replay the step from one CUDA graph (default) or launch it kernel by kernel")
    ap.add_argument("--kernel-table", default=None, help="write the per-kernel CUDA-event table (CSV) here")
    args = ap.parse_args()
    with StdoutToStderr():
        line = run_reference(args) if args.impl == "reference" else run_ours(args)
    if line is not None:
        print(line, flush=True)


if __name__ == "__main__":
    main()
