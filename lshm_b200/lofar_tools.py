"""Loader functions with the signatures of /root/reference/src/lofar_tools.py, on CUDA kernels.

``get_data_minibatch`` (:51-211), ``get_data_for_baseline`` (:214-349),
``get_data_for_baseline_flat`` (:352-406), ``get_metadata`` (:410-426), ``get_fileSAP``
(:430-463), ``torch_fftshift`` (:24-30).  HDF5 I/O stays on the host (h5py when installed);
an entry of ``file_list`` may also be an already-open mapping with the same group layout
(see :mod:`lshm_b200.synthetic`).  The int8 visibilities of the selected baselines are
uploaded as int8 (1 byte/sample, pinned staging) and scale * patchify * clamp * statistics
run as one kernel, the z-score as a second one (lshm_patchify_scale_i8, lshm_normalise).

The numpy RNG calls are made in the same order as the reference (:71, :88) so a seeded run
draws the same file and baselines.
"""
from __future__ import annotations

import glob
import math
import os
from typing import Mapping

import numpy as np
import torch

from ._lib import lib

rec_file_search = True
C_LIGHT = 2.99792458e8


def _device():
    if not torch.cuda.is_available():
        raise RuntimeError("lshm_b200.lofar_tools needs a CUDA device (no CPU path)")
    return torch.device("cuda", torch.cuda.current_device())


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _open(filename):
    if isinstance(filename, Mapping):
        return filename
    try:
        import h5py
    except ImportError as e:  # pragma: no cover - depends on the image
        raise RuntimeError("h5py is required to read %r" % (filename,)) from e
    return h5py.File(filename, "r")


def torch_fftshift(real, imag):
    """src/lofar_tools.py:24-30: roll dims >= 2 by size//2."""
    for dim in range(2, len(real.size())):
        real = torch.roll(real, dims=dim, shifts=real.size(dim) // 2)
        imag = torch.roll(imag, dims=dim, shifts=imag.size(dim) // 2)
    return real, imag


def fft_features(x: torch.Tensor, xhat: torch.Tensor = None, clamp: float = 10.0, mode: str = "reim") -> torch.Tensor:
    """Demo.ipynb:169-174 as one kernel: cat(Re,Im)(fftshift(fft2_ortho(x - xhat))).clamp(+-10).
    x [N,C,128,128] -> [N,2C,128,128].  mode="magphase": cat(|F|.clamp(max=clamp), angle(F)) of the same F."""
    if not x.is_cuda or x.dtype != torch.float32 or tuple(x.shape[2:]) != (128, 128):
        raise RuntimeError("lshm_b200: fft_features expects a CUDA float32 [N,C,128,128] tensor")
    if mode not in ("reim", "magphase"):
        raise ValueError("fft_features: mode must be 'reim' or 'magphase'")
    x = x.contiguous()
    if xhat is not None:
        xhat = xhat.contiguous()
    N, C = x.shape[:2]
    out = torch.empty(N, 2 * C, 128, 128, dtype=torch.float32, device=x.device)
    lib().fft2_features(x.data_ptr(), None if xhat is None else xhat.data_ptr(), out.data_ptr(),
                        N, C, float(clamp), 0 if mode == "reim" else 1, _stream())
    return out


def _to_device_i8(arr: np.ndarray, dev) -> torch.Tensor:
    t = torch.from_numpy(np.ascontiguousarray(arr))
    return t.pin_memory().to(dev, non_blocking=True)


def patchify_device(vis: torch.Tensor, scale: torch.Tensor, sel: torch.Tensor, patch_size: int,
                    num_channels: int, clamp: float, normalize: bool, out: torch.Tensor = None,
                    stats: torch.Tensor = None, group=None, n_global: int = None):
    """vis int8 [nbase,T,F,4,2], scale fp32 [nbase,F,4], sel int32 [nb] (all on device) ->
    (patchx, patchy, y [nb*px*py, C, P, P]).  `out` / `stats` (fp64 [2]): optional preallocated
    destination and scratch, so a staging loop allocates nothing.

    `group` + `n_global`: the minibatch is sharded over the ranks of a torch.distributed group; the z-score
    (src/lofar_tools.py:190-193: mean / unbiased std over the WHOLE minibatch) then uses the all-reduced
    (sum, sum of squares) of all `n_global` elements.  Without a group every rank normalises its own shard
    with its own statistics (N independent reference loaders)."""
    nbase, T, F = vis.shape[:3]
    nb, P = sel.numel(), patch_size
    s = P // 2
    px = (max(T, P) - P) // s + 1
    py = (max(F, P) - P) // s + 1
    shape = (nb * px * py, num_channels, P, P)
    if out is None:
        y = torch.empty(shape, dtype=torch.float32, device=vis.device)
    else:
        if tuple(out.shape) != shape or out.dtype != torch.float32 or not out.is_contiguous() or out.device != vis.device:
            raise ValueError(f"patchify_device: out must be a contiguous float32 {shape} tensor on {vis.device}")
        y = out
    if stats is None:
        stats = torch.zeros(2, dtype=torch.float64, device=vis.device)
    else:
        stats.zero_()
    st = _stream()
    lib().patchify_scale_i8(vis.data_ptr(), scale.data_ptr(), sel.data_ptr(), nb, T, F, num_channels, P,
                            float(clamp), y.data_ptr(), stats.data_ptr(), st)
    if normalize:
        if group is not None:
            import torch.distributed as dist
            dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
            lib().normalise_n(y.data_ptr(), y.numel(), stats.data_ptr(), int(n_global), st)
        else:
            lib().normalise(y.data_ptr(), y.numel(), stats.data_ptr(), st)
    return px, py, y


class DevicePrefetcher:
    """Runs a batch loader (pinned-host -> device copies + the loader kernels) on a side stream so the
    next minibatch is staged while the current closure / optimiser step computes.

        pf = DevicePrefetcher(device); pf.submit(load)       # load() returns tensors / tuples of tensors
        batch = pf.get(); pf.submit(load); ...train on batch...
    """

    def __init__(self, device, record_streams: bool = True):
        # record_streams=False: the loader writes into caller-owned, preallocated buffers (no allocator
        # hand-over between the two streams; the caller alternates buffer sets)
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(self.device)
        self.record_streams = record_streams
        self._out, self._ev = None, None

    def submit(self, load):
        self.stream.wait_stream(torch.cuda.current_stream(self.device))   # buffers the loader reuses are free
        with torch.cuda.stream(self.stream):
            self._out = load()
            self._ev = torch.cuda.Event()
            self._ev.record(self.stream)

    def get(self):
        if self._ev is None:
            raise RuntimeError("DevicePrefetcher.get() before submit()")
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(self._ev)

        def mark(o):
            if isinstance(o, torch.Tensor):
                if o.is_cuda:
                    o.record_stream(cur)
            elif isinstance(o, (tuple, list)):
                for e in o:
                    mark(e)
        if self.record_streams:
            mark(self._out)
        out, self._out, self._ev = self._out, None, None
        return out


def _uv_rotation(f, SAP):
    # src/lofar_tools.py:90-106
    hms = f["measurement"]["info"]["start_time"][0].decode("ascii").split()[1].split(sep=":")
    start_time = float(hms[0]) + float(hms[1]) / 60.0 + float(hms[2]) / 3600
    theta = start_time / 24.0 * (2 * math.pi)
    frq = f["measurement"]["saps"][SAP]["central_frequencies"]
    freq0 = frq[frq.shape[0] // 2]
    inv_lambda = freq0 / C_LIGHT
    return math.cos(theta) * inv_lambda, math.sin(theta) * inv_lambda


def _uv_of(f, SAP, baselinelist, rot00, rot01):
    # src/lofar_tools.py:143-151
    baselines = f["measurement"]["saps"][SAP]["baselines"]
    xyz = f["measurement"]["saps"][SAP]["antenna_locations"]["XYZ"]
    uv = np.zeros((len(baselinelist), 2), np.float32)
    for ck, mybase in enumerate(baselinelist):
        xx = xyz[baselines[mybase][0]][0] - xyz[baselines[mybase][1]][0]
        yy = xyz[baselines[mybase][0]][1] - xyz[baselines[mybase][1]][1]
        uv[ck, 0] = xx * rot00 + yy * rot01
        uv[ck, 1] = -xx * rot01 + yy * rot00
    return uv


_PINNED = {}     # (tag, shape, dtype) -> [pinned host tensor, event of its last upload]


def _pinned_like(tag, shape, dtype):
    """Reusable page-locked staging buffer (pinning 76 MB per minibatch costs more than copying it).  A buffer
    is handed out again only after the upload that last used it has completed."""
    key = (tag, tuple(shape), dtype)
    ent = _PINNED.get(key)
    if ent is None:
        ent = _PINNED[key] = [torch.empty(shape, dtype=dtype).pin_memory(), None]
    elif ent[1] is not None:
        ent[1].synchronize()
    return ent


def _load_selected(g, h, baselinelist, dev):
    """int8 visibilities [nb,T,F,4,2] and fp32 scale factors [nb,F,4] of the selected baselines: read straight
    into pinned staging buffers (one row per baseline, no intermediate stack) and uploaded asynchronously."""
    nb = len(baselinelist)
    first_v, first_s = np.asarray(g[int(baselinelist[0])]), np.asarray(h[int(baselinelist[0])])
    ev, es = _pinned_like("vis", (nb,) + first_v.shape, torch.int8), _pinned_like("scale", (nb,) + first_s.shape, torch.float32)
    hv, hs = ev[0].numpy(), es[0].numpy()
    for k, b in enumerate(baselinelist):
        hv[k] = first_v if k == 0 else np.asarray(g[int(b)])
        hs[k] = first_s if k == 0 else np.asarray(h[int(b)])
    vis, sc = ev[0].to(dev, non_blocking=True), es[0].to(dev, non_blocking=True)
    done = torch.cuda.Event()
    done.record(torch.cuda.current_stream(dev))
    ev[1] = es[1] = done
    return vis, sc


def get_data_minibatch(file_list, SAP_list, batch_size=2, patch_size=32, normalize_data=False,
                       num_channels=8, transform=None, uvdist=False):
    """src/lofar_tools.py:51-211.  Rows are ordered patch-major, n=(ci*py+cj)*nb+k (:169-173),
    uv rows baseline-major (:175-178) - both exactly as the reference."""
    assert len(file_list) == len(SAP_list)
    assert num_channels == 4 or num_channels == 8
    dev = _device()
    file_id = np.random.randint(0, len(file_list))
    f = _open(file_list[file_id])
    SAP = SAP_list[file_id]
    g = f["measurement"]["saps"][SAP]["visibilities"]
    h = f["measurement"]["saps"][SAP]["visibility_scale_factors"]
    nbase = g.shape[0]
    baselinelist = np.random.randint(0, nbase, batch_size)
    vis, sc = _load_selected(g, h, baselinelist, dev)
    sel = torch.arange(batch_size, dtype=torch.int32, device=dev)
    patchx, patchy, y = patchify_device(vis, sc, sel, patch_size, num_channels, 1e3, normalize_data)
    if uvdist:
        rot00, rot01 = _uv_rotation(f, SAP)
        uv = _uv_of(f, SAP, baselinelist, rot00, rot01)
        uv1 = torch.from_numpy(np.repeat(uv, patchx * patchy, axis=0)).to(dev)
    if transform:
        # src/lofar_tools.py:196-203: interleave original and transformed groups per baseline
        bpb = patchx * patchy
        y1 = torch.zeros(2 * y.shape[0], *y.shape[1:], dtype=y.dtype, device=dev)
        for ci in range(batch_size):
            y1[2 * ci * bpb:(2 * ci + 1) * bpb] = y[ci * bpb:(ci + 1) * bpb]
            y1[(2 * ci + 1) * bpb:(2 * ci + 2) * bpb] = transform(y[ci * bpb:(ci + 1) * bpb])
        y = y1
    if uvdist:
        return patchx, patchy, y, uv1
    return patchx, patchy, y


def get_data_for_baseline(filename, SAP, baseline_id, patch_size=32, num_channels=8, give_baseline=False,
                          uvdist=False):
    """src/lofar_tools.py:214-349: one baseline, clamp +-1e6, always z-scored."""
    assert num_channels == 4 or num_channels == 8
    dev = _device()
    f = _open(filename)
    g = f["measurement"]["saps"][SAP]["visibilities"]
    h = f["measurement"]["saps"][SAP]["visibility_scale_factors"]
    vis, sc = _load_selected(g, h, [baseline_id], dev)
    sel = torch.zeros(1, dtype=torch.int32, device=dev)
    patchx, patchy, y = patchify_device(vis, sc, sel, patch_size, num_channels, 1e6, True)
    out = [patchx, patchy, y]
    if uvdist:
        rot00, rot01 = _uv_rotation(f, SAP)
        uv = _uv_of(f, SAP, [baseline_id], rot00, rot01)
        out.append(torch.from_numpy(np.repeat(uv, patchx * patchy, axis=0)).to(dev))
    if give_baseline:
        out.insert(0, f["measurement"]["saps"][SAP]["baselines"][baseline_id])
    return tuple(out)


def get_data_for_baseline_flat(filename, SAP, baseline_id, patch_size=32, num_channels=8):
    """src/lofar_tools.py:352-406: the un-patched [1,C,T,F] spectrogram (display helper)."""
    assert num_channels == 4 or num_channels == 8
    dev = _device()
    f = _open(filename)
    g = f["measurement"]["saps"][SAP]["visibilities"]
    h = f["measurement"]["saps"][SAP]["visibility_scale_factors"]
    vis, sc = _load_selected(g, h, [baseline_id], dev)
    pols = (0, 1, 2, 3) if num_channels == 8 else (0, 3)
    chans = [vis[0, :, :, ci, ri].float() * sc[0, :, ci][None, :] for ci in pols for ri in (0, 1)]
    return torch.stack(chans)[None].clamp_(-1e6, 1e6)


def get_metadata(filename, SAP, give_baseline=False):
    """src/lofar_tools.py:410-426."""
    f = _open(filename)
    g = f["measurement"]["saps"][SAP]["visibilities"]
    if give_baseline:
        baselines = f["measurement"]["saps"][SAP]["baselines"]
        bline = np.ndarray(baselines.shape, dtype=object)
        for ci in range(baselines.shape[0]):
            bline[ci] = baselines[ci]
        return bline, g.shape
    return g.shape


def get_fileSAP(pathname, pattern="L*.MS_extract.h5"):
    """src/lofar_tools.py:430-463: (file_list, sap_list) of usable (file, SAP) pairs."""
    file_list, sap_list = [], []
    if rec_file_search:
        rawlist = glob.glob(pathname + "**" + os.sep + pattern, recursive=True)
    else:
        rawlist = glob.glob(pathname + os.sep + pattern)
    for filename in rawlist:
        f = _open(filename)
        fileused = False
        for SAP in list(f["measurement"]["saps"]):
            try:
                nbase, ntime, nfreq, npol, reim = f["measurement"]["saps"][SAP]["visibilities"].shape
                if nbase > 1 and nfreq >= 90 and ntime >= 90 and npol == 4 and reim == 2:
                    file_list.append(filename)
                    sap_list.append(SAP)
                    fileused = True
            except Exception:
                print("Failed opening" + filename)
        if not fileused:
            print("File " + filename + " not used")
    return file_list, sap_list
