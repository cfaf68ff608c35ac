"""Inference / assignment loop of /root/reference/src/evaluate_clustering.py:75-119.

Per baseline: cascade forward -> Mu = cat(mu, muT, muF) -> kdist = mod(Mu) ->
dist[k] = mean_n ||Mu_n - M_k||^p -> argmin_k.  The t-SNE / agglomerative / PNG part
(:121-163) is CPU post-processing on the returned ``X [K,nbase]`` and is out of scope.

`encode_assign` is the batched form used for many baselines at once (BASELINE config 3):
the same arithmetic on all patches of a shard, then one grouped reduction.
"""
from __future__ import annotations

import numpy as np
import torch

from ._lib import lib
from .lofar_models import AutoEncoder1DCNN, AutoEncoderCNN2, Kmeans


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


_SIDE = {}


def _side_stream(device) -> torch.cuda.Stream:
    key = torch.device(device).index
    if key not in _SIDE:
        _SIDE[key] = torch.cuda.Stream(device)
    return _SIDE[key]


@torch.no_grad()
def cascade_latents(net: AutoEncoderCNN2, net1D1: AutoEncoder1DCNN, net1D2: AutoEncoder1DCNN,
                    x: torch.Tensor, uv: torch.Tensor) -> torch.Tensor:
    """src/evaluate_clustering.py:81-89,108: Mu [N, L+2Lt] for patches x [N,C,128,128]."""
    N, C = x.shape[:2]
    L, Lt = net.latent_dim, net1D1.latent_dim
    st = _stream()
    x = x.contiguous()
    uv = uv.contiguous()
    Mu = torch.empty(N, L + 2 * Lt, dtype=torch.float32, device=x.device)
    scales = net.harmonic_scales.to(x.device).float().contiguous()
    e0, e1, e2 = net.engine(), net1D1.engine(), net1D2.engine()
    ws0 = e0.workspace(N, x.device, False)
    x1, _ = e0.forward(x.view(N, -1), uv, scales, net.named_param_dict(), ws0, st, mu_out=Mu[:, :L])
    iy1 = torch.empty(N, C * 16384, dtype=torch.float32, device=x.device)
    iy2 = torch.empty_like(iy1)
    lib().residual_split(x.data_ptr(), x1.data_ptr(), iy1.data_ptr(), iy2.data_ptr(), N, C, 128, st)
    del ws0
    # the two 1-D encoders are independent: second stream for the frequency-axis net (fork / join by events)
    ws1 = e1.workspace(N, x.device, False)
    ws2 = e2.workspace(N, x.device, False)
    cur = torch.cuda.current_stream(x.device)
    side = _side_stream(x.device)
    ev = torch.cuda.Event(); ev.record(cur); side.wait_event(ev)
    with torch.cuda.stream(side):
        e2.forward(iy2, uv, scales, net1D2.named_param_dict(), ws2, side.cuda_stream, mu_out=Mu[:, L + Lt:], decode=False)
        ev2 = torch.cuda.Event(); ev2.record(side)
    e1.forward(iy1, uv, scales, net1D1.named_param_dict(), ws1, st, mu_out=Mu[:, L:L + Lt], decode=False)
    cur.wait_event(ev2)
    iy2.record_stream(side)         # allocated on the current stream, last used on the side stream
    for v in vars(ws2).values():
        for t in (v if isinstance(v, (list, tuple)) else (v,)):
            if isinstance(t, torch.Tensor):
                t.record_stream(side)
    return Mu


@torch.no_grad()
def encode_assign(net, net1D1, net1D2, mod: Kmeans, x, uv, batch_per_bline: int):
    """Returns (dist [nbase,K] fp32, baseline cluster id [nbase] int32, per-patch id [N] int32,
    Mu [N,Ltot])."""
    Mu = cascade_latents(net, net1D1, net1D2, x, uv)
    dist, gid = mod.group_distances(Mu, batch_per_bline)
    return dist, gid, mod.assign(Mu), Mu


@torch.no_grad()
def graph_features(net, net1D1, net1D2, mod: Kmeans, x, uv, batch_per_bline: int):
    """Node features / labels of the graph classifiers (src/train_graph.py:137-158, src/train_graph_stat.py:
    197-218) for many baselines at once: node_data[g] = mean of the baseline's latents, node_label[g,k] = mean
    Euclidean distance of its patches to centre k.  x holds whole baselines, `batch_per_bline` patches each
    (rows of one baseline consecutive)."""
    Mu = cascade_latents(net, net1D1, net1D2, x, uv)
    G = Mu.shape[0] // batch_per_bline
    node_data = Mu.view(G, batch_per_bline, -1).mean(dim=1)
    node_label, _ = mod.group_distances(Mu, batch_per_bline, power=1.0)
    return node_data, node_label


@torch.no_grad()
def graph_stat_features(net, net1D1, net1D2, mod: Kmeans, x, uv):
    """Station-graph features of src/train_graph_stat.py:197-218 for every patch given (the script picks ONE random
    patch per baseline, :191-194, and calls the nets on it): attr[n] = Mu_n (node attribute of an autocorrelation,
    edge attribute of a cross-correlation) and label[n] = softmax_k(-d_nk / mean_k d_nk) with
    d_nk = ||Mu_n - M_k||^p (:207-210, 'smallest dist ~ highest prob').  The distances come from the grouped
    distance kernel with groups of one patch; the K-wide softmax is a torch op on the [N,K] result."""
    Mu = cascade_latents(net, net1D1, net1D2, x, uv)
    dist, _ = mod.group_distances(Mu, 1)
    label = torch.softmax(-dist / dist.mean(dim=1, keepdim=True), dim=1)
    return Mu, label


@torch.no_grad()
def evaluate(net, net1D1, net1D2, mod: Kmeans, baseline_loader, nbase: int, log=None):
    """The per-baseline loop of src/evaluate_clustering.py:75-119.

    baseline_loader(nb) -> (patchx, patchy, x, uv) (e.g. a partial of get_data_for_baseline with
    uvdist=True).  Returns (X [K,nbase] float64, clusid [nbase] float64) like the reference.
    """
    X = np.zeros([mod.K, nbase], dtype=np.float64)
    clusid = np.zeros(nbase, dtype=np.float64)
    for nb in range(nbase):
        patchx, patchy, x, uv = baseline_loader(nb)
        Mu = cascade_latents(net, net1D1, net1D2, x, uv)
        kdist = mod(Mu)
        dist, gid = mod.group_distances(Mu, Mu.shape[0])
        X[:, nb] = dist[0].double().cpu().numpy()
        clusid[nb] = int(gid[0])
        if log is not None:
            log("%d %e %d" % (nb, float(kdist), int(gid[0])))
    return X, clusid
