"""Synthetic LOFAR-shaped observation generator.

The real inputs are 4 GB HDF5 extracts that are not in the tree
(/root/reference/Demo.ipynb:43).  This builds an in-memory measurement with the same
group layout the reference loader walks (src/lofar_tools.py:76-110):

    measurement/saps/<SAP>/visibilities              int8  [nbase,T,F,4,2]
    measurement/saps/<SAP>/visibility_scale_factors  fp32  [nbase,F,4]
    measurement/saps/<SAP>/central_frequencies       fp64  [F]
    measurement/saps/<SAP>/baselines                 str   [nbase,2]
    measurement/saps/<SAP>/antenna_locations/XYZ/<station>  fp64 [3]
    measurement/info/start_time                      [b'<date> hh:mm:ss']

numpy only: the generator is host-side test/bench plumbing.
"""
from __future__ import annotations

import numpy as np


def make_measurement(nbase: int, ntime: int, nfreq: int, seed: int = 0, sap: str = "0",
                     nstations: int = 0, start_time: str = "12:30:00", f0: float = 140e6) -> dict:
    rng = np.random.default_rng(seed)
    vis = rng.integers(-128, 128, size=(nbase, ntime, nfreq, 4, 2), dtype=np.int16)
    # low-rank fringe so patches are not pure noise
    t = np.arange(ntime, dtype=np.float32)[None, :, None]
    f = np.arange(nfreq, dtype=np.float32)[None, None, :]
    rate = rng.uniform(0.01, 0.2, size=(nbase, 1, 1)).astype(np.float32)
    delay = rng.uniform(0.01, 0.2, size=(nbase, 1, 1)).astype(np.float32)
    phase = rate * t + delay * f
    for pol in range(4):
        amp = 40.0 if pol in (0, 3) else 8.0
        vis[..., pol, 0] = vis[..., pol, 0] // 4 + (amp * np.cos(phase)).astype(np.int16)
        vis[..., pol, 1] = vis[..., pol, 1] // 4 + (amp * np.sin(phase)).astype(np.int16)
    vis = np.clip(vis, -128, 127).astype(np.int8)
    scale = (1.0 - rng.random(size=(nbase, nfreq, 4), dtype=np.float32)).astype(np.float32)  # (0,1]
    if nstations <= 0:
        nstations = max(2, int(np.ceil((1 + np.sqrt(1 + 8 * nbase)) / 2)))
    names = [f"CS{idx:03d}" for idx in range(nstations)]
    xyz = {nm: rng.normal(0.0, 500.0, size=3) for nm in names}
    pairs = [(a, b) for a in range(nstations) for b in range(a + 1, nstations)]
    while len(pairs) < nbase:
        pairs = pairs + pairs
    baselines = np.array([[names[a], names[b]] for a, b in pairs[:nbase]], dtype=object)
    freqs = f0 + (np.arange(nfreq) - nfreq // 2) * 195312.5 / 64
    return {
        "measurement": {
            "saps": {
                sap: {
                    "visibilities": vis,
                    "visibility_scale_factors": scale,
                    "central_frequencies": freqs.astype(np.float64),
                    "baselines": baselines,
                    "antenna_locations": {"XYZ": xyz},
                }
            },
            "info": {"start_time": [("2020-01-01 " + start_time).encode("ascii")]},
        }
    }


def make_patches(n: int, channels: int = 8, seed: int = 0, patch: int = 128) -> np.ndarray:
    """N(0,1) fp32 patches [n,C,P,P] plus a per-patch fringe; already-normalised scale."""
    rng = np.random.default_rng(seed)
    x = rng.standard_normal(size=(n, channels, patch, patch), dtype=np.float32)
    t = np.arange(patch, dtype=np.float32)[None, None, :, None]
    f = np.arange(patch, dtype=np.float32)[None, None, None, :]
    rate = rng.uniform(0.02, 0.3, size=(n, 1, 1, 1)).astype(np.float32)
    delay = rng.uniform(0.02, 0.3, size=(n, 1, 1, 1)).astype(np.float32)
    x += 0.7 * np.cos(rate * t + delay * f + np.arange(channels, dtype=np.float32)[None, :, None, None])
    return x


def make_uv(n: int, seed: int = 0, per_group: int = 1) -> np.ndarray:
    """(u,v) in wavelengths, one draw per baseline group, repeated ``per_group`` times."""
    rng = np.random.default_rng(seed + 7919)
    g = (n + per_group - 1) // per_group
    uv = rng.normal(0.0, 300.0, size=(g, 2)).astype(np.float32)
    return np.repeat(uv, per_group, axis=0)[:n].copy()
