"""Data-parallel layer: shard patches by baseline group, one all-reduce per closure evaluation.

The reference is single-process; this layer is new.  Every loss term
of /root/reference/src/kharmonic_lofar.py:154-172 is a sum over patches divided by a global
constant, so a rank that holds whole baseline groups (rows [g*bpb,(g+1)*bpb), which keeps the
augmentation groups of :101-102 and the multipliers y1..y3 rank-local) computes its partial sums
with the GLOBAL divisors; summing [all gradients | loss scalars] over ranks then reproduces the
single-process closure exactly.  The centre penalty depends only on the replicated M and is
scaled by 1/world on every rank.  No other collective is on the data path.

Host logic only (works with gloo on CPU tensors - tests/test_dp_gloo.py - and NCCL on GPUs).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_groups(n_groups: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [start, stop) of baseline groups owned by `rank` (sizes differ by <= 1)."""
    base, rem = divmod(n_groups, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_rows(n_patches: int, bpb: int, rank: int, world: int) -> Tuple[int, int]:
    """Row range of the shard in the baseline-major patch order the loss assumes."""
    if n_patches % bpb:
        raise ValueError("n_patches must be a multiple of batch_per_bline")
    g0, g1 = shard_groups(n_patches // bpb, rank, world)
    return g0 * bpb, g1 * bpb


@dataclass
class ShardPlan:
    """Divisors / scales a rank applies to its partial sums (see module docstring)."""
    n_local: int
    n_global: int
    world: int
    bpb: int
    channels: int
    K: int
    Ltot: int

    @property
    def numel_global(self) -> float:
        return float(self.n_global) * self.channels * 16384

    def khm_scale(self, alpha: float) -> float:
        return alpha / (float(self.n_global) * self.K * self.Ltot)

    def aug_scale(self, gamma: float) -> float:
        return gamma / (float(self.bpb) * (self.n_global // self.bpb) * self.bpb)

    def sim_scale(self, beta: float) -> float:
        return beta / self.world

    def rica_scale(self, lam: float) -> float:
        return lam * float(self.n_local) / float(self.n_global)


def exchange(buf: torch.Tensor, group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """THE collective of the path: sum the flat fp32 buffer [gradients | loss scalars] (or only
    the loss scalars for a forward-only closure) over the data-parallel ranks, in place."""
    dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    return buf


def global_mean_std(local_sum: torch.Tensor, local_sumsq: torch.Tensor, n_global: int,
                    group: Optional[dist.ProcessGroup] = None):
    """Loader z-score statistics (src/lofar_tools.py:190-193) when patchify is sharded: all-reduce
    (sum, sum of squares) and return (mean, unbiased std) of the global tensor."""
    st = torch.stack((local_sum.double().reshape(()), local_sumsq.double().reshape(())))
    dist.all_reduce(st, op=dist.ReduceOp.SUM, group=group)
    mean = st[0] / n_global
    var = (st[1] - st[0] * st[0] / n_global) / (n_global - 1)
    return mean, var.sqrt()
