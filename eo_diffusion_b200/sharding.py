"""Batch sharding of the sampling path across the GPUs of one box.

Samples are independent (GroupNorm statistics are per sample, attention is per sample and
head, the sampler is elementwise), so the only multi-GPU structure the path has is: give every
rank a contiguous slice of the batch and, after the loop, gather the finished images
(SURVEY.md section 8e).  One process per GPU; `torch.distributed` (NCCL over NVLink on the GPU
box, gloo in the CPU tests) carries the single collective.  The reference itself is
single-device (inference.py:55), so there is no reference file to cite for this module."""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist

__all__ = ["shard_bounds", "shard_batch", "gather_samples", "sample_sharded"]


def shard_bounds(n: int, rank: int, world: int) -> Tuple[int, int]:
    """[lo, hi) of rank's slice of n samples: contiguous, sizes differ by at most one, earlier
    ranks take the remainder."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(t: Optional[torch.Tensor], rank: int, world: int) -> Optional[torch.Tensor]:
    if t is None:
        return None
    lo, hi = shard_bounds(t.shape[0], rank, world)
    return t[lo:hi]


def gather_samples(local: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """all_gather of per-rank sample slices (possibly ragged by one) -> [n_total, ...] on every
    rank, in rank order."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local
    rank = dist.get_rank(group)
    sizes = [shard_bounds(n_total, r, world) for r in range(world)]
    cap = max(hi - lo for lo, hi in sizes)
    rest = tuple(local.shape[1:])
    if all(hi - lo == cap for lo, hi in sizes):
        out = torch.empty((n_total,) + rest, dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    pad = torch.zeros((cap,) + rest, dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([p[: hi - lo] for p, (lo, hi) in zip(parts, sizes)], 0)


def sample_sharded(sample_fn: Callable[[int, Optional[torch.Tensor], Optional[torch.Tensor]], torch.Tensor],
                   n_samples: int, cond: Optional[torch.Tensor] = None,
                   y: Optional[torch.Tensor] = None, group=None) -> torch.Tensor:
    """Run `sample_fn(n_local, cond_local, y_local)` (e.g. a closure over
    `EODiffusion.sampling`) on this rank's slice and return all `n_samples` images on every
    rank.  With a shared pre-drawn noise tape sliced the same way the result equals the
    single-process result."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    lo, hi = shard_bounds(n_samples, rank, world)
    c = None if cond is None else cond[:n_samples][lo:hi]
    yy = None if y is None else y[:n_samples][lo:hi]
    local = sample_fn(hi - lo, c, yy)
    return gather_samples(local, n_samples, group)
