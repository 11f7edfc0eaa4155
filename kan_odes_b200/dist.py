"""Data-parallel plumbing for the ensemble / batched-IC configs.

Trajectories (initial conditions) are independent: the batch is split contiguously over ranks (one process per GPU),
parameters are replicated, and the ONLY collective of a training step is the all-reduce of the per-rank gradient sum
(np floats), the loss sum and the trajectory count.  torch.distributed is used for that plumbing (NCCL over NVLink on
the GPU box, gloo in the CPU tests); the compute stays in libkanode_b200.so.
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def shard_bounds(batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous near-equal split of `batch` trajectories: the first batch % world ranks get one more."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, rem = divmod(batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


_COUNT_CACHE: dict = {}


def combine_loss_grad(loss_sum: torch.Tensor, grad_sum: torch.Tensor, local_count: int, nsave: int, n: int,
                      group=None, sync: bool = True):
    """All-reduce the UNNORMALISED sums of a step (what kanode_loss_grad_dev returns) and normalise once, globally:
        loss = sum_b sum_{s,i} (pred - X)^2 / (B * nsave * n)     (mean(abs2, ...) over the whole ensemble)
        grad = sum_b g_b(t0) / B
    Works for any world size (1 included) and for ragged shards.  grad_sum is reduced in place; the loss sum and the
    trajectory count travel in ONE 2-element float64 all-reduce (two collectives per step in total).
    `sync=False` keeps everything on the device (no host read-back inside a training / timing loop): the third return value
    is then the global count as a 1-element tensor instead of an int."""
    key = (loss_sum.device, int(local_count))
    cnt = _COUNT_CACHE.get(key)
    if cnt is None:                                            # built once per (device, shard size): no per-step host-to-device copy
        cnt = _COUNT_CACHE[key] = torch.tensor([float(local_count)], dtype=torch.float64, device=loss_sum.device)
    pair = torch.cat([loss_sum.reshape(1).to(torch.float64), cnt])
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(grad_sum, group=group)
        dist.all_reduce(pair, group=group)
    count = pair[1:2]
    loss = pair[0:1] / (count * (nsave * n))
    grad = grad_sum / count.to(grad_sum.dtype)
    if not sync:
        return loss, grad, count
    total = int(round(count.item()))
    if total == 0:
        raise ValueError("empty global batch")
    return loss, grad, total
