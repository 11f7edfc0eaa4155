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


_PACK_CACHE: dict = {}
PACK_MAX = 1 << 16      # gradients up to this many entries travel with the loss sum and the count in ONE fp64 all-reduce
PEER_MAX_ENTRIES = 4096  # KANODE_PEER_MAX_ENTRIES: [gradient | loss | count] up to this size go through the library's own peer-memory all-reduce
IPC_HANDLE_BYTES = 64


def peer_setup(ode, group=None) -> bool:
    """Connect the handles of all ranks for the library's own all-reduce over NVLink peer memory (kanode_peer.cu): every rank
    exports the CUDA IPC handle of its mailbox, the 64-byte handles are all-gathered through torch.distributed (host objects:
    works on any backend), and every rank maps its peers' mailboxes.  Returns False (and leaves the NCCL path in place) when the
    model is too large for the mailbox or a GPU cannot export / map a mailbox.  Every rank takes part in every gather whatever
    happened locally, and the decision is collective: either all ranks use the peer path or none does."""
    import ctypes as C
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank(group) if world > 1 else 0

    def everyone(flag: bool) -> bool:
        if world == 1:
            return flag
        flags = [None] * world
        dist.all_gather_object(flags, bool(flag), group=group)
        return all(flags)

    mine = (C.c_ubyte * IPC_HANDLE_BYTES)()
    ok = ode.np_ + 2 <= PEER_MAX_ENTRIES and world <= 16 and ode.lib.kanode_peer_export(ode.h, mine) == 0
    handles = [bytes(mine)]
    if world > 1:
        handles = [None] * world
        dist.all_gather_object(handles, bytes(mine), group=group)
    if not everyone(ok):                                       # some rank could not export: nobody attaches
        ode._peer_ok = False
        return False
    blob = (C.c_ubyte * (IPC_HANDLE_BYTES * world)).from_buffer_copy(b"".join(handles))
    ok = everyone(ode.lib.kanode_peer_attach(ode.h, rank, world, blob) == 0)
    ode._peer_ok = ok
    return ok


def peer_all_reduce(ode, loss_sum: torch.Tensor, grad_sum: torch.Tensor, local_count: int) -> torch.Tensor:
    """ONE kernel per step and rank: packs [gradient sum | loss sum | count], stores it into every rank's mailbox over NVLink,
    waits for the world and sums in rank order (kanode_pack_allreduce_dev).  Same result layout as packed_all_reduce."""
    npar = grad_sum.numel()
    key = ("peer", loss_sum.device, npar)
    buf = _PACK_CACHE.get(key)
    if buf is None:
        buf = _PACK_CACHE[key] = torch.zeros(npar + 2, dtype=torch.float64, device=loss_sum.device)
    f = ode.lib.kanode_pack_allreduce_dev_f64 if grad_sum.dtype == torch.float64 else ode.lib.kanode_pack_allreduce_dev
    rc = f(ode.h, grad_sum.data_ptr(), loss_sum.data_ptr(), int(local_count), buf.data_ptr())
    if rc != 0:
        from . import abi
        abi.check(ode.lib, ode.h, rc, "kanode_pack_allreduce_dev")
    return buf


def packed_all_reduce(ode, loss_sum: torch.Tensor, grad_sum: torch.Tensor, local_count: int, group=None) -> torch.Tensor:
    """The data-parallel step on the device path: kanode_pack_sums_dev writes [gradient sum | loss sum | count] into ONE persistent
    fp64 buffer with one kernel of the library (on the handle's stream = torch's current stream in the callers), then ONE
    all-reduce.  Returns the reduced buffer [np + 2]; nothing is normalised or read back here (kanode_train_apply_packed_dev
    consumes it on the device; `unpack` normalises for a host-side consumer)."""
    import ctypes as C
    if getattr(ode, "_peer_ok", False):                       # peer_setup succeeded on every rank: no NCCL call in the step
        return peer_all_reduce(ode, loss_sum, grad_sum, local_count)
    npar = grad_sum.numel()
    key = ("packed", loss_sum.device, npar)
    buf = _PACK_CACHE.get(key)
    if buf is None:
        buf = _PACK_CACHE[key] = torch.zeros(npar + 2, dtype=torch.float64, device=loss_sum.device)
    f = ode.lib.kanode_pack_sums_dev_f64 if grad_sum.dtype == torch.float64 else ode.lib.kanode_pack_sums_dev
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
    rc = f(ode.h, grad_sum.data_ptr(), loss_sum.data_ptr(), int(local_count), buf.data_ptr())
    if rc != 0:
        raise RuntimeError(f"kanode_pack_sums_dev failed ({rc})")
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(buf, group=group)
    return buf


def unpack(packed: torch.Tensor, nsave: int, n: int, dtype=torch.float32):
    """(loss, grad, count) of a reduced packed buffer: loss = sum / (B nsave n), grad = sum / B."""
    npar = packed.numel() - 2
    count = packed[npar + 1:npar + 2]
    return packed[npar:npar + 1] / (count * (nsave * n)), (packed[:npar] / count).to(dtype), count


def combine_loss_grad(loss_sum: torch.Tensor, grad_sum: torch.Tensor, local_count: int, nsave: int, n: int,
                      group=None, sync: bool = True):
    """All-reduce the UNNORMALISED sums of a step (what kanode_loss_grad_dev returns) and normalise once, globally:
        loss = sum_b sum_{s,i} (pred - X)^2 / (B * nsave * n)     (mean(abs2, ...) over the whole ensemble)
        grad = sum_b g_b(t0) / B
    Works for any world size (1 included) and for ragged shards.  ONE collective per step for small models: the gradient sum
    (np <= PACK_MAX entries), the loss sum and the trajectory count are packed into one persistent fp64 buffer [np + 2]
    (at 8 ranks the step time of the LV ensemble is the latency of the rendezvous, not the bytes: a second tiny all-reduce
    costs as much as the first).  Larger gradients (the PDE surrogates) are reduced in their own dtype, in place, and the
    2-element fp64 pair follows in a second collective (noise next to their step time).
    `sync=False` keeps everything on the device (no host read-back inside a training / timing loop): the third return value
    is then the global count as a 1-element tensor instead of an int."""
    npar = grad_sum.numel()
    multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    key = (loss_sum.device, int(local_count), npar if npar <= PACK_MAX else 0)
    buf = _PACK_CACHE.get(key)
    if buf is None:                                            # built once per (device, shard size, np): no per-step allocation or H2D copy
        buf = _PACK_CACHE[key] = torch.zeros((npar if npar <= PACK_MAX else 0) + 2, dtype=torch.float64, device=loss_sum.device)
    if npar <= PACK_MAX:
        buf[:npar].copy_(grad_sum.reshape(-1))
        buf[npar:npar + 1].copy_(loss_sum.reshape(1))
        buf[npar + 1] = float(local_count)
        if multi:
            dist.all_reduce(buf, group=group)
        count = buf[npar + 1:npar + 2]
        loss = buf[npar:npar + 1] / (count * (nsave * n))
        grad = (buf[:npar] / count).to(grad_sum.dtype).reshape(grad_sum.shape)
    else:
        buf[0:1].copy_(loss_sum.reshape(1)); buf[1] = float(local_count)
        if multi:
            dist.all_reduce(grad_sum, group=group)
            dist.all_reduce(buf, group=group)
        count = buf[1:2]
        loss = buf[0:1] / (count * (nsave * n))
        grad = grad_sum / count.to(grad_sum.dtype)
    if not sync:
        return loss, grad, count
    total = int(round(count.item()))
    if total == 0:
        raise ValueError("empty global batch")
    return loss, grad, total
