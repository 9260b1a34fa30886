"""Multi-GPU plumbing for the batch-sharded path (one process per GPU, torch.distributed).

The eval forward has NO data-path collective (samples are independent: BN running stats, CBAM
pools per sample -- SURVEY 8e): ranks own disjoint slices of the batch and a full copy of the
16 MB of weights.  torch.distributed (NCCL on GPUs, gloo in the CPU tests) is used only for
rendezvous, barriers and the max-over-ranks timing reduction.  The one real exchange step of
the SmaAt-UNet path -- the training gradient all-reduce -- goes through `allreduce_flat_`.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def init_from_env(backend=None, device=None):
    """Initialise the default process group from torchrun's env (no-op for world size 1). Returns (rank, world, local)."""
    rank, world, local = env_rank_world()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        kw = {"device_id": device} if (backend == "nccl" and device is not None) else {}
        dist.init_process_group(backend, rank=rank, world_size=world, **kw)
    return rank, world, local


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous, balanced [lo, hi) slice of n_items for this rank (first n % world ranks get one extra)."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def barrier(device=None):
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        if device is not None and device.type == "cuda":
            dist.barrier(device_ids=[device.index])
        else:
            dist.barrier()
    if device is not None and device.type == "cuda":
        torch.cuda.synchronize(device)


def reduce_max(value: float, device=None) -> float:
    """max over ranks of a host scalar (timings are reported as the slowest rank's)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def allreduce_flat_(tensors, average=True):
    """One all-reduce of a flat bucket holding all `tensors` (training gradients: 4 033 537 fp32 = 16 MB for
    SmaAt-UNet, a single bucket -- SURVEY 5).  In place; returns the number of elements reduced."""
    tensors = [t for t in tensors if t is not None]
    if not tensors:
        return 0
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return sum(t.numel() for t in tensors)
    flat = torch.cat([t.reshape(-1) for t in tensors])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    if average:
        flat /= dist.get_world_size()
    off = 0
    for t in tensors:
        n = t.numel()
        t.copy_(flat[off:off + n].view_as(t))
        off += n
    return off
