"""Data-parallel plumbing for the head: flat gradient bucket + one all-reduce per step.

The reference has no distributed code at all (DistributedDataParallel is imported at train.py:18 and never
used); batch sharding is new work.  The head's samples are independent in forward and backward, so the only
exchange is the sum of parameter gradients.  Device-agnostic on purpose: the world_size-2 ``gloo`` tests run
this exact code on CPU tensors, the GPU path runs it over NCCL / NVLink.
"""
from __future__ import annotations

from typing import Iterable, List

import torch
import torch.distributed as dist


class FlatGradBucket:
    """One contiguous fp32 buffer whose slices are the ``.grad`` of the given parameters."""

    def __init__(self, params: Iterable[torch.nn.Parameter]):
        self.params: List[torch.nn.Parameter] = list(params)
        if not self.params:
            raise ValueError("FlatGradBucket needs at least one parameter")
        dev, dt = self.params[0].device, self.params[0].dtype
        sizes = [p.numel() for p in self.params]
        self.flat = torch.zeros(sum(sizes), device=dev, dtype=dt)
        self.views: List[torch.Tensor] = []
        off = 0
        for p, n in zip(self.params, sizes):
            v = self.flat[off: off + n].view_as(p)
            p.grad = v
            self.views.append(v)
            off += n

    def nbytes(self) -> int:
        return self.flat.numel() * self.flat.element_size()

    def all_reduce(self, group=None, async_op: bool = False):
        """Sum the bucket over the process group (no-op for a single process)."""
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return None
        return dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op)


def world_size(group=None) -> int:
    return dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1


def broadcast_parameters(params: Iterable[torch.Tensor], src: int = 0, group=None):
    """Every rank starts from rank ``src``'s weights (each process would otherwise draw its own init)."""
    if world_size(group) == 1:
        return
    for p in params:
        dist.broadcast(p.data if isinstance(p, torch.nn.Parameter) else p, src=src, group=group)


def broadcast_mask(mask, src: int = 0, group=None, device="cpu"):
    """Share rank ``src``'s host-drawn token mask so all shards see the same indices (parity with one big batch)."""
    if world_size(group) == 1:
        return list(mask)
    t = torch.tensor(list(mask), dtype=torch.int32, device=device)
    dist.broadcast(t, src=src, group=group)
    return t.cpu().tolist()


def shard_batch(global_batch: int, rank: int, world: int):
    """Contiguous batch shard [lo, hi) of rank ``rank``; remainders go to the lowest ranks."""
    base, rem = divmod(global_batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)
