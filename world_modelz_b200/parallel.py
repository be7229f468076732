"""Data-parallel plumbing: batch sharding and the single gradient exchange.

One process per GPU (``torchrun``), NCCL over NVLink on the device, gloo in the CPU tests.
The path shards by batch (clips are independent): the only collective is one all-reduce
of the flat gradient buffer per training step; sampling has none.
"""
from __future__ import annotations

import os
from typing import Tuple

import torch
import torch.distributed as dist


def init_from_env(backend: str = 'nccl') -> Tuple[int, int, int]:
    """Initialise the default process group from torchrun's environment.

    Returns ``(rank, local_rank, world_size)``; a no-op single-process setup when
    ``WORLD_SIZE`` is unset or 1.
    """
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        os.environ.setdefault('MASTER_PORT', '29500')
        if backend == 'nccl':
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend, rank=rank, world_size=world,
                                    device_id=torch.device('cuda', local_rank))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, local_rank, world


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced ``[begin, end)`` slice of ``total`` clips owned by ``rank``."""
    base, extra = divmod(total, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def allreduce_mean_(flat: torch.Tensor, group=None) -> torch.Tensor:
    """In-place average of a flat buffer over the group (SUM all-reduce, then 1/world)."""
    if dist.is_available() and dist.is_initialized():
        world = dist.get_world_size(group)
        if world > 1:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
            flat.mul_(1.0 / world)
    return flat
