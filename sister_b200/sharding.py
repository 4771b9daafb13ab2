"""Frame sharding of a batch of rigs over the GPUs of one box.

Rigs are independent (compute_disparities keeps no state between calls, hpp:26-119), so the batch is split into
contiguous blocks, one per rank, every rank runs its block through its own Engine with no data-path collective, and
the only communication is the gather of the H x W uint16 maps to one rank when the caller wants them in one place.
torch.distributed is the plumbing (NCCL on the GPU box, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Callable, List, Sequence, Tuple

import numpy as np


def frame_shard(n_rigs: int, world: int, rank: int) -> Tuple[int, int]:
    """[begin, end) of the rigs rank `rank` owns: contiguous blocks, the first n_rigs % world ranks get one extra."""
    if world <= 0 or not (0 <= rank < world) or n_rigs < 0:
        raise ValueError("bad shard request")
    base, extra = divmod(n_rigs, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def shard_counts(n_rigs: int, world: int) -> List[int]:
    return [frame_shard(n_rigs, world, r)[1] - frame_shard(n_rigs, world, r)[0] for r in range(world)]


def gather_maps(local_maps, n_rigs: int, dst: int = 0, group=None, async_op: bool = False):
    """Gather every rank's [n_local, H, W] maps (torch int16 tensor holding the uint16 bit patterns; NCCL has no u16)
    to rank `dst` in rig order. Returns the [n_rigs, H, W] tensor on dst, None elsewhere. Blocks are padded to the
    largest shard so a single collective moves everything.
    async_op: start the collective and return a function that completes it (and returns the same result), so that the
    caller can compute the next batch while the maps travel."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if world == 1:
        return (lambda: local_maps) if async_op else local_maps
    counts = shard_counts(n_rigs, world)
    cap = max(counts)
    h, w = local_maps.shape[1:]
    send = local_maps
    if local_maps.shape[0] < cap:
        send = torch.zeros((cap, h, w), dtype=local_maps.dtype, device=local_maps.device)
        send[: local_maps.shape[0]] = local_maps
    # the collective moves bytes: neither NCCL nor gloo carries 16-bit integers on every op
    send = send.contiguous().view(torch.uint8)
    recv = [torch.empty_like(send) for _ in range(world)] if rank == dst else None
    work = dist.gather(send, recv, dst=dst, group=group, async_op=async_op)

    def finish():
        if work is not None:
            work.wait()
        if rank != dst:
            return None
        return torch.cat([recv[r][: counts[r]] for r in range(world)], dim=0).view(local_maps.dtype)

    return finish if async_op else finish()


def compute_sharded(compute_block: Callable[[Sequence[int]], "object"], n_rigs: int, gather_to: int | None = 0, group=None):
    """Run `compute_block(rig_indices) -> [n_local, H, W] int16 tensor` on this rank's shard of `n_rigs` rigs and, if
    gather_to is not None, gather the maps to that rank. `compute_block` is where the Engine runs on the GPU box."""
    import torch.distributed as dist

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    begin, end = frame_shard(n_rigs, world, rank)
    local = compute_block(range(begin, end))
    if gather_to is None:
        return local
    return gather_maps(local, n_rigs, dst=gather_to, group=group)


def as_int16(maps_u16: np.ndarray):
    """uint16 numpy maps -> torch int16 tensor with the same bits."""
    import torch

    return torch.from_numpy(np.ascontiguousarray(maps_u16).view(np.int16))
