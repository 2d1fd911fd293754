"""Multi-GPU: the hot path shards on the batch / image dimension.

One process per GPU (torchrun), each rank runs the op on its own images, then EXACTLY ONE
collective: an all-gather of the per-shard outputs over NCCL (NVLink 5 / NVSwitch).  No reduce, no
all-to-all, nothing else.  Equal-size outputs (resize, roi_align, roi_pool, ps_roi_align,
deform_conv2d) use all_gather_into_tensor directly; NMS keep-lists (data-dependent length) are
padded to a static capacity with their length in slot 0, so it is still a single all-gather.

The helpers are backend-agnostic (nccl on GPUs; gloo in the CPU tests of the plumbing).
"""
from __future__ import annotations

from typing import Callable, Optional, Sequence

import torch
import torch.distributed as dist


def shard_bounds(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced split of n units: the first (n % world) ranks get one extra."""
    base, rem = divmod(n, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def _world(group=None) -> tuple[int, int]:
    if not dist.is_available() or not dist.is_initialized():
        return 0, 1
    return dist.get_rank(group), dist.get_world_size(group)


def all_gather_equal(local: torch.Tensor, group=None) -> torch.Tensor:
    """All ranks hold the same shape: returns cat over ranks along dim 0 (one collective)."""
    rank, world = _world(group)
    if world == 1:
        return local
    local = local.contiguous()
    out = torch.empty((world * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    if dist.get_backend(group) == "nccl":
        dist.all_gather_into_tensor(out, local, group=group)
    else:
        dist.all_gather(list(out.chunk(world, dim=0)), local, group=group)
    return out


def all_gather_varlen(local: torch.Tensor, capacity: int, group=None) -> list[torch.Tensor]:
    """1-D int64 lists of data-dependent length <= capacity: one padded all-gather.
    Returns the per-rank lists (rank order)."""
    rank, world = _world(group)
    assert local.dim() == 1 and local.dtype == torch.int64 and local.numel() <= capacity
    if world == 1:
        return [local]
    buf = torch.full((capacity + 1,), -1, dtype=torch.int64, device=local.device)
    buf[0] = local.numel()
    buf[1:1 + local.numel()] = local
    gathered = all_gather_equal(buf.unsqueeze(0), group=group)      # [world, capacity + 1]
    lens = gathered[:, 0].tolist()
    return [gathered[r, 1:1 + int(lens[r])] for r in range(world)]


def sharded_apply(fn: Callable[..., torch.Tensor], local_inputs: Sequence, group=None) -> torch.Tensor:
    """Run `fn(*local_inputs)` on this rank's shard and all-gather the equal-shaped outputs."""
    return all_gather_equal(fn(*local_inputs), group=group)


def sharded_batched_nms(fn: Callable[..., torch.Tensor], problems: Sequence[tuple], iou_threshold: float,
                        capacity: Optional[int] = None, group=None) -> list[list[torch.Tensor]]:
    """`problems` = this rank's images, each (boxes, scores, idxs).  A single image's NMS is never split
    (greedy dependency); images are the shard unit.  Returns keep lists for every rank's images.
    `capacity` = static upper bound, IDENTICAL on every rank, of (boxes + images) per rank; defaults to
    this rank's own total, which is only valid when all ranks hold equally many boxes (weak scaling)."""
    rank, world = _world(group)
    keeps = [fn(b, s, i, iou_threshold) for (b, s, i) in problems]
    if world == 1:
        return [keeps]
    cap = capacity if capacity is not None else sum(int(b.shape[0]) + 1 for (b, _, _) in problems)
    device = problems[0][0].device if problems else torch.device("cpu")
    packed = torch.cat([torch.cat([torch.tensor([k.numel()], dtype=torch.int64, device=device), k]) for k in keeps]) \
        if keeps else torch.empty(0, dtype=torch.int64, device=device)
    per_rank = all_gather_varlen(packed, cap, group=group)
    out = []
    for flat in per_rank:
        lst, pos = [], 0
        while pos < flat.numel():
            n = int(flat[pos])
            lst.append(flat[pos + 1:pos + 1 + n])
            pos += 1 + n
        out.append(lst)
    return out
