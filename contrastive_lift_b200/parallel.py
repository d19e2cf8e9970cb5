"""Data-parallel plumbing for the render path: one process per GPU, parameters replicated,
rays sharded, gradients mean-all-reduced once per backward over a flat fp32 arena.

This reproduces what the reference gets from Lightning's DDPStrategy (trainer/__init__.py:95-108): each rank renders
its own rays, every ``manual_backward`` (trainer:198, :220) is followed by a mean of the gradients over ranks.  The
path has no other exchange step (rays are independent; the slow-fast loss keeps whole images on one rank, as DDP's
DistributedSampler does), so torch.distributed's NCCL all-reduce over NVLink is the only collective.
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [begin, end) of ``n`` rays owned by ``rank`` (ray order is preserved across ranks)."""
    base, rem = divmod(n, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def shard_rays(rays: torch.Tensor, rank: Optional[int] = None, world: Optional[int] = None) -> torch.Tensor:
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    b, e = shard_range(rays.shape[0], rank, world)
    return rays[b:e]


def gather_rays_output(local: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """Inverse of shard_rays for an output map [n_local, K]: all-gather with the shard sizes of shard_range().
    Shards differ by at most one ray; the collective itself runs on equal-size buffers (gloo refuses uneven all_gather
    lists, NCCL falls back to one broadcast per rank for them), the padding row is dropped afterwards."""
    world = dist.get_world_size(group)
    sizes = [e - b for b, e in (shard_range(n_total, r, world) for r in range(world))]
    rows = max(sizes) if sizes else 0
    if local.shape[0] != sizes[dist.get_rank(group)]:
        raise ValueError(f"gather_rays_output: this rank holds {local.shape[0]} rays, shard_range() assigns it "
                         f"{sizes[dist.get_rank(group)]} of {n_total}")
    send = local.contiguous()
    if send.shape[0] < rows:
        send = torch.cat([send, send.new_zeros((rows - send.shape[0],) + tuple(send.shape[1:]))], 0)
    parts = [send.new_empty(send.shape) for _ in range(world)]
    dist.all_gather(parts, send, group=group)
    return torch.cat([p[:n] for p, n in zip(parts, sizes)], 0)


@torch.no_grad()
def allreduce_gradients(params: Iterable[torch.nn.Parameter], group=None, average: bool = True) -> int:
    """Sum (then 1/world) every parameter's ``.grad`` across ranks through ONE flat buffer / one collective.

    Parameters whose grad is None on this rank (heads not evaluated in this pass - the reference needs
    ``find_unused_parameters=True`` for the same reason) contribute zeros and stay None.  Returns the number of
    bytes reduced."""
    plist: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
    if not plist or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return 0
    dev = plist[0].device
    total = sum(p.numel() for p in plist)
    flat = torch.zeros((total,), dtype=torch.float32, device=dev)
    slots = flat.split([p.numel() for p in plist])                    # views into the arena, one per parameter
    live = [(s, p.grad) for s, p in zip(slots, plist) if p.grad is not None]
    if live:      # one multi-tensor copy in, one out (instead of a launch per parameter and direction)
        torch._foreach_copy_([s.view_as(g) for s, g in live], [g for _, g in live])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if average:
        flat.mul_(1.0 / dist.get_world_size(group))
    if live:
        torch._foreach_copy_([g for _, g in live], [s.view_as(g) for s, g in live])
    return total * 4


def broadcast_parameters(module: torch.nn.Module, src: int = 0, group=None) -> None:
    """Replicate rank ``src``'s parameters and buffers (what DDP does at construction)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    with torch.no_grad():
        for t in list(module.parameters()) + list(module.buffers()):
            dist.broadcast(t.data, src=src, group=group)
