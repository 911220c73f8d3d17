"""Multi-GPU plumbing (one process per GPU, torch.distributed; NCCL on GPUs, gloo in CPU tests).

The path shards over views: the view-independent stage runs once per scene on
`src`, its output -- encoder tokens plus the hoisted per-layer decoder K (pre-RoPE) and V^T --
is broadcast, every rank renders a contiguous slice of the views, images are gathered."""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.distributed as dist

from .engine import SceneState


def view_slice(n_views: int, world: int, rank: int) -> slice:
    """Contiguous, balanced partition of `n_views` over `world` ranks."""
    base, extra = divmod(n_views, world)
    start = rank * base + min(rank, extra)
    return slice(start, start + base + (1 if rank < extra else 0))


def broadcast_scene_state(state: SceneState, src: int = 0) -> SceneState:
    """In-place broadcast of every tensor of a SceneState (receivers pass alloc_scene_state())."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        for t in state.tensors():
            dist.broadcast(t, src=src)
    return state


def gather_images(img: torch.Tensor, dst: int = 0, sizes: Optional[List[int]] = None):
    """Gather per-rank image stacks [V_r, H, W, 3] on `dst` (returns the concatenation there, None
    elsewhere).  `sizes` lists V_r per rank when the split is uneven."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return img
    world, rank = dist.get_world_size(), dist.get_rank()
    if sizes is None:
        sizes = [img.shape[0]] * world
    vmax = max(sizes)
    pad = img
    if img.shape[0] < vmax:
        pad = torch.cat([img, img.new_zeros((vmax - img.shape[0],) + tuple(img.shape[1:]))], dim=0)
    bufs = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
    dist.gather(pad.contiguous(), bufs, dst=dst)
    if rank != dst:
        return None
    return torch.cat([b[:n] for b, n in zip(bufs, sizes)], dim=0)
