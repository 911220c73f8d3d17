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


@torch.no_grad()
def render_stream_sharded(pipe, scenes, resolution: int = 512, src: int = 0, ldr: Optional[str] = None):
    """Multi-GPU counterpart of `RenderFormerRenderingPipeline.render_stream`: every rank iterates the SAME
    sequence of host scene dicts (keys of `render`); rank `src` uploads geometry + texture and runs the
    view-independent stage, its SceneState is broadcast with NCCL, each rank renders its contiguous slice
    of the views and downloads its own images (no gather: a rank can write its own frames).  Uploads of
    scene i+1 and downloads of scene i overlap the kernels of the neighbouring scene.

    Yields (view_slice, pinned host tensor [B, V_rank, H, W, 3]) per scene; the buffer belongs to a ring of
    three."""
    dev = pipe.device
    world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
    rank = dist.get_rank() if world > 1 else 0
    main = torch.cuda.current_stream(dev)
    copy = torch.cuda.Stream(dev)

    def upload(sc):
        V = sc["c2w"].shape[1]
        mine = view_slice(V, world, rank)
        with torch.cuda.stream(copy):
            d = {"c2w": sc["c2w"][:, mine].contiguous().to(dev, non_blocking=True),
                 "fov": sc["fov"][:, mine].contiguous().to(dev, non_blocking=True)}
            if rank == src:
                for k in ("triangles", "texture", "mask", "vn"):
                    d[k] = sc[k].to(dev, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy)
        return d, ev, mine, tuple(sc["triangles"].shape[:2])

    it = iter(scenes)
    first = next(it, None)
    pending = upload(first) if first is not None else None
    ring, slot, prev = [None, None, None], 0, None
    while pending is not None:
        d, ev, mine, (B, N) = pending
        nxt = next(it, None)
        pending = upload(nxt) if nxt is not None else None
        main.wait_event(ev)
        if rank == src:
            st = pipe.encode(d["triangles"], d["texture"], d["mask"], d["vn"])
        else:
            st = pipe.static_scene_state(B, N)  # one persistent receive buffer per shape
        broadcast_scene_state(st, src=src)
        img = pipe.render_views(st, d["c2w"], d["fov"], resolution)
        for t in d.values():
            t.record_stream(main)
        if ldr is not None:
            img = pipe.hdr_to_ldr(img, ldr)
        if ring[slot] is None or ring[slot].shape != img.shape or ring[slot].dtype != img.dtype:
            ring[slot] = torch.empty(img.shape, dtype=img.dtype, pin_memory=True)
        host = ring[slot]
        slot = (slot + 1) % 3
        host.copy_(img, non_blocking=True)
        done = torch.cuda.Event()
        done.record(main)
        if prev is not None:
            prev[1].synchronize()
            yield prev[2], prev[0]
        prev = (host, done, mine)
    if prev is not None:
        prev[1].synchronize()
        yield prev[2], prev[0]
