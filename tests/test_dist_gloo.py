"""CPU, world_size 2, gloo: the N>1 host path (view sharding, SceneState broadcast, image gather)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from renderformer_b200.dist import broadcast_scene_state, gather_images, view_slice
from renderformer_b200.engine import SceneState


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _state(fill: bool) -> SceneState:
    g = torch.Generator().manual_seed(1)
    mk = (lambda *s: torch.randn(*s, generator=g)) if fill else (lambda *s: torch.zeros(*s))
    return SceneState(1, 24, 40, 40, mk(1, 40, 64), mk(1, 24, 9),
                      (mk(1, 24) > 0).to(torch.uint8), (mk(1, 4) * 100).to(torch.int32),
                      mk(1, 40, 3 * 64), mk(1, 3 * 64, 40).to(torch.bfloat16), 64)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        st = broadcast_scene_state(_state(fill=(rank == 0)), src=0)
        want = _state(fill=True)
        ok = all(torch.equal(a, b) for a, b in zip(st.tensors(), want.tensors()))
        sizes = [view_slice(5, world, r).stop - view_slice(5, world, r).start for r in range(world)]
        mine = view_slice(5, world, rank)
        img = torch.arange(5.0)[mine].view(-1, 1, 1, 1).expand(-1, 2, 2, 3).contiguous()
        out = gather_images(img, dst=0, sizes=sizes)
        if rank == 0:
            ok = ok and out.shape == (5, 2, 2, 3) and torch.equal(out[:, 0, 0, 0], torch.arange(5.0))
        else:
            ok = ok and out is None
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_view_slice_partitions():
    for n in (1, 4, 5, 32, 120):
        for w in (1, 2, 4, 8):
            idx = [i for r in range(w) for i in range(n)[view_slice(n, w, r)]]
            assert idx == list(range(n))
            sizes = [len(range(n)[view_slice(n, w, r)]) for r in range(w)]
            assert max(sizes) - min(sizes) <= 1


def test_broadcast_and_gather_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    assert res == {0: True, 1: True}
