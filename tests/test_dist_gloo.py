"""CPU, world_size 2, gloo: the N>1 host path (view sharding, SceneState broadcast, image gather)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from renderformer_b200.dist import broadcast_scene_state, gather_images, row_shard, view_slice
from renderformer_b200.engine import RowShard, SceneState


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
        # the per-layer collective of the row-sharded scene stage: in-place all-gather of the own chunk
        sh = row_shard()
        ok = ok and sh is not None and (sh.rank, sh.world) == (rank, world)
        ntp = 40
        S = sh.shard_rows(ntp)
        full = torch.full((world * S, 6), -1.0)
        r0, r1 = sh.my_rows(ntp)
        full[r0:r1] = torch.arange(r0, r1, dtype=torch.float32)[:, None].expand(-1, 6)
        sh.all_gather(full, full[rank * S:(rank + 1) * S])
        ok = ok and torch.equal(full[:ntp, 0], torch.arange(float(ntp)))
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


def test_row_shard_partitions():
    """Rows are partitioned contiguously in multiples of 8; trailing ranks may own nothing."""
    for ntp in (8, 24, 40, 1040, 4112, 5656, 8208):
        for w in (1, 2, 4, 8):
            shards = [RowShard(r, w, None) for r in range(w)]
            S = shards[0].shard_rows(ntp)
            assert S % 8 == 0 and w * S >= ntp
            rows = [sh.my_rows(ntp) for sh in shards]
            assert rows[0][0] == 0 and rows[-1][1] == ntp
            assert all(a[1] == b[0] for a, b in zip(rows, rows[1:]))
            assert all(0 <= r1 - r0 <= S for r0, r1 in rows)


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


def _worker_few_views(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        V = 1  # fewer views than ranks: rank 1 owns an empty slice and still takes part in the gather
        sizes = [view_slice(V, world, r).stop - view_slice(V, world, r).start for r in range(world)]
        mine = view_slice(V, world, rank)
        img = torch.full((mine.stop - mine.start, 2, 2, 3), 7.0)
        out = gather_images(img, dst=0, sizes=sizes)
        ok = sizes == [1, 0] and img.shape[0] == (1 if rank == 0 else 0)
        ok = ok and ((out.shape == (1, 2, 2, 3) and bool((out == 7.0).all())) if rank == 0 else out is None)
        h = gather_images(img, dst=0, sizes=sizes, async_op=True)  # the asynchronous form bench.py uses
        out2 = h.wait()
        ok = ok and ((out2.shape == (1, 2, 2, 3)) if rank == 0 else out2 is None)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_gather_with_fewer_views_than_ranks_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_few_views, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    assert res == {0: True, 1: True}
