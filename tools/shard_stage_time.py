#!/usr/bin/env python
"""How long does ONE rank's share of the row-sharded view-independent stage take?  (ONE GPU, development tool.)

The schedule of rank r in a world of W (`Engine._encode_scene_sharded`) is run on this GPU with the NCCL
all-gathers replaced by a no-op (the other ranks' rows stay whatever the buffer holds: timing does not depend on
the values), replayed from a CUDA graph like bench.py does, plus a per-kernel breakdown (CUDA events around
every launch).  Add the measured all-gather times (tools/nccl_probe.py) for the full picture.
usage: python tools/shard_stage_time.py [--tris 4096] [--worlds 1,2,4,8] [--breakdown 8]"""
import argparse
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from renderformer_b200 import ops  # noqa: E402
from renderformer_b200.config import RenderFormerConfig  # noqa: E402
from renderformer_b200.engine import RowShard  # noqa: E402
from renderformer_b200.model import RenderFormer, RenderFormerRenderingPipeline  # noqa: E402
from renderformer_b200.synth import init_state_dict, make_scene  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--tris", type=int, default=4096)
ap.add_argument("--worlds", default="1,2,4,8")
ap.add_argument("--breakdown", type=int, default=8)
ap.add_argument("--config", default="v1_1_swin_large")
a = ap.parse_args()
cfg = RenderFormerConfig.named(a.config)
model = RenderFormer(cfg)
model.load_state_dict(init_state_dict(cfg, 7))
pipe = RenderFormerRenderingPipeline(model)
pipe.to(torch.device("cuda:0"))
eng = model.engine()
sc = {k: v.cuda() for k, v in make_scene(a.tris, 1, seed=0).items()}
n_gather = [0]


def no_gather(full, chunk):
    n_gather[0] += 1


class FakeStore:
    """Stands in for dist.SymmKVStore on one GPU: local buffers, the barrier is a no-op."""

    def __init__(self, rows, width, dtype, device):
        self.t = torch.zeros((2, rows, width), dtype=dtype, device=device)

    def buf(self, i):
        return self.t[i]

    def dst(self, i):
        return [self.t[i].data_ptr()], False

    def barrier(self):
        n_gather[0] += 1


_stores = {}


def fake_store(rows, width, dtype, device):
    key = (rows, width, dtype)
    if key not in _stores:
        _stores[key] = FakeStore(rows, width, dtype, device)
    return _stores[key]


def stage(world, rank=0):
    sh = RowShard(rank, world, no_gather, fake_store) if world > 1 else None
    return eng.encode_scene(sc["triangles"], sc["texture"], sc["mask"], sc["vn"], shard=sh)


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for world in [int(x) for x in a.worlds.split(",")]:
    for _ in range(2):
        stage(world)
    torch.cuda.synchronize()
    eager = timed(lambda: stage(world), n=5)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        keep = stage(world)
    t = timed(g.replay)
    print(f"world {world}: rank 0's scene stage  graph replay {t * 1e3:8.1f} us   eager {eager * 1e3:8.1f} us", flush=True)
    del g, keep

if a.breakdown > 0:
    world = a.breakdown
    ops.PROFILE = []
    stage(world)
    torch.cuda.synchronize()
    prof, ops.PROFILE = ops.PROFILE, None
    agg = collections.OrderedDict()
    for kind, fl, s, e, tag in prof:
        t = agg.setdefault((kind, tag), [0.0, 0.0, 0])
        t[0] += s.elapsed_time(e)
        t[1] += fl
        t[2] += 1
    tot = sum(v[0] for v in agg.values())
    print(f"--- world {world}, rank 0: sum of kernel times {tot:.3f} ms over {len(prof)} launches")
    for (kind, tag), (ms, fl, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        tf = fl / (ms * 1e-3) / 1e12 if fl else 0.0
        print(f"{ms:8.3f} ms {100 * ms / tot:5.1f}%  n={n:3d}  {ms / n * 1e3:8.1f} us/launch  {tf:7.1f} TF/s  {kind:18s} {tag}")
