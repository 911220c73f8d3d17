#!/usr/bin/env python
"""Per-kernel time breakdown of one bench step (CUDA events around every launch).
usage: python tools/profile_step.py [--views 4] [--tris 4096] [--resolution 512] [--config v1_1_swin_large]"""
import argparse
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from renderformer_b200 import ops  # noqa: E402
from renderformer_b200.config import RenderFormerConfig  # noqa: E402
from renderformer_b200.model import RenderFormer, RenderFormerRenderingPipeline  # noqa: E402
from renderformer_b200.synth import init_state_dict, make_scene  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--views", type=int, default=4)
ap.add_argument("--tris", type=int, default=4096)
ap.add_argument("--resolution", type=int, default=512)
ap.add_argument("--config", default="v1_1_swin_large")
a = ap.parse_args()
cfg = RenderFormerConfig.named(a.config)
model = RenderFormer(cfg)
model.load_state_dict(init_state_dict(cfg, 7))
pipe = RenderFormerRenderingPipeline(model)
pipe.to(torch.device("cuda:0"))
pipe.view_chunk = a.views
sc = {k: v.cuda() for k, v in make_scene(a.tris, a.views, seed=0).items()}


def step():
    return pipe(sc["triangles"], sc["texture"], sc["mask"], sc["vn"], sc["c2w"], sc["fov"], resolution=a.resolution)


for _ in range(3):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
step()
e1.record()
torch.cuda.synchronize()
print(f"plain step: {e0.elapsed_time(e1):.3f} ms")
ops.PROFILE = []
step()
torch.cuda.synchronize()
prof, ops.PROFILE = ops.PROFILE, None
agg = collections.OrderedDict()
for kind, fl, s, e, tag in prof:
    k = (kind, tag)
    t = agg.setdefault(k, [0.0, 0.0, 0])
    t[0] += s.elapsed_time(e)
    t[1] += fl
    t[2] += 1
tot = sum(v[0] for v in agg.values())
print(f"sum of kernel times: {tot:.3f} ms over {len(prof)} launches")
bykind = collections.defaultdict(float)
for (kind, tag), (ms, fl, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    bykind[kind] += ms
    tf = fl / (ms * 1e-3) / 1e12 if fl else 0.0
    print(f"{ms:8.3f} ms {100 * ms / tot:5.1f}%  n={n:3d}  {ms / n * 1e3:8.1f} us/launch  {tf:7.1f} TF/s  {kind:18s} {tag}")
print("--- by kind")
for k, ms in sorted(bykind.items(), key=lambda kv: -kv[1]):
    print(f"{ms:8.3f} ms {100 * ms / tot:5.1f}%  {k}")
