#!/usr/bin/env python
"""BASELINE.json configs[1]: V1.1-swin-Large, examples/cbox.json (5633 triangles, tests/golden/cbox_scene.npz),
1 view 512x512 on one B200 -- single-frame latency and frames/s, plus a stress point (8192 triangles, 1024^2).
usage: python tools/run_config.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from renderformer_b200 import scene_io as sio  # noqa: E402
from renderformer_b200.config import RenderFormerConfig  # noqa: E402
from renderformer_b200.flops import job_flops  # noqa: E402
from renderformer_b200.model import RenderFormer, RenderFormerRenderingPipeline  # noqa: E402
from renderformer_b200.synth import init_state_dict, make_scene  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
cfg = RenderFormerConfig.named("v1_1_swin_large")
model = RenderFormer(cfg)
model.load_state_dict(init_state_dict(cfg, 7))
pipe = RenderFormerRenderingPipeline(model)
pipe.to(torch.device("cuda:0"))


def timed(sc, res, graphs, steps=10):
    pipe.cuda_graphs = graphs
    g = {k: v.cuda() for k, v in sc.items()}

    def step():
        return pipe(g["triangles"], g["texture"], g["mask"], g["vn"], g["c2w"], g["fov"], resolution=res)
    for _ in range(3):
        img = step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        img = step()
    e1.record()
    torch.cuda.synchronize()
    assert torch.isfinite(img).all()
    return e0.elapsed_time(e1) / steps


cbox = sio.load_npz(os.path.join(ROOT, "tests", "golden", "cbox_scene.npz"))
for name, sc, n, res in (("cbox.json (5633 tris), full 32x32 textures", sio.to_pipeline_inputs(cbox), 5633, 512),
                         ("cbox.json (5633 tris), constant-texture path", sio.to_pipeline_inputs(cbox, constant_texture=True), 5633, 512),
                         ("stress: 8192 synthetic tris, 1 view 1024x1024", make_scene(8192, 1, seed=0), 8192, 1024)):
    v = sc["c2w"].shape[1]
    tf = job_flops(cfg, n, res, 1, v) / 1e12
    for graphs in (False, True):
        ms = timed(sc, res, graphs)
        print(f"{name}: {v} view {res}x{res}, {'CUDA graph' if graphs else 'eager'}: {ms:.2f} ms/frame = "
              f"{v / ms * 1e3:.1f} frames/s ({tf:.2f} TFLOP, {tf / ms * 1e3:.0f} TFLOP/s)", flush=True)
    pipe._graphs.clear()
    torch.cuda.empty_cache()
