#!/usr/bin/env python
"""BASELINE.json configs[4]: triangles 512-4096 x resolution 256^2-1024^2 on one GPU against the
tensor roofline.  One step = one scene + `--views` views; prints a markdown table.
usage: python tools/sweep.py [--views 4] [--steps 8]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from renderformer_b200.config import RenderFormerConfig  # noqa: E402
from renderformer_b200.flops import job_flops  # noqa: E402
from renderformer_b200.model import RenderFormer, RenderFormerRenderingPipeline  # noqa: E402
from renderformer_b200.synth import init_state_dict, make_scene  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--views", type=int, default=4)
ap.add_argument("--steps", type=int, default=8)
ap.add_argument("--config", default="v1_1_swin_large")
ap.add_argument("--eager", action="store_true", help="launch from Python instead of CUDA-graph replay")
a = ap.parse_args()
peak = 1414.1
pk = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
if os.path.exists(pk):
    peak = json.load(open(pk)).get("bf16_tflops_sustained", peak)
cfg = RenderFormerConfig.named(a.config)
model = RenderFormer(cfg)
model.load_state_dict(init_state_dict(cfg, 7))
pipe = RenderFormerRenderingPipeline(model)
pipe.to(torch.device("cuda:0"))
pipe.view_chunk = a.views
pipe.cuda_graphs = not a.eager
pipe.max_cached_graphs = 1
print(f"launch mode: {'eager' if a.eager else 'CUDA-graph replay'}\n")
print(f"| triangles | resolution | ray tokens/view | ms/step ({a.views} views) | frames/s | TFLOP/step | TFLOP/s | frac of {peak:.0f} |")
print("|---|---|---|---|---|---|---|---|")
for n in (512, 1024, 2048, 4096):
    for r in (256, 512, 1024):
        sc = {k: v.cuda() for k, v in make_scene(n, a.views, seed=0).items()}

        def step():
            return pipe(sc["triangles"], sc["texture"], sc["mask"], sc["vn"], sc["c2w"], sc["fov"], resolution=r)
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.steps
        tf = job_flops(cfg, n, r, 1, a.views) / 1e12
        print(f"| {n} | {r}x{r} | {(r // 8) ** 2} | {ms:.2f} | {a.views / ms * 1e3:.1f} | {tf:.2f} | {tf / ms * 1e3:.0f} | "
              f"{tf / ms * 1e3 / peak:.2f} |", flush=True)
        del sc
        torch.cuda.empty_cache()
