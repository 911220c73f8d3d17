#!/usr/bin/env python
"""BASELINE.json configs[3] and configs[4] on N GPUs (N = WORLD_SIZE; also runs on one GPU):

* configs[3]: V1.1-Large video render -- ONE scene, 120-frame camera orbit, views split over the ranks;
* configs[4]: sweep triangles 512-4096 x resolution 256^2-1024^2, one scene + 32 views per step.

Same multi-GPU schedule as bench.py (row-sharded scene stage with one NCCL all-gather per encoder layer, view
slices, image gather on rank 0, one CUDA-graph replay per step and rank).  Timing: CUDA events, barrier +
synchronize on both sides, max over ranks.  Prints markdown on rank 0.

    python tools/scale_configs.py                                  # 1 GPU
    python -m torch.distributed.run --nproc-per-node 8 ... tools/scale_configs.py
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
os.environ.setdefault("NCCL_DEBUG", "WARN")

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from renderformer_b200.config import RenderFormerConfig  # noqa: E402
from renderformer_b200.dist import render_sharded, view_slice  # noqa: E402
from renderformer_b200.flops import job_flops  # noqa: E402
from renderformer_b200.model import RenderFormer, RenderFormerRenderingPipeline  # noqa: E402
from renderformer_b200.synth import init_state_dict, make_scene  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--orbit-views", type=int, default=120)
    ap.add_argument("--sweep-views", type=int, default=32)
    ap.add_argument("--view-chunk", type=int, default=8)
    ap.add_argument("--no-sweep", action="store_true")
    a = ap.parse_args()
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    peak = 1414.1
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = json.load(open(pk)).get("bf16_tflops_sustained", peak)
    cfg = RenderFormerConfig.named("v1_1_swin_large")
    model = RenderFormer(cfg)
    model.load_state_dict(init_state_dict(cfg, 7))
    pipe = RenderFormerRenderingPipeline(model)
    pipe.to(dev)
    pipe.cuda_graphs = True
    pipe.max_cached_graphs = 1

    def time_job(n_tris, res, views):
        pipe.view_chunk = max(1, min(a.view_chunk if res < 1024 else 2, -(-views // world)))
        sc = {k: v.to(dev) for k, v in make_scene(n_tris, views, seed=0).items()}

        def step():
            return render_sharded(pipe, sc["triangles"], sc["texture"], sc["mask"], sc["vn"], sc["c2w"], sc["fov"],
                                  resolution=res, dst=0, torch_dtype=torch.bfloat16)
        for _ in range(3):
            step()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.steps):
            step()
        e1.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / a.steps], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        pipe._graphs.clear()
        del sc
        torch.cuda.empty_cache()
        return ms.item()

    def out(*x):
        if rank == 0:
            print(*x, flush=True)

    V = a.orbit_views
    ms = time_job(4096, 512, V)
    tf = job_flops(cfg, 4096, 512, 1, V) / 1e12
    out(f"## configs[3]: V1.1-Large, 4096 triangles, {V}-frame orbit at 512x512, {world} GPU(s)\n")
    out(f"{ms:.1f} ms per orbit = **{V / ms * 1e3:.0f} frames/s** ({mine_text(V, world)}; scene stage once, row-sharded; "
        f"{tf:.0f} TFLOP per orbit = {tf / ms * 1e3 / world:.0f} TFLOP/s per GPU = {tf / ms * 1e3 / world / peak:.2f} of the measured "
        f"sustained bf16 peak {peak:.0f})\n")
    if not a.no_sweep:
        V = a.sweep_views
        out(f"## configs[4]: one scene + {V} views per step, {world} GPU(s), CUDA-graph replay, bf16 operands\n")
        out(f"| triangles | resolution | ray tokens/view | ms/step | frames/s | TFLOP/step | TFLOP/s per GPU | frac of {peak:.0f} |")
        out("|---|---|---|---|---|---|---|---|")
        for n in (512, 1024, 2048, 4096):
            for r in (256, 512, 1024):
                ms = time_job(n, r, V)
                tf = job_flops(cfg, n, r, 1, V) / 1e12
                out(f"| {n} | {r}x{r} | {(r // 8) ** 2} | {ms:.2f} | {V / ms * 1e3:.0f} | {tf:.1f} | {tf / ms * 1e3 / world:.0f} | "
                    f"{tf / ms * 1e3 / world / peak:.2f} |")
    if world > 1:
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        os._exit(0)


def mine_text(V, world):
    sizes = [view_slice(V, world, r).stop - view_slice(V, world, r).start for r in range(world)]
    return f"{min(sizes)}-{max(sizes)} views per GPU" if min(sizes) != max(sizes) else f"{sizes[0]} views per GPU"


if __name__ == "__main__":
    main()
