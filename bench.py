#!/usr/bin/env python
"""Headline benchmark: frames/sec of the RenderFormer-V1.1-swin-Large forward pass on synthetic
4096-triangle scenes at 512x512 (BASELINE.json metric), N GPUs of one node.

    python bench.py --gpus 1 --steps 10 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...   # CPU arm: the oracle port of the reference's PyTorch path

One step = one scene through both stages: rank 0 runs the view-independent stage once, the
per-layer triangle K/V is broadcast with NCCL, every rank renders `--views-per-gpu` views of a
camera orbit, images are gathered on rank 0 (weak scaling: per-GPU work is fixed).
Prints ONE JSON line on rank 0 (contract: see the task statement / DESIGN.md §Measurement).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# NCCL writes its debug / version banner to stdout by default; the contract is ONE JSON line there
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
    os.environ["NCCL_DEBUG"] = "WARN"

import torch  # noqa: E402

METRIC = "frames/sec (512^2, 4096 tris)"
UNIT = "frames/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="v1_1_swin_large")
    ap.add_argument("--tris", type=int, default=4096)
    ap.add_argument("--resolution", type=int, default=512)
    ap.add_argument("--views-per-gpu", type=int, default=4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-roofline", action="store_true")
    ap.add_argument("--no-cuda-graphs", action="store_true", help="launch every kernel from Python")
    ap.add_argument("--view-chunk", type=int, default=0, help="views per decoder pass (0 = all of a GPU's views)")
    ap.add_argument("--view-streams", type=int, default=1, help="CUDA streams the view chunks are spread over")
    ap.add_argument("--torch-cuda", action="store_true",
                    help="also time the oracle port on cuda:0 with torch kernels under bf16 autocast "
                         "(context: what the reference's PyTorch CUDA path costs on this GPU)")
    return ap.parse_args()


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return p.get("bf16_tflops_sustained", p.get("bf16_tflops")), p.get("hbm_gbs"), "measured (MEASURED_PEAKS.json, sustained)"
    return 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


# --------------------------------------------------------------------------- CPU arm
def cpu_sample(cfg_name: str, steps: int, warmup: int, full_tris: int, full_res: int, views_per_step: int):
    """Time the oracle (CPU fp32 port of the reference's PyTorch path) on a bounded sample of the
    workload and scale to the metric by algorithmic FLOPs.  Returns (frames/s, ms/step, cores, sample)."""
    from oracle import renderformer_oracle as orc
    from renderformer_b200.config import RenderFormerConfig
    from renderformer_b200.flops import job_flops
    from renderformer_b200.synth import init_state_dict, make_scene

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = RenderFormerConfig.named(cfg_name)
    s_tris, s_res = min(full_tris, 1024), min(full_res, 128)
    sd = init_state_dict(cfg, 7)
    sc = make_scene(s_tris, 1, seed=0)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        orc.render(sd, cfg, sc["triangles"], sc["texture"], sc["mask"], sc["vn"], sc["c2w"], sc["fov"], s_res)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    dt = statistics.median(times)
    cpu_flops_per_s = job_flops(cfg, s_tris, s_res, 1, 1) / dt
    step_flops = job_flops(cfg, full_tris, full_res, 1, views_per_step)
    fps = views_per_step / (step_flops / cpu_flops_per_s)
    sample = (f"oracle fp32 (torch CPU, {cores} threads): {cfg_name}, {s_tris} tris, 1 view {s_res}x{s_res} "
              f"= {job_flops(cfg, s_tris, s_res, 1, 1) / 1e12:.3f} TFLOP in {dt:.2f} s "
              f"({cpu_flops_per_s / 1e12:.3f} TFLOP/s), scaled by algorithmic FLOPs to "
              f"{full_tris} tris, {views_per_step} views {full_res}x{full_res} per step")
    return fps, dt * 1e3, cores, sample


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    fps, ms, cores, sample = cpu_sample(args.config, max(1, args.steps), min(args.warmup, 1), args.tris,
                                        args.resolution, args.views_per_gpu)
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.config}, synthetic {args.tris}-tri scene, {args.resolution}x{args.resolution}, "
                               f"{args.views_per_gpu} views/step (CPU arm: bounded sample, FLOP-scaled)"},
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    """SM clock / power / throttle reasons sampled with NVML every 20 ms from a thread while the
    timed region runs (nvidia-smi's own loop is too coarse for a region of a few hundred ms)."""

    def __init__(self, index: int):
        import threading
        self.rows, self.on, self.n = [], True, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.n = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:  # noqa: BLE001
            self.err = repr(e)
            return
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()

    def _run(self):
        n = self.n
        while self.on:
            try:
                self.rows.append((n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM),
                                  n.nvmlDeviceGetPowerUsage(self.h) / 1e3,
                                  n.nvmlDeviceGetCurrentClocksEventReasons(self.h)))
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.02)

    def stop(self):
        if self.n is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable: " + getattr(self, "err", "")]}
        self.on = False
        self.t.join(timeout=2)
        n, rows = self.n, self.rows
        names = {"hw_slowdown": n.nvmlClocksEventReasonHwSlowdown,
                 "hw_thermal_slowdown": n.nvmlClocksEventReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": n.nvmlClocksEventReasonSwThermalSlowdown,
                 "sw_power_cap": n.nvmlClocksEventReasonSwPowerCap,
                 "hw_power_brake": n.nvmlClocksEventReasonHwPowerBrakeSlowdown}
        bits = 0
        for r in rows:
            bits |= r[2]
        return {"sm_mhz": statistics.median(r[0] for r in rows) if rows else None, "sm_max_mhz": self.max,
                "power_w_max": max((r[1] for r in rows), default=None), "samples": len(rows),
                "reasons": sorted(k for k, v in names.items() if bits & v)}


# --------------------------------------------------------------------------- GPU arm
def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    import torch.distributed as dist
    from renderformer_b200 import lib, ops
    from renderformer_b200.config import RenderFormerConfig
    from renderformer_b200.dist import broadcast_scene_state
    from renderformer_b200.flops import job_flops, scene_flops, view_flops
    from renderformer_b200.model import RenderFormer, RenderFormerRenderingPipeline
    from renderformer_b200.synth import init_state_dict, make_scene

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    cfg = RenderFormerConfig.named(args.config)
    model = RenderFormer(cfg)
    model.load_state_dict(init_state_dict(cfg, 7))
    pipe = RenderFormerRenderingPipeline(model)
    pipe.to(dev)
    eng = model.engine()
    Vl, R, N = args.views_per_gpu, args.resolution, args.tris
    pipe.view_chunk = args.view_chunk or Vl
    pipe.view_streams = args.view_streams

    scene = make_scene(N, Vl * world, seed=0)
    host = {k: v.pin_memory() for k, v in scene.items()}
    my = slice(rank * Vl, (rank + 1) * Vl)
    host["c2w_local"] = scene["c2w"][:, my].contiguous().pin_memory()
    host["fov_local"] = scene["fov"][:, my].contiguous().pin_memory()
    d_in = {k: v.to(dev) for k, v in host.items()}
    out_host = torch.empty((1, Vl, R, R, 3), dtype=torch.float32).pin_memory()
    gather_list = [torch.empty((1, Vl, R, R, 3), dtype=torch.float32, device=dev) for _ in range(world)] \
        if (world > 1 and rank == 0) else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    graphs = not args.no_cuda_graphs
    recv_state = None

    def step_device(inp):
        """Both stages with inputs resident in HBM; returns this rank's images."""
        if world == 1:
            if pipe.cuda_graphs:  # one CUDA-graph replay of the whole call (inputs copied into its static buffers)
                return pipe.render(inp["triangles"], inp["texture"], inp["mask"], inp["vn"], inp["c2w_local"],
                                   inp["fov_local"], resolution=R, torch_dtype=torch.bfloat16)
            st = pipe.encode(inp["triangles"], inp["texture"], inp["mask"], inp["vn"])
        else:
            nonlocal recv_state
            if rank != 0 and recv_state is None:
                recv_state = pipe.static_scene_state(1, N)  # persistent receive buffers
            st = pipe.encode(inp["triangles"], inp["texture"], inp["mask"], inp["vn"]) if rank == 0 else recv_state
            broadcast_scene_state(st, src=0)  # per-layer triangle K/V + tokens over NVLink (NCCL)
        img = pipe.render_views(st, inp["c2w_local"], inp["fov_local"], R)
        if world > 1:
            dist.gather(img, gather_list, dst=0)
        return img

    def step_e2e():
        """Public API with HOST buffers: H2D of the step's inputs, render, D2H of the images."""
        inp = {}
        if rank == 0 or world == 1:
            for k in ("triangles", "texture", "mask", "vn"):
                inp[k] = host[k].to(dev, non_blocking=True)
        else:
            inp = {k: None for k in ("triangles", "texture", "mask", "vn")}
        inp["c2w_local"] = host["c2w_local"].to(dev, non_blocking=True)
        inp["fov_local"] = host["fov_local"].to(dev, non_blocking=True)
        if world == 1:
            img = pipe.render(inp["triangles"], inp["texture"], inp["mask"], inp["vn"], inp["c2w_local"],
                              inp["fov_local"], resolution=R, torch_dtype=torch.bfloat16)
        else:
            img = step_device(inp)
        out_host.copy_(img, non_blocking=False)

    def timed(fn, steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item() / steps

    pipe.cuda_graphs = graphs
    for _ in range(max(args.warmup, 3)):
        step_device(d_in)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches0 = lib.launch_count() + pipe.replayed_launches
    ms_step = timed(lambda: step_device(d_in), args.steps)
    launches = (lib.launch_count() + pipe.replayed_launches - launches0)
    clocks = sampler.stop() if sampler else None
    frames = Vl * world
    value = frames / (ms_step * 1e-3)

    e2e = None
    if not args.no_e2e:
        for _ in range(2):
            step_e2e()
        ms_e2e = timed(step_e2e, args.steps)
        h2d = sum(host[k].numel() * host[k].element_size() for k in ("triangles", "texture", "mask", "vn"))
        h2d += world * (host["c2w_local"].numel() + host["fov_local"].numel()) * 4
        e2e = {"value": frames / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e,
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(world * out_host.numel() * 4),
               "api": "RenderFormerRenderingPipeline.render, one blocking call per step"}
        if world == 1:
            # the batch API (batch_infer.py use case): every step still uploads its own inputs from
            # pinned host memory and downloads its own images inside the timed region, but the upload
            # of step i+1 and the download of step i overlap the kernels of the neighbouring step
            host_scene = {"triangles": host["triangles"], "texture": host["texture"], "mask": host["mask"],
                          "vn": host["vn"], "c2w": host["c2w_local"], "fov": host["fov_local"]}

            def stream_steps(n):
                sink = 0.0
                for img in pipe.render_stream((host_scene for _ in range(n)), resolution=R,
                                              torch_dtype=torch.bfloat16):
                    sink += float(img[0, 0, 0, 0, 0])  # touch the host result of every step
                return sink

            stream_steps(3)
            ms_stream = timed(lambda: stream_steps(args.steps), 1) / args.steps
            e2e = {"value": frames / (ms_stream * 1e-3), "unit": UNIT, "ms_per_step": ms_stream,
                   "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(out_host.numel() * 4),
                   "api": "RenderFormerRenderingPipeline.render_stream (host scenes in, pinned host images out; "
                          "copies of neighbouring steps overlap the kernels)",
                   "single_call": {"value": frames / (ms_e2e * 1e-3), "ms_per_step": ms_e2e,
                                   "api": "RenderFormerRenderingPipeline.render, one blocking call per step"}}
        else:
            from renderformer_b200.dist import render_stream_sharded
            host_scene = {k: host[k] for k in ("triangles", "texture", "mask", "vn", "c2w", "fov")}

            def stream_steps(n):
                sink = 0.0
                for _mine, img in render_stream_sharded(pipe, (host_scene for _ in range(n)), resolution=R):
                    sink += float(img[0, 0, 0, 0, 0])
                return sink

            stream_steps(3)
            ms_stream = timed(lambda: stream_steps(args.steps), 1) / args.steps
            e2e = {"value": frames / (ms_stream * 1e-3), "unit": UNIT, "ms_per_step": ms_stream,
                   "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(world * out_host.numel() * 4),
                   "api": "renderformer_b200.dist.render_stream_sharded (rank 0 uploads the scene, NCCL broadcast of "
                          "the scene state, every rank renders and downloads its own views; copies overlap kernels)",
                   "single_call": {"value": frames / (ms_e2e * 1e-3), "ms_per_step": ms_e2e,
                                   "api": "one blocking step per scene"}}

    roofline = None
    if not args.no_roofline and rank == 0:
        peak_tf, _, peak_src = load_peaks()
        pipe.cuda_graphs = False  # per-launch events need eager launches
        ops.PROFILE = []
        torch.cuda.synchronize()
        step_device(d_in) if world == 1 else (pipe.render_views(
            pipe.encode(d_in["triangles"], d_in["texture"], d_in["mask"], d_in["vn"]), d_in["c2w_local"],
            d_in["fov_local"], R))
        torch.cuda.synchronize()
        prof, ops.PROFILE = ops.PROFILE, None
        pipe.cuda_graphs = graphs
        by = {}
        for kind, fl, a, b, _tag in prof:
            t = a.elapsed_time(b) * 1e-3
            cur = by.setdefault(kind, [0.0, 0.0, 0])
            cur[0] += fl
            cur[1] += t
            cur[2] += 1
        g = by.get("gemm", [0.0, 1.0, 1])
        att = by.get("attention", [0.0, 1.0, 1])
        ach = g[0] / g[1] / 1e12
        traffic, traffic_src = None, None
        tpath = os.path.join(ROOT, "profiles", "r01e_gemm_traffic.json")
        if os.path.exists(tpath) and args.config == "v1_1_swin_large" and (N, R, Vl) == (4096, 512, 4):
            with open(tpath) as f:
                tj = json.load(f)
            traffic, traffic_src = tj["dram_bytes_per_launch"], "profiles/r01e_gemm_traffic.json (ncu dram__bytes_read+write, avg over one step's GEMM launches)"
        roofline = {
            "kernel": "gemm_tc_kernel (tcgen05 GEMM / implicit-GEMM conv)", "bound": "tensor",
            "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf, "traffic": traffic,
            "traffic_source": traffic_src, "flop_per_launch": g[0] / g[2],
            "peak_source": peak_src, "launches_per_step": g[2], "avg_launch_us": g[1] / g[2] * 1e6,
            "share_of_step": g[1] * 1e3 / ms_step,
            "method": "per-launch CUDA events on the launch stream, one instrumented step after the timed region",
            "attention": {"kernel": "attn_tc_kernel", "achieved": att[0] / att[1] / 1e12, "unit": "TFLOP/s",
                          "frac": att[0] / att[1] / 1e12 / peak_tf, "launches_per_step": att[2],
                          "share_of_step": att[1] * 1e3 / ms_step},
        }
    if world > 1:
        dist.barrier()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        fps, ms, cores, sample = cpu_sample(args.config, 2, 1, N, R, Vl)
        cpu = {"value": fps, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}

    torch_cuda = None
    if rank == 0 and world == 1 and args.torch_cuda:
        # context only: the reference's algorithm executed by torch's own CUDA kernels (SDPA, cuBLAS,
        # cuDNN) on this GPU.  The unmodified reference cannot travel to the GPU box (it imports
        # roma / needs /root/reference), so this is the oracle port under bf16 autocast -- a checker
        # being timed beside the product, never on the product path.
        from oracle import renderformer_oracle as orc
        sd_gpu = {k: v.to(dev) for k, v in init_state_dict(cfg, 7).items()}
        g_in = {k: v.to(dev) for k, v in scene.items()}

        def torch_step():
            with torch.autocast("cuda", dtype=torch.bfloat16):
                return orc.render(sd_gpu, cfg, g_in["triangles"], g_in["texture"], g_in["mask"], g_in["vn"],
                                  g_in["c2w"], g_in["fov"], R, view_chunk=Vl)
        try:
            for _ in range(2):
                torch_step()
            ms_t = timed(torch_step, max(2, min(args.steps, 5)))
            torch_cuda = {"value": frames / (ms_t * 1e-3), "unit": UNIT, "ms_per_step": ms_t, "kind": "port",
                          "what": "oracle (functional restatement of the reference) on cuda:0, torch SDPA/cuBLAS/cuDNN "
                                  "kernels, bf16 autocast, K/V recomputed per view like the reference"}
        except Exception as e:  # noqa: BLE001
            torch_cuda = {"unavailable": repr(e)[:200]}
        del sd_gpu, g_in
        torch.cuda.empty_cache()

    if rank == 0:
        step_tflop = job_flops(cfg, N, R, 1, frames) / 1e12
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {
                "workload": f"{args.config} (483M, random-init), synthetic {N}-triangle scene, {R}x{R}, "
                            f"{Vl} views/GPU per step ({frames} views total; 32-view batch at 8 GPUs = BASELINE configs[2]); "
                            "scene stage once per step on rank 0 + NCCL broadcast of per-layer K/V",
                "l2": "inputs + weights (~1.3 GB per step) exceed the 126 MB L2; no explicit flush",
                "launch": (("one CUDA-graph replay per step" if world == 1 else "CUDA-graph replay of each stage, NCCL eager")
                           + " (pipeline.cuda_graphs)" if graphs else "eager launches from Python"),
                "algorithmic_tflop_per_step": step_tflop,
                "model_flops_utilisation": step_tflop / (ms_step * 1e-3) / world / load_peaks()[0],
                "scene_tflop": scene_flops(cfg, N) / 1e12, "view_tflop": view_flops(cfg, N, R) / 1e12,
            },
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
        }
        if torch_cuda is not None:
            line["torch_cuda_port"] = torch_cuda
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
