#!/usr/bin/env python
"""Per-kernel breakdown of a bench step under SUSTAINED load (clocks / power sampled with NVML).
Runs `--steps` back-to-back steps (event-timed as a whole), then one instrumented step while the
GPU is still hot, and prints the cold (first steps after start) and hot breakdowns side by side.
usage: python tools/sustained_profile.py [--steps 40] [--views 4]"""
import argparse
import collections
import os
import statistics
import sys
import threading
import time

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
ap.add_argument("--steps", type=int, default=40)
ap.add_argument("--config", default="v1_1_swin_large")
a = ap.parse_args()


class Nvml(threading.Thread):
    def __init__(self, period=0.01):
        super().__init__(daemon=True)
        import pynvml
        self.n = pynvml
        pynvml.nvmlInit()
        self.h = pynvml.nvmlDeviceGetHandleByIndex(0)
        self.period, self.rows, self.on = period, [], True

    def run(self):
        n = self.n
        while self.on:
            self.rows.append((time.perf_counter(), n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM),
                              n.nvmlDeviceGetPowerUsage(self.h) / 1e3,
                              n.nvmlDeviceGetCurrentClocksEventReasons(self.h)))
            time.sleep(self.period)


cfg = RenderFormerConfig.named(a.config)
model = RenderFormer(cfg)
model.load_state_dict(init_state_dict(cfg, 7))
pipe = RenderFormerRenderingPipeline(model)
pipe.to(torch.device("cuda:0"))
pipe.view_chunk = a.views
sc = {k: v.cuda() for k, v in make_scene(a.tris, a.views, seed=0).items()}


def step():
    return pipe(sc["triangles"], sc["texture"], sc["mask"], sc["vn"], sc["c2w"], sc["fov"], resolution=a.resolution)


def instrumented():
    ops.PROFILE = []
    step()
    torch.cuda.synchronize()
    prof, ops.PROFILE = ops.PROFILE, None
    agg = collections.OrderedDict()
    for kind, fl, s, e, tag in prof:
        t = agg.setdefault((kind, tag), [0.0, 0.0, 0])
        t[0] += s.elapsed_time(e)
        t[1] += fl
        t[2] += 1
    return agg


for _ in range(3):
    step()
torch.cuda.synchronize()
cold = instrumented()
time.sleep(1.0)
mon = Nvml()
mon.start()
evs = [torch.cuda.Event(enable_timing=True) for _ in range(a.steps + 1)]
t_start = time.perf_counter()
evs[0].record()
for i in range(a.steps):
    step()
    evs[i + 1].record()
torch.cuda.synchronize()
t_end = time.perf_counter()
hot = instrumented()
mon.on = False
per = [evs[i].elapsed_time(evs[i + 1]) for i in range(a.steps)]
print("per-step ms:", " ".join(f"{x:.1f}" for x in per))
rows = [r for r in mon.rows if t_start <= r[0] <= t_end]
if rows:
    print(f"nvml during sustained region: {len(rows)} samples, sm MHz median {statistics.median(r[1] for r in rows)}"
          f" min {min(r[1] for r in rows)} max {max(r[1] for r in rows)}, power W median "
          f"{statistics.median(r[2] for r in rows):.0f} max {max(r[2] for r in rows):.0f}, reasons OR "
          f"{hex(__import__('functools').reduce(lambda x, y: x | y, (r[3] for r in rows)))}")
    print("clock trace (every 10th):", " ".join(str(r[1]) for r in rows[::10]))
    print("power trace (every 10th):", " ".join(f"{r[2]:.0f}" for r in rows[::10]))
tc, th = sum(v[0] for v in cold.values()), sum(v[0] for v in hot.values())
print(f"cold step kernels {tc:.3f} ms, hot step kernels {th:.3f} ms")
print(f"{'cold ms':>9s} {'hot ms':>9s} {'ratio':>6s} {'n':>4s} {'hot TF/s':>9s}  kernel")
for k, (ms, fl, n) in sorted(hot.items(), key=lambda kv: -kv[1][0]):
    c = cold.get(k, [0, 0, 0])[0]
    tf = fl / (ms * 1e-3) / 1e12 if fl else 0.0
    print(f"{c:9.3f} {ms:9.3f} {ms / c if c else 0:6.2f} {n:4d} {tf:9.1f}  {k[0]} {k[1]}")
