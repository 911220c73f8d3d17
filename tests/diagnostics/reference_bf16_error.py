#!/usr/bin/env python
"""What does bf16 cost the REFERENCE ITSELF?  (CPU; test infrastructure, imports the unmodified
reference through oracle/reference_loader.py.)

Renders one scene with the unmodified reference twice -- `torch_dtype=float32` (the grading reference) and
`torch_dtype=bfloat16` (CPU autocast: every Linear / conv / SDPA in bf16, norms and residuals as autocast leaves
them) -- with the same seeded weights, and prints the north_star metrics of the second against the first.  This is
the error floor of "bf16 tensor-core math" for that scene: a from-scratch bf16 implementation cannot be expected to
sit far below it.
usage: python tests/diagnostics/reference_bf16_error.py [cbox | <n_tris>] [resolution] [config]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from oracle.reference_loader import build_reference_pipeline  # noqa: E402
from renderformer_b200.config import RenderFormerConfig  # noqa: E402
from renderformer_b200.metrics import hdr_rel_err, log_psnr  # noqa: E402
from renderformer_b200.synth import init_state_dict, make_scene  # noqa: E402


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "cbox"
    R = int(sys.argv[2]) if len(sys.argv) > 2 else 512
    cfg = RenderFormerConfig.named(sys.argv[3] if len(sys.argv) > 3 else "v1_1_swin_large")
    torch.set_num_threads(os.cpu_count() or 1)
    if what == "cbox":
        from renderformer_b200 import scene_io as sio
        sc = sio.to_pipeline_inputs(sio.load_npz(os.path.join(ROOT, "tests", "golden", "cbox_scene.npz")))
    else:
        sc = make_scene(int(what), 1, seed=0)
    pipe, ref = build_reference_pipeline(cfg, init_state_dict(cfg, 7))
    out = {}
    for name, dt in (("float32", torch.float32), ("bfloat16", torch.bfloat16)):
        t0 = time.time()
        with torch.no_grad():
            img = pipe(sc["triangles"].clone(), sc["texture"].clone(), sc["mask"].clone(), sc["vn"].clone(),
                       sc["c2w"].clone(), sc["fov"].clone(), resolution=R, torch_dtype=dt)
        out[name] = img.float()
        print(f"reference ({ref}) torch_dtype={name}: {time.time() - t0:.1f} s, image range {img.min().item():.4f}..{img.max().item():.4f}", flush=True)
    a, b = out["bfloat16"], out["float32"]
    print(f"{what} {R}x{R} {cfg.name if hasattr(cfg, 'name') else ''}: reference bf16 autocast vs reference fp32: "
          f"hdr rel {hdr_rel_err(a, b):.3e}  log-PSNR {log_psnr(a, b):.1f} dB")


if __name__ == "__main__":
    main()
