"""GPU: the reference's own command-line programs, UNCHANGED (baseline/_ref/infer.py and batch_infer.py, installed
by tools/install_reference.py), run against this repo's drop-in `renderformer` package and
`renderformer_liger_kernel` hook.  Third-party packages the CLIs import and this image lacks are replaced by
minimal stand-ins that only move data (h5py -> reads an .npz written under an .h5 name, imageio -> records
what would have been written, natsort -> sorted, simple_ocio -> unused with --tone_mapper none).  The HDR
frames the CLIs "write" are compared with the fp32 oracle on the same scene file.

`infer.py:43` hard-codes `cuda:1`: that program is exercised on boxes with two GPUs (and then also covers a
model living on a device that is not the current one)."""
import importlib
import os
import sys
import types

import numpy as np
import pytest
import torch

from renderformer_b200.config import RenderFormerConfig
from renderformer_b200.metrics import PSNR_MIN, REL_TOL, hdr_rel_err, log_psnr
from renderformer_b200.synth import init_state_dict, make_scene

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")
needs_ref = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "batch_infer.py")),
                               reason="baseline/_ref missing: run tools/install_reference.py in the build container")


class _H5File:
    def __init__(self, path, mode="r"):
        self.d = np.load(path)

    def __enter__(self):
        return self.d

    def __exit__(self, *a):
        self.d.close()


@pytest.fixture()
def cli_env(tmp_path, monkeypatch):
    written = {}
    h5py = types.ModuleType("h5py")
    h5py.File = _H5File
    imageio = types.ModuleType("imageio")
    imageio.v3 = types.SimpleNamespace(imwrite=lambda path, arr, **kw: written.__setitem__(os.path.basename(path), np.asarray(arr).copy()))
    natsort = types.ModuleType("natsort")
    natsort.natsorted = sorted
    ocio = types.ModuleType("simple_ocio")
    ocio.ToneMapper = lambda name: (_ for _ in ()).throw(RuntimeError("tone mapper not expected in this test"))
    for name, mod in (("h5py", h5py), ("imageio", imageio), ("natsort", natsort), ("simple_ocio", ocio)):
        monkeypatch.setitem(sys.modules, name, mod)
    # the drop-in package must win over the reference's own `renderformer` inside baseline/_ref
    monkeypatch.setattr(sys, "path", [ROOT] + [p for p in sys.path if os.path.abspath(p or ".") != ROOT] + [REF])
    for k in [k for k in sys.modules if k == "renderformer" or k.startswith("renderformer.")]:
        monkeypatch.delitem(sys.modules, k)
    import renderformer
    assert os.path.abspath(renderformer.__file__).startswith(os.path.join(ROOT, "renderformer") + os.sep)

    cfg = RenderFormerConfig.named("tiny_swin")
    sd = init_state_dict(cfg, 21)
    model = renderformer.RenderFormer(cfg)
    model.load_state_dict(sd)
    model_dir = tmp_path / "model"
    model.save_pretrained(str(model_dir))
    return dict(cfg=cfg, sd=sd, model_dir=str(model_dir), written=written, tmp=tmp_path)


def _write_scene(path, n_tris, views, seed):
    sc = make_scene(n_tris, views, seed=seed)
    with open(path, "wb") as f:  # the stand-in h5py.File reads an .npz under the .h5 name
        np.savez(f, triangles=sc["triangles"][0].numpy(), texture=sc["texture"][0].numpy().astype(np.float16),
                 vn=sc["vn"][0].numpy(), c2w=sc["c2w"][0].numpy(), fov=sc["fov"][0, :, 0].numpy())
    sc["texture"] = sc["texture"].half().float()  # what the CLI will see after the fp16 round trip
    return sc


def _check(env, sc, got_hdr, got_ldr, view, res):
    from oracle import renderformer_oracle as orc
    ref = orc.render(env["sd"], env["cfg"], sc["triangles"], sc["texture"], sc["mask"], sc["vn"], sc["c2w"][:, view:view + 1],
                     sc["fov"][:, view:view + 1], res)[0, 0]
    got = torch.from_numpy(got_hdr)
    assert got.shape == ref.shape and got.dtype == torch.float32
    rel, psnr = hdr_rel_err(got, ref), log_psnr(got, ref)
    assert rel <= REL_TOL and psnr >= PSNR_MIN, (rel, psnr)
    assert got_ldr.dtype == np.uint8 and np.array_equal(got_ldr, (np.clip(got_hdr, 0, 1) * 255).astype(np.uint8))


@needs_ref
def test_batch_infer_cli_unchanged(cli_env, monkeypatch):
    """batch_infer.py:61-174: dataset of .h5 files padded to a common length, DataLoader batches of two scenes,
    apply_kernels hook, pipeline(...), per-view EXR/PNG and the MP4."""
    env = cli_env
    folder = env["tmp"] / "scenes"
    folder.mkdir()
    scenes = {f"s{i}": _write_scene(str(folder / f"s{i}.h5"), n, 2, seed=30 + i) for i, n in enumerate((40, 57, 33))}
    out_dir = env["tmp"] / "out"
    monkeypatch.setattr(sys, "argv", ["batch_infer.py", "--h5_folder", str(folder), "--model_id", env["model_dir"],
                                      "--precision", "fp16", "--resolution", "64", "--batch_size", "2",
                                      "--padding_length", "64", "--output_dir", str(out_dir)])
    monkeypatch.delitem(sys.modules, "batch_infer", raising=False)
    cli = importlib.import_module("batch_infer")
    assert os.path.abspath(cli.__file__) == os.path.join(REF, "batch_infer.py")
    cli.main()
    w = env["written"]
    assert "video.mp4" in w and w["video.mp4"].shape == (6, 64, 64, 3)
    for name, sc in scenes.items():
        for v in range(2):
            _check(env, sc, w[f"{name}_view_{v}.exr"], w[f"{name}_view_{v}.png"], v, 64)


@needs_ref
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="infer.py:43 hard-codes cuda:1")
def test_infer_cli_unchanged(cli_env, monkeypatch):
    """infer.py:33-106 on cuda:1 while cuda:0 stays the current device."""
    env = cli_env
    path = str(env["tmp"] / "scene.h5")
    sc = _write_scene(path, 77, 3, seed=5)
    monkeypatch.setattr(sys, "argv", ["infer.py", "--h5_file", path, "--model_id", env["model_dir"], "--precision", "bf16",
                                      "--resolution", "128", "--output_dir", str(env["tmp"] / "o")])
    monkeypatch.delitem(sys.modules, "infer", raising=False)
    cli = importlib.import_module("infer")
    assert os.path.abspath(cli.__file__) == os.path.join(REF, "infer.py")
    assert torch.cuda.current_device() == 0
    cli.main()
    for v in range(3):
        _check(env, sc, env["written"][f"scene_view_{v}.exr"], env["written"][f"scene_view_{v}.png"], v, 128)
