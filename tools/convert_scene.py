#!/usr/bin/env python
"""Scene JSON (reference examples/*.json format) -> compressed .npz of pipeline tensors, numpy only.
usage: python tools/convert_scene.py examples/cbox.json out.npz"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from renderformer_b200.scene_io import load_scene, save_npz  # noqa: E402

src, dst = sys.argv[1], sys.argv[2]
scene = load_scene(src)
save_npz(scene, dst)
print(f"{src}: {scene['triangles'].shape[0]} triangles, {scene['c2w'].shape[0]} camera(s) -> {dst}")
