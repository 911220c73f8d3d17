"""Deterministic synthetic weights and scenes (no network: no checkpoints, no datasets).

`init_state_dict` produces a state_dict with exactly the reference's key names and shapes
(verified by `load_state_dict(strict=True)` into the unmodified reference
in oracle/make_golden.py), filled from a seeded generator that does not depend on module
construction order.  `make_scene` follows the input conventions of the reference's
scene_processor/to_h5.py:37-92 (13 texture channels x 32 x 32, triangular texel mask,
Blender-style look-at cameras).
"""
from __future__ import annotations

import math
from typing import Dict

import numpy as np
import torch

from .config import RenderFormerConfig


def state_dict_shapes(cfg: RenderFormerConfig) -> Dict[str, tuple]:
    d, f, L = cfg.latent_dim, cfg.dim_feedforward, cfg.num_layers
    dv, fv, Lv = cfg.view_transformer_latent_dim, cfg.view_transformer_ffn_hidden_dim, cfg.view_transformer_n_layers
    tex_in = cfg.texture_channels * cfg.texture_encode_patch_size ** 2
    vn_in = 9 + 9 * cfg.vn_pe_num_freqs * 2
    ray_in = 3 * cfg.patch_size ** 2
    F = cfg.dpt_features
    C = list(cfg.dpt_out_channels)
    s: Dict[str, tuple] = {}
    s["tri_token"] = (1, 1, d)
    s["reg_tokens"] = (1, cfg.num_register_tokens, d)
    s["vn_encoding_proj.weight"] = (d, vn_in)
    s["vn_encoding_proj.bias"] = (d,)
    s["vn_encoder_norm.weight"] = (d,)
    s["texture_encoder.weight"] = (d, tex_in)
    s["texture_encoder.bias"] = (d,)
    s["texture_encoder_norm.weight"] = (d,)
    for i in range(L):
        p = f"transformer.layers.{i}."
        s[p + "multihead_attn.in_proj.weight"] = (3 * d, d)
        s[p + "multihead_attn.out_proj.weight"] = (d, d)
        s[p + "multihead_attn.q_norm.weight"] = (d,)
        s[p + "multihead_attn.k_norm.weight"] = (d,)
        s[p + "query_norm.weight"] = (d,)
        s[p + "ffn.w1.weight"] = (f, d)
        s[p + "ffn.w2.weight"] = (d, f)
        s[p + "ffn.w3.weight"] = (f, d)
        s[p + "ffn_norm.weight"] = (d,)
    s["transformer.rope_emb.freqs"] = (cfg.vertex_pe_num_freqs // 2,)
    v = "view_transformer."
    s[v + "ray_map_patch_token"] = (1, 1, dv)
    s[v + "ray_map_encoder.weight"] = (dv, ray_in)
    s[v + "ray_map_encoder.bias"] = (dv,)
    s[v + "ray_map_encoder_norm.weight"] = (dv,)
    for i in range(Lv):
        p = v + f"transformer.layers.{i}."
        s[p + "multihead_attn.q_proj.weight"] = (dv, dv)
        s[p + "multihead_attn.k_proj.weight"] = (dv, d)
        s[p + "multihead_attn.v_proj.weight"] = (dv, d)
        s[p + "multihead_attn.out_proj.weight"] = (dv, dv)
        s[p + "multihead_attn.q_norm.weight"] = (dv,)
        s[p + "multihead_attn.k_norm.weight"] = (dv,)
        s[p + "query_norm.weight"] = (dv,)
        s[p + "kv_norm.weight"] = (d,)
        s[p + "self_attn.in_proj.weight"] = (3 * dv, dv)
        s[p + "self_attn.out_proj.weight"] = (dv, dv)
        s[p + "self_attn.q_norm.weight"] = (dv,)
        s[p + "self_attn.k_norm.weight"] = (dv,)
        s[p + "self_attn_norm.weight"] = (dv,)
        s[p + "ffn.w1.weight"] = (fv, dv)
        s[p + "ffn.w2.weight"] = (dv, fv)
        s[p + "ffn.w3.weight"] = (fv, dv)
        s[p + "ffn_norm.weight"] = (dv,)
    s[v + "transformer.rope_emb.freqs"] = (cfg.view_rope_dim // 2,)
    o = v + "out_dpt."
    for i in range(4):
        s[o + f"projects.{i}.weight"] = (C[i], dv, 1, 1)
        s[o + f"projects.{i}.bias"] = (C[i],)
    s[o + "resize_layers.0.weight"] = (C[0], C[0], 4, 4)
    s[o + "resize_layers.0.bias"] = (C[0],)
    s[o + "resize_layers.1.weight"] = (C[1], C[1], 2, 2)
    s[o + "resize_layers.1.bias"] = (C[1],)
    s[o + "resize_layers.3.weight"] = (C[3], C[3], 3, 3)
    s[o + "resize_layers.3.bias"] = (C[3],)
    for i in range(4):
        s[o + f"scratch.layer{i + 1}_rn.weight"] = (F, C[i], 3, 3)
    for r in (1, 2, 3, 4):
        p = o + f"scratch.refinenet{r}."
        s[p + "out_conv.weight"] = (F, F, 1, 1)
        s[p + "out_conv.bias"] = (F,)
        for u in ((1, 2) if r != 4 else (2,)):
            for c in (1, 2):
                s[p + f"resConvUnit{u}.conv{c}.weight"] = (F, F, 3, 3)
                s[p + f"resConvUnit{u}.conv{c}.bias"] = (F,)
    s[o + "scratch.output_conv1.weight"] = (F // 2, F, 3, 3)
    s[o + "scratch.output_conv1.bias"] = (F // 2,)
    s[o + "scratch.output_conv2.0.weight"] = (32, F // 2, 3, 3)
    s[o + "scratch.output_conv2.0.bias"] = (32,)
    s[o + "scratch.output_conv2.2.weight"] = (4 if cfg.include_alpha else 3, 32, 1, 1)
    s[o + "scratch.output_conv2.2.bias"] = (4 if cfg.include_alpha else 3,)
    return s


def rope_freqs(dim: int) -> torch.Tensor:
    # reference: renderformer/encodings/rope.py:170-174 (log-spaced, 2**linspace(0, log2(dim/2-1), dim/2))
    return 2.0 ** torch.linspace(0, math.log(dim // 2 - 1, 2), dim // 2)


def init_state_dict(cfg: RenderFormerConfig, seed: int = 7) -> Dict[str, torch.Tensor]:
    """Random-init weights of the named architecture, fp32, on CPU."""
    shapes = state_dict_shapes(cfg)
    sd: Dict[str, torch.Tensor] = {}
    for idx, key in enumerate(sorted(shapes)):
        shp = shapes[key]
        g = torch.Generator().manual_seed(seed * 1_000_003 + idx)
        if key.endswith("rope_emb.freqs"):
            t = rope_freqs(shp[0] * 2)
        elif key.endswith("norm.weight"):
            t = 1.0 + 0.1 * torch.randn(shp, generator=g)
        elif key.endswith(".bias"):
            t = 0.05 * torch.randn(shp, generator=g)
        elif key in ("tri_token", "reg_tokens") or key.endswith("ray_map_patch_token"):
            t = torch.randn(shp, generator=g)
        elif key.endswith("output_conv2.2.weight"):
            t = torch.randn(shp, generator=g) * 0.015  # keeps the log-HDR image in a sane range
        else:  # linear / conv weights: unit-gain fan-in scaling
            if "resize_layers.0" in key or "resize_layers.1" in key:
                fan_in = shp[0]  # ConvTranspose2d weight is [in, out, k, k]; stride == k
            else:
                fan_in = int(np.prod(shp[1:]))
            t = torch.randn(shp, generator=g) / math.sqrt(fan_in)
        sd[key] = t.float().contiguous()
    return sd


def look_at_c2w(position, target=(0.0, 0.0, 0.0), up=(0.0, 0.0, 1.0)) -> np.ndarray:
    """Camera-to-world for a -Z-forward, +Y-up camera (scene_processor/to_h5.py:10-34)."""
    pos = np.asarray(position, dtype=np.float64)
    back = pos - np.asarray(target, dtype=np.float64)
    back /= np.linalg.norm(back)
    right = np.cross(np.asarray(up, dtype=np.float64), back)
    right /= np.linalg.norm(right)
    upv = np.cross(back, right)
    c2w = np.eye(4)
    c2w[:3, 0], c2w[:3, 1], c2w[:3, 2], c2w[:3, 3] = right, upv, back, pos
    return c2w.astype(np.float32)


def make_scene(n_tris: int, n_views: int = 1, seed: int = 0, pad_to: int | None = None,
               fov_deg: float = 37.5, radius: float = 2.0) -> Dict[str, torch.Tensor]:
    """Synthetic scene in the reference's documented input range.

    Returns batch-1 tensors: triangles [1,N,3,3], texture [1,N,13,32,32], mask [1,N] bool,
    vn [1,N,3,3], c2w [1,V,4,4], fov [1,V,1] (degrees).
    """
    rng = np.random.default_rng(seed)
    cen = rng.uniform(-0.5, 0.5, size=(n_tris, 1, 3))
    tri = (cen + rng.normal(0.0, 0.03, size=(n_tris, 3, 3))).astype(np.float32)
    fn = np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0])
    fn /= np.linalg.norm(fn, axis=-1, keepdims=True) + 1e-12
    vn = fn[:, None, :] + 0.1 * rng.normal(size=(n_tris, 3, 3))
    vn = (vn / np.linalg.norm(vn, axis=-1, keepdims=True)).astype(np.float32)

    diffuse = rng.uniform(0.0, 0.8, size=(n_tris, 3))
    specular = np.repeat(rng.uniform(0.0, 0.2, size=(n_tris, 1)), 3, axis=1)
    rough = rng.uniform(0.01, 0.99, size=(n_tris, 1))
    normal = np.tile(np.array([[0.5, 0.5, 1.0]]), (n_tris, 1))
    emission = np.zeros((n_tris, 3))
    n_lights = int(rng.integers(1, 9))
    emission[rng.choice(n_tris, size=min(n_lights, n_tris), replace=False)] = 5000.0
    tex13 = np.concatenate([diffuse, specular, rough, normal, emission], axis=1).astype(np.float32)
    ii, jj = np.meshgrid(np.arange(32), np.arange(32), indexing="ij")
    texel_mask = (ii + jj <= 32).astype(np.float32)
    texture = tex13[:, :, None, None] * texel_mask[None, None]

    mask = np.ones(n_tris, dtype=bool)
    if pad_to is not None and pad_to > n_tris:
        extra = pad_to - n_tris
        tri = np.concatenate([tri, np.zeros((extra, 3, 3), np.float32)])
        vn = np.concatenate([vn, np.zeros((extra, 3, 3), np.float32)])
        texture = np.concatenate([texture, np.zeros((extra, 13, 32, 32), np.float32)])
        mask = np.concatenate([mask, np.zeros(extra, dtype=bool)])

    c2w = []
    for v in range(n_views):
        ang = 2.0 * math.pi * v / max(n_views, 1) - math.pi / 2
        c2w.append(look_at_c2w((radius * math.cos(ang), radius * math.sin(ang), 0.3 * math.sin(2 * ang))))
    c2w = np.stack(c2w)
    fov = np.full((n_views, 1), fov_deg, dtype=np.float32)
    return {
        "triangles": torch.from_numpy(tri)[None],
        "texture": torch.from_numpy(texture.astype(np.float32))[None],
        "mask": torch.from_numpy(mask)[None],
        "vn": torch.from_numpy(vn)[None],
        "c2w": torch.from_numpy(c2w)[None],
        "fov": torch.from_numpy(fov)[None],
    }
