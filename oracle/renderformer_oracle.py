"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Never imported by the product path.

CPU fp32 restatement (plain torch ops, functional style) of the reference's inference forward
pass, used as the checker in tests/, in __graft_entry__.smoke() and as the timed CPU baseline
of bench.py.  Every function cites the reference file:line it follows (paths relative to
the reference root).  The reference ships no tests or golden vectors, so this
restatement is pinned against the *unmodified reference itself*: oracle/make_golden.py
imports /root/reference in the build container, runs it on seeded inputs and commits the
outputs to tests/golden/; tests/test_oracle_golden.py checks this file against them.

Third-party arithmetic used by the reference on this path and restated here through the
same torch primitives: torch (unpinned in requirements.txt:14; 2.11.0 here) for
linear/conv/rms_norm/SDPA, einops rearrange (pure index math), roma.Rigid (requirements.txt:3,
absent here; restated as R^T (x - t), transform.py:22-27).
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

EPS = 1e-6  # layers/attention.py:16


# --------------------------------------------------------------------------- encodings
def nerf_encode(x: torch.Tensor, num_freqs: int) -> torch.Tensor:
    """encodings/nerf_encoding.py:63-84 with include_input=True: [x | sin(x*2^j) | cos(x*2^j)]."""
    if num_freqs == 0:
        return x
    freqs = 2.0 ** torch.linspace(0.0, num_freqs - 1, num_freqs, device=x.device)
    sc = (x[..., None] * freqs).reshape(*x.shape[:-1], -1)
    enc = torch.sin(torch.cat([sc, sc + math.pi / 2.0], dim=-1))
    return torch.cat([x, enc], dim=-1)


def rope_tables(pos9: torch.Tensor, freqs: torch.Tensor, head_dim: int):
    """encodings/rope.py:152-206 + :78-103.  pos9 [B,N,9] -> cos, sin [B,1,N,head_dim]."""
    ang = pos9.float()[..., None] * freqs.float()  # [B,N,9,F]  index = coord*F + k
    ang = ang.reshape(*pos9.shape[:-1], -1)
    half = head_dim // 2
    ang = F.pad(ang, (0, half - ang.shape[-1]))  # pairs beyond 9*F rotate by angle 0
    ang = torch.cat([ang, ang], dim=-1)[:, None]
    return ang.cos(), ang.sin()


def rope_rotate(x: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor) -> torch.Tensor:
    """encodings/rope.py:41-45,131-149: x*cos + rotate_half(x)*sin, pairs (i, i+hd/2)."""
    h = x.shape[-1] // 2
    rot = torch.cat([-x[..., h:], x[..., :h]], dim=-1)
    return x * cos + rot * sin


def rms(x: torch.Tensor, w: torch.Tensor, eps: Optional[float]) -> torch.Tensor:
    """nn.RMSNorm; eps=None means torch.finfo(dtype).eps."""
    return F.rms_norm(x, (x.shape[-1],), w, eps)


# --------------------------------------------------------------------------- geometry
def to_camera_space(c2w: torch.Tensor, tris: torch.Tensor) -> torch.Tensor:
    """utils/transform.py:7-27: apply the inverse rigid transform.  c2w [B,4,4], tris [B,N,3,3]."""
    R, t = c2w[:, :3, :3], c2w[:, :3, 3]
    return torch.einsum("bji,bnvj->bnvi", R, tris - t[:, None, None, :])


def camera_rays(fov_rad: torch.Tensor, res: int) -> torch.Tensor:
    """utils/ray_generator.py:13-50 with identity c2w (camera space).  fov [B,1] -> [B,res,res,3]."""
    px = torch.linspace(0.5, res - 0.5, res, dtype=fov_rad.dtype, device=fov_rad.device)
    x, y = torch.meshgrid(px, px, indexing="xy")
    c = res / 2
    fl = res / 2 / torch.tan(0.5 * fov_rad[..., 0, None, None])
    d = torch.stack([(x - c) / fl, -(y - c) / fl, -torch.ones_like(x).expand_as((x - c) / fl)], dim=-1)
    return F.normalize(d, dim=-1, p=2)


def centroid_positions(pos9: torch.Tensor, mask: torch.Tensor, n_reg: int):
    """models/renderformer.py:103-124: register tokens sit at the masked vertex centroid."""
    w = (mask.float() / (mask.sum(dim=1, keepdim=True) + 1e-5))[..., None]
    cen = (w * pos9).sum(dim=1).reshape(-1, 3, 3).mean(dim=1, keepdim=True).repeat(1, n_reg, 3)
    pos = torch.cat([cen, pos9], dim=1)
    m = torch.cat([torch.ones((pos9.shape[0], n_reg), dtype=torch.bool, device=mask.device), mask], dim=1)
    return pos, m


# --------------------------------------------------------------------------- layers
def attention(q, k, v, heads: int, key_mask=None, attn_mask=None):
    """layers/attention.py:131-161: head split + SDPA (scale 1/sqrt(head_dim)), True = attend."""
    B, Nq, D = q.shape
    Nk = k.shape[1]
    qh = q.view(B, Nq, heads, -1).transpose(1, 2)
    kh = k.view(B, Nk, heads, -1).transpose(1, 2)
    vh = v.view(B, Nk, heads, -1).transpose(1, 2)
    m = attn_mask
    if key_mask is not None:
        m = key_mask.view(B, 1, 1, Nk).expand(-1, heads, -1, -1)
    o = F.scaled_dot_product_attention(qh, kh, vh, attn_mask=m)
    return o.transpose(1, 2).reshape(B, Nq, D)


def swiglu(sd, p, x):
    """layers/attention.py:56-57."""
    return F.linear(F.silu(F.linear(x, sd[p + "w1.weight"])) * F.linear(x, sd[p + "w3.weight"]), sd[p + "w2.weight"])


def swin_region_mask(H: int, W: int, ws: int, shift: int, device) -> torch.Tensor:
    """layers/attention.py:238-271: [nW, ws*ws, ws*ws] bool on the rolled grid (True = attend)."""
    img = torch.zeros((H, W), device=device)
    cnt = 0
    for hs in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
        for wsl in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
            img[hs, wsl] = cnt
            cnt += 1
    win = img.view(H // ws, ws, W // ws, ws).permute(0, 2, 1, 3).reshape(-1, ws * ws)
    return win[:, None, :] == win[:, :, None]


def swin_self_attention(sd, p, x, heads: int, Hp: int, Wp: int, shift: int, ws: int = 8):
    """layers/attention.py:316-370 (no RoPE)."""
    B, N, C = x.shape
    g = x.view(B, Hp, Wp, C)
    if shift > 0:
        g = torch.roll(g, shifts=(-shift, -shift), dims=(1, 2))
    win = g.view(B, Hp // ws, ws, Wp // ws, ws, C).permute(0, 1, 3, 2, 4, 5).reshape(-1, ws * ws, C)
    q, k, v = F.linear(win, sd[p + "in_proj.weight"]).chunk(3, dim=-1)
    q = rms(q, sd[p + "q_norm.weight"], EPS)
    k = rms(k, sd[p + "k_norm.weight"], EPS)
    am = None
    if shift > 0:
        am = swin_region_mask(Hp, Wp, ws, shift, x.device).repeat(B, 1, 1)[:, None]
    o = attention(q, k, v, heads, attn_mask=am)
    o = F.linear(o, sd[p + "out_proj.weight"])
    g = o.view(B, Hp // ws, Wp // ws, ws, ws, C).permute(0, 1, 3, 2, 4, 5).reshape(B, Hp, Wp, C)
    if shift > 0:
        g = torch.roll(g, shifts=(shift, shift), dims=(1, 2))
    return g.reshape(B, N, C)


def encoder_layer(sd, p, x, heads, key_mask, cos, sin):
    """layers/attention.py:484-527 (self-attention branch) + :115-202."""
    h = rms(x, sd[p + "query_norm.weight"], EPS)
    q, k, v = F.linear(h, sd[p + "multihead_attn.in_proj.weight"]).chunk(3, dim=-1)
    q = rms(q, sd[p + "multihead_attn.q_norm.weight"], EPS)
    k = rms(k, sd[p + "multihead_attn.k_norm.weight"], EPS)
    B, N, D = q.shape
    qh = rope_rotate(q.view(B, N, heads, -1).transpose(1, 2), cos, sin).transpose(1, 2).reshape(B, N, D)
    kh = rope_rotate(k.view(B, N, heads, -1).transpose(1, 2), cos, sin).transpose(1, 2).reshape(B, N, D)
    x = x + F.linear(attention(qh, kh, v, heads, key_mask=key_mask), sd[p + "multihead_attn.out_proj.weight"])
    return x + swiglu(sd, p + "ffn.", rms(x, sd[p + "ffn_norm.weight"], EPS))


def decoder_layer(sd, p, x, ctx, heads, key_mask, q_cs, k_cs, swin: bool, shift: int, Hp: int, Wp: int):
    """layers/attention.py:484-527 (cross-attn -> [swin|full] self-attn -> SwiGLU)."""
    B, Nq, D = x.shape
    h = rms(x, sd[p + "query_norm.weight"], EPS)
    c = rms(ctx, sd[p + "kv_norm.weight"], EPS)
    q = rms(F.linear(h, sd[p + "multihead_attn.q_proj.weight"]), sd[p + "multihead_attn.q_norm.weight"], EPS)
    k = rms(F.linear(c, sd[p + "multihead_attn.k_proj.weight"]), sd[p + "multihead_attn.k_norm.weight"], EPS)
    v = F.linear(c, sd[p + "multihead_attn.v_proj.weight"])
    Nk = k.shape[1]
    q = rope_rotate(q.view(B, Nq, heads, -1).transpose(1, 2), *q_cs).transpose(1, 2).reshape(B, Nq, D)
    k = rope_rotate(k.view(B, Nk, heads, -1).transpose(1, 2), *k_cs).transpose(1, 2).reshape(B, Nk, D)
    x = x + F.linear(attention(q, k, v, heads, key_mask=key_mask), sd[p + "multihead_attn.out_proj.weight"])

    h = rms(x, sd[p + "self_attn_norm.weight"], EPS)
    if swin:
        x = x + swin_self_attention(sd, p + "self_attn.", h, heads, Hp, Wp, shift)
    else:
        q, k, v = F.linear(h, sd[p + "self_attn.in_proj.weight"]).chunk(3, dim=-1)
        q = rms(q, sd[p + "self_attn.q_norm.weight"], EPS)
        k = rms(k, sd[p + "self_attn.k_norm.weight"], EPS)
        q = rope_rotate(q.view(B, Nq, heads, -1).transpose(1, 2), *q_cs).transpose(1, 2).reshape(B, Nq, D)
        k = rope_rotate(k.view(B, Nq, heads, -1).transpose(1, 2), *q_cs).transpose(1, 2).reshape(B, Nq, D)
        x = x + F.linear(attention(q, k, v, heads), sd[p + "self_attn.out_proj.weight"])
    return x + swiglu(sd, p + "ffn.", rms(x, sd[p + "ffn_norm.weight"], EPS))


# --------------------------------------------------------------------------- DPT head
def _rcu(sd, p, x):
    """layers/dpt.py:76-92."""
    y = F.conv2d(F.silu(x), sd[p + "conv1.weight"], sd[p + "conv1.bias"], padding=1)
    y = F.conv2d(F.silu(y), sd[p + "conv2.weight"], sd[p + "conv2.bias"], padding=1)
    return y + x


def _fusion(sd, p, size, x0, x1=None):
    """layers/dpt.py:133-159."""
    out = x0
    if x1 is not None:
        out = out + _rcu(sd, p + "resConvUnit1.", x1)
    out = _rcu(sd, p + "resConvUnit2.", out)
    out = F.interpolate(out, size=size, mode="bilinear", align_corners=True)
    return F.conv2d(out, sd[p + "out_conv.weight"], sd[p + "out_conv.bias"])


def dpt_head(sd, p, feats, Hp: int, Wp: int, patch: int):
    """layers/dpt.py:242-273.  feats: 4 x [B, Hp*Wp, D] -> [B, 3, Hp*patch, Wp*patch]."""
    maps = []
    for i, f in enumerate(feats):
        x = f.permute(0, 2, 1).reshape(f.shape[0], f.shape[-1], Hp, Wp)
        x = F.conv2d(x, sd[p + f"projects.{i}.weight"], sd[p + f"projects.{i}.bias"])
        if i == 0:
            x = F.conv_transpose2d(x, sd[p + "resize_layers.0.weight"], sd[p + "resize_layers.0.bias"], stride=4)
        elif i == 1:
            x = F.conv_transpose2d(x, sd[p + "resize_layers.1.weight"], sd[p + "resize_layers.1.bias"], stride=2)
        elif i == 3:
            x = F.conv2d(x, sd[p + "resize_layers.3.weight"], sd[p + "resize_layers.3.bias"], stride=2, padding=1)
        maps.append(F.conv2d(x, sd[p + f"scratch.layer{i + 1}_rn.weight"], None, padding=1))
    l1, l2, l3, l4 = maps
    s = p + "scratch."
    p4 = _fusion(sd, s + "refinenet4.", l3.shape[2:], l4)
    p3 = _fusion(sd, s + "refinenet3.", l2.shape[2:], p4, l3)
    p2 = _fusion(sd, s + "refinenet2.", l1.shape[2:], p3, l2)
    p1 = _fusion(sd, s + "refinenet1.", (l1.shape[2] * 2, l1.shape[3] * 2), p2, l1)
    out = F.conv2d(p1, sd[s + "output_conv1.weight"], sd[s + "output_conv1.bias"], padding=1)
    out = F.interpolate(out, (Hp * patch, Wp * patch), mode="bilinear", align_corners=True)
    out = F.silu(F.conv2d(out, sd[s + "output_conv2.0.weight"], sd[s + "output_conv2.0.bias"], padding=1))
    return F.conv2d(out, sd[s + "output_conv2.2.weight"], sd[s + "output_conv2.2.bias"])


# --------------------------------------------------------------------------- stages
def scene_stage(sd, cfg, tri9, texture, mask, vn9):
    """models/renderformer.py:126-169,185-186: token construction + view-independent encoder.
    `texture` must already be log-encoded (pipelines/rendering_pipeline.py:67-68)."""
    B, N = tri9.shape[:2]
    vn_emb = rms(F.linear(nerf_encode(vn9, cfg.vn_pe_num_freqs), sd["vn_encoding_proj.weight"],
                          sd["vn_encoding_proj.bias"]), sd["vn_encoder_norm.weight"], None)
    tex_emb = rms(F.linear(texture.reshape(B, N, -1), sd["texture_encoder.weight"], sd["texture_encoder.bias"]),
                  sd["texture_encoder_norm.weight"], None)
    tri_emb = sd["tri_token"] + tex_emb + vn_emb
    seq = torch.cat([sd["reg_tokens"].expand(B, -1, -1), tri_emb], dim=1)
    pos, key_mask = centroid_positions(tri9, mask, cfg.num_register_tokens)
    cos, sin = rope_tables(pos, sd["transformer.rope_emb.freqs"], cfg.latent_dim // cfg.num_heads)
    for i in range(cfg.num_layers):
        seq = encoder_layer(sd, f"transformer.layers.{i}.", seq, cfg.num_heads, key_mask, cos, sin)
    return seq, key_mask


def view_stage(sd, cfg, seq, key_mask, tri9_cam, mask, rays_d, taps: Optional[dict] = None):
    """models/view_transformer.py:88-127 for a batch of views sharing nothing but weights.
    seq [BV,Nt,d], key_mask [BV,Nt], tri9_cam [BV,N,9], mask [BV,N], rays_d [BV,R,R,3]."""
    p = "view_transformer."
    BV, R = rays_d.shape[0], rays_d.shape[1]
    P = cfg.patch_size
    Hp = Wp = R // P
    dv, heads = cfg.view_transformer_latent_dim, cfg.view_transformer_n_heads
    rm = nerf_encode(rays_d, cfg.vdir_num_freqs)
    # 'b (h1 p1) (w1 p2) c -> b (h1 w1) (c p1 p2)'   view_transformer.py:105
    tok = rm.view(BV, Hp, P, Wp, P, -1).permute(0, 1, 3, 5, 2, 4).reshape(BV, Hp * Wp, -1)
    x = sd[p + "ray_map_patch_token"] + rms(
        F.linear(tok, sd[p + "ray_map_encoder.weight"], sd[p + "ray_map_encoder.bias"]),
        sd[p + "ray_map_encoder_norm.weight"], None)
    pos_k, _ = centroid_positions(tri9_cam, mask, cfg.num_register_tokens)
    fr = sd[p + "transformer.rope_emb.freqs"]
    hd = dv // heads
    q_cs = rope_tables(torch.zeros((BV, Hp * Wp, 9), device=x.device), fr, hd)  # camera origin = 0
    k_cs = rope_tables(pos_k, fr, hd)
    out_layers = (list(range(cfg.view_transformer_n_layers - 4, cfg.view_transformer_n_layers))
                  if cfg.dpt_out_layers is None else list(cfg.dpt_out_layers))
    feats = []
    for i in range(cfg.view_transformer_n_layers):
        x = decoder_layer(sd, p + f"transformer.layers.{i}.", x, seq, heads, key_mask, q_cs, k_cs,
                          cfg.view_transformer_use_swin_attn, 0 if i % 2 == 0 else 4, Hp, Wp)
        if i in out_layers:
            feats.append(x)
    if taps is not None:
        taps["dec_feats"] = [f.clone() for f in feats]
    img = dpt_head(sd, p + "out_dpt.", feats, Hp, Wp, P)
    return F.elu(img, alpha=1e-3)  # view_transformer.py:86,122


@torch.no_grad()
def render(sd: Dict[str, torch.Tensor], cfg, triangles, texture, mask, vn, c2w, fov, resolution: int = 512,
           taps: Optional[dict] = None, view_chunk: int = 4) -> torch.Tensor:
    """pipelines/rendering_pipeline.py:28-125 in true fp32.  Does NOT mutate `texture`.
    Returns HDR [B, V, R, R, 3]."""
    B, V = c2w.shape[:2]
    N = triangles.shape[1]
    tex = texture.clone()
    tex[:, :, -3:] = torch.log10(tex[:, :, -3:] + 1.0)
    tri9 = triangles.reshape(B, N, 9).float()
    seq, key_mask = scene_stage(sd, cfg, tri9, tex.float(), mask, vn.reshape(B, N, 9).float())
    if taps is not None:
        taps["seq"] = seq.clone()
    out = []
    for b in range(B):
        imgs = []
        for v0 in range(0, V, view_chunk):
            v1 = min(V, v0 + view_chunk)
            n = v1 - v0
            tri_cam = to_camera_space(c2w[b, v0:v1], triangles[b:b + 1].expand(n, -1, -1, -1)).reshape(n, N, 9)
            rays = camera_rays(fov[b, v0:v1] / 180.0 * math.pi, resolution)
            t = {} if (taps is not None and b == 0 and v0 == 0) else None
            img = view_stage(sd, cfg, seq[b:b + 1].expand(n, -1, -1), key_mask[b:b + 1].expand(n, -1),
                             tri_cam, mask[b:b + 1].expand(n, -1), rays, t)
            if t:
                taps.update(t)
            imgs.append(img)
        out.append(torch.cat(imgs, dim=0))
    log_img = torch.stack(out, dim=0).permute(0, 1, 3, 4, 2)  # [B,V,R,R,3]
    if taps is not None:
        taps["log_img"] = log_img.clone()
    return torch.pow(10.0, log_img) - 1.0
