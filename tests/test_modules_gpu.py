"""GPU: the reference's module boundaries (renderformer.layers.* / encodings.* / utils.* /
models.view_transformer) called ON THEIR OWN, class by class, against the fp32 oracle's restatement of the
same reference lines.  The classes are imported through the drop-in package paths a reference user has."""
import math

import pytest
import torch
import torch.nn.functional as F

from oracle import renderformer_oracle as orc
from renderformer_b200.config import RenderFormerConfig
from renderformer_b200.metrics import rel_l2
from renderformer_b200.synth import init_state_dict, make_scene

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _sub(sd, prefix):
    return {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}


def _cfg_sd(name="tiny_swin", seed=7):
    cfg = RenderFormerConfig.named(name)
    return cfg, init_state_dict(cfg, seed)


def _rand(*shape, seed=0, scale=1.0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed)) * scale


def _encoder_layer_module(cfg, sd, i=0):
    from renderformer.layers.attention import AttentionLayer
    m = AttentionLayer(query_dim=cfg.latent_dim, num_heads=cfg.num_heads, ffn_hidden_dim=cfg.dim_feedforward, dropout=0.0,
                       bias=False, activation="swiglu", norm_type="rms_norm", qk_norm=True)
    m.load_state_dict(_sub(sd, f"transformer.layers.{i}."), strict=True)
    return m.to(DEV).eval()


def test_feed_forward_swiglu():
    from renderformer.layers.attention import FeedForwardSwiGLU
    cfg, sd = _cfg_sd()
    p = "transformer.layers.0.ffn."
    m = FeedForwardSwiGLU(cfg.latent_dim, cfg.dim_feedforward, dropout=0.0, bias=False)
    m.load_state_dict(_sub(sd, p), strict=True)
    m.to(DEV).eval()
    x = _rand(2, 50, cfg.latent_dim, seed=1)
    ref = orc.swiglu(sd, p, x)
    got = m(x.to(DEV)).cpu()
    assert got.shape == ref.shape and rel_l2(got, ref) < 4e-3


def test_attention_layer_self_with_rope_and_mask():
    """AttentionLayer (encoder flavour): query_norm -> MHA(in_proj, QK-norm, RoPE tables, key-padding mask) -> FFN."""
    from renderformer.encodings.rope import TriangleRotaryEmbedding, freqs_to_cos_sin
    cfg, sd = _cfg_sd()
    m = _encoder_layer_module(cfg, sd, 1)
    B, N = 2, 72
    x = _rand(B, N, cfg.latent_dim, seed=2)
    pos = _rand(B, N, 9, seed=3, scale=0.5)
    mask = torch.ones(B, N, dtype=torch.bool)
    mask[1, 50:] = False
    rope = TriangleRotaryEmbedding(dim=cfg.vertex_pe_num_freqs)
    rope.load_state_dict({"freqs": sd["transformer.rope_emb.freqs"]})
    cos, sin = freqs_to_cos_sin(rope.get_triangle_freqs(pos), head_dim=128)
    rc, rs = orc.rope_tables(pos, sd["transformer.rope_emb.freqs"], 128)
    assert torch.allclose(cos, rc, atol=1e-6) and torch.allclose(sin, rs, atol=1e-6)
    ref = orc.encoder_layer(sd, "transformer.layers.1.", x, cfg.num_heads, mask, rc, rs)
    got = m(x.to(DEV), src_key_padding_mask=mask.to(DEV), rope_cos=cos.to(DEV), rope_sin=sin.to(DEV)).cpu()
    valid = mask[..., None].expand_as(ref)
    assert rel_l2(got[valid], ref[valid]) < 4e-3


def test_multi_head_attention_and_rope_helpers():
    from renderformer.encodings.rope import apply_rotary_emb_cossin, rotate_half_hf
    cfg, sd = _cfg_sd()
    m = _encoder_layer_module(cfg, sd, 0).multihead_attn
    B, N, d = 1, 40, cfg.latent_dim
    x = _rand(B, N, d, seed=4)
    ang = _rand(B, 1, N, 64, seed=5)
    cos, sin = torch.cat([ang.cos(), ang.cos()], -1), torch.cat([ang.sin(), ang.sin()], -1)
    p = "transformer.layers.0.multihead_attn."
    q, k, v = F.linear(x, sd[p + "in_proj.weight"]).chunk(3, dim=-1)
    q, k = orc.rms(q, sd[p + "q_norm.weight"], 1e-6), orc.rms(k, sd[p + "k_norm.weight"], 1e-6)
    H = cfg.num_heads
    hs = lambda t: t.view(B, N, H, 128).transpose(1, 2)  # noqa: E731
    qh, kh = orc.rope_rotate(hs(q), cos, sin), orc.rope_rotate(hs(k), cos, sin)
    att = F.scaled_dot_product_attention(qh, kh, hs(v)).transpose(1, 2).reshape(B, N, d)
    ref = F.linear(att, sd[p + "out_proj.weight"])
    got = m(x.to(DEV), x.to(DEV), x.to(DEV), rope_cos=cos.to(DEV), rope_sin=sin.to(DEV)).cpu()
    assert rel_l2(got, ref) < 6e-3
    # the rotation helper alone (kernel, rounded to 16 bit) vs the formula
    t = _rand(B, H, N, 128, seed=6)
    want = t * cos + rotate_half_hf(t) * sin
    rq, _ = apply_rotary_emb_cossin(t.to(DEV), t.to(DEV), cos.to(DEV), sin.to(DEV))
    assert rel_l2(rq.cpu(), want) < 3e-3


@pytest.mark.parametrize("shift", [0, 4])
def test_swin_self_attention(shift):
    from renderformer.layers.attention import SwinSelfAttention, get_swin_attn_mask
    cfg, sd = _cfg_sd()
    layer = 1 if shift else 0
    p = f"view_transformer.transformer.layers.{layer}.self_attn."
    dv = cfg.view_transformer_latent_dim
    m = SwinSelfAttention(dv, cfg.view_transformer_n_heads, window_size=8, shift_size=shift, bias=False, qk_norm=True,
                          norm_type="rms_norm")
    m.load_state_dict(_sub(sd, p), strict=True)
    m.to(DEV).eval()
    B, Hp = 2, 16
    x = _rand(B, Hp, Hp, dv, seed=7)
    ref = orc.swin_self_attention(sd, p, x.view(B, Hp * Hp, dv), cfg.view_transformer_n_heads, Hp, Hp, shift)
    got = m(x.to(DEV)).cpu().view(B, Hp * Hp, dv)
    assert rel_l2(got, ref) < 6e-3
    if shift:
        assert torch.equal(get_swin_attn_mask(Hp, Hp, 8, shift, "cpu"), orc.swin_region_mask(Hp, Hp, 8, shift, "cpu"))


def test_transformer_encoder_stack():
    from renderformer.layers.attention import TransformerEncoder
    cfg, sd = _cfg_sd()
    m = TransformerEncoder(num_layers=cfg.num_layers, num_heads=cfg.num_heads, hidden_dim=cfg.latent_dim,
                           ffn_hidden_dim=cfg.dim_feedforward, dropout=0.0, activation="swiglu", norm_type="rms_norm",
                           rope_dim=cfg.vertex_pe_num_freqs, bias=False, qk_norm=True)
    m.load_state_dict(_sub(sd, "transformer."), strict=True)
    m.to(DEV).eval()
    B, N = 2, 61  # not a multiple of 8
    x = _rand(B, N, cfg.latent_dim, seed=8)
    pos = _rand(B, N, 9, seed=9, scale=0.5)
    mask = torch.ones(B, N, dtype=torch.bool)
    mask[0, 40:] = False
    cos, sin = orc.rope_tables(pos, sd["transformer.rope_emb.freqs"], 128)
    ref = x
    for i in range(cfg.num_layers):
        ref = orc.encoder_layer(sd, f"transformer.layers.{i}.", ref, cfg.num_heads, mask, cos, sin)
    got = m(x.to(DEV), src_key_padding_mask=mask.to(DEV), triangle_pos=pos.to(DEV)).cpu()
    valid = mask[..., None].expand_as(ref)
    assert got.shape == ref.shape and rel_l2(got[valid], ref[valid]) < 6e-3


def _view_inputs(cfg, sd, n_tris=90, V=2, R=64):
    sc = make_scene(n_tris, V, seed=12, pad_to=96)
    N = sc["triangles"].shape[1]
    tex = sc["texture"].clone()
    tex[:, :, -3:] = torch.log10(tex[:, :, -3:] + 1.0)
    seq, key_mask = orc.scene_stage(sd, cfg, sc["triangles"].reshape(1, N, 9), tex, sc["mask"], sc["vn"].reshape(1, N, 9))
    tri_cam = orc.to_camera_space(sc["c2w"][0], sc["triangles"].expand(V, -1, -1, -1)).reshape(V, N, 9)
    rays = orc.camera_rays(sc["fov"][0] / 180.0 * math.pi, R)
    pos_k, _ = orc.centroid_positions(tri_cam, sc["mask"].expand(V, -1), cfg.num_register_tokens)
    return sc, seq.expand(V, -1, -1), key_mask.expand(V, -1), tri_cam, rays, pos_k


@pytest.mark.parametrize("cfg_name", ["tiny_swin", "tiny_full"])
def test_view_transformer_decoder_and_dpt_head(cfg_name):
    """ViewTransformer.forward end to end, and its two halves -- TransformerDecoder.forward (4 feature maps) and
    DPTHead.forward (pre-ELU head output) -- each against the oracle."""
    from renderformer.layers.attention import TransformerDecoder
    from renderformer.layers.dpt import DPTHead
    from renderformer.models.view_transformer import ViewTransformer
    cfg, sd = _cfg_sd(cfg_name)
    V, R = 2, 64
    sc, seq, key_mask, tri_cam, rays, pos_k = _view_inputs(cfg, sd, V=V, R=R)
    taps = {}
    ref_log = orc.view_stage(sd, cfg, seq, key_mask, tri_cam, sc["mask"].expand(V, -1), rays, taps)  # [V,3,R,R]
    vt = ViewTransformer(cfg)
    vt.load_state_dict(_sub(sd, "view_transformer."), strict=True)
    vt.to(DEV).eval()
    assert isinstance(vt.transformer, TransformerDecoder) and isinstance(vt.out_dpt, DPTHead)
    got = vt(torch.zeros(V, 3, device=DEV), rays.to(DEV), seq.to(DEV), pos_k.to(DEV), key_mask.to(DEV)).cpu()
    assert got.shape == ref_log.shape
    assert (got - ref_log).abs().max().item() < 2e-2 * (ref_log.max() - ref_log.min()).item()

    # decoder alone: ray tokens in, the four DPT feature maps out
    Hp = R // 8
    p = "view_transformer."
    tok = rays.view(V, Hp, 8, Hp, 8, 3).permute(0, 1, 3, 5, 2, 4).reshape(V, Hp * Hp, -1)
    x = sd[p + "ray_map_patch_token"] + orc.rms(F.linear(tok, sd[p + "ray_map_encoder.weight"], sd[p + "ray_map_encoder.bias"]),
                                                sd[p + "ray_map_encoder_norm.weight"], None)
    feats = vt.transformer(x.to(DEV), seq.to(DEV), src_key_padding_mask=key_mask.to(DEV), triangle_pos=pos_k.to(DEV),
                           ray_pos=torch.zeros(V, Hp * Hp, 9, device=DEV), out_layers=vt.out_layers, patch_h=Hp, patch_w=Hp)
    assert len(feats) == 4
    for f, r in zip(feats, taps["dec_feats"]):
        assert rel_l2(f[0].cpu(), r) < 8e-3

    # DPT head alone on the oracle's features
    raw = vt.out_dpt([[f.to(DEV)] for f in taps["dec_feats"]], Hp, Hp, patch_size=8).cpu()
    ref_raw = orc.dpt_head(sd, p + "out_dpt.", taps["dec_feats"], Hp, Hp, 8)
    assert raw.shape == ref_raw.shape
    assert (raw - ref_raw).abs().max().item() < 1e-2 * (ref_raw.max() - ref_raw.min()).item()


def test_ray_generator_transform_and_nerf_encoding():
    from renderformer.encodings.nerf_encoding import NeRFEncoding
    from renderformer.utils.ray_generator import RayGenerator
    from renderformer.utils.transform import trans_to_cam_coord
    sc = make_scene(33, 3, seed=2)
    c2w, fov = sc["c2w"], sc["fov"] / 180.0 * math.pi            # [1,3,4,4], [1,3,1] radians
    rays_o, rays_d = RayGenerator()(c2w.to(DEV), fov.to(DEV), 64)
    assert rays_d.shape == (1, 3, 64, 64, 3) and torch.equal(rays_o.cpu(), c2w[..., :3, 3])
    cam = orc.camera_rays(fov[0], 64)                               # camera-space rays
    want = torch.einsum("vij,vhwj->vhwi", c2w[0, :, :3, :3], cam)   # rotated into world space
    assert (rays_d[0].cpu() - want).abs().max().item() < 2e-6
    tri_rep = sc["triangles"].expand(3, -1, -1, -1).contiguous()
    vn_rep = sc["vn"].expand(3, -1, -1, -1).contiguous()
    tri_cam, c2w_id, vn_cam = trans_to_cam_coord(c2w[0].to(DEV), tri_rep.to(DEV), vn_rep.to(DEV))
    assert (tri_cam.cpu() - orc.to_camera_space(c2w[0], tri_rep)).abs().max().item() < 1e-5
    assert torch.equal(c2w_id.cpu(), torch.eye(4).repeat(3, 1, 1))
    Rt = c2w[0, :, :3, :3].transpose(1, 2)
    assert (vn_cam.cpu() - torch.einsum("vij,vntj->vnti", Rt, vn_rep)).abs().max().item() < 1e-5
    enc = NeRFEncoding(in_dim=9, num_frequencies=6, include_input=True)
    vn9 = sc["vn"].reshape(1, -1, 9)
    got = enc(vn9.to(DEV)).cpu()
    assert got.shape[-1] == enc.get_out_dim() == 117
    assert (got - orc.nerf_encode(vn9, 6)).abs().max().item() < 2e-3   # fp16 kernel output
    ident = NeRFEncoding(in_dim=3, num_frequencies=0, include_input=True)
    assert ident.get_out_dim() == 3 and torch.equal(ident(rays_d), rays_d)


def test_model_tree_helpers_and_strict_state_dict():
    """RenderFormer composes the same sub-modules as the reference: attribute paths exist, construct_seq /
    process_tri_vpos_list run on the kernels."""
    from renderformer import RenderFormer
    from renderformer.layers.attention import SwinSelfAttention, TransformerEncoder
    cfg, sd = _cfg_sd()
    model = RenderFormer(cfg)
    model.load_state_dict(sd, strict=True)
    model.to(DEV)
    assert isinstance(model.transformer, TransformerEncoder)
    assert isinstance(model.view_transformer.transformer.layers[1].self_attn, SwinSelfAttention)
    assert model.view_transformer.transformer.layers[1].self_attn.shift_size == 4
    sc = make_scene(40, 1, seed=3, pad_to=48)
    N = 48
    tri9 = sc["triangles"].reshape(1, N, 9)
    pos, mask = model.process_tri_vpos_list(tri9.to(DEV), sc["mask"].to(DEV))
    rpos, rmask = orc.centroid_positions(tri9, sc["mask"], cfg.num_register_tokens)
    assert torch.equal(mask.cpu(), rmask) and (pos.cpu() - rpos).abs().max().item() < 1e-5
    tex = sc["texture"].clone()
    tex[:, :, -3:] = torch.log10(tex[:, :, -3:] + 1.0)
    seq, mask2, pos2 = model.construct_seq(tri9.to(DEV), tex.to(DEV), sc["mask"].to(DEV), sc["vn"].reshape(1, N, 9).to(DEV))
    # oracle: the token construction part of scene_stage (before the encoder layers)
    vn_emb = orc.rms(F.linear(orc.nerf_encode(sc["vn"].reshape(1, N, 9), cfg.vn_pe_num_freqs), sd["vn_encoding_proj.weight"],
                              sd["vn_encoding_proj.bias"]), sd["vn_encoder_norm.weight"], None)
    tex_emb = orc.rms(F.linear(tex.reshape(1, N, -1), sd["texture_encoder.weight"], sd["texture_encoder.bias"]),
                      sd["texture_encoder_norm.weight"], None)
    ref_seq = torch.cat([sd["reg_tokens"], sd["tri_token"] + tex_emb + vn_emb], dim=1)
    valid = rmask[..., None].expand_as(ref_seq)
    assert seq.shape == ref_seq.shape and rel_l2(seq.cpu()[valid], ref_seq[valid]) < 3e-3
    assert torch.equal(mask2.cpu(), rmask) and (pos2.cpu() - rpos).abs().max().item() < 1e-5
