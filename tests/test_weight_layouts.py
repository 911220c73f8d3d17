"""CPU: the one-time weight re-layouts of `Engine._prepare` against the reference's math in plain
torch fp32.  Each check restates what the consuming KERNEL does with the re-laid-out matrix (the contract in
include/rfb200.h) and compares with the reference formulation the oracle follows -- so a wrong permutation,
fold or padding is caught here, without a GPU.  The engine object is built without its CUDA check for this."""
import torch
import torch.nn.functional as F

from renderformer_b200.config import RenderFormerConfig
from renderformer_b200.engine import EPS, Engine
from renderformer_b200.synth import init_state_dict


def _cpu_engine(name="tiny_swin", seed=3):
    cfg = RenderFormerConfig.named(name)
    sd = {k: v.float() for k, v in init_state_dict(cfg, seed).items()}
    eng = Engine.__new__(Engine)  # no CUDA device here: only the layout code is exercised
    eng.cfg, eng.op, eng.parts, eng.device = cfg, torch.float32, ("tokens", "ray", "encoder", "decoder", "dpt"), torch.device("cpu")
    eng.fused_dec = bool(cfg.view_transformer_use_swin_attn)
    eng.w, eng._maps = {}, {}
    # op = fp32 keeps the transformer "operands" unrounded, so the comparison sees the layout, not 16-bit rounding
    # (the token encoders and the DPT head are fp16 in every mode: those checks round the reference weight alike)
    eng._prepare(sd)
    return cfg, sd, {k: v.float() for k, v in eng.w.items()}


def _rms(x, w=None, eps=EPS):
    y = x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps)
    return y if w is None else y * w


def test_norm_weight_folds_into_the_next_projection():
    """RMSNorm(x) W^T = r * (x (W . w)^T), r = rsqrt(mean(x^2) + eps): the kernels multiply by r in the epilogue
    (in_sumsq), the norm weight w lives inside the re-laid-out matrix (layers/attention.py:496-503)."""
    cfg, sd, w = _cpu_engine()
    x = torch.randn(37, cfg.latent_dim)
    r = torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + EPS)
    p = "transformer.layers.1."
    want = F.linear(_rms(x, sd[p + "query_norm.weight"]), sd[p + "multihead_attn.in_proj.weight"])
    got = r * F.linear(x, w["enc1.wqkv"])
    assert torch.allclose(got, want, atol=2e-5, rtol=1e-4)
    # decoder: q side folded only in the fused (swin) schedule, K / V of all layers concatenated [K_0..K_L | V_0..V_L]
    dv, Lv = cfg.view_transformer_latent_dim, cfg.view_transformer_n_layers
    assert w["dec.wkv_all"].shape == (2 * Lv * dv, cfg.latent_dim)
    for layer in (0, Lv - 1):
        p = f"view_transformer.transformer.layers.{layer}."
        ctx = _rms(x, sd[p + "kv_norm.weight"])
        kv = r * F.linear(x, w["dec.wkv_all"])
        assert torch.allclose(kv[:, layer * dv:(layer + 1) * dv], F.linear(ctx, sd[p + "multihead_attn.k_proj.weight"]), atol=2e-5, rtol=1e-4)
        assert torch.allclose(kv[:, (Lv + layer) * dv:(Lv + layer + 1) * dv], F.linear(ctx, sd[p + "multihead_attn.v_proj.weight"]), atol=2e-5, rtol=1e-4)
    assert torch.equal(w["dec.kn_all"], torch.cat([sd[f"view_transformer.transformer.layers.{i}.multihead_attn.k_norm.weight"] for i in range(Lv)]))


def test_swiglu_interleave_matches_the_epilogue_contract():
    """RFB_EPI_SWIGLU (include/rfb200.h): W rows interleaved [16 gate | 16 up] per 32, out[:, n/2] = silu(g) * u.
    With the folded ffn_norm this must equal w2(silu(w1 n(x)) * w3 n(x)) (layers/attention.py:56-57)."""
    cfg, sd, w = _cpu_engine()
    x = torch.randn(29, cfg.latent_dim)
    r = torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + EPS)
    p = "transformer.layers.0."
    acc = r * F.linear(x, w["enc0.w13"])                       # what the GEMM accumulates, rows scaled by r
    acc = acc.view(x.shape[0], -1, 2, 16)                      # 32-column groups: gate half, up half
    h = (F.silu(acc[:, :, 0]) * acc[:, :, 1]).reshape(x.shape[0], -1)
    xn = _rms(x, sd[p + "ffn_norm.weight"])
    want_h = F.silu(F.linear(xn, sd[p + "ffn.w1.weight"])) * F.linear(xn, sd[p + "ffn.w3.weight"])
    assert h.shape == want_h.shape and torch.allclose(h, want_h, atol=2e-5, rtol=1e-4)
    assert torch.allclose(F.linear(h, w["enc0.w2"]), F.linear(want_h, sd[p + "ffn.w2.weight"]), atol=2e-5, rtol=1e-4)


def test_conv_weights_follow_the_implicit_gemm_tap_order():
    """3x3 conv as a GEMM over K = (ky*3 + kx)*Ci + ci with NHWC activations (layers/dpt.py:57-92)."""
    cfg, sd, w = _cpu_engine()
    key = "view_transformer.out_dpt.scratch.refinenet2.resConvUnit1.conv1."
    wt, b = sd[key + "weight"].half().float(), sd[key + "bias"]  # the DPT head's operands are fp16
    Ci = wt.shape[1]
    x = torch.randn(2, Ci, 6, 5)
    want = F.conv2d(x, wt, b, padding=1)
    xp = F.pad(x.permute(0, 2, 3, 1), (0, 0, 1, 1, 1, 1))      # NHWC, zero halo (TMA's out-of-bounds fill)
    cols = torch.cat([xp[:, ky:ky + 6, kx:kx + 5] for ky in range(3) for kx in range(3)], dim=-1)
    got = (F.linear(cols, w["dpt.rf2.u1c1.w"]) + w["dpt.rf2.u1c1.b"]).permute(0, 3, 1, 2)
    assert torch.allclose(got, want, atol=2e-5, rtol=1e-4)
    # stride-2 conv of resize_layers.3 through rfb_im2col_s2: output pixel (y, x) reads input (2y+ky-1, 2x+kx-1)
    wt, b = sd["view_transformer.out_dpt.resize_layers.3.weight"].half().float(), sd["view_transformer.out_dpt.resize_layers.3.bias"]
    x = torch.randn(1, wt.shape[1], 8, 8)
    want = F.conv2d(x, wt, b, stride=2, padding=1)
    xp = F.pad(x.permute(0, 2, 3, 1), (0, 0, 1, 1, 1, 1))
    cols = torch.cat([xp[:, ky:ky + 8:2, kx:kx + 8:2] for ky in range(3) for kx in range(3)], dim=-1)
    got = (F.linear(cols, w["dpt.down3.w"]) + w["dpt.down3.b"]).permute(0, 3, 1, 2)
    assert torch.allclose(got, want, atol=2e-5, rtol=1e-4)


def test_conv_transpose_is_a_gemm_plus_pixel_shuffle():
    """ConvTranspose2d(kernel = stride = s) (layers/dpt.py:195-206): GEMM with W[(i*s+j)*Co + co, ci], bias
    repeated s*s times, then rfb_pixel_shuffle: column (i*s+j)*Co + co of pixel (y, x) -> out[y*s+i, x*s+j, co]."""
    cfg, sd, w = _cpu_engine()
    for idx, s in ((0, 4), (1, 2)):
        wt, b = sd[f"view_transformer.out_dpt.resize_layers.{idx}.weight"].half().float(), sd[f"view_transformer.out_dpt.resize_layers.{idx}.bias"]
        Ci, Co = wt.shape[:2]
        x = torch.randn(2, Ci, 3, 4)
        want = F.conv_transpose2d(x, wt, b, stride=s)
        y = F.linear(x.permute(0, 2, 3, 1), w[f"dpt.up{idx}.w"]) + w[f"dpt.up{idx}.b"]  # [B, h, w, s*s*Co]
        y = y.view(2, 3, 4, s, s, Co).permute(0, 1, 3, 2, 4, 5).reshape(2, 3 * s, 4 * s, Co)
        assert torch.allclose(y.permute(0, 3, 1, 2), want, atol=2e-5, rtol=1e-4)


def test_constant_texture_weights_and_padding():
    """Constant-texture fast path: a texture that is 13 constants times the triangular texel mask
    x + y <= P (scene_processor/to_h5.py:42-45) projects through the texel-summed weight [d, 13] (K padded to 16);
    vn weight K-padded 117 -> 128 with zeros."""
    cfg, sd, w = _cpu_engine()
    P, C = cfg.texture_encode_patch_size, cfg.texture_channels
    ii = torch.arange(P)
    tmask = (ii[:, None] + ii[None, :] <= P).float()
    consts = torch.rand(7, C)
    full = (consts[:, :, None, None] * tmask).reshape(7, -1)
    want = F.linear(full, sd["texture_encoder.weight"])
    got = F.linear(F.pad(consts, (0, w["tex.wr"].shape[1] - C)), w["tex.wr"])
    # the summed weight is rounded to fp16 once (<= 2^-11 relative per entry), the texel weights are not summed in fp16
    assert w["tex.wr"].shape[1] % 16 == 0 and (got - want).abs().max() <= 2e-3 * want.abs().max()
    vw = sd["vn_encoding_proj.weight"].half().float()
    assert w["vn.w"].shape[1] % 64 == 0 and torch.equal(w["vn.w"][:, :vw.shape[1]], vw) and not w["vn.w"][:, vw.shape[1]:].any()


def test_full_attention_decoder_keeps_explicit_norms():
    """V1-Base (full ray self-attention): q / self-attention / ffn norms stay explicit kernels, so those matrices
    are NOT folded; the hoisted K / V still are."""
    cfg, sd, w = _cpu_engine("tiny_full")
    p = "view_transformer.transformer.layers.0."
    dv = cfg.view_transformer_latent_dim
    assert torch.equal(w["dec0.wq"], sd[p + "multihead_attn.q_proj.weight"])
    assert torch.equal(torch.cat([w["dec0.s.wqk"], w["dec0.s.wv"]]), sd[p + "self_attn.in_proj.weight"])
    assert torch.allclose(w["dec.wkv_all"][:dv], sd[p + "multihead_attn.k_proj.weight"] * sd[p + "kv_norm.weight"][None, :])
