"""CPU: host-side logic -- config contract, state_dict key/shape contract, swin index maps,
FLOP formulas, drop-in import surface, checkpoint round trip."""
import json
import os

import pytest
import torch

from oracle import renderformer_oracle as orc
from renderformer_b200.config import RenderFormerConfig
from renderformer_b200.engine import swin_window_maps
from renderformer_b200.flops import dpt_flops, scene_flops, view_flops
from renderformer_b200.synth import init_state_dict, state_dict_shapes


def test_default_config_is_v1_base():
    c = RenderFormerConfig()
    assert (c.latent_dim, c.num_heads, c.num_layers, c.view_transformer_n_layers) == (768, 6, 12, 6)
    assert c.to_dict() == RenderFormerConfig.named("v1_base").to_dict()
    assert RenderFormerConfig.named("v1_1_swin_large").view_rope_dim == 12
    assert RenderFormerConfig.named("v1_1_swin_large").out_layers == [8, 9, 10, 11]


def test_unsupported_branches_raise():
    with pytest.raises(NotImplementedError):
        RenderFormerConfig(pe_type="nerf").check_supported()
    with pytest.raises(NotImplementedError):
        RenderFormerConfig(activation="gelu").check_supported()
    RenderFormerConfig.named("v1_1_swin_large").check_supported()
    RenderFormerConfig.named("v1_base").check_supported()


def test_param_counts():
    n = lambda name: sum(int(torch.Size(s).numel()) for s in state_dict_shapes(RenderFormerConfig.named(name)).values())
    assert n("v1_base") == 205173391
    assert n("v1_1_swin_large") == 483472079


@pytest.mark.parametrize("shift", [0, 4])
@pytest.mark.parametrize("hw", [(8, 8), (16, 16), (16, 24)])
def test_swin_maps_match_reference_partition(shift, hw):
    """perm/region reproduce roll + window_partition + get_swin_attn_mask of the oracle restatement."""
    H, W = hw
    perm, region = swin_window_maps(H, W, shift)
    tok = torch.arange(H * W).view(1, H, W, 1).float()
    g = torch.roll(tok, shifts=(-shift, -shift), dims=(1, 2)) if shift else tok
    win = g.view(1, H // 8, 8, W // 8, 8, 1).permute(0, 1, 3, 2, 4, 5).reshape(-1)
    assert torch.equal(win.long(), perm.long())
    assert sorted(perm.tolist()) == list(range(H * W))
    if shift:
        mask = orc.swin_region_mask(H, W, 8, shift, "cpu")  # [nW, 64, 64]
        r = region.view(-1, 64)
        assert torch.equal(mask, r[:, None, :] == r[:, :, None])
    else:
        assert int(region.max()) == 0


def test_flop_formulas_match_survey():
    c = RenderFormerConfig.named("v1_1_swin_large")
    assert abs(scene_flops(c, 4096) / 1e9 - (2599 + 207)) < 2
    assert abs(view_flops(c, 4096, 512) / 1e9 - 2936) < 2
    assert abs(dpt_flops(c, 512) / 1e9 - 237.8) < 0.2


def test_dropin_import_surface_and_checkpoint_roundtrip(tmp_path):
    import renderformer
    import renderformer_liger_kernel
    from renderformer import RenderFormer, RenderFormerRenderingPipeline
    from renderformer.models.config import RenderFormerConfig as C2
    assert C2 is RenderFormerConfig and callable(renderformer_liger_kernel.apply_kernels)
    cfg = RenderFormerConfig.named("tiny_full")
    m = RenderFormer(cfg)
    sd = init_state_dict(cfg, 3)
    m.load_state_dict(sd)
    assert set(m.state_dict().keys()) == set(state_dict_shapes(cfg).keys())
    m.save_pretrained(str(tmp_path))
    with open(tmp_path / "config.json") as f:
        assert json.load(f)["latent_dim"] == 384
    pipe = RenderFormerRenderingPipeline.from_pretrained(str(tmp_path))
    for k, v in pipe.model.state_dict().items():
        assert torch.equal(v, sd[k]), k
    assert pipe.to(torch.device("cpu")) is None and pipe.device.type == "cpu"


def test_encoder_key_split_rule():
    """Engine.enc_kv_split_tiles: chunk length of the key-split encoder attention.  It must be a function of the key
    count only (every schedule -- one GPU or N ranks -- has to do the same arithmetic per row), give at most 4
    chunks (rfb_attention accepts 8) and leave short sequences alone."""
    from renderformer_b200.engine import Engine
    for ntp in (8, 136, 1040, 2688):            # up to 21 key tiles: no split
        assert Engine.enc_kv_split_tiles(ntp) == 0
    for ntp, chunks in ((2920, 2), (4112, 3), (5648, 4), (8208, 4), (16400, 4)):
        kst = Engine.enc_kv_split_tiles(ntp)
        n_tiles = (ntp + 127) // 128
        assert kst > 0 and -(-n_tiles // kst) == chunks, (ntp, kst)
        assert kst * chunks >= n_tiles and kst * (chunks - 1) < n_tiles   # every chunk holds at least one tile


def test_precision_argument_follows_the_reference():
    """rendering_pipeline.py:98: bfloat16 / float16 / float32 are accepted, anything else fails the same assertion;
    float32 maps to fp16 operands and warns once (there is no fp32 tensor-core path)."""
    import warnings
    import pytest
    import torch
    from renderformer_b200 import model as M
    assert M.operand_dtype(torch.bfloat16) == torch.bfloat16 and M.operand_dtype(torch.float16) == torch.float16
    with pytest.raises(AssertionError, match="Invalid precision"):
        M.operand_dtype(torch.float64)
    M._warned_fp32[0] = False
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        assert M.operand_dtype(torch.float32) == torch.float16 and M.operand_dtype(torch.float32) == torch.float16
    assert len([x for x in w if "float32" in str(x.message)]) == 1


def test_no_cpu_fallback_errors_are_loud():
    """The product path has no CPU fallback: a pipeline on the CPU refuses to render instead of computing elsewhere."""
    import pytest
    import torch
    from renderformer_b200 import lib
    from renderformer_b200.config import RenderFormerConfig
    from renderformer_b200.model import RenderFormer, RenderFormerRenderingPipeline
    from renderformer_b200.synth import init_state_dict, make_scene
    cfg = RenderFormerConfig.named("tiny_swin")
    model = RenderFormer(cfg)
    model.load_state_dict(init_state_dict(cfg, 1))
    pipe = RenderFormerRenderingPipeline(model)
    sc = make_scene(8, 1, seed=0)
    with pytest.raises(lib.RfbError, match="CUDA"):
        pipe(sc["triangles"], sc["texture"], sc["mask"], sc["vn"], sc["c2w"], sc["fov"], resolution=64)
    with pytest.raises(lib.RfbError, match="CUDA"):
        next(pipe.render_stream([sc], resolution=64))
