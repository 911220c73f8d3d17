#!/usr/bin/env python
"""Where does the image error of the CUDA path come from?  (GPU box; test infrastructure: uses the oracle as the checker; not on the product path.)

Runs the fp32 CPU oracle and the CUDA engine on one scene and compares stage by stage, then swaps stages:
  A  ours end to end
  B  our decoder + DPT on the ORACLE's encoder output        (isolates the view stage)
  C  our DPT on the ORACLE's decoder features                 (isolates the DPT head)
  D  the ORACLE's DPT on OUR decoder features                 (our transformer stacks without our DPT)
usage: python tests/diagnostics/error_budget.py [cbox | <n_tris>] [resolution] [config]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from oracle import renderformer_oracle as orc  # noqa: E402
from renderformer_b200.config import RenderFormerConfig  # noqa: E402
from renderformer_b200.metrics import hdr_rel_err, log_psnr, rel_l2  # noqa: E402
from renderformer_b200.model import RenderFormer, RenderFormerRenderingPipeline  # noqa: E402
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
    sd = init_state_dict(cfg, 7)
    taps = {}
    ref = orc.render(sd, cfg, sc["triangles"], sc["texture"], sc["mask"], sc["vn"], sc["c2w"], sc["fov"], R, taps=taps)
    ref_log = taps["log_img"]
    dev = torch.device("cuda:0")
    model = RenderFormer(cfg)
    model.load_state_dict(sd)
    pipe = RenderFormerRenderingPipeline(model)
    pipe.to(dev)
    eng = model.engine()
    g = {k: v.to(dev) for k, v in sc.items()}
    N = sc["triangles"].shape[1]
    Nt = N + cfg.num_register_tokens
    Hp = R // 8

    def report(tag, img):
        img = img.cpu()
        d = (img.double() - ref.double()).abs()
        i = int(d.argmax())
        print(f"{tag:58s} hdr rel {hdr_rel_err(img, ref):.3e}  log-PSNR {log_psnr(img, ref):5.1f} dB   "
              f"worst pixel: ref {ref.flatten()[i]:.3f} got {img.flatten()[i]:.3f}", flush=True)

    # A: end to end, with taps
    st = eng.encode_scene(g["triangles"], g["texture"], g["mask"], g["vn"])
    t = {}
    img = eng.render_views(st, 0, g["c2w"][0], g["fov"][0], R, taps=t)[None]
    print(f"seq relL2 {rel_l2(st.seq[:, :Nt], taps['seq']):.3e}")
    for i, (a, b) in enumerate(zip(t["dec_feats"], taps["dec_feats"])):
        print(f"dec_feat[{i}] relL2 {rel_l2(a, b):.3e}   max|d| {(a.cpu() - b).abs().max():.3e}  (max|ref| {b.abs().max():.2f})")
    report("A  ours end to end", img)

    # B: our view stage on the oracle's tokens
    key_mask = torch.cat([torch.ones((1, cfg.num_register_tokens), dtype=torch.bool), sc["mask"]], dim=1)
    stB = eng.scene_state_from_tokens(taps["seq"].to(dev), None, key_mask.to(dev))
    stB.tri = st.tri
    tB = {}
    imgB = eng.render_views(stB, 0, g["c2w"][0], g["fov"][0], R, taps=tB)[None]
    for i, (a, b) in enumerate(zip(tB["dec_feats"], taps["dec_feats"])):
        print(f"   B dec_feat[{i}] relL2 {rel_l2(a, b):.3e}")
    report("B  our decoder + DPT on the oracle's encoder output", imgB)

    # C: our DPT on the oracle's decoder features
    feats16 = [f.reshape(-1, f.shape[-1]).to(dev, torch.float16).contiguous() for f in taps["dec_feats"]]
    imgC = eng._dpt(feats16, 1, Hp, Hp)[None]
    report("C  our DPT head on the oracle's decoder features", imgC)
    rawC = eng._dpt(feats16, 1, Hp, Hp, raw_out=True)[None].cpu()
    refraw = orc.dpt_head(sd, "view_transformer.out_dpt.", taps["dec_feats"], Hp, Hp, 8).permute(0, 2, 3, 1)[None]
    print(f"   C pre-ELU head output: max|d| {(rawC - refraw).abs().max():.3e}  relL2 {rel_l2(rawC, refraw):.3e}  (range {refraw.min():.3f}..{refraw.max():.3f})")

    # D: the oracle's DPT on our decoder features
    ours = [f.cpu().float() for f in t["dec_feats"]]
    logD = F.elu(orc.dpt_head(sd, "view_transformer.out_dpt.", ours, Hp, Hp, 8), alpha=1e-3).permute(0, 2, 3, 1)[None]
    report("D  the oracle's DPT head on OUR decoder features", torch.pow(10.0, logD) - 1.0)
    logDB = F.elu(orc.dpt_head(sd, "view_transformer.out_dpt.", [f.cpu().float() for f in tB["dec_feats"]], Hp, Hp, 8),
                  alpha=1e-3).permute(0, 2, 3, 1)[None]
    report("E  the oracle's DPT head on decoder features of run B", torch.pow(10.0, logDB) - 1.0)
    del ref_log


if __name__ == "__main__":
    main()
