#!/usr/bin/env python
"""Which operator of the row-sharded schedule stops being bit-identical to the single-GPU schedule?  (ONE GPU,
development tool.)  Layer 0 of the encoder is run on all rows and on one rank's rows, operator by operator, each
time from the SAME inputs, and the outputs are compared bit for bit.
usage: python tools/debug_shard.py [--tris 2900] [--world 4] [--rank 0]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from renderformer_b200 import lib as L  # noqa: E402
from renderformer_b200 import ops  # noqa: E402
from renderformer_b200.config import RenderFormerConfig  # noqa: E402
from renderformer_b200.engine import EPS, RowShard  # noqa: E402
from renderformer_b200.model import RenderFormer, RenderFormerRenderingPipeline  # noqa: E402
from renderformer_b200.synth import init_state_dict, make_scene  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--tris", type=int, default=2900)
ap.add_argument("--world", type=int, default=4)
ap.add_argument("--rank", type=int, default=0)
ap.add_argument("--config", default="v1_1_swin_large")
a = ap.parse_args()
cfg = RenderFormerConfig.named(a.config)
model = RenderFormer(cfg)
model.load_state_dict(init_state_dict(cfg, 5))
pipe = RenderFormerRenderingPipeline(model)
pipe.to(torch.device("cuda:0"))
eng = model.engine()
w = eng.w
sc = {k: v.cuda() for k, v in make_scene(a.tris, 1, seed=11).items()}
x, pos, bits, words, (B, N, Nt, Ntp), tri, mask_u8 = eng.construct_seq(sc["triangles"], sc["texture"], sc["mask"], sc["vn"])
d, H = cfg.latent_dim, cfg.num_heads
P = d // 128
bf, f32 = eng.op, torch.float32
nrm = dict(norm_dim=d, norm_eps=EPS)
sh = RowShard(a.rank, a.world, None)
r0, r1 = sh.my_rows(Ntp)
rows = r1 - r0
print(f"Ntp {Ntp}, rank {a.rank}/{a.world}: rows [{r0}, {r1})  key tiles {(Ntp + 127) // 128}  kv_split_tiles {eng.enc_kv_split_tiles(Ntp)}")


def E(shape, dt):
    return torch.empty(shape, dtype=dt, device="cuda")


def same(name, full, part):
    ok = torch.equal(full, part)
    extra = "" if ok else f"   max |d| {(full.float() - part.float()).abs().max().item():.3e}, {int((full != part).sum())} of {full.numel()} differ"
    print(f"  {'ok  ' if ok else 'DIFF'} {name}{extra}", flush=True)


o = "enc0."
# ---- all rows (the single-GPU schedule's operators)
xb, xsq = E((Ntp, d), bf), E((Ntp, P), f32)
ops.rowstat(x, xb, xsq, rows=Ntp, d=d)
vt = E((1, d, Ntp), bf)
qk = ops.gemm(xb, w[o + "wqkv"], out=E((Ntp, 2 * d), f32), in_sumsq=xsq, vt_out=vt, vt_split=2 * d, vt_rows_per_batch=Ntp, **nrm)
qkr = ops.qknorm_rope(qk, w[o + "qkn"], E((Ntp, 2 * d), bf), rows=Ntp, d=d, nseg=2, ldx=2 * d, ldo=2 * d, pos=pos, freqs=w["enc.freqs"], eps=EPS)
kst = eng.enc_kv_split_tiles(Ntp)
ws = E((max(ops.attention_ws_elems(1, H, Ntp, Ntp, kst), 1),), f32)
att = E((Ntp, d), bf)
ops.attention(qkr, qkr[:, d:], vt, att, B=1, H=H, Nq=Ntp, Nk=Ntp, ldq=2 * d, ldk=2 * d, ldvt=Ntp, ldo=d, mask_bits=bits, mask_bs=words,
              kv_split_tiles=kst, split_ws=ws)
att_nosplit = E((Ntp, d), bf)
ops.attention(qkr, qkr[:, d:], vt, att_nosplit, B=1, H=H, Nq=Ntp, Nk=Ntp, ldq=2 * d, ldk=2 * d, ldvt=Ntp, ldo=d, mask_bits=bits, mask_bs=words)
x1 = x.clone()
xb2, xsq2 = E((Ntp, d), bf), E((Ntp, P), f32)
ops.gemm(att, w[o + "wo"], out=x1, res1=x1, out_sumsq=xsq2, out16=xb2)
g = ops.gemm(xb2, w[o + "w13"], epi=L.EPI_SWIGLU, out_dtype=bf, in_sumsq=xsq2, **nrm)
x2 = x1.clone()
xb3, xsq3 = E((Ntp, d), bf), E((Ntp, P), f32)
ops.gemm(g, w[o + "w2"], out=x2, res1=x2, out_sumsq=xsq3, out16=xb3)
torch.cuda.synchronize()
print(f"split vs no split (tolerance only): max |d| {(att.float() - att_nosplit.float()).abs().max().item():.3e}")

# ---- one rank's rows, every operator from the all-rows inputs
print("own rows, operator by operator:")
qkv32 = E((rows, 3 * d), f32)
ops.gemm(xb[r0:r1], w[o + "wqkv"], out=qkv32, in_sumsq=xsq[r0:r1], **nrm)
same("fused qkv GEMM: q|k fp32", qk[r0:r1], qkv32[:, :2 * d])
kv = torch.zeros((Ntp, 2 * d), dtype=bf, device="cuda")
qr = E((rows, d), bf)
ops.qkv_post(qkv32, w[o + "qkn"], qr, [kv.data_ptr()], ldkv=2 * d, row0=r0, rows=rows, d=d, pos=pos[0, r0:r1], freqs=w["enc.freqs"], eps=EPS)
same("qkv_post: q", qkr[r0:r1, :d], qr)
same("qkv_post: k", qkr[r0:r1, d:], kv[r0:r1, :d])
same("qkv_post: v", vt[0].t()[r0:r1], kv[r0:r1, d:])
kv[:, :d].copy_(qkr[:, d:])
kv[:, d:].copy_(vt[0].t())
vt2 = E((1, d, Ntp), bf)
ops.transpose16(kv[:Ntp, d:], vt2[0], rows=Ntp, cols=d)
same("transpose16", vt, vt2)
ws2 = E((max(ops.attention_ws_elems(1, H, rows, Ntp, kst), 1),), f32)
att_p = E((rows, d), bf)
ops.attention(qkr[r0:r1], kv, vt2, att_p, B=1, H=H, Nq=rows, Nk=Ntp, ldq=2 * d, ldk=2 * d, ldvt=Ntp, ldo=d, mask_bits=bits, mask_bs=words,
              kv_split_tiles=kst, split_ws=ws2)
same("attention (key split)", att[r0:r1], att_p)
att_p2 = E((rows, d), bf)
ops.attention(qkr[r0:r1], kv, vt2, att_p2, B=1, H=H, Nq=rows, Nk=Ntp, ldq=2 * d, ldk=2 * d, ldvt=Ntp, ldo=d, mask_bits=bits, mask_bs=words)
same("attention (no split)", att_nosplit[r0:r1], att_p2)
xp = x[r0:r1].clone()
xb2p, xsq2p = E((rows, d), bf), E((rows, P), f32)
ops.gemm(att[r0:r1], w[o + "wo"], out=xp, res1=xp, out_sumsq=xsq2p, out16=xb2p, M=rows)
same("wo: x", x1[r0:r1], xp)
same("wo: 16-bit copy", xb2[r0:r1], xb2p)
same("wo: sums of squares", xsq2[r0:r1], xsq2p)
gp = ops.gemm(xb2[r0:r1], w[o + "w13"], epi=L.EPI_SWIGLU, out_dtype=bf, in_sumsq=xsq2[r0:r1], **nrm)
same("w13 SwiGLU", g[r0:r1], gp)
xp2 = x1[r0:r1].clone()
xb3p, xsq3p = E((rows, d), bf), E((rows, P), f32)
ops.gemm(g[r0:r1], w[o + "w2"], out=xp2, res1=xp2, out_sumsq=xsq3p, out16=xb3p, M=rows)
same("w2: x", x2[r0:r1], xp2)
same("w2: 16-bit copy", xb3[r0:r1], xb3p)
same("w2: sums of squares", xsq3[r0:r1], xsq3p)

# ---- run-to-run determinism of the attention kernels (a timing-dependent race shows up here)
for name, kw, ref_out in (("key split", dict(kv_split_tiles=kst, split_ws=ws2), att_p.clone()), ("no split", {}, att_p2.clone())):
    bad = 0
    for _ in range(40):
        out = E((rows, d), bf)
        ops.attention(qkr[r0:r1], kv, vt2, out, B=1, H=H, Nq=rows, Nk=Ntp, ldq=2 * d, ldk=2 * d, ldvt=Ntp, ldo=d, mask_bits=bits,
                      mask_bs=words, **kw)
        bad += 0 if torch.equal(out, ref_out) else 1
    print(f"  attention ({name}), own rows: {bad} of 40 repeats differ from the first run")
for name, kw, ref_out in (("key split", dict(kv_split_tiles=kst, split_ws=ws), att.clone()), ("no split", {}, att_nosplit.clone())):
    bad = 0
    for _ in range(40):
        out = E((Ntp, d), bf)
        ops.attention(qkr, qkr[:, d:], vt, out, B=1, H=H, Nq=Ntp, Nk=Ntp, ldq=2 * d, ldk=2 * d, ldvt=Ntp, ldo=d, mask_bits=bits,
                      mask_bs=words, **kw)
        bad += 0 if torch.equal(out, ref_out) else 1
    print(f"  attention ({name}), all rows: {bad} of 40 repeats differ from the first run")
