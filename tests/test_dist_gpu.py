"""GPU: the multi-GPU data plane.

* `test_row_sharded_encoder_emulated`: ONE GPU.  The row-sharded view-independent stage of every rank
  of a world of W is run in turn; its all-gathers are replaced by the per-layer K / V (and the final 16-bit
  stream) recorded from the single-GPU schedule, and what the rank contributes is compared bit for bit
  with the recorded rows.
  Rank r's arithmetic never sees which device computed the other rows, so this pins the sharded
  schedule without needing W devices.
* `test_sharded_render_equals_single_gpu`: TWO (or more) GPUs, `torch.distributed.run`, NCCL.  The
  row-sharded + view-sharded render must equal (`torch.equal`) the single-GPU render of the same
  scene, eager and replayed from a CUDA graph (NCCL all-gathers inside the graph); the streaming form
  must agree as well.  Skipped when the box has one GPU (the driver's scaling run covers N > 1 then).
"""
import os
import socket
import subprocess
import sys

import pytest
import torch

from renderformer_b200.config import RenderFormerConfig
from renderformer_b200.synth import init_state_dict, make_scene

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _pipe(cfg, seed, dev="cuda:0"):
    from renderformer_b200.model import RenderFormer, RenderFormerRenderingPipeline
    model = RenderFormer(cfg)
    model.load_state_dict(init_state_dict(cfg, seed))
    pipe = RenderFormerRenderingPipeline(model)
    pipe.to(torch.device(dev))
    return pipe


@pytest.mark.parametrize("use_store", [False, True], ids=["all-gather", "kv-store"])
@pytest.mark.parametrize("cfg_name,n_tris,pad_to,world", [
    ("tiny_swin", 100, None, 2),
    ("tiny_swin", 37, 48, 4),          # ragged count, padding, more ranks than 128-row tiles
    ("tiny_swin", 5, None, 8),         # fewer rows than ranks x 8: trailing ranks own nothing
    ("v1_1_swin_large", 1000, 1024, 8),
    ("v1_1_swin_large", 2900, None, 4),  # 23 key tiles: the key-split attention (2 chunks) in both schedules
])
def test_row_sharded_encoder_emulated(cfg_name, n_tris, pad_to, world, use_store):
    from renderformer_b200.engine import RowShard
    cfg = RenderFormerConfig.named(cfg_name)
    pipe = _pipe(cfg, 5)
    eng = pipe.model.engine()
    sc = {k: v.cuda() for k, v in make_scene(n_tris, 1, seed=11, pad_to=pad_to).items()}
    taps = {}
    ref = eng.encode_scene(sc["triangles"], sc["texture"], sc["mask"], sc["vn"], taps=taps)
    xb_fin, xsq_fin = taps["enc_stream"][-1]        # final 16-bit stream + row sums (input of the K/V hoist)
    enc_kv = taps["enc_kv"]                         # per layer: K [Ntp,d] after QK-norm + RoPE, V^T [1,d,Ntp]
    d = cfg.latent_dim
    P = d // 128
    assert len(enc_kv) == cfg.num_layers
    for rank in range(world):
        step = [0]

        def fake_all_gather(full, chunk, rank=rank):
            S = full.shape[0] // world
            r0, r1 = min(rank * S, ref.Ntp), min((rank + 1) * S, ref.Ntp)
            if full.dtype == torch.float32:         # the optional fp32 token gather
                mine = full[rank * S:(rank + 1) * S].clone()
                full[:ref.Ntp].copy_(ref.seq[0])
                assert torch.equal(mine[:r1 - r0], ref.seq[0, r0:r1]), f"rank {rank}: own fp32 rows differ"
                return
            if step[0] < cfg.num_layers:            # per-layer gather of the ranks' [k | v] rows
                k, vt = enc_kv[step[0]]
                v = vt[0].t()                       # [Ntp, d]
                assert full.shape[1] == 2 * d
                assert torch.equal(full[r0:r1, :d], k[r0:r1]), f"rank {rank} layer {step[0]}: own K rows differ"
                assert torch.equal(full[r0:r1, d:], v[r0:r1]), f"rank {rank} layer {step[0]}: own V rows differ"
                full[:ref.Ntp, :d].copy_(k)         # what the other ranks would have contributed
                full[:ref.Ntp, d:].copy_(v)
            else:                                   # final gather: 16-bit stream + row sums of squares
                f32 = full.view(torch.float32)
                assert torch.equal(full[r0:r1, :d], xb_fin[r0:r1]), f"rank {rank}: final 16-bit rows differ"
                assert torch.equal(f32[r0:r1, d // 2:d // 2 + P], xsq_fin[r0:r1]), f"rank {rank}: final row sums differ"
                full[:ref.Ntp, :d].copy_(xb_fin)
                f32[:ref.Ntp, d // 2:d // 2 + P].copy_(xsq_fin)
            step[0] += 1

        class FakeStore:  # stands in for dist.SymmKVStore: the barrier "receives" the other ranks' rows
            def __init__(self, rows, width, dtype, device):
                self.t = torch.zeros((2, rows, width), dtype=dtype, device=device)

            def buf(self, i):
                return self.t[i]

            def dst(self, i):
                return [self.t[i].data_ptr()], False

            def barrier(self):
                fake_all_gather(self.t[step[0] & 1], None)

        stores = {}

        def kv_store(rows, width, dtype, device):
            return stores.setdefault((rows, width, dtype), FakeStore(rows, width, dtype, device))

        sh = RowShard(rank, world, fake_all_gather, kv_store if use_store else None)
        st = eng.encode_scene(sc["triangles"], sc["texture"], sc["mask"], sc["vn"], shard=sh, gather_seq=True)
        assert step[0] == cfg.num_layers + 1 and st.complete
        for a, b, name in zip(st.tensors(), ref.tensors(), ("seq", "tri", "mask", "bits", "k_all", "v_all")):
            assert a.shape == b.shape, name
            assert torch.equal(a, b), f"rank {rank}: SceneState.{name} differs from the single-GPU schedule"
        # own-rows-only texture upload
        t0, t1 = eng.own_triangles(sc["triangles"].shape[1], sh)
        step[0] = 0
        st2 = eng.encode_scene(sc["triangles"], sc["texture"][:, t0:t1].contiguous(), sc["mask"], sc["vn"], shard=sh,
                               texture_own_rows=True)
        assert torch.equal(st2.k_all, ref.k_all) and torch.equal(st2.v_all, ref.v_all) and not (world > 1 and st2.complete)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("kv_push", ["2", "1", "0"], ids=["multicast-stores", "peer-stores", "nccl-all-gather"])
def test_sharded_render_equals_single_gpu(kv_push):
    """`kv_push` = RFB_KV_PUSH: how a layer's k | v rows reach the other ranks -- stores through the NVLS multicast
    address fused into the producing kernel (default), plain stores into every peer's mapped memory, or the NCCL
    all-gather."""
    n = 2  # two ranks exercise every code path; bench.py re-checks bit-identity at whatever N it runs on
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tests", "dist_worker.py")]
    env = dict(os.environ, NCCL_DEBUG="WARN", RFB_KV_PUSH=kv_push)
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=420, env=env, cwd=ROOT)
    tail = (r.stdout[-3000:] + "\n" + r.stderr[-3000:])
    assert r.returncode == 0, tail
    assert "DIST_WORKER_OK" in r.stdout, tail
