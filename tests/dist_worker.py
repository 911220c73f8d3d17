"""Worker of tests/test_dist_gpu.py::test_sharded_render_equals_single_gpu (launched by torch.distributed.run,
one rank per GPU, NCCL).  Every rank renders the whole scene alone on its own GPU (the single-GPU
schedule) and then takes part in the sharded render; rank 0 compares the gathered images, every rank
compares its own slice.  Prints DIST_WORKER_OK on rank 0."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    from renderformer_b200.config import RenderFormerConfig
    from renderformer_b200.dist import render_sharded, render_stream_sharded, view_slice
    from renderformer_b200.model import RenderFormer, RenderFormerRenderingPipeline
    from renderformer_b200.synth import init_state_dict, make_scene

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ok = True
    for cfg_name, n_tris, pad_to, views, res in (("tiny_swin", 100, 120, 5, 64), ("tiny_swin", 300, None, 2 * world, 128),
                                                 ("v1_1_swin_large", 700, None, world, 128)):
        cfg = RenderFormerConfig.named(cfg_name)
        model = RenderFormer(cfg)
        model.load_state_dict(init_state_dict(cfg, 7))
        pipe = RenderFormerRenderingPipeline(model)
        pipe.to(dev)
        # one view per decoder pass everywhere: the DPT picks its convolution kernel by the number of tiles in a
        # pass, so bit-identity needs the same view chunking in the single-GPU and in the sharded render
        pipe.view_chunk = 1
        host = make_scene(n_tris, views, seed=3, pad_to=pad_to)
        sc = {k: v.to(dev) for k, v in host.items()}
        single = pipe(sc["triangles"], sc["texture"], sc["mask"], sc["vn"], sc["c2w"], sc["fov"], resolution=res).clone()
        mine = view_slice(views, world, rank)
        for graphs in (False, True):
            pipe.cuda_graphs = graphs
            for rep in range(2):  # second call = graph replay
                out = render_sharded(pipe, sc["triangles"], sc["texture"], sc["mask"], sc["vn"], sc["c2w"], sc["fov"],
                                     resolution=res, dst=0)
                local_img = render_sharded(pipe, sc["triangles"], sc["texture"], sc["mask"], sc["vn"], sc["c2w"],
                                           sc["fov"], resolution=res, dst=None)
                same_local = torch.equal(local_img, single[:, mine])
                same_all = torch.equal(out, single) if rank == 0 else True
                if not (same_local and same_all):
                    d = (local_img - single[:, mine]).abs().max().item() if local_img.numel() else 0.0
                    print(f"[rank {rank}] {cfg_name} graphs={graphs} rep={rep}: sharded != single (max |d| {d:.3e})", flush=True)
                    ok = False
        # streaming form: host scenes in, own views out; own-rows-only texture upload
        pipe.cuda_graphs = True
        pinned = {k: v.pin_memory() for k, v in host.items()}
        n_scenes = 0
        for sl, img in render_stream_sharded(pipe, (pinned for _ in range(3)), resolution=res):
            n_scenes += 1
            if not torch.equal(img.to(dev), single[:, sl]):
                print(f"[rank {rank}] {cfg_name}: streamed views differ", flush=True)
                ok = False
        ok = ok and n_scenes == 3
        pipe.cuda_graphs = False
        del pipe, model
        torch.cuda.empty_cache()
    from renderformer_b200 import dist as rdist
    how = "nccl all-gather"
    if rdist._KV_STORES:
        how = "multicast stores" if next(iter(rdist._KV_STORES.values())).multicast else "peer stores"
    if rank == 0:
        print(f"k|v exchange of the row-sharded stage: {how} (RFB_KV_PUSH={os.environ.get('RFB_KV_PUSH', '2')})", flush=True)
    if os.environ.get("RFB_KV_PUSH", "2") != "0" and not rdist._KV_STORES:
        print(f"[rank {rank}] symmetric-memory store was requested but is not in use", flush=True)
        ok = False
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    good = flag.item() == 1
    if rank == 0 and good:
        print("DIST_WORKER_OK", flush=True)
    torch.cuda.synchronize()
    dist.barrier()
    sys.stdout.flush()
    os._exit(0 if good else 1)  # no communicator tear-down: see bench.py


if __name__ == "__main__":
    main()
