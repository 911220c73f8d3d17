#!/usr/bin/env python
"""All-gather / broadcast timing probe (torchrun): what transport NCCL picked and what one per-layer gather costs.
usage: NCCL_DEBUG=INFO python -m torch.distributed.run --nproc-per-node N tools/nccl_probe.py"""
import os
import sys

import torch
import torch.distributed as dist

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank)))
dist.init_process_group("nccl", device_id=dev)
for mb in (0.5, 2.0, 8.6, 34.0, 320.0):
    n = int(mb * 1e6 / 2) // (8 * world) * 8 * world
    full = torch.empty(n, dtype=torch.bfloat16, device=dev)
    chunk = full[rank * (n // world):(rank + 1) * (n // world)]
    for _ in range(5):
        dist.all_gather_into_tensor(full, chunk)
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        dist.all_gather_into_tensor(full, chunk)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    # same thing replayed from a CUDA graph
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        dist.all_gather_into_tensor(full, chunk)
    torch.cuda.current_stream().wait_stream(s)
    with torch.cuda.graph(g, capture_error_mode="thread_local"):
        for _ in range(20):
            dist.all_gather_into_tensor(full, chunk)
    g.replay()
    torch.cuda.synchronize()
    dist.barrier()
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    usg = e0.elapsed_time(e1) / 20 * 1e3
    if rank == 0:
        print(f"all_gather {mb:6.1f} MB total over {world} ranks: eager {us:7.1f} us, in-graph {usg:7.1f} us "
              f"({n * 2 * (world - 1) / world / usg / 1e3:.0f} GB/s received per rank)", flush=True)
torch.cuda.synchronize()
dist.barrier()
sys.stdout.flush()
os._exit(0)
