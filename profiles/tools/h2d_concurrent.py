"""Pinned host -> device bandwidth with all ranks copying at once (what the frame uploads of an
N-GPU step see).  torchrun --nproc-per-node N profiles/tools/h2d_concurrent.py"""
import os
import time

import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", 0))
local = int(os.environ.get("LOCAL_RANK", 0))
world = int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
n = 1 << 30
host = torch.empty(n, dtype=torch.uint8, pin_memory=True)
host.fill_(7)
d = torch.empty(n, dtype=torch.uint8, device=dev)
for rep in range(4):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    d.copy_(host, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if rep:
        print("rank %d of %d: 1 GiB pinned H2D in %.1f ms = %.1f GB/s" % (rank, world, 1e3 * dt, n / dt / 1e9), flush=True)
if world > 1:
    dist.destroy_process_group()
