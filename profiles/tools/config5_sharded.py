"""BASELINE config 5 without the linker: find + refine on ONE dense-cluster 2D video sharded over the
ranks of one box, through the public API, gather on rank 0 inside the timed region.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \\
        profiles/tools/config5_sharded.py [total_frames]

Every rank owns total_frames / N consecutive frames (rendered on its device, copied to pinned host
memory: the timed region starts from HOST frames), runs ``find_features`` on them (preprocess and
characterize on host cores, the maxima search on its GPU) and passes the table to
``parallel.refine_leastsq_sharded(presharded=True, gather='root')``.  The sequential linking step
of ``find_link`` is out of scope."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import bench
import clustertracking_b200 as ctb
from clustertracking_b200 import artificial, parallel

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
device = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=device)
total = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
n_frames = total // world
first = rank * n_frames
pos, frame, signal, start = bench.video_geometry(n_frames, seed=7 + rank)
d_stack = bench.render_video_torch(pos, frame, signal, n_frames, device, seed=100 + rank)
host = torch.empty(d_stack.shape, dtype=torch.uint8, pin_memory=True)
host.copy_(d_stack); torch.cuda.synchronize()
del d_stack
reader = artificial.FrameStack(host.numpy(), first_frame=first)


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def reduce_max(x):
    if world == 1:
        return x
    t = torch.tensor([x], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


best = None
for rep in range(3):
    barrier()
    t0 = time.perf_counter()
    table = ctb.find_features(reader.stack, separation=9, diameter=11, minmass=0, noise_size=1,
                              first_frame=first)
    t1 = time.perf_counter()
    if world > 1:
        out = parallel.refine_leastsq_sharded(table, reader, 11, presharded=True, gather='root')
    else:
        out = ctb.refine_leastsq(table, reader, 11)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    dt, dfind = reduce_max(t2 - t0), reduce_max(t1 - t0)
    if best is None or dt < best[0]:
        n_feat = len(table)
        best = (dt, dfind, n_feat, None if out is None else (len(out), int(np.isnan(out['cost'].values).sum())))
if world > 1:
    t = torch.tensor([best[2]], dtype=torch.float64, device=device)
    dist.all_reduce(t)
    n_total = int(t.item())
else:
    n_total = best[2]
if rank == 0:
    print(json.dumps(dict(config="config5 without the linker: find_features + refine_leastsq_sharded, ONE "
                                 "video of %d frames 1024x1024 over %d GPU(s)" % (n_frames * world, world),
                          n_gpus=world, frames=n_frames * world, features=n_total,
                          merged_rows=best[3][0], failed_features=best[3][1],
                          seconds=best[0], find_seconds=best[1],
                          frames_per_s=n_frames * world / best[0], features_per_s=n_total / best[0])),
          flush=True)
if world > 1:
    dist.destroy_process_group()
