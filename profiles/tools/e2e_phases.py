"""Host phase breakdown of the pipelined refine_leastsq on the bench workload (GPU box)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import bench
import clustertracking_b200 as ctb
from clustertracking_b200 import artificial, refine

n_frames = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
dev = torch.device("cuda", 0)
pos, frame, signal, start = bench.video_geometry(n_frames, seed=7)
d_stack = bench.render_video_torch(pos, frame, signal, n_frames, dev, seed=100)
host = torch.empty(d_stack.shape, dtype=torch.uint8, pin_memory=True)
host.copy_(d_stack); torch.cuda.synchronize()
reader = artificial.FrameStack(host.numpy())
f0 = bench.start_dataframe(start, frame)
for rep in range(6):
    torch.cuda.synchronize()
    t = time.perf_counter()
    out = ctb.refine_leastsq(f0, reader, bench.DIAMETER)
    torch.cuda.synchronize()
    dt = 1e3 * (time.perf_counter() - t)
    print("total %.1f ms | " % dt + " ".join("%s %.1f" % kv for kv in refine.LAST_CALL["phases_ms"].items()),
          "| chunks", refine.LAST_CALL["chunks"], flush=True)
