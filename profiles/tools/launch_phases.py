"""Where the host time of launch_cuda goes (per chunk), on the bench workload."""
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
acc = {}
def timed(name, fn):
    def wrapper(*a, **k):
        t = time.perf_counter()
        out = fn(*a, **k)
        acc[name] = acc.get(name, 0.) + 1e3 * (time.perf_counter() - t)
        return out
    return wrapper
refine.DeviceSession.__init__ = timed("session_init", refine.DeviceSession.__init__)
refine.DeviceSession.schedule = timed("schedule", refine.DeviceSession.schedule)
refine.DeviceSession.run = timed("run", refine.DeviceSession.run)
refine.launch_cuda = timed("launch_cuda", refine.launch_cuda)
refine.prepare_common = timed("prepare_common", refine.prepare_common)
refine.FrameSet.upload_async = timed("upload_async", refine.FrameSet.upload_async)
refine.FrameInfo.__init__ = timed("frame_info", refine.FrameInfo.__init__)
for rep in range(5):
    acc.clear()
    torch.cuda.synchronize()
    t = time.perf_counter()
    out = ctb.refine_leastsq(f0, reader, bench.DIAMETER)
    torch.cuda.synchronize()
    print("total %.1f ms |" % (1e3 * (time.perf_counter() - t)), " ".join("%s %.1f" % kv for kv in acc.items()), flush=True)
