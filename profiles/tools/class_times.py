"""Per size class: duration of the refine launches (config 2, inputs resident).
python profiles/tools/class_times.py [frames]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import bench
from clustertracking_b200 import artificial, refine

n_frames = int(sys.argv[1]) if len(sys.argv) > 1 else 300
dev = torch.device("cuda", 0)
pos, frame, signal, start = bench.video_geometry(n_frames, seed=7)
d_stack = bench.render_video_torch(pos, frame, signal, n_frames, dev, seed=100)
reader = artificial.FrameStack(d_stack.cpu().numpy())
f0 = bench.start_dataframe(start, frame)
plan = refine.prepare(f0.copy(), reader, bench.DIAMETER)
session = refine.DeviceSession(plan, dev)
session.frames.register(d_stack, 0)
session.frames.launch_frame_max(0, n_frames)
slices = session.schedule()
for _ in range(2):
    session.run(slices)
torch.cuda.synchronize()
events = []
session.run(slices, events)
torch.cuda.synchronize()
counts = {cap: count for cap, _, count in slices}
for kind, a, b, label in events:
    if kind == "refine":
        print("class %3d %-8s %7d clusters %8.3f ms" % (label[0], label[1], counts.get(label[0], 0), a.elapsed_time(b)))
print("total %.2f ms" % sum(a.elapsed_time(b) for kind, a, b, _ in events if kind == "refine"))
