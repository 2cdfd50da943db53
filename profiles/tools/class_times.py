"""Per size class: duration of the refine launch with the warp-per-cluster and with the
thread-per-cluster kernel (config 2, inputs resident).  python profiles/tools/class_times.py [frames]"""
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
os.environ['CTK_THREAD_MERGE'] = '0'
for mode in ('0', '1'):
    os.environ['CTK_THREAD_KERNEL'] = mode
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
    times = [a.elapsed_time(b) for kind, a, b in events if kind == "refine"]
    # launches come in pairs (class, its overflow relaunch)
    print("thread kernel" if mode == '1' else "warp kernel  ", " ".join(
        "cap%d:%d clusters %.2f+%.2f ms" % (cap, count, times[2 * k], times[2 * k + 1])
        for k, (cap, start_, count) in enumerate(slices)), "| total %.2f ms" % sum(times), flush=True)
