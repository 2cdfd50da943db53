"""cProfile of the main thread over repeated end-to-end calls (config 2, pinned host frames).
python profiles/tools/e2e_profile.py [frames] [calls]"""
import cProfile
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
import clustertracking_b200 as ctb  # noqa: E402
from clustertracking_b200 import artificial, refine  # noqa: E402

n_frames = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
calls = int(sys.argv[2]) if len(sys.argv) > 2 else 25
dev = torch.device("cuda", 0)
pos, frame, signal, start = bench.video_geometry(n_frames, seed=7)
d_stack = bench.render_video_torch(pos, frame, signal, n_frames, dev, seed=100)
host = torch.empty(d_stack.shape, dtype=torch.uint8, pin_memory=True)
host.copy_(d_stack)
torch.cuda.synchronize()
del d_stack
reader = artificial.FrameStack(host.numpy())
f0 = bench.start_dataframe(start, frame)
for _ in range(3):
    ctb.refine_leastsq(f0, reader, bench.DIAMETER)
prof = cProfile.Profile()
times = []
for _ in range(calls):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    prof.enable()
    out = ctb.refine_leastsq(f0, reader, bench.DIAMETER)
    prof.disable()
    torch.cuda.synchronize()
    times.append(1e3 * (time.perf_counter() - t0))
    out = None
print("ms per call:", " ".join("%.0f" % t for t in times))
print("setup parts of the last call:", refine.LAST_CALL.get("setup_ms"))
pstats.Stats(prof).sort_stats("tottime").print_stats(18)
