"""One ctk_refine_batch launch over all clusters of a golden fixture with a given capacity
(debugging aid): python profiles/tools/single_launch.py <fixture> <capacity> <rigorous 0|1> [precision]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import golden_io
import clustertracking_b200 as ctb
from clustertracking_b200 import refine, _lib
name, cap, rig = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
prec = sys.argv[4] if len(sys.argv) > 4 else 'float32'
d = golden_io.load(name)
f0, reader, diameter, kwargs = golden_io.refine_inputs(d, ctb.constraints)
plan = refine.prepare(f0.copy(), reader, diameter, precision=prec, **kwargs)
frames = refine.FrameSet(plan.frame_info).upload_async()
session = refine.DeviceSession(plan, frames=frames)
session.rigorous = refine.rigorous_problem(plan.problem)
session.frames.wait_for_frames(0)
session.launch_refine(cap, None, plan.n_clusters, problem=session.rigorous if rig else None)
torch.cuda.synchronize()
print("sizes ", plan.cluster_sizes().tolist())
print("status", session.d_status.cpu().numpy().tolist())
print("evals ", session.d_stats.cpu().numpy()[:, 0].tolist())
