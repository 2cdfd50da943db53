"""Per-launch times and iteration counts of BASELINE config 4 (3D anisotropic stacks, size 'var')."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from clustertracking_b200 import artificial, refine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
reader, f0 = artificial.confocal_video(n)
plan = refine.prepare(f0.copy(), reader, (9, 13, 13), param_mode=dict(signal='var', size='var'))
res = refine.execute_cuda(plan)
session = res.session
slices = session.schedule()
torch.cuda.synchronize()
events = []
session.run(slices, events)
torch.cuda.synchronize()
counts = {cap: count for cap, _, count in slices}
for kind, a, b, label in events:
    if kind == "refine":
        print("class %3d %-8s %7d clusters %8.3f ms" % (label[0], label[1], counts.get(label[0], 0), a.elapsed_time(b)))
st = res.stats
print("evals %.2f accums %.2f grad %.2f outer %.2f  failed %d" % (st[:, 0].mean(), st[:, 1].mean(), st[:, 7].mean(), st[:, 2].mean(), int((res.status != 0).sum())))
print("status histogram", np.bincount(res.status))
