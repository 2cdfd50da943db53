"""Status census of the bench workload (config 2): which clusters fail on the device and what the
CPU oracle does with the same clusters.  Run on a GPU box: python profiles/tools/failure_census.py [frames]"""
import collections
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402
import clustertracking_b200 as ctb  # noqa: E402
from clustertracking_b200 import artificial, refine as R, _lib  # noqa: E402

n_frames = int(sys.argv[1]) if len(sys.argv) > 1 else 40
dev = torch.device("cuda", 0)
pos, frame, signal, start = bench.video_geometry(n_frames, seed=7)
d_stack = bench.render_video_torch(pos, frame, signal, n_frames, dev, seed=100)
reader = artificial.FrameStack(d_stack.cpu().numpy())
f0 = bench.start_dataframe(start, frame)
for precision in ("float32", "float64"):
    plan = R.prepare(f0.copy(), reader, bench.DIAMETER, precision=precision)
    res = R.execute_cuda(plan)
    sizes = plan.cluster_sizes()
    print(precision, "clusters", plan.n_clusters, "status histogram",
          sorted(collections.Counter(res.status.tolist()).items()))
    bad = np.flatnonzero(res.status != 0)
    print("  failed by size", sorted(collections.Counter(sizes[bad].tolist()).items()),
          "all by size", sorted(collections.Counter(sizes.tolist()).items()))
    print("  stats of failed (evals, accums, outer, M, E, Q, V, grad):")
    for c in bad[:12]:
        print("   cluster", c, "frame", plan.cluster_frame[c], "n", sizes[c], "status", res.status[c],
              res.stats[c].tolist())
    if precision == "float32":
        keep_bad, keep_plan = bad, plan

# the oracle on the failed clusters
from oracle import cluster_oracle  # noqa: E402
plan = keep_plan
f = plan.f
n_show = 0
for c in keep_bad[:10]:
    rows = plan.order[plan.cluster_offset[c]:plan.cluster_offset[c + 1]]
    sub = f.iloc[rows][['y', 'x', 'signal', 'size', 'background', 'frame']].copy()
    fr = int(sub['frame'].iloc[0])
    img = reader[fr]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        want = cluster_oracle.refine_leastsq(sub.copy(), img, bench.DIAMETER)
        got = ctb.refine_leastsq(sub.copy(), img, bench.DIAMETER)
        got64 = ctb.refine_leastsq(sub.copy(), img, bench.DIAMETER, precision='float64')
    print("cluster", c, "oracle cost", want['cost'].values[0], "ours alone", got['cost'].values[0],
          "f64", got64['cost'].values[0])
    print("  start", sub[['y', 'x']].values.round(3).tolist())
    print("  oracle", want[['y', 'x', 'signal', 'background']].values.round(3).tolist())
    print("  ours64", got64[['y', 'x', 'signal', 'background']].values.round(3).tolist())
