"""Cluster labelling of the config-2 video (1000 frames x ~2100 features): the device kernel
(ctk_label_frames, one warp per frame) against the host threads (ctk_cluster_frames).
    python profiles/tools/label_bench.py [frames]
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

import bench  # noqa: E402
from clustertracking_b200 import _lib  # noqa: E402
from clustertracking_b200.utils import validate_tuple  # noqa: E402


def main():
    n_frames = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    pos, frame, signal, start = bench.video_geometry(n_frames, seed=7)
    starts = np.searchsorted(frame, np.arange(n_frames)).astype(np.int64)
    stops = np.concatenate((starts[1:], [len(frame)])).astype(np.int64)
    separation = np.asarray(validate_tuple(bench.DIAMETER, 2), dtype=np.float64)   # refine.py:259: separation = diameter
    start = np.ascontiguousarray(start)
    dev = torch.device("cuda", 0)
    d_pos = [torch.from_numpy(np.ascontiguousarray(start[:, k])).to(dev) for k in range(2)]
    d_starts, d_stops = torch.from_numpy(starts).to(dev), torch.from_numpy(stops).to(dev)
    max_points = int((stops - starts).max())
    nbytes = _lib.label_frames_scratch_bytes(max_points, 2, n_frames)
    scratch = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    labels = torch.empty(len(start), dtype=torch.int32, device=dev)
    flags = torch.empty(n_frames, dtype=torch.int32, device=dev)

    def launch():
        _lib.label_frames_device([t.data_ptr() for t in d_pos], 2, d_starts.data_ptr(), d_stops.data_ptr(),
                                 n_frames, max_points, separation, labels.data_ptr(), flags.data_ptr(),
                                 scratch.data_ptr(), nbytes, None)

    for _ in range(3):
        launch()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    e0.record()
    for _ in range(reps):
        launch()
    e1.record()
    torch.cuda.synchronize()
    device_ms = e0.elapsed_time(e1) / reps
    timing = scratch[64:128].cpu().numpy().view(np.uint64)        # CTK_LABEL_TIMING builds only
    got = labels.cpu().numpy()
    got_flags = flags.cpu().numpy()
    host = {}
    for threads in (1, 4, 16):
        t0 = time.perf_counter()
        want, _, _, _ = _lib.cluster_frames(start, starts, stops, separation, threads)
        host[threads] = 1e3 * (time.perf_counter() - t0)
    ok = got_flags == 0
    same = all(np.array_equal(got[a:b], want[a:b]) for a, b, k in zip(starts, stops, ok) if k)
    print(json.dumps(dict(frames=n_frames, features=int(len(start)), max_points=max_points,
                          scratch_mb=nbytes / 1e6, device_ms=device_ms,
                          host_ms_by_threads=host, flagged_frames=int((~ok).sum()),
                          labels_identical=bool(same),
                          phase_mcycles_per_frame=dict(zip(
                              ("init", "build", "query", "set_order", "union_total", "build_root",
                               "query_leaf_blocks", "leaf_blocks_count"),
                              (timing / 1e6 / n_frames).round(3).tolist())) if timing.any() else None,
                          clusters=int(sum(len(np.unique(want[a:b])) for a, b in zip(starts, stops))))))


if __name__ == "__main__":
    main()
