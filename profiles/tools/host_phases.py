"""Time the host-side phases of refine_leastsq on the config-2 geometry (no GPU work)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench
from clustertracking_b200 import find, refine

n_frames = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
pos, frame, signal, start = bench.video_geometry(n_frames, 7)
f0 = bench.start_dataframe(start, frame)
sep = np.array([11., 11.])
for rep in range(3):
    t = [time.perf_counter()]
    frames = f0['frame'].values
    mono = np.all(frames[1:] >= frames[:-1]); t.append(time.perf_counter())
    p = np.ascontiguousarray(f0[['y', 'x']].values, dtype=np.float64); t.append(time.perf_counter())
    cuts = np.flatnonzero(frames[1:] != frames[:-1]) + 1
    starts = np.concatenate(([0], cuts)).astype(np.int64); stops = np.concatenate((cuts, [len(p)])).astype(np.int64)
    t.append(time.perf_counter())
    job = find._LabelJob(p, starts, stops, sep); t.append(time.perf_counter())
    out = f0.copy(); t.append(time.perf_counter())
    cluster, size, by = job.result(); t.append(time.perf_counter())
    out['cluster'] = cluster; out['cluster_size'] = size; t.append(time.perf_counter())
    names = ['monotonic check', 'pos extract', 'cuts', 'job submit (shm copy)', 'f.copy', 'job.result (wait+copy)', 'column insert']
    print(' | '.join('%s %.1f' % (n, 1e3 * (b - a)) for n, a, b in zip(names, t[:-1], t[1:])), '| total %.1f ms' % (1e3 * (t[-1] - t[0])))
print('workers', find._pool_workers(), 'cpus', os.cpu_count())
