"""Throughput of ctk_find_maxima (local maxima of a batch of frames) on config-2 frames, inputs
resident in HBM, CUDA-event timed.  HBM-bound: reported against MEASURED_PEAKS.json."""
import ctypes, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import bench
from clustertracking_b200 import _lib

n_frames = int(sys.argv[1]) if len(sys.argv) > 1 else 200
dev = torch.device("cuda", 0)
pos, frame, signal, start = bench.video_geometry(n_frames, seed=7)
d_stack = bench.render_video_torch(pos, frame, signal, n_frames, dev, seed=100)
lib = _lib.load()
n_pixels = 1024 * 1024
ptrs = torch.from_numpy(d_stack.data_ptr() + n_pixels * np.arange(n_frames, dtype=np.int64)).to(dev)
ws = torch.empty(int(lib.ctk_find_workspace_bytes(n_frames, n_pixels, 0)), dtype=torch.uint8, device=dev)
cap = 16384
coords = torch.empty((n_frames, cap, 2), dtype=torch.int32, device=dev)
values = torch.empty((n_frames, cap), dtype=torch.int32, device=dev)
count = torch.empty(n_frames, dtype=torch.int32, device=dev)
thr = torch.empty(n_frames, dtype=torch.float64, device=dev)
shape = (ctypes.c_int64 * 3)(1024, 1024, 1)
size = (ctypes.c_int32 * 3)(int(2 * 5 / np.sqrt(2)), int(2 * 5 / np.sqrt(2)), 1)
margin = (ctypes.c_int32 * 3)(2, 2, 0)
stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)

def run():
    rc = lib.ctk_find_maxima(ptrs.data_ptr(), n_frames, shape, 2, 0, size, 90.0, margin, cap,
                             coords.data_ptr(), values.data_ptr(), count.data_ptr(), thr.data_ptr(),
                             ws.data_ptr(), 1, stream)
    assert rc == 0, lib.ctk_find_last_error()

for _ in range(3):
    run()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
steps = 5
for _ in range(steps):
    run()
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / steps
peaks = bench.measured_peaks()
frame_bytes = n_frames * n_pixels
# passes over frame-sized arrays: histogram 1 read; two filter passes 2 x (read + write); count and
# write 2 x (image + dilation reads)
moved = 9 * frame_bytes
print(json.dumps(dict(kernel="ctk_find_maxima (7 launches)", frames=n_frames, ms=ms,
                      frames_per_s=n_frames / (ms * 1e-3), maxima_per_frame=float(count.float().mean()),
                      algorithmic_GBps=frame_bytes / (ms * 1e-3) / 1e9,
                      moved_GBps=moved / (ms * 1e-3) / 1e9, peak_GBps=peaks["hbm_gbs"],
                      frac_of_peak_moved=moved / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"])))
