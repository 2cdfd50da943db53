"""Cluster labels from the GPU (ctk_label_frames, csrc/ctk_label.cuh) against the host restatement
(ctk_cluster_frames, itself verified against scipy): identical label VALUES (find.py:41-48, 87-91)."""
import numpy as np
import pytest

from clustertracking_b200 import _lib
from label_cases import label_cases

pytestmark = pytest.mark.gpu


def device_labels(pos, starts, stops, separation):
    import torch
    dev = torch.device('cuda', 0)
    n, m = pos.shape
    d_pos = [torch.from_numpy(np.ascontiguousarray(pos[:, k])).to(dev) for k in range(m)]
    d_starts = torch.from_numpy(np.asarray(starts, np.int64)).to(dev)
    d_stops = torch.from_numpy(np.asarray(stops, np.int64)).to(dev)
    max_points = int(np.max(np.asarray(stops) - np.asarray(starts))) if len(starts) else 0
    nbytes = _lib.label_frames_scratch_bytes(max_points, m, len(starts))
    scratch = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    labels = torch.full((max(1, n),), -7, dtype=torch.int32, device=dev)
    flags = torch.full((max(1, len(starts)),), -7, dtype=torch.int32, device=dev)
    _lib.label_frames_device([t.data_ptr() for t in d_pos], m, d_starts.data_ptr(), d_stops.data_ptr(),
                             len(starts), max_points, separation, labels.data_ptr(),
                             flags.data_ptr(), scratch.data_ptr(), nbytes, None)
    torch.cuda.synchronize()
    return labels.cpu().numpy()[:n], flags.cpu().numpy()[:len(starts)]


@pytest.mark.parametrize("name", sorted(label_cases()))
def test_device_labels_equal_host_labels(name):
    pos, starts, stops, separation = label_cases()[name]
    want_l, want_s, _, _ = _lib.cluster_frames(pos, starts, stops, separation, 4)
    got_l, flags = device_labels(pos, starts, stops, separation)
    for f, (a, b) in enumerate(zip(starts, stops)):
        if flags[f] == 1:            # capacity: the host path labels this frame (exercised below)
            continue
        assert flags[f] == 0
        assert np.array_equal(got_l[a:b], want_l[a:b]), (name, f)
    assert np.count_nonzero(flags) <= label_cases.expected_flagged.get(name, 0)


def test_refine_leastsq_labels_do_not_depend_on_where_they_are_computed(monkeypatch):
    """The public call with the labelling on the device and on the host threads: same table."""
    import clustertracking_b200 as ctb
    from clustertracking_b200 import artificial
    frames, f0 = [], []
    for k in range(12):
        frame, f, _ = artificial.clustered_frame((256, 256), 44, 2.75, 8, seed=50 + k)
        f['frame'] = k
        frames.append(frame)
        f0.append(f)
    import pandas as pd
    f0 = pd.concat(f0, ignore_index=True)
    reader = artificial.FrameStack(np.stack(frames))
    monkeypatch.setenv('CTK_LABEL_DEVICE', '2')            # forced: the table is below the automatic threshold
    on_device = ctb.refine_leastsq(f0.copy(), reader, 11)
    monkeypatch.setenv('CTK_LABEL_DEVICE', '0')
    on_host = ctb.refine_leastsq(f0.copy(), reader, 11)
    from clustertracking_b200 import refine
    pd.testing.assert_frame_equal(on_device, on_host)
    assert refine.LAST_CALL['labelling']['where'] == 'host threads'
    monkeypatch.setenv('CTK_LABEL_DEVICE', '2')
    ctb.refine_leastsq(f0.copy(), reader, 11)
    assert refine.LAST_CALL['labelling']['where'] == 'device'
    assert refine.LAST_CALL['labelling']['frames_relabelled_on_host'] == 0


def test_non_finite_positions_raise_like_scipy(monkeypatch):
    """cKDTree refuses non-finite data (ValueError); so does the labelling on either side."""
    import clustertracking_b200 as ctb
    from clustertracking_b200 import artificial
    import pandas as pd
    frames, f0 = [], []
    for k in range(6):
        frame, f, _ = artificial.clustered_frame((128, 128), 44, 2.75, 8, seed=70 + k)
        f['frame'] = k
        frames.append(frame)
        f0.append(f)
    f0 = pd.concat(f0, ignore_index=True)
    f0.loc[len(f0) // 2, 'x'] = np.nan
    reader = artificial.FrameStack(np.stack(frames))
    for mode in ('2', '0'):
        monkeypatch.setenv('CTK_LABEL_DEVICE', mode)
        with pytest.raises(ValueError):
            ctb.refine_leastsq(f0.copy(), reader, 11)
    # ... and the next call is not disturbed by the failed one
    monkeypatch.setenv('CTK_LABEL_DEVICE', '2')
    good = f0.dropna()
    out = ctb.refine_leastsq(good.copy(), reader, 11)
    monkeypatch.setenv('CTK_LABEL_DEVICE', '0')
    pd.testing.assert_frame_equal(out, ctb.refine_leastsq(good.copy(), reader, 11))


def test_config2_video_labels_equal_host_labels():
    """64 frames of the bench's config-2 geometry (~2100 features per frame in clusters of 2-6, the
    case where the label VALUES depend on the pair order): every label identical, nothing flagged."""
    import sys
    import os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    n_frames = 64
    _, frame, _, start = bench.video_geometry(n_frames, seed=11)
    starts = np.searchsorted(frame, np.arange(n_frames)).astype(np.int64)
    stops = np.concatenate((starts[1:], [len(frame)])).astype(np.int64)
    separation = np.array([float(bench.DIAMETER)] * 2)
    start = np.ascontiguousarray(start)
    want, _, _, _ = _lib.cluster_frames(start, starts, stops, separation, 4)
    got, flags = device_labels(start, starts, stops, separation)
    assert not flags.any()
    assert np.array_equal(got, want)
    sizes = np.bincount(np.bincount(want[starts[0]:stops[0]]))      # components of 3+ exist
    assert len(sizes) > 3
