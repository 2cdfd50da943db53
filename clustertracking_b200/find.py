"""Cluster grouping on the host (reference: clustertracking/find.py:12-163).

``find_clusters`` decides which features are fitted together, so membership, row order and even the
label values must equal the reference's.  The close pairs come from the same scipy
``cKDTree.query_pairs`` as upstream and are visited in the same order -- the iteration order of the
python ``set`` upstream receives (find.py:87-91).  That order is replayed in C from the array form
of the result (``ctk_pairs_set_order``; no python objects), the union step runs in C
(``ctk_label_clusters``) instead of the reference's dict-of-sets loop (find.py:12-60), frames are
cut from one stable sort instead of a pandas ``groupby`` + ``concat``, and frames are processed by
a small thread pool (scipy's kd-tree and the ctypes calls release the GIL).
"""
import os
from concurrent.futures import ThreadPoolExecutor
from itertools import chain

import numpy as np
from scipy.spatial import cKDTree

from . import _lib
from .utils import guess_pos_columns, validate_tuple

_replay_checked = None      # None = not checked yet, True = replay matches this interpreter's sets


def _pairs_via_set(tree):
    pairs = tree.query_pairs(1)                                     # a python set, as upstream
    flat = np.fromiter(chain.from_iterable(pairs), dtype=np.int64, count=2 * len(pairs))
    return flat.reshape(-1, 2)


def _pairs_via_replay(tree):
    return _lib.pairs_set_order(tree.query_pairs(1, output_type='ndarray'))


def _replay_is_exact():
    """One-time self check: the C replay of CPython's set order must reproduce a real set."""
    global _replay_checked
    if _replay_checked is None:
        rng = np.random.RandomState(12345)
        ok = True
        for n in (40, 700, 3000):
            tree = cKDTree(rng.uniform(0, 4 * np.sqrt(n), (n, 2)))
            ok = ok and np.array_equal(_pairs_via_set(tree), _pairs_via_replay(tree))
        _replay_checked = bool(ok)
    return _replay_checked


def _label_frame(pos, separation):
    """ids, sizes (int64 arrays) for the points of one frame (find.py:72-93)."""
    n = len(pos)
    if n == 0:
        return np.zeros(0, np.int64), np.zeros(0, np.int64)
    tree = cKDTree(pos / separation)
    pairs = _pairs_via_replay(tree) if _replay_is_exact() else _pairs_via_set(tree)
    return _lib.label_clusters(pairs, n)


def find_iter(f, separation, pos_columns=None, t_column='frame'):
    """Generator of ``(frame_no, DataFrame)`` with ``cluster`` and ``cluster_size`` columns added
    (find.py:96-129)."""
    if pos_columns is None:
        pos_columns = guess_pos_columns(f)
    separation = np.asarray(validate_tuple(separation, len(pos_columns)), dtype=np.float64)
    next_id = 0
    for frame_no, part in f.groupby(t_column):
        ids, sizes = _label_frame(part[pos_columns].values.astype(np.float64), separation)
        part = part.copy()
        part['cluster'] = ids + next_id
        part['cluster_size'] = sizes
        next_id = int(part['cluster'].max()) + 1
        yield frame_no, part


def label_frames(pos, starts, stops, separation):
    """Per-frame labels for frame-sorted positions: -> (cluster ids with the running offset of
    find.py:127-128 applied, cluster sizes, permutation that sorts the rows by (frame, cluster)
    keeping the row order inside a cluster -- the group order of refine.py:336)."""
    n_frames = len(starts)
    cluster = np.empty(len(pos), dtype=np.int64)
    size = np.empty(len(pos), dtype=np.int64)
    by_cluster = np.empty(len(pos), dtype=np.int64)

    def work(k):
        a, b = starts[k], stops[k]
        ids, sizes = _label_frame(pos[a:b], separation)
        cluster[a:b] = ids
        size[a:b] = sizes
        by_cluster[a:b] = a + np.argsort(ids, kind='stable')
        return int(ids.max()) + 1 if b > a else 0

    _replay_is_exact()                      # decide the path once, before threads start
    workers = min(32, os.cpu_count() or 1, max(1, n_frames // 4))
    if workers > 1:
        with ThreadPoolExecutor(workers) as pool:
            spans = list(pool.map(work, range(n_frames)))
    else:
        spans = [work(k) for k in range(n_frames)]
    offsets = np.concatenate(([0], np.cumsum(spans)[:-1]))
    cluster += np.repeat(offsets, np.asarray(stops) - np.asarray(starts))
    return cluster, size, by_cluster


def find_clusters(f, separation, pos_columns=None, t_column='frame'):
    """Group features closer than ``separation`` (number or per-axis tuple) into clusters, frame by
    frame (find.py:132-163).

    Returns a frame-sorted COPY of ``f`` (original index labels, original order inside a frame) with
    int64 columns ``cluster`` (unique over all frames) and ``cluster_size``.
    """
    return cluster_table(f, separation, pos_columns, t_column)[0]


def cluster_table(f, separation, pos_columns=None, t_column='frame'):
    """``find_clusters`` plus the row permutation that lists the result by (frame, cluster)."""
    if pos_columns is None:
        pos_columns = guess_pos_columns(f)
    separation = np.asarray(validate_tuple(separation, len(pos_columns)), dtype=np.float64)
    if t_column in f:
        frames = f[t_column].values
    else:
        frames = np.zeros(len(f), dtype=np.int64)            # find.py:151-161 adds, then deletes it
    if len(frames) > 1 and np.all(frames[1:] >= frames[:-1]):
        order = None                                          # already frame-sorted
        pos = np.ascontiguousarray(f[pos_columns].values, dtype=np.float64)
        sorted_frames = frames
    else:
        order = np.argsort(frames, kind='stable')
        pos = f[pos_columns].values.astype(np.float64)[order]
        sorted_frames = frames[order]
    cuts = np.flatnonzero(sorted_frames[1:] != sorted_frames[:-1]) + 1
    starts = np.concatenate(([0], cuts)).astype(np.int64)
    stops = np.concatenate((cuts, [len(pos)])).astype(np.int64)
    cluster, size, by_cluster = label_frames(pos, starts, stops, separation)
    out = f.copy() if order is None else f.iloc[order].copy()
    if t_column not in f:
        out[t_column] = 0                                     # the copies keep the temporary column
    out['cluster'] = cluster
    out['cluster_size'] = size
    return out, by_cluster
