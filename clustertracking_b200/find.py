"""Cluster grouping on the host (reference: clustertracking/find.py:12-163).

``find_clusters`` decides which features are fitted together, so membership, row order and even the
label values must equal the reference's.  The close pairs come from the same
``cKDTree.query_pairs`` call as upstream and are visited in the same order (the iteration order of
the python ``set`` it returns); the union step itself runs in C (``ctk_label_clusters``) instead of
the reference's dict-of-sets loop (find.py:12-60), and frames are cut from one stable sort instead
of a pandas ``groupby`` + ``concat``.
"""
from itertools import chain

import numpy as np
from scipy.spatial import cKDTree

from . import _lib
from .utils import guess_pos_columns, validate_tuple


def _label_frame(pos, separation):
    """ids, sizes (int64 arrays) for the points of one frame (find.py:72-93)."""
    n = len(pos)
    if n == 0:
        return np.zeros(0, np.int64), np.zeros(0, np.int64)
    pairs = cKDTree(pos / separation).query_pairs(1)                 # a python set, as upstream
    flat = np.fromiter(chain.from_iterable(pairs), dtype=np.int64, count=2 * len(pairs))
    return _lib.label_clusters(flat.reshape(-1, 2), n)


def find_iter(f, separation, pos_columns=None, t_column='frame'):
    """Generator of ``(frame_no, DataFrame)`` with ``cluster`` and ``cluster_size`` columns added
    (find.py:96-129)."""
    if pos_columns is None:
        pos_columns = guess_pos_columns(f)
    next_id = 0
    for frame_no, part in f.groupby(t_column):
        ids, sizes = _label_frame(part[pos_columns].values.astype(np.float64), separation)
        part = part.copy()
        part['cluster'] = ids + next_id
        part['cluster_size'] = sizes
        next_id = int(part['cluster'].max()) + 1
        yield frame_no, part


def find_clusters(f, separation, pos_columns=None, t_column='frame'):
    """Group features closer than ``separation`` (number or per-axis tuple) into clusters, frame by
    frame (find.py:132-163).

    Returns a frame-sorted COPY of ``f`` (original index labels, original order inside a frame) with
    int64 columns ``cluster`` (unique over all frames) and ``cluster_size``.
    """
    if pos_columns is None:
        pos_columns = guess_pos_columns(f)
    separation = np.asarray(validate_tuple(separation, len(pos_columns)), dtype=np.float64)
    if t_column in f:
        frames = f[t_column].values
    else:
        frames = np.zeros(len(f), dtype=np.int64)            # find.py:151-161 adds, then deletes it
    order = np.argsort(frames, kind='stable')
    pos = f[pos_columns].values.astype(np.float64)[order]
    sorted_frames = frames[order]
    cuts = np.flatnonzero(sorted_frames[1:] != sorted_frames[:-1]) + 1
    starts = np.concatenate(([0], cuts))
    stops = np.concatenate((cuts, [len(order)]))
    cluster = np.empty(len(order), dtype=np.int64)
    size = np.empty(len(order), dtype=np.int64)
    next_id = 0
    for a, b in zip(starts, stops):
        if b == a:
            continue
        ids, sizes = _label_frame(pos[a:b], separation)
        cluster[a:b] = ids + next_id                          # find.py:127-128
        size[a:b] = sizes
        next_id = int(cluster[a:b].max()) + 1
    out = f.iloc[order].copy()
    if t_column not in f:
        out[t_column] = 0                                     # the copies keep the temporary column
    out['cluster'] = cluster
    out['cluster_size'] = size
    return out
