"""ctypes binding of ``libctk.so`` (include/ctk.h).

The library is built in-tree by ``__graft_entry__.build()`` / ``python -m clustertracking_b200.build``.
There is no fallback: if the shared object is missing or a symbol is absent, importing this module's
users fails with an explicit error.
"""
import ctypes
import os

import numpy as np

CTK_MAX_PARAMS = 12
CTK_MAX_CLUSTER_FEATURES = 32
CTK_MAX_BIG_FEATURES = 256
CTK_MAX_RADIUS = 30
CTK_MAX_TAPS = 33

PIXEL_CODES = {np.dtype(np.uint8): 0, np.dtype(np.uint16): 1, np.dtype(np.float32): 2,
               np.dtype(np.float64): 3, np.dtype(np.int16): 4, np.dtype(np.int32): 5}
COMPUTE_F32, COMPUTE_F64 = 0, 1
MODE_CONST, MODE_VAR, MODE_GLOBAL, MODE_CLUSTER = 0, 1, 2, 3
FAMILY_GAUSS, FAMILY_RING, FAMILY_DISC = 0, 1, 2
CONSTRAINT_DIMER, CONSTRAINT_TRIMER, CONSTRAINT_TETRAMER = 1, 2, 4
LAUNCH_APPEND_OVERFLOW = 1        # ctk_refine_batch_ex flags

STATUS_NAMES = {0: 'ok', 1: 'non-finite initial parameters', 2: 'cluster outside of the image',
                3: 'solver did not converge', 4: 'rms deviation above max_rms_dev',
                5: 'lower bound above upper bound', 6: 'cluster exceeds the kernel capacity',
                7: 'non-finite value during the fit'}
STATUS_TOO_LARGE = 6


class Problem(ctypes.Structure):
    """``ctk_problem_t`` of include/ctk.h."""
    _fields_ = [
        ("ndim", ctypes.c_int32), ("isotropic", ctypes.c_int32), ("family", ctypes.c_int32),
        ("n_params", ctypes.c_int32), ("modes", ctypes.c_int32 * CTK_MAX_PARAMS),
        ("radius", ctypes.c_int32 * 3), ("pixel_dtype", ctypes.c_int32),
        ("compute_dtype", ctypes.c_int32), ("max_iter", ctypes.c_int32),
        ("lm_max_iter", ctypes.c_int32), ("max_shift", ctypes.c_double),
        ("max_rms_dev", ctypes.c_double), ("residual_factor", ctypes.c_double),
        ("xtol", ctypes.c_double), ("chord_tol", ctypes.c_double),
        ("constraint_mask", ctypes.c_int32),
        ("capacity_mode", ctypes.c_int32), ("dimer_dist", ctypes.c_double * 3),
        ("trimer_dist", ctypes.c_double * 3),
        ("tetramer_dist", ctypes.c_double * 3),
        ("bounds_abs", (ctypes.c_double * CTK_MAX_PARAMS) * 2),
        ("bounds_diff", (ctypes.c_double * CTK_MAX_PARAMS) * 2),
        ("bounds_rel", (ctypes.c_double * CTK_MAX_PARAMS) * 2),
        ("lowpass", ctypes.c_int32), ("lowpass_half", ctypes.c_int32 * 3),
        ("lowpass_threshold", ctypes.c_double),
        ("lowpass_sigma", ctypes.c_double * 3),
        ("probe_step", ctypes.c_double), ("probe_sweeps", ctypes.c_int32),
        ("reserved_", ctypes.c_int32),
    ]


LIB_PATH = os.environ.get("CTK_LIB_PATH") or os.path.join(os.path.dirname(os.path.abspath(__file__)),
                                                          "libctk.so")
_lib = None
_WORKSPACE = {}       # cached host scratch arrays (internal to one call at a time)


def workspace(key, n, dtype):
    """A reusable scratch array of at least ``n`` items: fresh multi-megabyte numpy arrays are new
    mappings whose pages fault in one by one -- per call and per rank, at the same moment on every
    rank of a box.  Only for arrays that never leave the call that uses them."""
    arr = _WORKSPACE.get(key)
    if arr is None or arr.dtype != np.dtype(dtype) or len(arr) < n:
        arr = np.empty(max(int(n), 1), dtype=dtype)
        _WORKSPACE[key] = arr
    return arr[:n]

_vp, _i32, _i64, _sz = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_size_t
_PROTOTYPES = {
    "ctk_version": (ctypes.c_int, []),
    "ctk_problem_bytes": (ctypes.c_size_t, []),
    "ctk_last_error": (ctypes.c_char_p, []),
    "ctk_frame_max": (ctypes.c_int, [_vp, _i32, _i64, _i32, _vp, _vp]),
    "ctk_refine_workspace_bytes": (_sz, []),
    "ctk_refine_shared_bytes": (_sz, [ctypes.POINTER(Problem), _i32]),
    "ctk_refine_workspace_bytes_for": (_sz, [ctypes.POINTER(Problem), _i32]),
    "ctk_refine_batch": (ctypes.c_int, [ctypes.POINTER(Problem), _vp, ctypes.POINTER(_i64), _vp,
                                        _i32, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                        _vp, _vp, _vp]),
    "ctk_refine_batch_chained": (ctypes.c_int, [ctypes.POINTER(Problem), _vp, ctypes.POINTER(_i64),
                                                _vp, _i32, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp,
                                                _vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp]),
    "ctk_refine_batch_ex": (ctypes.c_int, [ctypes.POINTER(Problem), _vp, ctypes.POINTER(_i64),
                                           _vp, _i32, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp,
                                           _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _vp]),
    "ctk_global_pass": (ctypes.c_int, [ctypes.POINTER(Problem), _vp, ctypes.POINTER(_i64),
                                       ctypes.c_double, _i32, _i32, _vp, _vp, _vp, _vp, _i32,
                                       ctypes.c_double, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "ctk_label_clusters": (ctypes.c_int, [_vp, _i64, _i64, _vp, _vp]),
    "ctk_pairs_set_order": (ctypes.c_int, [_vp, _i64, _vp]),
    "ctk_group_chunk": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _i32, _vp, _vp, _vp,
                                       _vp, _vp, _vp]),
    "ctk_schedule": (ctypes.c_int, [_vp, _i64, _vp, _i32, _vp, _vp, _vp, _vp, _vp]),
    "ctk_gather_rows": (ctypes.c_int, [_vp, _vp, _vp, _i64, _i32, _vp, _i32]),
    "ctk_scatter_rows": (ctypes.c_int, [_vp, _vp, _vp, _i64, _i64, _i32, _vp, _vp, _vp, _i64, _vp,
                                        _vp, _i32, _vp]),
    "ctk_cluster_pack_frames": (ctypes.c_int, [_vp, _i64, _i32, _vp, _vp, _i64, _vp, _i32, _vp, _vp,
                                               _vp, _vp, _vp, _vp, _i32, _i64, _vp, _vp, _vp]),
    "ctk_cluster_pack_columns": (ctypes.c_int, [_vp, _vp, _i64, _i32, _vp, _vp, _i64, _vp, _i32, _vp,
                                                _vp, _vp, _vp, _vp, _vp, _i32, _i64, _vp, _vp, _vp]),
    "ctk_cluster_pack_labelled": (ctypes.c_int, [_vp, _vp, _i64, _i32, _vp, _vp, _i64, _vp, _i32, _vp,
                                                 _vp, _vp, _vp, _vp, _vp, _i32, _i64, _vp, _vp, _vp,
                                                 _vp, _vp]),
    "ctk_frame_runs": (ctypes.c_int, [_vp, _i64, _vp, _i64, _vp, _vp]),
    "ctk_wait_flags": (ctypes.c_int, [_vp, _i64, _i64]),
    "ctk_label_frames_scratch": (ctypes.c_int, [_i64, _i32, _i64, _vp]),
    "ctk_label_frames": (ctypes.c_int, [_vp, _i32, _vp, _vp, _i64, _i64, _vp, _vp, _vp, _vp, _i64, _vp]),
    "ctk_concat_groups": (ctypes.c_int, [_vp, _vp, _vp, _vp, _i64, _i32, _vp, _vp, _vp]),
    "ctk_apply_label_offsets": (ctypes.c_int, [_vp, _vp, _vp, _vp, _i64, _i32, _vp]),
    "ctk_find_workspace_bytes": (_sz, [_i32, _i64, _i32]),
    "ctk_find_maxima": (ctypes.c_int, [_vp, _i32, ctypes.POINTER(_i64), _i32, _i32, _vp, ctypes.c_double,
                                       _vp, _i32, _vp, _vp, _vp, _vp, _vp, _i32, _vp]),
    "ctk_find_last_error": (ctypes.c_char_p, []),
    "ctk_drop_close_frames": (ctypes.c_int, [_vp, _vp, _vp, _i64, _i32, _i32, _vp, _i32, _vp, _vp]),
    "ctk_query_pairs_within": (ctypes.c_int, [_vp, _i64, _i32, ctypes.c_double, _vp, _i64, _vp]),
    "ctk_query_pairs": (ctypes.c_int, [_vp, _i64, _i32, _vp, _i64, _vp]),
    "ctk_cluster_frames": (ctypes.c_int, [_vp, _i64, _i32, _vp, _vp, _i64, _vp, _i32, _vp, _vp, _vp,
                                          _vp]),
}


def load():
    """Load (once) and return the ctypes handle; raises if the CUDA library has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "clustertracking_b200: %s not found. Build the CUDA library first "
            "(python -m clustertracking_b200.build). There is no CPU fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in _PROTOTYPES.items():
        fn = getattr(lib, name)                 # AttributeError if the symbol is missing
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.ctk_problem_bytes() != ctypes.sizeof(Problem):
        raise ImportError("clustertracking_b200: %s was built from a different include/ctk.h "
                          "(ctk_problem_t is %d bytes there, %d here); rebuild it"
                          % (LIB_PATH, lib.ctk_problem_bytes(), ctypes.sizeof(Problem)))
    _lib = lib
    return lib


def check(code, what):
    if code == -5:                          # CTK_E_NONFINITE: what scipy's cKDTree raises
        raise ValueError("data must be finite, check for nan or inf values")
    if code != 0:
        msg = load().ctk_last_error()
        raise RuntimeError("%s failed (%d): %s" % (what, code, msg.decode() if msg else "?"))


def label_clusters(pairs, n):
    """Host helper ``ctk_label_clusters``: labels and sizes from close pairs in visiting order."""
    pairs = np.ascontiguousarray(pairs, dtype=np.int64)
    labels = np.empty(n, dtype=np.int64)
    sizes = np.empty(n, dtype=np.int64)
    check(load().ctk_label_clusters(pairs.ctypes.data, len(pairs), n, labels.ctypes.data,
                                    sizes.ctypes.data), "ctk_label_clusters")
    return labels, sizes


def pairs_set_order(pairs):
    """Host helper ``ctk_pairs_set_order``: ``pairs`` (in ``query_pairs(output_type='ndarray')``
    order) reordered into the iteration order of the python set scipy would have built."""
    pairs = np.ascontiguousarray(pairs, dtype=np.int64)
    order = np.empty(len(pairs), dtype=np.int64)
    check(load().ctk_pairs_set_order(pairs.ctypes.data, len(pairs), order.ctypes.data),
          "ctk_pairs_set_order")
    return pairs[order]


def query_pairs(data, r=1.0):
    """Host helper ``ctk_query_pairs_within``: pairs of ``data`` [n, ndim] closer than ``r``; for
    r = 1 in the order of ``scipy.spatial.cKDTree(data).query_pairs(1, output_type='ndarray')``."""
    data = np.ascontiguousarray(data, dtype=np.float64)
    n, ndim = data.shape
    count = ctypes.c_int64(0)
    lib = load()
    capacity = max(16, 4 * n)
    while True:
        out = np.empty((capacity, 2), dtype=np.int64)
        code = lib.ctk_query_pairs_within(data.ctypes.data, n, ndim, float(r), out.ctypes.data,
                                          capacity, ctypes.byref(count))
        if code == -4:                      # CTK_E_CAPACITY: count holds the size needed
            capacity = int(count.value)
            continue
        check(code, "ctk_query_pairs")
        return out[:count.value]


def cluster_frames(pos, starts, stops, separation, n_threads):
    """Host helper ``ctk_cluster_frames`` -> (cluster, size, by_cluster, spans), int64 arrays."""
    pos = np.ascontiguousarray(pos, dtype=np.float64)
    n, ndim = pos.shape
    starts = np.ascontiguousarray(starts, dtype=np.int64)
    stops = np.ascontiguousarray(stops, dtype=np.int64)
    separation = np.ascontiguousarray(separation, dtype=np.float64)
    cluster = np.empty(n, dtype=np.int64)
    size = np.empty(n, dtype=np.int64)
    by_cluster = np.empty(n, dtype=np.int64)
    spans = np.zeros(len(starts), dtype=np.int64)
    check(load().ctk_cluster_frames(pos.ctypes.data, n, ndim, starts.ctypes.data, stops.ctypes.data,
                                    len(starts), separation.ctypes.data, int(n_threads),
                                    cluster.ctypes.data, size.ctypes.data, by_cluster.ctypes.data,
                                    spans.ctypes.data), "ctk_cluster_frames")
    return cluster, size, by_cluster, spans


def group_chunk(local, by_cluster, starts, stops, spans, next_id, row_base, frame_base, cluster_out):
    """``ctk_group_chunk`` -> (order, group_offset int32 [g + 1], group_frame int32 [g], next_id);
    ``cluster_out`` (int64 [m], a view into the table's cluster column) is filled in place."""
    m = len(local)
    order = np.empty(m, dtype=np.int64)
    goff = np.empty(m + 1, dtype=np.int32)
    gframe = np.empty(max(m, 1), dtype=np.int32)
    n_groups, nxt = ctypes.c_int64(0), ctypes.c_int64(0)
    assert cluster_out.flags.c_contiguous and cluster_out.dtype == np.int64
    check(load().ctk_group_chunk(local.ctypes.data, by_cluster.ctypes.data, starts.ctypes.data,
                                 stops.ctypes.data, spans.ctypes.data, len(starts), int(next_id),
                                 int(row_base), int(frame_base), cluster_out.ctypes.data,
                                 order.ctypes.data, goff.ctypes.data, gframe.ctypes.data,
                                 ctypes.byref(n_groups), ctypes.byref(nxt)), "ctk_group_chunk")
    g = n_groups.value
    return order, goff[:g + 1], gframe[:g], nxt.value


def gather_rows(sources, rows, out, n_threads):
    """``ctk_gather_rows``: out[r, j] = sources[j][rows[r]] (array) or sources[j] (scalar)."""
    n_cols = len(sources)
    ptrs = (ctypes.c_void_p * n_cols)()
    scalars = (ctypes.c_double * n_cols)()
    for j, src in enumerate(sources):
        if isinstance(src, np.ndarray):
            assert src.dtype == np.float64 and src.flags.c_contiguous
            ptrs[j] = src.ctypes.data
        else:
            ptrs[j] = None
            scalars[j] = float(src)
    assert out.flags.c_contiguous and out.dtype == np.float64 and rows.dtype == np.int64
    check(load().ctk_gather_rows(ptrs, scalars, rows.ctypes.data, len(rows), n_cols,
                                 out.ctypes.data, int(n_threads)), "ctk_gather_rows")


def scatter_rows(params, params_in, rows, group_offset, group_cost, group_status, block, cost_out,
                 n_threads, row_base=0):
    """``ctk_scatter_rows``: write one chunk back into the table-order column block [P, N] and the
    cost column (table row = row_base + rows[r]); returns the number of failed clusters."""
    n_cols = len(block)                       # [P, N] array or a list of P column arrays
    ptrs = (ctypes.c_void_p * n_cols)(*[block[j].ctypes.data for j in range(n_cols)])
    failed = ctypes.c_int64(0)
    for arr in [params, params_in, rows, group_offset, group_cost, group_status, cost_out] + list(block):
        assert arr.flags.c_contiguous
    check(load().ctk_scatter_rows(params.ctypes.data, params_in.ctypes.data, rows.ctypes.data,
                                  int(row_base), len(rows), n_cols, group_offset.ctypes.data,
                                  group_cost.ctypes.data, group_status.ctypes.data,
                                  len(group_status), ptrs, cost_out.ctypes.data, int(n_threads),
                                  ctypes.byref(failed)), "ctk_scatter_rows")
    return failed.value


def schedule(cluster_offset, caps, class_target):
    """``ctk_schedule`` -> (work ids grouped by target class with the expensive clusters first,
    clusters per target class, ids that cannot run)."""
    cluster_offset = np.ascontiguousarray(cluster_offset, dtype=np.int32)
    caps = np.ascontiguousarray(caps, dtype=np.int32)
    class_target = np.ascontiguousarray(class_target, dtype=np.int32)
    n = len(cluster_offset) - 1
    work = np.empty(max(n, 1), dtype=np.int32)
    counts = np.zeros(len(caps), dtype=np.int64)
    not_run = np.empty(max(n, 1), dtype=np.int32)
    n_not = ctypes.c_int64(0)
    check(load().ctk_schedule(cluster_offset.ctypes.data, n, caps.ctypes.data, len(caps),
                              class_target.ctypes.data, work.ctypes.data, counts.ctypes.data,
                              not_run.ctypes.data, ctypes.byref(n_not)), "ctk_schedule")
    n_run = int(counts.sum())
    return work[:n_run], counts, not_run[:n_not.value]


def column_pointers(sources):
    """(pointer array, scalar array) describing table columns for the gather helpers: an ndarray
    (float64, contiguous) or a constant per column.  Keep the returned objects alive during the call."""
    n_cols = len(sources)
    ptrs = (ctypes.c_void_p * n_cols)()
    scalars = (ctypes.c_double * n_cols)()
    for j, src in enumerate(sources):
        if isinstance(src, np.ndarray):
            assert src.dtype == np.float64 and src.flags.c_contiguous
            ptrs[j] = src.ctypes.data
        else:
            ptrs[j] = None
            scalars[j] = float(src)
    return ptrs, scalars


def frame_runs(frames):
    """``ctk_frame_runs`` for a contiguous int64 column -> (run starts int64, sorted?)."""
    n = len(frames)
    capacity = 1 << 16
    while True:
        starts = np.empty(capacity, dtype=np.int64)
        runs, is_sorted = ctypes.c_int64(0), ctypes.c_int32(0)
        check(load().ctk_frame_runs(frames.ctypes.data, n, starts.ctypes.data, capacity,
                                    ctypes.byref(runs), ctypes.byref(is_sorted)), "ctk_frame_runs")
        if runs.value <= capacity:
            return starts[:runs.value].copy(), bool(is_sorted.value)
        capacity = int(runs.value)


def label_frames_scratch_bytes(max_points, ndim, n_frames):
    """``ctk_label_frames_scratch``: bytes of device scratch ``ctk_label_frames`` wants."""
    out = ctypes.c_int64(0)
    check(load().ctk_label_frames_scratch(int(max_points), int(ndim), int(n_frames), ctypes.byref(out)),
          "ctk_label_frames_scratch")
    return int(out.value)


def label_frames_device(pos_ptrs, ndim, d_starts, d_stops, n_frames, max_points, separation, d_labels,
                        d_flags, d_scratch, scratch_bytes, stream):
    """``ctk_label_frames`` (asynchronous): device pointers as ints; ``pos_ptrs`` = ndim column pointers."""
    cols = (ctypes.c_void_p * 3)(*([int(p) for p in pos_ptrs] + [None] * (3 - ndim)))
    separation = np.ascontiguousarray(separation, dtype=np.float64)
    check(load().ctk_label_frames(cols, int(ndim), int(d_starts), int(d_stops), int(n_frames),
                                  int(max_points), separation.ctypes.data, int(d_labels),
                                  int(d_flags), int(d_scratch),
                                  int(scratch_bytes), int(stream) if stream else None),
          "ctk_label_frames")


def cluster_pack_frames(pos, starts, stops, separation, n_threads, sources, row_base, params_out,
                        labels=None, flags=None, cluster_out=None, size_out=None, by_cluster_out=None,
                        group_start_out=None):
    """``ctk_cluster_pack_columns`` -> (labels local to each frame, sizes, by_cluster, spans,
    group counts per frame, group starts); ``params_out`` [n, P] receives the packed rows.
    ``pos``: [n, ndim] array of this call's rows, or a list of ndim table-order float64 columns
    (then this call's rows are rows ``row_base .. row_base + len(params_out)`` of the table).
    ``labels`` (int32, this call's rows) and ``flags`` (int32 per frame): labels computed on the
    device by ``ctk_label_frames``; frames with a non-zero flag are labelled here."""
    pos_cols = None
    if isinstance(pos, (list, tuple)):
        ndim, n = len(pos), len(params_out)
        pos_cols, _ = column_pointers(list(pos))
        pos_ptr = None
    else:
        pos = np.ascontiguousarray(pos, dtype=np.float64)
        n, ndim = pos.shape
        pos_ptr = pos.ctypes.data
    starts = np.ascontiguousarray(starts, dtype=np.int64)
    stops = np.ascontiguousarray(stops, dtype=np.int64)
    separation = np.ascontiguousarray(separation, dtype=np.float64)
    cluster = np.empty(n, dtype=np.int64) if cluster_out is None else cluster_out
    size = np.empty(n, dtype=np.int64) if size_out is None else size_out
    for arr in (cluster, size):                    # optional caller-owned outputs (views are fine)
        assert arr.dtype == np.int64 and arr.flags.c_contiguous and len(arr) == n
    by_cluster = np.empty(n, dtype=np.int64) if by_cluster_out is None else by_cluster_out
    spans = np.zeros(len(starts), dtype=np.int64)
    gcount = np.zeros(len(starts), dtype=np.int32)
    gstart = np.empty(max(n, 1), dtype=np.int32) if group_start_out is None else group_start_out
    assert by_cluster.dtype == np.int64 and len(by_cluster) == n and by_cluster.flags.c_contiguous
    assert gstart.dtype == np.int32 and len(gstart) >= n and gstart.flags.c_contiguous
    ptrs, scalars = column_pointers(sources)
    assert params_out.flags.c_contiguous and params_out.dtype == np.float64
    if labels is not None:
        assert labels.dtype == np.int32 and labels.flags.c_contiguous and len(labels) == n
        assert flags.dtype == np.int32 and flags.flags.c_contiguous and len(flags) == len(starts)
        check(load().ctk_cluster_pack_labelled(
            pos_ptr, pos_cols, n, ndim, starts.ctypes.data, stops.ctypes.data, len(starts),
            separation.ctypes.data, int(n_threads), cluster.ctypes.data, size.ctypes.data,
            by_cluster.ctypes.data, spans.ctypes.data, ptrs, scalars, len(sources), int(row_base),
            params_out.ctypes.data, gcount.ctypes.data, gstart.ctypes.data, labels.ctypes.data,
            flags.ctypes.data), "ctk_cluster_pack_labelled")
        return cluster, size, by_cluster, spans, gcount, gstart
    check(load().ctk_cluster_pack_columns(
        pos_ptr, pos_cols, n, ndim, starts.ctypes.data, stops.ctypes.data, len(starts),
        separation.ctypes.data, int(n_threads), cluster.ctypes.data, size.ctypes.data,
        by_cluster.ctypes.data, spans.ctypes.data, ptrs, scalars, len(sources), int(row_base),
        params_out.ctypes.data, gcount.ctypes.data, gstart.ctypes.data), "ctk_cluster_pack_columns")
    return cluster, size, by_cluster, spans, gcount, gstart


def concat_groups(starts, stops, gcount, gstart, frame_base):
    """``ctk_concat_groups`` -> (group_offset int32 [g + 1], group_frame int32 [g])."""
    total = int(gcount.sum())
    goff = np.empty(total + 1, dtype=np.int32)
    gframe = np.empty(max(total, 1), dtype=np.int32)
    n_groups = ctypes.c_int64(0)
    check(load().ctk_concat_groups(starts.ctypes.data, stops.ctypes.data, gcount.ctypes.data,
                                   gstart.ctypes.data, len(starts), int(frame_base), goff.ctypes.data,
                                   gframe.ctypes.data, ctypes.byref(n_groups)), "ctk_concat_groups")
    return goff, gframe[:total]


def apply_label_offsets(local, starts, stops, frame_offset, n_threads, out):
    """``ctk_apply_label_offsets``: out[i] = local[i] + frame_offset[frame of row i]."""
    for arr in (local, starts, stops, frame_offset, out):
        assert arr.flags.c_contiguous and arr.dtype == np.int64
    check(load().ctk_apply_label_offsets(local.ctypes.data, starts.ctypes.data, stops.ctypes.data,
                                         frame_offset.ctypes.data, len(starts), int(n_threads),
                                         out.ctypes.data), "ctk_apply_label_offsets")


def drop_close_frames(coords, values, counts, separation, n_threads):
    """``ctk_drop_close_frames`` -> keep flags [n_frames, capacity] (bool) for the maxima arrays of
    ``ctk_find_maxima`` (coords int32 [n_frames, capacity, ndim], values int32, counts int32)."""
    n_frames, capacity, ndim = coords.shape
    for arr in (coords, values, counts):
        assert arr.flags.c_contiguous and arr.dtype == np.int32
    separation = np.ascontiguousarray(separation, dtype=np.float64)
    keep = np.zeros((n_frames, capacity), dtype=np.uint8)
    kept = np.zeros(n_frames, dtype=np.int32)
    check(load().ctk_drop_close_frames(coords.ctypes.data, values.ctypes.data, counts.ctypes.data,
                                       n_frames, capacity, ndim, separation.ctypes.data,
                                       int(n_threads), keep.ctypes.data, kept.ctypes.data),
          "ctk_drop_close_frames")
    return keep.view(np.bool_)
