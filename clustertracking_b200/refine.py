"""``refine_leastsq`` -- drop-in for clustertracking/refine.py:82-452 with the per-cluster fits on a B200.

Host side (this file): argument handling exactly as the reference (refine.py:242-305), cluster
grouping (find.py), vectorised per-feature bounds (fitfunc.py:538-551), packing of every
``(frame, cluster)`` group into flat ragged buffers, binning of the clusters by size, and the
write-back of the result columns (refine.py:408-427).  Device side (csrc/, through the C ABI of
include/ctk.h): everything inside the reference's ``for _, f_iter in iterable`` loop.

There is no CPU fallback: options the CUDA solver does not carry raise ``NotImplementedError``.
"""
import logging
import os

import numpy as np

from . import _lib
from . import constraints as _constraints
from .find import ChunkLabeller, DeviceLabels, Prestaged, cluster_table, device_labelling_enabled
from .fitfunc import FitFunctions
from .utils import guess_pos_columns, host_threads, is_isotropic, validate_tuple

logger = logging.getLogger(__name__)

# clusters are launched in bins of at most this many features (shared memory is sized per bin)
_BINS = (1, 2, 3, 4, 6, 8, 12, 16, 24, 32, 64, 128, 256)   # > 32: large-cluster kernels (workspace)
_FRAME_BATCH_BYTES = 128 << 20        # frames per upload batch (uploads overlap the kernels)
# scaled step below which the factorised normal matrix is reused (chord iterations)
_CHORD_TOL = float(os.environ.get('CTK_CHORD_TOL', 0.02))
_BIG_WORKSPACE_LIMIT = 8 << 30         # device scratch a large-cluster launch may take
# basin search of the ring / disc families (ctk.h: probe_step, probe_sweeps)
_PROBE_STEP = float(os.environ.get('CTK_PROBE_STEP', 0.25))
_PROBE_SWEEPS = int(os.environ.get('CTK_PROBE_SWEEPS', 2))


_THREADS = None


def _parallel(fn, items):
    """Run ``fn`` over ``items`` on a small thread pool (numpy's take/put loops release the GIL)."""
    global _THREADS
    items = list(items)
    if len(items) < 2:
        for item in items:
            fn(item)
        return
    if _THREADS is None:
        from concurrent.futures import ThreadPoolExecutor
        _THREADS = ThreadPoolExecutor(host_threads(8))
    list(_THREADS.map(fn, items))


def _spans(n, target=1 << 18):
    """Row ranges of about ``target`` rows."""
    k = max(1, min(8, n // target))
    cuts = np.linspace(0, n, k + 1).astype(np.int64)
    return list(zip(cuts[:-1], cuts[1:]))


class Plan(object):
    """Everything ``refine_leastsq`` knows after the host-side preparation; plain numpy.

    ``order`` maps packed rows to rows of ``f`` (the frame-sorted copy): packed row k is
    ``f.iloc[order[k]]``; clusters are consecutive packed rows, sorted by (frame, cluster id)."""

    def __init__(self):
        self.f = None
        self.ff = None
        self.problem = None
        self.frame_numbers = None     # sorted unique frame numbers
        self.frame_source = None      # mapping frame number -> image
        self.frame_shape = None
        self.pixel_dtype = None
        self.order = None
        self.cluster_offset = None    # int32 [n_clusters + 1]
        self.cluster_frame = None     # int32 [n_clusters] index into frame_numbers
        self.params_in = None         # float64 [N, P]
        self.bounds_lo = None         # optional float64 [N, P]; None = device derives them from
        self.bounds_hi = None         # the tables in ``problem`` (fitfunc.py:538-551)

    @property
    def n_clusters(self):
        return len(self.cluster_frame)

    def cluster_sizes(self):
        return np.diff(self.cluster_offset)


class Result(object):
    def __init__(self, plan, allocate=True):
        if allocate:
            self.params_out = plan.params_in.copy()
            self.cost = np.full(plan.n_clusters, np.nan)
            self.status = np.full(plan.n_clusters, -1, dtype=np.int32)
            self.stats = np.zeros((plan.n_clusters, 8), dtype=np.int32)   # CTK_STAT_* counters

    @property
    def iters(self):
        return self.stats[:, 0]


# --------------------------------------------------------------------------------------------------
# host-side preparation
# --------------------------------------------------------------------------------------------------
def _normalise_reader(f, reader, t_column):
    """refine.py:252-283.  Returns (frame source, ndim).  May add ``t_column`` to the caller's
    ``f`` in place, exactly like the reference does."""
    try:
        return reader, len(reader.frame_shape)
    except AttributeError:
        pass
    try:
        ndim = reader.ndim
    except AttributeError:
        raise ValueError('For multiple frames, the reader should be a FramesSequence object '
                         'exposing the "frame_shape" attribute')
    frame_no = getattr(reader, 'frame_no', None)
    if frame_no is not None:
        frame_no = int(frame_no)
        if t_column in f:
            assert np.all(f[t_column] == frame_no)
        else:
            f[t_column] = frame_no
        return {frame_no: reader}, ndim
    if t_column in f:
        assert f[t_column].nunique() == 1
        return {int(f[t_column].iloc[0]): reader}, ndim
    f[t_column] = 0
    return {0: reader}, ndim


def _solver_options(kwargs, compute_default):
    """``**kwargs`` of the reference go to scipy.optimize.minimize (refine.py:225-228, 242-244).
    The CUDA solver honours ``options['maxiter']``; ``tol`` only matters to SLSQP and is accepted
    and ignored (the device solver always converges tighter than SLSQP's default).  Extra keys:
    ``precision`` ('float32' | 'float64' pixel arithmetic), ``xtol`` (step tolerance) and
    ``chord_tol`` (scaled step size below which the factorised normal matrix is reused),
    ``probe_step`` / ``probe_sweeps`` (basin search of the ring / disc families, see ctk.h)."""
    kwargs = dict(kwargs)
    method = kwargs.pop('method', 'SLSQP')
    if method != 'SLSQP':
        raise NotImplementedError("only the reference's default method='SLSQP' is mirrored")
    kwargs.pop('tol', None)
    options = dict(kwargs.pop('options', None) or {})
    lm_max_iter = int(options.pop('maxiter', 100))
    options.pop('disp', None)
    precision = kwargs.pop('precision', compute_default)
    xtol = float(kwargs.pop('xtol', 0.))
    chord_tol = float(kwargs.pop('chord_tol', _CHORD_TOL))
    probe = (float(kwargs.pop('probe_step', _PROBE_STEP)), int(kwargs.pop('probe_sweeps', _PROBE_SWEEPS)))
    if kwargs or options:
        raise TypeError("unsupported keyword arguments: %r" % sorted(list(kwargs) + list(options)))
    if precision not in ('float32', 'float64'):
        raise ValueError("precision must be 'float32' or 'float64'")
    return (lm_max_iter, (_lib.COMPUTE_F64 if precision == 'float64' else _lib.COMPUTE_F32), xtol,
            chord_tol, probe)


class Prepared(object):
    """What ``refine_leastsq`` knows before the clustering: the validated arguments, the model
    (``ff``), the filled problem struct and the frames of the call."""


def prepare_common(f, reader, diameter, separation=None, fit_function='gauss', param_mode=None,
                   param_val=None, constraints=None, bounds=None, pos_columns=None,
                   t_column='frame', noise_size=None, threshold=None, max_iter=10, max_shift=1,
                   max_rms_dev=1., residual_factor=100000., compute_error=False, frames_hook=None,
                   allow_global=False, **kwargs):
    """Argument handling of refine.py:242-315 (no clustering, no GPU work) -> :class:`Prepared`.
    ``frames_hook`` is called with the :class:`FrameInfo` as soon as the frames of the call are
    known: ``refine_leastsq`` uses it to start the uploads while the host is still busy."""
    lm_max_iter, compute_dtype, xtol, chord_tol, probe = _solver_options(kwargs, 'float32')
    if pos_columns is None:
        pos_columns = guess_pos_columns(f)
    if compute_error:
        raise NotImplementedError("compute_error (numdifftools Hessian, refine.py:400-406) is not "
                                  "available in the CUDA solver")
    source, ndim = _normalise_reader(f, reader, t_column)
    assert ndim == len(pos_columns)
    if ndim not in (2, 3):
        raise ValueError("only 2D and 3D images are supported")

    diameter = validate_tuple(diameter, ndim)                          # refine.py:285-289
    radius = tuple([int(d // 2) for d in diameter])
    isotropic = is_isotropic(diameter)
    if separation is None:
        separation = diameter

    ff = FitFunctions(fit_function, ndim, isotropic, param_mode)       # refine.py:291
    if any(m == 2 for m in ff.modes) and not allow_global:
        raise NotImplementedError("param_mode 'global' couples all clusters into one problem "
                                  "(refine.py:319-332): call refine_leastsq, which runs it through "
                                  "clustertracking_b200.global_fit")
    if any(m > 3 for m in ff.modes):
        raise NotImplementedError("param modes 'particle' and 'frame' are not implemented")
    if len(ff.params) > _lib.CTK_MAX_PARAMS:
        raise NotImplementedError("too many parameters per feature")
    if max(radius) > _lib.CTK_MAX_RADIUS or min(radius) < 1:
        raise NotImplementedError("mask radius (diameter // 2) must be within [1, %d]"
                                  % _lib.CTK_MAX_RADIUS)
    cons = _constraints.parse(constraints, ndim)
    if len(f) == 0:
        raise ValueError("no features to refine")

    info = FrameInfo(source, f[t_column].values, ndim)
    if frames_hook is not None:
        frames_hook(info)
    tables = ff.validate_bounds(bounds, radius=radius)                 # refine.py:315

    prob = _lib.Problem()
    prob.ndim, prob.isotropic, prob.family = ndim, int(isotropic), ff.family
    prob.n_params = len(ff.params)
    for j, m in enumerate(ff.modes):
        prob.modes[j] = m
    for k, r in enumerate(radius):
        prob.radius[k] = r
    prob.pixel_dtype = _lib.PIXEL_CODES[np.dtype(info.dtype)]
    prob.compute_dtype = compute_dtype
    prob.max_iter, prob.lm_max_iter = int(max_iter), lm_max_iter
    prob.max_shift, prob.max_rms_dev = float(max_shift), float(max_rms_dev)
    prob.residual_factor, prob.xtol, prob.chord_tol = float(residual_factor), xtol, chord_tol
    prob.probe_step, prob.probe_sweeps = probe
    mask = 0
    if cons['dimer'] is not None:
        mask |= _lib.CONSTRAINT_DIMER
        for k in range(ndim):
            prob.dimer_dist[k] = cons['dimer'][k]
    if cons['trimer'] is not None:
        mask |= _lib.CONSTRAINT_TRIMER
        for k in range(ndim):
            prob.trimer_dist[k] = cons['trimer'][k]
    if cons['tetramer'] is not None:
        mask |= _lib.CONSTRAINT_TETRAMER
        for k in range(ndim):
            prob.tetramer_dist[k] = cons['tetramer'][k]
    prob.constraint_mask = mask
    if noise_size is not None:                                         # refine.py:36-40
        # lowpass of the cluster's sub-image with the taps of trackpy.masks.gaussian_kernel(sigma, 4)
        # (preprocessing.py:41-44, built on the device); a size <= 0 leaves that axis unfiltered
        prob.lowpass = 1
        prob.lowpass_threshold = 0. if threshold is None else float(threshold)
        for k, sigma in enumerate(validate_tuple(noise_size, ndim)):
            if not sigma > 0:
                prob.lowpass_half[k] = -1
                continue
            lw = int(4.0 * sigma + 0.5)
            if 2 * lw + 1 > _lib.CTK_MAX_TAPS:
                raise NotImplementedError("noise_size %r: the lowpass kernel is limited to a half "
                                          "width of %d pixels" % (sigma, _lib.CTK_MAX_TAPS // 2))
            prob.lowpass_half[k] = lw
            prob.lowpass_sigma[k] = float(sigma)
    for which, table in zip(("bounds_abs", "bounds_diff", "bounds_rel"), tables):
        dst = getattr(prob, which)
        for side in range(2):
            for j in range(_lib.CTK_MAX_PARAMS):
                dst[side][j] = table[side, j] if j < len(ff.params) else np.nan

    pre = Prepared()
    pre.ff, pre.problem, pre.info = ff, prob, info
    pre.pos_columns, pre.t_column, pre.separation = list(pos_columns), t_column, separation
    pre.param_val, pre.ndim = param_val, ndim
    return pre


def _plan_for(pre, order, cluster_offset, cluster_frame, params_in):
    plan = Plan()
    plan.ff, plan.problem = pre.ff, pre.problem
    plan.order, plan.cluster_offset, plan.cluster_frame = order, cluster_offset, cluster_frame
    plan.params_in = params_in
    info = pre.info
    plan.frame_numbers, plan.frame_source = info.numbers, info.source
    plan.frame_shape, plan.pixel_dtype, plan.frame_info = info.shape, info.dtype, info
    return plan


def prepare(f, reader, diameter, separation=None, fit_function='gauss', param_mode=None,
            param_val=None, constraints=None, bounds=None, pos_columns=None, t_column='frame',
            noise_size=None, threshold=None, max_iter=10, max_shift=1, max_rms_dev=1.,
            residual_factor=100000., compute_error=False, frames_hook=None, empty=np.empty,
            allow_global=False, **kwargs):
    """Host half of ``refine_leastsq`` for the whole table at once: returns a :class:`Plan` (no GPU
    work).  ``empty(shape, dtype)`` allocates the packed parameter table."""
    pre = prepare_common(f, reader, diameter, separation, fit_function, param_mode, param_val,
                         constraints, bounds, pos_columns, t_column, noise_size, threshold,
                         max_iter, max_shift, max_rms_dev, residual_factor, compute_error,
                         frames_hook=frames_hook, allow_global=allow_global, **kwargs)
    ff, info = pre.ff, pre.info
    import time as _time
    _t0 = _time.perf_counter()
    f, order = cluster_table(f, pre.separation, pre.pos_columns, t_column)   # refine.py:297 (a copy)
    _t1 = _time.perf_counter()
    if param_val is not None:                                          # refine.py:300-302
        for col in param_val:
            f[col] = param_val[col]
    for col in ff.params:                                              # refine.py:303-305
        if col not in f.columns:
            f[col] = ff.default[col]

    # (frame, cluster) groups in the order of f.groupby(['frame', 'cluster'])     refine.py:336
    # ``order`` lists the rows by (frame, cluster), row order kept inside a cluster
    frames_s, cluster_s = f[t_column].values[order], f['cluster'].values[order]
    n = len(order)
    if n == 0:
        raise ValueError("no features to refine")
    new_group = np.empty(n, dtype=bool)
    new_group[0] = True
    new_group[1:] = (frames_s[1:] != frames_s[:-1]) | (cluster_s[1:] != cluster_s[:-1])
    starts = np.flatnonzero(new_group)
    params_in = empty((n, len(ff.params)), np.float64)
    columns = [np.asarray(f[col].values, dtype=np.float64) for col in ff.params]

    def gather(span):                                                  # packed, group order
        a, b = span
        rows = order[a:b]
        for j, col in enumerate(columns):
            params_in[a:b, j] = col[rows]

    _parallel(gather, _spans(n))
    plan = _plan_for(pre, order, np.concatenate((starts, [n])).astype(np.int32),
                     np.searchsorted(info.sorted_numbers, frames_s[starts]).astype(np.int32),
                     params_in)
    plan.f = f
    plan.solver = dict(max_iter=int(max_iter), max_shift=float(max_shift),
                       max_rms_dev=float(max_rms_dev), residual_factor=float(residual_factor))
    plan.timing = dict(cluster_ms=1e3 * (_t1 - _t0), pack_ms=1e3 * (_time.perf_counter() - _t1))
    return plan


class FrameInfo(object):
    """The frames one call touches: source, sorted unique frame numbers, shape and pixel type."""

    def __init__(self, source, frame_column, ndim):
        frames = np.asarray(frame_column)
        self.run_starts = None                  # rows at which a new frame starts (sorted tables)
        if len(frames) > 1 and frames.dtype == np.int64 and frames.flags.c_contiguous:
            starts, is_sorted = _lib.frame_runs(frames)             # one pass in C
            if is_sorted:
                self.run_starts = starts
        elif len(frames) > 1:
            change = frames[1:] != frames[:-1]
            cuts = np.flatnonzero(change) + 1
            # sorted <=> every change is an increase
            if np.all(frames[cuts] > frames[cuts - 1]):
                self.run_starts = np.concatenate(([0], cuts)).astype(np.int64)
        if self.run_starts is not None:
            uniq = frames[self.run_starts]
        else:
            uniq = np.unique(frames)
        self.source = source
        self.sorted_numbers = uniq
        self.numbers = [int(x) if float(x).is_integer() else x for x in uniq]
        first = np.asarray(source[self.numbers[0]])
        self.shape = tuple(first.shape)
        if len(self.shape) != ndim:
            raise ValueError("frames must have %d dimensions" % ndim)
        self.dtype = first.dtype if first.dtype in _lib.PIXEL_CODES else np.dtype(np.float64)
        # attribute names shared with Plan, so that load_frame() serves both
        self.frame_source, self.frame_shape, self.pixel_dtype = source, self.shape, self.dtype
        self.frame_numbers = self.numbers


def load_frame(plan, frame_no):
    """One frame as a C-contiguous array of the plan's pixel type."""
    image = np.asarray(plan.frame_source[frame_no])
    if image.shape != plan.frame_shape:
        raise ValueError("frame %r has shape %r, expected %r" % (frame_no, image.shape,
                                                                  plan.frame_shape))
    return np.ascontiguousarray(image, dtype=plan.pixel_dtype)


def rigorous_problem(problem):
    """Copy of the problem struct that sizes shared memory for the worst case (relaunches)."""
    import ctypes
    clone = _lib.Problem()
    ctypes.memmove(ctypes.byref(clone), ctypes.byref(problem), ctypes.sizeof(_lib.Problem))
    clone.capacity_mode = 1
    return clone


def bin_clusters(sizes, cluster_ids):
    """-> list of (capacity, ids) with ids sorted by descending size (expensive clusters first)."""
    out = []
    lower = 0
    for cap in _BINS:
        sel = cluster_ids[(sizes[cluster_ids] > lower) & (sizes[cluster_ids] <= cap)]
        if len(sel):
            sel = sel[np.argsort(-sizes[sel], kind='stable')]
            out.append((cap, sel.astype(np.int32)))
        lower = cap
    return out


def run_bins(sizes, cluster_ids, status, launch):
    """Launch every bin; clusters that overflowed their bin's typical-case capacity (status
    TOO_LARGE) are relaunched once with rigorous capacities, and what even those cannot hold (dense
    clusters with more overlapping pairs than 4 n) once more in the first large-cluster class --
    the host-driven mirror of ``DeviceSession._run`` (used by the emulated backend of the tests).
    ``launch(capacity, ids, rigorous)`` must fill ``status[ids]``."""
    too_big = cluster_ids[sizes[cluster_ids] > _BINS[-1]]
    status[too_big] = _lib.STATUS_TOO_LARGE
    first_big = min(c for c in _BINS if c > _lib.CTK_MAX_CLUSTER_FEATURES)
    leftover = []
    for cap, ids in bin_clusters(sizes, cluster_ids):
        launch(cap, ids, False)
        retry = ids[status[ids] == _lib.STATUS_TOO_LARGE]
        if len(retry) and cap < _lib.CTK_MAX_CLUSTER_FEATURES:
            launch(cap, retry, True)              # same class, rigorous capacities
            retry = retry[status[retry] == _lib.STATUS_TOO_LARGE]
        if len(retry) and cap <= _lib.CTK_MAX_CLUSTER_FEATURES:
            leftover.append(retry)
    if leftover:
        launch(first_big, np.concatenate(leftover), False)


def finalize(plan, result):
    """Write the fitted parameters and ``cost`` into the DataFrame (refine.py:408-427).  The device
    already copies the input parameters through for failed clusters, so the whole table is written
    with one scatter per column."""
    f, ff = plan.f, plan.ff
    sizes = plan.cluster_sizes()
    ok = result.status == 0
    block = np.empty((len(ff.params), len(f)), dtype=np.float64)
    source = result.params_out
    if not ok.all():                       # belt and braces: failed clusters keep their input exactly
        rows = np.repeat(~ok, sizes)
        source = source.copy()
        source[rows] = plan.params_in[rows]

    def scatter(j):
        block[j, plan.order] = source[:, j]

    _parallel(scatter, range(len(ff.params)))
    for j, col in enumerate(ff.params):
        f[col] = block[j]
    cost = np.empty(len(f), dtype=np.float64)
    cost[plan.order] = np.repeat(np.where(ok, result.cost, np.nan), sizes)
    f['cost'] = cost
    failed = np.flatnonzero(~ok)
    if len(failed):
        first_row = plan.order[plan.cluster_offset[:-1][failed]]
        ids = f['cluster'].values[first_row]
        for c, cid in zip(failed[:20], ids[:20]):
            logger.warning("RefineException: cluster %d: %s", int(cid),
                           _lib.STATUS_NAMES.get(int(result.status[c]), "status %d" % result.status[c]))
        if len(failed) > 20:
            logger.warning("RefineException: ... and %d more clusters failed", len(failed) - 20)
    return f


# --------------------------------------------------------------------------------------------------
# device execution
# --------------------------------------------------------------------------------------------------
_PINNED = {}          # cached pinned host buffers for the result download (not thread-safe)


def _pinned_buffer(torch, key, nbytes):
    """Reusable pinned staging buffer: pinning costs ~0.6 ms per MB, a pageable device->host copy
    runs at ~2 GB/s, a pinned one at ~50 GB/s (measured on the B200 boxes)."""
    buf = _PINNED.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, pin_memory=True)
        _PINNED[key] = buf
    return buf


class _PinnedArena(object):
    """Bump allocator over one cached pinned buffer for the small per-launch uploads (cluster
    offsets, frame indices, work ids).  A pageable source would make every ``.to(device)`` wait for
    the kernels already queued on the stream; from pinned memory the copies are asynchronous.
    ``reset()`` is called at the start of a ``refine_leastsq`` call, after the previous call has
    synchronised, so the memory is not reused while a copy is in flight."""

    def __init__(self):
        self.offset = 0
        self.fallback = []

    def reset(self):
        self.offset = 0
        self.fallback = []

    def stage(self, torch, array):
        """-> pinned torch tensor holding a copy of ``array`` (1-D view of its bytes as its dtype)."""
        array = np.ascontiguousarray(array)
        nbytes = array.nbytes
        buf = _PINNED.get("arena")
        if buf is None:
            buf = _pinned_buffer(torch, "arena", 64 << 20)
        start = (self.offset + 255) & ~255
        if start + nbytes > buf.numel():          # does not fit: a one-off pinned tensor
            t = torch.empty(max(nbytes, 1), dtype=torch.uint8, pin_memory=True)
            self.fallback.append(t)
            view = t[:nbytes]
        else:
            view = buf[start:start + nbytes]
            self.offset = start + nbytes
        host = view.numpy().view(array.dtype).reshape(array.shape)
        host[...] = array
        return torch.from_numpy(host)


_ARENA = _PinnedArena()
_FITS_CACHE = {}      # (problem bytes) -> per-class capacity answers of the library


class FrameSet(object):
    """The frames of one call on the device: a table of frame pointers and the per-frame maxima.
    ``upload_async`` enqueues the copies batch by batch on a copy stream and the ``ctk_frame_max``
    launches behind them; nothing blocks the host, so clustering and packing run meanwhile."""

    def __init__(self, info, device=None):
        import torch
        self.torch = torch
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise RuntimeError("clustertracking_b200 needs a CUDA device (B200, sm_100a); "
                               "there is no CPU fallback")
        self.dev = (torch.device('cuda', torch.cuda.current_device()) if device is None
                    else torch.device(device))
        self.info = info
        self.n_frames = len(info.numbers)
        self.n_pixels = int(np.prod(info.shape))
        self.frame_bytes = self.n_pixels * np.dtype(info.dtype).itemsize
        self.pixel_code = _lib.PIXEL_CODES[np.dtype(info.dtype)]
        self.torch_dtype = {
            np.dtype(np.uint8): torch.uint8, np.dtype(np.uint16): torch.uint16,
            np.dtype(np.float32): torch.float32, np.dtype(np.float64): torch.float64,
            np.dtype(np.int16): torch.int16, np.dtype(np.int32): torch.int32,
        }[np.dtype(info.dtype)]
        self.launches = 0
        self.h2d_bytes = 0
        self.tensors = []                 # keeps the uploaded batches alive
        self.big_workspaces = {}          # large-cluster scratch, shared by the chunks of the call
        self.batch_ready, self.batch_last = [], []       # per upload batch: event, last frame
        self.recorded, self.thread, self.upload_error = None, None, None   # background staging
        with torch.cuda.device(self.dev):
            self.d_ptrs = torch.zeros(self.n_frames, dtype=torch.int64, device=self.dev)
            self.d_fmax = torch.empty(self.n_frames, dtype=torch.float64, device=self.dev)

    def stream_ptr(self):
        return _lib.ctypes.c_void_p(self.torch.cuda.current_stream(self.dev).cuda_stream)

    def _direct_view(self, f0, f1):
        """Zero-copy host view of frames f0..f1-1 when the reader is backed by one contiguous array
        of the right type (``reader.stack``); pinned memory then goes to the device by DMA."""
        src = self.info.source
        stack = getattr(src, 'stack', None)
        first = getattr(src, 'first_frame', 0)
        if not isinstance(stack, np.ndarray) or stack.dtype != self.info.dtype:
            return None
        if stack.shape[1:] != tuple(self.info.shape) or not stack.flags.c_contiguous:
            return None
        idx = np.asarray(self.info.numbers[f0:f1], dtype=np.int64) - first
        if idx[0] < 0 or idx[-1] >= len(stack) or np.any(np.diff(idx) != 1):
            return None
        return stack[idx[0]:idx[-1] + 1]

    def register(self, d_frames, f0):
        """Frames f0 .. f0+n-1 are resident in ``d_frames`` [n, *frame_shape]."""
        n = int(d_frames.shape[0])
        ptrs = d_frames.data_ptr() + self.frame_bytes * np.arange(n, dtype=np.int64)
        self.d_ptrs[f0:f0 + n] = self.torch.from_numpy(ptrs).to(self.dev)
        self.tensors.append(d_frames)

    def launch_frame_max(self, f0, n, events=None, stream=None):
        if events is not None:
            start, stop = (self.torch.cuda.Event(enable_timing=True) for _ in range(2))
            start.record()
        stream_ptr = (self.stream_ptr() if stream is None
                      else _lib.ctypes.c_void_p(stream.cuda_stream))
        _lib.check(self.lib.ctk_frame_max(self.d_ptrs.data_ptr() + 8 * f0, n, self.n_pixels,
                                          self.pixel_code, self.d_fmax.data_ptr() + 8 * f0,
                                          stream_ptr), "ctk_frame_max")
        self.launches += 1
        if events is not None:
            stop.record()
            events.append(("frame_max", start, stop, None))

    def wait_for_frames(self, last_frame):
        """Make the current stream wait until frames 0 .. last_frame are resident and their maxima
        computed (a launch that touches only the first frames need not wait for the whole video)."""
        if not self.batch_last:
            return
        k = int(np.searchsorted(self.batch_last, last_frame))
        k = min(k, len(self.batch_last) - 1)
        if self.recorded is not None:
            # staged upload on a background thread: the event exists once that thread has enqueued
            # the batch (waiting on an event that was never recorded would be a no-op)
            self.recorded[k].wait()
            if self.upload_error is not None:
                raise self.upload_error
        self.torch.cuda.current_stream(self.dev).wait_event(self.batch_ready[k])

    def close(self):
        """Join the staging thread (if any) before the pinned ring can be reused."""
        thread, self.thread = self.thread, None
        if thread is not None:
            thread.join()
        if self.upload_error is not None:
            error, self.upload_error = self.upload_error, None
            raise error

    def upload_async(self):
        """Allocate the device frames, upload the pointer table once (the only host-synchronous
        step, done while the device is idle), then enqueue copy -> frame max per batch on a copy
        stream; ``wait_for_frames`` orders later launches behind the batches they read.

        Frames that already sit in pinned host memory go to the device by DMA straight from the
        caller's array.  Anything else (a pageable numpy stack, a reader that produces frames one
        by one) is staged on a BACKGROUND thread through a ring of two pinned buffers: while batch
        k travels to the device, batch k + 1 is copied (or read) into the other buffer, and the
        caller goes on with the cluster labelling meanwhile.  (A pageable ``copy_`` would block the
        host for the whole upload at ~2-6 GB/s: config 4 lost 6.5x to that.)"""
        torch = self.torch
        per_batch = max(1, min(self.n_frames, _FRAME_BATCH_BYTES // max(self.frame_bytes, 1)))
        cuts = list(range(0, self.n_frames, per_batch)) + [self.n_frames]
        self.recorded, self.thread, self.upload_error = None, None, None
        with torch.cuda.device(self.dev):
            compute = torch.cuda.current_stream(self.dev)
            copy_stream = torch.cuda.Stream(device=self.dev)
            batches, ptrs = [], np.empty(self.n_frames, dtype=np.int64)
            try:
                for f0, f1 in zip(cuts[:-1], cuts[1:]):
                    d_frames = torch.empty((f1 - f0,) + tuple(self.info.shape), dtype=self.torch_dtype,
                                           device=self.dev)
                    ptrs[f0:f1] = d_frames.data_ptr() + self.frame_bytes * np.arange(f1 - f0, dtype=np.int64)
                    batches.append(d_frames)
            except torch.cuda.OutOfMemoryError:
                # (asked only now: cudaMemGetInfo costs 3 ms per call and stalls behind other threads)
                del batches
                free, _ = torch.cuda.mem_get_info(self.dev)
                raise MemoryError("the %d frames of this call need %.1f GB of device memory, %.1f GB "
                                  "are free: refine the video in blocks of frames"
                                  % (self.n_frames, self.n_frames * self.frame_bytes / 1e9, free / 1e9))
            self.d_ptrs.copy_(torch.from_numpy(ptrs))
            self.tensors.extend(batches)
            copy_stream.wait_stream(compute)
            self.copy_stream = copy_stream
            self.batch_last = [f1 - 1 for f1 in cuts[1:]]
            self.batch_ready = [torch.cuda.Event() for _ in batches]
            views = [self._direct_view(f0, f1) for f0, f1 in zip(cuts[:-1], cuts[1:])]
            pinned = all(v is not None and torch.from_numpy(v).is_pinned() for v in views)

            def enqueue(k, src):
                n = cuts[k + 1] - cuts[k]
                with torch.cuda.stream(copy_stream):
                    batches[k].copy_(src, non_blocking=True)
                    self.launch_frame_max(cuts[k], n, stream=copy_stream)
                    self.batch_ready[k].record(copy_stream)
                self.h2d_bytes += n * self.frame_bytes

            if pinned:
                for k, view in enumerate(views):
                    enqueue(k, torch.from_numpy(view))
                return self

            import threading
            ring = [_pinned_buffer(torch, "frames%d" % j, per_batch * self.frame_bytes)
                    for j in range(2)]
            self.recorded = [threading.Event() for _ in batches]

            def stage():
                try:
                    torch.cuda.set_device(self.dev)
                    for k, view in enumerate(views):
                        n = cuts[k + 1] - cuts[k]
                        if k >= 2:
                            self.batch_ready[k - 2].synchronize()    # its ring slot is free again
                        slot = ring[k % 2][:n * self.frame_bytes].view(self.torch_dtype)
                        slot = slot.reshape((n,) + tuple(self.info.shape))
                        host = slot.numpy()
                        if view is not None:
                            _copy_frames(host, view)
                        else:
                            for j in range(n):
                                host[j] = load_frame(self.info, self.info.numbers[cuts[k] + j])
                        enqueue(k, slot)
                        self.recorded[k].set()
                except BaseException as exc:                  # surface it in the calling thread
                    self.upload_error = exc
                    for flag in self.recorded:
                        flag.set()

            self.thread = threading.Thread(target=stage, daemon=True)
            self.thread.start()
        return self


def _copy_frames(dst, src):
    """dst[...] = src on a few threads (numpy's copy releases the GIL; one thread moves ~10 GB/s)."""
    n = len(src)
    k = max(1, min(n, host_threads(4)))
    if k == 1 or src.nbytes < (8 << 20):
        dst[...] = src
        return
    cuts = np.linspace(0, n, k + 1).astype(int)

    def one(span):
        dst[span[0]:span[1]] = src[span[0]:span[1]]
    _parallel(one, list(zip(cuts[:-1], cuts[1:])))


class DeviceSession(object):
    """Device-side state of one plan.  torch is used for device memory, streams and events only;
    all arithmetic happens inside libctk (C ABI, include/ctk.h).

    Everything is enqueued asynchronously on the current stream: one refine launch per size class
    over all clusters of the call (work ids: expensive clusters first), no host synchronisation
    until the results are downloaded.  Clusters whose pixel lists overflowed their size class
    (status TOO_LARGE) are relaunched once with rigorous capacities from a device-side list."""

    def __init__(self, plan, device=None, frames=None):
        self.frames = frames if frames is not None else FrameSet(plan.frame_info, device)
        torch = self.torch = self.frames.torch
        self.lib = self.frames.lib
        self.dev = self.frames.dev
        self.plan = plan
        self.sizes = plan.cluster_sizes()
        self.n_frames = self.frames.n_frames
        self.shape_arr = (_lib.ctypes.c_int64 * 3)(
            *(list(plan.frame_shape) + [1] * (3 - len(plan.frame_shape))))
        self.launches = 0
        self.h2d_bytes = 0
        self.d2h_bytes = 0
        with torch.cuda.device(self.dev):
            dev = self.dev
            self.workspace = torch.empty(int(self.lib.ctk_refine_workspace_bytes()),
                                         dtype=torch.uint8, device=dev)
            self.d_offset = self._up(plan.cluster_offset)
            self.d_cframe = self._up(plan.cluster_frame)
            self.d_params = self._up(plan.params_in)
            self.d_lo = self._up(plan.bounds_lo) if plan.bounds_lo is not None else None
            self.d_hi = self._up(plan.bounds_hi) if plan.bounds_hi is not None else None
            self.d_out = torch.empty_like(self.d_params)
            self.d_cost = torch.empty(plan.n_clusters, dtype=torch.float64, device=dev)
            self.d_status = torch.full((plan.n_clusters,), -1, dtype=torch.int32, device=dev)
            self.d_stats = torch.zeros((plan.n_clusters, 8), dtype=torch.int32, device=dev)

    def _up(self, array):
        t = self.torch.from_numpy(np.ascontiguousarray(array))
        if not t.is_pinned() and self.torch.cuda.is_available():
            t = _ARENA.stage(self.torch, array)
        self.h2d_bytes += t.numel() * t.element_size()
        return t.to(self.dev, non_blocking=True)

    def stream_ptr(self):
        return _lib.ctypes.c_void_p(self.torch.cuda.current_stream(self.dev).cuda_stream)

    def schedule(self):
        """Work ids of every size class, concatenated and uploaded once: -> list of
        (capacity, start, count) slices.  Inside a class the expensive clusters come first.
        A class whose per-cluster arrays do not fit the shared memory (large masks) runs in the
        first large-cluster class instead (global-memory workspace)."""
        caps = np.asarray(_BINS)
        self.rigorous = rigorous_problem(self.plan.problem)
        small = caps <= _lib.CTK_MAX_CLUSTER_FEATURES
        key = bytes(self.plan.problem)
        if key not in _FITS_CACHE:                 # the capacity queries only depend on the problem
            prob = _lib.ctypes.byref(self.plan.problem)
            rig = _lib.ctypes.byref(self.rigorous)
            fits = np.array([(self.lib.ctk_refine_shared_bytes(prob, int(c)) > 0
                              if c <= _lib.CTK_MAX_CLUSTER_FEATURES else
                              0 < self.lib.ctk_refine_workspace_bytes_for(prob, int(c)) <= _BIG_WORKSPACE_LIMIT)
                             for c in caps] + [False])
            retry_fits = {int(c): self.lib.ctk_refine_shared_bytes(rig, int(c)) > 0
                          for c in caps[small]}
            if len(_FITS_CACHE) > 64:
                _FITS_CACHE.clear()
            _FITS_CACHE[key] = (fits, retry_fits)
        # classes whose typical-case capacities can overflow, and whether the rigorous ones fit
        fits, self.retry_fits = _FITS_CACHE[key]
        first_big = int(np.flatnonzero(~small)[0])
        self.big_fallback = int(caps[first_big]) if fits[first_big] else None
        # class a cluster of class k runs in: k, the first large class when k's arrays do not fit
        # the shared memory, or -1
        target = np.where(fits[:len(caps)], np.arange(len(caps)),
                          np.where(small & bool(self.big_fallback), first_big, -1)).astype(np.int32)
        ids, counts, self.never_run = _lib.schedule(self.plan.cluster_offset, caps, target)
        self.d_work = self._up(ids)
        slices, at = [], 0
        for k, count in enumerate(counts):
            if count:
                slices.append((int(caps[k]), at, int(count)))
                at += int(count)
        # one overflow list [count, ids...] per small class: filled by the class's launch, consumed by
        # its relaunch with rigorous capacities -- no host round trip in between
        self.overflow_at = {}
        words = 0
        for cap, _, count in slices:
            if cap <= _lib.CTK_MAX_CLUSTER_FEATURES:
                self.overflow_at[cap] = (words, count)
                words += 1 + count
        # the FINAL list [count, ids...]: clusters that are still too large for the rigorous
        # capacities of their class (dense clusters: more overlapping pairs than 4 n, a box wider
        # than 1023 pixels) or for the largest shared-memory class; one large-cluster launch
        # (global-memory workspace) behind all classes consumes it
        self.final_at = (words, sum(count for cap, _, count in slices
                                    if cap <= _lib.CTK_MAX_CLUSTER_FEATURES))
        words += 1 + self.final_at[1]
        self.d_overflow = self.torch.zeros(max(words, 1), dtype=self.torch.int32, device=self.dev)
        return slices

    def launch_refine(self, cap, work_ptr, count, events=None, problem=None, n_work_ptr=None,
                      overflow_ptr=None, overflow_cap=0, flags=0, label=None):
        if events is not None:
            start, stop = (self.torch.cuda.Event(enable_timing=True) for _ in range(2))
            start.record()
        _lib.check(self.lib.ctk_refine_batch_ex(
            _lib.ctypes.byref(problem if problem is not None else self.plan.problem),
            self.frames.d_ptrs.data_ptr(), self.shape_arr,
            self.frames.d_fmax.data_ptr(), count, work_ptr, int(cap), self.d_cframe.data_ptr(),
            self.d_offset.data_ptr(), self.d_params.data_ptr(),
            self.d_lo.data_ptr() if self.d_lo is not None else None,
            self.d_hi.data_ptr() if self.d_hi is not None else None, self.d_out.data_ptr(),
            self.d_cost.data_ptr(), self.d_status.data_ptr(), self.d_stats.data_ptr(),
            self._workspace_for(cap).data_ptr(), n_work_ptr, overflow_ptr, int(overflow_cap),
            int(flags), self.stream_ptr()), "ctk_refine_batch_ex")
        self.launches += 1
        if events is not None:
            stop.record()
            events.append(("refine", start, stop, label if label is not None else (cap, "main")))

    def _workspace_for(self, cap):
        """Small classes share one scratch word; a large-cluster class gets its own workspace."""
        if cap <= _lib.CTK_MAX_CLUSTER_FEATURES:
            return self.workspace
        # one workspace per class and CALL (kept by the FrameSet): the chunks of a call run on one
        # stream, so their large-cluster launches can share it
        pool = self.frames.big_workspaces
        if cap not in pool:
            nbytes = int(self.lib.ctk_refine_workspace_bytes_for(
                _lib.ctypes.byref(self.plan.problem), int(cap)))
            pool[cap] = self.torch.empty(nbytes, dtype=self.torch.uint8, device=self.dev)
        return pool[cap]

    def run(self, slices, events=None):
        self.frames.wait_for_frames(int(self.plan.cluster_frame.max()) if self.plan.n_clusters else 0)
        self._run(slices, events)

    def _run(self, slices, events=None):
        """One refine launch per size class; behind it, for the classes provisioned for the typical
        case, the relaunch of the clusters that overflowed (device-side list, usually empty or a
        percent of the class): with rigorous capacities if those fit the shared memory, else in the
        first large-cluster class.  What even the rigorous capacities cannot hold (dense clusters
        with more than 4 n overlapping pairs, boxes wider than 1023 pixels) is appended to the final
        list, which ONE large-cluster launch at the end consumes -- still no host round trip."""
        base = self.d_overflow.data_ptr()
        final = base + 4 * self.final_at[0]
        final_cap = self.final_at[1] if self.big_fallback is not None else 0
        used_final = False

        def to_final():
            """Overflow arguments of a launch that feeds the final list: the first such launch of
            a pass resets the list's count, the later ones append."""
            nonlocal used_final
            if not final_cap:
                return dict()
            flags = _lib.LAUNCH_APPEND_OVERFLOW if used_final else 0
            used_final = True
            return dict(overflow_ptr=final, overflow_cap=final_cap, flags=flags)

        for cap, start, count in slices:
            at = self.overflow_at.get(cap)
            if at is None:
                self.launch_refine(cap, self.d_work.data_ptr() + 4 * start, count, events)
                continue
            lst = base + 4 * at[0]
            if self.retry_fits.get(cap) and cap < _lib.CTK_MAX_CLUSTER_FEATURES:
                self.launch_refine(cap, self.d_work.data_ptr() + 4 * start, count, events,
                                   overflow_ptr=lst, overflow_cap=count)
                self.launch_refine(cap, lst + 4, count, events, problem=self.rigorous,
                                   n_work_ptr=lst, label=(cap, "rigorous"), **to_final())
            elif self.retry_fits.get(cap):               # the largest class is rigorous already
                self.launch_refine(cap, self.d_work.data_ptr() + 4 * start, count, events,
                                   **to_final())
            else:
                self.launch_refine(cap, self.d_work.data_ptr() + 4 * start, count, events,
                                   overflow_ptr=lst, overflow_cap=count)
                if self.big_fallback is not None:
                    self.launch_refine(self.big_fallback, lst + 4, count, events, n_work_ptr=lst,
                                       label=(cap, "large"))
        if used_final:
            self.launch_refine(self.big_fallback, final + 4, final_cap, events, n_work_ptr=final,
                               label=(0, "final"))

    def download(self, want_stats=True):
        """Results through cached pinned buffers.  The arrays of the returned Result are views of
        those buffers: they stay valid until the next ``download`` of this process."""
        torch = self.torch
        plan = self.plan
        result = Result(plan, allocate=False)
        result.stats = None
        stream = torch.cuda.current_stream(self.dev)
        pieces = [("params", self.d_out, np.float64, plan.params_in.shape),
                  ("cost", self.d_cost, np.float64, (plan.n_clusters,)),
                  ("status", self.d_status, np.int32, (plan.n_clusters,))]
        if want_stats:
            pieces.append(("stats", self.d_stats, np.int32, (plan.n_clusters, 8)))
        out = {}
        for name, tensor, dtype, shape in pieces:
            nbytes = tensor.numel() * tensor.element_size()
            buf = _pinned_buffer(torch, name, nbytes)[:nbytes]
            buf.copy_(tensor.view(torch.uint8).reshape(-1), non_blocking=True)
            out[name] = buf.numpy().view(dtype).reshape(shape)
            self.d2h_bytes += nbytes
        stream.synchronize()
        result.params_out, result.cost, result.status = out["params"], out["cost"], out["status"]
        if want_stats:
            result.stats = out["stats"]
        if len(self.never_run):               # too many features for the kernel: never launched
            result.status[self.never_run] = _lib.STATUS_TOO_LARGE
            rows = np.repeat(result.status == _lib.STATUS_TOO_LARGE, self.sizes)
            result.params_out[rows] = plan.params_in[rows]
        return result


def execute_cuda(plan, device=None, want_stats=True, frames=None):
    """Run the plan on the current (or given) CUDA device through the C ABI.  ``frames``: a
    :class:`FrameSet` whose uploads were started earlier (else they are started here)."""
    import time as _time
    _t0 = _time.perf_counter()
    _ARENA.reset()
    if frames is None:
        frames = FrameSet(plan.frame_info, device).upload_async()
    session = DeviceSession(plan, frames=frames)
    torch = session.torch
    with torch.cuda.device(session.dev):
        slices = session.schedule()
        _t1 = _time.perf_counter()
        session.run(slices)
        _t2 = _time.perf_counter()
        _t3 = _t2
        result = session.download(want_stats)
    frames.close()
    session.launches += frames.launches
    session.h2d_bytes += frames.h2d_bytes
    result.session = session
    result.timing = dict(session_ms=1e3 * (_t1 - _t0), enqueue_ms=1e3 * (_t2 - _t1),
                         wait_ms=1e3 * (_t3 - _t2), download_ms=1e3 * (_time.perf_counter() - _t3))
    return result


class Pending(object):
    """A launched chunk: the kernels and the result copies are enqueued; ``result()`` waits for the
    copies (an event, not a device-wide synchronisation) and returns the :class:`Result`."""

    def __init__(self, session, result, event):
        self.session, self._result, self.event = session, result, event

    def result(self):
        if self.event is not None:
            self.event.synchronize()
            self.event = None
        res = self._result
        session = self.session
        if len(session.never_run):            # too many features for any kernel: never launched
            res.status[session.never_run] = _lib.STATUS_TOO_LARGE
            rows = np.repeat(res.status == _lib.STATUS_TOO_LARGE, session.sizes)
            res.params_out[rows] = session.plan.params_in[rows]
        return res


def launch_cuda(plan, frames, out_params, out_cost, out_status, stream=None):
    """Enqueue one plan on the current CUDA device (uploads, one launch per size class, result
    copies into the given pinned host arrays) and return a :class:`Pending` without waiting.
    ``stream``: torch stream to enqueue on (default: the current one)."""
    import contextlib
    import torch
    scope = torch.cuda.stream(stream) if stream is not None else contextlib.nullcontext()
    with scope:
        return _launch_cuda(plan, frames, out_params, out_cost, out_status)


def _launch_cuda(plan, frames, out_params, out_cost, out_status):
    session = DeviceSession(plan, frames=frames)
    torch = session.torch
    with torch.cuda.device(session.dev):
        session.run(session.schedule())
        result = Result(plan, allocate=False)
        result.stats = None
        for tensor, dst in ((session.d_out, out_params), (session.d_cost, out_cost),
                            (session.d_status, out_status)):
            host = torch.from_numpy(dst)
            host.copy_(tensor.view(host.dtype).reshape(host.shape), non_blocking=True)
            session.d2h_bytes += dst.nbytes
        event = torch.cuda.Event()
        event.record()
    result.params_out, result.cost, result.status = out_params, out_cost, out_status
    result.session = session
    return Pending(session, result, event)


def _chunk_streams(frameset, count):
    """Streams for the chunk launches; all of them are ordered behind the current stream (which
    holds the pointer-table upload) -- or [None] (current stream) when there is no device."""
    torch = getattr(frameset, 'torch', None)
    if torch is None or count < 2:
        return [None]
    current = torch.cuda.current_stream(frameset.dev)
    streams = [torch.cuda.Stream(device=frameset.dev) for _ in range(count)]
    for st in streams:
        st.wait_stream(current)
    return streams


_CHUNK_ROWS = int(os.environ.get('CTK_CHUNK_ROWS', 1 << 18))   # features per pipeline chunk
_MAX_CHUNKS = 16


def _pinned_array(key, shape, dtype):
    import torch
    nbytes = max(1, int(np.prod(shape)) * np.dtype(dtype).itemsize)
    return _pinned_buffer(torch, key, nbytes)[:nbytes].numpy().view(dtype).reshape(shape)


_CALL_LOCK = __import__('threading').Lock()
_LABELLERS = []       # the running call's labelling thread(s), closed when the call ends
_FRAMESETS = []       # ... and its frame uploads (staging thread)


def refine_leastsq(f, reader, diameter, separation=None, fit_function='gauss', param_mode=None,
                   param_val=None, constraints=None, bounds=None, pos_columns=None,
                   t_column='frame', noise_size=None, threshold=None, max_iter=10, max_shift=1,
                   max_rms_dev=1., residual_factor=100000., compute_error=False, **kwargs):
    """Refine cluster coordinates by least-squares fitting of radial model functions, on the GPU.
    See :func:`_refine_leastsq` for the parameters.  The cached pinned staging buffers are
    process-wide, so concurrent calls are serialised here; whatever happens inside, the labelling
    thread has stopped writing into them before the call returns or raises."""
    alloc = kwargs.pop('_alloc', None)       # internal: where the result columns live (parallel.py)
    with _CALL_LOCK:
        try:
            return _refine_leastsq(f, reader, diameter, separation, fit_function, param_mode,
                                   param_val, constraints, bounds, pos_columns, t_column,
                                   noise_size, threshold, max_iter, max_shift, max_rms_dev,
                                   residual_factor, compute_error, _alloc=alloc, **kwargs)
        finally:
            while _LABELLERS:
                _LABELLERS.pop().close()
            while _FRAMESETS:
                _FRAMESETS.pop().close()


def _has_global(fit_function, param_mode):
    return bool(param_mode) and any(v == 'global' or v == 2 for v in param_mode.values())


def _refine_global(f, reader, diameter, separation, fit_function, param_mode, param_val, constraints,
                   bounds, pos_columns, t_column, noise_size, threshold, max_iter, max_shift,
                   max_rms_dev, residual_factor, compute_error, passes_factory=None, reducer=None,
                   **kwargs):
    """``param_mode`` with a 'global' column: ONE problem over the whole table (refine.py:319-332),
    solved as a block-arrow system -- see :mod:`clustertracking_b200.global_fit`."""
    from . import global_fit
    if constraints:
        raise NotImplementedError("constraints on a global-level fit (constraints.dimer_global) are "
                                  "not available in the CUDA solver")
    lm_max_iter, _, xtol, _, _ = _solver_options(dict(kwargs), 'float32')
    started = []
    hook = None if passes_factory is not None else (
        lambda info: started.append(_track(FrameSet(info)).upload_async()))
    plan = prepare(f, reader, diameter, separation, fit_function, param_mode, param_val, None, bounds,
                   pos_columns, t_column, noise_size, threshold, max_iter, max_shift, max_rms_dev,
                   residual_factor, compute_error, frames_hook=hook, allow_global=True, **kwargs)
    if plan.cluster_sizes().max() > _lib.CTK_MAX_CLUSTER_FEATURES:
        raise NotImplementedError("global-level fits take clusters of up to %d features"
                                  % _lib.CTK_MAX_CLUSTER_FEATURES)
    passes = (passes_factory(plan) if passes_factory is not None
              else global_fit.CudaPasses(plan, started[0], reducer))
    if reducer is not None:
        passes.reducer = reducer
    ok, params, rms_dev = global_fit.solve(plan, passes, plan.ff, int(max_iter), float(max_shift),
                                           float(max_rms_dev), float(residual_factor), lm_max_iter, xtol)
    LAST_CALL.clear()
    LAST_CALL.update(launches=getattr(passes, 'launches', 0), chunks=1, h2d_bytes=0, d2h_bytes=0)
    return global_fit.write_back(plan, ok, params, rms_dev)


def _cuda_present():
    import torch
    return torch.cuda.is_available()


def _track(frameset):
    _FRAMESETS.append(frameset)
    return frameset


def _refine_leastsq(f, reader, diameter, separation=None, fit_function='gauss', param_mode=None,
                    param_val=None, constraints=None, bounds=None, pos_columns=None,
                    t_column='frame', noise_size=None, threshold=None, max_iter=10, max_shift=1,
                    max_rms_dev=1., residual_factor=100000., compute_error=False, _alloc=None,
                    **kwargs):
    """Refine cluster coordinates by least-squares fitting of radial model functions, on the GPU.

    Same signature, same returned columns and the same failure convention as the reference
    (clustertracking/refine.py:82-241): the result is a frame-sorted copy of ``f`` with the columns
    ``cluster``, ``cluster_size``, every model parameter (``background``, ``signal``, positions,
    ``size`` or ``size_<axis>``, ``thickness`` / ``disc_size``) and ``cost``; a cluster whose fit
    fails keeps its parameters and gets ``cost = NaN``.

    Parameters are those of the reference.  Differences:

    * ``fit_function``: 'gauss', 'ring' or 'disc' (custom dicts and 'inv_series_<n>' raise);
    * ``param_mode`` values 'const', 'var', 'cluster' ('global' raises);
    * ``constraints``: ``constraints.dimer`` / ``trimer`` / ``tetramer`` descriptors (others raise);
    * ``compute_error`` raises ``NotImplementedError``;
    * ``**kwargs``: ``options=dict(maxiter=...)`` caps the inner iterations; ``tol`` is accepted and
      ignored; ``precision='float64'`` switches the pixel arithmetic from float32 to float64.
    """
    import time
    import pandas as pd
    t0 = time.perf_counter()
    _ARENA.reset()
    if _has_global(fit_function, param_mode):
        return _refine_global(f, reader, diameter, separation, fit_function, param_mode, param_val,
                              constraints, bounds, pos_columns, t_column, noise_size, threshold,
                              max_iter, max_shift, max_rms_dev, residual_factor, compute_error,
                              **kwargs)
    started = []          # the uploads start as soon as the frames are known, before the clustering
    early_labels = []     # ... and so does the cluster labelling on the GPU (frame-sorted tables)

    # the position columns start towards the pinned staging buffer right away (threads)
    prestaged = None
    label_cols = pos_columns if pos_columns is not None else guess_pos_columns(f)
    if (len(f) >= 4096 and all(col in f for col in label_cols) and device_labelling_enabled(len(f), 4)
            and _cuda_present()):
        prestaged = Prestaged([np.ascontiguousarray(f[col].values, dtype=np.float64) for col in label_cols])
        _LABELLERS.append(prestaged)

    def frames_known(info):
        frameset = _track(FrameSet(info))
        n_rows = len(f)
        if info.run_starts is not None and device_labelling_enabled(n_rows, len(info.run_starts)):
            cols = label_cols
            sep = np.asarray(validate_tuple(diameter if separation is None else separation, len(cols)),
                             dtype=np.float64)
            first = info.run_starts.astype(np.int64)
            last = np.concatenate((first[1:], [n_rows])).astype(np.int64)
            columns = (prestaged.columns if prestaged is not None else
                       [np.ascontiguousarray(f[col].values, dtype=np.float64) for col in cols])
            labels = DeviceLabels(columns, first, last, sep, frameset.dev, prestaged=prestaged)
            _LABELLERS.append(labels)
            labels.start()         # stages + enqueues its 34 MB ahead of the gigabyte of frames
            early_labels.append(labels)
        started.append(frameset.upload_async())

    pre = prepare_common(f, reader, diameter, separation, fit_function, param_mode, param_val,
                         constraints, bounds, pos_columns, t_column, noise_size, threshold,
                         max_iter, max_shift, max_rms_dev, residual_factor, compute_error,
                         frames_hook=frames_known, **kwargs)
    ff, info = pre.ff, pre.info
    frameset = started[0]
    P = len(ff.params)
    _s1 = time.perf_counter()

    # ---- frame-sorted view of the table (find.py:122-129: the result is sorted by frame) ----------
    frames_col = f[t_column].values
    n = len(f)
    if n > 1 and info.run_starts is None:
        order0 = np.argsort(frames_col, kind='stable')
        base = f.iloc[order0]
        frames_col = frames_col[order0]
        cuts = np.flatnonzero(frames_col[1:] != frames_col[:-1]) + 1
        starts = np.concatenate(([0], cuts)).astype(np.int64)
    else:
        order0, base = None, f
        starts = info.run_starts if info.run_starts is not None else np.zeros(1, dtype=np.int64)
    # positions as table-order columns: the labelling workers read them in place
    pos = [np.ascontiguousarray(base[col].values, dtype=np.float64) for col in pre.pos_columns]
    stops = np.concatenate((starts[1:], [n])).astype(np.int64)
    n_frames = len(starts)
    # chunks of whole frames with about _CHUNK_ROWS features each
    k_chunks = int(max(1, min(_MAX_CHUNKS, n // max(1, _CHUNK_ROWS), n_frames)))
    # equal chunks, except that the last one is split in two: what is left to do after the
    # labelling has finished (launch, kernels, write-back of the final chunk) is on the critical path
    weights = np.ones(k_chunks) if k_chunks < 4 else np.array([1.] * (k_chunks - 1) + [.6, .4])
    frame_cuts = np.unique(np.searchsorted(starts, np.cumsum(weights)[:-1] / weights.sum() * n))
    frame_cuts = [0] + [int(c) for c in frame_cuts if 0 < c < n_frames] + [n_frames]
    separation = np.asarray(validate_tuple(pre.separation, pre.ndim), dtype=np.float64)

    # ---- parameter columns: from the table, from param_val, or the model's defaults ---------------
    sources = []                                   # per column: 1-D float64 array or a scalar
    for col in ff.params:
        if param_val is not None and col in param_val:
            sources.append(float(param_val[col]))
        elif col in base.columns:
            sources.append(np.ascontiguousarray(base[col].values, dtype=np.float64))
        else:
            sources.append(float(ff.default[col]))
    _s2 = time.perf_counter()
    params_in = _pinned_array("params_in", (n, P), np.float64)     # packed, group order
    # labelling, packing and the group tables run on host threads from here on
    # The result columns.  ``_alloc(name, dtype)`` lets the sharded path place them in a block all
    # ranks share, so that the final gather copies nothing (parallel.py).
    alloc = _alloc if _alloc is not None else (lambda name, dtype: np.empty(n, dtype=dtype))
    csize = alloc('cluster_size', np.int64)
    local_all = _lib.workspace("local_labels", n, np.int64)        # labels local to each frame
    labeller = ChunkLabeller(pos, starts, stops, frame_cuts, separation, sources, params_in,
                             device=getattr(frameset, "dev", None), size_out=csize, label_out=local_all,
                             device_labels=early_labels[0] if early_labels else None)
    _LABELLERS.append(labeller)
    frame_cuts = labeller.frame_cuts
    t1 = time.perf_counter()
    setup_parts = dict(prepare=1e3 * (_s1 - t0), columns=1e3 * (_s2 - _s1), labeller=1e3 * (t1 - _s2))
    out_params = _pinned_array("params", (n, P), np.float64)
    out_cost = _pinned_array("cost", (n,), np.float64)             # one entry per cluster (<= n)
    out_status = _pinned_array("status", (n,), np.int32)
    block = [alloc(col, np.float64) for col in ff.params]          # fitted columns, table order
    cost = alloc('cost', np.float64)
    cluster = alloc('cluster', np.int64)
    threads = host_threads(8)
    frame_offset = np.zeros(n_frames, dtype=np.int64)                 # find.py:127-128
    chunks = []                                    # (a, b, rows by group (chunk-local), plan, pending, c0)
    totals = dict(h2d=0, d2h=0, launches=0, failed=0)
    next_id, c0 = 0, 0
    failures = []

    def finish(chunk):
        a, b, rows_c, plan, pending, _ = chunk
        res = pending.result()
        n_failed = _lib.scatter_rows(res.params_out, plan.params_in, rows_c, plan.cluster_offset,
                                     res.cost, res.status, block, cost, threads, row_base=a)
        if n_failed:
            failed = np.flatnonzero(res.status != 0)
            failures.extend((a + int(rows_c[plan.cluster_offset[c]]), int(res.status[c]))
                            for c in failed[:max(0, 20 - totals['failed'])])
            totals['failed'] += n_failed
        session = res.session
        totals['h2d'] += session.h2d_bytes
        totals['d2h'] += session.d2h_bytes
        totals['launches'] += session.launches

    # CTK_CHUNK_STREAMS > 1 sends consecutive chunks to alternating streams.  Measured on B200: the
    # per-stream pools of torch's caching allocator make every chunk pay device allocations
    # (74 ms -> 190-240 ms per call), so one stream is the default.
    streams = _chunk_streams(frameset, int(os.environ.get('CTK_CHUNK_STREAMS', 1)))
    lap = dict(label_wait=0., index=0., launch=0., finish=0.)
    # The write-back of a chunk (wait for its results, scatter into the table columns: C code on
    # host threads, no interpreter lock) runs on its own thread, so that the main thread goes
    # straight on to the next chunk's launch and the device never waits for a scatter.
    import queue
    import threading
    done_queue, finish_errors = queue.Queue(), []
    timeline = []                                   # ms after the call began: launch of every chunk

    def finisher():
        while True:
            chunk = done_queue.get()
            if chunk is None:
                return
            if finish_errors:
                continue
            try:
                _t = time.perf_counter()
                finish(chunk)
                lap['finish'] += 1e3 * (time.perf_counter() - _t)
            except BaseException as exc:          # re-raised on the main thread
                finish_errors.append(exc)

    finish_thread = threading.Thread(target=finisher, daemon=True)
    finish_thread.start()
    try:
        for k, (fa, fb) in enumerate(zip(frame_cuts[:-1], frame_cuts[1:])):
            a, b = int(starts[fa]), int(stops[fb - 1])
            _ta = time.perf_counter()
            local, size, by_cluster, spans, g_offset, g_frame = labeller.get(k)
            _tb = time.perf_counter()
            if size is not csize:                  # (the native labeller writes both in place)
                csize[a:b] = size
                local_all[a:b] = local
            frame_offset[fa:fb] = next_id + np.concatenate(([0], np.cumsum(spans)[:-1]))
            next_id += int(np.sum(spans))
            _tc = time.perf_counter()
            plan = _plan_for(pre, by_cluster, g_offset, g_frame, params_in[a:b])
            n_c = len(g_frame)
            pending = launch_cuda(plan, frameset, out_params[a:b], out_cost[c0:c0 + n_c],
                                  out_status[c0:c0 + n_c], stream=streams[k % len(streams)])
            chunks.append((a, b, by_cluster, plan, pending, c0))
            done_queue.put(chunks[-1])
            timeline.append(round(1e3 * (time.perf_counter() - t0), 1))
            c0 += n_c
            _te = time.perf_counter()
            for name, dt in (("label_wait", _tb - _ta), ("index", _tc - _tb), ("launch", _te - _tc)):
                lap[name] += 1e3 * dt
            if finish_errors:
                break
    except BaseException:
        done_queue.put(None)
        finish_thread.join()
        raise
    done_queue.put(None)
    t2 = time.perf_counter()

    # ---- while the device works on the last chunks and the write-back thread drains: the running
    # cluster ids (find.py:127-128) and the result table, a frame-sorted copy of f with the new
    # columns (refine.py:296-305).  The fitted columns and the cost are the arrays the write-back
    # thread fills in place; the table only references them.
    try:
        _lib.apply_label_offsets(local_all, starts, stops, frame_offset, threads, cluster)
        data = {}
        fitted = {col: block[j] for j, col in enumerate(ff.params)}

        def filled(col, values):
            """A result column holding ``values`` (array or scalar): an independent copy, like f.copy()."""
            values = np.asarray(values)
            dst = alloc(col, values.dtype)
            dst[...] = values
            return dst

        for col in base.columns:
            if col in fitted:
                data[col] = fitted[col]
            elif param_val is not None and col in param_val:
                data[col] = filled(col, param_val[col])
            else:
                data[col] = filled(col, base[col].values)
        data['cluster'] = cluster
        data['cluster_size'] = csize
        if param_val is not None:
            for col in param_val:
                if col not in data:
                    data[col] = fitted[col] if col in fitted else filled(col, param_val[col])
        for col in ff.params:
            if col not in data:
                data[col] = fitted[col]
        data['cost'] = cost
        out = pd.DataFrame(data, index=base.index, copy=False)
    finally:
        t3 = time.perf_counter()
        finish_thread.join()
    if finish_errors:
        raise finish_errors[0]
    for row, status in failures:
        logger.warning("RefineException: cluster %d: %s", int(cluster[row]),
                       _lib.STATUS_NAMES.get(status, "status %d" % status))
    if totals['failed'] > 20:
        logger.warning("RefineException: ... and %d more clusters failed", totals['failed'] - 20)
    t4 = time.perf_counter()
    LAST_CALL.clear()
    dl = labeller.device_labels            # cluster labels from the GPU (ctk_label_frames): one launch
    LAST_CALL.update(h2d_bytes=totals['h2d'] + frameset.h2d_bytes + (dl.h2d_bytes if dl else 0),
                     d2h_bytes=totals['d2h'] + (dl.d2h_bytes if dl else 0),
                     launches=totals['launches'] + frameset.launches + (1 if dl else 0),
                     chunks=len(chunks),
                     labelling=dict(where='device' if dl else 'host threads',
                                    frames_relabelled_on_host=labeller.flagged_frames,
                                    **(dl.ms if dl else {})),
                     setup_ms=setup_parts, launched_at_ms=timeline,
                     returned_at_ms=round(1e3 * (time.perf_counter() - t0), 1),
                     phases_ms=dict(setup=1e3 * (t1 - t0), chunks=1e3 * (t2 - t1),
                                    table=1e3 * (t3 - t2), last_chunk=1e3 * (t4 - t3), **lap))
    return out


refine_leastsq.__doc__ = _refine_leastsq.__doc__ + """
    Concurrent calls are serialised (the pinned staging buffers are process-wide)."""

# diagnostics of the most recent refine_leastsq call (bytes copied, kernel launches, host phases)
LAST_CALL = {}
