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
import warnings

import numpy as np

from . import _lib
from . import constraints as _constraints
from .find import cluster_table, find_clusters
from .fitfunc import FitFunctions
from .utils import guess_pos_columns, is_isotropic, validate_tuple

logger = logging.getLogger(__name__)

# clusters are launched in bins of at most this many features (shared memory is sized per bin)
_BINS = (1, 2, 3, 4, 6, 8, 12, 16, 24, 32)
_FRAME_BATCH_BYTES = 128 << 20        # frames per upload batch (uploads overlap the kernels)
# scaled step below which the factorised normal matrix is reused (chord iterations)
_CHORD_TOL = float(os.environ.get('CTK_CHORD_TOL', 0.02))


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
    def __init__(self, plan):
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
    ``chord_tol`` (scaled step size below which the factorised normal matrix is reused)."""
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
    if kwargs or options:
        raise TypeError("unsupported keyword arguments: %r" % sorted(list(kwargs) + list(options)))
    if precision not in ('float32', 'float64'):
        raise ValueError("precision must be 'float32' or 'float64'")
    return (lm_max_iter, (_lib.COMPUTE_F64 if precision == 'float64' else _lib.COMPUTE_F32), xtol,
            chord_tol)


def prepare(f, reader, diameter, separation=None, fit_function='gauss', param_mode=None,
            param_val=None, constraints=None, bounds=None, pos_columns=None, t_column='frame',
            noise_size=None, threshold=None, max_iter=10, max_shift=1, max_rms_dev=1.,
            residual_factor=100000., compute_error=False, **kwargs):
    """Host half of ``refine_leastsq``: returns a :class:`Plan` (no GPU work)."""
    lm_max_iter, compute_dtype, xtol, chord_tol = _solver_options(kwargs, 'float32')
    if pos_columns is None:
        pos_columns = guess_pos_columns(f)
    if compute_error:
        raise NotImplementedError("compute_error (numdifftools Hessian, refine.py:400-406) is not "
                                  "available in the CUDA solver")
    if noise_size is not None:
        raise NotImplementedError("noise_size / threshold (lowpass on the sub-image, "
                                  "refine.py:36-40) is not available in the CUDA solver yet")
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
    if any(m == 2 for m in ff.modes):
        raise NotImplementedError("param_mode 'global' couples all clusters into one problem "
                                  "(refine.py:319-332) and is out of scope of the CUDA solver")
    if any(m > 3 for m in ff.modes):
        raise NotImplementedError("param modes 'particle' and 'frame' are not implemented")
    if len(ff.params) > _lib.CTK_MAX_PARAMS:
        raise NotImplementedError("too many parameters per feature")
    if max(radius) > _lib.CTK_MAX_RADIUS or min(radius) < 1:
        raise NotImplementedError("mask radius (diameter // 2) must be within [1, %d]"
                                  % _lib.CTK_MAX_RADIUS)
    cons = _constraints.parse(constraints, ndim)

    f, order = cluster_table(f, separation, pos_columns, t_column)     # refine.py:297 (a copy)
    if param_val is not None:                                          # refine.py:300-302
        for col in param_val:
            f[col] = param_val[col]
    for col in ff.params:                                              # refine.py:303-305
        if col not in f.columns:
            f[col] = ff.default[col]
    tables = ff.validate_bounds(bounds, radius=radius)                 # refine.py:315

    plan = Plan()
    plan.f, plan.ff = f, ff
    # column order of the parameter table follows ff.params, but positions are read from the
    # user's pos_columns (refine.py:345 reads ff.params; they coincide for the default names)
    params = np.ascontiguousarray(f[ff.params].values, dtype=np.float64)

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
    plan.order = order
    plan.cluster_offset = np.concatenate((starts, [n])).astype(np.int32)
    frame_numbers, frame_index = np.unique(frames_s[starts], return_inverse=True)
    plan.frame_numbers = [int(x) if float(x).is_integer() else x for x in frame_numbers]
    plan.cluster_frame = frame_index.astype(np.int32)
    plan.params_in = np.ascontiguousarray(params[order])
    plan.frame_source = source

    first = np.asarray(source[plan.frame_numbers[0]])
    plan.frame_shape = tuple(first.shape)
    if len(plan.frame_shape) != ndim:
        raise ValueError("frames must have %d dimensions" % ndim)
    plan.pixel_dtype = first.dtype if first.dtype in _lib.PIXEL_CODES else np.dtype(np.float64)

    prob = _lib.Problem()
    prob.ndim, prob.isotropic, prob.family = ndim, int(isotropic), ff.family
    prob.n_params = len(ff.params)
    for j, m in enumerate(ff.modes):
        prob.modes[j] = m
    for k, r in enumerate(radius):
        prob.radius[k] = r
    prob.pixel_dtype = _lib.PIXEL_CODES[np.dtype(plan.pixel_dtype)]
    prob.compute_dtype = compute_dtype
    prob.max_iter, prob.lm_max_iter = int(max_iter), lm_max_iter
    prob.max_shift, prob.max_rms_dev = float(max_shift), float(max_rms_dev)
    prob.residual_factor, prob.xtol, prob.chord_tol = float(residual_factor), xtol, chord_tol
    mask = 0
    if cons['dimer'] is not None:
        mask |= _lib.CONSTRAINT_DIMER
        for k in range(ndim):
            prob.dimer_dist[k] = cons['dimer'][k]
    if cons['trimer'] is not None:
        mask |= _lib.CONSTRAINT_TRIMER
        for k in range(ndim):
            prob.trimer_dist[k] = cons['trimer'][k]
    prob.constraint_mask = mask
    for which, table in zip(("bounds_abs", "bounds_diff", "bounds_rel"), tables):
        dst = getattr(prob, which)
        for side in range(2):
            for j in range(_lib.CTK_MAX_PARAMS):
                dst[side][j] = table[side, j] if j < len(ff.params) else np.nan
    plan.problem = prob
    return plan


def load_frame(plan, frame_no):
    """One frame as a C-contiguous array of the plan's pixel type."""
    image = np.asarray(plan.frame_source[frame_no])
    if image.shape != plan.frame_shape:
        raise ValueError("frame %r has shape %r, expected %r" % (frame_no, image.shape,
                                                                  plan.frame_shape))
    return np.ascontiguousarray(image, dtype=plan.pixel_dtype)


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
    """Launch every bin; clusters that overflowed their bin's capacity (status TOO_LARGE) are retried
    once in the next larger bin.  ``launch(capacity, ids)`` must fill ``status[ids]``."""
    too_big = cluster_ids[sizes[cluster_ids] > _BINS[-1]]
    status[too_big] = _lib.STATUS_TOO_LARGE
    for cap, ids in bin_clusters(sizes, cluster_ids):
        launch(cap, ids)
        retry = ids[status[ids] == _lib.STATUS_TOO_LARGE]
        bigger = [c for c in _BINS if c > cap]
        if len(retry) and bigger:
            launch(bigger[min(1, len(bigger) - 1)], retry)


def finalize(plan, result):
    """Write the fitted parameters and ``cost`` into the DataFrame (refine.py:408-427).  The device
    already copies the input parameters through for failed clusters, so the whole table is written
    with one scatter per column."""
    f, ff = plan.f, plan.ff
    sizes = plan.cluster_sizes()
    ok = result.status == 0
    block = np.empty((len(ff.params), len(f)), dtype=np.float64)
    block[:, plan.order] = result.params_out.T
    if not ok.all():                       # belt and braces: failed clusters keep their input exactly
        rows = np.repeat(~ok, sizes)
        block[:, plan.order[rows]] = plan.params_in[rows].T
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
class DeviceSession(object):
    """Device-side state of one plan: feature-level buffers stay resident for the whole call, frames
    are uploaded in batches through pinned host memory.  torch is used for device memory and the
    stream only; all arithmetic happens inside libctk (C ABI, include/ctk.h)."""

    def __init__(self, plan, device=None):
        import torch
        self.torch = torch
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise RuntimeError("clustertracking_b200 needs a CUDA device (B200, sm_100a); "
                               "there is no CPU fallback")
        self.dev = (torch.device('cuda', torch.cuda.current_device()) if device is None
                    else torch.device(device))
        self.plan = plan
        self.sizes = plan.cluster_sizes()
        self.n_frames = len(plan.frame_numbers)
        self.n_pixels = int(np.prod(plan.frame_shape))
        self.frame_bytes = self.n_pixels * np.dtype(plan.pixel_dtype).itemsize
        self.shape_arr = (_lib.ctypes.c_int64 * 3)(
            *(list(plan.frame_shape) + [1] * (3 - len(plan.frame_shape))))
        self.torch_dtype = {
            np.dtype(np.uint8): torch.uint8, np.dtype(np.uint16): torch.uint16,
            np.dtype(np.float32): torch.float32, np.dtype(np.float64): torch.float64,
            np.dtype(np.int16): torch.int16, np.dtype(np.int32): torch.int32,
        }[np.dtype(plan.pixel_dtype)]
        self.first_cluster_of_frame = np.searchsorted(plan.cluster_frame,
                                                      np.arange(self.n_frames + 1))
        self.launches = 0
        self.h2d_bytes = 0
        self.d2h_bytes = 0
        self.host_status = np.full(plan.n_clusters, -1, dtype=np.int32)
        with torch.cuda.device(self.dev):
            dev = self.dev
            self.workspace = torch.empty(int(self.lib.ctk_refine_workspace_bytes()),
                                         dtype=torch.uint8, device=dev)
            self.d_offset = self._up(plan.cluster_offset)
            self.d_params = self._up(plan.params_in)
            self.d_lo = self._up(plan.bounds_lo) if plan.bounds_lo is not None else None
            self.d_hi = self._up(plan.bounds_hi) if plan.bounds_hi is not None else None
            self.d_out = torch.empty_like(self.d_params)
            self.d_cost = torch.empty(plan.n_clusters, dtype=torch.float64, device=dev)
            self.d_status = torch.full((plan.n_clusters,), -1, dtype=torch.int32, device=dev)
            self.d_stats = torch.zeros((plan.n_clusters, 8), dtype=torch.int32, device=dev)

    def _up(self, array):
        t = self.torch.from_numpy(np.ascontiguousarray(array))
        self.h2d_bytes += t.numel() * t.element_size()
        return t.to(self.dev, non_blocking=False)

    def stream_ptr(self):
        return _lib.ctypes.c_void_p(self.torch.cuda.current_stream(self.dev).cuda_stream)

    def attach_frames(self, d_frames, f0):
        """Describe a batch of frames already resident on the device: ``d_frames`` is a torch
        tensor [n, *frame_shape] holding frames f0 .. f0 + n - 1 of the plan."""
        torch = self.torch
        n = int(d_frames.shape[0])
        ptrs = d_frames.data_ptr() + self.frame_bytes * np.arange(n, dtype=np.int64)
        batch = dict(frames=d_frames, f0=f0, n=n, d_ptrs=self._up(ptrs),
                     d_fmax=torch.empty(n, dtype=torch.float64, device=self.dev),
                     d_cframe=self._up(self.plan.cluster_frame - f0),
                     c0=int(self.first_cluster_of_frame[f0]),
                     c1=int(self.first_cluster_of_frame[f0 + n]))
        ids = np.arange(batch['c0'], batch['c1'])
        batch['bins'] = [(cap, sel, self._up(sel)) for cap, sel in bin_clusters(self.sizes, ids)]
        return batch

    def _direct_view(self, f0, f1):
        """Zero-copy host view of frames f0..f1-1 when the reader is backed by one contiguous array
        of the right type (``reader.stack``); pinned memory then goes to the device by DMA."""
        src = self.plan.frame_source
        stack = getattr(src, 'stack', None)
        first = getattr(src, 'first_frame', 0)
        numbers = self.plan.frame_numbers[f0:f1]
        if not isinstance(stack, np.ndarray) or stack.dtype != self.plan.pixel_dtype:
            return None
        if stack.shape[1:] != tuple(self.plan.frame_shape) or not stack.flags.c_contiguous:
            return None
        idx = np.asarray(numbers, dtype=np.int64) - first
        if idx[0] < 0 or idx[-1] >= len(stack) or np.any(np.diff(idx) != 1):
            return None
        return stack[idx[0]:idx[-1] + 1]

    def upload_frames(self, f0, f1, staging=None):
        """Copy frames f0..f1-1 of the plan's source to the device: straight from the reader's
        array when it has one, else frame by frame through a pinned staging buffer."""
        torch = self.torch
        n = f1 - f0
        view = self._direct_view(f0, f1)
        if view is not None:
            d_frames = torch.from_numpy(view).to(self.dev, non_blocking=True)
            self.h2d_bytes += n * self.frame_bytes
            return self.attach_frames(d_frames, f0), staging
        if staging is None or staging.shape[0] < n:
            staging = torch.empty((n,) + tuple(self.plan.frame_shape), dtype=self.torch_dtype,
                                  pin_memory=True)
        view = staging.numpy()
        for k in range(f0, f1):
            view[k - f0] = load_frame(self.plan, self.plan.frame_numbers[k])
        d_frames = staging[:n].to(self.dev, non_blocking=True)
        self.h2d_bytes += n * self.frame_bytes
        return self.attach_frames(d_frames, f0), staging

    def _launch(self, batch, cap, ids, d_ids, events=None):
        prob = self.plan.problem
        if events is not None:
            start, stop = (self.torch.cuda.Event(enable_timing=True) for _ in range(2))
            start.record()
        _lib.check(self.lib.ctk_refine_batch(
            _lib.ctypes.byref(prob), batch['d_ptrs'].data_ptr(), self.shape_arr,
            batch['d_fmax'].data_ptr(), len(ids), d_ids.data_ptr(), int(cap),
            batch['d_cframe'].data_ptr(), self.d_offset.data_ptr(), self.d_params.data_ptr(),
            self.d_lo.data_ptr() if self.d_lo is not None else None,
            self.d_hi.data_ptr() if self.d_hi is not None else None, self.d_out.data_ptr(),
            self.d_cost.data_ptr(), self.d_status.data_ptr(), self.d_stats.data_ptr(),
            self.workspace.data_ptr(), self.stream_ptr()), "ctk_refine_batch")
        self.launches += 1
        if events is not None:
            stop.record()
            events.append(("refine", start, stop))

    def run_batch(self, batch, retry=True, events=None):
        """Frame maxima, then one refine launch per size bin; clusters whose pixel lists overflowed
        their bin (status TOO_LARGE) are relaunched once with a larger capacity.  ``events``: list
        that receives (kind, start, stop) CUDA-event triples of every launch (bench.py)."""
        prob = self.plan.problem
        if events is not None:
            start, stop = (self.torch.cuda.Event(enable_timing=True) for _ in range(2))
            start.record()
        _lib.check(self.lib.ctk_frame_max(batch['d_ptrs'].data_ptr(), batch['n'], self.n_pixels,
                                          prob.pixel_dtype, batch['d_fmax'].data_ptr(),
                                          self.stream_ptr()), "ctk_frame_max")
        self.launches += 1
        if events is not None:
            stop.record()
            events.append(("frame_max", start, stop))
        for cap, ids, d_ids in batch['bins']:
            self._launch(batch, cap, ids, d_ids, events)
        ids = np.arange(batch['c0'], batch['c1'])
        self.host_status[ids[self.sizes[ids] > _BINS[-1]]] = _lib.STATUS_TOO_LARGE
        if not retry:
            return
        # one small device->host read decides whether anything has to be relaunched
        status = self.d_status[batch['c0']:batch['c1']].cpu().numpy()
        self.d2h_bytes += status.nbytes
        over = ids[(status == _lib.STATUS_TOO_LARGE) & (self.sizes[ids] <= _BINS[-1])]
        for cap in _BINS:
            sel = over[(self.sizes[over] <= cap)]
            over = over[self.sizes[over] > cap]
            bigger = [c for c in _BINS if c > cap]
            if len(sel) and bigger:
                sel = np.ascontiguousarray(sel, dtype=np.int32)
                self._launch(batch, bigger[min(1, len(bigger) - 1)], sel, self._up(sel))

    def download(self):
        result = Result(self.plan)
        self.torch.cuda.current_stream(self.dev).synchronize()
        result.params_out = self.d_out.cpu().numpy()
        result.cost = self.d_cost.cpu().numpy()
        status = self.d_status.cpu().numpy()
        result.status = np.where(self.host_status == _lib.STATUS_TOO_LARGE, self.host_status, status)
        result.stats = self.d_stats.cpu().numpy()
        self.d2h_bytes += (result.params_out.nbytes + result.cost.nbytes + status.nbytes
                           + result.stats.nbytes)
        # clusters that never ran (too many features) keep their input parameters
        never = np.repeat(result.status == _lib.STATUS_TOO_LARGE, self.sizes)
        result.params_out[never] = self.plan.params_in[never]
        return result


def execute_cuda(plan, device=None):
    """Run the plan on the current (or given) CUDA device through the C ABI.

    Frames go to the device in batches on a separate copy stream, so the upload of batch k+1
    overlaps the kernels of batch k: straight from the reader's array when it has one (a DMA when
    that memory is pinned), else frame by frame through a pinned staging buffer."""
    session = DeviceSession(plan, device)
    torch = session.torch
    per_batch = max(1, min(session.n_frames, _FRAME_BATCH_BYTES // max(session.frame_bytes, 1)))
    with torch.cuda.device(session.dev):
        compute = torch.cuda.current_stream(session.dev)
        direct = session._direct_view(0, session.n_frames) is not None
        if direct:
            copy_stream = torch.cuda.Stream(device=session.dev)
            pending = []
            for f0 in range(0, session.n_frames, per_batch):
                f1 = min(session.n_frames, f0 + per_batch)
                with torch.cuda.stream(copy_stream):
                    view = session._direct_view(f0, f1)
                    d_frames = torch.from_numpy(view).to(session.dev, non_blocking=True)
                    done = torch.cuda.Event()
                    done.record(copy_stream)
                session.h2d_bytes += (f1 - f0) * session.frame_bytes
                pending.append((f0, d_frames, done))
            for f0, d_frames, done in pending:
                compute.wait_event(done)
                d_frames.record_stream(compute)
                session.run_batch(session.attach_frames(d_frames, f0))
        else:
            staging = None
            for f0 in range(0, session.n_frames, per_batch):
                f1 = min(session.n_frames, f0 + per_batch)
                batch, staging = session.upload_frames(f0, f1, staging)
                session.run_batch(batch)
                compute.synchronize()                    # staging is reused by the next batch
        result = session.download()
    result.session = session
    return result


def refine_leastsq(f, reader, diameter, separation=None, fit_function='gauss', param_mode=None,
                   param_val=None, constraints=None, bounds=None, pos_columns=None,
                   t_column='frame', noise_size=None, threshold=None, max_iter=10, max_shift=1,
                   max_rms_dev=1., residual_factor=100000., compute_error=False, **kwargs):
    """Refine cluster coordinates by least-squares fitting of radial model functions, on the GPU.

    Same signature, same returned columns and the same failure convention as the reference
    (clustertracking/refine.py:82-241): the result is a frame-sorted copy of ``f`` with the columns
    ``cluster``, ``cluster_size``, every model parameter (``background``, ``signal``, positions,
    ``size`` or ``size_<axis>``, ``thickness`` / ``disc_size``) and ``cost``; a cluster whose fit
    fails keeps its parameters and gets ``cost = NaN``.

    Parameters are those of the reference.  Differences:

    * ``fit_function``: 'gauss', 'ring' or 'disc' (custom dicts and 'inv_series_<n>' raise);
    * ``param_mode`` values 'const', 'var', 'cluster' ('global' raises);
    * ``constraints``: ``constraints.dimer`` / ``constraints.trimer`` descriptors (others raise);
    * ``noise_size`` and ``compute_error`` raise ``NotImplementedError``;
    * ``**kwargs``: ``options=dict(maxiter=...)`` caps the inner iterations; ``tol`` is accepted and
      ignored; ``precision='float64'`` switches the pixel arithmetic from float32 to float64.
    """
    import time
    t0 = time.perf_counter()
    plan = prepare(f, reader, diameter, separation, fit_function, param_mode, param_val,
                   constraints, bounds, pos_columns, t_column, noise_size, threshold, max_iter,
                   max_shift, max_rms_dev, residual_factor, compute_error, **kwargs)
    t1 = time.perf_counter()
    result = execute_cuda(plan)
    t2 = time.perf_counter()
    out = finalize(plan, result)
    t3 = time.perf_counter()
    LAST_CALL.clear()
    LAST_CALL.update(h2d_bytes=result.session.h2d_bytes, d2h_bytes=result.session.d2h_bytes,
                     launches=result.session.launches,
                     phases_ms=dict(prepare=1e3 * (t1 - t0), device=1e3 * (t2 - t1),
                                    finalize=1e3 * (t3 - t2)))
    return out


# diagnostics of the most recent refine_leastsq call (bytes copied, kernel launches, host phases)
LAST_CALL = {}
