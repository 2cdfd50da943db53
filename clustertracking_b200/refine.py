"""``refine_leastsq`` -- drop-in for clustertracking/refine.py:82-452 with the per-cluster fits on a B200.

Host side (this file): argument handling exactly as the reference (refine.py:242-305), cluster
grouping (find.py), vectorised per-feature bounds (fitfunc.py:538-551), packing of every
``(frame, cluster)`` group into flat ragged buffers, binning of the clusters by size, and the
write-back of the result columns (refine.py:408-427).  Device side (csrc/, through the C ABI of
include/ctk.h): everything inside the reference's ``for _, f_iter in iterable`` loop.

There is no CPU fallback: options the CUDA solver does not carry raise ``NotImplementedError``.
"""
import logging
import warnings

import numpy as np

from . import _lib
from . import constraints as _constraints
from .find import find_clusters
from .fitfunc import FitFunctions
from .utils import guess_pos_columns, is_isotropic, validate_tuple

logger = logging.getLogger(__name__)

# clusters are launched in bins of at most this many features (shared memory is sized per bin)
_BINS = (1, 2, 3, 4, 6, 8, 12, 16, 24, 32)
_FRAME_BATCH_BYTES = 1 << 30          # frames resident on the device per batch


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
        self.bounds_lo = None
        self.bounds_hi = None

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
        self.iters = np.zeros(plan.n_clusters, dtype=np.int32)


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
    ``precision`` ('float32' | 'float64' pixel arithmetic) and ``xtol`` (step tolerance)."""
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
    if kwargs or options:
        raise TypeError("unsupported keyword arguments: %r" % sorted(list(kwargs) + list(options)))
    if precision not in ('float32', 'float64'):
        raise ValueError("precision must be 'float32' or 'float64'")
    return lm_max_iter, (_lib.COMPUTE_F64 if precision == 'float64' else _lib.COMPUTE_F32), xtol


def prepare(f, reader, diameter, separation=None, fit_function='gauss', param_mode=None,
            param_val=None, constraints=None, bounds=None, pos_columns=None, t_column='frame',
            noise_size=None, threshold=None, max_iter=10, max_shift=1, max_rms_dev=1.,
            residual_factor=100000., compute_error=False, **kwargs):
    """Host half of ``refine_leastsq``: returns a :class:`Plan` (no GPU work)."""
    lm_max_iter, compute_dtype, xtol = _solver_options(kwargs, 'float32')
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

    f = find_clusters(f, separation, pos_columns, t_column)            # refine.py:297 (a copy)
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
    lo, hi = ff.feature_bounds(tables, params)

    # (frame, cluster) groups in the order of f.groupby(['frame', 'cluster'])     refine.py:336
    frames = f[t_column].values
    cluster = f['cluster'].values
    order = np.lexsort((cluster, frames))                              # stable: keeps row order
    frames_s, cluster_s = frames[order], cluster[order]
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
    plan.bounds_lo = np.ascontiguousarray(lo[order])
    plan.bounds_hi = np.ascontiguousarray(hi[order])
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
    prob.residual_factor, prob.xtol = float(residual_factor), xtol
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
    """Write the fitted parameters and ``cost`` into the DataFrame (refine.py:408-427)."""
    f, ff = plan.f, plan.ff
    sizes = plan.cluster_sizes()
    ok_rows = np.repeat(result.status == 0, sizes)
    values = np.ascontiguousarray(f[ff.params].values, dtype=np.float64)
    values[plan.order[ok_rows]] = result.params_out[ok_rows]
    for j, col in enumerate(ff.params):
        f[col] = values[:, j]
    cost = np.empty(len(f), dtype=np.float64)
    cost[plan.order] = np.repeat(np.where(result.status == 0, result.cost, np.nan), sizes)
    f['cost'] = cost
    failed = np.flatnonzero(result.status != 0)
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
def execute_cuda(plan, device=None):
    """Run the plan on the current (or given) CUDA device through the C ABI.  Frames are staged
    through pinned host memory in batches of about 1 GiB."""
    import torch
    lib = _lib.load()
    if not torch.cuda.is_available():
        raise RuntimeError("clustertracking_b200 needs a CUDA device (B200, sm_100a); "
                           "there is no CPU fallback")
    dev = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
    result = Result(plan)
    sizes = plan.cluster_sizes()
    prob = plan.problem
    n_frames = len(plan.frame_numbers)
    frame_bytes = int(np.prod(plan.frame_shape)) * np.dtype(plan.pixel_dtype).itemsize
    per_batch = max(1, min(n_frames, _FRAME_BATCH_BYTES // max(frame_bytes, 1)))
    shape_arr = (_lib.ctypes.c_int64 * 3)(*(list(plan.frame_shape) + [1] * (3 - len(plan.frame_shape))))
    torch_dtype = {np.dtype(np.uint8): torch.uint8, np.dtype(np.uint16): torch.uint16,
                   np.dtype(np.float32): torch.float32, np.dtype(np.float64): torch.float64,
                   np.dtype(np.int16): torch.int16, np.dtype(np.int32): torch.int32}[np.dtype(plan.pixel_dtype)]

    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream()
        sptr = _lib.ctypes.c_void_p(stream.cuda_stream)
        workspace = torch.empty(int(lib.ctk_refine_workspace_bytes()), dtype=torch.uint8, device=dev)
        # feature-level buffers live on the device for the whole call
        d_offset = torch.from_numpy(plan.cluster_offset).to(dev)
        d_params = torch.from_numpy(plan.params_in).to(dev)
        d_lo = torch.from_numpy(plan.bounds_lo).to(dev)
        d_hi = torch.from_numpy(plan.bounds_hi).to(dev)
        d_out = d_params.clone()
        d_cost = torch.full((plan.n_clusters,), float('nan'), dtype=torch.float64, device=dev)
        d_status = torch.full((plan.n_clusters,), -1, dtype=torch.int32, device=dev)
        d_iters = torch.zeros(plan.n_clusters, dtype=torch.int32, device=dev)
        first_cluster_of_frame = np.searchsorted(plan.cluster_frame, np.arange(n_frames + 1))
        staging = torch.empty((per_batch,) + plan.frame_shape, dtype=torch_dtype, pin_memory=True)
        staging_np = staging.numpy()
        for f0 in range(0, n_frames, per_batch):
            f1 = min(n_frames, f0 + per_batch)
            for k in range(f0, f1):
                staging_np[k - f0] = load_frame(plan, plan.frame_numbers[k])
            d_frames = staging[:f1 - f0].to(dev, non_blocking=True)
            ptrs = d_frames.data_ptr() + frame_bytes * np.arange(f1 - f0, dtype=np.int64)
            d_ptrs = torch.from_numpy(ptrs).to(dev)
            d_fmax = torch.empty(f1 - f0, dtype=torch.float64, device=dev)
            _lib.check(lib.ctk_frame_max(d_ptrs.data_ptr(), f1 - f0, int(np.prod(plan.frame_shape)),
                                         prob.pixel_dtype, d_fmax.data_ptr(), sptr), "ctk_frame_max")
            c0, c1 = int(first_cluster_of_frame[f0]), int(first_cluster_of_frame[f1])
            d_cframe = torch.from_numpy(plan.cluster_frame - f0).to(dev)   # batch-relative index
            host_status = result.status

            def launch(cap, ids):
                d_ids = torch.from_numpy(np.ascontiguousarray(ids)).to(dev)
                _lib.check(lib.ctk_refine_batch(
                    _lib.ctypes.byref(prob), d_ptrs.data_ptr(), shape_arr, d_fmax.data_ptr(),
                    len(ids), d_ids.data_ptr(), int(cap), d_cframe.data_ptr(), d_offset.data_ptr(),
                    d_params.data_ptr(), d_lo.data_ptr(), d_hi.data_ptr(), d_out.data_ptr(),
                    d_cost.data_ptr(), d_status.data_ptr(), d_iters.data_ptr(),
                    workspace.data_ptr(), sptr), "ctk_refine_batch")
                host_status[ids] = d_status[torch.from_numpy(ids.astype(np.int64)).to(dev)].cpu().numpy()

            run_bins(sizes, np.arange(c0, c1), host_status, launch)
            stream.synchronize()        # staging buffer is reused by the next batch
        result.params_out = d_out.cpu().numpy()
        result.cost = d_cost.cpu().numpy()
        result.status = np.where(result.status == _lib.STATUS_TOO_LARGE, result.status,
                                 d_status.cpu().numpy())
        result.iters = d_iters.cpu().numpy()
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
    plan = prepare(f, reader, diameter, separation, fit_function, param_mode, param_val,
                   constraints, bounds, pos_columns, t_column, noise_size, threshold, max_iter,
                   max_shift, max_rms_dev, residual_factor, compute_error, **kwargs)
    result = execute_cuda(plan)
    return finalize(plan, result)
