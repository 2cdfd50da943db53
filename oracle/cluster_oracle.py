"""CPU oracle for the ``refine_leastsq`` hot path of caspervdw/clustertracking.

TEST INFRASTRUCTURE -- NOT PART OF THE PRODUCT.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this module, and only as the
checker / the CPU baseline.  ``clustertracking_b200`` never imports it and has no CPU fallback.

It is a numpy restatement of the reference's cluster-level algorithm.  Every function names the
reference lines it follows (paths relative to ``/root/reference/clustertracking``).  The optimiser is
the third-party ``scipy.optimize.minimize(method='SLSQP')`` (reference call site ``refine.py:373-375``,
defaults ``refine.py:242-243``; scipy is unpinned in the reference's ``setup.py:22``; this image has
scipy 1.18.1) and is called here exactly as the reference calls it.

PARITY PINNED: ``tests/test_oracle_golden.py`` checks this module against outputs of the unmodified
reference run in the build container (``oracle/make_golden.py`` -> ``tests/golden/*.npz``), and against
the known-answer tests the reference holds for the path (``tests/test_fitfunc.py:65-83``,
``tests/test_mask.py:11-129``).

Deliberate divergences from reference quirks, all documented in DESIGN.md:
  * C1  each constraint is applied to clusters of its own ``cluster_size`` (the reference's closures
        bind the loop variable late, ``constraints.py:25-38``); compare with one kind per call.
  * C2  a failure without message gives ``cost = NaN`` instead of crashing the handler (``refine.py:417``).
  * global-mode fits, ``compute_error`` and custom dict fit functions are out of scope -> raise.
"""
import logging
import warnings

import numpy as np
import pandas as pd
from scipy.optimize import minimize
from scipy.spatial import cKDTree

logger = logging.getLogger(__name__)

CONST, VAR, GLOBAL, CLUSTER = 0, 1, 2, 3
_MODE_CODES = {0: 0, 1: 1, 2: 2, 3: 3, 'const': CONST, 'var': VAR, 'global': GLOBAL,
               'cluster': CLUSTER}          # fitfunc.py:9-11 (modes 4-6 are unimplemented upstream)


class FitFailure(Exception):
    """Per-cluster failure signal (utils.py:96-97)."""


def as_ndim_tuple(value, ndim):
    """trackpy.utils.validate_tuple (third party, absent; behaviour per SURVEY.md App. B.1)."""
    if not hasattr(value, '__iter__'):
        return (value,) * ndim
    if len(value) == ndim:
        return tuple(value)
    raise ValueError("List length should have same length as image dimensions.")


def position_columns(ndim):            # utils.py:40-41
    return ['z', 'y', 'x'][-ndim:]


def size_columns(ndim, isotropic):     # utils.py:44-49
    return ['size'] if isotropic else ['size_z', 'size_y', 'size_x'][-ndim:]


def all_equal(value):                  # utils.py:52-56
    if hasattr(value, '__iter__'):
        return bool(np.all(np.asarray(value[1:]) == np.asarray(value[:-1])))
    return True


# ----------------------------------------------------------------------------------------------
# cluster grouping                                                         find.py:12-163
# ----------------------------------------------------------------------------------------------
def label_clusters(pos, separation):
    """Cluster ids and sizes for the points of ONE frame (find.py:72-93 with Clusters, 12-60).

    Pairs closer than 1 in ``pos / separation`` units are merged; when the clusters of ``a`` and
    ``b`` merge, the label of ``a``'s cluster survives (find.py:41-48).  The pairs are visited in the
    iteration order of the python ``set`` that ``cKDTree.query_pairs`` returns, exactly as upstream,
    so the label VALUES (an artefact of that order) are reproduced as well as the membership.
    """
    pos = np.asarray(pos, dtype=np.float64)
    n = len(pos)
    label = list(range(n))
    members = {i: [i] for i in range(n)}
    for a, b in cKDTree(pos / separation).query_pairs(1):
        keep, drop = label[a], label[b]
        if keep == drop:
            continue
        moved = members.pop(drop)
        for k in moved:
            label[k] = keep
        members[keep].extend(moved)
    size = [0] * n
    for group in members.values():
        for k in group:
            size[k] = len(group)
    return label, size


def find_clusters(f, separation, pos_columns=None, t_column='frame'):
    """find.py:96-163: per frame labels + running offset; returns a frame-sorted COPY."""
    if pos_columns is None:
        pos_columns = ['z', 'y', 'x'] if 'z' in f else ['y', 'x']      # utils.py:24-29
    added = t_column not in f
    if added:
        f[t_column] = 0                                                 # find.py:151-153
    pieces, next_id = [], 0
    for _, part in f.groupby(t_column):                                 # find.py:121
        ids, sizes = label_clusters(part[pos_columns].values, separation)
        part = part.copy()
        part['cluster'] = ids
        part['cluster_size'] = sizes
        part['cluster'] += next_id                                      # find.py:127
        next_id = part['cluster'].max() + 1
        pieces.append(part)
    out = pd.concat(pieces)
    if added:
        del f[t_column]
    return out


# ----------------------------------------------------------------------------------------------
# pixel set of a cluster                                      masks.py:30-68, refine.py:28-58
# ----------------------------------------------------------------------------------------------
def bounding_box(coords, shape, radius):
    """masks.py:49-61.  Returns (lower, upper) integer lists, or None when no coordinate is in
    bounds.  Round-half-even, keep rows with -r <= c < shape + r, box [min-r, max+r+1) clipped."""
    ndim = len(shape)
    radius = as_ndim_tuple(radius, ndim)
    ci = np.atleast_2d(np.round(coords).astype(int))
    keep = np.ones(len(ci), dtype=bool)
    for k in range(ndim):                                               # masks.py:42-46
        keep &= (ci[:, k] >= -radius[k]) & (ci[:, k] < shape[k] + radius[k])
    ci = ci[keep]
    if len(ci) == 0:
        return None
    lower = [max(0, int(ci[:, k].min()) - radius[k]) for k in range(ndim)]          # masks.py:30-39
    upper = [min(shape[k], int(ci[:, k].max()) + radius[k] + 1) for k in range(ndim)]
    return lower, upper


def gaussian_kernel(sigma, truncate=4.0):
    """trackpy.masks.gaussian_kernel (third party, absent)."""
    lw = int(truncate * sigma + 0.5)
    x = np.arange(-lw, lw + 1)
    k = np.exp(x ** 2 / (-2 * sigma ** 2))
    return k / np.sum(k)


def lowpass(image, lshort, threshold=None):
    """preprocessing.py:12-49: separable gaussian, zero padded at the array edge, then cut."""
    from scipy.ndimage import correlate1d
    lshort = as_ndim_tuple(lshort, image.ndim)
    if threshold is None:
        threshold = 1 if np.issubdtype(image.dtype, np.integer) else 1 / 256.
    out = np.array(image, dtype=np.float64)
    for axis, sigma in enumerate(lshort):
        if sigma > 0:
            correlate1d(out, gaussian_kernel(sigma, 4), axis, output=out, mode='constant', cval=0.0)
    return np.where(out > threshold, out, 0)


def cluster_pixels(coords, image, radius, noise_size=None, threshold=None):
    """refine.py:28-58.  -> (values[M] f64, mesh[ndim, M] f64 absolute coords, masks[n, M] bool).

    Feature i covers box pixel idx iff sum_k ((idx_k - (c_ik - origin_k)) / r_k)**2 <= 1, evaluated
    in float64 with the axis order z, y, x; the pixel set is the union, in C order over the box."""
    ndim = image.ndim
    radius = as_ndim_tuple(radius, ndim)
    box = bounding_box(coords, image.shape, radius)
    if box is None:
        raise FitFailure("cluster is outside of the image")            # refine.py:33-34
    lower, upper = box
    sub = image[tuple(slice(a, b) for a, b in zip(lower, upper))]
    if noise_size is not None:                                          # refine.py:37-40
        sub = lowpass(sub, noise_size, 0 if threshold is None else threshold)
    grid_t = np.indices(sub.shape).T                                    # refine.py:43
    inside = [np.sum(((grid_t - (c - lower)) / radius) ** 2, -1) <= 1 for c in coords]
    union = np.any(inside, axis=0).T                                    # refine.py:47
    masks = np.empty((len(coords), int(union.sum())), dtype=bool)
    for i, one in enumerate(inside):
        masks[i] = one.T[union]
    mesh = np.indices(sub.shape, dtype=np.float64)[:, union]            # refine.py:54
    mesh += np.array(lower)[:, np.newaxis]
    return sub[union].astype(np.float64), mesh, masks


# ----------------------------------------------------------------------------------------------
# radial models                                                        fitfunc.py:14-204
# ----------------------------------------------------------------------------------------------
def _reduced_r2(mesh, p, ndim, isotropic, safe):
    """fitfunc.py:14-109 (the r2_* family).  ``safe``: NaN where the PIXEL distance**2 < 1."""
    c = p[2:2 + ndim]
    if ndim == 2:
        y, x = mesh
        cy, cx = c
        if isotropic:
            size = p[4]
            if not safe:
                return ((x - cx) ** 2 + (y - cy) ** 2) / size ** 2               # :14-17
            dist = (x - cx) ** 2 + (y - cy) ** 2                                 # :20-26
            dist[dist < 1.] = np.nan
            dist /= size ** 2
            return dist
        size_y, size_x = p[4:6]
        out = (x - cx) ** 2 / size_x ** 2 + (y - cy) ** 2 / size_y ** 2          # :62-65
        if safe:
            out[(x - cx) ** 2 + (y - cy) ** 2 < 1.] = np.nan                     # :68-74
        return out
    z, y, x = mesh
    cz, cy, cx = c
    if isotropic:
        size = p[5]
        if not safe:
            return ((x - cx) ** 2 + (y - cy) ** 2 + (z - cz) ** 2) / size ** 2   # :37-40
        dist = (x - cx) ** 2 + (y - cy) ** 2 + (z - cz) ** 2                     # :43-49
        dist[dist < 1.] = np.nan
        dist /= size ** 2
        return dist
    size_z, size_y, size_x = p[5:8]
    out = (x - cx) ** 2 / size_x ** 2 + (y - cy) ** 2 / size_y ** 2 + \
          (z - cz) ** 2 / size_z ** 2                                            # :87-90
    if safe:
        out[(x - cx) ** 2 + (y - cy) ** 2 + (z - cz) ** 2 < 1.] = np.nan         # :93-101
    return out


def _reduced_r2_grad(mesh, p, ndim, isotropic):
    """fitfunc.py:29-34, 52-59, 77-84, 104-109 (the dr2_* family): rows = centres then size(s)."""
    c = p[2:2 + ndim]
    if isotropic:
        size = p[2 + ndim]
        rows = [(c[k] - mesh[k]) * (2. / size ** 2) for k in range(ndim)]
        if ndim == 2:
            y, x = mesh
            rows.append(((x - c[1]) ** 2 + (y - c[0]) ** 2) * (-2. / size ** 3))
        else:
            z, y, x = mesh
            rows.append(((x - c[2]) ** 2 + (y - c[1]) ** 2 + (z - c[0]) ** 2) * (-2. / size ** 3))
        return np.vstack(rows)
    sizes = p[2 + ndim:2 + 2 * ndim]
    rows = [(c[k] - mesh[k]) * (2. / sizes[k] ** 2) for k in range(ndim)]
    rows += [(mesh[k] - c[k]) ** 2 * (-2. / sizes[k] ** 3) for k in range(ndim)]
    return np.vstack(rows)


def gauss_value(r2, extra, ndim):                                       # fitfunc.py:112-113
    return np.exp(-0.5 * ndim * r2)


def gauss_value_grad(r2, extra, ndim):                                  # fitfunc.py:116-118
    g = np.exp(-0.5 * ndim * r2)
    return g, [-0.5 * ndim * g]


def disc_value(r2, extra, ndim):                                        # fitfunc.py:121-131
    out = np.ones_like(r2)
    d = extra[0]
    if d <= 0:
        return gauss_value(r2, None, ndim)
    elif d >= 1.:
        d = 0.999
    outer = r2 > d ** 2
    out[outer] = np.exp(((r2[outer] ** 0.5 - d) / (1 - d)) ** 2 * ndim / -2)
    return out


def ring_value(r2, extra, ndim):                                        # fitfunc.py:134-137
    t = extra[0]
    r = r2 ** 0.5
    return np.exp(-0.5 * ndim * ((r - 1 + t) / t) ** 2)


def ring_value_grad(r2, extra, ndim):                                   # fitfunc.py:140-146
    t = extra[0]
    r = r2 ** 0.5
    num = r - 1 + t
    g = np.exp(-0.5 * ndim * (num / t) ** 2)
    return g, [g * (-0.5 * ndim / (r * t ** 2)) * num,
               g * ndim * (num ** 2 / t ** 3 - num / t ** 2)]


_FAMILIES = {                                                           # fitfunc.py:195-204
    'gauss': dict(extra=[], value=gauss_value, value_grad=gauss_value_grad, continuous=True,
                  default={}),
    'ring': dict(extra=['thickness'], value=ring_value, value_grad=ring_value_grad,
                 continuous=False, default=dict(thickness=0.5)),
    'disc': dict(extra=['disc_size'], value=disc_value, value_grad=None, continuous=False,
                 default=dict(disc_size=0.5)),
}


# ----------------------------------------------------------------------------------------------
# vector <-> parameter table (cluster level: groups is None)          fitfunc.py:207-315
# ----------------------------------------------------------------------------------------------
def pack_vector(params, modes, reduce_op=None):
    """fitfunc.py:207-263 with ``groups=None``: const skipped, var -> n entries, else one entry."""
    chunks = []
    for j, mode in enumerate(modes):
        if mode == CONST:
            continue
        if mode == VAR:
            chunks.append(params[:, j])
        elif reduce_op is None:
            chunks.append([params[0, j]])
        else:
            chunks.append([reduce_op(params[:, j])])
    if not chunks:
        return np.empty((0,))
    return np.concatenate(chunks)


def unpack_vector(vect, params, modes):
    """fitfunc.py:266-315 with ``groups=None``."""
    n = params.shape[0]
    out = params.copy()
    at = 0
    for j, mode in enumerate(modes):
        if mode == CONST:
            continue
        if mode == VAR:
            out[:, j] = vect[at:at + n]
            at += n
        else:
            out[:, j] = vect[at]
            at += 1
    return out


class ModelSpec(object):
    """Parameter list, modes, objective and bounds (fitfunc.py:318-558, class FitFunctions)."""

    def __init__(self, fit_function='gauss', ndim=2, isotropic=True, param_mode=None):
        if isinstance(fit_function, dict) or fit_function not in _FAMILIES:
            raise NotImplementedError("oracle covers the gauss / ring / disc families only")
        fam = _FAMILIES[fit_function]
        self.family = fit_function
        self.ndim, self.isotropic = ndim, isotropic
        self.pos_columns = position_columns(ndim)
        self.size_columns = size_columns(ndim, isotropic)
        self.extra = list(fam['extra'])
        self.value, self.value_grad = fam['value'], fam['value_grad']
        self.safe = not fam['continuous']                               # fitfunc.py:396-411
        self.default = dict(background=0., **fam['default'])            # fitfunc.py:349
        self.params = ['background', 'signal'] + self.pos_columns + self.size_columns + self.extra

        mode = dict(signal='var', background='cluster')                 # fitfunc.py:356-360
        if param_mode is not None:
            mode.update(param_mode)
        if 'pos' in mode:                                               # fitfunc.py:362-367
            for col in self.pos_columns:
                mode.setdefault(col, mode['pos'])
            del mode['pos']
        if (not isotropic) and ('size' in mode):                        # fitfunc.py:368-373
            for col in self.size_columns:
                mode.setdefault(col, mode['size'])
            del mode['size']
        mode = {k: _MODE_CODES[v] for k, v in mode.items()}             # fitfunc.py:375-377
        for col in self.pos_columns:
            mode.setdefault(col, VAR)                                   # fitfunc.py:379-382
        for col in self.params:
            mode.setdefault(col, CONST)                                 # fitfunc.py:383-387
        if mode['background'] == VAR:                                   # fitfunc.py:389-392
            warnings.warn('The background param mode cannot vary per feature. '
                          'Varying per cluster now.')
            mode['background'] = CLUSTER
        self.param_mode = mode
        self.modes = [int(mode[p]) for p in self.params]

    # -- objective ---------------------------------------------------------------------------
    def objective(self, values, mesh, masks, params_const, norm=1.):
        """fitfunc.py:421-489 for one cluster.  Returns (fun(vect), grad(vect) or None)."""
        n, n_cols = params_const.shape
        n_extra = len(self.extra)
        ndim, iso, safe, modes = self.ndim, self.isotropic, self.safe, self.modes
        n_pix = len(values)

        def fun(vect):                                                  # fitfunc.py:436-450
            if np.any(np.isnan(vect)):
                raise FitFailure("non-finite parameter vector")
            p = unpack_vector(vect, params_const, modes)
            diff = values - p[0, 0]
            for i in range(n):
                r2 = _reduced_r2(mesh[:, masks[i]], p[i], ndim, iso, safe)
                diff[masks[i]] -= p[i, 1] * self.value(r2, p[i, n_cols - n_extra:], ndim)
            return np.nansum(diff ** 2) / n_pix / norm

        if self.value_grad is None:                                     # fitfunc.py:452-453
            return fun, None

        def grad(vect):                                                 # fitfunc.py:455-487
            if np.any(np.isnan(vect)):
                raise FitFailure("non-finite parameter vector")
            p = unpack_vector(vect, params_const, modes)
            out = p.copy()
            diff = values - p[0, 0]
            derivs = np.zeros((n, n_cols - 1, n_pix))
            for i in range(n):
                m = masks[i]
                r2 = _reduced_r2(mesh[:, m], p[i], ndim, iso, safe)
                dr2 = _reduced_r2_grad(mesh[:, m], p[i], ndim, iso)
                model, dmodel = self.value_grad(r2, p[i, n_cols - n_extra:], ndim)
                diff[m] -= p[i, 1] * model
                derivs[i, 0, m] = model
                derivs[i, 1:1 + len(dr2), m] = p[i, 1] * (dmodel[0] * dr2).T
                if n_extra > 0:
                    derivs[i, -n_extra:, m] = p[i, 1] * np.array(dmodel[1:]).T
            out[:, 1:] = np.nansum(-2 * diff * derivs, axis=2) / n_pix
            out[:, 0] = np.nansum(-2 * diff) / (n * n_pix)
            return pack_vector(out, modes, np.sum) / norm

        return fun, grad

    # -- bounds ------------------------------------------------------------------------------
    def bounds_tables(self, bounds=None, radius=None):
        """fitfunc.py:492-533 -> (abs, diff, rel_diff), each shape (2, P)."""
        bounds = dict() if bounds is None else bounds
        n_p = len(self.params)
        abs_t, diff_t, rel_t = (np.empty((2, n_p)) for _ in range(3))
        for j, name in enumerate(self.params):
            a = bounds.get(name, np.nan)
            d = bounds.get(name + '_diff', np.nan)
            r = bounds.get(name + '_rel_diff', np.nan)
            for group, cols in (('pos', self.pos_columns), ('size', self.size_columns)):
                if name in cols:
                    if a is np.nan:
                        a = bounds.get(group, np.nan)
                    if d is np.nan:
                        d = bounds.get(group + '_diff', np.nan)
                    if r is np.nan:
                        r = bounds.get(group + '_rel_diff', np.nan)
            if a is np.nan and name in ['background', 'signal'] + self.size_columns:
                a = (0., np.nan)                                        # fitfunc.py:518-521
            if d is np.nan and name in self.pos_columns:
                half = float(radius[self.pos_columns.index(name)])      # fitfunc.py:523-527
                d = (half, half)
            abs_t[:, j], diff_t[:, j], rel_t[:, j] = a, d, r
        return abs_t, diff_t, rel_t

    def feature_bounds(self, tables, params):
        """fitfunc.py:538-551: per-feature, per-column (low, high) before packing."""
        abs_t, diff_t, rel_t = tables
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            low = np.nanmax([params - diff_t[0], params * (1 - rel_t[0])], axis=0)
            low = np.fmax(low, abs_t[0])
            low[np.isnan(low)] = -np.inf
            high = np.nanmin([params + diff_t[1], params * (1 + rel_t[1])], axis=0)
            high = np.fmin(high, abs_t[1])
            high[np.isnan(high)] = np.inf
        return low, high

    def cluster_bounds(self, tables, params):
        """fitfunc.py:535-558 -> (V, 2): shared entries take the widest bound."""
        low, high = self.feature_bounds(tables, params)
        return np.array([pack_vector(low, self.modes, np.min),
                         pack_vector(high, self.modes, np.max)], dtype=np.float64).T


# ----------------------------------------------------------------------------------------------
# global level: one problem over the whole table                       refine.py:319-332
# ----------------------------------------------------------------------------------------------
def pack_vector_groups(params, modes, groups, reduce_op=None):
    """fitfunc.py:207-263 WITH groups (the clusters): const skipped, var -> n entries, global -> one
    entry over all rows, cluster -> one entry per group."""
    chunks = []
    for j, mode in enumerate(modes):
        if mode == CONST:
            continue
        if mode == VAR:
            chunks.append(params[:, j])
        elif mode == GLOBAL:
            chunks.append([params[0, j] if reduce_op is None else reduce_op(params[:, j])])
        elif reduce_op is None:
            chunks.append(params[[g[0] for g in groups], j])
        else:
            chunks.append([reduce_op(params[g, j]) for g in groups])
    if not chunks:
        return np.empty((0,))
    return np.concatenate([np.asarray(c, dtype=np.float64) for c in chunks])


def unpack_vector_groups(vect, params, modes, groups):
    """fitfunc.py:266-315 WITH groups."""
    n = params.shape[0]
    out = params.copy()
    at = 0
    for j, mode in enumerate(modes):
        if mode == CONST:
            continue
        if mode == VAR:
            out[:, j] = vect[at:at + n]
            at += n
        elif mode == GLOBAL:
            out[:, j] = vect[at]
            at += 1
        else:
            for g, value in zip(groups, vect[at:at + len(groups)]):
                out[g, j] = value
            at += len(groups)
    return out


def _refine_global(f, frames, spec, tables, radius, ndim, t_column, noise_size, threshold,
                   max_iter, max_shift, max_rms_dev, residual_factor, constraints, solver):
    """refine.py:319-332 + 343-430 for level == 'global': ONE minimisation over all rows; the
    objective is the sum of the per-cluster terms of fitfunc.py:436-450 (each divided by its own
    pixel count), the norm comes from the brightest frame."""
    if constraints:
        raise NotImplementedError("constraints on a global fit (dimer_global) are out of scope")
    params = f[spec.params].values.astype(np.float64)
    # refine.py:349-350: groups = row positions per cluster, in the order of the cluster ids
    groups = [np.asarray(g) for g in f.reset_index().groupby('cluster').indices.values()]
    frame_nos = f[t_column].values
    norm = max(float(np.asarray(frames[int(i)]).max()) for i in np.unique(frame_nos)) ** 2 / residual_factor
    modes, n_cols, n_extra = spec.modes, params.shape[1], len(spec.extra)
    iso, safe = spec.isotropic, spec.safe
    try:
        if not np.isfinite(params).all():
            raise FitFailure("non-finite initial parameters")
        coords = params[:, 2:2 + ndim]
        x0 = pack_vector_groups(params, modes, groups, np.mean)              # refine.py:361
        low, high = spec.feature_bounds(tables, params)
        box = np.array([pack_vector_groups(low, modes, groups, np.min),
                        pack_vector_groups(high, modes, groups, np.max)]).T  # fitfunc.py:552-558
        for _ in range(max_iter):                                            # refine.py:365
            pix = [cluster_pixels(coords[g], np.asarray(frames[int(frame_nos[g[0]])]), radius,
                                  noise_size, threshold) for g in groups]    # refine.py:61-79

            def fun(vect):                                                   # fitfunc.py:436-450
                if np.any(np.isnan(vect)):
                    raise FitFailure("non-finite parameter vector")
                p = unpack_vector_groups(vect, params, modes, groups)
                total = 0.
                for g, (values, mesh, masks) in zip(groups, pix):
                    diff = values - p[g[0], 0]
                    for i, mask in zip(g, masks):
                        r2 = _reduced_r2(mesh[:, mask], p[i], ndim, iso, safe)
                        diff[mask] -= p[i, 1] * spec.value(r2, p[i, n_cols - n_extra:], ndim)
                    total += np.nansum(diff ** 2) / len(values)
                return total / norm

            def grad(vect):                                                  # fitfunc.py:455-487
                if np.any(np.isnan(vect)):
                    raise FitFailure("non-finite parameter vector")
                p = unpack_vector_groups(vect, params, modes, groups)
                out = p.copy()
                for g, (values, mesh, masks) in zip(groups, pix):
                    diff = values - p[g[0], 0]
                    derivs = np.zeros((len(g), n_cols - 1, len(values)))
                    for k, (i, m) in enumerate(zip(g, masks)):
                        r2 = _reduced_r2(mesh[:, m], p[i], ndim, iso, safe)
                        dr2 = _reduced_r2_grad(mesh[:, m], p[i], ndim, iso)
                        model, dmodel = spec.value_grad(r2, p[i, n_cols - n_extra:], ndim)
                        diff[m] -= p[i, 1] * model
                        derivs[k, 0, m] = model
                        derivs[k, 1:1 + len(dr2), m] = p[i, 1] * (dmodel[0] * dr2).T
                        if n_extra > 0:
                            derivs[k, -n_extra:, m] = p[i, 1] * np.array(dmodel[1:]).T
                    out[g, 1:] = np.nansum(-2 * diff * derivs, axis=2) / len(values)
                    out[g, 0] = np.nansum(-2 * diff) / (len(g) * len(values))
                return pack_vector_groups(out, modes, groups, np.sum) / norm

            res = minimize(fun, x0, bounds=box, jac=grad if spec.value_grad is not None else None,
                           **solver)
            if not res['success']:
                raise FitFailure(res['message'])
            rms_dev = np.sqrt(res['fun'] / residual_factor)
            params = unpack_vector_groups(res['x'], params, modes, groups)
            moved = params[:, 2:2 + ndim]
            if np.all(np.sum((moved - coords) ** 2, 1) < max_shift ** 2):
                break
            coords = moved
        if rms_dev > max_rms_dev:
            raise FitFailure("rms deviation %.4f above the maximum %.4f" % (rms_dev, max_rms_dev))
    except FitFailure as exc:                                                # refine.py:409-411
        f['cost'] = np.nan
        logger.warning('RefineException: %s', exc.args[0] if exc.args else '')
    else:                                                                    # refine.py:420-422
        f[spec.params] = params
        f['cost'] = rms_dev
    return f


# ----------------------------------------------------------------------------------------------
# constraints                                                          constraints.py:17-137
# ----------------------------------------------------------------------------------------------
def _pair_defect(pos, a, b, dist):
    return 1 - np.sum(((pos[:, a] - pos[:, b]) / dist) ** 2, axis=1)


def dimer_defect(x, dist, ndim):                                        # constraints.py:59-61
    return _pair_defect(x[..., 2:2 + ndim], 0, 1, dist)


def trimer_defect(x, dist, ndim):                                       # constraints.py:79-83
    pos = x[..., 2:2 + ndim]
    return np.concatenate((_pair_defect(pos, 0, 1, dist), _pair_defect(pos, 1, 2, dist),
                           _pair_defect(pos, 0, 2, dist)))


def tetramer_defect_2d(x, dist):                                        # constraints.py:102-114
    pos = x[..., 2:4]
    d = np.vstack([1 - _pair_defect(pos, a, b, dist)
                   for a, b in ((0, 1), (1, 2), (0, 2), (1, 3), (0, 3), (2, 3))])
    return np.ravel(1 - np.sort(d, axis=0)[:4])


def tetramer_defect_3d(x, dist):                                        # constraints.py:117-125
    pos = x[..., 2:5]
    return np.concatenate([_pair_defect(pos, a, b, dist)
                           for a, b in ((0, 1), (1, 2), (0, 2), (1, 3), (0, 3), (2, 3))])


def dimer(dist, ndim=2):                                                # constraints.py:70-76
    return (dict(type='eq', cluster_size=2, fun=dimer_defect,
                 args=(np.array(as_ndim_tuple(dist, ndim)), ndim)),)


def trimer(dist, ndim=2):                                               # constraints.py:93-99
    return (dict(type='eq', cluster_size=3, fun=trimer_defect,
                 args=(np.array(as_ndim_tuple(dist, ndim)), ndim)),)


def tetramer(dist, ndim=2):                                             # constraints.py:127-137
    dist = np.array(as_ndim_tuple(dist, ndim))
    fun = {2: tetramer_defect_2d, 3: tetramer_defect_3d}[ndim]
    return (dict(type='eq', cluster_size=4, fun=fun, args=(dist,)),)


def bind_constraints(constraints, params_const, modes):
    """constraints.py:17-56 with ``groups=None``: keep a constraint when its ``cluster_size`` is None
    or equals the cluster's size; the callable sees ``params[np.newaxis]``.  (Divergence C1: each
    closure binds its OWN constraint.)"""
    bound = []
    for cons in (constraints or ()):
        size = cons.get('cluster_size', None)
        if size is not None and len(params_const) != size:
            continue

        def call(vect, *args, _f=cons['fun']):
            return _f(unpack_vector(vect, params_const, modes)[np.newaxis, :, :], *args)

        item = {k: v for k, v in cons.items() if k != 'jac'}            # constraints.py:53-55
        item['fun'] = call
        bound.append(item)
    return bound


# ----------------------------------------------------------------------------------------------
# driver                                                                 refine.py:82-452
# ----------------------------------------------------------------------------------------------
def _frame_lookup(f, reader, pos_columns, t_column):
    """refine.py:252-283: FramesSequence-like reader, or a single ndarray wrapped in a dict."""
    try:
        ndim = len(reader.frame_shape)
        return reader, ndim
    except AttributeError:
        pass
    try:
        ndim = reader.ndim
    except AttributeError:
        raise ValueError('For multiple frames, the reader should be a FramesSequence object '
                         'exposing the "frame_shape" attribute')
    frame_no = getattr(reader, 'frame_no', None)
    frame_no = int(frame_no) if frame_no is not None else None
    if frame_no is not None and t_column in f:
        assert np.all(f['frame'] == frame_no)
        return {frame_no: reader}, ndim
    if frame_no is not None:
        f[t_column] = frame_no
        return {frame_no: reader}, ndim
    if t_column in f:
        assert f[t_column].nunique() == 1
        return {int(f[t_column].iloc[0]): reader}, ndim
    f[t_column] = 0
    return {0: reader}, ndim


# cluster id -> (outer iterations used, settled) of the most recent refine_leastsq call
OUTER_LOOP = {}


def refine_leastsq(f, reader, diameter, separation=None, fit_function='gauss', param_mode=None,
                   param_val=None, constraints=None, bounds=None, pos_columns=None,
                   t_column='frame', noise_size=None, threshold=None, max_iter=10, max_shift=1,
                   max_rms_dev=1., residual_factor=100000., compute_error=False, **kwargs):
    """Cluster-level ``refine_leastsq`` (refine.py:82-452; global branch 319-332 out of scope)."""
    solver = dict(method='SLSQP', tol=1E-6, options=dict(maxiter=100, disp=False))   # :242-244
    solver.update(kwargs)
    OUTER_LOOP.clear()
    if compute_error:
        raise NotImplementedError("compute_error is out of scope (SURVEY.md section 2 row 13)")
    if pos_columns is None:
        pos_columns = ['z', 'y', 'x'] if 'z' in f else ['y', 'x']
    frames, ndim = _frame_lookup(f, reader, pos_columns, t_column)
    assert ndim == len(pos_columns)

    diameter = as_ndim_tuple(diameter, ndim)                            # refine.py:285-289
    radius = tuple([d // 2 for d in diameter])
    isotropic = all_equal(diameter)
    if separation is None:
        separation = diameter

    spec = ModelSpec(fit_function, ndim, isotropic, param_mode)
    if any(m > CLUSTER for m in spec.modes):
        raise NotImplementedError("modes 'particle' and 'frame' are not implemented upstream")

    f = find_clusters(f, separation, pos_columns, t_column)             # refine.py:297 (copy)
    if param_val is not None:                                           # refine.py:300-302
        for col in param_val:
            f[col] = param_val[col]
    for col in [p for p in spec.params if p not in f.columns]:          # refine.py:303-305
        f[col] = spec.default[col]
    tables = spec.bounds_tables(bounds, radius)                         # refine.py:315
    if any(m == GLOBAL for m in spec.modes):                            # refine.py:319-332
        return _refine_global(f, frames, spec, tables, radius, ndim, t_column, noise_size, threshold,
                              max_iter, max_shift, max_rms_dev, residual_factor, constraints, solver)

    for _, group in f.groupby(['frame', 'cluster']):                    # refine.py:336, 343
        params = group[spec.params].values.astype(np.float64)
        frame = frames[group[t_column].values[0]]
        norm = float(frame.max()) ** 2 / residual_factor                # refine.py:354
        try:
            if not np.isfinite(params).all():                           # refine.py:356-357
                raise FitFailure("non-finite initial parameters")
            coords = params[:, 2:2 + ndim]
            x0 = pack_vector(params, spec.modes, np.mean)               # refine.py:361
            cons = bind_constraints(constraints, params, spec.modes)    # refine.py:363
            box = spec.cluster_bounds(tables, params)                   # refine.py:364
            settled = False
            for n_outer in range(1, max_iter + 1):                      # refine.py:365
                values, mesh, masks = cluster_pixels(coords, np.asarray(frame), radius,
                                                     noise_size, threshold)
                fun, grad = spec.objective(values, mesh, masks, params, norm)
                res = minimize(fun, x0, bounds=box, constraints=cons, jac=grad, **solver)
                if not res['success']:                                  # refine.py:376-377
                    raise FitFailure(res['message'])
                rms_dev = np.sqrt(res['fun'] / residual_factor)         # refine.py:379
                params = unpack_vector(res['x'], params, spec.modes)
                moved = params[:, 2:2 + ndim]
                if np.all(np.sum((moved - coords) ** 2, 1) < max_shift ** 2):    # :383-385
                    settled = True
                    break
                coords = moved                                          # refine.py:388
            if rms_dev > max_rms_dev:                                   # refine.py:391-394
                raise FitFailure("rms deviation %.4f above the maximum %.4f"
                                 % (rms_dev, max_rms_dev))
        except FitFailure as exc:                                       # refine.py:408-418
            f.loc[group.index, 'cost'] = np.nan
            logger.warning('RefineException: %s', exc.args[0] if exc.args else '')
        else:                                                           # refine.py:426-427
            f.loc[group.index, spec.params] = params
            f.loc[group.index, 'cost'] = rms_dev
            # diagnostics (not part of the reference's output): did the re-mask loop settle, or was
            # the last of max_iter results taken while the mask was still moving (refine.py:365-388)?
            OUTER_LOOP[int(group['cluster'].values[0])] = (n_outer, settled)
    return f
