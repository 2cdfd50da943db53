"""Host mirror of the reference's ``FitFunctions`` (clustertracking/fitfunc.py:318-558).

Only the bookkeeping lives on the host: the parameter list, the modes, the defaults and the bounds
tables.  The model functions, the residual and its analytic Jacobian (fitfunc.py:14-146, 421-489)
are evaluated inside the CUDA solver (csrc/ctk_solver.cuh); there is no numpy evaluation path here.
"""
import warnings

import numpy as np

from .utils import default_pos_columns, default_size_columns

MODE_DICT = {0: 0, 1: 1, 2: 2, 3: 3, 4: 4, 5: 5, 6: 6,
             'const': 0, 'var': 1, 'global': 2, 'cluster': 3, 'particle': 4, 'frame': 5}

# name -> (family code of include/ctk.h, extra parameter names, defaults)     fitfunc.py:195-204
_FAMILIES = {
    'gauss': (0, [], {}),
    'ring': (1, ['thickness'], dict(thickness=0.5)),
    'disc': (2, ['disc_size'], dict(disc_size=0.5)),
}


class FitFunctions(object):
    """Parameter bookkeeping for one call of ``refine_leastsq``.

    Attributes (same names as the reference): ``params`` (column order background, signal,
    positions, sizes, extras), ``modes`` (0 const, 1 var, 2 global, 3 cluster), ``param_mode``,
    ``default``, ``pos_columns``, ``size_columns``, ``ndim``, ``isotropic``.
    """

    def __init__(self, fit_function='gauss', ndim=2, isotropic=True, param_mode=None):
        if isinstance(fit_function, dict):
            raise NotImplementedError("custom (dict) fit functions cannot run in the CUDA solver")
        if fit_function not in _FAMILIES:
            if str(fit_function).startswith('inv_series'):
                raise NotImplementedError("the inv_series family is not built into the CUDA solver")
            raise ValueError("Unknown fit function {}".format(fit_function))
        self.name = fit_function
        self.family, extra, extra_default = _FAMILIES[fit_function]
        self.ndim = ndim
        self.isotropic = isotropic
        self.pos_columns = default_pos_columns(ndim)
        self.size_columns = default_size_columns(ndim, isotropic)
        self._params = list(extra)
        self.default = dict(background=0., **extra_default)               # fitfunc.py:349
        self.continuous = fit_function == 'gauss'
        self.has_jacobian = True          # the CUDA solver differentiates every family analytically
        self.params = ['background', 'signal'] + self.pos_columns + self.size_columns + self._params

        mode = dict(signal='var', background='cluster')                   # fitfunc.py:356-360
        mode.update(param_mode or {})
        for group, cols in (('pos', self.pos_columns),                    # fitfunc.py:362-373
                            ('size', self.size_columns if not isotropic else [])):
            if group in mode and (group == 'pos' or not isotropic):
                for col in cols:
                    mode.setdefault(col, mode[group])
                del mode[group]
        mode = {key: MODE_DICT[val] for key, val in mode.items()}          # fitfunc.py:375-377
        for col in self.pos_columns:
            mode.setdefault(col, 1)
        for col in self.params:
            mode.setdefault(col, 0)
        if mode['background'] == 1:                                       # fitfunc.py:389-392
            warnings.warn('The background param mode cannot vary per feature. '
                          'Varying per cluster now.')
            mode['background'] = 3
        self.param_mode = mode
        self.modes = [int(mode[p]) for p in self.params]

    def get_residual(self, *args, **kwargs):
        raise NotImplementedError(
            "the residual and Jacobian are evaluated on the GPU (ctk_refine_batch); "
            "this class only keeps the parameter bookkeeping")

    # ---- bounds ---------------------------------------------------------------------------------
    def validate_bounds(self, bounds=None, radius=None):
        """User dict -> (abs, diff, rel_diff) tables of shape (2, P)  (fitfunc.py:492-533).

        Keys: ``'<param>'`` absolute (low, high); ``'<param>_diff'`` and ``'<param>_rel_diff'``
        one- or two-sided; ``'pos*'`` / ``'size*'`` broadcast to the position / size columns.
        Defaults: background, signal and sizes >= 0; positions within the mask radius."""
        bounds = {} if bounds is None else bounds
        n_p = len(self.params)
        tables = [np.empty((2, n_p), dtype=np.float64) for _ in range(3)]
        suffixes = ('', '_diff', '_rel_diff')
        for j, name in enumerate(self.params):
            entry = [bounds.get(name + s, np.nan) for s in suffixes]
            for group, cols in (('pos', self.pos_columns), ('size', self.size_columns)):
                if name in cols:
                    entry = [bounds.get(group + s, np.nan) if e is np.nan else e
                             for e, s in zip(entry, suffixes)]
            if entry[0] is np.nan and name in ['background', 'signal'] + self.size_columns:
                entry[0] = (0., np.nan)
            if entry[1] is np.nan and name in self.pos_columns:
                half = float(radius[self.pos_columns.index(name)])
                entry[1] = (half, half)
            for table, value in zip(tables, entry):
                table[:, j] = value
        return tuple(tables)

    def feature_bounds(self, tables, params):
        """Per-feature (low, high) arrays, shape (N, P): the narrowest of the absolute, difference
        and relative bounds (fitfunc.py:538-551).  Vectorised over every feature of the call; the
        packing of shared entries (widest bound, fitfunc.py:552-558) happens on the device."""
        abs_t, diff_t, rel_t = tables
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            low = np.fmax(np.fmax(params - diff_t[0], params * (1 - rel_t[0])), abs_t[0])
            high = np.fmin(np.fmin(params + diff_t[1], params * (1 + rel_t[1])), abs_t[1])
        low = np.where(np.isnan(low), -np.inf, low)
        high = np.where(np.isnan(high), np.inf, high)
        return low, high
