"""TEST INFRASTRUCTURE: the randomised option sweep of ``profiles/tools/fuzz_parity.py`` as a
generator, so that the tool, the golden generator (``oracle/make_golden.py``) and the GPU tests
(``tests/test_gpu_fuzz.py``) draw exactly the same cases.

Case k of a seed depends on the draws of cases 0..k-1 (one random stream per seed): ``cases(seed,
n)`` walks the stream and yields every case; rendering is skipped for the cases not in ``only``.
"""
import numpy as np
import pandas as pd

from clustertracking_b200 import artificial


def cases(seed, n_cases, only=None):
    """Yields dict(case, frame, f0, diameter, kwargs, constraint, meta) for the cases of ``seed``
    (all, or those listed in ``only``).  ``constraint`` is None or ('dimer', dist): bind it with the
    constraints module of the implementation under test."""
    rng = np.random.default_rng(seed)
    for case in range(n_cases):
        ndim = int(rng.choice([2, 2, 3]))
        family = str(rng.choice(['gauss', 'gauss', 'ring', 'disc']))
        iso = bool(rng.random() < 0.6)
        if ndim == 2:
            shape = (int(rng.integers(90, 140)), int(rng.integers(90, 140)))
            size = (4., 4.) if iso else (4.5, 3.)
            pitch = 40
        else:
            shape = (40, 72, 72)
            size = (2.5, 2.5, 2.5) if iso else (2.25, 3.25, 3.25)
            pitch = 26
        diameter = tuple(int(4 * s) for s in size)
        centres = artificial.jittered_grid(shape, pitch, 14 if ndim == 3 else pitch // 2 + 2, 2, rng)
        kmax = int(rng.integers(1, 4))
        counts = rng.integers(1, kmax + 1, len(centres))
        pos, _ = artificial.grow_clusters(rng, centres, counts, tuple(2 * s for s in size))
        extra = {}
        if family == 'ring':
            extra = dict(thickness=0.25)
        if family == 'disc':
            extra = dict(disc_size=0.5)
        noise = int(rng.choice([0, 3, 8]))
        # the draws below must happen for every case (they advance the stream)
        frame = artificial.draw_features(shape, pos, size, rng.uniform(100, 180, len(pos)),
                                         feat_func=family, noise=noise, rng=rng, **extra)
        cols = ['z', 'y', 'x'][-ndim:]
        start = pos + rng.uniform(-0.4, 0.4, pos.shape)
        if rng.random() < 0.3:
            start = np.round(start)
        f0 = pd.DataFrame(start, columns=cols)
        f0['signal'] = 140.
        if iso:
            f0['size'] = size[0]
        else:
            for c, s in zip(cols, size):
                f0['size_' + c] = s
        kwargs = dict(fit_function=family)
        if extra:
            kwargs['param_val'] = extra
        mode = {}
        if rng.random() < 0.4:
            mode['size'] = 'var'
        if rng.random() < 0.2:
            mode['signal'] = 'cluster'
        if rng.random() < 0.15 and family != 'gauss':
            mode[list(extra)[0]] = 'var'
        if mode:
            kwargs['param_mode'] = mode
        if rng.random() < 0.25:
            kwargs['bounds'] = dict(pos_diff=3.0, signal=(10, 400))
        if rng.random() < 0.25:
            kwargs['noise_size'] = float(rng.choice([0.7, 1.0]))
        constraint = None
        if rng.random() < 0.2 and ndim == 2:
            constraint = ('dimer', tuple(2 * s for s in size))
        if only is not None and case not in only:
            continue
        yield dict(case=case, seed=seed, frame=frame, f0=f0, diameter=diameter, kwargs=kwargs,
                   constraint=constraint, cols=cols,
                   meta=dict(ndim=ndim, family=family, iso=iso, n=len(f0), noise=noise))


def bind(case, constraints_module):
    """kwargs of the case with its constraint bound through ``constraints_module.dimer``."""
    kwargs = dict(case['kwargs'])
    if case['constraint'] is not None:
        kind, dist = case['constraint']
        kwargs['constraints'] = getattr(constraints_module, kind)(dist, len(case['cols']))
    return kwargs


def judge(got, ref, tight, cols, pos_tol=1e-3, cost_slack=1e-6, unsettled=()):
    """Per-cluster verdict of one case.  A cluster passes when its positions are within ``pos_tol``
    of the reference at its default tolerance, OR its cost is not above the reference's cost at
    BOTH tolerances (a failed reference fit counts as infinite cost).  -> dict(ok, dpos, worse):
    ``worse`` = clusters that fail both tests.

    ``unsettled``: ids of clusters whose re-mask loop (refine.py:365-388) did NOT settle in the
    reference -- SLSQP's first step threw a centre against its bound, the next mask was cut around
    that far point, the fit of the (empty) region there returned the start vector, and so on until
    ``max_iter`` ran out; the reference then reports whichever of the two states the last iteration
    held.  Such a cluster is no target (its answer flips with the parity of ``max_iter``): it is
    listed in ``cycling`` and counts neither as a pass nor as a failure."""
    assert np.array_equal(got['cluster'].values, ref['cluster'].values)
    d = np.abs(got[cols].values - ref[cols].values).max(axis=1)
    c_got = got['cost'].values
    c_ref = np.where(np.isnan(ref['cost'].values), np.inf, ref['cost'].values)
    c_tight = np.where(np.isnan(tight['cost'].values), np.inf, tight['cost'].values)
    close = np.where(np.isfinite(c_ref), d < pos_tol, False) & ~np.isnan(c_got)
    lower = (c_got <= c_ref * (1 + cost_slack)) & (c_got <= c_tight * (1 + cost_slack))
    cyc = np.isin(got['cluster'].values, list(unsettled))
    ok_rows = close | lower | cyc
    both = np.isfinite(c_ref) & ~np.isnan(c_got) & ~cyc
    return dict(ok=bool(ok_rows.all()), dpos=float(d[both].max()) if both.any() else 0.,
                within=bool(close[both].all()) if both.any() else True,
                worse=sorted(set(int(c) for c in got['cluster'].values[~ok_rows])),
                cycling=sorted(set(int(c) for c in got['cluster'].values[cyc & ~(close | lower)])),
                fail_ours=int(np.isnan(c_got).sum()), fail_ref=int(np.isinf(c_ref).sum()))
