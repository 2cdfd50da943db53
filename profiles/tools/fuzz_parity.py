"""Randomised parity sweep: random model family / dimension / modes / bounds / constraints / lowpass on
small synthetic frames, CUDA path against the CPU oracle.  python profiles/tools/fuzz_parity.py [cases] [seed]"""
import json, os, sys, warnings
import numpy as np
import pandas as pd
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import clustertracking_b200 as ctb
from clustertracking_b200 import artificial, constraints
from oracle import cluster_oracle

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 30
only = [int(v) for v in os.environ.get('FUZZ_ONLY', '').split(',') if v]
if os.environ.get('FUZZ_EMUL'):          # CPU: the one-lane host build of the device solver (tests/emul)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import emul_backend
    run_ours = lambda f, fr, d, **kw: emul_backend.refine_leastsq(f, fr, d, **kw)[0]
else:
    run_ours = lambda f, fr, d, **kw: ctb.refine_leastsq(f, fr, d, **kw)
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
worst = dict(dpos=0., dsig=0.)
bad = 0
worse = 0
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
    frame = artificial.draw_features(shape, pos, size, rng.uniform(100, 180, len(pos)), feat_func=family,
                                     noise=noise, rng=rng, **extra)
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
    okw = dict(kwargs)
    if rng.random() < 0.2 and ndim == 2:
        d = tuple(2 * s for s in size)
        kwargs['constraints'] = constraints.dimer(d, ndim)
        okw['constraints'] = cluster_oracle.dimer(d, ndim)
    if only and case not in only:
        continue
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        got = run_ours(f0.copy(), frame, diameter, **kwargs)
        want = cluster_oracle.refine_leastsq(f0.copy(), frame, diameter, **okw)
        if os.environ.get('FUZZ_TIGHT'):
            tight = cluster_oracle.refine_leastsq(f0.copy(), frame, diameter, tol=1e-12,
                                                  options=dict(maxiter=1000), **okw)
            print("   cost ours", got['cost'].values.round(7).tolist())
            print("   cost ref ", want['cost'].values.round(7).tolist())
            print("   cost tight", tight['cost'].values.round(7).tolist())
            print("   dpos vs tight", float(np.nanmax(np.abs(got[cols].values - tight[cols].values))),
                  " ref vs tight", float(np.nanmax(np.abs(want[cols].values - tight[cols].values))))
    same_clusters = np.array_equal(got['cluster'].values, want['cluster'].values)
    both = ~np.isnan(got['cost'].values) & ~np.isnan(want['cost'].values)
    dpos = np.abs(got[cols].values[both] - want[cols].values[both]).max() if both.any() else 0.
    dsig = np.abs(got['signal'].values[both] / np.maximum(want['signal'].values[both], 1e-9) - 1).max() if both.any() else 0.
    ok = same_clusters and dpos < 1e-3
    # where the answers differ: is ours the better minimum?  (cost = rms residual / frame max)
    not_worse = bool(np.all(got['cost'].values[both] <= want['cost'].values[both] * (1 + 1e-5) + 1e-9))
    bad += not ok
    worst['dpos'] = max(worst['dpos'], float(dpos))
    print(("ok  " if ok else "BAD ") + json.dumps(dict(case=case, ndim=ndim, family=family, iso=iso, n=len(f0),
          kwargs={k: (v if k != 'constraints' else 'dimer') for k, v in kwargs.items()}, noise=noise,
          fail_ours=int(np.isnan(got['cost']).sum()), fail_oracle=int(np.isnan(want['cost']).sum()),
          dpos=float(dpos), dsignal=float(dsig), ours_cost_not_above_ref=not_worse)), flush=True)
    worse += (not ok) and (not not_worse)
print("cases", n_cases, "bad", bad, "worst dpos", worst['dpos'], "outliers where our cost is above the reference's:", worse)
