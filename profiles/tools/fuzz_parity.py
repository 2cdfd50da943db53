"""Randomised parity sweep: random model family / dimension / modes / bounds / constraints / lowpass on
small synthetic frames (tests/fuzz_cases.py), the CUDA path against the CPU oracle at the reference's
default tolerance AND at tol=1e-12.

    python profiles/tools/fuzz_parity.py [cases] [seed]

A case passes when every cluster is within 1e-3 px of the default-tolerance reference or ends at a
cost not above the reference's at both tolerances (tests/fuzz_cases.judge).
FUZZ_EMUL=1: run the one-lane host build of the device solver instead of the GPU (build container).
FUZZ_ONLY=3,17: only these cases.  FUZZ_CACHE=dir: keep the oracle's answers between runs.
FUZZ_PRECISION=float64: pixel arithmetic of our side."""
import json
import os
import pickle
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import fuzz_cases
from oracle import cluster_oracle

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 30
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
only = [int(v) for v in os.environ.get('FUZZ_ONLY', '').split(',') if v] or None
cache_dir = os.environ.get('FUZZ_CACHE')
precision = os.environ.get('FUZZ_PRECISION')
if os.environ.get('FUZZ_EMUL'):          # CPU: the one-lane host build of the device solver (tests/emul)
    import emul_backend
    from clustertracking_b200 import constraints
    run_ours = lambda f, fr, d, **kw: emul_backend.refine_leastsq(f, fr, d, **kw)[0]
else:
    import clustertracking_b200 as ctb
    from clustertracking_b200 import constraints
    run_ours = lambda f, fr, d, **kw: ctb.refine_leastsq(f, fr, d, **kw)


def oracle_answers(case):
    path = cache_dir and os.path.join(cache_dir, "oracle_%d_%d.pkl" % (case['seed'], case['case']))
    if path and os.path.exists(path):
        with open(path, 'rb') as fh:
            return pickle.load(fh)
    okw = fuzz_cases.bind(case, cluster_oracle)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = cluster_oracle.refine_leastsq(case['f0'].copy(), case['frame'], case['diameter'], **okw)
        loops = dict(cluster_oracle.OUTER_LOOP)
        tight = cluster_oracle.refine_leastsq(case['f0'].copy(), case['frame'], case['diameter'],
                                              tol=1e-12, options=dict(maxiter=1000), **okw)
        # clusters whose re-mask loop did not settle in either reference run (see fuzz_cases.judge)
        unsettled = sorted(set(c for run in (loops, cluster_oracle.OUTER_LOOP)
                               for c, (_, settled) in run.items() if not settled))
    if path:
        os.makedirs(cache_dir, exist_ok=True)
        with open(path, 'wb') as fh:
            pickle.dump((ref, tight, unsettled), fh)
    return ref, tight, unsettled


bad = outside = 0
worst = 0.
for case in fuzz_cases.cases(seed, n_cases, only):
    kwargs = fuzz_cases.bind(case, constraints)
    if precision:
        kwargs['precision'] = precision
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        got = run_ours(case['f0'].copy(), case['frame'], case['diameter'], **kwargs)
    ref, tight, unsettled = oracle_answers(case)
    verdict = fuzz_cases.judge(got, ref, tight, case['cols'], unsettled=unsettled)
    bad += not verdict['ok']
    outside += not verdict['within']
    worst = max(worst, verdict['dpos'])
    if os.environ.get('FUZZ_VERBOSE') and not verdict['within']:
        print("   cost ours ", got['cost'].values.round(7).tolist())
        print("   cost ref  ", ref['cost'].values.round(7).tolist())
        print("   cost tight", tight['cost'].values.round(7).tolist())
    tag = "ok  " if verdict['within'] else ("low " if verdict['ok'] else "BAD ")
    print(tag + json.dumps(dict(case=case['case'], **case['meta'],
                                kwargs={k: v for k, v in case['kwargs'].items()},
                                constraint=case['constraint'], **verdict)), flush=True)
print("seed", seed, "cases", n_cases, "beyond 1e-3 px:", outside,
      "of those above the reference's cost (BAD):", bad, "worst dpos", worst)
