"""Full-frame parity on the BASELINE.json configurations (``-m gpu``): the CUDA path against the CPU
oracle (scipy SLSQP, ``oracle/cluster_oracle.py``, pinned to the unmodified reference by
tests/golden) on one whole frame / stack of each workload, at BASELINE's sizes:

  config 2  1024x1024 uint8, ~2100 gaussian features in clusters of 2-6, diameter 11
  config 3  512x512, 60 dimers + 40 trimers, ``constraints.dimer(8) + constraints.trimer(8)`` in ONE call
  config 4  64x256x256 anisotropic stack, clusters of 1-4, per-axis size free ('var')
  config 5  find -> refine: ``grey_dilation`` maxima of a config-2 frame, refined from the integer pixels

plus size-independent properties on a multi-frame video of each (membership and order identical to
the per-frame result, chunk invariance, no failures).  Tolerance (BASELINE.json north_star): 1e-3 px
in position, 1e-3 relative in signal and size, identical cluster membership and feature order."""
import warnings

import numpy as np
import pandas as pd
import pytest
from numpy.testing import assert_allclose, assert_array_equal

pytestmark = pytest.mark.gpu

POS_TOL, REL_TOL = 1e-3, 1e-3


def _oracle(f0, image, diameter, **kwargs):
    from oracle import cluster_oracle
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return cluster_oracle.refine_leastsq(f0.copy(), image, diameter, **kwargs)


def _check(got, want, pos_cols, rel_cols, outliers=0):
    """Membership, order and sizes identical; every cluster within tolerance of the oracle, except
    at most ``outliers`` FEATURES in clusters that fail in one solver only or end in another basin
    (see config 5).  -> number of features compared within tolerance."""
    assert_array_equal(got.index.values, want.index.values)
    assert_array_equal(got['cluster'].values, want['cluster'].values)
    assert_array_equal(got['cluster_size'].values, want['cluster_size'].values)
    bad = np.isnan(got['cost'].values) != np.isnan(want['cost'].values)
    both = ~np.isnan(want['cost'].values) & ~np.isnan(got['cost'].values)
    for col in pos_cols:
        bad |= both & ~(np.abs(got[col].values - want[col].values) <= POS_TOL)
    for col in rel_cols:
        bad |= both & ~(np.abs(got[col].values - want[col].values)
                        <= REL_TOL * np.abs(want[col].values) + 1e-9)
    bad |= both & ~(np.abs(got['cost'].values - want['cost'].values)
                    <= REL_TOL * np.abs(want['cost'].values) + 1e-6)
    bad = np.isin(got['cluster'].values, got['cluster'].values[bad])       # whole clusters
    assert bad.sum() <= outliers, (int(bad.sum()), got[bad], want[bad])
    return int((both & ~bad).sum())


@pytest.mark.parametrize("precision", ["float32", "float64"])
def test_config2_full_frame(precision):
    import clustertracking_b200 as ctb
    from clustertracking_b200 import artificial
    frame, f0, _ = artificial.clustered_frame((1024, 1024), seed=11)
    got = ctb.refine_leastsq(f0.copy(), frame, 11, precision=precision)
    want = _oracle(f0, frame, 11)
    assert len(got) > 1800 and got['cluster_size'].max() >= 5
    assert _check(got, want, ['y', 'x'], ['signal']) == len(got)


@pytest.mark.parametrize("precision", ["float32", "float64"])
def test_config3_dimers_and_trimers_in_one_call(precision):
    """Both constraint kinds in ONE call, each applied to clusters of its own size.  The reference
    cannot do that (late-binding closures), so the oracle gets them the way this
    repository applies them; the reference-pinned single-kind cases are the goldens."""
    import clustertracking_b200 as ctb
    from clustertracking_b200 import artificial, constraints
    from oracle import cluster_oracle
    reader, f0 = artificial.dimer_trimer_video(1)
    kwargs = dict(param_mode=dict(signal='var', size='const'))
    got = ctb.refine_leastsq(f0.copy(), reader, 16, precision=precision,
                             constraints=constraints.dimer(8.0) + constraints.trimer(8.0), **kwargs)
    want = _oracle(f0, reader, 16, constraints=cluster_oracle.dimer(8.0, 2) + cluster_oracle.trimer(8.0, 2),
                   **kwargs)
    assert sorted(np.unique(got['cluster_size'].values)) == [2, 3]
    assert _check(got, want, ['y', 'x'], ['signal']) == len(got)
    for _, g in got.groupby('cluster'):                         # the constraints hold
        p = g[['y', 'x']].values
        for a in range(len(p)):
            for b in range(a + 1, len(p)):
                assert abs(1 - np.sum(((p[a] - p[b]) / 8.) ** 2)) < 1e-6


@pytest.mark.parametrize("precision", ["float32", "float64"])
def test_config4_anisotropic_stack(precision):
    import clustertracking_b200 as ctb
    from clustertracking_b200 import artificial
    reader, f0 = artificial.confocal_video(1)
    kwargs = dict(param_mode=dict(signal='var', size='var'))
    got = ctb.refine_leastsq(f0.copy(), reader, (9, 13, 13), precision=precision, **kwargs)
    want = _oracle(f0, reader, (9, 13, 13), **kwargs)
    assert got['cluster_size'].max() >= 3
    n_ok = _check(got, want, ['z', 'y', 'x'], ['signal', 'size_z', 'size_y', 'size_x'])
    assert n_ok >= len(got) - 4


def test_config5_find_then_refine():
    """find -> refine on a config-2 frame: the maxima equal the oracle's exactly (integer work) and
    the refinement from those integer pixels meets the tolerance."""
    import clustertracking_b200 as ctb
    from clustertracking_b200 import artificial, find
    from oracle import find_oracle
    frame, _, truth = artificial.clustered_frame((1024, 1024), seed=12)
    maxima = find.grey_dilation(frame, 5, percentile=95, margin=6)
    assert_array_equal(maxima, find_oracle.grey_dilation(frame, 5, percentile=95, margin=6))
    assert len(maxima) > 1500
    f0 = pd.DataFrame(dict(y=maxima[:, 0].astype(float), x=maxima[:, 1].astype(float), signal=120.,
                           size=2.75))
    got = ctb.refine_leastsq(f0.copy(), frame, 11)
    want = _oracle(f0, frame, 11)
    # Starting from integer pixels with a blind guess of the amplitude (120, background 0), a few
    # maxima are not clean features: a dim neighbour merged into the maximum gives the objective
    # two minima inside the mask, a noise maximum drifts to its position bound.  Which end point a
    # solver reaches there depends on its path through the re-mask loop (refine.py:365-388); on this
    # frame 2 of 1584 features differ from SLSQP's answer (one fails here only, one ends 0.09 px
    # away).  At most 0.2 % of the features may do so; all others must meet the tolerance.
    n_ok = _check(got, want, ['y', 'x'], ['signal'], outliers=int(0.002 * len(got)))
    assert n_ok >= 0.98 * len(got)
    # and the refined positions are the rendered features (rms error well below a pixel)
    from scipy.spatial import cKDTree
    ok = ~np.isnan(got['cost'].values)
    dist, _ = cKDTree(truth).query(got[['y', 'x']].values[ok])
    assert np.median(dist) < 0.5 and (dist < 2).mean() > 0.9


@pytest.mark.parametrize("config", [2, 3, 4])
def test_video_equals_frame_by_frame(config):
    """Size-independent property at video scale: refining a whole video gives, frame by frame,
    exactly the table that refining that frame alone gives (same clusters, same order, same
    numbers -- every (frame, cluster) group is an independent problem, refine.py:333-343), with the
    cluster ids running on across frames (find.py:127-128)."""
    import clustertracking_b200 as ctb
    from clustertracking_b200 import artificial, constraints
    if config == 2:
        reader, f0 = artificial.clustered_video(6)
        diameter, kwargs = 11, {}
    elif config == 3:
        reader, f0 = artificial.dimer_trimer_video(8)
        diameter = 16
        kwargs = dict(constraints=constraints.dimer(8.0) + constraints.trimer(8.0))
    else:
        reader, f0 = artificial.confocal_video(3)
        diameter, kwargs = (9, 13, 13), dict(param_mode=dict(size='var'))
    whole = ctb.refine_leastsq(f0.copy(), reader, diameter, **kwargs)
    assert_array_equal(whole['frame'].values, np.sort(f0['frame'].values))
    next_id = 0
    for t in np.unique(f0['frame'].values):
        sub = f0[f0['frame'] == t]
        single = ctb.refine_leastsq(sub.copy(), reader, diameter, **kwargs)
        part = whole[whole['frame'] == t]
        assert_array_equal(part.index.values, single.index.values)
        assert_array_equal(part['cluster'].values, single['cluster'].values + next_id)
        next_id = int(part['cluster'].max()) + 1
        for col in single.columns:
            if col != 'cluster':
                assert_array_equal(part[col].values, single[col].values, err_msg=col)


@pytest.mark.parametrize("precision", ["float32", "float64"])
@pytest.mark.parametrize("ny,nx", [(3, 4), (4, 4), (4, 5), (4, 7), (5, 6)])
def test_dense_cluster_is_fitted_not_dropped(ny, nx, precision):
    """12-30 narrow features on a 4 px lattice at diameter 11: more overlapping pairs than the 4 n
    lists even the rigorous shared-memory capacities provision.  They used to come back with
    cost = NaN; the final large-cluster launch of ``DeviceSession._run`` takes them
    without a host round trip."""
    import clustertracking_b200 as ctb
    from test_solver_emul import dense_cluster_case
    image, f0, pos = dense_cluster_case(ny, nx)
    got = ctb.refine_leastsq(f0.copy(), image, 11, precision=precision)
    assert got['cluster_size'].values[0] == ny * nx
    assert np.isfinite(got['cost'].values).all()
    assert np.abs(got[['y', 'x']].values - pos).max() < 0.1
