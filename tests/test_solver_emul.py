"""CPU tests of the solver LOGIC through the one-lane host build of the device source
(tests/emul): the same ctk_solver.cuh that nvcc compiles for sm_100a, with a 1-lane "warp".
This is a logic check for the GPU-less container, not a product path; the parity tests proper are
tests/test_gpu_parity.py (``-m gpu``).  Tolerances as there: 1e-3 px / 1e-3 relative against the
reference's default run, 2e-5 against its tol=1e-12 run."""
import warnings

import numpy as np
import pytest
from numpy.testing import assert_allclose, assert_array_equal

import golden_io
import emul_backend
import clustertracking_b200 as ctb
from test_gpu_parity import (_compare, check_tetramer2d, check_basins, POS_TOL, REL_TOL, POS_TOL_TIGHT,
                             REL_TOL_TIGHT, STRICT, BASIN)

CASES = STRICT


@pytest.mark.parametrize("name", CASES)
def test_emulated_solver_matches_reference_golden(name):
    d = golden_io.load(name)
    f0, reader, diameter, kwargs = golden_io.refine_inputs(d, ctb.constraints)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        got, plan = emul_backend.refine_leastsq(f0, reader, diameter, **kwargs)
    _compare(got, golden_io.frame(d, "ref_"), POS_TOL, REL_TOL)
    _compare(got, golden_io.frame(d, "tight_"), POS_TOL_TIGHT, REL_TOL_TIGHT)


@pytest.mark.parametrize("name", ["refine_gauss2d_clusters", "refine_trimer2d_constrained",
                                  "refine_gauss3d_aniso", "refine_ring2d"])
def test_emulated_solver_float64(name):
    d = golden_io.load(name)
    f0, reader, diameter, kwargs = golden_io.refine_inputs(d, ctb.constraints)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        got, _ = emul_backend.refine_leastsq(f0, reader, diameter, precision='float64', **kwargs)
    _compare(got, golden_io.frame(d, "tight_"), 1e-6, 1e-6)


@pytest.mark.parametrize("name", [n for n in BASIN if "3d" not in n])
def test_emulated_ring_disc_basins(name):
    d = golden_io.load(name)
    f0, reader, diameter, kwargs = golden_io.refine_inputs(d, ctb.constraints)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        got, _ = emul_backend.refine_leastsq(f0, reader, diameter, **kwargs)
    check_basins(got, d)


# a few 2D cases of the randomised sweep (the full fixed-seed subset runs on the GPU,
# tests/test_gpu_fuzz.py; the one-lane build is slow on 3D)
@pytest.mark.parametrize("name", ["fuzz_1_35", "fuzz_2_22", "fuzz_3_17", "fuzz_3_24", "fuzz_4_40"])
def test_emulated_fuzz_subset(name):
    d = golden_io.load(name)
    f0, reader, diameter, kwargs = golden_io.refine_inputs(d, ctb.constraints)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        got, _ = emul_backend.refine_leastsq(f0, reader, diameter, **kwargs)
    check_basins(got, d)


@pytest.mark.parametrize("precision", ["float32", "float64"])
def test_emulated_tetramer2d(precision):
    d = golden_io.load("refine_tetramer2d_constrained")
    f0, reader, diameter, kwargs = golden_io.refine_inputs(d, ctb.constraints)
    got, _ = emul_backend.refine_leastsq(f0.copy(), reader, diameter, precision=precision, **kwargs)
    check_tetramer2d(got, d, f0, np.asarray(reader), 1e-6 if precision == 'float64' else 2e-5)


def test_constraints_are_satisfied():
    d = golden_io.load("refine_trimer2d_constrained")
    f0, reader, diameter, kwargs = golden_io.refine_inputs(d, ctb.constraints)
    got, _ = emul_backend.refine_leastsq(f0, reader, diameter, **kwargs)
    for _, g in got.groupby('cluster'):
        p = g[['y', 'x']].values
        if len(p) == 3:
            for a, b in ((0, 1), (1, 2), (0, 2)):
                assert abs(np.linalg.norm(p[a] - p[b]) - 8.0) < 1e-6


def test_failure_statuses():
    import pandas as pd
    img = np.zeros((40, 40), np.uint8)
    img[18:23, 18:23] = 100
    f = pd.DataFrame(dict(y=[20., 500., 20.], x=[20., 500., 32.], signal=[100., 100., np.nan],
                          size=2.))
    got, plan = emul_backend.refine_leastsq(f, img, 9, separation=4)
    res = emul_backend.execute(plan)
    by_row = dict(zip(plan.order[plan.cluster_offset[:-1]], res.status))
    assert by_row[0] == 0                        # fine
    assert by_row[1] == 2                        # outside of the image  (refine.py:33-34)
    assert by_row[2] == 1                        # non-finite parameters (refine.py:356-357)
    assert np.isnan(got['cost'].values[1:]).all() and np.isfinite(got['cost'].values[0])
    assert got['y'].values[1] == 500.            # failed fits keep their parameters


def test_cluster_larger_than_every_capacity_fails_loudly():
    import pandas as pd
    rng = np.random.RandomState(0)
    n = 300                                   # above CTK_MAX_BIG_FEATURES
    f = pd.DataFrame(dict(y=60 + rng.uniform(-30, 30, n), x=60 + rng.uniform(-30, 30, n),
                          signal=50., size=2.))
    img = rng.randint(0, 50, (128, 128)).astype(np.uint8)
    got, plan = emul_backend.refine_leastsq(f, img, 9)
    assert plan.n_clusters == 1 and np.isnan(got['cost'].values).all()
    assert (got['y'].values == f['y'].values).all()


def test_large_cluster_path():
    """More than 32 overlapping features percolate into one cluster (TestMultiple,
    tests/test_refine.py:913-922): the large-cluster kernels (global workspace, multi-word pixel
    masks) take it; every feature ends within 0.1 px of the truth."""
    import pandas as pd
    from clustertracking_b200 import artificial
    rng = np.random.RandomState(7)
    pos = []
    while len(pos) < 70:
        cand = rng.uniform(21, 200 - 21, 2)
        if all(np.linalg.norm(cand - p) >= 15 for p in pos):
            pos.append(cand)
    pos = np.array(pos)
    image = artificial.draw_features((200, 200), pos, 5.25, 200.)
    f0 = pd.DataFrame(pos + rng.random_sample(pos.shape) * 7, columns=['y', 'x'])
    f0['signal'] = 200.
    f0['size'] = 5.25
    got, plan = emul_backend.refine_leastsq(f0, image, 21, separation=24)
    assert plan.cluster_sizes().max() > 32
    assert np.isfinite(got['cost'].values).all()
    assert np.abs(got[['y', 'x']].values - pos).max() < 0.1


def test_edge_clipped_and_overlapping_masks():
    """Features at the image border and coincident features still give the oracle's answer."""
    from clustertracking_b200 import artificial
    from oracle import cluster_oracle
    import pandas as pd
    rng = np.random.default_rng(5)
    pos = np.array([[2.3, 3.1], [2.9, 37.2], [36.8, 20.4], [20.2, 20.9], [23.1, 24.8]])
    img = artificial.draw_features((40, 40), pos, 2.0, 150., noise=4, rng=rng)
    f0 = pd.DataFrame(pos + rng.uniform(-0.4, 0.4, pos.shape), columns=['y', 'x'])
    f0['signal'] = 120.
    f0['size'] = 2.0
    got, _ = emul_backend.refine_leastsq(f0.copy(), img, 9, precision='float64')
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        want = cluster_oracle.refine_leastsq(f0.copy(), img, 9, tol=1e-12)
    assert_array_equal(got['cluster'].values, want['cluster'].values)
    assert_allclose(got[['y', 'x']].values, want[['y', 'x']].values, atol=1e-5)
    assert_allclose(got['cost'].values, want['cost'].values, rtol=1e-6)


def dense_cluster_case(ny, nx, pitch=4.0, seed=3):
    """A lattice of ny x nx narrow features ``pitch`` apart, fitted with diameter 11: every mask
    overlaps many others (more overlapping pairs than the 4 n lists the shared-memory classes
    provision even with rigorous capacities)."""
    import pandas as pd
    from clustertracking_b200 import artificial
    rng = np.random.default_rng(seed)
    gy, gx = np.meshgrid(np.arange(ny) * pitch, np.arange(nx) * pitch, indexing='ij')
    pos = np.stack([gy.ravel(), gx.ravel()], axis=1) + 20. + rng.uniform(-0.3, 0.3, (ny * nx, 2))
    shape = (int(pos[:, 0].max()) + 21, int(pos[:, 1].max()) + 21)
    image = artificial.draw_features(shape, pos, 1.5, 180., noise=2, rng=rng)
    f0 = pd.DataFrame(pos + rng.uniform(-0.2, 0.2, pos.shape), columns=['y', 'x'])
    f0['signal'] = 150.
    f0['size'] = 1.5
    return image, f0, pos


@pytest.mark.parametrize("ny,nx", [(3, 4), (4, 4), (4, 5), (4, 7), (5, 6)])
def test_dense_cluster_is_fitted_not_dropped(ny, nx):
    """12-30 features with more than 4 n overlapping pairs overflowed even the rigorous
    capacities and came back with cost = NaN where the reference fits them; they now take the
    large-cluster path."""
    from clustertracking_b200 import _lib, refine
    image, f0, pos = dense_cluster_case(ny, nx)
    plan = refine.prepare(f0.copy(), image, 11, precision='float64')
    assert plan.n_clusters == 1 and plan.cluster_sizes()[0] == ny * nx
    result = emul_backend.execute(plan)
    assert result.status[0] == 0, _lib.STATUS_NAMES.get(int(result.status[0]))
    got = refine.finalize(plan, result)
    assert np.isfinite(got['cost'].values).all()
    assert np.abs(got[['y', 'x']].values - pos).max() < 0.1
