"""GPU parity tests (run on the B200 box with ``-m gpu``): the CUDA path, called through the C ABI by
``clustertracking_b200.refine_leastsq``, against (a) the reference's own outputs stored in
tests/golden and (b) the CPU oracle on the same seeded inputs.

Tolerances (BASELINE.json north_star): 1e-3 px in position, 1e-3 relative in signal and size,
identical cluster membership, feature order and failure set.  The reference run with tol=1e-12 is
matched much tighter (1e-5 px, float32 pixel arithmetic) -- that residue is ours, the rest of the
1e-3 budget is the reference's own SLSQP termination noise.
"""
import warnings

import numpy as np
import pytest
from numpy.testing import assert_allclose, assert_array_equal

import golden_io

pytestmark = pytest.mark.gpu

POS_TOL = 1e-3          # px, vs reference at its default tol=1e-6
REL_TOL = 1e-3          # signal, size (relative)
POS_TOL_TIGHT = 2e-5    # px, vs reference at tol=1e-12
REL_TOL_TIGHT = 2e-5


def _compare(got, want, pos_tol, rel_tol):
    assert sorted(got.columns) == sorted(want.columns)
    assert_array_equal(got.index.values, want.index.values)
    assert_array_equal(got['cluster'].values, want['cluster'].values)
    assert_array_equal(got['cluster_size'].values, want['cluster_size'].values)
    assert_array_equal(np.isnan(got['cost'].values), np.isnan(want['cost'].values))
    for col in want.columns:
        if col in ('cluster', 'cluster_size', 'frame'):
            assert_array_equal(got[col].values, want[col].values)
        elif col in ('z', 'y', 'x'):
            assert_allclose(got[col].values, want[col].values, rtol=0, atol=pos_tol, err_msg=col)
        elif col == 'cost':
            assert_allclose(got[col].values, want[col].values, rtol=rel_tol, atol=1e-6, err_msg=col)
        elif col == 'background':
            # absolute, in units of the signal scale (background is often exactly 0)
            assert_allclose(got[col].values, want[col].values, rtol=0,
                            atol=rel_tol * max(1., float(np.nanmax(np.abs(want['signal'].values)))),
                            err_msg=col)
        else:
            assert_allclose(got[col].values, want[col].values, rtol=rel_tol, atol=1e-9, err_msg=col)


# The 2D tetramer constraint is built on a sort of the six pair distances (constraints.py:102-114),
# which SLSQP differentiates numerically; the reference stops at slightly sub-optimal points there
# (and fails two of four clusters at tol=1e-12), so that fixture gets its own test below.
# Ring and disc have piecewise-smooth objectives with many shallow basins (the "safe" pixels of
# fitfunc.py:20-26 enter and leave the sums; the disc switches branch at disc_size = 0); outside the
# reference's default modes in 2D the two solvers need not end in the same basin.  Those fixtures
# are judged basin-aware (check_basins below); everything else must meet the strict tolerances.
BASIN = [n for n in golden_io.names("refine_ring") + golden_io.names("refine_disc")
         if n not in ("refine_ring2d", "refine_ring2d_sizevar", "refine_disc2d")]
STRICT = [n for n in golden_io.names("refine_")
          if n != "refine_tetramer2d_constrained" and n not in BASIN]


def check_basins(got, d):
    """Every cluster is within 1e-3 px of the reference at its default tolerance, or ends at a cost
    not above the reference's at BOTH tolerances (tests/fuzz_cases.judge).  Clusters whose re-mask
    loop cycles in the reference (meta['unsettled']) are no target."""
    import fuzz_cases
    ref, tight = golden_io.frame(d, "ref_"), golden_io.frame(d, "tight_")
    assert_array_equal(got.index.values, ref.index.values)
    assert_array_equal(got['cluster_size'].values, ref['cluster_size'].values)
    assert not np.isnan(got['cost'].values).any()
    cols = [c for c in ('z', 'y', 'x') if c in ref.columns]
    verdict = fuzz_cases.judge(got, ref, tight, cols, unsettled=golden_io.meta(d).get('unsettled', ()))
    assert verdict['ok'], verdict
    return verdict


def check_tetramer2d(got, d, f0, image, pos_tol):
    """The 2D tetramer constraint (four shortest pair distances = bond: a rhombus) is a sorted,
    non-smooth function that SLSQP differentiates numerically; the reference stops 2e-3 .. 1e-2 px
    short of the minimum.  Proof, not assertion: an independent minimiser on the exactly
    parametrised feasible set (tests/tetramer_check.py) lands on the DEVICE answer to ``pos_tol``
    and at the device's cost, and the reference's cost is not lower.  The 2e-2 px left against the
    reference's stored answer is therefore the reference's residue."""
    import tetramer_check
    want = golden_io.frame(d, "ref_")
    assert_array_equal(got['cluster'].values, want['cluster'].values)
    assert not np.isnan(got['cost'].values).any()
    assert (got['cost'].values <= want['cost'].values * (1 + 1e-6)).all()
    assert_allclose(got[['y', 'x']].values, want[['y', 'x']].values, rtol=0, atol=2e-2)
    for _, g in got.groupby('cluster'):
        p = g[['y', 'x']].values
        d2 = sorted(np.sum(((p[a] - p[b]) / 8.) ** 2) for a in range(4) for b in range(a + 1, 4))
        assert max(abs(1 - x) for x in d2[:4]) < 1e-6
    return tetramer_check.check_against_independent_minimum(got, d, f0, image, pos_tol)


@pytest.mark.parametrize("precision", ["float32", "float64"])
def test_cuda_tetramer2d(precision):
    import clustertracking_b200 as ctb
    d = golden_io.load("refine_tetramer2d_constrained")
    f0, reader, diameter, kwargs = golden_io.refine_inputs(d, ctb.constraints)
    got = ctb.refine_leastsq(f0.copy(), reader, diameter, precision=precision, **kwargs)
    check_tetramer2d(got, d, f0, np.asarray(reader), 1e-6 if precision == 'float64' else 2e-5)


@pytest.mark.parametrize("precision", ["float32", "float64"])
@pytest.mark.parametrize("name", STRICT)
def test_cuda_matches_reference_golden(name, precision):
    import clustertracking_b200 as ctb
    d = golden_io.load(name)
    f0, reader, diameter, kwargs = golden_io.refine_inputs(d, ctb.constraints)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        got = ctb.refine_leastsq(f0, reader, diameter, precision=precision, **kwargs)
    _compare(got, golden_io.frame(d, "ref_"), POS_TOL, REL_TOL)
    _compare(got, golden_io.frame(d, "tight_"), POS_TOL_TIGHT, REL_TOL_TIGHT)


@pytest.mark.parametrize("precision", ["float32", "float64"])
@pytest.mark.parametrize("name", BASIN)
def test_cuda_ring_disc_basins(name, precision):
    """Ring / disc x {2D anisotropic, 3D, 3D anisotropic} x {default modes, shape parameter free}: the
    combinations the reference's own suite runs (tests/test_refine.py:768-881)."""
    import clustertracking_b200 as ctb
    d = golden_io.load(name)
    f0, reader, diameter, kwargs = golden_io.refine_inputs(d, ctb.constraints)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        got = ctb.refine_leastsq(f0, reader, diameter, precision=precision, **kwargs)
    check_basins(got, d)
