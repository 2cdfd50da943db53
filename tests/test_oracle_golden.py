"""Pin ``oracle/cluster_oracle.py`` to the reference.

The fixtures hold outputs of the UNMODIFIED reference run in the build container
(``oracle/make_golden.py``).  The oracle calls the same scipy SLSQP with the same arithmetic, so the
expected agreement is at rounding level; tolerances below are stated per test.  Also restates the
known-answer tests the reference holds for the path (tests/test_fitfunc.py:65-83,
tests/test_mask.py:11-129, tests/test_find.py:34-125).
"""
import json
import warnings

import numpy as np
import pytest
from numpy.testing import assert_allclose, assert_array_equal

import golden_io
from oracle import cluster_oracle as oracle


# ---- residual / jacobian (fitfunc.py:421-489) -----------------------------------------------
@pytest.mark.parametrize("name", golden_io.names("fitfunc_"))
def test_objective_matches_reference(name):
    d = golden_io.load(name)
    spec = oracle.ModelSpec(str(d["family"]), int(d["ndim"]), bool(d["isotropic"]),
                            json.loads(str(d["param_mode"])))
    assert spec.params == [str(p) for p in d["param_names"]]
    assert_array_equal(spec.modes, d["modes"])
    fun, grad = spec.objective(d["image"], d["mesh"], d["masks"], d["params"], float(d["norm"]))
    vect = oracle.pack_vector(d["params"], spec.modes, np.mean)
    assert_array_equal(vect, d["vect"])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        assert_allclose(fun(vect), d["fun"], rtol=1e-13)        # same expressions: rounding level
        if grad is None:
            assert d["jac"].size == 0
        else:
            assert_allclose(grad(vect), d["jac"], rtol=1e-12, atol=1e-12)


def test_2d_gauss_known_answer():
    """tests/test_fitfunc.py:65-83: the residual equals a hand-written gaussian to 1e-7."""
    rng = np.random.RandomState(0)
    spec = oracle.ModelSpec('gauss', 2, True)
    params = np.array([[5, 200, 4, 5, 6]], dtype=float)
    image = rng.random_sample(100) * 200
    mesh = rng.random_sample((2, 100)) * 10
    masks = np.ones((1, 100), dtype=bool)
    fun, _ = spec.objective(image, mesh, masks, params)
    y, x = mesh
    model = 5 + 200 * np.exp(-((y - 4) ** 2 / 36. + (x - 5) ** 2 / 36.))
    assert_allclose(fun(oracle.pack_vector(params, spec.modes, np.mean)),
                    np.sum((image - model) ** 2) / 100, atol=1e-7)


@pytest.mark.parametrize("family,ndim,iso,n,custom", [
    ('gauss', 2, True, 1, {}), ('gauss', 2, False, 1, {}), ('gauss', 3, True, 1, {}),
    ('gauss', 3, False, 1, {}), ('ring', 2, True, 1, {}), ('gauss', 2, True, 2, {}),
    ('gauss', 2, True, 2, dict(signal='cluster'))])
def test_gradient_vs_finite_differences(family, ndim, iso, n, custom):
    """tests/test_fitfunc.py:29-42 (compare_jacobian): epsilon 1e-7, rtol 0.01, atol 0.001."""
    rng = np.random.RandomState(4)
    mode = {p: 'var' for p in oracle.ModelSpec(family, ndim, iso).params}
    mode['background'] = 'cluster'
    mode.update(custom)
    spec = oracle.ModelSpec(family, ndim, iso, mode)
    params = rng.random_sample((n, len(spec.params))) * 10
    image = rng.random_sample(100) * 200
    mesh = rng.random_sample((ndim, 100)) * 10
    masks = rng.random_sample((n, 100)) > 0.5
    fun, grad = spec.objective(image, mesh, masks, params)
    v = oracle.pack_vector(params, spec.modes, np.mean)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        f0 = fun(v)
        fd = np.array([(fun(v + 1e-7 * np.eye(len(v))[k]) - f0) / 1e-7 for k in range(len(v))])
        assert_allclose(grad(v), fd, rtol=0.01, atol=0.001)


# ---- pixel sets (masks.py:30-68, refine.py:28-58) -------------------------------------------
def test_bounding_boxes_match_reference():
    d = golden_io.load("pixels_boxes")
    for i in range(int(d["n_cases"])):
        box = oracle.bounding_box(d["coords_%d" % i], tuple(d["shape_%d" % i]),
                                  tuple(np.atleast_1d(d["radius_%d" % i])) if d["radius_%d" % i].ndim
                                  else int(d["radius_%d" % i]))
        if box is None:
            assert (d["lo_%d" % i] == -1).all()
        else:
            assert_array_equal(box[0], d["lo_%d" % i])
            assert_array_equal(box[1], d["hi_%d" % i])


def test_slicing_known_answers():
    """tests/test_mask.py:11-129, exact integers."""
    for ndim in (2, 3):
        shape = (9,) * ndim
        for r in range(1, 5):
            lo, hi = oracle.bounding_box([[4] * ndim], shape, r)
            assert lo == [4 - r] * ndim and hi == [4 + r + 1] * ndim
            lo, hi = oracle.bounding_box([[0] + [4] * (ndim - 1)], shape, r)
            assert [b - a for a, b in zip(lo, hi)] == [r + 1] + [2 * r + 1] * (ndim - 1)
            lo, hi = oracle.bounding_box([[0] * ndim], shape, r)
            assert [b - a for a, b in zip(lo, hi)] == [r + 1] * ndim
        for r in range(2, 5):
            lo, hi = oracle.bounding_box([[-1] + [4] * (ndim - 1)], shape, r)
            assert [b - a for a, b in zip(lo, hi)] == [r] + [2 * r + 1] * (ndim - 1)
            assert oracle.bounding_box([[-10, 20, 30][:ndim]], shape, r) is None
    lo, hi = oracle.bounding_box([[4, 2], [4, 6]], (9, 9), 2)
    assert (lo, hi) == ([2, 0], [7, 9])
    lo, hi = oracle.bounding_box([[2, 4], [6, 4], [-10, 20]], (9, 9), 2)
    assert (lo, hi) == ([0, 2], [9, 7])
    lo, hi = oracle.bounding_box([[4, 2, 6], [4, 6, 2]], (9, 9, 9), 2)
    assert (lo, hi) == ([2, 0, 0], [7, 9, 9])


@pytest.mark.parametrize("name", [n for n in golden_io.names("pixels_") if n != "pixels_boxes"])
def test_pixel_sets_match_reference(name):
    d = golden_io.load(name)
    radius = tuple(int(r) for r in np.atleast_1d(d["radius"]))
    radius = radius if len(radius) > 1 else radius[0]
    values, mesh, masks = oracle.cluster_pixels(d["coords"], d["image"], radius)
    assert_array_equal(values, d["values"])
    assert_array_equal(mesh, d["mesh"])
    assert_array_equal(masks, d["masks"])


@pytest.mark.parametrize("name", golden_io.names("lowpass_pixels_"))
def test_lowpass_pixel_sets_match_reference(name):
    """``noise_size`` / ``threshold``: the filter runs on the cluster's box, zero padded at the BOX
    edge (refine.py:36-40, preprocessing.py:12-49)."""
    d = golden_io.load(name)
    radius = tuple(int(r) for r in np.atleast_1d(d["radius"]))
    radius = radius if len(radius) > 1 else radius[0]
    noise = tuple(float(v) for v in np.atleast_1d(d["noise_size"]))
    noise = noise if len(noise) > 1 else noise[0]
    threshold = None if np.isnan(d["threshold"]) else float(d["threshold"])
    values, mesh, masks = oracle.cluster_pixels(d["coords"], d["image"], radius, noise, threshold)
    assert_array_equal(values, d["values"])
    assert_array_equal(mesh, d["mesh"])
    assert_array_equal(masks, d["masks"])


# ---- clustering (find.py:12-163) -------------------------------------------------------------
@pytest.mark.parametrize("name", golden_io.names("clusters_"))
def test_find_clusters_matches_reference(name):
    d = golden_io.load(name)
    f = golden_io.frame(d, "in_")
    want = golden_io.frame(d, "out_")
    sep = d["separation"]
    sep = tuple(sep) if sep.ndim else float(sep)
    got = oracle.find_clusters(f, sep)
    assert list(got.columns) == list(want.columns)
    assert_array_equal(got.index.values, want.index.values)
    for col in want.columns:
        assert_array_equal(got[col].values, want[col].values)


def test_find_clusters_known_answers():
    """tests/test_find.py:34-125 style: a line of touching points is one cluster, far points not."""
    import pandas as pd
    line = pd.DataFrame(dict(y=np.zeros(10), x=np.arange(10) * 0.9))
    out = oracle.find_clusters(line, 1.0)
    assert out['cluster'].nunique() == 1 and (out['cluster_size'] == 10).all()
    assert 'frame' not in line                       # find.py:151-161: temporary column removed
    far = pd.DataFrame(dict(y=np.zeros(10), x=np.arange(10) * 1.1))
    out = oracle.find_clusters(far, 1.0)
    assert out['cluster'].nunique() == 10 and (out['cluster_size'] == 1).all()


# ---- end to end (refine.py:82-452) ------------------------------------------------------------
@pytest.mark.parametrize("name", golden_io.names("refine_"))
def test_refine_matches_reference(name):
    """Same SLSQP, same objective: positions within 1e-6 px, other columns 1e-6 relative.
    (Iterates are rounding-sensitive, so bit equality is not demanded.)"""
    d = golden_io.load(name)
    f0, reader, diameter, kwargs = golden_io.refine_inputs(d, oracle)
    want = golden_io.frame(d, "ref_")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        got = oracle.refine_leastsq(f0, reader, diameter, **kwargs)
    assert sorted(got.columns) == sorted(want.columns)
    assert_array_equal(got.index.values, want.index.values)
    assert_array_equal(got['cluster'].values, want['cluster'].values)
    assert_array_equal(got['cluster_size'].values, want['cluster_size'].values)
    assert_array_equal(np.isnan(got['cost'].values), np.isnan(want['cost'].values))
    # the sort-based 2D tetramer constraint makes SLSQP's iterates rounding-sensitive
    tol = 1e-3 if name == "refine_tetramer2d_constrained" else 1e-6
    for col in want.columns:
        if col in ('cluster', 'cluster_size', 'frame'):
            continue
        assert_allclose(got[col].values, want[col].values, rtol=tol, atol=tol, err_msg=col)


# ---- feature finding (find.py:166-277) -----------------------------------------------------------
@pytest.mark.parametrize("name", golden_io.names("find_"))
def test_grey_dilation_oracle_matches_reference(name):
    import json
    from oracle import find_oracle
    d = golden_io.load(name)
    kwargs = json.loads(str(d["kwargs"]))
    for key in ("separation", "margin"):
        if isinstance(kwargs.get(key), list):
            kwargs[key] = tuple(kwargs[key])
    got = find_oracle.grey_dilation(d["image"], **kwargs)
    assert_array_equal(np.asarray(got).reshape(-1, d["image"].ndim), d["pos"])
