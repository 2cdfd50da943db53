"""CPU tests of the host side: clustering, parameter bookkeeping, bounds, constraint descriptors,
packing, argument errors, and that the C-ABI library loads and exports every declared symbol
(no compute calls -- there is no GPU here)."""
import ctypes
import os
import re
import warnings

import numpy as np
import pandas as pd
import pytest
from numpy.testing import assert_allclose, assert_array_equal

import golden_io
import clustertracking_b200 as ctb
from clustertracking_b200 import _lib, refine
from oracle import cluster_oracle as oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ---- library --------------------------------------------------------------------------------------
def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "ctk.h")).read()
    declared = set(re.findall(r"\b(ctk_[a-z_]+)\s*\(", header))
    assert {"ctk_refine_batch", "ctk_frame_max", "ctk_label_clusters", "ctk_version"} <= declared
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert _lib.load().ctk_version() >= 100


def test_problem_struct_matches_header(tmp_path):
    """The ctypes mirror of ctk_problem_t has the size and the field offsets the C compiler gives
    the struct of include/ctk.h."""
    import subprocess
    names = [name for name, _ in _lib.Problem._fields_]
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "ctk.h"\nint main(void) {\n'
                   '  printf("%zu\\n", sizeof(ctk_problem_t));\n'
                   + "".join('  printf("%%zu\\n", offsetof(ctk_problem_t, %s));\n' % n for n in names)
                   + '  return 0;\n}\n')
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    out = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    assert out[0] == ctypes.sizeof(_lib.Problem)
    assert out[1:] == [getattr(_lib.Problem, n).offset for n in names]


def test_shared_bytes_query_needs_no_gpu():
    plan = _tiny_plan()
    lib = _lib.load()
    small = lib.ctk_refine_shared_bytes(ctypes.byref(plan.problem), 2)
    big = lib.ctk_refine_shared_bytes(ctypes.byref(plan.problem), 32)
    assert 0 < small < big
    assert lib.ctk_refine_shared_bytes(ctypes.byref(plan.problem), 33) == 0


def test_label_clusters_rule():
    # (0,1) then (2,1): label of the first argument's cluster survives (find.py:41-48)
    labels, sizes = _lib.label_clusters(np.array([[0, 1], [2, 1], [4, 5]]), 6)
    assert_array_equal(labels, [2, 2, 2, 3, 4, 4])
    assert_array_equal(sizes, [3, 3, 3, 1, 2, 2])
    labels, sizes = _lib.label_clusters(np.zeros((0, 2), np.int64), 3)
    assert_array_equal(labels, [0, 1, 2])


# ---- clustering ------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", golden_io.names("clusters_"))
def test_find_clusters_matches_reference(name):
    d = golden_io.load(name)
    f = golden_io.frame(d, "in_")
    want = golden_io.frame(d, "out_")
    sep = d["separation"]
    got = ctb.find_clusters(f, tuple(sep) if sep.ndim else float(sep))
    assert list(got.columns) == list(want.columns)
    assert_array_equal(got.index.values, want.index.values)
    for col in want.columns:
        assert_array_equal(got[col].values, want[col].values)
        assert got[col].dtype == want[col].dtype


def test_find_clusters_without_frame_column():
    f = pd.DataFrame(dict(y=np.zeros(10), x=np.arange(10) * 0.9))
    out = ctb.find_clusters(f, 1.0)
    assert out['cluster'].nunique() == 1 and (out['cluster_size'] == 10).all()
    assert 'frame' in out and 'frame' not in f
    far = pd.DataFrame(dict(y=np.zeros(10), x=np.arange(10) * 1.1))
    assert ctb.find_clusters(far, 1.0)['cluster'].nunique() == 10


# ---- FitFunctions ---------------------------------------------------------------------------------
@pytest.mark.parametrize("family,ndim,iso,mode", [
    ('gauss', 2, True, None), ('gauss', 2, False, dict(size='var')),
    ('gauss', 3, False, dict(size='cluster', signal='const')), ('ring', 2, True, dict(thickness='var')),
    ('disc', 3, True, dict(pos='const', size='var')), ('gauss', 3, True, dict(background='var'))])
def test_fitfunctions_bookkeeping_matches_oracle(family, ndim, iso, mode):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ours = ctb.FitFunctions(family, ndim, iso, mode)
        ref = oracle.ModelSpec(family, ndim, iso, mode)
    assert ours.params == ref.params
    assert ours.modes == ref.modes
    assert ours.default == ref.default


@pytest.mark.parametrize("bounds", [
    None, dict(signal=(20, 2000), size=(.9, 9)), dict(pos_diff=2.0, signal_rel_diff=0.5),
    dict(x=(2, 60), x_diff=(5, 3), size_rel_diff=(0.2, 0.4), background=(1, 5))])
def test_feature_bounds_match_oracle(bounds):
    rng = np.random.RandomState(3)
    for ndim, iso in ((2, True), (3, False)):
        ours = ctb.FitFunctions('gauss', ndim, iso)
        ref = oracle.ModelSpec('gauss', ndim, iso)
        radius = (5,) * ndim
        if bounds and 'size' in bounds and not iso:
            pass
        params = rng.uniform(0.5, 50, (7, len(ours.params)))
        lo, hi = ours.feature_bounds(ours.validate_bounds(bounds, radius), params)
        lo_r, hi_r = ref.feature_bounds(ref.bounds_tables(bounds, radius), params)
        assert_array_equal(lo, lo_r)
        assert_array_equal(hi, hi_r)


def test_unsupported_options_raise():
    f = pd.DataFrame(dict(y=[10.], x=[10.], signal=[100.], size=[2.]))
    img = np.zeros((32, 32), np.uint8)
    with pytest.raises(NotImplementedError):
        refine.prepare(f.copy(), img, 9, param_mode=dict(signal='global'))
    with pytest.raises(NotImplementedError):
        refine.prepare(f.copy(), img, 9, fit_function=dict(params=[], func=None))
    with pytest.raises(NotImplementedError):
        refine.prepare(f.copy(), img, 9, fit_function='inv_series_3')
    with pytest.raises(NotImplementedError):
        refine.prepare(f.copy(), img, 9, noise_size=5)      # taps wider than the device table
    plan = refine.prepare(f.copy(), img, 9, noise_size=(1, 0), threshold=3)
    assert plan.problem.lowpass == 1 and list(plan.problem.lowpass_half)[:2] == [4, -1]
    assert plan.problem.lowpass_threshold == 3. and plan.problem.lowpass_sigma[0] == 1.
    with pytest.raises(NotImplementedError):
        refine.prepare(f.copy(), img, 9, compute_error=True)
    with pytest.raises(NotImplementedError):
        refine.prepare(f.copy(), img, 9, constraints=[dict(type='eq', fun=lambda x: x)])
    with pytest.raises(ValueError):
        refine.prepare(f.copy(), [img], 9)                  # refine.py:260-262
    with pytest.raises(AssertionError):
        refine.prepare(f.copy(), np.zeros((4, 32, 32), np.uint8), 9)   # refine.py:283


def test_constraint_descriptors():
    (d,) = ctb.constraints.dimer(8.0, 2)
    assert d['type'] == 'eq' and d['cluster_size'] == 2
    x = np.zeros((1, 2, 5))
    x[0, 1, 2:4] = (8.0, 0.)
    assert_allclose(d['fun'](x, *d['args']), 0.)
    (o,) = oracle.dimer(8.0, 2)
    assert_allclose(o['fun'](x, *o['args']), d['fun'](x, *d['args']))
    (t,) = ctb.constraints.trimer((6., 8.), 2)
    x = np.random.RandomState(0).uniform(0, 10, (1, 3, 6))
    (ot,) = oracle.trimer((6., 8.), 2)
    assert_allclose(t['fun'](x, *t['args']), ot['fun'](x, *ot['args']))
    parsed = ctb.constraints.parse(ctb.constraints.dimer(8.) + ctb.constraints.trimer(7.), 2)
    assert_array_equal(parsed['dimer'], [8., 8.])
    assert_array_equal(parsed['trimer'], [7., 7.])


# ---- packing ---------------------------------------------------------------------------------------
def _tiny_plan():
    d = golden_io.load("refine_gauss2d_video")
    f0, reader, diameter, kwargs = golden_io.refine_inputs(d, ctb.constraints)
    return refine.prepare(f0, reader, diameter, **kwargs)


def test_plan_groups_follow_reference_order():
    d = golden_io.load("refine_gauss2d_video")
    f0, reader, diameter, kwargs = golden_io.refine_inputs(d, ctb.constraints)
    plan = refine.prepare(f0, reader, diameter, **kwargs)
    f = plan.f
    want_groups = [list(g.index) for _, g in f.groupby(['frame', 'cluster'])]
    got_groups = [list(f.index[plan.order[a:b]])
                  for a, b in zip(plan.cluster_offset[:-1], plan.cluster_offset[1:])]
    assert got_groups == want_groups
    assert_array_equal(plan.params_in, f[plan.ff.params].values[plan.order])
    assert plan.frame_numbers == [0, 1, 2]
    assert plan.cluster_frame.dtype == np.int32 and (np.diff(plan.cluster_frame) >= 0).all()
    assert plan.problem.n_params == 5 and list(plan.problem.modes)[:5] == [3, 1, 1, 1, 0]


def test_finalize_failure_semantics():
    plan = _tiny_plan()
    res = refine.Result(plan)
    res.params_out = plan.params_in + 1.0
    res.status[:] = 0
    res.cost[:] = 0.01
    res.status[1] = 3                      # one failed cluster: unchanged parameters, NaN cost
    before = plan.f[plan.ff.params].values.copy()
    out = refine.finalize(plan, res)
    a, b = plan.cluster_offset[1], plan.cluster_offset[2]
    failed_rows = plan.order[a:b]
    assert np.isnan(out['cost'].values[failed_rows]).all()
    assert_array_equal(out[plan.ff.params].values[failed_rows], before[failed_rows])
    ok = np.setdiff1d(np.arange(len(out)), failed_rows)
    assert_allclose(out[plan.ff.params].values[ok], before[ok] + 1.0)
    assert (out['cost'].values[ok] == 0.01).all()
    assert list(out.columns)[-1] == 'cost'


def test_frame_column_is_added_in_place_like_the_reference():
    f = pd.DataFrame(dict(y=[10., 20.], x=[10., 20.], signal=[100., 100.], size=[2., 2.]))
    img = np.zeros((32, 32), np.uint8)
    plan = refine.prepare(f, img, 9)
    assert 'frame' in f and (f['frame'] == 0).all()          # refine.py:279
    assert list(plan.f.columns[:7]) == ['y', 'x', 'signal', 'size', 'frame', 'cluster', 'cluster_size']


# ---- the chunked pipeline of refine_leastsq (host logic; the solver is the one-lane emulation) -----
class _NoFrames(object):
    h2d_bytes = launches = 0

    def __init__(self, info, device=None):
        pass

    def upload_async(self):
        return self

    def close(self):
        pass


class _Done(object):
    def __init__(self, result):
        self._result = result

    def result(self):
        return self._result


def _emulated_launch(plan, frames, out_params, out_cost, out_status, stream=None):
    import emul_backend
    result = emul_backend.execute(plan)
    out_params[...], out_cost[...], out_status[...] = result.params_out, result.cost, result.status
    result.params_out, result.cost, result.status = out_params, out_cost, out_status
    result.session = type("S", (), dict(h2d_bytes=0, d2h_bytes=0, launches=0))()
    return _Done(result)


@pytest.mark.parametrize("shuffle", [False, True])
def test_pipelined_chunks_equal_whole_table(monkeypatch, shuffle):
    import torch
    import emul_backend
    from clustertracking_b200 import artificial
    stack, rows = [], []
    for t in range(6):
        frame, f0, _ = artificial.clustered_frame((96, 96), pitch=44, seed=70 + t)
        f0['frame'] = 2 * t + 1
        f0['tag'] = np.arange(len(f0)) + 100 * t                # a column the fit does not touch
        stack.append(frame)
        rows.append(f0)
    reader = {2 * t + 1: stack[t] for t in range(6)}

    class Reader(dict):
        frame_shape = (96, 96)

    reader = Reader(reader)
    f = pd.concat(rows, ignore_index=True)
    if shuffle:
        f = f.sample(frac=1., random_state=3)
    want, _ = emul_backend.refine_leastsq(f.copy(), reader, 11, param_val=dict(size=2.75))
    monkeypatch.setattr(refine, "FrameSet", _NoFrames)
    monkeypatch.setattr(refine, "_pinned_buffer",
                        lambda torch_, key, nbytes: torch.empty(max(nbytes, 1), dtype=torch.uint8))
    monkeypatch.setattr(refine, "launch_cuda", _emulated_launch)
    monkeypatch.setattr(refine, "_CHUNK_ROWS", 8)
    got = refine.refine_leastsq(f.copy(), reader, 11, param_val=dict(size=2.75))
    assert refine.LAST_CALL["chunks"] > 2
    assert list(got.columns) == list(want.columns)
    assert_array_equal(got.index.values, want.index.values)
    for col in want.columns:
        assert got[col].dtype == want[col].dtype, col
        assert_array_equal(got[col].values, want[col].values, err_msg=col)
    got.loc[got.index[0], 'x'] = 1.0                            # the result is writable


def test_schedule_groups_by_class_expensive_first():
    sizes = np.array([1, 5, 2, 6, 300, 33, 2, 4, 40, 1])
    offset = np.concatenate(([0], np.cumsum(sizes))).astype(np.int32)
    caps = np.array(refine._BINS, dtype=np.int32)
    target = np.arange(len(caps), dtype=np.int32)
    target[list(caps).index(4)] = list(caps).index(64)        # class 4 does not fit: runs in class 64
    ids, counts, not_run = _lib.schedule(offset, caps, target)
    assert list(not_run) == [4]                               # 300 features: no class
    by_cap = dict(zip(caps.tolist(), counts.tolist()))
    assert by_cap[1] == 2 and by_cap[2] == 2 and by_cap[6] == 2 and by_cap[4] == 0 and by_cap[64] == 3
    assert list(ids) == [0, 9, 2, 6, 3, 1, 8, 5, 7]           # per class: larger clusters first


# ---- feature finding, host part --------------------------------------------------------------------
def test_where_close_matches_oracle():
    from clustertracking_b200 import find
    from oracle import find_oracle
    rng = np.random.RandomState(8)
    for ndim, sep in ((2, 9), (2, (8, 12)), (3, (5, 9, 9))):
        pos = rng.randint(0, 120, (400, ndim))
        inten = rng.randint(50, 60, 400)                    # many ties
        assert_array_equal(find.where_close(pos, sep, inten), find_oracle.where_close(pos, sep, inten))
        assert_array_equal(find.where_close(pos, sep), find_oracle.where_close(pos, sep))
        assert_array_equal(find.drop_close(pos, sep, inten),
                           np.delete(pos, find_oracle.where_close(pos, sep, inten), axis=0))
    assert find.where_close(np.empty((0, 2)), 5) == []


def test_query_pairs_within_is_the_scipy_set():
    from scipy.spatial import cKDTree
    rng = np.random.RandomState(9)
    for r in (1 - 1e-7, 0.5, 1.0, 2.0):
        data = rng.randint(0, 60, (500, 2)) / 7.
        want = cKDTree(data, 30).query_pairs(r, output_type='ndarray')
        got = _lib.query_pairs(data, r)
        assert set(map(tuple, got.tolist())) == set(map(tuple, want.tolist()))


def test_pipelined_path_failures_and_tiny_inputs(monkeypatch, caplog):
    """Failure bookkeeping of the pipelined path (refine.py:408-418 semantics: cost NaN, parameters
    untouched, a warning) and inputs of one or two rows."""
    import logging
    import torch
    monkeypatch.setattr(refine, "FrameSet", _NoFrames)
    monkeypatch.setattr(refine, "_pinned_buffer",
                        lambda torch_, key, nbytes: torch.empty(max(nbytes, 1), dtype=torch.uint8))
    monkeypatch.setattr(refine, "launch_cuda", _emulated_launch)
    img = np.zeros((40, 40), np.uint8)
    img[18:23, 18:23] = 100
    f = pd.DataFrame(dict(y=[20., 500., 20.], x=[20., 500., 32.], signal=[100., 100., np.nan], size=2.))
    with caplog.at_level(logging.WARNING, logger="clustertracking_b200.refine"):
        out = refine.refine_leastsq(f.copy(), img, 9, separation=4)
    assert np.isfinite(out['cost'].values[0]) and np.isnan(out['cost'].values[1:]).all()
    assert out['y'].values[1] == 500. and np.isnan(out['signal'].values[2])
    assert sum("RefineException" in r.message for r in caplog.records) == 2
    assert list(out['cluster']) == [0, 1, 2] and list(out['cluster_size']) == [1, 1, 1]
    one = refine.refine_leastsq(f.iloc[:1].copy(), img, 9)
    assert len(one) == 1 and np.isfinite(one['cost'].values[0])
    two_frames = pd.DataFrame(dict(y=[20., 20.], x=[20., 20.], signal=100., size=2., frame=[7, 3]))

    class Reader(dict):
        frame_shape = (40, 40)

    out2 = refine.refine_leastsq(two_frames, Reader({3: img, 7: img}), 9)
    assert list(out2['frame']) == [3, 7] and list(out2.index) == [1, 0]
    assert list(out2['cluster']) == [0, 1]
    assert out2['y'].values[0] == out2['y'].values[1]


def test_drop_close_frames_matches_oracle():
    from oracle import find_oracle
    rng = np.random.RandomState(5)
    for ndim, sep in ((2, (9., 9.)), (2, (8., 12.)), (3, (5., 9., 9.))):
        n_frames, cap = 5, 250
        counts = rng.randint(0, cap, n_frames).astype(np.int32)
        counts[0], counts[1] = 0, 1
        coords = rng.randint(0, 100, (n_frames, cap, ndim)).astype(np.int32)
        values = rng.randint(50, 58, (n_frames, cap)).astype(np.int32)       # many ties
        keep = _lib.drop_close_frames(coords, values, counts, sep, 3)
        for f in range(n_frames):
            pos = coords[f, :counts[f]].astype(np.int64)
            want = np.ones(counts[f], bool)
            want[list(find_oracle.where_close(pos, sep, values[f, :counts[f]]))] = False
            assert_array_equal(keep[f, :counts[f]], want)


def test_frame_runs_helper():
    """ctk_frame_runs: first rows of the frame groups (find.py:122) and sortedness, in one pass."""
    from clustertracking_b200 import _lib
    frames = np.array([3, 3, 3, 5, 5, 9, 10, 10], dtype=np.int64)
    starts, is_sorted = _lib.frame_runs(frames)
    assert is_sorted and starts.tolist() == [0, 3, 5, 6]
    starts, is_sorted = _lib.frame_runs(np.array([4, 4, 2, 2, 7], dtype=np.int64))
    assert not is_sorted and starts.tolist() == [0, 2, 4]
    starts, is_sorted = _lib.frame_runs(np.zeros(0, dtype=np.int64))
    assert is_sorted and len(starts) == 0
    many = np.arange(200000, dtype=np.int64)                 # more runs than the first capacity
    starts, is_sorted = _lib.frame_runs(many)
    assert is_sorted and np.array_equal(starts, many)
