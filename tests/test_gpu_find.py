"""GPU tests of the feature-finding step (``-m gpu``): ``clustertracking_b200.find.grey_dilation``
(image work in ``ctk_find_maxima``) against the reference's own outputs (tests/golden/find_*.npz)
and against the CPU oracle on a full-size frame.  Integer work: results must be identical."""
import json

import numpy as np
import pytest
from numpy.testing import assert_array_equal

import golden_io

pytestmark = pytest.mark.gpu


def _kwargs(d):
    kwargs = json.loads(str(d["kwargs"]))
    for key in ("separation", "margin"):
        if isinstance(kwargs.get(key), list):
            kwargs[key] = tuple(kwargs[key])
    return kwargs


@pytest.mark.parametrize("name", golden_io.names("find_"))
def test_grey_dilation_matches_reference_golden(name):
    from clustertracking_b200 import find
    d = golden_io.load(name)
    got = find.grey_dilation(d["image"], **_kwargs(d))
    assert_array_equal(np.asarray(got).reshape(-1, d["image"].ndim), d["pos"])


def test_grey_dilation_batch_equals_single_frames():
    from clustertracking_b200 import find
    names = ["find_2d_u8_fast", "find_2d_u8_fast"]
    d = golden_io.load(names[0])
    image = d["image"]
    flipped = np.ascontiguousarray(image[::-1])
    out = find.grey_dilation_batch([image, flipped], **_kwargs(d))
    assert_array_equal(out[0], d["pos"])
    assert_array_equal(out[1], find.grey_dilation(flipped, **_kwargs(d)))


def test_full_size_frame_equals_oracle():
    """1024x1024 config-2 frame: identical maxima; and the maxima sit on the rendered features."""
    from clustertracking_b200 import artificial, find
    from oracle import find_oracle
    frame, f0, truth = artificial.clustered_frame((1024, 1024), seed=3)
    got = find.grey_dilation(frame, 5, percentile=90)
    want = find_oracle.grey_dilation(frame, 5, percentile=90)
    assert_array_equal(got, want)
    assert len(got) > 1000
    for dtype in (np.uint16,):
        assert_array_equal(find.grey_dilation(frame.astype(dtype) * 3, 5, percentile=90),
                           find_oracle.grey_dilation(frame.astype(dtype) * 3, 5, percentile=90))


def test_unsupported_pixel_types_raise():
    from clustertracking_b200 import find
    with pytest.raises(NotImplementedError):
        find.grey_dilation(np.zeros((32, 32), np.float32), 5)


def test_find_then_refine_matches_oracle():
    """BASELINE config 5 in miniature: integer pixel maxima from the find step are the start
    coordinates of the refinement (the reference's documented pipeline, doc/source/api.rst:10-12);
    both steps on the GPU against both steps of the CPU oracle.  Maxima that are not on a rendered
    feature (noise peaks) are dropped first: fits of noise have many local minima and no solver
    agrees with another on them."""
    import warnings
    import pandas as pd
    from scipy.spatial import cKDTree
    import clustertracking_b200 as ctb
    from clustertracking_b200 import artificial, find
    from oracle import cluster_oracle, find_oracle
    frame, _, truth = artificial.clustered_frame((220, 220), pitch=44, size=2.75, noise=4, seed=21)
    pos = find.grey_dilation(frame, 5, percentile=95, margin=6)
    assert_array_equal(pos, find_oracle.grey_dilation(frame, 5, percentile=95, margin=6))
    dist, _ = cKDTree(truth).query(pos)
    pos = pos[dist < 1.5]
    assert len(pos) >= 0.7 * len(truth)
    f0 = pd.DataFrame(dict(y=pos[:, 0].astype(float), x=pos[:, 1].astype(float), signal=120., size=2.75))
    got = ctb.refine_leastsq(f0.copy(), frame, 11)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        want = cluster_oracle.refine_leastsq(f0.copy(), frame, 11)
    assert_array_equal(got['cluster'].values, want['cluster'].values)
    both = ~np.isnan(got['cost'].values) & ~np.isnan(want['cost'].values)
    assert both.mean() > 0.95
    dpos = np.abs(got[['y', 'x']].values[both] - want[['y', 'x']].values[both]).max(axis=1)
    # clusters with a missed member are ill-posed fits with flat directions, where SLSQP's default
    # tolerance stops early: the bulk must agree to the contract, the tail to a tenth of a pixel
    assert np.percentile(dpos, 90) < 1e-3 and dpos.max() < 0.1, (np.percentile(dpos, 90), dpos.max())


def test_find_features_equals_reference_steps():
    """``find_features`` = the find half of find_link (find_link.py:954-969) frame by frame: the same
    coordinates, mass, signal and size as the oracle / reference-pinned pieces give one frame at a
    time, and a table ``refine_leastsq`` takes as it is."""
    import clustertracking_b200 as ctb
    from clustertracking_b200 import artificial, preprocessing
    from oracle import find_oracle
    reader, _ = artificial.clustered_video(3, shape=(256, 256), seed=40)
    table = ctb.find_features(reader.stack, separation=9, diameter=11, minmass=200, noise_size=1)
    assert list(table.columns) == ['y', 'x', 'mass', 'signal', 'size', 'frame']
    for t in range(3):
        raw = reader.stack[t]
        proc = np.asarray(preprocessing.preprocess(raw, 1, (9, 9), None))
        coords = find_oracle.grey_dilation(proc, (9, 9), 64, (5, 5), precise=True)
        extra = preprocessing.characterize(coords, raw, (5, 5), True)
        keep = extra['mass'] >= 200
        part = table[table['frame'] == t]
        assert_array_equal(part[['y', 'x']].values, coords[keep].astype(float))
        for key in ('mass', 'signal', 'size'):
            assert_array_equal(part[key].values, extra[key][keep])
    out = ctb.refine_leastsq(table, reader, 11)
    assert np.isfinite(out['cost'].values).mean() > 0.9
