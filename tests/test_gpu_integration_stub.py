"""The C ABI as documented, executed standalone (``-m gpu``): the two code blocks of INTEGRATION.md
section 2 -- a pure ``ctypes`` + ``torch`` binding, ONE ``ctk_refine_batch`` call with
``d_work_ids = NULL`` -- run verbatim in a subprocess that never imports ``clustertracking_b200``,
on golden fixtures of the unmodified reference (refine.py:343-430 is what the call replaces).
Proves that ``include/ctk.h``, not ``clustertracking_b200/_lib.py``, is the contract."""
import os
import subprocess
import sys

import numpy as np
import pytest
from numpy.testing import assert_allclose, assert_array_equal

import golden_io

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("name", ["refine_gauss2d_clusters", "refine_dimer2d_constrained"])
def test_integration_stub_verbatim(name, tmp_path):
    out = str(tmp_path / "stub.npz")
    env = dict(os.environ)
    env["LD_LIBRARY_PATH"] = os.path.join(ROOT, "clustertracking_b200") + ":" + env.get("LD_LIBRARY_PATH", "")
    proc = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "integration_stub_runner.py"),
                           name, out], env=env, capture_output=True, text=True, timeout=600)
    assert proc.returncode == 0, proc.stderr[-4000:]
    got = np.load(out)
    want = golden_io.frame(golden_io.load(name), "ref_")
    assert_array_equal(got["index"], want.index.values)
    assert_array_equal(got["col_cluster"], want["cluster"].values)
    assert_array_equal(got["col_cluster_size"], want["cluster_size"].values)
    assert not np.isnan(got["col_cost"]).any() and not np.isnan(want["cost"].values).any(), proc.stderr[-2000:]
    for col in ("y", "x"):
        assert_allclose(got["col_" + col], want[col].values, rtol=0, atol=1e-3, err_msg=col)
    assert_allclose(got["col_signal"], want["signal"].values, rtol=1e-3)
    assert_allclose(got["col_cost"], want["cost"].values, rtol=1e-3, atol=1e-6)
    if name == "refine_dimer2d_constrained":
        p = np.stack([got["col_y"], got["col_x"]], axis=1)
        for c in np.unique(got["col_cluster"]):
            a, b = p[got["col_cluster"] == c]
            assert abs(1 - np.sum(((a - b) / 8.) ** 2)) < 1e-6
