"""GPU parity on the fixed-seed subset of the randomised option sweep (``-m gpu``).

``tests/golden/fuzz_<seed>_<case>.npz`` hold 42 cases of ``tests/fuzz_cases.py`` (seeds 1-4 of
``profiles/tools/fuzz_parity.py``): every ring / disc case with a free shape parameter or with
``noise_size``, plus every case in which the device solver and the reference end further than
1e-3 px apart -- answered by the UNMODIFIED reference at its default tolerance and at tol=1e-12
(``oracle/make_golden.py fuzz``).  Bar (tests/fuzz_cases.judge): every cluster within 1e-3 px of the
default-tolerance reference, or at a cost not above the reference's at both tolerances.
"""
import warnings

import pytest

import golden_io
from test_gpu_parity import check_basins

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("precision", ["float32", "float64"])
@pytest.mark.parametrize("name", golden_io.names("fuzz_"))
def test_cuda_fuzz_subset(name, precision):
    import clustertracking_b200 as ctb
    d = golden_io.load(name)
    f0, reader, diameter, kwargs = golden_io.refine_inputs(d, ctb.constraints)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        got = ctb.refine_leastsq(f0, reader, diameter, precision=precision, **kwargs)
    check_basins(got, d)
