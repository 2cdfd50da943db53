"""Minimal stand-in for the `trackpy` package (absent from this image).

TEST INFRASTRUCTURE ONLY.  It exists so that `oracle/ref_loader.py` can import the
unmodified reference (`/root/reference/clustertracking`) in the build container in
order to generate the golden vectors under `tests/golden/`.  Only the handful of
names the reference imports at module scope are provided; everything that is not
needed on the `refine_leastsq` path raises.
"""
from . import utils, masks, preprocessing  # noqa: F401


def refine(*args, **kwargs):
    raise NotImplementedError("trackpy.refine is not available in the shim")
