"""TEST INFRASTRUCTURE: run a ``clustertracking_b200.refine.Plan`` through the one-lane host build of
the device solver (tests/emul/ctk_emul.cpp).  Lets the CPU test-suite check the solver logic and the
host packing in the GPU-less build container.  Never imported by the package."""
import ctypes
import hashlib
import os
import subprocess

import numpy as np

from clustertracking_b200 import _lib
from clustertracking_b200 import refine as _refine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "emul", "ctk_emul.cpp")
BUILD = os.path.join(ROOT, "tests", "emul", "_build")
_handle = None


def _build():
    h = hashlib.sha256()
    for path in (SRC, os.path.join(ROOT, "include", "ctk.h"),
                 os.path.join(ROOT, "clustertracking_b200", "csrc", "ctk_solver.cuh"),
                 os.path.join(ROOT, "clustertracking_b200", "csrc", "ctk_layout.h")):
        with open(path, "rb") as fh:
            h.update(fh.read())
    tag = h.hexdigest()[:16]
    out = os.path.join(BUILD, "libctk_emul_%s.so" % tag)
    if not os.path.exists(out):
        os.makedirs(BUILD, exist_ok=True)
        for stale in os.listdir(BUILD):                     # one build at a time: they travel with the repo
            if stale.startswith("libctk_emul_") and stale.endswith(".so"):
                os.remove(os.path.join(BUILD, stale))
        opt = os.environ.get("CTK_EMUL_OPT", "-O1")
        subprocess.check_call(["g++", opt, "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared",
                               "-I", os.path.join(ROOT, "include"),
                               "-I", os.path.join(ROOT, "clustertracking_b200", "csrc"),
                               SRC, "-o", out])
    return out


def lib():
    global _handle
    if _handle is None:
        _handle = ctypes.CDLL(_build())
        _handle.ctk_emul_refine_batch.restype = ctypes.c_int
        _handle.ctk_emul_refine_batch.argtypes = (
            [ctypes.POINTER(_lib.Problem), ctypes.c_void_p, ctypes.POINTER(ctypes.c_int64),
             ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p, ctypes.c_int32] + [ctypes.c_void_p] * 9)
    return _handle


def execute(plan):
    """Emulated counterpart of ``refine.execute_cuda``."""
    handle = lib()
    result = _refine.Result(plan)
    frames = [_refine.load_frame(plan, no) for no in plan.frame_numbers]
    ptrs = np.array([fr.ctypes.data for fr in frames], dtype=np.uint64)
    fmax = np.array([float(fr.max()) for fr in frames], dtype=np.float64)
    shape = (ctypes.c_int64 * 3)(*(list(plan.frame_shape) + [1] * (3 - len(plan.frame_shape))))
    sizes = plan.cluster_sizes()

    def launch(cap, ids, rigorous):
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        problem = _refine.rigorous_problem(plan.problem) if rigorous else plan.problem
        code = handle.ctk_emul_refine_batch(
            ctypes.byref(problem), ptrs.ctypes.data, shape, fmax.ctypes.data, len(ids),
            ids.ctypes.data, int(cap), plan.cluster_frame.ctypes.data,
            plan.cluster_offset.ctypes.data, plan.params_in.ctypes.data,
            plan.bounds_lo.ctypes.data if plan.bounds_lo is not None else None,
            plan.bounds_hi.ctypes.data if plan.bounds_hi is not None else None,
            result.params_out.ctypes.data, result.cost.ctypes.data,
            result.status.ctypes.data, result.stats.ctypes.data)
        assert code == 0, "emulated launch failed: %d" % code

    _refine.run_bins(sizes, np.arange(plan.n_clusters), result.status, launch)
    return result


def refine_leastsq(f, reader, diameter, **kwargs):
    plan = _refine.prepare(f, reader, diameter, **kwargs)
    return _refine.finalize(plan, execute(plan)), plan
