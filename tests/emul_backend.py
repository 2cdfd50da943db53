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
        _handle.ctk_emul_global_pass.restype = ctypes.c_int
        _handle.ctk_emul_global_pass.argtypes = (
            [ctypes.POINTER(_lib.Problem), ctypes.c_void_p, ctypes.POINTER(ctypes.c_int64),
             ctypes.c_double, ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p,
             ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32, ctypes.c_double, ctypes.c_int32]
            + [ctypes.c_void_p] * 5)
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
    if _refine._has_global(kwargs.get('fit_function', 'gauss'), kwargs.get('param_mode')):
        return refine_leastsq_global(f, reader, diameter, **kwargs), None
    plan = _refine.prepare(f, reader, diameter, **kwargs)
    return _refine.finalize(plan, execute(plan)), plan


class GlobalPasses(object):
    """Emulated counterpart of ``global_fit.CudaPasses`` (host arrays, ``ctk_emul_global_pass``)."""

    def __init__(self, plan):
        self.plan, self.handle = plan, lib()
        self.frames = [_refine.load_frame(plan, no) for no in plan.frame_numbers]
        self.ptrs = np.array([fr.ctypes.data for fr in self.frames], dtype=np.uint64)
        self.shape = (ctypes.c_int64 * 3)(*(list(plan.frame_shape) + [1] * (3 - len(plan.frame_shape))))
        P = plan.problem.n_params
        self.G = sum(1 for m in list(plan.problem.modes)[:P] if m == _lib.MODE_GLOBAL)
        self.cap = int(plan.cluster_sizes().max())
        self.launches = 0

    def frame_max(self):
        return max(float(fr.max()) for fr in self.frames)

    def run(self, phase, params, centres, norm, lam, newton, step=None):
        plan = self.plan
        params = np.ascontiguousarray(params, dtype=np.float64)
        centres = np.ascontiguousarray(centres, dtype=np.float64)
        acc = np.zeros(8 + self.G + self.G * (self.G + 1) // 2)
        out = np.empty_like(params)
        cost = np.empty(plan.n_clusters)
        status = np.zeros(plan.n_clusters, dtype=np.int32)
        step = np.ascontiguousarray(step if step is not None else np.zeros(max(self.G, 1)), dtype=np.float64)
        code = self.handle.ctk_emul_global_pass(
            ctypes.byref(plan.problem), self.ptrs.ctypes.data, self.shape, float(norm), plan.n_clusters,
            self.cap, plan.cluster_frame.ctypes.data, plan.cluster_offset.ctypes.data,
            params.ctypes.data, centres.ctypes.data, int(phase), float(lam), int(bool(newton)),
            step.ctypes.data, out.ctypes.data, acc.ctypes.data, cost.ctypes.data, status.ctypes.data)
        assert code == 0, "emulated global pass failed: %d" % code
        self.launches += 1
        reducer = getattr(self, 'reducer', None)
        if reducer is not None:
            reducer.accumulator(acc)
        return acc, (out if phase == 2 else None)


def refine_leastsq_global(f, reader, diameter, **kwargs):
    """``clustertracking_b200.refine_leastsq`` for a global-level fit with the emulated passes."""
    return _refine.refine_leastsq(f, reader, diameter, passes_factory=GlobalPasses, **kwargs)
