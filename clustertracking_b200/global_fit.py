"""Global-level fits: ``param_mode`` with a ``'global'`` column (reference: refine.py:319-332 and the
``level == 'global'`` branches of refine.py:343-430; objective fitfunc.py:421-489 with ``groups``).

A global column is ONE unknown shared by all features of the table, so the per-(frame, cluster)
independence of the cluster level is gone: the reference hands the whole table to a single SLSQP
run (a dense BFGS over every unknown of every cluster).  The normal matrix of that problem is a
block arrow -- one block per cluster, a thin border for the shared unknowns -- and that is how it is
solved here:

* device (``ctk_global_pass``, csrc/ctk_solver.cuh ``run_global``), one warp per cluster, all
  clusters of all frames in one launch: pixel set, residuals and normal equations exactly as in the
  per-cluster fit, elimination of the cluster's own unknowns, and the cluster's share of the Schur
  complement and reduced right-hand side of the shared unknowns added to a small accumulator -- the
  only reduction across clusters on this path (an all-reduce of a few dozen doubles when frames are
  sharded over ranks);
* host (this file): the G x G damped system of the shared unknowns (G is 1-4), the
  Levenberg-Marquardt logic (gain ratio, damping, acceptance) and the re-mask loop of
  refine.py:365-388.  Two launches per iteration; no per-cluster host work.

``dimer_global`` (constraints.py:140-171) is not available.
"""
import logging

import numpy as np

from . import _lib

logger = logging.getLogger(__name__)

HEADER = 8                      # CTK_GLOBAL_HEADER
F0, FT, PRED, STEP, FAILED, SINGULAR = 0, 1, 2, 3, 4, 5


class Reducer(object):
    """Reductions over the ranks that share one global-level fit (frames sharded over GPUs): the
    accumulator of ``ctk_global_pass`` -- a few dozen doubles -- is the only thing that crosses
    ranks on this path.  Without a process group everything is the identity."""

    def __init__(self, group=None, sharded=False):
        self.group, self.sharded = group, sharded

    def host(self, values, op):
        """All-reduce a small host array (op: 'sum' | 'max' | 'min')."""
        values = np.asarray(values, dtype=np.float64)
        if not self.sharded:
            return values
        import torch
        import torch.distributed as dist
        dev = ('cuda' if dist.get_backend(self.group) == 'nccl' else 'cpu')
        t = torch.from_numpy(values.copy()).to(dev)
        dist.all_reduce(t, op=dict(sum=dist.ReduceOp.SUM, max=dist.ReduceOp.MAX,
                                   min=dist.ReduceOp.MIN)[op], group=self.group)
        dist.broadcast(t, src=dist.get_global_rank(self.group, 0) if self.group else 0, group=self.group)
        return t.cpu().numpy()

    def accumulator(self, acc):
        """All-reduce the pass accumulator (a torch tensor on the device or a host array): sums,
        except the step-size slot, which is a maximum."""
        if not self.sharded:
            return acc
        import torch
        import torch.distributed as dist
        t = acc if isinstance(acc, torch.Tensor) else torch.from_numpy(acc)
        step = t[STEP:STEP + 1].clone()
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        dist.all_reduce(step, op=dist.ReduceOp.MAX, group=self.group)
        t[STEP:STEP + 1] = step
        # every rank must take the SAME accept / reject decisions: make the sums bit-identical
        dist.broadcast(t, src=dist.get_global_rank(self.group, 0) if self.group else 0, group=self.group)
        return acc


class CudaPasses(object):
    """Device side of the iteration: buffers as torch tensors, ``ctk_global_pass`` through ctypes."""

    def __init__(self, plan, frames, reducer=None):
        import torch
        self.torch, self.plan, self.frames = torch, plan, frames
        self.reducer = reducer or Reducer()
        self.lib = _lib.load()
        self.dev = frames.dev
        n, P = plan.params_in.shape
        self.G = sum(1 for m in list(plan.problem.modes)[:P] if m == _lib.MODE_GLOBAL)
        self.n_acc = HEADER + self.G + self.G * (self.G + 1) // 2
        with torch.cuda.device(self.dev):
            self.d_offset = torch.from_numpy(plan.cluster_offset).to(self.dev)
            self.d_cframe = torch.from_numpy(plan.cluster_frame).to(self.dev)
            self.d_in = torch.empty((n, P), dtype=torch.float64, device=self.dev)
            self.d_out = torch.empty((n, P), dtype=torch.float64, device=self.dev)
            self.d_centres = torch.empty((n, plan.problem.ndim), dtype=torch.float64, device=self.dev)
            self.d_acc = torch.zeros(self.n_acc, dtype=torch.float64, device=self.dev)
            self.d_step = torch.zeros(max(self.G, 1), dtype=torch.float64, device=self.dev)
            self.d_cost = torch.empty(plan.n_clusters, dtype=torch.float64, device=self.dev)
            self.d_status = torch.zeros(plan.n_clusters, dtype=torch.int32, device=self.dev)
            self.workspace = torch.zeros(256, dtype=torch.uint8, device=self.dev)
        self.shape_arr = (_lib.ctypes.c_int64 * 3)(
            *(list(plan.frame_shape) + [1] * (3 - len(plan.frame_shape))))
        self.cap = int(plan.cluster_sizes().max())
        self.launches = 0

    def frame_max(self):
        self.frames.wait_for_frames(self.frames.n_frames - 1)
        self.torch.cuda.current_stream(self.dev).synchronize()
        return float(self.frames.d_fmax.max().item())

    def run(self, phase, params, centres, norm, lam, newton, step=None):
        torch = self.torch
        with torch.cuda.device(self.dev):
            self.d_in.copy_(torch.from_numpy(np.ascontiguousarray(params)))
            self.d_centres.copy_(torch.from_numpy(np.ascontiguousarray(centres)))
            if step is not None:
                self.d_step[:self.G].copy_(torch.from_numpy(np.ascontiguousarray(step, dtype=np.float64)))
            self.d_acc.zero_()
            stream = _lib.ctypes.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)
            _lib.check(self.lib.ctk_global_pass(
                _lib.ctypes.byref(self.plan.problem), self.frames.d_ptrs.data_ptr(), self.shape_arr,
                float(norm), self.plan.n_clusters, self.cap, self.d_cframe.data_ptr(),
                self.d_offset.data_ptr(), self.d_in.data_ptr(), self.d_centres.data_ptr(), int(phase),
                float(lam), int(bool(newton)), self.d_step.data_ptr(), self.d_out.data_ptr(),
                self.d_acc.data_ptr(), self.d_cost.data_ptr(), self.d_status.data_ptr(),
                self.workspace.data_ptr(), stream), "ctk_global_pass")
            self.launches += 1
            self.reducer.accumulator(self.d_acc)
            acc = self.d_acc.cpu().numpy()
            trial = self.d_out.cpu().numpy() if phase == 2 else None
        return acc, trial


def _unpack_schur(acc, G):
    r = acc[HEADER:HEADER + G].copy()
    S = np.zeros((G, G))
    at = HEADER + G
    for u in range(G):
        for v in range(u + 1):
            S[u, v] = S[v, u] = acc[at]
            at += 1
    return S, r


def _global_step(S, r, lam, g, lo, hi):
    """Damped step of the shared unknowns, projected on their box (a variable that sits on a bound
    and is pushed outward stays there)."""
    G = len(g)
    free = ~(((g <= lo) & (r < 0)) | ((g >= hi) & (r > 0)) | ~(lo < hi))
    step = np.zeros(G)
    if free.any():
        K = S[np.ix_(free, free)].copy()
        K[np.diag_indices_from(K)] += lam * np.abs(np.diag(K)) + 1e-300
        try:
            step[free] = np.linalg.solve(K, r[free])
        except np.linalg.LinAlgError:
            return None
    return np.clip(g + step, lo, hi) - g


def minimise(passes, params, centres, norm, gcols, lo_g, hi_g, f32, lm_max_iter, xtol, allow_newton):
    """Levenberg-Marquardt on the block-arrow system.  -> (ok, objective, params)."""
    G = len(gcols)
    xtol = xtol if xtol > 0. else (2e-6 if f32 else 1e-9)
    eps_f = 4e-6 if f32 else 1e-13
    lam, nu, newton, rejects = 1e-3, 2., False, 0
    g = params[0, gcols].astype(np.float64)
    acc, _ = passes.run(1, params, centres, norm, lam, newton)
    for _ in range(lm_max_iter + 1):
        if acc[FAILED] > 0 or not np.isfinite(acc[F0]):
            return False, np.nan, params
        if acc[SINGULAR] > 0:
            if newton:                       # the exact Hessian is indefinite here: Gauss-Newton
                newton, allow_newton = False, False
            else:
                lam = max(lam * 10., 1e-8)
                rejects += 1
                if rejects > 60:
                    return False, np.nan, params
            acc, _ = passes.run(1, params, centres, norm, lam, newton)
            continue
        F = acc[F0]
        S, r = _unpack_schur(acc, G)
        step = _global_step(S, r, lam, g, lo_g, hi_g)
        if step is None or not np.isfinite(step).all():
            lam = max(lam * 10., 1e-8)
            rejects += 1
            if rejects > 60:
                return False, np.nan, params
            acc, _ = passes.run(1, params, centres, norm, lam, newton)
            continue
        acc2, trial = passes.run(2, params, centres, norm, lam, newton, step)
        if acc2[FAILED] > 0 or acc2[SINGULAR] > 0:
            lam = max(lam * 10., 1e-8)
            rejects += 1
            if rejects > 60:
                return False, np.nan, params
            acc, _ = passes.run(1, params, centres, norm, lam, newton)
            continue
        Ft, pred = acc2[FT], acc2[PRED]
        worst = max(float(acc2[STEP]), float(np.max(np.abs(step) / np.maximum(1., np.abs(g)))) if G else 0.)
        if worst <= xtol:
            return True, F, params           # stationary: the step left is below the tolerance
        noise = pred > 0. and pred <= eps_f * abs(F) and abs(Ft - F) <= 8. * eps_f * abs(F)
        if np.isfinite(Ft) and pred > 0. and (Ft < F or noise):
            if not noise:
                rho = (F - Ft) / pred
                lam = max(lam * max(1. / 3., 1. - (2. * rho - 1.) ** 3), 1e-12)
            nu, rejects = 2., 0
            params = trial
            g = g + step
            newton = allow_newton
        else:
            lam *= nu
            nu *= 2.
            rejects += 1
            if rejects > 40 or lam > 1e18:
                return True, F, params       # no representable descent step is left
        acc, _ = passes.run(1, params, centres, norm, lam, newton)
    return False, np.nan, params             # iteration limit (refine.py:376-377)


def solve(plan, passes, ff, max_iter, max_shift, max_rms_dev, residual_factor, lm_max_iter, xtol):
    """The re-mask loop of refine.py:365-388 around :func:`minimise`.
    -> (ok, params [n, P] in group order, rms_dev)."""
    prob = plan.problem
    P, ndim = prob.n_params, prob.ndim
    modes = list(prob.modes)[:P]
    gcols = [c for c in range(P) if modes[c] == _lib.MODE_GLOBAL]
    f32 = prob.compute_dtype == _lib.COMPUTE_F32
    start = np.array(plan.params_in, dtype=np.float64, copy=True)
    red0 = getattr(passes, 'reducer', None) or Reducer()
    if red0.host([float(not np.isfinite(start).all())], 'max')[0] > 0.:    # refine.py:356-357
        return False, start, np.nan
    # bounds of the shared unknowns: as broad as possible over the rows (fitfunc.py:552-557)
    tables = [np.array([[t[side][j] for j in range(P)] for side in range(2)])
              for t in (prob.bounds_abs, prob.bounds_diff, prob.bounds_rel)]
    abs_t, diff_t, rel_t = tables
    with np.errstate(invalid='ignore'):
        low = np.fmax(np.fmax(start - diff_t[0], start * (1 - rel_t[0])), abs_t[0])
        high = np.fmin(np.fmin(start + diff_t[1], start * (1 + rel_t[1])), abs_t[1])
    low[np.isnan(low)] = -np.inf
    high[np.isnan(high)] = np.inf
    red = getattr(passes, 'reducer', None) or Reducer()
    lo_g = red.host(low[:, gcols].min(axis=0), 'min')
    hi_g = red.host(high[:, gcols].max(axis=0), 'max')
    sums = red.host(np.concatenate((start[:, gcols].sum(axis=0), [len(start)])), 'sum')
    for k, c in enumerate(gcols):                                      # refine.py:361: the mean
        start[:, c] = sums[k] / sums[-1]
    start[:, gcols] = np.clip(start[:, gcols], lo_g, hi_g)             # scipy clips x0 into the box
    fmax = float(red.host([passes.frame_max()], 'max')[0])
    norm = fmax ** 2 / residual_factor                                 # refine.py:325-331
    if not norm > 0.:
        return False, start, np.nan
    centres = start[:, 2:2 + ndim].copy()
    allow_newton = ff.family == _lib.FAMILY_GAUSS
    params, F = start, np.nan
    for _ in range(max_iter):                                          # refine.py:365
        ok, F, params = minimise(passes, start.copy(), centres, norm, gcols, lo_g, hi_g, f32,
                                 lm_max_iter, xtol, allow_newton)
        if not ok:
            return False, start, np.nan
        moved = params[:, 2:2 + ndim]
        far = float(np.any(~(np.sum((moved - centres) ** 2, axis=1) < max_shift ** 2)))
        if red.host([far], 'max')[0] == 0.:                            # refine.py:383-385
            break
        centres = moved.copy()                                         # refine.py:388
    rms_dev = float(np.sqrt(F / residual_factor))                      # refine.py:379
    if not np.isfinite(rms_dev) or rms_dev > max_rms_dev:              # refine.py:391-394
        return False, start, rms_dev
    return True, params, rms_dev


def write_back(plan, ok, params, rms_dev):
    """refine.py:409-422 for level == 'global': the whole table succeeds or fails together."""
    f, ff = plan.f, plan.ff
    if ok:
        block = np.empty((len(ff.params), len(f)), dtype=np.float64)
        for j in range(len(ff.params)):
            block[j, plan.order] = params[:, j]
        for j, col in enumerate(ff.params):
            f[col] = block[j]
        f['cost'] = rms_dev
    else:
        f['cost'] = np.nan
        logger.warning("RefineException: the global fit failed%s",
                       "" if not np.isfinite(rms_dev) else " (rms deviation %.4f)" % rms_dev)
    return f
