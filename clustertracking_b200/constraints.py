"""Distance constraints of the refine path (reference: clustertracking/constraints.py:59-137).

The reference hands python callables to SLSQP, which differentiates them numerically.  On the GPU
the dimer / trimer / tetramer constraints are built into the solver (augmented-Lagrangian rows), so the dicts
returned here are *descriptors*: ``refine_leastsq`` recognises them by their ``fun`` and reads the
distances from ``args``.  They keep the reference's keys (``type``, ``cluster_size``, ``fun``,
``args``) and the callables evaluate the same expressions, so user code that inspects or calls them
keeps working.
"""
import numpy as np

from .utils import validate_tuple


def _pair_term(pos, a, b, dist):
    return 1 - np.sum(((pos[:, a] - pos[:, b]) / dist) ** 2, axis=1)


def _dimer_fun(x, dist, ndim):
    """1 - |p0 - p1|^2 (in units of dist) per cluster; x has axes (cluster, feature, parameter)
    with the positions in columns 2..2+ndim (constraints.py:59-61)."""
    return _pair_term(x[..., 2:2 + ndim], 0, 1, dist)


def _trimer_fun(x, dist, ndim):
    """The three pair terms (0,1), (1,2), (0,2) (constraints.py:79-83)."""
    pos = x[..., 2:2 + ndim]
    return np.concatenate([_pair_term(pos, a, b, dist) for a, b in ((0, 1), (1, 2), (0, 2))])


def _tetramer_fun(x, dist, ndim):
    """2D: the four shortest of the six pair distances (the sides of a square); 3D: all six
    (constraints.py:102-125)."""
    pos = x[..., 2:2 + ndim]
    terms = [1 - _pair_term(pos, a, b, dist)
             for a, b in ((0, 1), (1, 2), (0, 2), (1, 3), (0, 3), (2, 3))]
    if ndim == 2:
        return np.ravel(1 - np.sort(np.vstack(terms), axis=0)[:4])
    return np.concatenate([1 - t for t in terms])


def dimer(dist, ndim=2):
    """Constrain clusters of 2 to the given centre distance; a tuple gives per-axis distances
    (constraints.py:64-76)."""
    dist = np.array(validate_tuple(dist, ndim), dtype=np.float64)
    return (dict(type='eq', cluster_size=2, fun=_dimer_fun, args=(dist, ndim)),)


def trimer(dist, ndim=2):
    """Constrain clusters of 3: all three distances equal ``dist`` (constraints.py:86-99)."""
    dist = np.array(validate_tuple(dist, ndim), dtype=np.float64)
    return (dict(type='eq', cluster_size=3, fun=_trimer_fun, args=(dist, ndim)),)


def tetramer(dist, ndim=2):
    """Constrain clusters of 4: a square in 2D (4 constraints), a tetrahedron in 3D (6)
    (constraints.py:127-137)."""
    if ndim not in (2, 3):
        raise NotImplementedError
    dist = np.array(validate_tuple(dist, ndim), dtype=np.float64)
    return (dict(type='eq', cluster_size=4, fun=_tetramer_fun, args=(dist, ndim)),)


def parse(constraints, ndim):
    """-> dict(dimer=dist | None, trimer=dist | None) from an iterable of constraint dicts.

    Accepts this module's descriptors and the reference's own (recognised by function name, so
    dicts built by ``clustertracking.constraints`` work unchanged).  Anything the CUDA solver does
    not carry raises NotImplementedError: there is no CPU fallback."""
    out = dict(dimer=None, trimer=None, tetramer=None)
    if not constraints:
        return out
    for cons in constraints:
        name = getattr(cons.get('fun', None), '__name__', '')
        size = cons.get('cluster_size', None)
        kind = {('_dimer_fun', 2): 'dimer', ('_trimer_fun', 3): 'trimer',
                ('_tetramer_fun', 4): 'tetramer', ('_tetramer_fun_2d', 4): 'tetramer',
                ('_tetramer_fun_3d', 4): 'tetramer'}.get((name, size))
        if kind is None or cons.get('type', 'eq') != 'eq':
            raise NotImplementedError(
                "constraint %r (cluster_size=%r) is not available in the CUDA solver; supported: "
                "constraints.dimer, constraints.trimer, constraints.tetramer"
                % (name or cons.get('fun'), size))
        dist = np.asarray(validate_tuple(cons['args'][0] if np.ndim(cons['args'][0]) == 0
                                         else tuple(cons['args'][0]), ndim), dtype=np.float64)
        if out[kind] is not None and not np.array_equal(out[kind], dist):
            raise ValueError("two different %s constraints given" % kind)
        out[kind] = dist
    return out
