"""Small host-side helpers of the refine path (reference: clustertracking/utils.py:24-56, 96-97)."""
import numpy as np


class RefineException(Exception):
    """Per-cluster failure signal, kept for API compatibility (utils.py:96-97).  The CUDA path
    reports failures through per-cluster status codes instead of raising."""


def validate_tuple(value, ndim):
    """``trackpy.utils.validate_tuple`` as the reference uses it (utils.py:6): broadcast a scalar to
    an ndim-tuple, pass a sequence of the right length through."""
    if not hasattr(value, '__iter__'):
        return (value,) * ndim
    value = tuple(value)
    if len(value) != ndim:
        raise ValueError("List length should have same length as image dimensions.")
    return value


def guess_pos_columns(f):
    """utils.py:24-29: three coordinates when a 'z' column exists."""
    return ['z', 'y', 'x'] if 'z' in f else ['y', 'x']


def default_pos_columns(ndim):
    """utils.py:40-41."""
    return ['z', 'y', 'x'][-ndim:]


def default_size_columns(ndim, isotropic):
    """utils.py:44-49."""
    return ['size'] if isotropic else ['size_z', 'size_y', 'size_x'][-ndim:]


def is_isotropic(value):
    """utils.py:52-56."""
    if hasattr(value, '__iter__'):
        value = np.asarray(value)
        return bool(np.all(value[1:] == value[:-1]))
    return True


def host_threads(cap=16):
    """Host threads one process may use: the cores this process can run on, split between the
    ranks that share the box (one process per GPU, ``LOCAL_WORLD_SIZE`` set by torchrun)."""
    import os
    try:
        cores = len(os.sched_getaffinity(0))
    except (AttributeError, OSError):
        cores = os.cpu_count() or 1
    local_ranks = max(1, int(os.environ.get('LOCAL_WORLD_SIZE', '1') or 1))
    share = max(1, cores // local_ranks)
    if os.environ.get('CTK_HOST_THREADS'):               # explicit budget per process
        share = max(1, int(os.environ['CTK_HOST_THREADS']))
    return int(min(cap, share))
