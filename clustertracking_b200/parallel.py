"""Frame sharding across the GPUs of one box (one process per GPU, ``torch.distributed``).

At cluster level every (frame, cluster) group of the reference's loop is an independent problem
(refine.py:333-343), so the path shards by frames with NO collective on the data path: each rank
refines a contiguous block of frames on its own GPU.  Only the small result tables are gathered on
the host at the end (``all_gather_object``), and the per-shard cluster ids are shifted so that they
equal the running ids the reference assigns over the whole video (find.py:120-129).
"""
import numpy as np
import pandas as pd

from . import refine as _refine


def shard_bounds(n_items, world_size):
    """Contiguous, balanced blocks: -> int array of world_size + 1 cut points."""
    base, extra = divmod(n_items, world_size)
    sizes = np.full(world_size, base, dtype=np.int64)
    sizes[:extra] += 1
    return np.concatenate(([0], np.cumsum(sizes)))


def frame_shard(f, rank, world_size, t_column='frame'):
    """Rows of ``f`` whose frame falls in this rank's contiguous block of the sorted unique frames."""
    frames = np.unique(f[t_column].values)
    cuts = shard_bounds(len(frames), world_size)
    mine = frames[cuts[rank]:cuts[rank + 1]]
    if len(mine) == 0:
        return f.iloc[0:0]
    sel = (f[t_column].values >= mine[0]) & (f[t_column].values <= mine[-1])
    return f[sel]


def merge_shards(parts):
    """Concatenate per-rank results (rank order = frame order) and make the cluster ids run on
    across shards exactly as a single ``find_clusters`` pass would number them."""
    out, next_id = [], 0
    for part in parts:
        if part is None or len(part) == 0:
            continue
        part = part.copy()
        part['cluster'] = part['cluster'].values + next_id
        next_id = int(part['cluster'].max()) + 1
        out.append(part)
    return pd.concat(out) if out else None


def refine_leastsq_sharded(f, reader, diameter, t_column='frame', group=None, **kwargs):
    """``refine_leastsq`` over the ranks of a ``torch.distributed`` process group: every rank passes
    the same ``f`` and ``reader``, refines its own block of frames on its current CUDA device and
    returns the merged DataFrame (identical on all ranks, identical to a single-GPU call)."""
    import torch.distributed as dist
    if t_column not in f:
        raise ValueError("sharding needs a %r column" % t_column)
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    mine = frame_shard(f, rank, world, t_column)
    part = None
    if len(mine):
        part = _refine.refine_leastsq(mine, reader, diameter, t_column=t_column, **kwargs)
    parts = [None] * world
    dist.all_gather_object(parts, part, group=group)
    return merge_shards(parts)
