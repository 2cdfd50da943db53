"""Frame sharding across the GPUs of one box (one process per GPU, ``torch.distributed``).

At cluster level every (frame, cluster) group of the reference's loop is an independent problem
(refine.py:333-343), so the path shards by frames with NO collective on the data path: each rank
refines a contiguous block of frames on its own GPU.  What remains is the reference's "one table
out" (find.py:157-158 concatenates the per-frame tables):

* the running cluster ids (find.py:120-129) need one number per rank -- how many ids the ranks
  before it used -- which travels in a 2-element ``all_gather``;
* the result tables are gathered WITHOUT pickling: on one host (the case this repository is built
  for: 8 GPUs of one box) every rank writes its columns straight into its slice of one shared-memory
  block (``/dev/shm``) and the receiving rank(s) build the DataFrame over that block without a copy;
  across hosts the same packed blocks (one float64 and one int64 matrix) go through
  ``dist.gather`` / ``dist.all_gather`` as tensors.  Tables with non-numeric columns or a
  non-integer index fall back to ``all_gather_object``.
"""
import os
import socket
import uuid

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


# --------------------------------------------------------------------------------------------------
# packed gather
# --------------------------------------------------------------------------------------------------
def _layout(part):
    """Column split of a result table: -> (float columns, int columns) or None when the table
    cannot travel as two numeric blocks."""
    if not (isinstance(part.index, pd.RangeIndex) or part.index.dtype.kind in 'iu'):
        return None
    floats, ints = [], []
    for col in part.columns:
        kind = part[col].dtype.kind
        if kind == 'f':
            floats.append(col)
        elif kind in 'iub':
            ints.append(col)
        else:
            return None
    return floats, ints


def _dist_device(dist, group):
    """Tensors of a collective must live where the backend works: CUDA for NCCL, host for gloo."""
    import torch
    if dist.get_backend(group) == 'nccl':
        return torch.device('cuda', torch.cuda.current_device())
    return torch.device('cpu')


def _all_gather_ints(values, group):
    """Small all_gather of a few int64 per rank -> array [world, len(values)]."""
    import torch
    import torch.distributed as dist
    dev = _dist_device(dist, group)
    mine = torch.tensor([int(v) for v in values], dtype=torch.int64, device=dev)
    out = [torch.empty_like(mine) for _ in range(dist.get_world_size(group))]
    dist.all_gather(out, mine, group=group)
    return np.stack([t.cpu().numpy() for t in out])


class _SharedBlock(object):
    """One file in /dev/shm holding the merged table: float block [n_f, N], int block [n_i, N]
    (the last int row is the index).  Created by rank 0, written by every rank into its row range,
    unlinked as soon as every reader has mapped it (the mappings keep the memory alive)."""

    def __init__(self, path, n_f, n_i, total, create):
        self.path, self.n_f, self.n_i, self.total = path, n_f, n_i, total
        nbytes = max(8, 8 * total * (n_f + n_i))
        if create:
            with open(path, 'wb') as fh:
                fh.truncate(nbytes)
        self.nbytes = nbytes

    def arrays(self, mode):
        raw = np.memmap(self.path, dtype=np.uint8, mode=mode, shape=(self.nbytes,))
        split = 8 * self.total * self.n_f
        fblock = raw[:split].view(np.float64).reshape(self.n_f, self.total)
        iblock = raw[split:split + 8 * self.total * self.n_i].view(np.int64).reshape(self.n_i, self.total)
        return fblock, iblock


def _table_from_blocks(fblock, iblock, floats, ints, columns, dtypes):
    data = {}
    for col in columns:
        if col in floats:
            values = np.asarray(fblock[floats.index(col)])
        else:
            values = np.asarray(iblock[ints.index(col)])
        data[col] = values.astype(dtypes[col], copy=False)          # e.g. float32 / bool columns
    return pd.DataFrame(data, index=np.asarray(iblock[len(ints)]), copy=False)[list(columns)]


def _same_host(group):
    import torch.distributed as dist
    if os.environ.get('CTK_GATHER', '') == 'tensors' or not os.path.isdir('/dev/shm'):
        mine = "no-shm-%s" % uuid.uuid4().hex
    else:
        try:
            with open('/proc/sys/kernel/random/boot_id') as fh:
                boot = fh.read().strip()
        except OSError:
            boot = ''
        mine = socket.gethostname() + boot
    names = [None] * dist.get_world_size(group)
    dist.all_gather_object(names, mine, group=group)
    return all(name == names[0] for name in names)


def gather_tables(part, group=None, gather='all', ids_done=False):
    """Merge the per-rank result tables (rank order = frame order) with running cluster ids.
    ``gather``: 'all' -> every rank returns the merged table; 'root' -> rank 0 does, the others
    return None; 'none' -> every rank returns its own part (ids already running on)."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    n_rows = 0 if part is None else len(part)
    n_ids = 0 if n_rows == 0 else int(part['cluster'].values.max()) + 1
    layout = _layout(part) if n_rows else ([], [])
    counts = _all_gather_ints([n_rows, n_ids, 0 if layout is None else 1], group)
    row_off = np.concatenate(([0], np.cumsum(counts[:, 0])))
    id_off = np.concatenate(([0], np.cumsum(counts[:, 1])))
    total = int(row_off[-1])
    if total == 0:
        return None
    if n_rows and not ids_done:
        part['cluster'] = part['cluster'].values + int(id_off[rank])     # find.py:127-128
    if gather == 'none':
        return part
    if not counts[:, 2].all():
        # non-numeric columns or labels: the general (pickling) path
        parts = [None] * world
        dist.all_gather_object(parts, part, group=group)
        parts = [p for p in parts if p is not None and len(p)]
        return pd.concat(parts) if (gather == 'all' or rank == 0) else None

    # column order and split: taken from the first non-empty rank (all ranks ran the same call)
    meta = [None] * world
    dist.all_gather_object(meta, None if n_rows == 0 else
                           (list(part.columns), layout, {c: str(part[c].dtype) for c in part.columns}),
                           group=group)
    columns, (floats, ints), dtypes = next(m for m in meta if m is not None)
    n_f, n_i = len(floats), len(ints) + 1                     # + the index
    a, b = int(row_off[rank]), int(row_off[rank + 1])

    def fill(fblock, iblock, lo, hi):
        """This rank's columns -> rows lo..hi of the blocks."""
        jobs = [(fblock[j], part[col].values) for j, col in enumerate(floats)]
        jobs += [(iblock[j], part[col].values) for j, col in enumerate(ints)]
        jobs.append((iblock[len(ints)], part.index.values))

        def copy(job):
            job[0][lo:hi] = job[1]
        _refine._parallel(copy, jobs)

    if _same_host(group):
        name = [None]
        if rank == 0:
            name[0] = os.path.join('/dev/shm', 'ctk_%s' % uuid.uuid4().hex)
            block = _SharedBlock(name[0], n_f, n_i, total, create=True)
        dist.broadcast_object_list(name, src=0, group=group)
        if rank != 0:
            block = _SharedBlock(name[0], n_f, n_i, total, create=False)
        out = None
        try:
            if n_rows:
                fblock, iblock = block.arrays('r+')
                fill(fblock, iblock, a, b)
                del fblock, iblock
            dist.barrier(group=group)                          # every slice is written
            if gather == 'all' or rank == 0:
                # rank 0 owns the block; other readers map it copy-on-write (private tables)
                fblock, iblock = block.arrays('r+' if rank == 0 else 'c')
                out = _table_from_blocks(fblock, iblock, floats, ints, columns, dtypes)
            dist.barrier(group=group)                          # every reader has mapped it
        finally:
            if rank == 0 and os.path.exists(name[0]):
                os.unlink(name[0])
        return out

    # across hosts: the packed blocks as tensors
    dev = _dist_device(dist, group)
    fmine = torch.empty((n_f, n_rows), dtype=torch.float64)
    imine = torch.empty((n_i, n_rows), dtype=torch.int64)
    if n_rows:
        fill(fmine.numpy(), imine.numpy(), 0, n_rows)
    out = None
    blocks = []
    for mine, dtype in ((fmine, torch.float64), (imine, torch.int64)):
        rows = mine.shape[0]
        # equal-sized pieces (padded to the largest shard): what gather / all_gather need
        width = int(counts[:, 0].max())
        padded = torch.zeros((rows, width), dtype=dtype, device=dev)
        padded[:, :n_rows] = mine.to(dev)
        if gather == 'all':
            pieces = [torch.empty_like(padded) for _ in range(world)]
            dist.all_gather(pieces, padded, group=group)
        else:
            pieces = [torch.empty_like(padded) for _ in range(world)] if rank == 0 else None
            dist.gather(padded, pieces, dst=0, group=group)
        if pieces is not None:
            blocks.append(torch.cat([p[:, :int(counts[r, 0])] for r, p in enumerate(pieces)],
                                    dim=1).cpu().numpy())
    if blocks:
        out = _table_from_blocks(blocks[0], blocks[1], floats, ints, columns, dtypes)
    return out


LAST_GATHER = {}      # diagnostics of the most recent shared-block call (host phases, this rank)


class _SharedColumns(object):
    """A [n_slots, total] matrix of 8-byte cells in /dev/shm that all ranks of the host map: column
    k of the merged result table is row k, and a rank's rows are cells row_off[rank] ..
    row_off[rank + 1].  ``refine_leastsq`` writes its result columns straight into these cells
    (its ``_alloc`` hook), so that "gathering" the table costs no copy at all."""

    def __init__(self, path, n_slots, total, a, b, create):
        self.path, self.n_slots, self.total = path, n_slots, total
        self.nbytes = max(8, 8 * n_slots * total)
        if create:
            with open(path, 'wb') as fh:
                fh.truncate(self.nbytes)
        self.mapped = np.memmap(path, dtype=np.int64, mode='r+', shape=(n_slots, total))
        self.private = None                        # copy-on-write mapping (readers other than rank 0)
        self.fresh = True                          # pages not touched yet
        self.begin(a, b)

    def begin(self, a, b):
        """Start a call: rows a..b are this rank's.  Every array handed out during the call derives
        from ONE token object, so that "is any result of that call still alive?" is a weak
        reference (the block is reused by the next call only when the answer is no)."""
        import weakref
        self.a, self.b = a, b
        token = _Token(self.mapped)
        self.cells = np.asarray(token)             # base chain: views -> this array -> token
        self.alive = weakref.ref(token)
        self.private_alive = None
        self.columns = []                          # (name, dtype) in allocation order
        self.spilled = False                       # a column did not fit the 8-byte cells
        self.thread = None
        if self.fresh:
            self._start_populate()

    def end(self):
        """Drop this object's own references to the call's arrays."""
        self.cells = None

    def in_use(self):
        return self.alive() is not None or (self.private_alive is not None and
                                            self.private_alive() is not None)

    def _start_populate(self):
        # Fresh tmpfs pages fault one 4 KB page at a time (no transparent huge pages): ~30 ms per
        # rank for a 2 M-row table if the faults happen where the columns are first written.  A
        # background thread populates the cells of every column as soon as it is allocated, while
        # the host labels clusters and the device works (libc's madvise through ctypes: the call
        # releases the GIL).
        import queue
        import threading
        self.todo = queue.Queue()
        self.thread = threading.Thread(target=self._populate, args=(self.cells,), daemon=True)
        self.thread.start()

    def _populate(self, cells):
        import ctypes
        import mmap
        page = mmap.PAGESIZE
        populate_write = 23                        # MADV_POPULATE_WRITE (Linux 5.14+)
        try:
            libc = ctypes.CDLL(None, use_errno=True)
            libc.madvise.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]
        except (OSError, AttributeError):
            libc = None
        base = cells.ctypes.data
        while True:
            k = self.todo.get()
            if k is None:
                return
            start = 8 * (k * self.total + self.a)
            stop = 8 * (k * self.total + self.b)
            lo, hi = start // page * page, min(self.nbytes, -(-stop // page) * page)
            if libc is None or libc.madvise(base + lo, hi - lo, populate_write) != 0:
                # fallback: READ one word per page (allocates and zeroes the page; a write here
                # could land after refine_leastsq has filled the column)
                int(np.asarray(cells[k, self.a:self.b:page // 8]).sum())

    def ready(self):
        thread, self.thread = self.thread, None
        if thread is not None:
            self.todo.put(None)
            thread.join()
        self.fresh = False

    def alloc(self, name, dtype):
        dtype = np.dtype(dtype)
        k = len(self.columns)
        if dtype.itemsize != 8 or dtype.kind not in 'fiu' or k >= self.n_slots - 1:
            self.spilled = True
            return np.empty(self.b - self.a, dtype=dtype)
        self.columns.append((name, dtype.str))
        if self.thread is not None:
            self.todo.put(k)
        return self.cells[k, self.a:self.b].view(dtype)

    def table(self, columns, private):
        cells = self.cells
        if private:
            import weakref
            if self.private is None:
                self.private = np.memmap(self.path, dtype=np.int64, mode='c',
                                         shape=(self.n_slots, self.total))
            token = _Token(self.private)
            cells = np.asarray(token)
            self.private_alive = weakref.ref(token)
        data = {name: cells[k].view(np.dtype(dt)) for k, (name, dt) in enumerate(columns)}
        index = cells[self.n_slots - 1]
        return pd.DataFrame(data, index=index, copy=False)


class _Token(object):
    """Owner of the arrays of one call: exposes a mapping through the array interface and keeps it
    alive; weakly referenceable, which numpy arrays viewed from a memmap are not usefully."""

    def __init__(self, array):
        self.__array_interface__ = dict(array.__array_interface__)
        self._keep = array


# The shared block of the previous call, kept mapped (and its file linked) so that the next call of
# the same size writes into pages that are already resident: populating a fresh tmpfs mapping costs
# ~1 ms per 2 MB and was the largest host cost of the sharded path at 8 ranks.  It is reused only
# when NO array of the previous result is alive any more on any rank (weak references above).
_CACHE = {}


def _drop_cache(unlink):
    block = _CACHE.pop('block', None)
    if block is not None and unlink and os.path.exists(block.path):
        os.unlink(block.path)


def _cache_signature():
    block = _CACHE.get('block')
    if block is None:
        return [0, 0, 0, 0]
    tag = int.from_bytes(block.path.encode()[-7:], 'little')
    return [0 if block.in_use() else 1, block.total, block.n_slots, tag]


def _refine_into_shared(mine, reader, diameter, t_column, group, gather, kwargs):
    """The one-host path: result columns allocated in a shared block, no copy at gather time.
    -> (done, result); done = False: fall back to gather_tables (inputs this scheme cannot hold)."""
    import time
    import torch.distributed as dist
    lap = [time.perf_counter()]
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    index_ok = isinstance(mine.index, pd.RangeIndex) or mine.index.dtype.kind in 'iu'
    numeric = all(mine[c].dtype.kind in 'fiu' and mine[c].dtype.itemsize == 8 for c in mine.columns)
    counts = _all_gather_ints([len(mine), 1 if (index_ok and numeric) else 0, len(mine.columns)]
                              + _cache_signature(), group)
    host_ok = _same_host(group)
    if not (host_ok and counts[:, 1].all()) or counts[:, 0].sum() == 0:
        return False, None
    row_off = np.concatenate(([0], np.cumsum(counts[:, 0])))
    total, a, b = int(row_off[-1]), int(row_off[rank]), int(row_off[rank + 1])
    n_slots = int(counts[:, 2].max()) + 24               # input columns + model columns + index
    # every rank evaluates the same predicate on the gathered signatures: the cached block is free
    # everywhere, has this call's shape and is the same file on all ranks
    reuse = bool(counts[:, 3].all() and (counts[:, 4] == total).all() and (counts[:, 5] == n_slots).all()
                 and (counts[:, 6] == counts[0, 6]).all() and os.environ.get('CTK_SHARED_CACHE', '1') != '0')
    if reuse:
        shared = _CACHE['block']
        shared.begin(a, b)
    else:
        _drop_cache(unlink=rank == 0)
        name = [None]
        if rank == 0:
            name[0] = os.path.join('/dev/shm', 'ctk_%d_%s' % (os.getpid(), uuid.uuid4().hex[:12]))
            shared = _SharedColumns(name[0], n_slots, total, a, b, create=True)
        dist.broadcast_object_list(name, src=0, group=group)
        if rank != 0:
            shared = _SharedColumns(name[0], n_slots, total, a, b, create=False)
        _CACHE['block'] = shared
        if rank == 0 and not _CACHE.get('atexit'):
            import atexit
            atexit.register(_drop_cache, True)
            _CACHE['atexit'] = True
    LAST_GATHER.clear()
    LAST_GATHER['block_reused'] = reuse
    try:
        part = None
        lap.append(time.perf_counter())
        if b > a:
            part = _refine.refine_leastsq(mine, reader, diameter, t_column=t_column,
                                          _alloc=shared.alloc, **kwargs)
            lap.append(time.perf_counter())
            shared.ready()
            complete = (not shared.spilled and [c for c, _ in shared.columns] != [] and
                        set(part.columns) == set(c for c, _ in shared.columns))
            shared.cells[n_slots - 1, a:b] = part.index.values
        else:
            complete = True
            shared.ready()
        n_ids = 0 if part is None else int(part['cluster'].values.max()) + 1
        ids = _all_gather_ints([n_ids, 1 if complete else 0], group)
        id_off = int(ids[:rank, 0].sum())
        if part is not None and id_off:                            # find.py:127-128
            if complete:                                           # the column IS a view of the cells
                k = [c for c, _ in shared.columns].index('cluster')
                shared.cells[k, a:b] += id_off
            else:
                part['cluster'] = part['cluster'].values + id_off
        if not ids[:, 1].all():
            return True, gather_tables(part, group, gather, ids_done=True)
        if gather == 'none':
            dist.barrier(group=group)
            return True, part
        meta = [None] * world
        dist.all_gather_object(meta, None if part is None else (shared.columns, list(part.columns)),
                               group=group)                        # doubles as "all slices written"
        columns, order = next(m for m in meta if m is not None)
        out = None
        lap.append(time.perf_counter())
        if gather == 'all' or rank == 0:
            out = shared.table(columns, private=rank != 0)[order]
        lap.append(time.perf_counter())
        dist.barrier(group=group)                                  # every reader has mapped it
        lap.append(time.perf_counter())
        if len(lap) == 6:
            LAST_GATHER.update(zip(("setup_ms", "refine_ms", "ids_and_wait_ms", "table_ms", "barrier_ms"),
                                   (1e3 * (y - x) for x, y in zip(lap[:-1], lap[1:]))))
        return True, out
    except BaseException:
        _drop_cache(unlink=rank == 0)              # whatever state the block is in: not reused
        raise
    finally:
        part = None
        shared.end()


def refine_leastsq_sharded(f, reader, diameter, t_column='frame', group=None, presharded=False,
                           gather='all', **kwargs):
    """``refine_leastsq`` over the ranks of a ``torch.distributed`` process group.  Every rank
    refines a contiguous block of frames on its current CUDA device; the result equals a single-GPU
    call on the whole table (rows, order, cluster ids, numbers).

    ``presharded=False``: every rank passes the same whole table ``f`` and takes its block of the
    sorted unique frames.  ``presharded=True``: ``f`` holds THIS rank's frames only; rank order must
    be frame order (rank 0 the earliest frames) -- no rank ever holds the whole input.
    ``gather``: 'all' (default) every rank returns the merged DataFrame; 'root' only rank 0 does (the
    others return None); 'none' every rank returns its own part.  ``reader`` must serve the frames
    of this rank's rows."""
    import torch.distributed as dist
    if t_column not in f:
        raise ValueError("sharding needs a %r column" % t_column)
    if gather not in ('all', 'root', 'none'):
        raise ValueError("gather must be 'all', 'root' or 'none'")
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    mine = f if presharded else frame_shard(f, rank, world, t_column)
    if _refine._has_global(kwargs.get('fit_function', 'gauss'), kwargs.get('param_mode')):
        # a global-level fit is ONE problem over all ranks: the accumulator of every pass is
        # all-reduced (global_fit.Reducer) -- the only collective on a data path of this package
        from . import global_fit
        if world > 1 and len(mine) == 0:
            raise ValueError("global-level fits need at least one frame per rank")
        part = _refine.refine_leastsq(mine, reader, diameter, t_column=t_column,
                                      reducer=global_fit.Reducer(group, sharded=world > 1), **kwargs)
        return gather_tables(part, group, gather)
    if os.environ.get('CTK_GATHER', '') not in ('tensors', 'copy'):
        done, out = _refine_into_shared(mine, reader, diameter, t_column, group, gather, kwargs)
        if done:
            return out
    part = None
    if len(mine):
        part = _refine.refine_leastsq(mine, reader, diameter, t_column=t_column, **kwargs)
    return gather_tables(part, group, gather)
