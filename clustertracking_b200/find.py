"""Cluster grouping on the host (reference: clustertracking/find.py:12-163).

``find_clusters`` decides which features are fitted together, so membership, row order and even the
label values must equal the reference's.  The close pairs come from the same scipy
``cKDTree.query_pairs`` as upstream and are visited in the same order -- the iteration order of the
python ``set`` upstream receives (find.py:87-91).  That order is replayed in C from the array form
of the result (``ctk_pairs_set_order``; no python objects), the union step runs in C
(``ctk_label_clusters``) instead of the reference's dict-of-sets loop (find.py:12-60), frames are
cut from one stable sort instead of a pandas ``groupby`` + ``concat``, and frames are processed by
host threads inside libctk (``ctk_cluster_frames``), which restates scipy's kd-tree for this one
call so that the pairs come out in scipy's order; the restatement is verified against the installed
scipy once per process, and scipy itself (in worker processes for long videos) is the fallback.
"""
import atexit
import os
import pickle
import subprocess
import sys
from itertools import chain
from multiprocessing import shared_memory

import numpy as np
from scipy.spatial import cKDTree

from . import _lib
from .utils import guess_pos_columns, validate_tuple

_replay_checked = None      # None = not checked yet, True = replay matches this interpreter's sets


def _pairs_via_set(tree):
    pairs = tree.query_pairs(1)                                     # a python set, as upstream
    flat = np.fromiter(chain.from_iterable(pairs), dtype=np.int64, count=2 * len(pairs))
    return flat.reshape(-1, 2)


def _pairs_via_replay(tree):
    return _lib.pairs_set_order(tree.query_pairs(1, output_type='ndarray'))


def _replay_is_exact():
    """One-time self check: the C replay of CPython's set order must reproduce a real set."""
    global _replay_checked
    if _replay_checked is None:
        rng = np.random.RandomState(12345)
        ok = True
        for n in (40, 700, 3000):
            tree = cKDTree(rng.uniform(0, 4 * np.sqrt(n), (n, 2)))
            ok = ok and np.array_equal(_pairs_via_set(tree), _pairs_via_replay(tree))
        _replay_checked = bool(ok)
    return _replay_checked


_native_checked = None      # None = not checked yet, True = ctk_query_pairs reproduces scipy's order


def _native_is_exact():
    """One-time self check of the kd-tree restatement in libctk (ctk_query_pairs) against the
    installed scipy: same pairs in the same order on random, integer-grid and duplicate-heavy data.
    If scipy ever changes its traversal the labelling falls back to scipy itself."""
    global _native_checked
    if _native_checked is None:
        rng = np.random.RandomState(54321)
        ok = _replay_is_exact()
        cases = [rng.uniform(0, 30, (900, 2)), rng.uniform(0, 9, (700, 3)),
                 rng.randint(0, 200, (800, 2)) / 11., rng.randint(0, 60, (600, 3)) / [9., 13., 13.],
                 rng.randint(0, 4, (300, 2)) * 0.5, rng.uniform(0, 0.6, (40, 2))]
        for data in cases:
            data = np.ascontiguousarray(data, dtype=np.float64)
            want = cKDTree(data).query_pairs(1, output_type='ndarray')
            ok = ok and np.array_equal(want, _lib.query_pairs(data))
        _native_checked = bool(ok)
    return _native_checked


def _label_frame(pos, separation):
    """ids, sizes (int64 arrays) for the points of one frame (find.py:72-93)."""
    n = len(pos)
    if n == 0:
        return np.zeros(0, np.int64), np.zeros(0, np.int64)
    tree = cKDTree(pos / separation)
    pairs = _pairs_via_replay(tree) if _replay_is_exact() else _pairs_via_set(tree)
    return _lib.label_clusters(pairs, n)


def find_iter(f, separation, pos_columns=None, t_column='frame'):
    """Generator of ``(frame_no, DataFrame)`` with ``cluster`` and ``cluster_size`` columns added
    (find.py:96-129)."""
    if pos_columns is None:
        pos_columns = guess_pos_columns(f)
    separation = np.asarray(validate_tuple(separation, len(pos_columns)), dtype=np.float64)
    next_id = 0
    for frame_no, part in f.groupby(t_column):
        ids, sizes = _label_frame(part[pos_columns].values.astype(np.float64), separation)
        part = part.copy()
        part['cluster'] = ids + next_id
        part['cluster_size'] = sizes
        next_id = int(part['cluster'].max()) + 1
        yield frame_no, part


_POOL = None
_POOL_MIN_FRAMES = 64


def _pool_workers():
    env = os.environ.get('CTK_FIND_WORKERS')
    if env is not None:
        return max(0, int(env))
    # one process per GPU shares the host: split the cores between the local ranks
    from .utils import host_threads
    return host_threads(16)


class _WorkerPool(object):
    """Persistent worker processes, each a fresh interpreter running
    ``python -m clustertracking_b200._find_worker`` (no fork of this process, no re-import of the
    caller's ``__main__``, no CUDA state).  Tasks and replies are pickles on the workers' pipes;
    the bulk data travels through shared memory."""

    def __init__(self, n):
        env = dict(os.environ)
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        env['PYTHONPATH'] = root + os.pathsep + env.get('PYTHONPATH', '')
        env['CTK_FIND_WORKERS'] = '0'
        for var in ('OMP_NUM_THREADS', 'OPENBLAS_NUM_THREADS', 'MKL_NUM_THREADS'):
            env[var] = '1'
        self.procs = [subprocess.Popen([sys.executable, '-m', 'clustertracking_b200._find_worker'],
                                       stdin=subprocess.PIPE, stdout=subprocess.PIPE, env=env)
                      for _ in range(n)]

    def __len__(self):
        return len(self.procs)

    def send(self, tasks):
        """One task per worker (len(tasks) <= len(self))."""
        for proc, task in zip(self.procs, tasks):
            pickle.dump(task, proc.stdin)
            proc.stdin.flush()

    def receive(self, n_tasks):
        """Replies of the last ``send`` in order; an exception in a worker is re-raised."""
        replies = [pickle.load(proc.stdout) for proc in self.procs[:n_tasks]]
        for reply in replies:
            if isinstance(reply, Exception):
                raise reply
        return replies

    def close(self):
        for proc in self.procs:
            try:
                proc.stdin.close()
                proc.terminate()
            except Exception:
                pass
        self.procs = []


def _get_pool():
    global _POOL
    workers = _pool_workers()
    if workers < 2:
        return None
    if _POOL is None or any(p.poll() is not None for p in _POOL.procs):
        if _POOL is not None:
            _POOL.close()
        _POOL = _WorkerPool(workers)
        atexit.register(_close_pool)
    return _POOL


def _close_pool():
    global _POOL
    if _POOL is not None:
        _POOL.close()
        _POOL = None


def _label_range(pos, ranges, separation, cluster, size, by_cluster):
    """Label the frames ``ranges`` = [(a, b), ...] of frame-sorted ``pos`` into the output arrays;
    returns the per-frame label spans (max label + 1)."""
    spans = []
    for a, b in ranges:
        ids, sizes = _label_frame(pos[a:b], separation)
        cluster[a:b] = ids
        size[a:b] = sizes
        by_cluster[a:b] = a + np.argsort(ids, kind='stable')
        spans.append(int(ids.max()) + 1 if b > a else 0)
    return spans


def _pool_task(args):
    in_name, out_name, n, ndim, ranges, separation = args
    shm_in = shared_memory.SharedMemory(name=in_name)
    shm_out = shared_memory.SharedMemory(name=out_name)
    for shm in (shm_in, shm_out):        # the parent owns the segments: do not track them here
        try:
            from multiprocessing import resource_tracker
            resource_tracker.unregister(shm._name, 'shared_memory')
        except Exception:
            pass
    try:
        pos = np.ndarray((n, ndim), dtype=np.float64, buffer=shm_in.buf)
        out = np.ndarray((3, n), dtype=np.int64, buffer=shm_out.buf)
        return _label_range(pos, ranges, separation, out[0], out[1], out[2])
    finally:
        shm_in.close()
        shm_out.close()


class _LabelJob(object):
    """Labelling of all frames, running on host threads inside libctk (ctk_cluster_frames; ctypes
    releases the GIL) while the caller does other host work; ``result()`` waits and returns
    (cluster, cluster_size, by_cluster).  If the library's kd-tree restatement does not reproduce
    the installed scipy (self check), scipy itself is used: in worker processes for long videos."""

    def __init__(self, pos, starts, stops, separation):
        self.starts, self.stops = np.asarray(starts), np.asarray(stops)
        n_frames, n = len(starts), len(pos)
        self.pool = None
        self.thread = None
        self.n = n
        if os.environ.get('CTK_FIND_NATIVE', '1') != '0' and _native_is_exact():
            self.error = None
            args = (np.ascontiguousarray(pos, dtype=np.float64), self.starts.astype(np.int64),
                    self.stops.astype(np.int64), separation, max(1, _pool_workers()))

            def work():
                try:
                    self.cluster, self.size, self.by_cluster, spans = _lib.cluster_frames(*args)
                    self.spans = spans.tolist()
                except Exception as exc:       # re-raised in result()
                    self.error = exc

            if n_frames >= 8:
                import threading
                self.thread = threading.Thread(target=work)
                self.thread.start()
            else:
                work()
            return
        ranges = list(zip((int(a) for a in starts), (int(b) for b in stops)))
        self.pool = _get_pool() if n_frames >= _POOL_MIN_FRAMES else None
        if self.pool is None:
            self.cluster = np.empty(n, dtype=np.int64)
            self.size = np.empty(n, dtype=np.int64)
            self.by_cluster = np.empty(n, dtype=np.int64)
            self.spans = _label_range(pos, ranges, separation, self.cluster, self.size,
                                      self.by_cluster)
            return
        self.shm_in = shared_memory.SharedMemory(create=True, size=max(1, pos.nbytes))
        self.shm_out = shared_memory.SharedMemory(create=True, size=max(1, 3 * n * 8))
        np.ndarray(pos.shape, dtype=np.float64, buffer=self.shm_in.buf)[:] = pos
        per_task = max(1, -(-n_frames // len(self.pool)))
        self.tasks = [(self.shm_in.name, self.shm_out.name, n, pos.shape[1],
                       ranges[k:k + per_task], separation) for k in range(0, n_frames, per_task)]
        self.pool.send(self.tasks)

    def result(self):
        if self.thread is not None:
            self.thread.join()
            self.thread = None
        if getattr(self, 'error', None) is not None:
            raise self.error
        if self.pool is not None:
            try:
                spans = [s for part in self.pool.receive(len(self.tasks)) for s in part]
                out = np.ndarray((3, self.n), dtype=np.int64, buffer=self.shm_out.buf)
                self.cluster, self.size, self.by_cluster = out[0].copy(), out[1].copy(), out[2].copy()
                self.spans = spans
            finally:
                for shm in (self.shm_in, self.shm_out):
                    shm.close()
                    shm.unlink()
                self.pool = None
        offsets = np.concatenate(([0], np.cumsum(self.spans)[:-1])).astype(np.int64)
        cluster = self.cluster + np.repeat(offsets, self.stops - self.starts)
        return cluster, self.size, self.by_cluster


_PINNED = {}        # cached pinned staging buffers of the device labelling (one call at a time)
_SCRATCH = {}       # cached device scratch per device
_STREAMS = {}       # the labelling stream of each device


def _pinned(torch, key, nbytes):
    buf = _PINNED.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, pin_memory=True)
        _PINNED[key] = buf
    return buf[:nbytes]


class Prestaged(object):
    """The position columns on their way into the pinned staging buffer, on threads, started the
    moment ``refine_leastsq`` is entered (the copy needs nothing the argument handling computes)."""

    def __init__(self, columns):
        import threading
        import torch
        from concurrent.futures import ThreadPoolExecutor
        from .utils import host_threads
        self.columns = columns
        ndim, n = len(columns), len(columns[0])
        # room for the frame bounds behind the columns (at most one frame per row)
        self.stage = _pinned(torch, "pos", (ndim * n + 2 * n) * 8)
        cols = self.stage.numpy()[:ndim * n * 8].view(np.float64).reshape(ndim, n)
        pieces = [(k, a, min(n, a + (1 << 20))) for k in range(ndim) for a in range(0, n, 1 << 20)]
        workers = max(1, min(4, host_threads(4), len(pieces)))

        def run():
            with ThreadPoolExecutor(workers) as pool:
                list(pool.map(lambda p: np.copyto(cols[p[0], p[1]:p[2]], columns[p[0]][p[1]:p[2]]), pieces))

        self.thread = threading.Thread(target=run, daemon=True)
        self.thread.start()

    def close(self):
        thread, self.thread = self.thread, None
        if thread is not None:
            thread.join()


class DeviceLabels(object):
    """Labels of all frames of a frame-sorted table from ``ctk_label_frames`` (one warp per frame on
    the GPU, label values identical to the reference's).  Both directions go through MAPPED pinned
    The staged position columns are copied to the device ahead of the frame uploads; labels and
    per-frame flags are written by the kernel straight into MAPPED pinned host memory, so nothing
    waits behind the frame uploads on the copy engines and a chunk of frames can be consumed as
    soon as ITS flags have arrived.

    ``start()`` (on the labelling thread) stages the columns and launches; ``wait_frames(fa, fb)``
    blocks until frames fa..fb-1 are labelled -> (labels int32 [n], flags int32 [n_frames]) views.
    Frames whose flag is 1 exceeded a scratch capacity on the device and are labelled by the host
    path."""

    def __init__(self, pos, starts, stops, separation, device, prestaged=None):
        self.prestaged = prestaged if prestaged is not None and prestaged.columns is pos else None
        self.pos, self.starts, self.stops = pos, starts, stops
        self.separation, self.device = separation, device
        n, n_frames, ndim = len(pos[0]), len(starts), len(pos)
        self.h2d_bytes = n * ndim * 8 + n_frames * 16
        self.d2h_bytes = n * 4 + n_frames * 4
        self.ms = {}

    def start(self):
        import time
        import torch
        from concurrent.futures import ThreadPoolExecutor
        from .utils import host_threads
        pos, device = self.pos, self.device
        n, n_frames, ndim = len(pos[0]), len(self.starts), len(pos)
        t0 = time.perf_counter()
        max_points = int((self.stops - self.starts).max())
        if self.prestaged is not None:                     # copied while the arguments were handled
            self.prestaged.close()
            stage = self.prestaged.stage[:(ndim * n + 2 * n_frames) * 8]
            host = stage.numpy()
        else:
            stage = _pinned(torch, "pos", (ndim * n + 2 * n_frames) * 8)
            host = stage.numpy()
            cols = host[:ndim * n * 8].view(np.float64).reshape(ndim, n)
            pieces = [(k, a, min(n, a + (1 << 20))) for k in range(ndim) for a in range(0, n, 1 << 20)]
            with ThreadPoolExecutor(max(1, min(4, host_threads(4), len(pieces)))) as pool:
                list(pool.map(lambda p: np.copyto(cols[p[0], p[1]:p[2]], pos[p[0]][p[1]:p[2]]), pieces))
        bounds = host[ndim * n * 8:(ndim * n + 2 * n_frames) * 8].view(np.int64).reshape(2, n_frames)
        bounds[0], bounds[1] = self.starts, self.stops
        out = _pinned(torch, "labels", (n + n_frames) * 4)
        result = out.numpy().view(np.int32)
        self.labels, self.flags = result[:n], result[n:n + n_frames]
        self.flags[:] = -1                                  # "not labelled yet"
        t1 = time.perf_counter()
        with torch.cuda.device(device):
            self.stream = _STREAMS.get(device)        # one labelling stream per device, kept
            if self.stream is None:
                self.stream = _STREAMS[device] = torch.cuda.Stream(device=device, priority=-1)
            nbytes = _lib.label_frames_scratch_bytes(max_points, ndim, n_frames)
            scratch = _SCRATCH.get(device)
            if scratch is None or scratch.numel() < nbytes:
                _SCRATCH.clear()
                scratch = torch.empty(nbytes, dtype=torch.uint8, device=device)
                _SCRATCH[device] = scratch
                torch.cuda.current_stream(device).synchronize()
            # positions and frame bounds go to the device by DMA (the caller issues this BEFORE the
            # frame uploads: mapped-memory reads of the kernel would crawl behind a gigabyte of
            # queued uploads -- measured at 8 ranks: all frames labelled after 51 ms instead of 12);
            # labels and flags come back through mapped pinned memory (posted writes, other direction)
            d_in = torch.empty(stage.numel(), dtype=torch.uint8, device=device)   # (current stream's pool)
            self.stream.wait_stream(torch.cuda.current_stream(device))
            with torch.cuda.stream(self.stream):
                d_in.copy_(stage, non_blocking=True)
            base, obase = d_in.data_ptr(), out.data_ptr()        # pinned: the same address on the device
            _lib.label_frames_device([base + k * n * 8 for k in range(ndim)], ndim,
                                     base + ndim * n * 8, base + ndim * n * 8 + n_frames * 8,
                                     n_frames, max_points, self.separation, obase, obase + n * 4,
                                     scratch.data_ptr(), nbytes, self.stream.cuda_stream)
        self.keep = (stage, d_in, out, scratch)
        self.t_launch = time.perf_counter()
        self.ms = dict(stage=1e3 * (t1 - t0), launch=1e3 * (self.t_launch - t1), launched_at=self.t_launch)

    def start_async(self):
        """``start()`` on a thread of its own (so that the caller's set-up goes on meanwhile)."""
        import threading
        self.error = None

        def run():
            try:
                self.start()
            except BaseException as exc:
                self.error = exc

        self.thread = threading.Thread(target=run, daemon=True)
        self.thread.start()
        return self

    def join(self):
        thread = getattr(self, 'thread', None)
        if thread is not None:
            thread.join()
            self.thread = None
        if getattr(self, 'error', None) is not None:
            raise self.error

    def close(self):
        """End of the call (also a failed one): the kernel writes into cached mapped buffers the
        next call reuses, so it must have finished."""
        try:
            self.join()
        finally:
            stream = getattr(self, 'stream', None)
            if stream is not None:
                stream.synchronize()

    def wait_frames(self, fa, fb):
        import time
        flags = self.flags[fa:fb]
        while len(flags) and _lib.load().ctk_wait_flags(flags.ctypes.data, len(flags), 20000) != 0:
            if self.stream.query():                      # the kernel has ended (or failed)
                self.stream.synchronize()
                if flags.min() < 0:
                    raise RuntimeError("ctk_label_frames finished without labelling frames %d..%d" % (fa, fb))
        if fb == len(self.flags):
            self.ms['all_frames_labelled'] = 1e3 * (time.perf_counter() - self.t_launch)
        return self.labels, self.flags


def device_labelling_enabled(n_rows, n_frames):
    """Where the cluster labels are computed.  CTK_LABEL_DEVICE=0: always on the host threads;
    2: always on the GPU; default: on the GPU unless the table is so small (a few frames, a few
    thousand features) that the launch would cost more than the host threads need."""
    mode = os.environ.get('CTK_LABEL_DEVICE', '1')
    if mode == '0':
        return False
    return mode == '2' or (n_frames >= 4 and n_rows >= 4096)


class ChunkLabeller(object):
    """Labels a frame-sorted video chunk by chunk on a background thread, so that the caller can
    launch chunk k while chunk k+1 is being labelled.  ``frame_cuts`` = frame indices at which the
    chunks begin and end (len K+1).  While a frame's rows are hot in a worker's cache the worker
    also gathers the packed parameter rows of the frame (``sources`` -> ``params_out``, in
    (cluster, row) order) and writes the frame's group table.  ``get(k)`` waits for chunk k and
    returns, for its rows, (labels local to each frame, cluster sizes, permutation that lists the
    chunk's rows by (frame, label) as indices into the chunk, per-frame label spans, group offsets
    int32 [g + 1], frame index of every group int32 [g]).

    Without the verified native labelling (see ``_native_is_exact``) everything is one chunk
    labelled through scipy and packed with numpy."""

    def __init__(self, pos, starts, stops, frame_cuts, separation, sources, params_out, device=None,
                 size_out=None, label_out=None, device_labels=None):
        import threading
        self.native = os.environ.get('CTK_FIND_NATIVE', '1') != '0' and _native_is_exact()
        self.device_labels = None
        self.flagged_frames = 0
        self.starts, self.stops = np.asarray(starts, np.int64), np.asarray(stops, np.int64)
        self.frame_cuts = list(frame_cuts) if self.native else [0, len(starts)]
        self.results = [None] * (len(self.frame_cuts) - 1)
        self.events = [threading.Event() for _ in self.results]
        self.error = None
        if not self.native:
            if isinstance(pos, (list, tuple)):
                pos = np.stack(pos, axis=1)
            job = _LabelJob(pos, starts, stops, separation)
            job.result()
            rows = job.by_cluster
            for j, src in enumerate(sources):
                params_out[:, j] = src[rows] if isinstance(src, np.ndarray) else src
            labels = job.cluster[rows]
            frame_of_row = np.repeat(np.arange(len(starts), dtype=np.int32), self.stops - self.starts)
            new_group = np.concatenate(([True], (labels[1:] != labels[:-1]) |
                                        (frame_of_row[rows][1:] != frame_of_row[rows][:-1])))
            g_starts = np.flatnonzero(new_group)
            self.results[0] = (job.cluster, job.size, rows, np.asarray(job.spans, np.int64),
                               np.concatenate((g_starts, [len(rows)])).astype(np.int32),
                               frame_of_row[rows[g_starts]])
            self.events[0].set()
            return
        workers = max(1, _pool_workers())
        self.cancelled = False
        early = device_labels is not None and np.array_equal(device_labels.starts, self.starts)
        if early:
            self.device_labels = device_labels            # launched by the caller during its set-up
        elif (device is not None and isinstance(pos, (list, tuple)) and len(starts) > 0
                and len(pos[0]) > 0 and device_labelling_enabled(len(pos[0]), len(starts))):
            # the labels themselves come from the GPU (one warp per frame); the host threads below
            # only count, order and pack
            self.device_labels = DeviceLabels(pos, self.starts, self.stops, separation, device)

        n_rows = int(self.stops[-1]) if len(self.stops) else 0
        order_all = _lib.workspace("by_cluster", n_rows, np.int64)
        gstart_all = _lib.workspace("group_start", n_rows + 1, np.int32)

        def work():
            try:
                labels = flags = None
                if self.device_labels is not None:
                    if early:
                        self.device_labels.join()
                    else:
                        self.device_labels.start()
                for k, (fa, fb) in enumerate(zip(self.frame_cuts[:-1], self.frame_cuts[1:])):
                    if self.cancelled:
                        raise RuntimeError("labelling cancelled")
                    if self.device_labels is not None:
                        labels, flags = self.device_labels.wait_frames(fa, fb)
                        self.flagged_frames += int(np.count_nonzero(flags[fa:fb]))
                    a = int(self.starts[fa]) if fb > fa else 0
                    b = int(self.stops[fb - 1]) if fb > fa else 0
                    st, sp = self.starts[fa:fb] - a, self.stops[fa:fb] - a
                    chunk_pos = pos if isinstance(pos, (list, tuple)) else pos[a:b]
                    local, size, by_cluster, spans, gcount, gstart = _lib.cluster_pack_frames(
                        chunk_pos, st, sp, separation, workers, sources, a, params_out[a:b],
                        labels=None if labels is None else labels[a:b],
                        flags=None if flags is None else flags[fa:fb],
                        cluster_out=None if label_out is None else label_out[a:b],
                        size_out=None if size_out is None else size_out[a:b],
                        by_cluster_out=order_all[a:b], group_start_out=gstart_all[a:b + 1])
                    if size_out is not None:       # written in place: tell the caller not to copy
                        size, local = size_out, label_out
                    goff, gframe = _lib.concat_groups(st, sp, gcount, gstart, fa)
                    self.results[k] = (local, size, by_cluster, spans, goff, gframe)
                    self.events[k].set()
            except Exception as exc:
                self.error = exc
                for ev in self.events:
                    ev.set()

        self.thread = threading.Thread(target=work, daemon=True)
        self.thread.start()

    def __len__(self):
        return len(self.results)

    def close(self):
        """Stop after the chunk in flight and wait for the thread: it writes into the caller's
        (cached, pinned) ``params_out``, which must not be reused while it runs."""
        thread = getattr(self, 'thread', None)
        if thread is not None:
            self.cancelled = True
            thread.join()
            self.thread = None

    def get(self, k):
        self.events[k].wait()
        if self.error is not None:
            raise self.error
        return self.results[k]


def label_frames(pos, starts, stops, separation):
    """Per-frame labels for frame-sorted positions: -> (cluster ids with the running offset of
    find.py:127-128 applied, cluster sizes, permutation that sorts the rows by (frame, cluster)
    keeping the row order inside a cluster -- the group order of refine.py:336)."""
    return _LabelJob(pos, starts, stops, separation).result()


def find_clusters(f, separation, pos_columns=None, t_column='frame'):
    """Group features closer than ``separation`` (number or per-axis tuple) into clusters, frame by
    frame (find.py:132-163).

    Returns a frame-sorted COPY of ``f`` (original index labels, original order inside a frame) with
    int64 columns ``cluster`` (unique over all frames) and ``cluster_size``.
    """
    return cluster_table(f, separation, pos_columns, t_column)[0]


def cluster_table(f, separation, pos_columns=None, t_column='frame'):
    """``find_clusters`` plus the row permutation that lists the result by (frame, cluster)."""
    if pos_columns is None:
        pos_columns = guess_pos_columns(f)
    separation = np.asarray(validate_tuple(separation, len(pos_columns)), dtype=np.float64)
    if t_column in f:
        frames = f[t_column].values
    else:
        frames = np.zeros(len(f), dtype=np.int64)            # find.py:151-161 adds, then deletes it
    if len(frames) > 1 and np.all(frames[1:] >= frames[:-1]):
        order = None                                          # already frame-sorted
        pos = np.ascontiguousarray(f[pos_columns].values, dtype=np.float64)
        sorted_frames = frames
    else:
        order = np.argsort(frames, kind='stable')
        pos = f[pos_columns].values.astype(np.float64)[order]
        sorted_frames = frames[order]
    cuts = np.flatnonzero(sorted_frames[1:] != sorted_frames[:-1]) + 1
    starts = np.concatenate(([0], cuts)).astype(np.int64)
    stops = np.concatenate((cuts, [len(pos)])).astype(np.int64)
    job = _LabelJob(pos, starts, stops, separation)          # may run in the worker pool ...
    out = f.copy() if order is None else f.iloc[order].copy()   # ... while the copy is made
    if t_column not in f:
        out[t_column] = 0                                     # the copies keep the temporary column
    cluster, size, by_cluster = job.result()
    out['cluster'] = cluster
    out['cluster_size'] = size
    return out, by_cluster


# --------------------------------------------------------------------------------------------------
# feature finding: the step in front of the refinement (reference: clustertracking/find.py:166-277)
# --------------------------------------------------------------------------------------------------
def where_close(pos, separation, intensity=None):
    """Indices of features closer than ``separation`` to another feature: of every close pair the
    dimmer one (``intensity`` given) or the one nearer the top left (find.py:166-199).  The pairs
    come from libctk's kd-tree (``ctk_query_pairs_within``); the result does not depend on their
    order."""
    if len(pos) == 0:
        return []
    pos = np.asarray(pos)
    separation = validate_tuple(separation, pos.shape[1])
    if any([s == 0 for s in separation]):
        return []
    pos_rescaled = pos / separation                                          # find.py:179
    pairs = _lib.query_pairs(np.ascontiguousarray(pos_rescaled, dtype=np.float64), 1 - 1e-7)
    if len(pairs) == 0:
        return []
    index_0, index_1 = pairs[:, 0], pairs[:, 1]
    top_left = np.sum(pos_rescaled[index_0], 1) > np.sum(pos_rescaled[index_1], 1)
    if intensity is None:
        to_drop = np.where(top_left, index_1, index_0)
    else:
        intensity = np.asarray(intensity)
        intensity_0, intensity_1 = intensity[index_0], intensity[index_1]
        to_drop = np.where(intensity_0 > intensity_1, index_1, index_0)
        ties = intensity_0 == intensity_1
        to_drop[ties] = np.where(top_left, index_1, index_0)[ties]
    return np.unique(to_drop)


def drop_close(pos, separation, intensity=None):
    """``pos`` without the features ``where_close`` names (find.py:202-207)."""
    return np.delete(pos, where_close(pos, separation, intensity), axis=0)


def _maxima_on_device(frames, size, percentile, margin):
    """Maxima of equally shaped integer frames (a list of arrays or one stacked array) on the current
    CUDA device through ``ctk_find_maxima`` -> (coords int32 [n_frames, capacity, ndim], values
    int32 [n_frames, capacity], counts int32 [n_frames], thresholds float64 [n_frames])."""
    import torch
    lib = _lib.load()
    if not torch.cuda.is_available():
        raise RuntimeError("clustertracking_b200 needs a CUDA device (B200, sm_100a); "
                           "there is no CPU fallback")
    if isinstance(frames, np.ndarray):
        stack = np.ascontiguousarray(frames)                 # already [n_frames, *shape]: no copy
    else:
        stack = np.ascontiguousarray(np.stack([np.asarray(fr) for fr in frames]))
    if stack.dtype not in (np.dtype(np.uint8), np.dtype(np.uint16)):
        raise NotImplementedError("grey_dilation on the GPU takes uint8 or uint16 frames, got %s"
                                  % stack.dtype)
    shape = stack.shape[1:]
    ndim = len(shape)
    if ndim not in (2, 3):
        raise ValueError("only 2D and 3D images are supported")
    dev = torch.device('cuda', torch.cuda.current_device())
    d_stack = torch.from_numpy(stack).to(dev)
    n_frames, n_pixels = len(stack), int(np.prod(shape))
    ptrs = d_stack.data_ptr() + stack[0].nbytes * np.arange(n_frames, dtype=np.int64)
    d_ptrs = torch.from_numpy(ptrs).to(dev)
    code = _lib.PIXEL_CODES[stack.dtype]
    d_ws = torch.empty(int(lib.ctk_find_workspace_bytes(n_frames, n_pixels, code)), dtype=torch.uint8,
                       device=dev)
    shape_arr = (_lib.ctypes.c_int64 * 3)(*(list(shape) + [1] * (3 - ndim)))
    size_arr = (_lib.ctypes.c_int32 * 3)(*(list(size) + [1] * (3 - ndim)))
    margin_arr = (_lib.ctypes.c_int32 * 3)(*(list(margin) + [0] * (3 - ndim)))
    capacity = max(1024, n_pixels // 64)
    while True:
        d_coords = torch.empty((n_frames, capacity, ndim), dtype=torch.int32, device=dev)
        d_values = torch.empty((n_frames, capacity), dtype=torch.int32, device=dev)
        d_count = torch.empty(n_frames, dtype=torch.int32, device=dev)
        d_thr = torch.empty(n_frames, dtype=torch.float64, device=dev)
        rc = lib.ctk_find_maxima(d_ptrs.data_ptr(), n_frames, shape_arr, ndim, code, size_arr,
                                 float(percentile), margin_arr, capacity, d_coords.data_ptr(),
                                 d_values.data_ptr(), d_count.data_ptr(), d_thr.data_ptr(),
                                 d_ws.data_ptr(), int(bool(np.all(ptrs % 4 == 0))),
                                 _lib.ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
        if rc != 0:
            raise RuntimeError("ctk_find_maxima failed (%d): %s"
                               % (rc, lib.ctk_find_last_error().decode()))
        counts = d_count.cpu().numpy()
        if counts.max(initial=0) <= capacity:
            break
        capacity = int(counts.max())
    used = max(1, int(counts.max(initial=0)))
    coords = np.ascontiguousarray(d_coords[:, :used].cpu().numpy())
    values = np.ascontiguousarray(d_values[:, :used].cpu().numpy())
    return coords, values, counts, d_thr.cpu().numpy()


def grey_dilation_batch(frames, separation, percentile=64, margin=None, precise=True):
    """``grey_dilation`` for equally shaped frames (a sequence of arrays or one stacked array
    [n_frames, *shape]): the image work on the GPU in one batch, ``drop_close`` for all frames on
    host threads inside libctk -> list of position arrays."""
    if not isinstance(frames, np.ndarray):
        frames = list(frames)
    if len(frames) == 0:
        return []
    ndim = np.asarray(frames[0]).ndim
    separation = validate_tuple(separation, ndim)
    if margin is None:
        margin = tuple([int(s / 2) for s in separation])                     # find.py:246-247
    margin = validate_tuple(margin, ndim)
    size = [int(2 * s / np.sqrt(ndim)) for s in separation]                   # find.py:255
    coords, values, counts, thresholds = _maxima_on_device(frames, size, percentile, margin)
    if precise:                                                               # find.py:275-276
        keep = _lib.drop_close_frames(coords, values, counts, separation, max(1, _pool_workers()))
    out = []
    for k in range(len(counts)):
        if np.isnan(thresholds[k]) or counts[k] == 0:
            out.append(np.empty((0, ndim)))                                   # find.py:251-252, 262
            continue
        pos = coords[k, :counts[k]]
        if precise:
            pos = pos[keep[k, :counts[k]]]
        out.append(pos.astype(np.int64))
    return out


def grey_dilation(image, separation, percentile=64, margin=None, precise=True):
    """Local maxima brighter than the given percentile of the non-black pixels, at least
    ``separation`` apart -- same signature and result as the reference's ``grey_dilation``
    (find.py:219-277).  The percentile, the dilation and the comparison run on the GPU
    (``ctk_find_maxima``); uint8 and uint16 images."""
    return grey_dilation_batch([image], separation, percentile, margin, precise)[0]
