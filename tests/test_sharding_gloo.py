"""World-size-2 test of the frame sharding (gloo backend, CPU).  The per-shard refinement is done by
the one-lane host build of the device solver (test infrastructure, see tests/emul): the subject of
this test is the host-side sharding, the no-collective data path and the final gather, which must
reproduce the single-process result exactly, cluster ids included."""
import os
import socket
import sys

import numpy as np
import pandas as pd
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _video(n_frames=4):
    sys.path.insert(0, ROOT)
    from clustertracking_b200 import artificial
    stack, rows = [], []
    for t in range(n_frames):
        frame, f0, _ = artificial.clustered_frame((96, 96), pitch=44, seed=50 + t)
        f0['frame'] = t + 3                      # frame numbers need not start at 0
        stack.append(frame)
        rows.append(f0)
    return artificial.FrameStack(np.array(stack), first_frame=3), pd.concat(rows, ignore_index=True)


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    import emul_backend
    from clustertracking_b200 import parallel, refine
    refine.FrameSet = _NoFrames          # no GPU here: the emulated solver reads the host frames
    refine._pinned_buffer = lambda torch, key, nbytes: torch.empty(max(nbytes, 1), dtype=torch.uint8)
    refine.launch_cuda = _emulated_launch
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank,
                            world_size=world)
    reader, f0 = _video()
    # (1) every rank passes the whole table, every rank gets the merged result (shared-memory gather)
    out = parallel.refine_leastsq_sharded(f0, reader, 11)
    out.to_pickle(os.path.join(out_dir, "rank%d.pkl" % rank))
    # (2) the same through the tensor gather (what runs across hosts)
    os.environ['CTK_GATHER'] = 'tensors'
    out = parallel.refine_leastsq_sharded(f0, reader, 11)
    out.to_pickle(os.path.join(out_dir, "tensors%d.pkl" % rank))
    del os.environ['CTK_GATHER']
    # (3) every rank passes only ITS frames; the merged table goes to rank 0 alone
    mine = parallel.frame_shard(f0, rank, world)
    for mode in ('shm', 'copy', 'tensors'):     # zero-copy shared block | copy into one | tensors
        if mode != 'shm':
            os.environ['CTK_GATHER'] = mode
        out = parallel.refine_leastsq_sharded(mine, reader, 11, presharded=True, gather='root')
        assert (out is None) == (rank != 0)
        if out is not None:
            out.to_pickle(os.path.join(out_dir, "root_%s.pkl" % mode))
    os.environ.pop('CTK_GATHER', None)
    # (5) a global-level fit over both ranks: the accumulator of every pass is all-reduced
    out = parallel.refine_leastsq_sharded(mine, reader, 11, presharded=True, gather='root',
                                          param_mode=dict(signal='var', size='global'),
                                          passes_factory=emul_backend.GlobalPasses)
    if out is not None:
        out.to_pickle(os.path.join(out_dir, "global.pkl"))
    # (6) the shared result block is reused by the next call -- but only when no array of the
    # previous result is alive on any rank
    import json
    keep = parallel.refine_leastsq_sharded(f0, reader, 11)
    snapshot = keep.copy()
    second = parallel.refine_leastsq_sharded(f0.assign(signal=f0['signal'] * 1.5), reader, 11)
    reused_while_alive = bool(parallel.LAST_GATHER['block_reused'])
    untouched = all(np.array_equal(keep[c].values, snapshot[c].values, equal_nan=True) for c in keep.columns)
    del keep, second
    third = parallel.refine_leastsq_sharded(f0, reader, 11)
    reused_when_dead = bool(parallel.LAST_GATHER['block_reused'])
    third.to_pickle(os.path.join(out_dir, "reused%d.pkl" % rank))
    with open(os.path.join(out_dir, "reuse%d.json" % rank), "w") as fh:
        json.dump(dict(reused_while_alive=reused_while_alive, untouched=untouched,
                       reused_when_dead=reused_when_dead), fh)
    del third
    # (4) no gather: every rank keeps its part, cluster ids already running on across ranks
    out = parallel.refine_leastsq_sharded(mine, reader, 11, presharded=True, gather='none')
    out.to_pickle(os.path.join(out_dir, "part%d.pkl" % rank))
    dist.destroy_process_group()


class _NoFrames(object):
    h2d_bytes = launches = 0

    def __init__(self, info, device=None):
        pass

    def upload_async(self):
        return self

    def close(self):
        pass


class _Session(object):
    h2d_bytes = d2h_bytes = launches = 0


class _Done(object):
    def __init__(self, result):
        self._result = result

    def result(self):
        return self._result


def _emulated_launch(plan, frames, out_params, out_cost, out_status, stream=None):
    """Stand-in for ``refine.launch_cuda``: the one-lane host build of the device solver."""
    import emul_backend
    result = emul_backend.execute(plan)
    out_params[...], out_cost[...], out_status[...] = result.params_out, result.cost, result.status
    result.params_out, result.cost, result.status = out_params, out_cost, out_status
    result.session = _Session()
    return _Done(result)


def test_two_rank_sharding_matches_single_process(tmp_path):
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import emul_backend
    emul_backend.lib()                           # build once, before the workers start
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    read = lambda name: pd.read_pickle(os.path.join(str(tmp_path), name))
    parts = [read("rank%d.pkl" % r) for r in range(2)] + [read("tensors%d.pkl" % r) for r in range(2)]
    parts += [read("root_shm.pkl"), read("root_copy.pkl"), read("root_tensors.pkl"),
              pd.concat([read("part%d.pkl" % r) for r in range(2)])]
    parts += [read("reused%d.pkl" % r) for r in range(2)]
    import json
    for r in range(2):
        with open(os.path.join(str(tmp_path), "reuse%d.json" % r)) as fh:
            flags = json.load(fh)
        assert flags == dict(reused_while_alive=False, untouched=True, reused_when_dead=True), flags
    reader, f0 = _video()
    single, _ = emul_backend.refine_leastsq(f0, reader, 11)
    for part in parts:                           # every variant: the full, identical result
        assert list(part.columns) == list(single.columns)
        assert np.array_equal(part.index.values, single.index.values)
        for col in single.columns:
            assert part[col].dtype == single[col].dtype, col
            assert np.array_equal(part[col].values, single[col].values, equal_nan=True), col


def test_sharded_global_fit_matches_single_process(tmp_path):
    """param_mode 'global' over two ranks (frames sharded, Schur accumulators all-reduced) gives
    the single-process answer."""
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import emul_backend
    emul_backend.lib()
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    got = pd.read_pickle(os.path.join(str(tmp_path), "global.pkl"))
    reader, f0 = _video()
    single = emul_backend.refine_leastsq_global(f0, reader, 11, param_mode=dict(signal='var', size='global'))
    assert np.array_equal(got.index.values, single.index.values)
    assert np.array_equal(got['cluster'].values, single['cluster'].values)
    assert np.ptp(got['size'].values) == 0.
    for col in ('y', 'x', 'signal', 'size', 'background', 'cost'):
        assert np.allclose(got[col].values, single[col].values, rtol=1e-7, atol=1e-7), col


def test_shard_bounds_and_frame_shard():
    sys.path.insert(0, ROOT)
    from clustertracking_b200 import parallel
    assert list(parallel.shard_bounds(10, 4)) == [0, 3, 6, 8, 10]
    assert list(parallel.shard_bounds(2, 4)) == [0, 1, 2, 2, 2]
    f = pd.DataFrame(dict(frame=[5, 5, 7, 9, 9, 9, 12], x=np.arange(7.)))
    got = [list(parallel.frame_shard(f, r, 3)['frame'].unique()) for r in range(3)]
    assert got == [[5, 7], [9], [12]]
    assert len(parallel.frame_shard(f.iloc[:1], 1, 2)) == 0
