"""The product's sharded path on real devices (``-m gpu``): two ranks refine the two halves of ONE
video through ``parallel.refine_leastsq_sharded`` and the merged table must equal the single-GPU
call exactly (rows, order, cluster ids, numbers).  With two GPUs the ranks use one each over NCCL;
on a one-GPU box both ranks share cuda:0 and the process group is gloo (NCCL refuses two ranks on
one device) -- the kernels, the frame sharding and the gather are the same."""
import os
import socket
import sys

import numpy as np
import pandas as pd
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _video():
    sys.path.insert(0, ROOT)
    from clustertracking_b200 import artificial
    return artificial.clustered_video(6, shape=(256, 256), seed=70)


def _worker(rank, world, port, out_dir, nccl):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from clustertracking_b200 import parallel
    torch.cuda.set_device(rank if nccl else 0)
    os.environ['LOCAL_WORLD_SIZE'] = str(world)
    dist.init_process_group("nccl" if nccl else "gloo", init_method="tcp://127.0.0.1:%d" % port,
                            rank=rank, world_size=world)
    reader, f0 = _video()
    whole = parallel.refine_leastsq_sharded(f0, reader, 11)                   # whole table in, all out
    whole.to_pickle(os.path.join(out_dir, "all%d.pkl" % rank))
    mine = parallel.frame_shard(f0, rank, world)
    root = parallel.refine_leastsq_sharded(mine, reader, 11, presharded=True, gather='root')
    assert (root is None) == (rank != 0)
    if root is not None:
        root.to_pickle(os.path.join(out_dir, "root.pkl"))
    os.environ['CTK_GATHER'] = 'tensors'                                      # the cross-host transport
    root = parallel.refine_leastsq_sharded(mine, reader, 11, presharded=True, gather='root')
    if root is not None:
        root.to_pickle(os.path.join(out_dir, "root_tensors.pkl"))
    dist.destroy_process_group()


def test_two_ranks_equal_single_gpu(tmp_path):
    import torch
    import torch.multiprocessing as mp
    import clustertracking_b200 as ctb
    nccl = torch.cuda.device_count() >= 2
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path), nccl), nprocs=2, join=True)
    reader, f0 = _video()
    single = ctb.refine_leastsq(f0, reader, 11)
    names = ["all0.pkl", "all1.pkl", "root.pkl", "root_tensors.pkl"]
    for name in names:
        part = pd.read_pickle(os.path.join(str(tmp_path), name))
        assert list(part.columns) == list(single.columns), name
        assert np.array_equal(part.index.values, single.index.values), name
        for col in single.columns:
            assert part[col].dtype == single[col].dtype, (name, col)
            assert np.array_equal(part[col].values, single[col].values, equal_nan=True), (name, col)
