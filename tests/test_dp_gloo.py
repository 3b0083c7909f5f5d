"""CPU, world_size 2 over gloo: the host-side logic of the data-parallel step -- batch sharding and the single
gradient all-reduce on the flat buffer (the CUDA kernels themselves are covered by the -m gpu tests)."""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from _util import ROOT


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    sys.path.insert(0, os.path.join(ROOT, 'layered-scene-inference_b200'))
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from lsi.nnutils import train_utils
    batch = {'imgs_src': torch.arange(8 * 3, dtype=torch.float32).reshape(8, 3), 'k_s': torch.arange(8.0).reshape(8, 1)}
    shard = train_utils.shard_batch(batch, rank, world)
    flat = torch.full((1000,), float(rank + 1))
    n = train_utils.allreduce_sum_(flat)
    out[rank] = (shard['imgs_src'][:, 0].tolist(), shard['k_s'].flatten().tolist(), flat[:3].tolist(), n)
    dist.destroy_process_group()


def test_shard_and_allreduce_world2():
    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert out[0][0] == [0.0, 3.0, 6.0, 9.0] and out[1][0] == [12.0, 15.0, 18.0, 21.0]      # disjoint, complete, ordered
    assert out[0][1] + out[1][1] == [float(i) for i in range(8)]
    assert out[0][2] == [3.0, 3.0, 3.0] and out[1][2] == [3.0, 3.0, 3.0] and out[0][3] == 2  # 1 + 2 on both ranks


def test_allreduce_is_noop_without_process_group():
    sys.path.insert(0, os.path.join(ROOT, 'layered-scene-inference_b200'))
    from lsi.nnutils import train_utils
    flat = torch.ones(4)
    assert train_utils.allreduce_sum_(flat) == 1 and flat.tolist() == [1.0] * 4
    import pytest
    with pytest.raises(ValueError):
        train_utils.shard_batch({'x': torch.zeros(5, 2)}, 0, 2)
