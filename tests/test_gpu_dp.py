"""2-GPU data-parallel correctness (pytest -m gpu on a box with >= 2 GPUs; skipped otherwise): the flat-gradient
all-reduce over NCCL makes both ranks apply the same update, and the averaged gradient equals the mean of the two
per-shard gradients computed on one device."""
import os
import socket
import sys

import pytest
import torch

from _util import ROOT

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _batch(B, H, W, seed):
    import numpy as np
    from oracle import gen_inputs
    rs = np.random.RandomState(seed)
    s = gen_inputs.scene(1, B, H, W, 'synth', seed, 1.0)
    return dict(imgs_src=rs.uniform(0, 1, (B, H, W, 3)).astype('float32'), imgs_trg=rs.uniform(0, 1, (B, H, W, 3)).astype('float32'),
                k_s=s['k_s'], k_t=s['k_t'], rot_mat=s['rot'], trans_mat=s['t'])


def _worker(rank, world, port, out):
    sys.path.insert(0, os.path.join(ROOT, 'layered-scene-inference_b200'))
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    from lsi.nnutils import nets, train_utils
    nets.set_conv_mode('fp32')
    L, B, H, W = 2, 4, 128, 128
    opts = train_utils.default_opts(n_layers=L, batch_size=B, img_height=H, img_width=W)
    full = {k: torch.tensor(v, device='cuda') for k, v in _batch(B, H, W, 11).items()}
    shard = train_utils.shard_batch(full, rank, world)
    tr = train_utils.Trainer(opts, store=nets.ParamStore(device='cuda', seed=2))
    before = None
    loss, _ = tr.train_step(shard)
    g = tr.store.flat_grad.clone()          # summed over ranks by the all-reduce inside train_step
    p = tr.store.flat.clone()
    out[rank] = (g.cpu(), p.cpu(), float(loss))
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs')
def test_two_rank_step_matches_single_device_shards():
    import torch.multiprocessing as mp
    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    g0, p0, _ = out[0]
    g1, p1, _ = out[1]
    assert torch.equal(g0, g1)               # same all-reduced gradient on both ranks
    assert torch.equal(p0, p1)               # hence the same parameters after Adam
    # reference: both shards on one device, gradients summed
    sys.path.insert(0, os.path.join(ROOT, 'layered-scene-inference_b200'))
    from lsi.nnutils import nets, train_utils
    nets.set_conv_mode('fp32')
    try:
        L, B, H, W = 2, 4, 128, 128
        opts = train_utils.default_opts(n_layers=L, batch_size=B, img_height=H, img_width=W)
        full = {k: torch.tensor(v, device='cuda') for k, v in _batch(B, H, W, 11).items()}
        total = None
        for r in range(2):
            tr = train_utils.Trainer(opts, store=nets.ParamStore(device='cuda', seed=2))
            tr.train_step(train_utils.shard_batch(full, r, 2))
            total = tr.store.flat_grad.clone() if total is None else total + tr.store.flat_grad
        err = float((total.cpu() - g0).norm() / g0.norm())
        assert err < 1e-3, err                # fp32 atomics / summation order only
    finally:
        nets.set_conv_mode('tf32')
