"""Data-parallel correctness with two ranks (pytest -m gpu).  On a box with >= 2 GPUs the ranks use one GPU each over NCCL; on a
1-GPU box both ranks share cuda:0 and talk over gloo (which all-reduces CUDA tensors through the host), so the driver's single-GPU
run exercises the multi-rank path too:

  * the flat-gradient all-reduce makes both ranks apply the same update, and the all-reduced gradient equals the sum of the two
    per-shard gradients computed on one device (per-replica batch norm, the default);
  * with synchronised batch norm (nets.set_sync_bn) the two-rank step IS the single-device step on the global batch -- the
    reference's semantics (it normalises over the whole batch on one device, nets.py:263-272)."""
import os
import socket
import sys

import pytest
import torch

from _util import ROOT

pytestmark = pytest.mark.gpu
L, B, H, W = 2, 4, 128, 128


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _batch(seed):
    import numpy as np
    from oracle import gen_inputs
    rs = np.random.RandomState(seed)
    s = gen_inputs.scene(1, B, H, W, 'synth', seed, 1.0)
    return dict(imgs_src=rs.uniform(0, 1, (B, H, W, 3)).astype('float32'), imgs_trg=rs.uniform(0, 1, (B, H, W, 3)).astype('float32'),
                k_s=s['k_s'], k_t=s['k_t'], rot_mat=s['rot'], trans_mat=s['t'])


def _worker(rank, world, port, backend, sync_bn, out):
    sys.path.insert(0, os.path.join(ROOT, 'layered-scene-inference_b200'))
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    import torch.distributed as dist
    dev = rank if backend == 'nccl' else 0
    torch.cuda.set_device(dev)
    if backend == 'nccl':
        dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', dev))
    else:
        dist.init_process_group('gloo', rank=rank, world_size=world)
    from lsi.nnutils import nets, train_utils
    nets.set_conv_mode('fp32')
    nets.set_sync_bn(sync_bn)
    opts = train_utils.default_opts(n_layers=L, batch_size=B, img_height=H, img_width=W)
    full = {k: torch.tensor(v, device='cuda') for k, v in _batch(11).items()}
    shard = train_utils.shard_batch(full, rank, world)
    tr = train_utils.Trainer(opts, store=nets.ParamStore(device='cuda', seed=2))
    loss, _, chk = tr.train_step(shard, dp_check=True)
    out[rank] = (tr.store.flat_grad.cpu(), tr.store.flat.cpu(), float(loss), chk)
    dist.destroy_process_group()


def _run_two_ranks(sync_bn):
    import torch.multiprocessing as mp
    backend = 'nccl' if torch.cuda.device_count() >= 2 else 'gloo'
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), backend, sync_bn, out), nprocs=2, join=True)
    return out[0], out[1], backend


def _single_device(shards):
    """One Trainer per entry of `shards` on this device; returns the summed gradients and the last trainer's parameters."""
    from lsi.nnutils import nets, train_utils
    opts = train_utils.default_opts(n_layers=L, batch_size=B, img_height=H, img_width=W)
    total, params = None, None
    for sh in shards:
        tr = train_utils.Trainer(opts, store=nets.ParamStore(device='cuda', seed=2))
        tr.train_step(sh)
        total = tr.store.flat_grad.clone() if total is None else total + tr.store.flat_grad
        params = tr.store.flat.clone()
    return total.cpu(), params.cpu()


def test_two_rank_step_matches_single_device_shards():
    (g0, p0, _, chk0), (g1, p1, _, chk1), backend = _run_two_ranks(False)
    assert torch.equal(g0, g1)               # same all-reduced gradient on both ranks
    assert torch.equal(p0, p1)               # hence bit-identical parameters after Adam
    assert chk0['world'] == 2 and abs(chk0['proj_sum_of_shards'] - chk0['proj_allreduced']) <= 1e-5 * max(abs(chk0['proj_allreduced']), 1e-12)
    sys.path.insert(0, os.path.join(ROOT, 'layered-scene-inference_b200'))
    from lsi.nnutils import nets, train_utils
    nets.set_conv_mode('fp32')
    try:
        full = {k: torch.tensor(v, device='cuda') for k, v in _batch(11).items()}
        total, _ = _single_device([train_utils.shard_batch(full, r, 2) for r in range(2)])
        err = float((total - g0).norm() / g0.norm())
        print('2 ranks over %s: |sum of shard gradients - all-reduced| / |.| = %.3g' % (backend, err))
        assert err < 1e-3, err                # fp32 atomics / summation order only
    finally:
        nets.set_conv_mode('tf32')


def test_sync_bn_two_ranks_equal_single_device_global_batch():
    (g0, p0, l0, _), (g1, p1, l1, _), backend = _run_two_ranks(True)
    assert torch.equal(g0, g1) and torch.equal(p0, p1)
    sys.path.insert(0, os.path.join(ROOT, 'layered-scene-inference_b200'))
    from lsi.nnutils import nets
    nets.set_conv_mode('fp32')
    try:
        full = {k: torch.tensor(v, device='cuda') for k, v in _batch(11).items()}
        g_full, p_full = _single_device([full])           # the reference's step: whole batch, one device
        # each rank's loss is the mean over its shard: the all-reduced sum is twice the gradient of the global-batch mean
        err_g = float((0.5 * g0 - g_full).norm() / g_full.norm())
        err_p = float((p0 - p_full).abs().max())
        print('sync BN, 2 ranks over %s vs single device, global batch %d: gradient rel err %.3g, max |d param| %.3g' % (backend, B, err_g, err_p))
        assert err_g < 2e-3, err_g
        assert err_p < 2.5e-4, err_p            # one Adam step moves every parameter by at most lr = 1e-4
        # without synchronisation the statistics differ per replica and the step is a different function
        (gn, _, _, _), _, _ = _run_two_ranks(False)
        assert float((0.5 * gn - g_full).norm() / g_full.norm()) > 10 * max(err_g, 1e-6)
    finally:
        nets.set_conv_mode('tf32')
