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
SYNC_HW = 256          # the sync-BN comparison runs at 256x256: at 128x128 the 1x1 bottleneck normalises over 4 samples and ReLU-mask
                       # flips turn 1e-7 summation-order differences into percent-level gradient changes


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _batch(seed, H=H, W=W):
    import numpy as np
    from oracle import gen_inputs
    rs = np.random.RandomState(seed)
    s = gen_inputs.scene(1, B, H, W, 'synth', seed, 1.0)
    return dict(imgs_src=rs.uniform(0, 1, (B, H, W, 3)).astype('float32'), imgs_trg=rs.uniform(0, 1, (B, H, W, 3)).astype('float32'),
                k_s=s['k_s'], k_t=s['k_t'], rot_mat=s['rot'], trans_mat=s['t'])


def _layer_case(nets, x, w, beta, g):
    """conv 3x3 + batch-stat BN + ReLU on x, loss = <y, g>: returns (y, dx, dw, dbeta)."""
    store = nets.ParamStore()
    store.load_state_dict({'t/weights': w, 't/BatchNorm/beta': beta})
    xg = x.clone().requires_grad_(True)
    y = nets._conv_layer(store, 't', xg, w.shape[3], 3, 1, reuse=True)
    dx, dw, db = torch.autograd.grad((y * g).sum(), [xg, store.vars['t/weights'], store.vars['t/BatchNorm/beta']])
    return y.detach(), dx, dw, db


def _layer_inputs():
    gen = torch.Generator().manual_seed(5)
    x = torch.randn(4, 12, 10, 32, generator=gen) + 3.0        # mean >> spread: the case fp32 E[x^2] - mean^2 gets wrong
    w = torch.randn(3, 3, 32, 64, generator=gen) / 17.0
    beta = torch.randn(64, generator=gen) * 0.3
    g = torch.randn(4, 12, 10, 64, generator=gen)
    return x, w, beta, g


def _worker(rank, world, port, backend, sync_bn, out, hw=H):
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
    if hw == 0:          # layer-level case: each rank holds half of the batch
        x, w, beta, g = (t.cuda() for t in _layer_inputs())
        sl = slice(rank * 2, rank * 2 + 2)
        y, dx, dw, db = _layer_case(nets, x[sl].contiguous(), w, beta, g[sl].contiguous())
        for t in (dw, db):
            dist.all_reduce(t)                 # what the trainer's flat-gradient all-reduce does
        out[rank] = (y.cpu(), dx.cpu(), dw.cpu(), db.cpu())
        dist.destroy_process_group()
        return
    opts = train_utils.default_opts(n_layers=L, batch_size=B, img_height=hw, img_width=hw)
    full = {k: torch.tensor(v, device='cuda') for k, v in _batch(11, hw, hw).items()}
    shard = train_utils.shard_batch(full, rank, world)
    tr = train_utils.Trainer(opts, store=nets.ParamStore(device='cuda', seed=2), sync_bn=sync_bn)
    loss, _, chk = tr.train_step(shard, dp_check=True)
    out[rank] = (tr.store.flat_grad.cpu(), tr.store.flat.cpu(), float(loss), chk)
    dist.destroy_process_group()


def _run_two_ranks(sync_bn, hw=H):
    import torch.multiprocessing as mp
    backend = 'nccl' if torch.cuda.device_count() >= 2 else 'gloo'
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), backend, sync_bn, out, hw), nprocs=2, join=True)
    return out[0], out[1], backend


def _single_device(shards, hw=H):
    """One Trainer per entry of `shards` on this device; returns the summed gradients and the last trainer's parameters."""
    from lsi.nnutils import nets, train_utils
    opts = train_utils.default_opts(n_layers=L, batch_size=B, img_height=hw, img_width=hw)
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


def test_sync_bn_layer_two_ranks_equal_single_device():
    """One conv + BN + ReLU layer, batch of 4 split over 2 ranks with synchronised statistics == the same layer on one device:
    outputs, data gradients (per shard) and the all-reduced weight / beta gradients."""
    (y0, dx0, dw0, db0), (y1, dx1, dw1, db1), backend = _run_two_ranks(True, hw=0)
    sys.path.insert(0, os.path.join(ROOT, 'layered-scene-inference_b200'))
    from lsi.nnutils import nets
    nets.set_conv_mode('fp32')
    try:
        x, w, beta, g = (t.cuda() for t in _layer_inputs())
        y, dx, dw, db = (t.cpu() for t in _layer_case(nets, x, w, beta, g))
        rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
        errs = [rel(torch.cat([y0, y1]), y), rel(torch.cat([dx0, dx1]), dx), rel(dw0, dw), rel(db0, db)]
        print('sync BN layer over %s: y %.2g dx %.2g dw %.2g dbeta %.2g' % ((backend,) + tuple(errs)))
        assert max(errs) < 2e-5, errs
        assert torch.equal(dw0, dw1) and torch.equal(db0, db1)
        # per-replica statistics are a different function
        (yn0, _, _, _), (yn1, _, _, _), _ = _run_two_ranks(False, hw=0)
        assert rel(torch.cat([yn0, yn1]), y) > 1e-3
    finally:
        nets.set_conv_mode('tf32')


def test_sync_bn_two_ranks_equal_single_device_global_batch():
    (g0, p0, l0, _), (g1, p1, l1, _), backend = _run_two_ranks(True, hw=SYNC_HW)
    assert torch.equal(g0, g1) and torch.equal(p0, p1)
    sys.path.insert(0, os.path.join(ROOT, 'layered-scene-inference_b200'))
    from lsi.nnutils import nets
    nets.set_conv_mode('fp32')
    try:
        full = {k: torch.tensor(v, device='cuda') for k, v in _batch(11, SYNC_HW, SYNC_HW).items()}
        g_full, p_full = _single_device([full], hw=SYNC_HW)           # the reference's step: whole batch, one device
        # each rank's loss is the mean over its shard: the all-reduced sum is twice the gradient of the global-batch mean
        err_g = float((0.5 * g0 - g_full).norm() / g_full.norm())
        err_p = float((p0 - p_full).abs().max())
        print('sync BN, 2 ranks over %s vs single device, global batch %d at %dx%d: gradient rel err %.3g, max |d param| %.3g'
              % (backend, B, SYNC_HW, SYNC_HW, err_g, err_p))
        (gn, _, _, _), _, _ = _run_two_ranks(False, hw=SYNC_HW)
        err_n = float((0.5 * gn - g_full).norm() / g_full.norm())
        print('per-replica BN instead: gradient rel err %.3g' % err_n)
        # the exact equivalence is asserted at layer level above; through ~20 BN layers with 16-sample statistics at the bottleneck,
        # summation-order differences flip ReLU masks (the oracle's own fp32 vs fp64 gradients differ by 2.5e-2 on such sizes)
        assert err_g < 5e-2, err_g
        assert err_n > 5 * err_g, (err_n, err_g)
    finally:
        nets.set_conv_mode('tf32')
