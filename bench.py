#!/usr/bin/env python
"""bench.py -- rendered views/sec of the B200-native LDI view-synthesis path at BASELINE.json's headline configuration
(256x832, 4-layer LDI, batch 64 per GPU): encoder-decoder CNN -> per-layer (texture, disparity) -> forward-splat
renderer, with the splat kernel's HBM roofline, the conv kernels' tensor throughput, the reference CPU path timed on
the box's host cores, and the end-to-end (host images in, rendered views out) number.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A "step" = one pass of the hot path over one batch of B synthetic source images already resident in HBM: the U-Net
trunk + L heads predict the LDI (ldi_enc_dec.py:196-213), lsi.geometry.ldi.forward_splat renders it into the target
camera (ldi_enc_dec.py:307-318, compose_layers=True).  Rank 0 prints ONE JSON line.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (os.path.join(ROOT, 'layered-scene-inference_b200'), ROOT):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np
import torch

H, W, L, B_PER_GPU = 256, 832, 4, 64
MAX_DISP, BG_DISP, ZBUF_SCALE, DS = 0.4, 1e-3, 50.0, 1.0      # kitti constants, ldi_enc_dec.py:421-425
METRIC = 'rendered views/sec at 256x832x4-layer'
NCU_SPLAT_DRAM_BYTES_PER_VIEW = (238.452736e6 + 10.389248e6) / 14     # profiles/r1_splat_stream_ncu_summary.txt
WORKLOAD = ('KITTI-like 256x832 image -> encoder-decoder U-Net + 4 LDI heads (W zero-padded to 896 for the U-Net, '
            'prediction cropped; tcgen05 convs with fp16 operands/activations and fp32 accumulation, batch-stat BN) -> forward_splat(compose_layers=True, '
            'trg_downsampling=1) -> rendered target view; batch %d per GPU' % B_PER_GPU)


def bytes_fwd_per_view(has_mask):
    """reads tex(3)+disp(1)(+mask(1)) per source pixel per layer, writes img(3)+wts(1) per
    target pixel (trg_disp is not requested on the training path)."""
    n, nt = H * W, int(H * DS) * int(W * DS)
    return 4 * ((5 if has_mask else 4) * L * n + 4 * nt)


def load_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    return 6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s)'


class ClockSampler(object):
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.rows, self.proc, self.thread = index, [], None, None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '100'], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except OSError:
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        self.thread.join(timeout=2)
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace('.', '').isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace('.', '').isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = sorted({n for r in self.rows if len(r) >= 7 for n, v in zip(names, r[3:7]) if v.lower().startswith('active')})
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': reasons, 'samples': len(sm)}


def make_inputs(batch, seed):
    """Synthetic KITTI-like LDIs: 8 procedurally generated scenes, tiled to the batch with a
    per-copy disparity scale so that no two views scatter identically."""
    from oracle import gen_inputs
    uniq = min(batch, 8)
    s = gen_inputs.scene(L, uniq, H, W, 'kitti', seed, MAX_DISP)
    reps = (batch + uniq - 1) // uniq
    out = {}
    for k, v in s.items():
        axis = 1 if k in ('tex', 'mask', 'disp') else 0
        out[k] = np.concatenate([v] * reps, axis=axis)
        out[k] = np.ascontiguousarray(out[k][:, :batch] if axis == 1 else out[k][:batch])
    scale = (1.0 - 0.01 * (np.arange(batch) // uniq)).astype(np.float32)
    out['disp'] = out['disp'] * scale[None, :, None, None, None]
    return out


def _oracle_view_fn(views):
    """The reference CPU path for `views` views: oracle CNN (lsi_oracle_nets) + oracle renderer (lsi_oracle), same
    padding policy as the B200 path."""
    from oracle import lsi_oracle as O
    from oracle import lsi_oracle_nets as N
    rs = np.random.RandomState(0)
    img = torch.tensor(rs.uniform(0, 1, (views, H, W, 3)).astype(np.float32))
    wp = -(-W // 128) * 128
    padded = torch.zeros(views, H, wp, 3)
    padded[:, :, :W] = img
    params = N.init_params(L, seed=0)
    s = make_inputs(views, 0)
    cam = [torch.tensor(s[k]) for k in ('k_s', 'k_t', 'rot', 't')]
    pc = O.pixel_coords(views, H, W)
    kw = dict(compose_layers=True, trg_downsampling=1, bg_layer_disp=BG_DISP, max_disp=MAX_DISP, zbuf_scale=ZBUF_SCALE)

    def fn():
        with torch.no_grad():
            tex, masks, disps = N.predict_ldi(params, padded, L, MAX_DISP)
            ldi = (tex[:, :, :, :W].contiguous(), masks[:, :, :, :W].contiguous(), disps[:, :, :, :W].contiguous())
            return O.forward_splat(ldi, pc, *cam, **kw)
    return fn


def run_reference(args, rank):
    """--impl reference: the reference's own CPU algorithm (TF-1.4 is not installable here, so this is the op-for-op
    CPU restatement under oracle/, `kind: port`) on the host cores, bounded sample of the same workload."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    views = 2                                    # bounded sample: 2 of the 64 views per step
    fn = _oracle_view_fn(views)
    steps, warm = max(1, min(args.steps, 3)), max(1, min(args.warmup, 1))
    for _ in range(warm):
        fn()
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    dt = (time.perf_counter() - t0) / steps
    v = views / dt
    sample = '%d of %d views per step, %d steps, torch CPU %d threads (oracle CNN + oracle renderer)' % (views, B_PER_GPU, steps, cores)
    print(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': 'views/s', 'n_gpus': args.gpus, 'steps': steps,
        'warmup': warm, 'ms_per_step': dt * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic', 'config': {'workload': WORKLOAD, 'sample': sample},
        'cpu_baseline': {'value': v, 'unit': 'views/s', 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': v, 'unit': 'views/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}))


def cpu_baseline():
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    views = 2
    fn = _oracle_view_fn(views)
    fn()
    reps, t0 = 0, time.perf_counter()
    while reps < 2 or (time.perf_counter() - t0 < 12.0 and reps < 20):
        fn()
        reps += 1
    dt = (time.perf_counter() - t0) / reps
    return {'value': views / dt, 'unit': 'views/s', 'cores': cores, 'kind': 'port',
            'sample': '%d of %d views x %d reps of the oracle (lsi_oracle_nets.predict_ldi + lsi_oracle.forward_splat in the '
                      'reference decomposition), torch CPU fp32' % (views, B_PER_GPU, reps)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if args.impl == 'reference':
        run_reference(args, rank)
        return
    args.warmup = max(args.warmup, 3)

    assert torch.cuda.is_available(), 'bench.py needs a CUDA device (no CPU fallback)'
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=dev)

    from lsi import _b200
    from lsi.geometry import ldi as ldi_utils
    from lsi.nnutils import helpers, nets, train_utils
    lib = _b200.lib()

    B = B_PER_GPU
    host = make_inputs(B, seed=rank)
    rs = np.random.RandomState(100 + rank)
    from oracle import gen_inputs
    img_host = np.stack([gen_inputs.band_limited(rs, (H, W), 3) for _ in range(8)] * (B // 8)).astype(np.float32)
    imgs = torch.tensor(img_host, device=dev)
    cam = [torch.tensor(host[k], device=dev) for k in ('k_s', 'k_t', 'rot', 't')]
    pc = helpers.pixel_coords(B, H, W, device=dev)
    opts = train_utils.default_opts(dataset='kitti', n_layers=L, batch_size=B, img_height=H, img_width=W,
                                    zbuf_scale=ZBUF_SCALE)
    # inference-only conv mode: fp16 activations in HBM + kind::f16 tcgen05 MMAs, fp32 accumulation and statistics (same
    # mantissa as a TF32 operand; the bench contract asks for >= bf16).  With autograd enabled (the training step measured
    # below) the mode is TF32.  BENCH_CONV_MODE=tf32 reproduces the all-fp32-activation numbers.
    nets.set_conv_mode(os.environ.get('BENCH_CONV_MODE', 'split'))
    store = nets.ParamStore(device=dev, seed=0)           # random-init weights of the reference architecture
    kw = dict(compose_layers=True, trg_downsampling=DS, bg_layer_disp=BG_DISP, max_disp=MAX_DISP, zbuf_scale=ZBUF_SCALE)
    with torch.no_grad():
        train_utils.predict_ldi(imgs[:1], opts, store, reuse=False)      # creates the variables

    def step(x=None):
        with torch.no_grad():
            ldi = train_utils.predict_ldi(imgs if x is None else x, opts, store, reuse=True)
            return ldi_utils.forward_splat(tuple(ldi), pc, *cam, **kw)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    n0 = _b200.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    launches = _b200.launch_count() - n0
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop()
    if dist is not None:
        tt = torch.tensor([ms], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = tt.item()
    ms_per_step = ms / args.steps
    value = world * B / (ms_per_step * 1e-3)

    # --- per-kernel timing (CUDA events on the launching stream, live) -------------------------------------------
    lib.lsi_b200_kernel_timing_enable(1)
    for _ in range(args.steps):
        step()
    torch.cuda.synchronize()
    kms, kn = (ctypes.c_double * 8)(), (ctypes.c_int * 8)()
    _b200.call('lsi_b200_kernel_timing_collect', ctypes.cast(kms, ctypes.c_void_p), ctypes.cast(kn, ctypes.c_void_p))
    lib.lsi_b200_kernel_timing_enable(0)
    peak, peak_src = load_peaks()
    n_src = H * W
    splat_bytes_per_step = 4.0 * 4 * L * n_src * B          # packed (r,g,b,disp) head output read once by the splat kernel
    splat_ms = kms[0] / args.steps
    achieved = splat_bytes_per_step / (splat_ms * 1e-3) / 1e9
    roofline = {'bound': 'hbm', 'kernel': 'splat_fwd_stream_kernel (forward splat; %d launches per step)' % (kn[0] // args.steps),
                'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak, 'peak_source': peak_src,
                # dram__bytes_read.sum + dram__bytes_write.sum of one launch (14 views) in profiles/r1_splat_stream_ncu_summary.txt,
                # scaled to the views of one launch here; ncu flushes L2 before each replay, so the 3.4 MB/view accumulator that
                # is L2-resident in a real run (memset just before) is re-fetched and shows up as extra reads
                'traffic': NCU_SPLAT_DRAM_BYTES_PER_VIEW * B / max(kn[0] // args.steps, 1),
                'algorithmic_bytes_per_launch': splat_bytes_per_step / max(kn[0] // args.steps, 1),
                'algorithmic_bytes_per_step': splat_bytes_per_step, 'kernel_ms_per_step': splat_ms,
                'normalize_ms_per_step': kms[1] / args.steps}
    wp = -(-W // 128) * 128
    conv_flops = (21.8e9 + 23.0e9 * L) * (H * wp) / (256.0 * 768.0) * B      # forward 2*MAC per step
    conv_ms = (kms[4] + kms[5]) / args.steps
    conv = {'bound': 'tensor', 'kernel': 'conv_tc_kernel + conv_halo_kernel (tcgen05 kind::%s, TMA) + fp32 stem' % ('f16' if nets.get_conv_mode() == 'f16' else 'tf32'), 'achieved': conv_flops / (conv_ms * 1e-3) / 1e12,
            'unit': 'TFLOP/s', 'kernel_ms_per_step': conv_ms, 'tc_ms_per_step': kms[4] / args.steps,
            'fp32_ms_per_step': kms[5] / args.steps, 'flops_per_step': conv_flops,
            'share_of_step': conv_ms / ms_per_step, 'note': 'dense fp16 peak = the measured bf16 figure in MEASURED_PEAKS.json (1653 TFLOP/s burst); TF32 about half of it'}

    # --- renderer slice alone (the kernel the roofline is about), LDIs resident in HBM ------------------------------
    with torch.no_grad():
        ldi_fixed = [t.contiguous() for t in train_utils.predict_ldi(imgs, opts, store, reuse=True)]
    ldi_fixed[1]._lsi_all_ones = True
    for _ in range(3):
        ldi_utils.forward_splat(tuple(ldi_fixed), pc, *cam, **kw)
    torch.cuda.synchronize()
    r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    r0.record()
    for _ in range(20):
        with torch.no_grad():
            ldi_utils.forward_splat(tuple(ldi_fixed), pc, *cam, **kw)
    r1.record()
    torch.cuda.synchronize()
    renderer_only = B / (r0.elapsed_time(r1) / 20 * 1e-3)
    del ldi_fixed

    # --- the same kernel on SURVEY.md 8(d)'s structured config-4 LDI (road-plane disparity ramps + smooth bumps, what a
    #     trained network predicts).  The random-init CNN above emits checkerboard noise from its untrained 4x4/2
    #     up-convolutions (neighbouring disparities differ by ~20 target pixels), i.e. a fully scattered splat. -----------
    s_tex = torch.tensor(host['tex'], device=dev)
    s_disp = torch.tensor(host['disp'], device=dev)
    s_mask = torch.ones(L, B, H, W, 1, device=dev)
    s_mask._lsi_all_ones = True
    for _ in range(3):
        with torch.no_grad():
            ldi_utils.forward_splat((s_tex, s_mask, s_disp), pc, *cam, **kw)
    torch.cuda.synchronize()
    lib.lsi_b200_kernel_timing_enable(1)
    for _ in range(10):
        with torch.no_grad():
            ldi_utils.forward_splat((s_tex, s_mask, s_disp), pc, *cam, **kw)
    torch.cuda.synchronize()
    kms2, kn2 = (ctypes.c_double * 8)(), (ctypes.c_int * 8)()
    _b200.call('lsi_b200_kernel_timing_collect', ctypes.cast(kms2, ctypes.c_void_p), ctypes.cast(kn2, ctypes.c_void_p))
    lib.lsi_b200_kernel_timing_enable(0)
    s_ms = kms2[0] / 10
    s_ach = splat_bytes_per_step / (s_ms * 1e-3) / 1e9
    roofline['structured_ldi'] = {'what': 'same kernel, same shapes, SURVEY.md 8(d) config-4 synthetic LDI (planar tex [.,3] + disp [.,1] '
                                          'tensors: 16 B per pixel-layer) instead of the random-init CNN output',
                                  'achieved': s_ach, 'frac': s_ach / peak, 'kernel_ms_per_step': s_ms,
                                  'normalize_ms_per_step': kms2[1] / 10}
    del s_tex, s_disp, s_mask

    # --- end to end through the public API with HOST buffers: every step copies its images + cameras host->device and its
    #     rendered views device->host (train_utils.HostViewPipeline: the copies of neighbouring steps overlap the kernels) ----
    del imgs
    torch.cuda.empty_cache()
    hbatch = {'img': torch.tensor(img_host).pin_memory()}
    for name, key in (('k_s', 'k_s'), ('k_t', 'k_t'), ('rot', 'rot'), ('t', 't')):
        hbatch[name] = torch.tensor(host[key]).pin_memory()
    pipe = train_utils.HostViewPipeline(opts, store, kw, B, H, W, dev, depth=2)
    checks = []
    e2e_steps = max(5, min(args.steps, 10))
    pipe.run([hbatch] * 3)
    barrier()
    t0 = time.perf_counter()
    pipe.run([hbatch] * e2e_steps, on_result=lambda k, im, wt: checks.append(float(im[0, 0, 0, 0, 0])))
    barrier()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    assert len(checks) == e2e_steps and all(np.isfinite(c) for c in checks)
    if dist is not None:
        tt = torch.tensor([e2e_s], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_s = tt.item()
    e2e = {'value': world * B / e2e_s, 'unit': 'views/s', 'h2d_bytes_per_step': pipe.h2d_bytes(hbatch),
           'd2h_bytes_per_step': pipe.d2h_bytes(), 'ms_per_step': e2e_s * 1e3, 'steps': e2e_steps,
           'api': 'lsi.nnutils.train_utils.HostViewPipeline.run (predict_ldi + lsi.geometry.ldi.forward_splat per batch) on pinned '
                  'host images/cameras, rendered views copied back to pinned host memory every step; H2D of step k+1 and D2H of '
                  'step k-1 overlap the kernels of step k'}
    del pipe

    # --- training step at BASELINE config 4's per-GPU shard (batch 8 per GPU, 256x832, L=4): two towers, view-synthesis
    #     loss, backward, ONE all-reduce of the flat gradient buffer (NCCL, when world > 1), fused Adam -------------------
    torch.cuda.empty_cache()
    tb = 8
    infer_mode = nets.get_conv_mode()
    nets.set_conv_mode('tf32')       # the training step runs the TF32 tcgen05 kernels (fwd / dgrad / wgrad)
    topts = train_utils.default_opts(dataset='kitti', n_layers=L, batch_size=tb, img_height=H, img_width=W)
    trainer = train_utils.Trainer(topts, store=nets.ParamStore(device=dev, seed=0))
    rs2 = np.random.RandomState(200 + rank)
    tbatch = {'imgs_src': torch.tensor(img_host[:tb], device=dev),
              'imgs_trg': torch.tensor(np.ascontiguousarray(img_host[tb:2 * tb]), device=dev),
              'k_s': cam[0][:tb].contiguous(), 'k_t': cam[1][:tb].contiguous(), 'rot_mat': cam[2][:tb].contiguous(),
              'trans_mat': cam[3][:tb].contiguous()}
    for _ in range(2):
        trainer.train_step(tbatch)
    barrier()
    t_steps = max(2, min(args.steps, 5))
    te0, te1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    te0.record()
    for _ in range(t_steps):
        tloss, _ = trainer.train_step(tbatch)
    te1.record()
    barrier()
    tms = te0.elapsed_time(te1) / t_steps
    if dist is not None:
        tt = torch.tensor([tms], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        tms = tt.item()
    train = {'ms_per_step': tms, 'image_pairs_per_s': world * tb / (tms * 1e-3), 'batch_per_gpu': tb, 'steps': t_steps,
             'loss': float(tloss), 'grad_allreduce_bytes': int(trainer.store.flat_grad.numel() * 4),
             'what': 'ldi_enc_dec.py training step: 2 U-Net towers + %d heads, self-consistency + 4 forward splats + smoothness '
                     '+ ordering losses, backward (tcgen05 dgrad/wgrad), %s, fused Adam' % (L, 'NCCL all-reduce of the flat '
                     'gradient buffer' if world > 1 else 'no collective at 1 GPU')}

    nets.set_conv_mode(infer_mode)
    cpu = cpu_baseline() if (rank == 0 and world == 1) else None
    if rank == 0:
        print(json.dumps({
            'metric': METRIC, 'value': value, 'unit': 'views/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': {'split': 'f32-equivalent convs: split fp16 (hi, lo) pairs = 22-bit mantissas, 3 exact tcgen05 kind::f16 products per fp32 product, fp32 accumulate',
                      'f16': 'f16 conv operands/activations (fp32 accumulate, fp32 batch statistics)',
                      'tf32': 'tf32 convs (fp32 accumulate)', 'fp32': 'f32 CUDA-core convs'}[nets.get_conv_mode()] + ' + f32 renderer', 'data': 'synthetic',
            'config': {'workload': WORKLOAD, 'h': H, 'w': W, 'layers': L, 'batch_per_gpu': B, 'global_batch': world * B,
                       'parallelism': 'dp%d (independent views per rank, no data-path collective at inference)' % world,
                       'l2_policy': 'per-step activations (several GB) exceed the 126 MB L2'},
            'clocks': clocks, 'e2e': e2e, 'gpu_launches': int(launches), 'roofline': roofline, 'conv': conv,
            'renderer_only_views_per_s': renderer_only, 'train': train, 'cpu_baseline': cpu}))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
