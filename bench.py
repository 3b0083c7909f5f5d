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
SPLAT_TRAFFIC_FILE = os.path.join(ROOT, 'profiles', 'r2_splat_rowgather_traffic.json')   # dram bytes of the splat + normalise launches, from a
                                                                                # committed ncu capture (--cache-control none); see tools/ncu_summary.py
WORKLOAD = ('KITTI-like 256x832 image -> encoder-decoder U-Net + 4 LDI heads (W zero-padded to 896 for the U-Net, '
            'prediction cropped; tcgen05 convs in the fp32-parity split-precision mode, batch-stat BN) -> forward_splat(compose_layers=True, '
            'trg_downsampling=1) -> rendered target view; batch %d per GPU' % B_PER_GPU)
DTYPES = {'split': 'f32-equivalent convs (split fp16 (hi, lo) pairs = 22-bit mantissas, 3 exact tcgen05 kind::f16 products per fp32 '
                   'product, fp32 accumulation: parity-green against the fp64 oracle at the fp32 bars)',
          'f16': 'f16 conv operands/activations (fp32 accumulate, fp32 batch statistics)',
          'tf32': 'tf32 convs (fp32 accumulate)', 'fp32': 'f32 CUDA-core convs'}


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


def make_images(batch, seed):
    """Band-limited synthetic source images in [0,1] (8 distinct ones, tiled)."""
    from oracle import gen_inputs
    rs = np.random.RandomState(100 + seed)
    uniq = [gen_inputs.band_limited(rs, (H, W), 3) for _ in range(min(batch, 8))]
    return np.stack([uniq[i % len(uniq)] for i in range(batch)]).astype(np.float32)


def _oracle_view_fn(views, img=None):
    """The reference CPU path for `views` views: oracle CNN (lsi_oracle_nets) + oracle renderer (lsi_oracle), same
    padding policy as the B200 path.  Returns (fn, state): fn() runs one pass and leaves the LDI prediction in state['pred']."""
    from oracle import lsi_oracle as O
    from oracle import lsi_oracle_nets as N
    if img is None:
        img = make_images(views, 0)
    img = torch.tensor(img[:views])
    wp = -(-W // 128) * 128
    padded = torch.zeros(views, H, wp, 3)
    padded[:, :, :W] = img
    params = N.init_params(L, seed=0)
    s = make_inputs(views, 0)
    cam = [torch.tensor(s[k]) for k in ('k_s', 'k_t', 'rot', 't')]
    pc = O.pixel_coords(views, H, W)
    kw = dict(compose_layers=True, trg_downsampling=1, bg_layer_disp=BG_DISP, max_disp=MAX_DISP, zbuf_scale=ZBUF_SCALE)
    state = {}

    def fn():
        with torch.no_grad():
            tex, masks, disps = N.predict_ldi(params, padded, L, MAX_DISP)
            ldi = (tex[:, :, :, :W].contiguous(), masks[:, :, :, :W].contiguous(), disps[:, :, :, :W].contiguous())
            state['pred'] = torch.cat([ldi[0], ldi[2]], dim=-1)
            return O.forward_splat(ldi, pc, *cam, **kw)
    return fn, state


def run_reference(args, rank):
    """--impl reference: the reference's own CPU algorithm (TF-1.4 is not installable here, so this is the op-for-op
    CPU restatement under oracle/, `kind: port`) on the host cores, bounded sample of the same workload."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    views = 2                                    # bounded sample: 2 of the 64 views per step
    fn, _ = _oracle_view_fn(views)
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


def cpu_baseline(img2, gpu_preds):
    """The oracle timed on the host cores (bounded sample: 2 views), and -- the oracle acting as the checker -- the absolute
    error of each conv mode's LDI prediction for the same 2 images and the same weights against the oracle's fp32 output."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    views = 2
    fn, state = _oracle_view_fn(views, img2)
    fn()
    reps, t0 = 0, time.perf_counter()
    while reps < 2 or (time.perf_counter() - t0 < 12.0 and reps < 20):
        fn()
        reps += 1
    dt = (time.perf_counter() - t0) / reps
    ref = state['pred'].numpy().astype(np.float64)
    acc = {}
    for mode, pred in gpu_preds.items():
        d = np.abs(pred.astype(np.float64) - ref)
        acc[mode] = {'max_abs_err': float(d.max()), 'mean_abs_err': float(d.mean()), 'max_rel_err': float(d.max() / np.abs(ref).max())}
    return ({'value': views / dt, 'unit': 'views/s', 'cores': cores, 'kind': 'port',
             'sample': '%d of %d views x %d reps of the oracle (lsi_oracle_nets.predict_ldi + lsi_oracle.forward_splat in the '
                       'reference decomposition), torch CPU fp32' % (views, B_PER_GPU, reps)},
            {'what': 'LDI prediction (tex, disp) of each conv mode for the cpu_baseline sample (2 views, batch-norm over those 2) against the '
                     'CPU oracle fp32 evaluation with the same weights; sigmoid outputs in [0,1] / [0,%.1f]' % MAX_DISP, 'modes': acc})


def timed_steps(step, n, barrier):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(n):
        step()
    e1.record()
    barrier()
    return e0.elapsed_time(e1) / n


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--no-extras', action='store_true', help='skip the config 3 / config 5 / backward-kernel measurements')
    ap.add_argument('--profile-step', action='store_true', help='run the warm-up and the timed steps only, then exit (launch lists under ncu)')
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
    numa_node = train_utils.bind_to_gpu_numa_node(local_rank)       # before any pinned allocation

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        tt = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return tt.item()

    def collect():
        kms, kn = (ctypes.c_double * 8)(), (ctypes.c_int * 8)()
        _b200.call('lsi_b200_kernel_timing_collect', ctypes.cast(kms, ctypes.c_void_p), ctypes.cast(kn, ctypes.c_void_p))
        return list(kms), list(kn)

    B = B_PER_GPU
    host = make_inputs(B, seed=rank)
    img_host = make_images(B, rank)
    imgs = torch.tensor(img_host, device=dev)
    cam = [torch.tensor(host[k], device=dev) for k in ('k_s', 'k_t', 'rot', 't')]
    pc = helpers.pixel_coords(B, H, W, _device=dev)
    opts = train_utils.default_opts(dataset='kitti', n_layers=L, batch_size=B, img_height=H, img_width=W,
                                    zbuf_scale=ZBUF_SCALE)
    # headline conv mode: 'split' -- the tensor-core mode that passes the fp32 parity bars against the oracle
    # (tests/test_gpu_split.py).  The narrower modes ('f16', 'tf32') are timed below as extra keys, each with its error.
    head_mode = os.environ.get('BENCH_CONV_MODE', 'split')
    nets.set_conv_mode(head_mode)
    store = nets.ParamStore(device=dev, seed=0)           # random-init weights of the reference architecture
    kw = dict(compose_layers=True, trg_downsampling=DS, bg_layer_disp=BG_DISP, max_disp=MAX_DISP, zbuf_scale=ZBUF_SCALE)
    with torch.no_grad():
        train_utils.predict_ldi(imgs[:1], opts, store, reuse=False)      # creates the variables

    def step(x=None):
        with torch.no_grad():
            ldi = train_utils.predict_ldi(imgs if x is None else x, opts, store, reuse=True)
            return ldi_utils.forward_splat(tuple(ldi), pc, *cam, **kw)

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
    clocks = sampler.stop()
    ms_per_step = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    value = world * B / (ms_per_step * 1e-3)
    if args.profile_step:
        if rank == 0:
            print(json.dumps({'metric': METRIC, 'value': value, 'unit': 'views/s', 'ms_per_step': ms_per_step, 'steps': args.steps,
                              'gpu_launches': int(launches), 'config': {'conv_mode': head_mode}, 'note': 'profile-step run'}))
        return

    # --- per-kernel timing (CUDA events on the launching stream, live) -------------------------------------------
    lib.lsi_b200_kernel_timing_enable(1)
    for _ in range(args.steps):
        step()
    torch.cuda.synchronize()
    kms, kn = collect()
    lib.lsi_b200_kernel_timing_enable(0)
    peak, peak_src = load_peaks()
    n_src, n_trg = H * W, int(H * DS) * int(W * DS)
    # SURVEY.md 8(d) bytes_fwd (packed form: mask == 1, no trg_disp): the splat kernel reads 16 B per source pixel-layer, the normalise
    # pass writes 16 B per target pixel; the accumulator between them stays in L2 by design and is not algorithmic traffic
    bytes_step = float(4 * (4 * L * n_src + 4 * n_trg) * B)
    read_bytes_step = float(4 * 4 * L * n_src * B)
    splat_ms, norm_ms = kms[0] / args.steps, kms[1] / args.steps
    n_launch = max(kn[0] // args.steps, 1)
    achieved = bytes_step / ((splat_ms + norm_ms) * 1e-3) / 1e9
    traffic = None
    if os.path.exists(SPLAT_TRAFFIC_FILE):
        with open(SPLAT_TRAFFIC_FILE) as f:
            tj = json.load(f)
        traffic = tj['dram_bytes_per_view'] * B / n_launch
    fused = kn[1] == 0      # row-gather kernel: splat + normalise in one launch (rectified poses); else the reduction kernels + normalise pass
    roofline = {'bound': 'hbm', 'kernel': ('splat_fwd_rowgather_kernel (target rows owned by CTAs, scatter inverted in shared memory, bg / divide_safe / '
                                           'store fused; %d launch per step)' % n_launch) if fused else
                                          ('splat_fwd_stream_kernel + normalize_fast_kernel (forward splat into the L2-resident accumulator, then '
                                           'bg / divide_safe / store; %d launches of each per step)' % n_launch),
                'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak, 'peak_source': peak_src,
                'traffic': traffic, 'traffic_source': 'profiles/r2_splat_rowgather_traffic.json (ncu dram__bytes_read+write, --cache-control none)' if traffic else None,
                'algorithmic_bytes_per_launch': bytes_step / n_launch, 'algorithmic_bytes_per_step': bytes_step,
                'definition': 'SURVEY 8(d) bytes_fwd, packed (mask==1, no trg_disp): 4*(4*L*N + 4*N_t) per view, over splat + normalise time',
                'kernel_ms_per_step': splat_ms + norm_ms, 'splat_ms_per_step': splat_ms, 'normalize_ms_per_step': norm_ms,
                'splat_kernel_read_only': {'achieved': read_bytes_step / (splat_ms * 1e-3) / 1e9, 'frac': read_bytes_step / (splat_ms * 1e-3) / 1e9 / peak,
                                           'what': 'round-1 definition: 16*L*N read bytes over the splat kernel alone'},
                'input': 'LDI predicted by the random-init CNN (checkerboard-noise disparities: fully scattered splat)'}
    wp = -(-W // 128) * 128
    flops_padded = (21.8e9 + 23.0e9 * L) * (H * wp) / (256.0 * 768.0) * B      # forward 2*MAC per step as executed (W padded to 896)
    flops_useful = (21.8e9 + 23.0e9 * L) * (H * W) / (256.0 * 768.0) * B       # the same network on the 832 columns that are asked for
    conv_ms = (kms[4] + kms[5]) / args.steps
    mma_factor = {'split': 3.0}.get(head_mode, 1.0)
    tpeak = None
    pk = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(pk):
        with open(pk) as f:
            tpeak = float(json.load(f).get('bf16_tflops_sustained', 0.0)) or None
    conv = {'bound': 'tensor', 'kernel': 'conv_tc_kernel + conv_halo_kernel (tcgen05 kind::%s, TMA) + fp32 CUDA-core stem' % ('tf32' if head_mode == 'tf32' else 'f16'),
            'mode': head_mode, 'achieved': flops_useful / (conv_ms * 1e-3) / 1e12, 'unit': 'TFLOP/s (useful fp32-equivalent FLOPs of the 256x832 network)',
            'achieved_padded': flops_padded / (conv_ms * 1e-3) / 1e12,
            'executed_mma_tflops': mma_factor * flops_padded / (conv_ms * 1e-3) / 1e12,
            'executed_mma_frac_of_sustained_bf16': (mma_factor * flops_padded / (conv_ms * 1e-3) / 1e12 / tpeak) if tpeak else None,
            'peak_sustained_bf16_tflops': tpeak,
            'mma_per_useful_product': mma_factor, 'kernel_ms_per_step': conv_ms, 'tc_ms_per_step': kms[4] / args.steps,
            'fp32_ms_per_step': kms[5] / args.steps, 'other_nn_ms_per_step': kms[7] / args.steps,
            'flops_per_step_useful': flops_useful, 'flops_per_step_padded': flops_padded,
            'share_of_step': conv_ms / ms_per_step}

    # --- the narrower conv modes as extra keys (same step, same inputs) ------------------------------------------------
    modes = {head_mode: {'ms_per_step': ms_per_step, 'views_per_s': value, 'dtype': DTYPES[head_mode]}}
    gpu_preds = {}
    for mode in ('split', 'f16', 'tf32'):
        nets.set_conv_mode(mode)
        with torch.no_grad():        # LDI prediction for the cpu_baseline sample (2 views) in this mode, for the error table
            ldi2 = train_utils.predict_ldi(imgs[:2].contiguous(), opts, store, reuse=True)
            gpu_preds[mode] = torch.cat([ldi2[0], ldi2[2]], dim=-1).float().cpu().numpy()
        if mode != head_mode:
            for _ in range(3):
                step()
            ms_m = max_over_ranks(timed_steps(step, max(3, args.steps // 2), barrier))
            modes[mode] = {'ms_per_step': ms_m, 'views_per_s': world * B / (ms_m * 1e-3), 'dtype': DTYPES[mode]}
    nets.set_conv_mode(head_mode)

    # --- renderer slice alone (the kernel the roofline is about), LDIs resident in HBM ------------------------------
    with torch.no_grad():
        ldi_fixed = [t.contiguous() for t in train_utils.predict_ldi(imgs, opts, store, reuse=True)]
    ldi_fixed[1]._lsi_all_ones = True

    def render_only():
        with torch.no_grad():
            ldi_utils.forward_splat(tuple(ldi_fixed), pc, *cam, **kw)
    for _ in range(3):
        render_only()
    renderer_only = B / (timed_steps(render_only, 20, barrier) * 1e-3)
    del ldi_fixed

    # --- the same kernels on SURVEY.md 8(d)'s structured config-4 LDI (road-plane disparity ramps + smooth bumps, what a
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
    kms2, _ = collect()
    lib.lsi_b200_kernel_timing_enable(0)
    s_ms, s_nm = kms2[0] / 10, kms2[1] / 10
    s_ach = bytes_step / ((s_ms + s_nm) * 1e-3) / 1e9
    roofline['structured_ldi'] = {'what': 'same kernels, same shapes, SURVEY.md 8(d) config-4 synthetic LDI (planar tex [.,3] + disp [.,1] '
                                          'tensors: 16 B per pixel-layer) instead of the random-init CNN output',
                                  'achieved': s_ach, 'frac': s_ach / peak, 'kernel_ms_per_step': s_ms + s_nm, 'splat_ms_per_step': s_ms,
                                  'normalize_ms_per_step': s_nm,
                                  'splat_kernel_read_only_frac': read_bytes_step / (s_ms * 1e-3) / 1e9 / peak}
    del s_tex, s_disp, s_mask

    # --- end to end through the public API with HOST buffers: every step copies its images + cameras host->device and its
    #     rendered views device->host (train_utils.HostViewPipeline: the copies of neighbouring steps overlap the kernels) ----
    del imgs
    torch.cuda.empty_cache()
    # the source images are 8-bit data (PNG / JPEG decoders produce uint8): they cross the bus as uint8 and are scaled on the device
    hbatch = {'img': torch.tensor(np.round(img_host * 255.0).astype(np.uint8)).pin_memory()}
    for name, key in (('k_s', 'k_s'), ('k_t', 'k_t'), ('rot', 'rot'), ('t', 't')):
        hbatch[name] = torch.tensor(host[key]).pin_memory()
    pipe = train_utils.HostViewPipeline(opts, store, kw, B, H, W, dev, depth=3, u8_input=True)
    checks = []
    e2e_steps = max(5, min(args.steps, 10))
    pipe.run([hbatch] * 3)
    barrier()
    t0 = time.perf_counter()
    pipe.run([hbatch] * e2e_steps, on_result=lambda k, im, wt: checks.append(float(im[0, 0, 0, 0, 0])))
    barrier()
    e2e_s = max_over_ranks((time.perf_counter() - t0) / e2e_steps)
    assert len(checks) == e2e_steps and all(np.isfinite(c) for c in checks)
    e2e = {'value': world * B / e2e_s, 'unit': 'views/s', 'h2d_bytes_per_step': pipe.h2d_bytes(hbatch),
           'd2h_bytes_per_step': pipe.d2h_bytes(), 'ms_per_step': e2e_s * 1e3, 'steps': e2e_steps, 'conv_mode': head_mode,
           'numa_node': numa_node,
           'api': 'lsi.nnutils.train_utils.HostViewPipeline.run (predict_ldi + lsi.geometry.ldi.forward_splat per batch) on pinned '
                  'host uint8 images + float cameras, rendered views copied back to pinned host memory every step; H2D of step k+1 '
                  'and D2H of step k-1 overlap the kernels of step k; 3 buffers deep; process bound to the GPU\'s NUMA node'}
    del pipe

    # --- strong scaling of BASELINE config 4 as written: global batch 64 split over the ranks (64 / world views per GPU) ----
    strong = None
    if B_PER_GPU % world == 0:
        bs = B_PER_GPU // world
        imgs_s = torch.tensor(img_host[:bs], device=dev)
        cam_s = [c[:bs].contiguous() for c in cam]
        pc_s = helpers.pixel_coords(bs, H, W, _device=dev)

        def step_s():
            with torch.no_grad():
                ldi = train_utils.predict_ldi(imgs_s, opts, store, reuse=True)
                return ldi_utils.forward_splat(tuple(ldi), pc_s, *cam_s, **kw)
        for _ in range(3):
            step_s()
        ms_s = max_over_ranks(timed_steps(step_s, max(3, args.steps // 2), barrier))
        strong = {'global_batch': B_PER_GPU, 'batch_per_gpu': bs, 'ms_per_step': ms_s, 'views_per_s': B_PER_GPU / (ms_s * 1e-3),
                  'what': 'inference, config 4 as written (global batch 64): views/s of the whole job'}
        del imgs_s, cam_s, pc_s

    # --- extras at N=1: config 3 variants, config 5 sweep, backward kernels ----------------------------------------------
    extras = None
    if world == 1 and not args.no_extras:
        import bench_extras
        torch.cuda.empty_cache()
        extras = bench_extras.run(dev, peak)
        torch.cuda.empty_cache()

    # --- training step at BASELINE config 4's per-GPU shard (batch 8 per GPU, 256x832, L=4): two towers, view-synthesis
    #     loss, backward, ONE all-reduce of the flat gradient buffer (NCCL, when world > 1), fused Adam -------------------
    torch.cuda.empty_cache()
    tb = 8
    nets.set_conv_mode('tf32')       # the training step runs the TF32 tcgen05 kernels (fwd / dgrad / wgrad)
    topts = train_utils.default_opts(dataset='kitti', n_layers=L, batch_size=tb, img_height=H, img_width=W)
    trainer = train_utils.Trainer(topts, store=nets.ParamStore(device=dev, seed=0))
    tbatch = {'imgs_src': torch.tensor(img_host[:tb], device=dev),
              'imgs_trg': torch.tensor(np.ascontiguousarray(img_host[tb:2 * tb]), device=dev),
              'k_s': cam[0][:tb].contiguous(), 'k_t': cam[1][:tb].contiguous(), 'rot_mat': cam[2][:tb].contiguous(),
              'trans_mat': cam[3][:tb].contiguous()}
    for _ in range(2):
        trainer.train_step(tbatch)
    t_steps = max(2, min(args.steps, 5))
    tloss = [None]

    def tstep():
        tloss[0], _ = trainer.train_step(tbatch)
    tms = max_over_ranks(timed_steps(tstep, t_steps, barrier))
    train = {'ms_per_step': tms, 'image_pairs_per_s': world * tb / (tms * 1e-3), 'batch_per_gpu': tb, 'global_batch': world * tb,
             'steps': t_steps, 'dtype': 'tf32 convs (fp32 accumulate)', 'scaling': 'weak (config 4 at 8 GPUs = global batch 64)',
             'loss': float(tloss[0]), 'grad_allreduce_bytes': int(trainer.store.flat_grad.numel() * 4),
             'what': 'ldi_enc_dec.py training step: 2 U-Net towers + %d heads, self-consistency + 4 forward splats + smoothness '
                     '+ ordering losses, backward (tcgen05 dgrad/wgrad), %s, fused Adam' % (L, 'NCCL all-reduce of the flat '
                     'gradient buffer' if world > 1 else 'no collective at 1 GPU')}
    # data-parallel correctness, seen by the driver: after the timed steps every rank must hold bit-identical parameters, and the
    # all-reduced gradient must equal the sum of the shard gradients (checked through a fixed random projection)
    chk = trainer.train_step(tbatch, dp_check=True)[2]
    if dist is not None:
        flat = trainer.store.flat
        ref = flat.clone()
        dist.broadcast(ref, 0)
        dmax = max_over_ranks(float((flat - ref).abs().max()))
        train['dp_check'] = {'max_abs_param_diff_vs_rank0': dmax, 'grad_projection_sum_of_shards': chk['proj_sum_of_shards'],
                             'grad_projection_allreduced': chk['proj_allreduced'],
                             'grad_projection_rel_diff': abs(chk['proj_sum_of_shards'] - chk['proj_allreduced']) / max(abs(chk['proj_sum_of_shards']), 1e-30),
                             'world': world}
        assert dmax == 0.0, 'ranks hold different parameters after the step: max |d| = %g' % dmax
        assert train['dp_check']['grad_projection_rel_diff'] < 1e-4, train['dp_check']
        # the same step with batch-norm statistics over the global batch (the reference's single-device semantics)
        tr_sync = train_utils.Trainer(topts, store=nets.ParamStore(device=dev, seed=0), sync_bn=True)
        for _ in range(2):
            tr_sync.train_step(tbatch)
        train['sync_bn_ms_per_step'] = max_over_ranks(timed_steps(lambda: tr_sync.train_step(tbatch), t_steps, barrier))
        del tr_sync
        nets.set_sync_bn(False)
    del trainer
    nets.set_conv_mode(head_mode)

    cpu, accuracy = (cpu_baseline(img_host[:2], gpu_preds) if (rank == 0 and world == 1) else (None, None))
    if rank == 0:
        print(json.dumps({
            'metric': METRIC, 'value': value, 'unit': 'views/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': DTYPES[head_mode] + ' + f32 renderer', 'data': 'synthetic',
            'config': {'workload': WORKLOAD, 'h': H, 'w': W, 'layers': L, 'batch_per_gpu': B, 'global_batch': world * B,
                       'conv_mode': head_mode,
                       'parallelism': 'dp%d (independent views per rank, no data-path collective at inference)' % world,
                       'l2_policy': 'per-step activations (several GB) exceed the 126 MB L2'},
            'clocks': clocks, 'e2e': e2e, 'gpu_launches': int(launches), 'roofline': roofline, 'conv': conv,
            'conv_modes': modes, 'accuracy_vs_oracle': accuracy, 'renderer_only_views_per_s': renderer_only,
            'strong_scaling_config4': strong, 'train': train, 'extras': extras, 'cpu_baseline': cpu}))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
