"""Measurements that ride along with bench.py at N=1 (extra keys of its JSON line):

  * config 3 (256x256, L=3, B=32, rotated + translated poses, random masks): the forward-splat variants on ITS shapes --
    0 shipped (streaming / fast kernels), 1 one-thread-per-pixel global atomics, 2 shared-memory row-owner tiles (the north
    star's SMEM-tile design), 3 block-per-segment reduction kernel;
  * config 5 (512x1664, L=5): the contention sweep t_x x {0, 0.25, 1, 4}, ds in {1, 0.5, 0.25} at the largest batch that is
    run here (16 views per launch set), HBM GB/s by SURVEY 8(d)'s bytes_fwd;
  * the backward kernels (splat_bwd_target + splat_bwd_source) at config 4 against bytes_bwd, at ds = 1 and ds = 0.5 (the
    training setting, ldi_enc_dec.py:101);
every number from CUDA events on the launching stream around the kernels themselves (lsi_b200_kernel_timing_*)."""
import ctypes

import numpy as np
import torch


def _collect(lib, _b200):
    kms, kn = (ctypes.c_double * 8)(), (ctypes.c_int * 8)()
    _b200.call('lsi_b200_kernel_timing_collect', ctypes.cast(kms, ctypes.c_void_p), ctypes.cast(kn, ctypes.c_void_p))
    return list(kms), list(kn)


def _scene_dev(gen_inputs, L, B, H, W, cam, seed, max_disp, dev, uniq=4, tx_scale=1.0):
    u = min(B, uniq)
    s = gen_inputs.scene(L, u, H, W, cam, seed, max_disp)
    reps = (B + u - 1) // u
    out = {}
    for k, v in s.items():
        axis = 1 if k in ('tex', 'mask', 'disp') else 0
        a = np.concatenate([v] * reps, axis=axis)
        out[k] = torch.tensor(np.ascontiguousarray(a[:, :B] if axis == 1 else a[:B]), device=dev)
    out['t'] = out['t'] * tx_scale
    return out


def bytes_fwd(L, n_src, n_trg, has_mask, nl_out=1, trg_disp=False):
    """reads tex(3)+disp(1)(+mask(1)) per source pixel per layer, writes img(3)+wts(1)(+disp(1)) per target pixel."""
    return 4 * ((5 if has_mask else 4) * L * n_src + (5 if trg_disp else 4) * nl_out * n_trg)


def bytes_bwd(L, n_src, n_trg, has_mask, nl_out=1):
    """read d_img (3 per target pixel), re-read the inputs, write d_tex / d_disp (/ d_mask)."""
    c = 5 if has_mask else 4
    return 4 * (3 * nl_out * n_trg + c * L * n_src + c * L * n_src)


def _time_fwd(lib, _b200, ldi_utils, ldi, pc, cam, kw, variant, reps=10):
    def step():
        with torch.no_grad():
            ldi_utils.forward_splat(ldi, pc, *cam, _variant=variant, **kw)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    lib.lsi_b200_kernel_timing_enable(1)
    for _ in range(reps):
        step()
    torch.cuda.synchronize()
    kms, _ = _collect(lib, _b200)
    lib.lsi_b200_kernel_timing_enable(0)
    return kms[0] / reps, kms[1] / reps


def run(dev, peak):
    from lsi import _b200
    from lsi.geometry import ldi as ldi_utils
    from lsi.nnutils import helpers
    from oracle import gen_inputs
    lib = _b200.lib()
    out = {}

    # ---- config 3: variant ablation on its own shapes ------------------------------------------------------------------
    L, B, H, W = 3, 32, 256, 256
    s = _scene_dev(gen_inputs, L, B, H, W, 'synth', 3, 1.0, dev, uniq=8)
    pc = helpers.pixel_coords(B, H, W, _device=dev)
    cam = [s[k] for k in ('k_s', 'k_t', 'rot', 't')]
    kw = dict(compose_layers=True, trg_downsampling=1, bg_layer_disp=0.2, max_disp=1.0, zbuf_scale=50.0)
    by = bytes_fwd(L, H * W, H * W, True) * B
    rows = {}
    names = {0: 'shipped (streaming / fast)', 1: 'global atomics, one thread per pixel-layer', 2: 'shared-memory row-owner tiles',
             3: 'block-per-segment reductions'}
    for v in (0, 1, 2, 3):
        try:
            sm, nm = _time_fwd(lib, _b200, ldi_utils, (s['tex'], s['mask'], s['disp']), pc, cam, kw, v)
            rows[str(v)] = {'kernel': names[v], 'splat_ms': sm, 'normalize_ms': nm, 'gbs': by / ((sm + nm) * 1e-3) / 1e9,
                            'frac': by / ((sm + nm) * 1e-3) / 1e9 / peak}
        except RuntimeError as e:
            rows[str(v)] = {'kernel': names[v], 'error': str(e)[:120]}
    out['config3_splat_variants'] = {'shape': '256x256 L3 B32, rotated poses, random masks, ds=1', 'bytes_fwd_per_batch': by, 'variants': rows}
    del s, pc, cam

    # ---- config 5: contention sweep -------------------------------------------------------------------------------------
    L, B, H, W = 5, 16, 512, 1664
    base = _scene_dev(gen_inputs, L, B, H, W, 'kitti', 5, 0.4, dev, uniq=2)
    pc = helpers.pixel_coords(B, H, W, _device=dev)
    mask1 = torch.ones(L, B, H, W, 1, device=dev)
    mask1._lsi_all_ones = True
    sweep = []
    for ds in (1.0, 0.5, 0.25):
        for txs in (0.0, 0.25, 1.0, 4.0):
            cam = [base['k_s'], base['k_t'], base['rot'], base['t'] * txs]
            kw = dict(compose_layers=True, trg_downsampling=ds, bg_layer_disp=1e-3, max_disp=0.4, zbuf_scale=50.0)
            sm, nm = _time_fwd(lib, _b200, ldi_utils, (base['tex'], mask1, base['disp']), pc, cam, kw, 0, reps=5)
            by = bytes_fwd(L, H * W, int(H * ds) * int(W * ds), False) * B
            sweep.append({'ds': ds, 'tx_scale': txs, 'splat_ms': sm, 'normalize_ms': nm, 'gbs': by / ((sm + nm) * 1e-3) / 1e9,
                          'frac': by / ((sm + nm) * 1e-3) / 1e9 / peak})
    out['config5_contention_sweep'] = {'shape': '512x1664 L5, 16 views per measurement (config 5 is 16/GPU), mask == 1', 'rows': sweep}
    del base, pc, mask1

    # ---- backward kernels at config 4 -----------------------------------------------------------------------------------
    L, B, H, W = 4, 16, 256, 832
    s = _scene_dev(gen_inputs, L, B, H, W, 'kitti', 4, 0.4, dev, uniq=4)
    pc = helpers.pixel_coords(B, H, W, _device=dev)
    cam = [s[k] for k in ('k_s', 'k_t', 'rot', 't')]
    bw = []
    for ds in (1.0, 0.5):
        kw = dict(compose_layers=True, trg_downsampling=ds, bg_layer_disp=1e-3, max_disp=0.4, zbuf_scale=50.0)
        leaves = [s[k].clone().requires_grad_(True) for k in ('tex', 'mask', 'disp')]

        def step():
            img, _ = ldi_utils.forward_splat(tuple(leaves), pc, *cam, **kw)
            g = torch.autograd.grad(img.sum(), leaves)
            return g
        for _ in range(2):
            step()
        torch.cuda.synchronize()
        lib.lsi_b200_kernel_timing_enable(1)
        reps = 5
        for _ in range(reps):
            step()
        torch.cuda.synchronize()
        kms, _ = _collect(lib, _b200)
        lib.lsi_b200_kernel_timing_enable(0)
        nt = int(H * ds) * int(W * ds)
        bb = bytes_bwd(L, H * W, nt, True) * B
        bf = bytes_fwd(L, H * W, nt, True) * B
        t_b = (kms[2] + kms[3]) / reps
        t_f = (kms[0] + kms[1]) / reps
        bw.append({'ds': ds, 'bwd_target_ms': kms[2] / reps, 'bwd_source_ms': kms[3] / reps, 'bytes_bwd': bb,
                   'bwd_gbs': bb / (t_b * 1e-3) / 1e9, 'bwd_frac': bb / (t_b * 1e-3) / 1e9 / peak,
                   'fwd_ms': t_f, 'fwd_gbs': bf / (t_f * 1e-3) / 1e9, 'fwd_frac': bf / (t_f * 1e-3) / 1e9 / peak})
    out['config4_backward'] = {'shape': '256x832 L4 B16, masks with gradient, compose', 'rows': bw}
    return out
