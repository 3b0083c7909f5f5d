"""Golden fixtures for the evaluation metrics: the reference's own, unmodified
`ldi_pred_eval.Tester.define_metrics` (ldi_pred_eval.py:297-548) and `projection.disocclusion_mask`
(projection.py:109-150) run over the eager TF-1 shim (oracle/tf1_shim) on a stand-in `self`.

Separate from gen_golden.py because ldi_pred_eval and ldi_enc_dec define the same absl flags (one process can import only one
of them).  `lsi.nnutils.test_utils` (matplotlib / scipy.misc / html plumbing) is replaced by an empty stand-in: only the
metric arithmetic is run.  Test infrastructure: run here (the reference tree exists only in the build container);
    python oracle/gen_golden_eval.py   ->  tests/golden/eval_synth.npz, tests/golden/eval_kitti.npz
"""
import builtins
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = '/root/reference'
GOLD = os.path.join(ROOT, 'tests', 'golden')


def _import_reference():
    sys.path.insert(0, os.path.join(HERE, 'tf1_shim'))
    sys.path.insert(1, REF)
    for name in ('lsi.data.kitti.data', 'lsi.data.syntheticPlanes.data'):
        sys.modules[name] = types.ModuleType(name)       # Python-2 syntax + absent datasets: not on the path
    tu = types.ModuleType('lsi.nnutils.test_utils')      # plotting / html plumbing of the reference: not on the path
    tu.define_default_flags = lambda flags: None
    tu.Tester = object
    sys.modules['lsi.nnutils.test_utils'] = tu
    import tensorflow as tf
    from lsi.nnutils import helpers
    helpers.range = lambda *a: list(builtins.range(*a))  # Python-2 range() semantics for helpers.transpose
    from lsi.geometry import projection
    import lsi.nnutils as _n
    _n.test_utils = tu
    import lsi.data.kitti as _k
    import lsi.data.syntheticPlanes as _s
    _k.data = sys.modules['lsi.data.kitti.data']
    _s.data = sys.modules['lsi.data.syntheticPlanes.data']
    import ldi_pred_eval
    return tf, helpers, projection, ldi_pred_eval


tf, helpers, projection, ldi_pred_eval = _import_reference()
T = tf.Tensor
sys.path.insert(0, ROOT)
from oracle import gen_inputs  # noqa: E402


def make_case(seed, L, B, H, W, cam, max_disp, bg):
    s = gen_inputs.scene(L, B, H, W, cam, seed, max_disp)
    rs = np.random.RandomState(seed + 500)
    case = dict(tex_s=s['tex'], disp_s=s['disp'], k_s=s['k_s'], k_t=s['k_t'], rot=s['rot'], t=s['t'],
                tex_t=rs.uniform(0, 1, s['tex'].shape), disp_t=rs.uniform(0.05, 0.95, s['disp'].shape) * max_disp,
                img_s=rs.uniform(0, 1, (B, H, W, 3)), img_t=rs.uniform(0, 1, (B, H, W, 3)))
    # ground-truth disparities: smooth fields with a background region (<= bg) and a few zeros (kitti: holes in the GT)
    yy, xx = np.meshgrid(np.linspace(0, 1, H), np.linspace(0, 1, W), indexing='ij')
    for nme, ph in (('gt_disp_s', 0.0), ('gt_disp_t', 0.7)):
        g = max_disp * (0.25 + 0.5 * np.sin(3.0 * xx + ph) ** 2 * yy)
        g = np.where(rs.uniform(0, 1, (B, H, W)) < 0.15, 0.0 if cam == 'kitti' else 0.5 * bg, g[None])
        case[nme] = g[..., None]
    case['gt_disp_bg_s'] = case['gt_disp_s'] * rs.uniform(0.3, 1.1, (B, H, W, 1))
    case['gt_disp_bg_t'] = case['gt_disp_t'] * rs.uniform(0.3, 1.1, (B, H, W, 1))
    case['gt_tex_bg_s'] = rs.uniform(0, 1, (B, H, W, 3))
    case['gt_tex_bg_t'] = rs.uniform(0, 1, (B, H, W, 3))
    return {k: np.asarray(v, dtype=np.float32).astype(np.float64) for k, v in case.items()}


def run_metrics(case, opts_kw, dtype):
    tf._set_float(dtype)
    L, B, H, W, _ = case['tex_s'].shape
    c = lambda k: T(torch.tensor(case[k], dtype=dtype))
    me = types.SimpleNamespace()
    me.opts = types.SimpleNamespace(**opts_kw)
    ones = torch.ones(L, B, H, W, 1, dtype=dtype)
    me.ldi_src = [c('tex_s'), T(ones), c('disp_s')]
    me.ldi_trg = [c('tex_t'), T(ones.clone()), c('disp_t')]
    me.imgs_src, me.imgs_trg = c('img_s'), c('img_t')
    me.k_s, me.k_t, me.rot_mat, me.trans_mat = c('k_s'), c('k_t'), c('rot'), c('t')
    me.pixel_coords = helpers.pixel_coords(B, H, W)
    me.focal_disps = None
    me.inv_rot_mat = helpers.transpose(me.rot_mat)                 # ldi_pred_eval.py:180-181
    me.inv_trans_mat = -tf.matmul(me.inv_rot_mat, me.trans_mat)
    me.src_gt_disp, me.trg_gt_disp = c('gt_disp_s'), c('gt_disp_t')
    me.visuals = {}
    out = {}
    if opts_kw['dataset'] == 'synthetic':                          # ldi_pred_eval.py:129-161
        me.src_gt_disp_bg, me.trg_gt_disp_bg = c('gt_disp_bg_s'), c('gt_disp_bg_t')
        me.src_gt_tex_bg, me.trg_gt_tex_bg = c('gt_tex_bg_s'), c('gt_tex_bg_t')
        src2trg = projection.forward_projection_matrix(me.k_s, me.k_t, me.rot_mat, me.trans_mat)
        trg2src = projection.inverse_projection_matrix(me.k_s, me.k_t, me.rot_mat, me.trans_mat)
        me.disocclusion_mask_src = projection.disocclusion_mask(me.src_gt_disp, me.trg_gt_disp, me.pixel_coords, src2trg)
        me.disocclusion_mask_trg = projection.disocclusion_mask(me.trg_gt_disp, me.src_gt_disp, me.pixel_coords, trg2src)
        out['disocc_src'] = me.disocclusion_mask_src.t.to(torch.float64).numpy()
        out['disocc_trg'] = me.disocclusion_mask_trg.t.to(torch.float64).numpy()
    else:                                                          # ldi_pred_eval.py:163-173
        me.disocclusion_mask_src = tf.equal(me.src_gt_disp, 0)
        me.disocclusion_mask_trg = tf.equal(me.trg_gt_disp, 0)
    ldi_pred_eval.Tester.define_metrics(me)                        # ldi_pred_eval.py:297-548, unmodified
    for k, v in me.metrics.items():
        out['m_' + k] = np.float64(v.t.item() if hasattr(v, 't') else v)
    for k, v in me.metrics_norm.items():
        out['n_' + k] = np.float64(v.t.item() if hasattr(v, 't') else v)
    return out


def main():
    cfgs = [
        ('eval_synth', 41, 2, 2, 16, 24, 'synth',
         dict(dataset='synthetic', kitti_dl_disparities=False, trg_splat_downsampling=0.5, zbuf_scale=50, bg_layer_disp=0.2,
              max_disp=1.0, splat_bdry_ignore=0.1, n_layers=2)),
        ('eval_kitti', 42, 3, 1, 8, 24, 'kitti',
         dict(dataset='kitti', kitti_dl_disparities=True, trg_splat_downsampling=1, zbuf_scale=50, bg_layer_disp=1e-3,
              max_disp=0.4, splat_bdry_ignore=0.05, n_layers=3)),
    ]
    for name, seed, L, B, H, W, cam, opts_kw in cfgs:
        opts_kw = dict(opts_kw, batch_size=B, img_height=H, img_width=W)
        case = make_case(seed, L, B, H, W, cam, opts_kw['max_disp'], opts_kw['bg_layer_disp'])
        blob = {'in_' + k: v.astype(np.float32) for k, v in case.items()}
        blob.update({'opt_' + k: (np.float64(v) if not isinstance(v, str) else np.asarray(v)) for k, v in opts_kw.items()})
        for dtype, sfx in ((torch.float32, '_f32'), (torch.float64, '_f64')):
            for k, v in run_metrics(case, opts_kw, dtype).items():
                blob[k + sfx] = v
        path = os.path.join(GOLD, name + '.npz')
        np.savez_compressed(path, **blob)
        print('%-14s %7.1f KB  %s' % (name, os.path.getsize(path) / 1024.0,
                                     {k: float(v) for k, v in blob.items() if k.startswith(('m_', 'n_')) and k.endswith('_f64')}))


if __name__ == '__main__':
    main()
