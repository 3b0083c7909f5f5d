"""TEST INFRASTRUCTURE -- generates tests/golden/*.npz by executing the reference's OWN, UNMODIFIED
Python sources (/root/reference/lsi/..., /root/reference/ldi_enc_dec.py) over the torch-backed TF-1 shim
in oracle/tf1_shim.  Run in the build container only (the GPU box has no /root/reference):

    python oracle/gen_golden.py            # rewrites tests/golden/*.npz

What is "reference" in a fixture: the WIRING (which ops, in which order, with which constants) is the
reference's; the per-op TF-1.4 semantics are the shim's restatement (see oracle/tf1_shim/tensorflow).
Gradients are torch autograd through that same wiring.  Two Python-2-isms are bridged without touching
the sources: `range()` returning a list (helpers.py:75-77 assigns into it) is injected into the module
namespace of lsi.nnutils.helpers, and the dataset modules (Python-2 syntax, external data) are stubbed
before `import ldi_enc_dec`.

Every case is produced twice: with the shim computing in float32 (what TF would do; key suffix `_f32`)
and in float64 (`_f64`, rounding-free reference used to bound the fp32 noise of both sides).
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
    import tensorflow as tf
    from lsi.nnutils import helpers
    helpers.range = lambda *a: list(builtins.range(*a))  # Python-2 range() semantics for helpers.transpose
    from lsi.geometry import ldi, projection, sampling
    from lsi.loss import loss
    import lsi.data.kitti as _k
    import lsi.data.syntheticPlanes as _s
    _k.data = sys.modules['lsi.data.kitti.data']
    _s.data = sys.modules['lsi.data.syntheticPlanes.data']
    import ldi_enc_dec
    return tf, helpers, ldi, projection, sampling, loss, ldi_enc_dec


tf, helpers, ldi, projection, sampling, loss, ldi_enc_dec = _import_reference()
T = tf.Tensor


# ---------------------------------------------------------------------------------------------------
# seeded procedural inputs (fp64 masters; cast per precision)
# ---------------------------------------------------------------------------------------------------
def rot_xyz(ax, ay, az):
    cx, sx, cy, sy, cz, sz = np.cos(ax), np.sin(ax), np.cos(ay), np.sin(ay), np.cos(az), np.sin(az)
    rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    return rz @ ry @ rx


def synth_k(h, w):
    return np.array([[w, 0, w / 2.0], [0, h, h / 2.0], [0, 0, 1.0]])


def kitti_k(h, w):
    return np.array([[721.54 * w / 1242.0, 0, 609.56 * w / 1242.0],
                     [0, 721.54 * h / 375.0, 172.85 * h / 375.0], [0, 0, 1.0]])


def make_case(name, seed, L, B, H, W, cam, ds, bg, max_disp, scale, masks, focal=False):
    rs = np.random.RandomState(seed)
    tex = rs.uniform(0, 1, (L, B, H, W, 3))
    disp = rs.uniform(0.05, 0.95, (L, B, H, W, 1)) * max_disp
    mask = rs.uniform(0.1, 1.0, (L, B, H, W, 1)) if masks else np.ones((L, B, H, W, 1))
    if cam == 'synth':
        k_s = np.stack([synth_k(H, W)] * B)
        k_t = np.stack([synth_k(H, W) * np.array([[1.05], [0.97], [1.0]])] * B)
        rot = np.stack([rot_xyz(*rs.uniform(-0.06, 0.06, 3)) for _ in range(B)])
        t = rs.uniform(-0.12, 0.12, (B, 3, 1))
    elif cam == 'kitti':
        k_s = np.stack([kitti_k(H, W)] * B)
        k_t = k_s.copy()
        rot = np.stack([np.eye(3)] * B)
        t = np.tile(np.array([[-0.5327], [0.0], [0.0]]), (B, 1, 1))
    else:
        raise ValueError(cam)
    h_t, w_t = int(H * ds), int(W * ds)
    case = dict(tex=tex, disp=disp, mask=mask, k_s=k_s, k_t=k_t, rot=rot, t=t,
                ds=np.float64(ds), bg=np.float64(bg), max_disp=np.float64(max_disp), scale=np.float64(scale))
    if focal:
        case['focal'] = rs.uniform(0.0, 0.1, (B, 1, 1, 1)) * max_disp
    # upstream gradients for the three outputs, for both compose modes
    for comp, nl in (('c', 1), ('i', L)):
        case['g_img_' + comp] = rs.normal(0, 1, (nl, B, h_t, w_t, 3))
        case['g_wts_' + comp] = rs.normal(0, 1, (nl, B, h_t, w_t, 1)) * 1e-3
        case['g_disp_' + comp] = rs.normal(0, 1, (nl, B, h_t, w_t, 1))
    return name, case


def run_forward_splat(case, dtype):
    """ldi.forward_splat through the reference sources; returns outputs + autograd gradients."""
    tf._set_float(dtype)
    out = {}
    L, B, H, W, _ = case['tex'].shape
    cam = [T(torch.tensor(case[k], dtype=dtype)) for k in ('k_s', 'k_t', 'rot', 't')]
    focal = T(torch.tensor(case['focal'], dtype=dtype)) if 'focal' in case else None
    pc = helpers.pixel_coords(B, H, W)
    ds = float(case['ds'])
    ds = int(ds) if ds == int(ds) else ds
    for comp, flag in (('c', True), ('i', False)):
        leaves = [torch.tensor(case[k], dtype=dtype, requires_grad=True) for k in ('tex', 'mask', 'disp')]
        img, wts, dsp = ldi.forward_splat(
            tuple(T(x) for x in leaves), pc, *cam, focal_disps=focal, compose_layers=flag,
            compute_trg_disp=True, trg_downsampling=ds, bg_layer_disp=float(case['bg']),
            max_disp=float(case['max_disp']), zbuf_scale=float(case['scale']))
        out['img_' + comp], out['wts_' + comp], out['disp_' + comp] = (x.t.detach().numpy() for x in (img, wts, dsp))
        # gradient of <img, g_img> only (the training path consumes img only, ldi_enc_dec.py:307,322) ...
        s = (img.t * torch.tensor(case['g_img_' + comp], dtype=dtype)).sum()
        gr = torch.autograd.grad(s, leaves, retain_graph=True)
        for nme, g in zip(('tex', 'mask', 'disp'), gr):
            out['d%s_img_%s' % (nme, comp)] = g.numpy()
        # ... and of all three outputs together
        s = s + (wts.t * torch.tensor(case['g_wts_' + comp], dtype=dtype)).sum() \
              + (dsp.t * torch.tensor(case['g_disp_' + comp], dtype=dtype)).sum()
        gr = torch.autograd.grad(s, leaves)
        for nme, g in zip(('tex', 'mask', 'disp'), gr):
            out['d%s_all_%s' % (nme, comp)] = g.numpy()
    return out


def save_case(name, case, out32, out64):
    blob = {}
    for k, v in case.items():
        blob['in_' + k] = np.asarray(v, dtype=np.float32)
    for k, v in out32.items():
        blob[k + '_f32'] = v.astype(np.float32)
    for k, v in out64.items():
        blob[k + '_f64'] = v.astype(np.float64)
    path = os.path.join(GOLD, name + '.npz')
    np.savez_compressed(path, **blob)
    print('%-28s %7.1f KB' % (name, os.path.getsize(path) / 1024.0))


def gen_forward_splat():
    cases = [
        make_case('fs_synth_ds05', 11, 2, 2, 12, 16, 'synth', 0.5, 0.2, 1.0, 50, masks=True),
        make_case('fs_synth_ds1', 12, 3, 1, 10, 14, 'synth', 1, 0.2, 1.0, 10, masks=True),
        make_case('fs_kitti_ds1', 13, 2, 1, 8, 26, 'kitti', 1, 1e-3, 0.4, 50, masks=False),
        make_case('fs_kitti_ds05', 14, 4, 2, 8, 26, 'kitti', 0.5, 1e-3, 0.4, 50, masks=False),
        make_case('fs_focal', 15, 2, 2, 8, 12, 'synth', 1, 0.2, 1.0, 50, masks=True, focal=True),
    ]
    for name, case in cases:
        # inputs are rounded to fp32 FIRST so that the f64 run sees exactly the values the f32 run sees
        case = {k: np.asarray(v, dtype=np.float32).astype(np.float64) for k, v in case.items()}
        save_case(name, case, run_forward_splat(case, torch.float32), run_forward_splat(case, torch.float64))


# ---------------------------------------------------------------------------------------------------
# primitives: splat, bilinear, disocclusion_mask, projection matrices, zbuffer_weights, losses
# ---------------------------------------------------------------------------------------------------
def gen_primitives():
    rs = np.random.RandomState(21)
    B, H, W, C, HT, WT = 2, 6, 7, 3, 5, 9
    src = rs.uniform(0, 1, (B, H, W, C)).astype(np.float32)
    coords = np.stack([rs.uniform(-1.5, WT + 1.5, (B, H, W)), rs.uniform(-1.5, HT + 1.5, (B, H, W))], -1).astype(np.float32)
    init = rs.uniform(0, 1, (B, HT, WT, C)).astype(np.float32)
    g = rs.normal(0, 1, (B, HT, WT, C)).astype(np.float32)
    img = rs.uniform(0, 1, (B, HT, WT, C)).astype(np.float32)
    gb = rs.normal(0, 1, (B, H, W, C)).astype(np.float32)
    blob = dict(in_src=src, in_coords=coords, in_init=init, in_g=g, in_img=img, in_gb=gb)
    for dtype, sfx in ((torch.float32, '_f32'), (torch.float64, '_f64')):
        tf._set_float(dtype)
        s = torch.tensor(src, dtype=dtype, requires_grad=True)
        c = torch.tensor(coords, dtype=dtype, requires_grad=True)
        i0 = torch.tensor(init, dtype=dtype, requires_grad=True)
        o = sampling.splat(T(s), T(c), T(i0))
        gs, gc, gi = torch.autograd.grad((o.t * torch.tensor(g, dtype=dtype)).sum(), [s, c, i0])
        blob.update({'splat' + sfx: o.t.detach().numpy(), 'splat_dsrc' + sfx: gs.numpy(),
                     'splat_dcoords' + sfx: gc.numpy(), 'splat_dinit' + sfx: gi.numpy()})
        im = torch.tensor(img, dtype=dtype, requires_grad=True)
        c2 = torch.tensor(coords, dtype=dtype, requires_grad=True)
        o = sampling.bilinear(T(im), T(c2))
        gi2, gc2 = torch.autograd.grad((o.t * torch.tensor(gb, dtype=dtype)).sum(), [im, c2])
        blob.update({'bilinear' + sfx: o.t.detach().numpy(), 'bilinear_dimg' + sfx: gi2.numpy(),
                     'bilinear_dcoords' + sfx: gc2.numpy()})
        # bilinear_wrapper on a 5-D input ([L,B,...]) -- sampling.py:135-168
        im5 = torch.tensor(np.stack([img, img[::-1]]), dtype=dtype)
        c5 = torch.tensor(np.stack([coords, coords * 0.9]), dtype=dtype)
        blob['bilinear_wrapper' + sfx] = sampling.bilinear_wrapper(T(im5), T(c5)).t.numpy()
        # compose=False (sampling.py:117-131): four masked corner samples + their raw weights, plain and through the wrapper
        ims_nc, wts_nc = sampling.bilinear(T(torch.tensor(img, dtype=dtype)), T(torch.tensor(coords, dtype=dtype)), compose=False)
        blob['bilinear_nc_ims' + sfx] = np.stack([v.t.numpy() for v in ims_nc])
        blob['bilinear_nc_wts' + sfx] = np.stack([v.t.numpy() for v in wts_nc])
        ims_w, wts_w = sampling.bilinear_wrapper(T(im5.clone()), T(c5.clone()), compose=False)
        blob['bilinear_wrapper_nc_ims' + sfx] = np.stack([v.t.numpy() for v in ims_w])
        blob['bilinear_wrapper_nc_wts' + sfx] = np.stack([v.t.numpy() for v in wts_w])
        # projection + disocclusion
        k_s = torch.tensor(np.stack([synth_k(H, W)] * B), dtype=dtype)
        k_t = torch.tensor(np.stack([synth_k(H, W) * np.array([[1.1], [0.9], [1.0]])] * B), dtype=dtype)
        rot = torch.tensor(np.stack([rot_xyz(0.03, -0.05, 0.02), rot_xyz(-0.02, 0.04, 0.01)]), dtype=dtype)
        t = torch.tensor(np.array([[[0.1], [-0.05], [0.02]], [[-0.07], [0.03], [0.04]]]), dtype=dtype)
        fwd = projection.forward_projection_matrix(T(k_s), T(k_t), T(rot), T(t))
        inv = projection.inverse_projection_matrix(T(k_s), T(k_t), T(rot), T(t))
        blob.update({'in_k_s': k_s.numpy().astype(np.float32), 'in_k_t': k_t.numpy().astype(np.float32),
                     'in_rot': rot.numpy().astype(np.float32), 'in_t': t.numpy().astype(np.float32),
                     'proj_fwd' + sfx: fwd.t.numpy(), 'proj_inv' + sfx: inv.t.numpy()})
        d_src = torch.tensor(rs.uniform(0.1, 0.9, (B, H, W, 1)).astype(np.float32), dtype=dtype)
        d_trg = torch.tensor(rs.uniform(0.1, 0.9, (B, H, W, 1)).astype(np.float32), dtype=dtype)
        if sfx == '_f32':
            blob['in_d_src'], blob['in_d_trg'] = d_src.numpy(), d_trg.numpy()
        else:
            d_src = torch.tensor(blob['in_d_src'], dtype=dtype)
            d_trg = torch.tensor(blob['in_d_trg'], dtype=dtype)
        pc = helpers.pixel_coords(B, H, W)
        blob['disocc' + sfx] = projection.disocclusion_mask(T(d_src), T(d_trg), pc, fwd, thresh=0.05).t.numpy()
        # zbuffer weights incl. the d<=0 and d>1 branches, and the scalar form used for bg_wt (ldi.py:115)
        zin = np.array([-0.3, 0.0, 1e-6, 0.2, 0.5, 0.999, 1.0, 1.7], dtype=np.float32)
        blob['in_zbw'] = zin
        blob['zbw50' + sfx] = helpers.zbuffer_weights(T(torch.tensor(zin, dtype=dtype)), scale=50).t.numpy()
        blob['zbw_bg_synth' + sfx] = helpers.zbuffer_weights(0.2 / 1.0, scale=50).t.numpy()
        blob['zbw_bg_kitti' + sfx] = helpers.zbuffer_weights(1e-3 / 0.4, scale=50).t.numpy()
        # soft_z_buffering / enforce_bg_occupied (helpers.py:140-177)
        lm = torch.tensor(rs.uniform(0, 1, (3, B, H, W, 1)).astype(np.float32), dtype=dtype)
        ld = torch.tensor(rs.uniform(-0.1, 1, (3, B, H, W, 1)).astype(np.float32), dtype=dtype)
        if sfx == '_f32':
            blob['in_lm'], blob['in_ld'] = lm.numpy(), ld.numpy()
        else:
            lm = torch.tensor(blob['in_lm'], dtype=dtype)
            ld = torch.tensor(blob['in_ld'], dtype=dtype)
        blob['softz' + sfx] = helpers.soft_z_buffering(T(lm), T(ld), depth_softmax_temp=0.4).t.numpy()
        blob['bg_occ' + sfx] = helpers.enforce_bg_occupied(T(lm)).t.numpy()
    path = os.path.join(GOLD, 'primitives.npz')
    np.savez_compressed(path, **blob)
    print('%-28s %7.1f KB' % ('primitives', os.path.getsize(path) / 1024.0))


# ---------------------------------------------------------------------------------------------------
# the whole view-synthesis loss: ldi_enc_dec.Trainer.define_loss_graph run on a stand-in `self`
# ---------------------------------------------------------------------------------------------------
def run_loss(case, opts_kw, dtype):
    tf._set_float(dtype)
    L, B, H, W, _ = case['tex_s'].shape
    names = ('tex_s', 'mask_s', 'disp_s', 'tex_t', 'mask_t', 'disp_t')
    leaves = [torch.tensor(case[k], dtype=dtype, requires_grad=True) for k in names]
    me = types.SimpleNamespace()
    me.opts = types.SimpleNamespace(**opts_kw)
    me.ldi_src = [T(x) for x in leaves[:3]]
    me.ldi_trg = [T(x) for x in leaves[3:]]
    me.imgs_src = T(torch.tensor(case['img_s'], dtype=dtype))
    me.imgs_trg = T(torch.tensor(case['img_t'], dtype=dtype))
    me.k_s, me.k_t = T(torch.tensor(case['k_s'], dtype=dtype)), T(torch.tensor(case['k_t'], dtype=dtype))
    me.rot_mat, me.trans_mat = T(torch.tensor(case['rot'], dtype=dtype)), T(torch.tensor(case['t'], dtype=dtype))
    me.pixel_coords = helpers.pixel_coords(B, H, W)
    me.focal_disps = None
    me.inv_rot_mat = helpers.transpose(me.rot_mat)                 # ldi_enc_dec.py:193-194
    me.inv_trans_mat = -tf.matmul(me.inv_rot_mat, me.trans_mat)
    ldi_enc_dec.Trainer.define_loss_graph(me)                     # ldi_enc_dec.py:265-410, unmodified
    total = me.total_loss.t
    out = dict(total=total.detach().numpy(), self_cons=me.self_cons_loss.t.detach().numpy(),
               indep_splat=me.indep_splat_loss.t.detach().numpy(),
               compose_splat=me.compose_splat_loss.t.detach().numpy(),
               smooth=me.disp_smoothness_loss.t.detach().numpy(),
               incr=me.incr_depth_loss.t.detach().numpy())
    for nme, g in zip(names, torch.autograd.grad(total, leaves)):
        out['d' + nme] = g.numpy()
    return out


def gen_loss():
    cfgs = [
        ('loss_synth', 31, 2, 2, 16, 16, 'synth',
         dict(self_cons_wt=1.0, indep_splat_wt=1.0, compose_splat_wt=1.0, splat_bdry_ignore=0.1, zbuf_scale=50,
              trg_splat_downsampling=0.5, disp_smoothness_wt=0.1, incr_depth_wt=10.0, bg_layer_disp=0.2,
              max_disp=1.0, l0_self_cons=False), True),
        ('loss_kitti', 32, 3, 1, 8, 24, 'kitti',
         dict(self_cons_wt=10.0, indep_splat_wt=1.0, compose_splat_wt=1.0, splat_bdry_ignore=0.05, zbuf_scale=50,
              trg_splat_downsampling=0.5, disp_smoothness_wt=0.1, incr_depth_wt=10.0, bg_layer_disp=1e-3,
              max_disp=0.4, l0_self_cons=False), False),
    ]
    for name, seed, L, B, H, W, cam, opts_kw, masks in cfgs:
        _, c = make_case(name, seed, L, B, H, W, cam, 1, opts_kw['bg_layer_disp'], opts_kw['max_disp'], 50, masks)
        rs = np.random.RandomState(seed + 100)
        case = dict(tex_s=c['tex'], mask_s=c['mask'], disp_s=c['disp'],
                    tex_t=rs.uniform(0, 1, c['tex'].shape),
                    mask_t=(rs.uniform(0.1, 1, c['mask'].shape) if masks else np.ones(c['mask'].shape)),
                    disp_t=rs.uniform(0.05, 0.95, c['disp'].shape) * opts_kw['max_disp'],
                    img_s=rs.uniform(0, 1, (B, H, W, 3)), img_t=rs.uniform(0, 1, (B, H, W, 3)),
                    k_s=c['k_s'], k_t=c['k_t'], rot=c['rot'], t=c['t'])
        case = {k: np.asarray(v, dtype=np.float32).astype(np.float64) for k, v in case.items()}
        blob = {'in_' + k: v.astype(np.float32) for k, v in case.items()}
        blob.update({'opt_' + k: np.float64(v) for k, v in opts_kw.items()})
        for dtype, sfx in ((torch.float32, '_f32'), (torch.float64, '_f64')):
            for k, v in run_loss(case, opts_kw, dtype).items():
                blob[k + sfx] = v
        # the zbuffer_composition_loss term on its own (loss.py:66-115), with its gradients
        for dtype, sfx in ((torch.float32, '_f32'), (torch.float64, '_f64')):
            tf._set_float(dtype)
            leaves = [torch.tensor(case[k], dtype=dtype, requires_grad=True) for k in ('tex_s', 'mask_s', 'disp_s')]
            v = loss.zbuffer_composition_loss(T(leaves[0]), T(leaves[1]), T(leaves[2]),
                                              T(torch.tensor(case['img_s'], dtype=dtype)),
                                              bg_layer_disp=opts_kw['bg_layer_disp'], max_disp=opts_kw['max_disp'],
                                              zbuf_scale=opts_kw['zbuf_scale'])
            blob['zcl' + sfx] = v.t.detach().numpy()
            for nme, g in zip(('tex', 'mask', 'disp'), torch.autograd.grad(v.t, leaves)):
                blob['zcl_d' + nme + sfx] = g.numpy()
        path = os.path.join(GOLD, name + '.npz')
        np.savez_compressed(path, **blob)
        print('%-28s %7.1f KB' % (name, os.path.getsize(path) / 1024.0))


# ---------------------------------------------------------------------------------------------------
# the CNN: the reference's own lsi/nnutils/nets.py wiring over the slim stand-in
# ---------------------------------------------------------------------------------------------------
def gen_nets():
    from lsi.nnutils import nets
    sys.path.insert(0, ROOT)
    from oracle import lsi_oracle_nets as N
    L, B, H, W, steps, max_disp = 2, 2, 128, 128, 3, 1.0
    rs = np.random.RandomState(41)
    img = rs.uniform(0, 1, (B, H, W, 3)).astype(np.float32)
    g_pred = np.random.RandomState(42).normal(0, 1, (L, B, H, W, 4)).astype(np.float32)   # regenerated by the tests
    blob = dict(in_img=img, meta=np.array([L, B, H, W, steps], dtype=np.int64), max_disp=np.float64(max_disp),
                param_seed=np.int64(7), g_seed=np.int64(42))
    for dtype, sfx in ((torch.float32, '_f32'), (torch.float64, '_f64')):
        tf._set_float(dtype)
        params = N.init_params(L, seed=7, n_layerwise_steps=steps, random_beta=True, dtype=dtype)
        leaves = {k: v.clone().requires_grad_(True) for k, v in params.items()}
        tf._VARS.clear()
        tf._VARS.update(leaves)
        x = torch.tensor(img, dtype=dtype, requires_grad=True)
        _, feat_dec, skip_feat, _ = nets.encoder_decoder_unet(T(x), nl_diff_enc_dec=steps)       # ldi_enc_dec.py:196-199
        ldi_pred = nets.ldi_predictor(feat_dec, n_layers=L, n_layerwise_steps=steps, skip_feat=skip_feat)
        tex, masks, disps = (v.t for v in ldi_pred)
        disps = disps * max_disp                                                                  # ldi_enc_dec.py:213
        pred = torch.cat([tex, disps], dim=-1)
        assert float(masks.min()) == 1.0 and float(masks.max()) == 1.0
        # fixtures stay small: strided samples of the big tensors plus fp64 checksums of the whole tensors
        blob['pred' + sfx] = pred.detach()[:, :, ::4, ::4, :].numpy().astype(np.float32)
        blob['feat_dec' + sfx] = feat_dec.t.detach()[:, ::2, ::2, ::4].numpy().astype(np.float32)
        blob['pred_stats' + sfx] = np.array([float(pred.double().sum()), float(pred.double().pow(2).sum().sqrt())])
        s = (pred * torch.tensor(g_pred, dtype=dtype)).sum()
        names = sorted(leaves)
        grads = torch.autograd.grad(s, [leaves[n] for n in names] + [x])
        blob['d_img' + sfx] = grads[-1][:, ::4, ::4].numpy().astype(np.float32)
        blob['d_img_stats' + sfx] = np.array([float(grads[-1].double().sum()), float(grads[-1].double().pow(2).sum().sqrt())])
        keep = ['encoder_decoder_unet/cnv1/weights', 'encoder_decoder_unet/cnv1/BatchNorm/beta',
                'encoder_decoder_unet/cnv7b/BatchNorm/beta', 'encoder_decoder_unet/icnv4/BatchNorm/beta',
                'ldi_tex_disp/pixelwise_pred/upsample_1/pred_1/weights', 'ldi_tex_disp/pixelwise_pred/upsample_1/pred_1/biases',
                'ldi_tex_disp/pixelwise_pred/upsample_0/decoder/upcnv1/weights',
                'ldi_tex_disp/pixelwise_pred/upsample_0/decoder/upcnv3b/BatchNorm/beta']
        for n, g in zip(names, grads[:-1]):
            if n in keep:
                blob['grad:' + n + sfx] = g.numpy().astype(np.float32)
        blob['grad_names'] = np.array(names)
        blob['grad_sum' + sfx] = np.array([float(g.double().sum()) for g in grads[:-1]])
        blob['grad_l2' + sfx] = np.array([float(g.double().pow(2).sum().sqrt()) for g in grads[:-1]])
        extra = [k for k in tf._VARS if k not in leaves]
        assert all(('/fc/' in k) or any(t in k for t in ('icnv3', 'icnv2', 'icnv1', 'unet/upcnv3', 'unet/upcnv2', 'unet/upcnv1'))
                   for k in extra), extra          # only the never-executed parts may be missing from the oracle's list
    path = os.path.join(GOLD, 'nets_unet_l2.npz')
    np.savez_compressed(path, **blob)
    print('%-28s %7.1f KB' % ('nets_unet_l2', os.path.getsize(path) / 1024.0))


def gen_nets_simple():
    """The non-U-Net variant: the reference's own encoder_decoder_simple (nets.py:211-241) + ldi_predictor without skips
    (ldi_enc_dec.py:202-213), over the slim stand-in."""
    from lsi.nnutils import nets
    sys.path.insert(0, ROOT)
    from oracle import lsi_oracle_nets as N
    L, B, H, W, steps, nz, max_disp = 1, 4, 128, 128, 3, 96, 1.0
    img = np.random.RandomState(43).uniform(0, 1, (B, H, W, 3)).astype(np.float32)
    blob = dict(in_img=img, meta=np.array([L, B, H, W, steps, nz], dtype=np.int64), max_disp=np.float64(max_disp), param_seed=np.int64(9))
    for dtype, sfx in ((torch.float32, '_f32'), (torch.float64, '_f64')):
        tf._set_float(dtype)
        params = N.init_params_simple(L, (H, W), seed=9, random_beta=True, dtype=dtype, nz=nz)
        tf._VARS.clear()
        tf._VARS.update({k: v.clone() for k, v in params.items()})
        x = torch.tensor(img, dtype=dtype)
        feat, feat_dec, skip_feat, _ = nets.encoder_decoder_simple(T(x), nz=nz, nl_diff_enc_dec=steps)     # ldi_enc_dec.py:202-205
        assert skip_feat is None
        tex, masks, disps = (v.t for v in nets.ldi_predictor(feat_dec, n_layers=L, n_layerwise_steps=steps, skip_feat=skip_feat))
        pred = torch.cat([tex, disps * max_disp], dim=-1)
        assert sorted(tf._VARS) == sorted(params), sorted(set(tf._VARS) ^ set(params))   # the oracle's variable list IS the reference's
        blob['feat' + sfx] = feat.t.numpy().astype(np.float32)
        blob['feat_dec' + sfx] = feat_dec.t[:, ::2, ::2, :].numpy().astype(np.float32)
        blob['pred' + sfx] = pred[:, :, ::8, ::8, :].numpy().astype(np.float32)
        blob['pred_stats' + sfx] = np.array([float(pred.double().sum()), float(pred.double().pow(2).sum().sqrt())])
    blob['var_names'] = np.array(sorted(params))
    path = os.path.join(GOLD, 'nets_simple_l1.npz')
    np.savez_compressed(path, **blob)
    print('%-28s %7.1f KB' % ('nets_simple_l1', os.path.getsize(path) / 1024.0))


if __name__ == '__main__':
    os.makedirs(GOLD, exist_ok=True)
    which = sys.argv[1:] or ['fs', 'prim', 'loss', 'nets']
    if 'fs' in which:
        gen_forward_splat()
    if 'prim' in which:
        gen_primitives()
    if 'loss' in which:
        gen_loss()
    if 'nets' in which:
        gen_nets()
    if 'nets_simple' in which:
        gen_nets_simple()
