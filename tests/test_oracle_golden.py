"""CPU: the oracle (oracle/lsi_oracle.py) against the fixtures produced by the reference's own sources
(oracle/gen_golden.py) and against the closed-form known answers of SURVEY.md section 8(c)."""
import math

import numpy as np
import pytest
import torch

from oracle import lsi_oracle as O
from _util import load_golden, rel_err, t32

FS_CASES = ['fs_synth_ds05', 'fs_synth_ds1', 'fs_kitti_ds1', 'fs_kitti_ds05', 'fs_focal']


def _fs_inputs(g, dtype=torch.float32):
    c = lambda k: torch.tensor(g['in_' + k], dtype=dtype)
    L, B, H, W, _ = g['in_tex'].shape
    kw = dict(trg_downsampling=float(g['in_ds']), bg_layer_disp=float(g['in_bg']),
              max_disp=float(g['in_max_disp']), zbuf_scale=float(g['in_scale']))
    if 'in_focal' in g:
        kw['focal_disps'] = c('focal')
    return (c('tex'), c('mask'), c('disp')), O.pixel_coords(B, H, W, dtype), (c('k_s'), c('k_t'), c('rot'), c('t')), kw


@pytest.mark.parametrize('name', FS_CASES)
@pytest.mark.parametrize('comp', ['c', 'i'])
def test_forward_splat_matches_reference_sources(name, comp):
    g = load_golden(name)
    ldi, pc, cam, kw = _fs_inputs(g)
    leaves = [x.clone().requires_grad_(True) for x in ldi]
    img, wts, dsp = O.forward_splat(tuple(leaves), pc, *cam, compose_layers=(comp == 'c'), compute_trg_disp=True, **kw)
    # same op sequence on the same library => expect (near) bit equality with the fp32 fixture
    assert rel_err(img.detach(), g['img_%s_f32' % comp]) < 1e-6
    assert rel_err(wts.detach(), g['wts_%s_f32' % comp]) < 1e-6
    assert rel_err(dsp.detach(), g['disp_%s_f32' % comp]) < 1e-6
    s = (img * t32(g['in_g_img_' + comp])).sum()
    for nme, gr in zip(('tex', 'mask', 'disp'), torch.autograd.grad(s, leaves)):
        assert rel_err(gr, g['d%s_img_%s_f32' % (nme, comp)]) < 1e-5, nme


@pytest.mark.parametrize('name', FS_CASES)
def test_fp32_fixture_close_to_fp64_fixture(name):
    """Bounds the fp32 noise of the reference arithmetic itself.  Forward values and d/dtex sit well inside the
    1e-4 parity bar; d/ddisp and d/dmask do NOT (up to 9.4e-5 on fs_kitti_ds1: the chain rule cancels
    <tex, gA_img> against gA_w), which is why the GPU parity tests bound the CUDA result against the fp64
    fixture and allow the fp32 fixture its own measured noise."""
    g = load_golden(name)
    for k in ('img_c', 'wts_c', 'disp_c', 'img_i', 'dtex_img_c'):
        assert rel_err(g[k + '_f32'], g[k + '_f64']) < 2e-5, k
    for k in ('ddisp_img_c', 'dmask_img_c', 'ddisp_img_i', 'dmask_img_i'):
        assert rel_err(g[k + '_f32'], g[k + '_f64']) < 5e-4, k


def test_primitives_match_reference_sources():
    g = load_golden('primitives')
    assert rel_err(O.splat(t32(g['in_src']), t32(g['in_coords']), t32(g['in_init'])), g['splat_f32']) < 1e-6
    assert rel_err(O.bilinear(t32(g['in_img']), t32(g['in_coords'])), g['bilinear_f32']) < 1e-6
    ims, wts = O.bilinear(t32(g['in_img']), t32(g['in_coords']), compose=False)
    assert rel_err(torch.stack(ims), g['bilinear_nc_ims_f32']) < 1e-6 and rel_err(torch.stack(wts), g['bilinear_nc_wts_f32']) < 1e-6
    cam = [t32(g['in_' + k]) for k in ('k_s', 'k_t', 'rot', 't')]
    fwd = O.forward_projection_matrix(*cam)
    assert rel_err(fwd, g['proj_fwd_f32']) < 1e-6
    assert rel_err(O.inverse_projection_matrix(*cam), g['proj_inv_f32']) < 1e-6
    B, H, W, _ = g['in_d_src'].shape
    dm = O.disocclusion_mask(t32(g['in_d_src']), t32(g['in_d_trg']), O.pixel_coords(B, H, W), fwd, thresh=0.05)
    assert np.array_equal(dm.numpy(), g['disocc_f32'])
    assert rel_err(O.zbuffer_weights(t32(g['in_zbw']), 50), g['zbw50_f32']) < 1e-6
    assert rel_err(O.soft_z_buffering(t32(g['in_lm']), t32(g['in_ld']), 0.4), g['softz_f32']) < 1e-6
    assert np.array_equal(O.enforce_bg_occupied(t32(g['in_lm'])).numpy(), g['bg_occ_f32'])


@pytest.mark.parametrize('name', ['loss_synth', 'loss_kitti'])
def test_view_synthesis_loss_matches_reference_sources(name):
    g = load_golden(name)
    names = ('tex_s', 'mask_s', 'disp_s', 'tex_t', 'mask_t', 'disp_t')
    leaves = [t32(g['in_' + k]).requires_grad_(True) for k in names]
    opts = O.LossOpts(**{k[4:]: float(v) for k, v in g.items() if k.startswith('opt_')})
    total, parts = O.view_synthesis_loss(tuple(leaves[:3]), tuple(leaves[3:]), t32(g['in_img_s']), t32(g['in_img_t']),
                                         t32(g['in_k_s']), t32(g['in_k_t']), t32(g['in_rot']), t32(g['in_t']), opts)
    assert abs(total.item() - float(g['total_f32'])) < 1e-6 * abs(float(g['total_f32']))
    for k, gk in (('self_cons', 'self_cons'), ('indep_splat', 'indep_splat'), ('compose_splat', 'compose_splat'),
                  ('disp_smoothness', 'smooth'), ('incr_depth', 'incr')):
        assert abs(float(parts[k]) - float(g[gk + '_f32'])) <= 1e-6 * max(abs(float(g[gk + '_f32'])), 1e-6), k
    for nme, gr in zip(names, torch.autograd.grad(total, leaves)):
        assert rel_err(gr, g['d%s_f32' % nme]) < 1e-5, nme
    zcl = O.zbuffer_composition_loss(*leaves[:3], t32(g['in_img_s']), bg_layer_disp=opts.bg_layer_disp,
                                     max_disp=opts.max_disp, zbuf_scale=opts.zbuf_scale)
    assert abs(zcl.item() - float(g['zcl_f32'])) < 1e-6 * abs(float(g['zcl_f32']))


# ---------------------------------------------------------------------------------------------------
# closed-form known answers (KAT 1-7)
# ---------------------------------------------------------------------------------------------------
def _kat_setup(h=8, w=8, tx=0.0, L=1, seed=0):
    torch.manual_seed(seed)
    tex = torch.rand(L, 1, h, w, 3)
    k = torch.tensor([[[float(w), 0, w / 2.0], [0, float(h), h / 2.0], [0, 0, 1.0]]])
    rot = torch.eye(3)[None]
    t = torch.tensor([[[tx], [0.0], [0.0]]])
    return tex, torch.ones(L, 1, h, w, 1), k, rot, t


def test_kat1_identity_pose():
    tex, mask, k, rot, t = _kat_setup()
    disp = torch.full((1, 1, 8, 8, 1), 0.5)
    img, wts, dsp = O.forward_splat((tex, mask, disp), O.pixel_coords(1, 8, 8), k, k, rot, t, compute_trg_disp=True,
                                    bg_layer_disp=0.2, max_disp=1, zbuf_scale=50)
    assert (img - tex).abs().max() < 1e-6
    assert (wts - 1.0).abs().max() < 1e-6          # e^0 + bg_wt (3.06e-7)
    assert (dsp - 0.5).abs().max() < 1e-6


def test_kat2_integer_shift_and_kat7_mass():
    tex, mask, k, rot, t = _kat_setup(tx=0.5)
    disp = torch.full((1, 1, 8, 8, 1), 0.5)        # shift = fx * tx * d = 8*0.5*0.5 = 2 px
    img, wts = O.forward_splat((tex, mask, disp), O.pixel_coords(1, 8, 8), k, k, rot, t, bg_layer_disp=0.2,
                               max_disp=1, zbuf_scale=50)
    assert (img[0, :, :, 2:] - tex[0, :, :, :-2]).abs().max() < 1e-6
    assert torch.equal(img[0, :, :, :2], torch.ones(1, 8, 2, 3))     # white canvas: bg_wt/bg_wt
    img, wts = O.forward_splat((tex, mask, disp), O.pixel_coords(1, 8, 8), k, k, rot, t, bg_layer_disp=0,
                               max_disp=1, zbuf_scale=50)
    assert wts.sum().item() == 48.0                 # KAT 7: (W-2)*H in-bounds weights of exactly 1


def test_kat3_half_pixel_shift():
    tex, mask, k, rot, t = _kat_setup(tx=0.125)    # 0.5 px
    disp = torch.full((1, 1, 8, 8, 1), 0.5)
    img, wts = O.forward_splat((tex, mask, disp), O.pixel_coords(1, 8, 8), k, k, rot, t, bg_layer_disp=0,
                               max_disp=1, zbuf_scale=50)
    assert (img[0, :, :, 1:] - 0.5 * (tex[0, :, :, 1:] + tex[0, :, :, :-1])).abs().max() < 1e-6
    assert (img[0, :, :, 0] - tex[0, :, :, 0]).abs().max() < 1e-6
    assert (wts[0, :, :, 0] - 0.5).abs().max() < 1e-6


def test_kat4_two_coincident_layers():
    tex, mask, k, rot, t = _kat_setup(L=2)
    tex = torch.stack([torch.zeros(1, 8, 8, 3), torch.ones(1, 8, 8, 3)])
    disp = torch.stack([torch.full((1, 8, 8, 1), 0.6), torch.full((1, 8, 8, 1), 0.5)])
    img, wts, dsp = O.forward_splat((tex, mask, disp), O.pixel_coords(1, 8, 8), k, k, rot, t, compute_trg_disp=True,
                                    bg_layer_disp=0, max_disp=1, zbuf_scale=10)
    assert (img - 1.0 / (1.0 + math.e)).abs().max() < 1e-6
    assert (dsp - 0.6).abs().max() < 1e-6


def test_kat5_scalars():
    assert O.zbuffer_weights(torch.tensor(0.0)).item() == 0.0
    assert abs(O.zbuffer_weights(torch.tensor(1.0), 50).item() / math.exp(25) - 1) < 1e-6
    assert abs(O.zbuffer_weights(0.2 / 1.0, 50).item() / 3.059e-7 - 1) < 1e-3
    assert abs(O.zbuffer_weights(1e-3 / 0.4, 50).item() / 1.574e-11 - 1) < 1e-3
    assert O.divide_safe(torch.tensor(1.0), torch.tensor(0.0)).item() == pytest.approx(1e8)
    assert O.pixel_coords(1, 2, 2)[0, 0, 0].tolist() == [0.5, 0.5, 1.0]
    g = load_golden('primitives')
    cam = [t32(g['in_' + k]) for k in ('k_s', 'k_t', 'rot', 't')]
    fwd, inv = O.forward_projection_matrix(*cam), O.inverse_projection_matrix(*cam)
    assert (fwd @ inv - torch.eye(4)).abs().max() < 2e-5
    assert torch.equal(fwd[:, 3], torch.tensor([[0.0, 0, 0, 1]] * 2))


def test_kat6_subthreshold_shift_does_not_blend():
    tex, mask, k, rot, t0 = _kat_setup(tx=0.0)
    _, _, _, _, t1 = _kat_setup(tx=1.25e-4)         # 5e-4 px: minor corner weight <= 1e-3 is zeroed
    disp = torch.full((1, 1, 8, 8, 1), 0.5)
    a, _ = O.forward_splat((tex, mask, disp), O.pixel_coords(1, 8, 8), k, k, rot, t0, bg_layer_disp=0, zbuf_scale=50)
    b, _ = O.forward_splat((tex, mask, disp), O.pixel_coords(1, 8, 8), k, k, rot, t1, bg_layer_disp=0, zbuf_scale=50)
    # the major weight (0.9995) cancels in img/wts up to 1 ulp; an un-thresholded splat would blend neighbours by ~5e-4
    assert (a - b).abs().max() < 2e-7


def test_kat8_gradients_vs_fp64_central_differences():
    g = load_golden('fs_synth_ds05')
    ldi, pc, cam, kw = _fs_inputs(g, torch.float64)
    gi = torch.tensor(g['in_g_img_c'], dtype=torch.float64)

    def f(tex, mask, disp):
        img, _ = O.forward_splat((tex, mask, disp), pc, *cam, compose_layers=True, **kw)
        return (img * gi).sum()

    leaves = [x.clone().requires_grad_(True) for x in ldi]
    grads = torch.autograd.grad(f(*leaves), leaves)
    rs = np.random.RandomState(0)
    eps = 1e-6
    for which in range(3):
        for _ in range(6):
            idx = tuple(rs.randint(0, s) for s in ldi[which].shape)
            plus = [x.clone() for x in ldi]
            minus = [x.clone() for x in ldi]
            plus[which][idx] += eps
            minus[which][idx] -= eps
            fd = (f(*plus) - f(*minus)).item() / (2 * eps)
            assert abs(fd - grads[which][idx].item()) < 1e-5 * max(1.0, abs(fd)), (which, idx)
