"""The GPU-native synthetic planar-room generator (lsi.data.syntheticPlanes, lsi.geometry.homography / layers).
CPU: world-layout helpers and view sampling against the fixture produced by the reference's own sources.  GPU: the fused
renderer and the per-plane API against the oracle (oracle/lsi_oracle_planes.py) and the same fixture; the loader's contract."""
import types

import numpy as np
import pytest
import torch

from _util import load_golden, rel_err


def test_layout_helpers_match_reference_sources():
    from lsi.data.syntheticPlanes import utils as U
    g = load_golden('planes_render')
    box = U.box_planes([-0.7, -0.5, 2.0, 0.7, 0.5, 3.5])
    assert np.allclose(np.stack([b['pt'] for b in box]), g['u_box_pt']) and np.allclose(np.stack([b['x_dir'] for b in box]), g['u_box_x'])
    assert np.allclose(np.stack([b['y_dir'] for b in box]), g['u_box_y']) and np.allclose(np.array([[b['w'], b['h']] for b in box]), g['u_box_wh'])
    assert np.allclose(U.dims2kmat(1.4, 1.5, 64, 48), g['u_kmat'], atol=1e-12)
    assert np.allclose(U.get_centre(box[1]['pt'], box[1]['x_dir'], box[1]['y_dir'], box[1]['w'], box[1]['h'], off_x=0, off_y=0), g['u_centre'])
    assert np.allclose(U.lookat_rotation(g['lookat_delta']), g['lookat_rot'], atol=1e-12)
    for ix in range(3):
        c = U.get_centre(box[ix]['pt'], box[ix]['x_dir'], box[ix]['y_dir'], box[ix]['w'], box[ix]['h'], off_x=0, off_y=0)
        rot, t = U.canonical_transform(c, box[ix]['x_dir'], box[ix]['y_dir'])
        assert np.allclose(rot, g['rot_w2s'][ix], atol=1e-12) and np.allclose(t, g['t_w2s'][ix], atol=1e-12)
    assert np.allclose(U.resize_instrinsic(np.arange(9.0).reshape(3, 3), 0.5, 0.25), np.diag([0.5, 0.25, 1]) @ np.arange(9.0).reshape(3, 3))


def test_sample_views_and_world_layout_statistics():
    from lsi.data.syntheticPlanes import data as D
    rs = np.random.RandomState(0)
    for rot, t in D.sample_views(20, _rs=rs):
        assert np.allclose(rot @ rot.T, np.eye(3), atol=1e-12) and abs(np.linalg.det(rot) - 1) < 1e-12
        cam = -rot.T @ t                                   # camera centre: on the z = 0 plane within +-0.5 (data.py:38-40)
        assert abs(cam[2, 0]) < 1e-12 and np.all(np.abs(cam[:2, 0]) <= 0.5)
    gen = D.WorldGenerator(h=32, w=32, n_obj_max=3, n_obj_min=1, n_box_planes=5, _device='cpu')
    rot_w2s, t_w2s, k_w, n_hat_w, a_w, waves, kind = gen.layout()
    assert rot_w2s.shape == (8, 3, 3) and t_w2s.shape == (8, 3, 1) and k_w.shape == (8, 3, 3) and n_hat_w.shape == (8, 1, 3)
    assert np.all(a_w == -1) and list(kind[:5]) == [0] * 5 and set(kind[5:]) <= {1, 2} and (kind[5:] == 1).sum() >= 1
    centres = t_w2s[:, :, 0] + rot_w2s[:, :, 2]          # canonical centre (0,0,1) mapped into the world
    assert np.all(centres[:, 2] >= 2.0 - 1e-9) and np.all(centres[:, 2] <= 3.5 + 1e-9)      # every plane sits inside the box depth range


@pytest.mark.gpu
def test_fused_renderer_matches_oracle_and_reference_fixture():
    from lsi.data.syntheticPlanes import data as D
    from oracle import lsi_oracle_planes as P
    g = load_golden('planes_render')
    n_box, n_obj, h, w = (int(v) for v in g['meta'])
    n = n_box + n_obj
    imgs = torch.tensor(g['imgs_w'], dtype=torch.float32, device='cuda')[None]
    masks = torch.tensor(g['masks_w'], dtype=torch.float32, device='cuda')[None]
    views = [(np.eye(3), np.zeros((3, 1))), (g['v1_rot'], g['v1_t'])]
    hom, dmat = [], []
    for rot, t in views:
        h_, d_, nh, a = D._t2w_matrices(g['k_w'], g['k_cam'], g['rot_w2s'], g['t_w2s'], g['n_hat_w'], g['a_w'], rot, t)
        hom.append(h_); dmat.append(d_)
    assert rel_err(hom[1].reshape(n, 3, 3), g['v1_inv_hom_f64']) < 1e-12
    assert rel_err(nh, g['v1_n_hat_t_f64']) < 1e-12 and rel_err(a, g['v1_a_t_f64']) < 1e-12
    img, fg, bg = D.render_views(imgs.expand(2, -1, -1, -1, -1).contiguous(), masks.expand(2, -1, -1, -1, -1).contiguous(),
                                 np.stack(hom), np.stack(dmat), h, w)
    for vi in range(2):
        for got, key in ((img[vi], 'render'), (fg[vi], 'disp_fg'), (bg[vi], 'disp_bg')):
            ref = g['v%d_%s_f64' % (vi, key)]
            bad = np.abs(got.cpu().numpy() - ref) > 2e-5 * max(np.abs(ref).max(), 1e-30)
            # hard arg-max selection: an fp32 tie at a layer boundary may fall either way for isolated pixels
            assert bad.mean() <= 0.006, (vi, key, float(bad.mean()))
    # Renderer API (data.py:293-516) on the same world
    r = D.Renderer(n, h=h, w=w)
    r.set_feed_dict(k_w=g['k_w'], k_s=g['k_cam'], k_t=g['k_cam'], rot_w2s=g['rot_w2s'], t_w2s=g['t_w2s'], n_hat_w=g['n_hat_w'],
                    a_w=g['a_w'], imgs_w=imgs[0], masks_w=masks[0])
    assert torch.equal(r.render_planes(g['v1_rot'], g['v1_t']), img[1])
    d_fg, d_bg = r.render_disps(g['v1_rot'], g['v1_t'])
    assert torch.equal(d_fg, fg[1]) and torch.equal(d_bg, bg[1])
    nh2, a2 = r.plane_geometry(g['v1_rot'], g['v1_t'])
    assert rel_err(nh2, g['v1_n_hat_t_f64']) < 1e-12 and rel_err(a2, g['v1_a_t_f64']) < 1e-12


@pytest.mark.gpu
def test_per_plane_api_matches_reference_fixture():
    """lsi.geometry.homography / lsi.geometry.layers (API parity for code written against the reference)."""
    from lsi.geometry import homography, layers
    g = load_golden('planes_render')
    n_box, n_obj, h, w = (int(v) for v in g['meta'])
    n = n_box + n_obj
    cu = lambda a: torch.tensor(np.asarray(a), dtype=torch.float32, device='cuda')
    ys, xs = np.meshgrid(np.arange(h) + 0.5, np.arange(w) + 0.5, indexing='ij')
    pc = cu(np.stack([xs, ys, np.ones_like(xs)], -1))
    rep = lambda x: x.unsqueeze(0).expand(n, *x.shape)
    rot, t = cu(g['v1_rot']), cu(g['v1_t'])
    rot_w2t, t_w2t = rep(rot) @ cu(g['rot_w2s']), rep(t) + rep(rot) @ cu(g['t_w2s'])
    args = (cu(g['k_w']), rep(cu(g['k_cam'])), rot_w2t, t_w2t, cu(g['n_hat_w']), cu(g['a_w']))
    assert rel_err(homography.inv_homography(*args).cpu(), g['v1_inv_hom_f64']) < 1e-4
    imgs = homography.transform_plane_imgs(cu(g['imgs_w']), rep(pc).contiguous(), *args)
    masks = homography.transform_plane_imgs(cu(g['masks_w']), rep(pc).contiguous(), *args)
    dm = homography.trg_disp_maps(rep(pc), *args[1:])
    assert rel_err(imgs.cpu(), g['v1_imgs_w2t_f64']) < 2e-4 and rel_err(masks.cpu(), g['v1_masks_w2t_f64']) < 2e-4
    assert rel_err(dm.cpu(), g['v1_dmats_f64']) < 1e-5
    i64, m64, d64 = cu(g['v1_imgs_w2t_f64']), cu(g['v1_masks_w2t_f64']), cu(g['v1_dmats_f64'])
    kw = dict(min_disp=2e-1, depth_softmax_temp=0.4)
    assert rel_err(layers.compose(i64, m64, d64, soft=True, **kw).cpu(), g['v1_render_soft_f64']) < 1e-5
    for got, key in ((layers.compose(i64, m64, d64, soft=False, **kw), 'render'), (layers.compose_depth(m64, d64, bg_layer=False, **kw), 'disp_fg'),
                     (layers.compose_depth(m64, d64, bg_layer=True, **kw), 'disp_bg')):
        ref = g['v1_%s_f64' % key]
        assert (np.abs(got.cpu().numpy() - ref) > 2e-5 * np.abs(ref).max()).mean() <= 0.006, key
    kc = cu(g['k_cam'])
    im2, mk2, dm2 = layers.planar_transform(i64, m64, pc, kc, kc, rot.T.contiguous(), -rot.T @ t, cu(g['v1_n_hat_t_f64']), cu(g['v1_a_t_f64']))
    assert rel_err(im2.cpu(), g['v1_pt_imgs_f64']) < 5e-4 and rel_err(dm2.cpu(), g['v1_pt_dmaps_f64']) < 1e-4


@pytest.mark.gpu
def test_data_loader_contract():
    """DataLoader.forward (data.py:642-673): shapes, value ranges, determinism per seed, geometric consistency of the pair."""
    from lsi.data.syntheticPlanes import data as D
    from lsi.geometry import projection
    opts = types.SimpleNamespace(img_height=64, img_width=64, synth_ds_factor=2, n_obj_max=2, n_obj_min=1, n_box_planes=5,
                                 data_split='train', synth_dl_eval_data=True, sun_imgs_dir=None, pascal_objects_dir=None)
    out = D.DataLoader(opts, _seed=3).forward(3)
    names = ['img_s', 'img_t', 'k_s', 'k_t', 'rot', 'trans', 'n_hat', 'a', 'disp_s_fg', 'disp_s_bg', 'disp_t_fg', 'disp_t_bg', 'img_s_bg', 'img_t_bg']
    assert len(out) == len(names)
    o = dict(zip(names, out))
    assert tuple(o['img_s'].shape) == (3, 64, 64, 3) and tuple(o['disp_t_bg'].shape) == (3, 64, 64, 1) and tuple(o['n_hat'].shape) == (3, 7, 1, 3)
    assert tuple(o['k_s'].shape) == (3, 3, 3) and tuple(o['trans'].shape) == (3, 3, 1)
    for k in ('img_s', 'img_t', 'img_s_bg'):
        assert float(o[k].min()) >= 0.0 and float(o[k].max()) <= 1.0 and float(o[k].std()) > 0.02
    assert float(o['disp_s_fg'].min()) >= 0.2 - 1e-6 and float(o['disp_s_fg'].max()) <= 0.5 + 1e-3      # box depth 2 .. 3.5, bg plane at 1/5
    assert bool((o['disp_s_fg'] >= o['disp_s_bg'] - 1e-6).all())       # the background-only world is never closer than the full world
    assert float((o['disp_s_fg'] - o['disp_s_bg']).abs().max()) > 0.02  # ... and the objects do stand in front of it somewhere
    assert np.allclose(o['k_s'][0].cpu().numpy(), [[64, 0, 32], [0, 64, 32], [0, 0, 1]])
    again = D.DataLoader(opts, _seed=3).forward(3)
    assert all(torch.equal(a, b) for a, b in zip(out, again))
    other = D.DataLoader(opts, _seed=4).forward(3)
    assert not torch.equal(out[0], other[0])
    # the pair is geometrically consistent: splatting the source view with its ground-truth disparity lands on the target view
    from lsi.geometry import ldi as ldi_utils
    from lsi.nnutils import helpers
    B, H, W = 3, 64, 64
    ldi = (o['img_s'][None].contiguous(), torch.ones(1, B, H, W, 1, device='cuda'), o['disp_s_fg'][None].contiguous())
    img, wts = ldi_utils.forward_splat(ldi, helpers.pixel_coords(B, H, W), o['k_s'], o['k_t'], o['rot'], o['trans'], compose_layers=True,
                                       bg_layer_disp=0.0, max_disp=1.0, zbuf_scale=0)      # scale 0: unit weights, wts = coverage
    seen = (wts[0, ..., 0] > 0.5)
    err = (img[0] - o['img_t']).abs().mean(dim=-1)[seen]
    assert float(seen.float().mean()) > 0.5 and float(err.median()) < 0.08, (float(seen.float().mean()), float(err.median()))
