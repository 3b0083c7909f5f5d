"""CPU: the oracle of the synthetic planar-room generator (oracle/lsi_oracle_planes.py) against the fixture produced by the
reference's own homography.py / layers.py / syntheticPlanes/utils.py (oracle/gen_golden_planes.py)."""
import numpy as np
import torch

from _util import load_golden, rel_err
from oracle import lsi_oracle_planes as P


def _world(g, dt):
    return {k: torch.tensor(g[k], dtype=dt) for k in ('rot_w2s', 't_w2s', 'k_w', 'n_hat_w', 'a_w', 'imgs_w', 'masks_w')}


def _pc(h, w, dt):
    ys, xs = np.meshgrid(np.arange(h) + 0.5, np.arange(w) + 0.5, indexing='ij')
    return torch.tensor(np.stack([xs, ys, np.ones_like(xs)], -1), dtype=dt)


def test_renderer_matches_reference_sources():
    g = load_golden('planes_render')
    n_box, n_obj, h, w = (int(v) for v in g['meta'])
    for dt, sfx, tol in ((torch.float64, '_f64', 1e-12), (torch.float32, '_f32', 1e-5)):
        wd = _world(g, dt)
        views = [(torch.eye(3, dtype=dt), torch.zeros(3, 1, dtype=dt)), (torch.tensor(g['v1_rot'], dtype=dt), torch.tensor(g['v1_t'], dtype=dt))]
        for vi, (rot, t) in enumerate(views):
            out = P.render_planes(wd, torch.tensor(g['k_cam'], dtype=dt), _pc(h, w, dt), rot, t)
            for key, name in (('render', 'render'), ('disp_fg', 'disp_fg'), ('disp_bg', 'disp_bg'), ('imgs_w2t', 'imgs_w2t'),
                              ('masks_w2t', 'masks_w2t'), ('dmats', 'dmats'), ('n_hat_t', 'n_hat_t'), ('a_t', 'a_t')):
                ref = g['v%d_%s%s' % (vi, name, sfx)]
                bad = np.abs(out[key].numpy() - ref) > tol * max(np.abs(ref).max(), 1e-30)
                # hard arg-max selection: in fp32 a tie at a layer boundary may fall either way for isolated pixels
                assert bad.mean() <= (0.0 if dt == torch.float64 else 0.004), (vi, key, sfx, float(bad.mean()))
    dt = torch.float64
    wd = _world(g, dt)
    n = n_box + n_obj
    rot, t = torch.tensor(g['v1_rot'], dtype=dt), torch.tensor(g['v1_t'], dtype=dt)
    rep = lambda x: x.unsqueeze(0).expand(n, *x.shape)
    rot_w2t, t_w2t = rep(rot) @ wd['rot_w2s'], rep(t) + rep(rot) @ wd['t_w2s']
    k_t = rep(torch.tensor(g['k_cam'], dtype=dt))
    assert rel_err(P.inv_homography(wd['k_w'], k_t, rot_w2t, t_w2t, wd['n_hat_w'], wd['a_w']), g['v1_inv_hom_f64']) < 1e-12
    imgs, masks, dm = (torch.tensor(g['v1_%s_f64' % k]) for k in ('imgs_w2t', 'masks_w2t', 'dmats'))
    assert rel_err(P.compose(imgs, masks, dm, soft=True, min_disp=2e-1, depth_softmax_temp=0.4), g['v1_render_soft_f64']) < 1e-12
    n_hat_t, a_t = torch.tensor(g['v1_n_hat_t_f64']), torch.tensor(g['v1_a_t_f64'])
    kc = torch.tensor(g['k_cam'], dtype=dt)
    im2, mk2, dm2 = P.planar_transform(imgs, masks, _pc(h, w, dt), kc, kc, rot.T, -rot.T @ t, n_hat_t, a_t)
    assert rel_err(im2, g['v1_pt_imgs_f64']) < 1e-12 and rel_err(mk2, g['v1_pt_masks_f64']) < 1e-12 and rel_err(dm2, g['v1_pt_dmaps_f64']) < 1e-12


def test_world_layout_helpers_match_reference_sources():
    g = load_golden('planes_render')
    box = P.box_planes([-0.7, -0.5, 2.0, 0.7, 0.5, 3.5])
    assert np.allclose(np.stack([b['pt'] for b in box]), g['u_box_pt']) and np.allclose(np.stack([b['x_dir'] for b in box]), g['u_box_x'])
    assert np.allclose(np.stack([b['y_dir'] for b in box]), g['u_box_y']) and np.allclose(np.array([[b['w'], b['h']] for b in box]), g['u_box_wh'])
    assert np.allclose(P.dims2kmat(1.4, 1.5, 64, 48), g['u_kmat'], atol=1e-12)
    pl = box[1]
    assert np.allclose(P.get_centre(pl['pt'], pl['x_dir'], pl['y_dir'], pl['w'], pl['h'], off_x=0, off_y=0), g['u_centre'], atol=1e-12)
    assert np.allclose(P.lookat_rotation(g['lookat_delta']), g['lookat_rot'], atol=1e-12)
    # the planes of the fixture world were placed with canonical_transform(get_centre(...)): reproduce the box part
    for ix in range(3):
        c = P.get_centre(box[ix]['pt'], box[ix]['x_dir'], box[ix]['y_dir'], box[ix]['w'], box[ix]['h'], off_x=0, off_y=0)
        rot, t = P.canonical_transform(c, box[ix]['x_dir'], box[ix]['y_dir'])
        assert np.allclose(rot, g['rot_w2s'][ix], atol=1e-12) and np.allclose(t, g['t_w2s'][ix], atol=1e-12)
