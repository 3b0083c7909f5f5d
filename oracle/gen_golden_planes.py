"""Golden fixtures for the synthetic planar-room generator: the reference's own, unmodified
lsi/geometry/homography.py and lsi/geometry/layers.py run over the eager TF-1 shim, wired exactly as the renderer of
lsi/data/syntheticPlanes/data.py:293-450 wires them (planar_rendering, rendering_disp_fg / _bg, plane geometry), plus the
world-layout helpers of lsi/data/syntheticPlanes/utils.py:36-203 (that file is Python 2 -- print statements -- and imports
matplotlib / absl: its source is read, the print statements are parenthesised IN MEMORY and the module is executed with stub
imports; nothing is copied).
    python oracle/gen_golden_planes.py  ->  tests/golden/planes_render.npz   (test infrastructure; needs /root/reference)"""
import os
import re
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = '/root/reference'
sys.path.insert(0, os.path.join(HERE, 'tf1_shim'))
sys.path.insert(1, REF)


def _load_utils():
    src = open(os.path.join(REF, 'lsi', 'data', 'syntheticPlanes', 'utils.py')).read()
    src = re.sub(r'^(\s*)print (.+)$', r'\1print(\2)', src, flags=re.M)
    for name in ('absl', 'absl.logging', 'matplotlib', 'matplotlib.pyplot'):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules['absl'].logging = sys.modules['absl.logging']
    sys.modules['matplotlib'].pyplot = sys.modules['matplotlib.pyplot']
    mod = types.ModuleType('ref_planes_utils')
    exec(compile(src, 'syntheticPlanes/utils.py', 'exec'), mod.__dict__)
    return mod


def make_world(rs, U, n_box, n_obj, h, w):
    """A random box world the way WorldGenerator.forward (data.py:207-291) lays it out, with random textures."""
    extent = [-0.7, -0.5, 2.0, 0.7, 0.5, 3.5]
    planes = U.box_planes(extent)[0:n_box]
    for ix in range(n_obj):           # billboards standing on the floor (random_obj_plane, data.py:145-205, fixed_plane=ix)
        w_obj, h_obj = rs.uniform(0.4, 0.6) * 1.4, rs.uniform(0.3, 0.5)
        planes.append({'pt': np.array([extent[0] + 0.35 + 0.35 * ix, extent[4], extent[2] + 0.3 * ix]), 'x_dir': np.array([1, 0, 0]),
                       'y_dir': np.array([0, 1, 0]), 'w': w_obj, 'h': h_obj, 'off_x': 0.5, 'off_y': 1})
    bs = len(planes)
    rot_w2s, t_w2s, k_w = np.zeros((bs, 3, 3)), np.zeros((bs, 3, 1)), np.zeros((bs, 3, 3))
    for ix, pl in enumerate(planes):
        c = U.get_centre(pl['pt'], pl['x_dir'], pl['y_dir'], pl['w'], pl['h'], off_x=pl['off_x'], off_y=pl['off_y'])
        rot_w2s[ix], t_w2s[ix] = U.canonical_transform(c, pl['x_dir'], pl['y_dir'])
        k_w[ix] = U.dims2kmat(pl['w'], pl['h'], h, w)
    n_hat_w = np.tile(np.array([[[0.0, 0.0, 1.0]]]), (bs, 1, 1))
    a_w = -np.ones((bs, 1, 1))
    imgs_w = rs.uniform(0, 1, (bs, h, w, 3))
    masks_w = np.ones((bs, h, w, 1))
    masks_w[n_box:] = (rs.uniform(0, 1, (n_obj, h, w, 1)) > 0.35).astype(np.float64)
    return dict(rot_w2s=rot_w2s, t_w2s=t_w2s, k_w=k_w, n_hat_w=n_hat_w, a_w=a_w, imgs_w=imgs_w, masks_w=masks_w, planes=planes)


def main():
    import builtins
    import tensorflow as tf
    from lsi.nnutils import helpers
    helpers.range = lambda *a: list(builtins.range(*a))       # Python-2 range() semantics for helpers.transpose (helpers.py:75-77)
    from lsi.geometry import homography, layers
    U = _load_utils()
    T = tf.Tensor
    rs = np.random.RandomState(17)
    n_box, n_obj, h, w = 3, 2, 20, 28
    n = n_box + n_obj
    wd = make_world(rs, U, n_box, n_obj, h, w)
    k_cam = np.array([[w, 0, w / 2.0], [0, h, h / 2.0], [0, 0, 1.0]])          # data.py:548-557
    ys, xs = np.meshgrid(np.arange(h) + 0.5, np.arange(w) + 0.5, indexing='ij')
    pc = np.stack([xs, ys, np.ones_like(xs)], -1)
    # one identity view and one sample_views-like view (data.py:29-52)
    cam = np.array([0.31, -0.22, 0.0]).reshape(3, 1)
    rot_v = U.lookat_rotation(np.array([-0.12, 0.18, 3.2]).reshape(3, 1) - cam)
    views = [(np.eye(3), np.zeros((3, 1))), (rot_v, -rot_v @ cam)]
    blob = {k: wd[k] for k in ('rot_w2s', 't_w2s', 'k_w', 'n_hat_w', 'a_w', 'imgs_w', 'masks_w')}
    blob.update(k_cam=k_cam, meta=np.array([n_box, n_obj, h, w], dtype=np.int64), lookat_delta=np.array([-0.43, 0.40, 3.2]),
                lookat_rot=rot_v)
    # world-layout helpers on their own
    pl = wd['planes'][1]
    blob['u_centre'] = U.get_centre(pl['pt'], pl['x_dir'], pl['y_dir'], pl['w'], pl['h'], off_x=pl['off_x'], off_y=pl['off_y'])
    blob['u_kmat'] = U.dims2kmat(1.4, 1.5, 64, 48)
    box = U.box_planes([-0.7, -0.5, 2.0, 0.7, 0.5, 3.5])
    blob['u_box_pt'] = np.stack([b['pt'] for b in box]); blob['u_box_x'] = np.stack([b['x_dir'] for b in box])
    blob['u_box_y'] = np.stack([b['y_dir'] for b in box]); blob['u_box_wh'] = np.array([[b['w'], b['h']] for b in box])
    for dtype, sfx in ((torch.float32, '_f32'), (torch.float64, '_f64')):
        tf._set_float(dtype)
        t = lambda a: T(torch.tensor(np.asarray(a), dtype=dtype))
        for vi, (rot_s2t, t_s2t) in enumerate(views):
            # data.py:372-420 ("w2t_rendering"), op for op
            rot_rep = t(np.tile(rot_s2t[None], (n, 1, 1)))
            t_rep = t(np.tile(t_s2t[None], (n, 1, 1)))
            k_t = t(np.tile(k_cam[None], (n, 1, 1)))
            pcs = t(np.tile(pc[None], (n, 1, 1, 1)))
            rot_w2t = tf.matmul(rot_rep, t(wd['rot_w2s']))
            t_w2t = t_rep + tf.matmul(rot_rep, t(wd['t_w2s']))
            imgs_w2t = homography.transform_plane_imgs(t(wd['imgs_w']), pcs, t(wd['k_w']), k_t, rot_w2t, t_w2t, t(wd['n_hat_w']), t(wd['a_w']))
            masks_w2t = homography.transform_plane_imgs(t(wd['masks_w']), pcs, t(wd['k_w']), k_t, rot_w2t, t_w2t, t(wd['n_hat_w']), t(wd['a_w']))
            dmats = homography.trg_disp_maps(pcs, k_t, rot_w2t, t_w2t, t(wd['n_hat_w']), t(wd['a_w']))
            n_hat_t, a_t = homography.transform_plane_eqns(rot_w2t, t_w2t, t(wd['n_hat_w']), t(wd['a_w']))
            kw = dict(min_disp=2e-1, depth_softmax_temp=0.4)
            blob['v%d_render%s' % (vi, sfx)] = layers.compose(imgs_w2t, masks_w2t, dmats, soft=False, **kw).t.numpy()
            blob['v%d_disp_fg%s' % (vi, sfx)] = layers.compose_depth(masks_w2t, dmats, bg_layer=False, **kw).t.numpy()
            blob['v%d_disp_bg%s' % (vi, sfx)] = layers.compose_depth(masks_w2t, dmats, bg_layer=True, **kw).t.numpy()
            blob['v%d_imgs_w2t%s' % (vi, sfx)] = imgs_w2t.t.numpy()
            blob['v%d_masks_w2t%s' % (vi, sfx)] = masks_w2t.t.numpy()
            blob['v%d_dmats%s' % (vi, sfx)] = dmats.t.numpy()
            blob['v%d_n_hat_t%s' % (vi, sfx)] = n_hat_t.t.numpy(); blob['v%d_a_t%s' % (vi, sfx)] = a_t.t.numpy()
            if vi == 1:
                blob['v1_rot'], blob['v1_t'] = rot_s2t, t_s2t
                blob['v1_inv_hom' + sfx] = homography.inv_homography(t(wd['k_w']), k_t, rot_w2t, t_w2t, t(wd['n_hat_w']), t(wd['a_w'])).t.numpy()
                # the soft composition and planar_transform (layers.py:29-74, 121-162) for API parity
                blob['v1_render_soft' + sfx] = layers.compose(imgs_w2t, masks_w2t, dmats, soft=True, **kw).t.numpy()
                im2, mk2, dm2 = layers.planar_transform(imgs_w2t, masks_w2t, t(pc), t(k_cam), t(k_cam), t(views[1][0].T),
                                                        t(-views[1][0].T @ views[1][1]), n_hat_t, a_t)
                blob['v1_pt_imgs' + sfx], blob['v1_pt_masks' + sfx], blob['v1_pt_dmaps' + sfx] = im2.t.numpy(), mk2.t.numpy(), dm2.t.numpy()
    path = os.path.join(ROOT, 'tests', 'golden', 'planes_render.npz')
    np.savez_compressed(path, **blob)
    print(path, os.path.getsize(path), 'bytes')


if __name__ == '__main__':
    main()
