"""ORACLE -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement (torch / numpy) of the reference's synthetic planar-room generator: lsi/geometry/homography.py:28-156
(plane-induced homographies, plane equations, per-pixel disparities), lsi/geometry/layers.py:29-162 (hard / soft layer
composition, depth composition, planar_transform), the renderer wiring of lsi/data/syntheticPlanes/data.py:372-420
("w2t_rendering") and the world-layout helpers of lsi/data/syntheticPlanes/utils.py:36-203.

Parity status: pinned to the reference's own homography.py / layers.py / utils.py through tests/golden/planes_render.npz
(oracle/gen_golden_planes.py runs those sources over the TF shim); [TF1.4] semantics restated in the shim: tf.argmax with
axis=None is axis 0, one_hot(axis=0), matrix_inverse, bilinear sampling as in lsi_oracle.bilinear.
"""
import math

import numpy as np
import torch

from oracle import lsi_oracle as O


def _T(x):
    return x.transpose(-1, -2)


# ---- lsi/geometry/homography.py ----------------------------------------------------------------------------------------
def inv_homography(k_s, k_t, rot, t, n_hat, a):
    """homography.py:28-52: K_s (R^T + R^T t n^T R^T / (a - n^T R^T t)) K_t^-1."""
    rot_t = _T(rot)
    denom = a - n_hat @ rot_t @ t
    numerator = rot_t @ t @ n_hat @ rot_t
    return k_s @ (rot_t + O.divide_safe(numerator, denom)) @ torch.linalg.inv(k_t)


def inv_homography_dmat(k_t, rot, t, n_hat, a):
    """homography.py:55-75: the row vector M with M (u, v, 1)^T = disparity in the target frame."""
    rot_t = _T(rot)
    denom = a - n_hat @ rot_t @ t
    return O.divide_safe(-1 * (n_hat @ rot_t @ torch.linalg.inv(k_t)), denom)


def transform_plane_imgs(imgs, pixel_coords_trg, k_s, k_t, rot, t, n_hat, a):
    """homography.py:95-118: inverse-warp every plane's image with its homography (bilinear, zero outside)."""
    hom = inv_homography(k_s, k_t, rot, t, n_hat, a)
    pts = O.transform_pts(pixel_coords_trg, hom)
    uv = O.divide_safe(pts[..., :2], pts[..., 2:3])
    lead = list(imgs.shape[:-3])
    out = O.bilinear(imgs.reshape([-1] + list(imgs.shape[-3:])), uv.reshape([-1] + list(uv.shape[-3:])))
    return out.reshape(lead + list(out.shape[-3:]))


def transform_plane_eqns(rot, t, n_hat, a):
    """homography.py:121-137."""
    rot_t = _T(rot)
    return n_hat @ rot_t, a - n_hat @ (rot_t @ t)


def trg_disp_maps(pixel_coords_trg, k_t, rot, t, n_hat, a):
    """homography.py:140-156."""
    dm = inv_homography_dmat(k_t, rot, t, n_hat, a)
    return (dm.unsqueeze(-2) * pixel_coords_trg).sum(dim=-1, keepdim=True)


# ---- lsi/geometry/layers.py ----------------------------------------------------------------------------------------------
def _with_bg(imgs, masks, dmaps, min_disp):
    dmaps = torch.relu(dmaps)
    one = torch.ones_like(masks[:1])
    out_imgs = None if imgs is None else torch.cat([imgs, torch.ones_like(imgs[:1])], 0)
    return out_imgs, torch.cat([masks, one], 0), torch.cat([dmaps, one * min_disp], 0)


def compose(imgs, masks, dmaps, soft=False, min_disp=1e-6, depth_softmax_temp=1):
    """layers.py:29-74: append a white background layer at disparity min_disp, soft z-buffer, hard selection unless soft."""
    n_layers = imgs.shape[0]
    imgs, masks, dmaps = _with_bg(imgs, masks, dmaps, min_disp)
    sel = O.soft_z_buffering(masks, dmaps, depth_softmax_temp=depth_softmax_temp)
    if not soft:
        sel = torch.nn.functional.one_hot(torch.argmax(sel, dim=0), n_layers + 1).to(imgs.dtype).movedim(-1, 0)
    return (sel * imgs).sum(dim=0)


def compose_depth(masks, dmaps, bg_layer=False, min_disp=1e-6, depth_softmax_temp=1):
    """layers.py:77-118: disparity of the selected layer; bg_layer=True selects by (global max disparity - disparity)."""
    n_layers = masks.shape[0]
    _, masks, dmaps = _with_bg(None, masks, dmaps, min_disp)
    if bg_layer:
        dsel = torch.cat([dmaps.max() - dmaps[0:n_layers], dmaps[n_layers:]], 0)
    else:
        dsel = dmaps
    sel = O.soft_z_buffering(masks, dsel, depth_softmax_temp=depth_softmax_temp)
    sel = torch.nn.functional.one_hot(torch.argmax(sel, dim=0), n_layers + 1).to(dmaps.dtype).movedim(-1, 0)
    return (sel * dmaps).sum(dim=0)


def planar_transform(imgs, masks, pixel_coords_trg, k_s, k_t, rot, t, n_hat, a):
    """layers.py:121-162."""
    L = imgs.shape[0]
    rep = lambda x: x.unsqueeze(0).expand(L, *x.shape)
    pc = rep(pixel_coords_trg)
    im = transform_plane_imgs(torch.cat([imgs, masks], dim=-1), pc, rep(k_s), rep(k_t), rep(rot), rep(t), n_hat, a)
    return im[..., :3], im[..., 3:4], trg_disp_maps(pc, rep(k_t), rep(rot), rep(t), n_hat, a)


# ---- lsi/data/syntheticPlanes/data.py:372-420 ("w2t_rendering") -------------------------------------------------------------
def render_planes(world, k_cam, pixel_coords, rot_s2t, t_s2t, min_disp=2e-1, depth_softmax_temp=0.4):
    """-> dict(render [h,w,3], disp_fg [h,w,1], disp_bg [h,w,1], n_hat_t [n,1,3], a_t [n,1,1])."""
    n = world['imgs_w'].shape[0]
    rep = lambda x: x.unsqueeze(0).expand(n, *x.shape)
    rot_w2t = rep(rot_s2t) @ world['rot_w2s']
    t_w2t = rep(t_s2t) + rep(rot_s2t) @ world['t_w2s']
    pc, k_t = rep(pixel_coords), rep(k_cam)
    args = (world['k_w'], k_t, rot_w2t, t_w2t, world['n_hat_w'], world['a_w'])
    imgs = transform_plane_imgs(world['imgs_w'], pc, *args)
    masks = transform_plane_imgs(world['masks_w'], pc, *args)
    dm = trg_disp_maps(pc, k_t, rot_w2t, t_w2t, world['n_hat_w'], world['a_w'])
    n_hat_t, a_t = transform_plane_eqns(rot_w2t, t_w2t, world['n_hat_w'], world['a_w'])
    kw = dict(min_disp=min_disp, depth_softmax_temp=depth_softmax_temp)
    return dict(render=compose(imgs, masks, dm, soft=False, **kw), disp_fg=compose_depth(masks, dm, bg_layer=False, **kw),
                disp_bg=compose_depth(masks, dm, bg_layer=True, **kw), n_hat_t=n_hat_t, a_t=a_t, imgs_w2t=imgs, masks_w2t=masks, dmats=dm)


# ---- lsi/data/syntheticPlanes/utils.py ----------------------------------------------------------------------------------------
def dims2kmat(w_plane, h_plane, w_tex, h_tex):
    """utils.py:36-51: intrinsics of a texture image glued on a fronto-parallel plane at z = 1."""
    return np.array([[w_tex / w_plane, 0, w_tex / 2.0], [0, h_tex / h_plane, h_tex / 2.0], [0, 0, 1.0]])


def _unit(v):
    v = np.asarray(v, dtype=np.float64).reshape(3, 1)
    return v / np.linalg.norm(v)


def get_centre(pt, x_dir, y_dir, w, h, off_x=0.5, off_y=0.5):
    """utils.py:54-75."""
    return np.asarray(pt, dtype=np.float64).reshape(3, 1) + w * _unit(x_dir) * (0.5 - off_x) + h * _unit(y_dir) * (0.5 - off_y)


def canonical_transform(centre_s, x_dir, y_dir):
    """utils.py:78-106: the rigid motion that takes the canonical plane (centre (0,0,1), axes x, y) to the given pose."""
    x, y = _unit(x_dir), _unit(y_dir)
    rot = np.concatenate([x, y, np.cross(x, y, axis=0)], axis=1)
    return rot, np.asarray(centre_s, dtype=np.float64).reshape(3, 1) - rot @ np.array([[0.0], [0.0], [1.0]])


def box_planes(extent):
    """utils.py:109-176: front wall, floor, ceiling, left wall, right wall of the box (in that order)."""
    x0, y0, z0, x1, y1, z1 = extent
    ex, ey, ez = np.eye(3)
    mk = lambda pt, xd, yd, w, h: {'pt': np.array(pt, dtype=np.float64), 'x_dir': xd, 'y_dir': yd, 'w': w, 'h': h, 'off_x': 0, 'off_y': 0}
    return [mk([x0, y0, z1], ex, ey, x1 - x0, y1 - y0), mk([x0, y1, z1], ex, -ez, x1 - x0, z1 - z0),
            mk([x0, y0, z1], ex, -ez, x1 - x0, z1 - z0), mk([x0, y0, z0], ez, ey, z1 - z0, y1 - y0),
            mk([x1, y0, z0], ez, ey, z1 - z0, y1 - y0)]


def lookat_rotation(delta):
    """utils.py:189-203: rotation that takes the direction delta onto the z axis."""
    d = np.asarray(delta, dtype=np.float64).reshape(3)
    theta, phi = math.atan2(d[0], d[2]), math.asin(d[1] / np.linalg.norm(d))
    ry = np.array([[math.cos(-theta), 0, math.sin(-theta)], [0, 1, 0], [-math.sin(-theta), 0, math.cos(-theta)]])
    rx = np.array([[1, 0, 0], [0, math.cos(phi), -math.sin(phi)], [0, math.sin(phi), math.cos(phi)]])
    return rx @ ry
