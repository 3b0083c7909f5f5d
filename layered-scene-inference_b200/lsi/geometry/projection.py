"""Mirror of lsi/geometry/projection.py (reference tree)."""
import torch

from lsi import _b200
from lsi.geometry import sampling
from lsi.nnutils import helpers as nn_helpers


def pad_intrinsic(k_mat):
    """projection.py:27-46 -- [...,3,3] -> [...,4,4]."""
    k = k_mat
    out = torch.zeros(*k.shape[:-2], 4, 4, dtype=k.dtype, device=k.device)
    out[..., :3, :3] = k
    out[..., 3, 3] = 1
    return out


def pad_extrinsic(rot_mat, trans_mat):
    """projection.py:49-68 -- [R t; 0 1]."""
    rot, trans = rot_mat, trans_mat
    out = torch.zeros(*rot.shape[:-2], 4, 4, dtype=rot.dtype, device=rot.device)
    out[..., :3, :3] = rot
    out[..., :3, 3:4] = trans
    out[..., 3, 3] = 1
    return out


def _cam(k_s, k_t, rot, t):
    k_s, k_t, rot, t = (_b200.dev_f32(x, n) for x, n in ((k_s, 'k_s'), (k_t, 'k_t'), (rot, 'rot'), (t, 't')))
    if k_s.dim() != 3 or k_s.shape[1:] != (3, 3) or k_t.shape != k_s.shape or rot.shape != k_s.shape:
        raise RuntimeError('lsi_b200: k_s, k_t, rot must be [B,3,3], got %s %s %s'
                           % (tuple(k_s.shape), tuple(k_t.shape), tuple(rot.shape)))
    if tuple(t.shape) not in ((k_s.shape[0], 3, 1), (k_s.shape[0], 3)):
        raise RuntimeError('lsi_b200: t must be [B,3,1], got %s' % (tuple(t.shape),))
    return k_s, k_t, rot, t


def _matrix(k_s, k_t, rot, t, inverse):
    k_s, k_t, rot, t = _cam(k_s, k_t, rot, t)
    out = torch.empty(k_s.shape[0], 4, 4, dtype=torch.float32, device=k_s.device)
    _b200.call('lsi_b200_projection_matrix', _b200.ptr(k_s), _b200.ptr(k_t), _b200.ptr(rot), _b200.ptr(t),
               k_s.shape[0], inverse, _b200.ptr(out), _b200.stream())
    return out


def forward_projection_matrix(k_s, k_t, rot, t):
    """projection.py:71-86 -- src->trg 4x4 matrices [B,4,4]."""
    return _matrix(k_s, k_t, rot, t, 0)


def inverse_projection_matrix(k_s, k_t, rot, t):
    """projection.py:89-106 -- trg->src 4x4 matrices [B,4,4]."""
    return _matrix(k_s, k_t, rot, t, 1)


def disocclusion_mask(disps_src, disps_trg, pixel_coords_src, src2trg_mat, thresh=1e-2):
    """projection.py:109-150 (eval path): 1 where the src pixel is visible in trg bounds but its projected
    disparity disagrees with the bilinearly sampled trg disparity."""
    _, h_t, w_t, _ = disps_trg.shape
    pts = nn_helpers.transform_pts(torch.cat([pixel_coords_src.as_subclass(torch.Tensor), disps_src], dim=-1), src2trg_mat)
    uv = nn_helpers.divide_safe(pts[..., 0:2], pts[..., 2:3])
    d12 = nn_helpers.divide_safe(pts[..., 3:4], pts[..., 2:3])
    u, v = uv[..., 0:1], uv[..., 1:2]
    trunc = ((u > w_t).float() + (v > h_t).float() + (u < 0).float() + (v < 0).float()) > 0
    sampled = sampling.bilinear_wrapper(disps_trg, uv, compose=True)
    disocc = (d12 - sampled).abs() > thresh
    return (1 - trunc.float()) * disocc.float()
