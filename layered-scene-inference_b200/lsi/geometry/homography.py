"""Mirror of lsi/geometry/homography.py (reference tree): plane-induced homographies for the synthetic planar-room generator.

Same names and arguments as the reference.  The 3x3 algebra runs as batched torch ops on whatever device the inputs live on
(a handful of matrices per scene); image warping goes through lsi.geometry.sampling.bilinear_wrapper (CUDA).  The data
generator itself does not call these per plane: lsi.data.syntheticPlanes renders a batch of scenes with one fused kernel
(lsi_b200_render_planes) fed with the matrices computed here.
"""
import torch

from lsi.geometry import sampling
from lsi.nnutils import helpers as nn_helpers


def inv_homography(k_s, k_t, rot, t, n_hat, a):
    """homography.py:28-52 -- [...,3,3] matrices taking target pixels to source pixels for the plane n_hat . x = a."""
    rot_t = nn_helpers.transpose(rot)
    denom = a - torch.matmul(torch.matmul(n_hat, rot_t), t)
    numerator = torch.matmul(torch.matmul(torch.matmul(rot_t, t), n_hat), rot_t)
    return torch.matmul(torch.matmul(k_s, rot_t + nn_helpers.divide_safe(numerator, denom)), torch.linalg.inv(k_t))


def inv_homography_dmat(k_t, rot, t, n_hat, a):
    """homography.py:55-75 -- [...,1,3] row vectors M with M (u, v, 1)^T = disparity of the plane at target pixel (u, v)."""
    rot_t = nn_helpers.transpose(rot)
    denom = a - torch.matmul(torch.matmul(n_hat, rot_t), t)
    return nn_helpers.divide_safe(-1 * torch.matmul(torch.matmul(n_hat, rot_t), torch.linalg.inv(k_t)), denom)


def normalize_homogeneous(pts_coords):
    """homography.py:78-92 -- divide by the last coordinate (divide_safe)."""
    return nn_helpers.divide_safe(pts_coords[..., :-1], pts_coords[..., -1:])


def transform_plane_imgs(imgs, pixel_coords_trg, k_s, k_t, rot, t, n_hat, a):
    """homography.py:95-118 -- warp [...,Hs,Ws,C] images into the target frame through the planes' homographies."""
    hom = inv_homography(k_s, k_t, rot, t, n_hat, a)
    coords = normalize_homogeneous(nn_helpers.transform_pts(pixel_coords_trg, hom))
    return sampling.bilinear_wrapper(imgs, coords.contiguous())


def transform_plane_eqns(rot, t, n_hat, a):
    """homography.py:121-137 -- plane equations in the target frame."""
    rot_t = nn_helpers.transpose(rot)
    return torch.matmul(n_hat, rot_t), a - torch.matmul(n_hat, torch.matmul(rot_t, t))


def trg_disp_maps(pixel_coords_trg, k_t, rot, t, n_hat, a):
    """homography.py:140-156 -- [...,Ht,Wt,1] inverse depth of every plane at every target pixel."""
    dmats_t = inv_homography_dmat(k_t, rot, t, n_hat, a)
    return (dmats_t.unsqueeze(-2) * pixel_coords_trg).sum(dim=-1, keepdim=True)
