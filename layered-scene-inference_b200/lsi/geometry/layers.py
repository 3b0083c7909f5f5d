"""Mirror of lsi/geometry/layers.py (reference tree): composition of planar layers.

Same names and arguments as the reference; elementwise torch ops over the (few) layers on the caller's device -- API parity for
code written against the reference.  The synthetic-data path does not go through these per-layer tensors: see
lsi.data.syntheticPlanes (fused lsi_b200_render_planes kernel).
"""
import torch

from lsi.geometry import homography
from lsi.nnutils import helpers as nn_helpers


def _hard(selection_mask):
    """tf.one_hot(tf.argmax(sel), L, axis=0) (layers.py:66-67): [TF1.4] argmax without axis is over axis 0."""
    idx = torch.argmax(selection_mask, dim=0, keepdim=True)
    return torch.zeros_like(selection_mask).scatter_(0, idx, 1.0)


def compose(imgs, masks, dmaps, soft=False, min_disp=1e-6, depth_softmax_temp=1):
    """layers.py:29-74 -- [L,...,C] layers -> [...,C]: a white background layer at disparity min_disp is appended, the layers are
    soft z-buffered and, unless soft, the most probable one is selected per pixel."""
    dmaps = torch.relu(dmaps)
    imgs = torch.cat([imgs, torch.ones_like(imgs[:1])], 0)
    masks = torch.cat([masks, torch.ones_like(masks[:1])], 0)
    dmaps = torch.cat([dmaps, torch.ones_like(dmaps[:1]) * min_disp], 0)
    sel = nn_helpers.soft_z_buffering(masks, dmaps, depth_softmax_temp=depth_softmax_temp)
    if not soft:
        sel = _hard(sel)
    return (sel * imgs).sum(dim=0)


def compose_depth(masks, dmaps, bg_layer=False, min_disp=1e-6, depth_softmax_temp=1):
    """layers.py:77-118 -- disparity of the selected layer; bg_layer=True selects the FARTHEST valid layer (selection by the
    global maximum disparity minus the layer's disparity)."""
    n_layers = masks.shape[0]
    dmaps = torch.relu(dmaps)
    masks = torch.cat([masks, torch.ones_like(masks[:1])], 0)
    dmaps = torch.cat([dmaps, torch.ones_like(dmaps[:1]) * min_disp], 0)
    dsel = torch.cat([dmaps.max() - dmaps[0:n_layers], dmaps[n_layers:]], 0) if bg_layer else dmaps
    sel = _hard(nn_helpers.soft_z_buffering(masks, dsel, depth_softmax_temp=depth_softmax_temp))
    return (sel * dmaps).sum(dim=0)


def planar_transform(imgs, masks, pixel_coords_trg, k_s, k_t, rot, t, n_hat, a):
    """layers.py:121-162 -- warp [L,...] layer images and masks into the target frame and compute their disparity maps."""
    n_layers = imgs.shape[0]
    rep = lambda x: x.unsqueeze(0).expand(n_layers, *x.shape)
    pc = rep(pixel_coords_trg)
    both = homography.transform_plane_imgs(torch.cat([imgs, masks], dim=-1).contiguous(), pc, rep(k_s), rep(k_t), rep(rot), rep(t), n_hat, a)
    return both[..., 0:3], both[..., 3:4], homography.trg_disp_maps(pc, rep(k_t), rep(rot), rep(t), n_hat, a)
