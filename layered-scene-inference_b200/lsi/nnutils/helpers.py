"""Mirror of lsi/nnutils/helpers.py (reference tree).  Same names, arguments and semantics, on torch tensors.

On the hot path these are fused into the CUDA kernels (csrc/common.cuh); the functions here exist so that
reference call sites keep working, and are thin elementwise wrappers (API parity, not the measured path).
`optimistic_restorer` (helpers.py:27-62) is implemented in lsi.nnutils.checkpoint (numpy archives keyed by the TF variable
names) and exposed here under the reference's name and signature.
"""
import torch

from lsi import _b200
from lsi.nnutils import checkpoint as _ckpt


def optimistic_restorer(save_file, vars_all=None, _store=None):
    """helpers.py:27-62 -- restorer for the variables that are present in save_file with the same shape.  The reference reads
    tf.global_variables(); here the variables live in a ParamStore (`_store`, default: the default store of lsi.nnutils.nets).
    Returns an object with `.restore(store)`."""
    if _store is None:
        from lsi.nnutils import nets
        _store = nets.get_default_store()
    return _ckpt.optimistic_restorer(save_file, _store, vars_all)


def transpose(rot):
    """helpers.py:65-79 -- swap the last two dimensions."""
    return rot.transpose(-1, -2)


def divide_safe(num, den, name=None):
    """helpers.py:82-85 -- eps only where den == 0 exactly."""
    den = den + 1e-8 * (den == 0).to(den.dtype)
    return num / den


class _StandardGrid(torch.Tensor):
    """Marker subclass: tells the fused renderer that the coordinates are the standard grid, so it can derive
    them from thread indices instead of reading 12 bytes per pixel."""
    pass


def pixel_coords(bs, h, w, _device='cuda'):
    """helpers.py:88-113 -- [bs,h,w,3] (x+0.5, y+0.5, 1)."""
    device = _device
    ys = (torch.arange(h, dtype=torch.float32, device=device) + 1).view(1, h, 1).expand(bs, h, w) - 0.5
    xs = (torch.arange(w, dtype=torch.float32, device=device) + 1).view(1, 1, w).expand(bs, h, w) - 0.5
    out = torch.stack([xs, ys, torch.ones(bs, h, w, dtype=torch.float32, device=device)], dim=3)
    out = out.as_subclass(_StandardGrid)
    out._lsi_standard_grid = (bs, h, w)
    return out


def is_standard_grid(t, bs, h, w):
    return getattr(t, '_lsi_standard_grid', None) == (bs, h, w)


def transform_pts(pts_coords_init, tform_mat):
    """helpers.py:116-137 -- pts [...,H,W,D] x mat [...,D,D]^T."""
    shp = pts_coords_init.shape
    flat = pts_coords_init.reshape(*tform_mat.shape[:-2], -1, tform_mat.shape[-1])
    return torch.matmul(flat, tform_mat.transpose(-1, -2)).reshape(shp)


def soft_z_buffering(layer_masks, layer_disps, depth_softmax_temp=1):
    """helpers.py:140-160."""
    eps = 1e-8
    depths = divide_safe(torch.ones_like(layer_disps), torch.relu(layer_disps))
    logp = torch.log(layer_masks + eps) - depths / depth_softmax_temp
    logp = logp - logp.amax(dim=0, keepdim=True)
    p = torch.exp(logp)
    return p / p.sum(dim=0, keepdim=True)


def enforce_bg_occupied(ldi_masks):
    """helpers.py:163-177 -- last layer's mask := 1."""
    n = ldi_masks.shape[0]
    if n == 1:
        return ldi_masks * 0 + 1
    return torch.cat([ldi_masks[:n - 1], ldi_masks[n - 1:] * 0 + 1], dim=0)


def zbuffer_weights(disps, scale=50):
    """helpers.py:180-193 -- exp((clip(d,0,1)-0.5)*scale)*[d>0]; accepts python scalars like the reference."""
    if not torch.is_tensor(disps):
        disps = torch.tensor(float(disps), dtype=torch.float32)
    pos = (disps > 0).to(disps.dtype)
    d = torch.clamp(disps, 0, 1)
    return torch.exp((d - 0.5) * scale) * pos
