"""Mirror of lsi/geometry/ldi.py (reference tree): the LDI renderer and the disparity smoothness loss.

`forward_splat` keeps the reference signature (ldi.py:71-83) but runs as fused sm_100a kernels behind the
C ABI (lsi_b200_forward_splat / _backward in include/lsi_b200.h): projection, z-buffer weights, the three
splats per layer, per-layer disparity normalisation, composition and image normalisation happen in one
pass per batch chunk instead of the reference's ~80*L scatter_nd+add pairs.
"""
import torch

from lsi import _b200
from lsi.nnutils import helpers as nn_helpers


def gradient(pred):
    """ldi.py:33-44 -- forward differences along W (dx) and H (dy) of [L,B,H,W,C]."""
    dy = pred[:, :, 1:, :, :] - pred[:, :, :-1, :, :]
    dx = pred[:, :, :, 1:, :] - pred[:, :, :, :-1, :]
    return dx, dy


class _DispSmoothness(torch.autograd.Function):
    @staticmethod
    def forward(ctx, disp):
        n_img, h, w = disp.shape[0] * disp.shape[1], disp.shape[2], disp.shape[3]
        out = torch.empty((), dtype=torch.float32, device=disp.device)
        _b200.call('lsi_b200_disp_smoothness_loss', _b200.ptr(disp), n_img, h, w, _b200.ptr(out),
                   _b200.ptr(_b200.partials(disp.device)), _b200.stream())
        ctx.save_for_backward(disp)
        return out

    @staticmethod
    def backward(ctx, g):
        disp, = ctx.saved_tensors
        g = g.contiguous().float()
        d = torch.empty_like(disp)
        _b200.call('lsi_b200_disp_smoothness_loss_backward', _b200.ptr(disp), disp.shape[0] * disp.shape[1],
                   disp.shape[2], disp.shape[3], _b200.ptr(g), _b200.ptr(d), _b200.stream())
        return d


def disp_smoothness_loss(pred_disp):
    """ldi.py:47-68 -- sum of the means of |d2/dx2|, |d2/dxdy|, |d2/dydx|, |d2/dy2|; pred_disp [L,B,H,W,1]."""
    disp = _b200.dev_f32(pred_disp, 'pred_disp')
    if disp.dim() != 5 or disp.shape[4] != 1:
        raise RuntimeError('lsi_b200: pred_disp must be [L,B,H,W,1], got %s' % (tuple(disp.shape),))
    return _DispSmoothness.apply(disp)


def _px_stride(t, name):
    """Accept dense [L,B,H,W,C] tensors and channel-slices of a packed [L,B,H,W,C'] tensor (nets.py:204) without
    copying: returns (tensor, pixel stride in elements)."""
    L, B, H, W, C = t.shape
    s = t.stride()
    ps = s[3]
    if s[4] == 1 and ps >= C and s[2] == W * ps and s[1] == H * W * ps and s[0] == B * H * W * ps:
        return t, ps
    return t.contiguous(), C


class _PoseClassCache(object):
    """Which camera sets are entirely in the rectified pose class (n == 1, y' independent of x and d: stereo pairs) -- the class
    the row-gather kernel renders on its own (csrc/render_rowgather.cuh).  The classification itself happens on the device in
    every call (proj_matrix_kernel writes one flag per image next to the matrices); this cache only remembers the flags of camera
    tensors it has seen before, so that a later call with the SAME tensor objects at the SAME versions can tell the library
    (variant 5) not to launch the fallback kernels for other pose classes at all.  Nothing ever synchronises: the flags of a
    first-seen camera set are copied back asynchronously and used once the copy has completed.  Entries are tied to the tensor
    objects by weak references (an address alone could be reused by a different tensor) and to their version counters (in-place
    updates); tensors that keep changing are marked volatile and not looked at again."""

    def __init__(self):
        self.entries = {}

    def lookup(self, cams):
        import weakref
        key = tuple(id(c) for c in cams)
        vers = tuple(c._version for c in cams)
        e = self.entries.get(key)
        if e is not None and not all(r() is c for r, c in zip(e['refs'], cams)):
            e = None                                   # an id was recycled by another object
        if e is None:
            if len(self.entries) > 64:
                self.entries.clear()
            e = dict(refs=tuple(weakref.ref(c) for c in cams), vers=vers, state=None, pending=None, volatile=False)
            self.entries[key] = e
            return e, None
        if e['vers'] != vers:                          # modified in place since: classify again, once
            e['volatile'] = e['state'] is not None or e['pending'] is not None or e['volatile']
            e.update(vers=vers, state=None, pending=None)
            return e, None
        self._resolve(e)
        return e, e['state']

    @staticmethod
    def _resolve(e):
        if e['pending'] is not None and e['pending'][1].query():
            flags = e['pending'][0]
            e['state'] = 'all' if bool(flags.all().item()) else ('none' if not bool(flags.any().item()) else 'mixed')
            e['pending'] = None

    def lookup_key(self, key):
        """The same memo for callers that identify a camera set by content (e.g. a hash of the host copies they upload every step)
        instead of by device tensor identity."""
        e = self.entries.get(key)
        if e is None:
            if len(self.entries) > 64:
                self.entries.clear()
            e = dict(refs=(), vers=(), state=None, pending=None, volatile=False)
            self.entries[key] = e
            return e, None
        self._resolve(e)
        return e, e['state']

    @staticmethod
    def record(e, ws, batch):
        """Start the asynchronous read-back of this call's per-image class flags (they follow the B 4x4 matrices in the workspace)."""
        if e is None or e['volatile'] or e['state'] is not None or e['pending'] is not None:
            return
        host = torch.empty(batch, dtype=torch.int32, pin_memory=True)
        host.copy_(ws[batch * 64: batch * 68].view(torch.int32), non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        e['pending'] = (host, ev)


_POSE_CACHE = _PoseClassCache()


class _ForwardSplat(torch.autograd.Function):
    @staticmethod
    def forward(ctx, tex, mask, disp, pc, k_s, k_t, rot, t, focal, cfg, pose_entry=None):
        L, B, H, W, _ = tex.shape
        h_t, w_t, ds, compose, want_disp, bg, max_disp, scale, variant = cfg
        tex_v, tex_s = _px_stride(tex, 'tex')
        disp_v, disp_s = _px_stride(disp, 'disp')
        mask_v, mask_s = (None, 1) if mask is None else _px_stride(mask, 'mask')
        desc = _b200.SplatDesc(L, B, H, W, h_t, w_t, ds, bg, max_disp, scale, int(compose), int(want_disp),
                               tex_s, disp_s, mask_s, variant)
        nl = 1 if compose else L
        dev = tex.device
        img = torch.empty(nl, B, h_t, w_t, 3, dtype=torch.float32, device=dev)
        wts = torch.empty(nl, B, h_t, w_t, 1, dtype=torch.float32, device=dev)
        dsp = torch.empty(nl, B, h_t, w_t, 1, dtype=torch.float32, device=dev) if want_disp else None
        needs_grad = any(ctx.needs_input_grad[:3])
        layer_acc = torch.empty(L, B, h_t, w_t, 2, dtype=torch.float32, device=dev) if (want_disp and needs_grad) else None
        ws_bytes = _b200.lib().lsi_b200_forward_splat_workspace_bytes(desc)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        _b200.call('lsi_b200_forward_splat', desc, _b200.ptr(tex_v), _b200.ptr(mask_v), _b200.ptr(disp_v),
                   _b200.ptr(pc), _b200.ptr(k_s), _b200.ptr(k_t), _b200.ptr(rot), _b200.ptr(t), _b200.ptr(focal),
                   _b200.ptr(img), _b200.ptr(wts), _b200.ptr(dsp), _b200.ptr(layer_acc), _b200.ptr(ws), ws_bytes,
                   _b200.stream())
        _PoseClassCache.record(pose_entry, ws, B)
        if needs_grad:
            ctx.save_for_backward(tex, mask, disp, pc, k_s, k_t, rot, t, focal, img, wts, layer_acc)
            ctx.cfg = cfg
        if want_disp:
            return img, wts, dsp
        return img, wts

    @staticmethod
    def backward(ctx, g_img, g_wts, g_disp=None):
        tex, mask, disp, pc, k_s, k_t, rot, t, focal, img, wts, layer_acc = ctx.saved_tensors
        L, B, H, W, _ = tex.shape
        h_t, w_t, ds, compose, want_disp, bg, max_disp, scale, variant = ctx.cfg
        tex_c, disp_c = tex.contiguous(), disp.contiguous()
        mask_c = None if mask is None else mask.contiguous()
        desc = _b200.SplatDesc(L, B, H, W, h_t, w_t, ds, bg, max_disp, scale, int(compose), int(want_disp), 3, 1, 1, variant)
        dev = tex.device
        g_img = None if g_img is None else g_img.contiguous()
        g_wts = None if g_wts is None else g_wts.contiguous()
        g_disp = None if g_disp is None else g_disp.contiguous()
        d_tex = torch.empty(L, B, H, W, 3, dtype=torch.float32, device=dev)
        d_disp = torch.empty(L, B, H, W, 1, dtype=torch.float32, device=dev)
        d_mask = torch.empty(L, B, H, W, 1, dtype=torch.float32, device=dev) if (mask is not None and ctx.needs_input_grad[1]) else None
        ws_bytes = _b200.lib().lsi_b200_forward_splat_backward_workspace_bytes(desc)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        _b200.call('lsi_b200_forward_splat_backward', desc, _b200.ptr(tex_c), _b200.ptr(mask_c), _b200.ptr(disp_c),
                   _b200.ptr(pc), _b200.ptr(k_s), _b200.ptr(k_t), _b200.ptr(rot), _b200.ptr(t), _b200.ptr(focal),
                   _b200.ptr(img), _b200.ptr(wts), _b200.ptr(layer_acc), _b200.ptr(g_img), _b200.ptr(g_wts),
                   _b200.ptr(g_disp), _b200.ptr(d_tex), _b200.ptr(d_mask), _b200.ptr(d_disp), _b200.ptr(ws), ws_bytes,
                   _b200.stream())
        return d_tex, d_mask, d_disp, None, None, None, None, None, None, None, None


def forward_splat(ldi_src, pixel_coords_src, k_s, k_t, rot, t, focal_disps=None, compose_layers=True,
                  compute_trg_disp=False, trg_downsampling=1, bg_layer_disp=0, max_disp=1, zbuf_scale=10,
                  _variant=0, _pose_key=None):
    """ldi.py:71-182.  Forward splat the source LDI into the target camera.

    Args (as the reference): ldi_src = (imgs [L,B,H,W,3], masks [L,B,H,W,1], disps [L,B,H,W,1]);
      pixel_coords_src [B,H,W,3]; k_s, k_t, rot [B,3,3]; t [B,3,1]; focal_disps optional [B,1,1,1].
    Returns: trg_img [nl_out,B,Ht,Wt,3], trg_wts [nl_out,B,Ht,Wt,1] (, trg_disp [nl_out,B,Ht,Wt,1]).
    Differentiable w.r.t. imgs, masks, disps.  `masks` tagged all-ones by nets.ldi_predictor (nets.py:205) and
    the standard `helpers.pixel_coords` grid are recognised and never read from memory.
    """
    imgs, masks, disps = ldi_src
    tex = _b200.dev_f32(imgs, 'ldi_src[0]', contiguous=False)
    disp = _b200.dev_f32(disps, 'ldi_src[2]', contiguous=False)
    if tex.dim() != 5 or tex.shape[4] != 3:
        raise RuntimeError('lsi_b200: ldi imgs must be [L,B,H,W,3], got %s' % (tuple(tex.shape),))
    L, B, H, W, _ = tex.shape
    if tuple(disp.shape) != (L, B, H, W, 1):
        raise RuntimeError('lsi_b200: ldi disps must be [L,B,H,W,1] = %s, got %s' % ((L, B, H, W, 1), tuple(disp.shape)))
    mask = None
    if masks is not None and not getattr(masks, '_lsi_all_ones', False):
        mask = _b200.dev_f32(masks, 'ldi_src[1]', contiguous=False)
        if tuple(mask.shape) != (L, B, H, W, 1):
            raise RuntimeError('lsi_b200: ldi masks must be %s, got %s' % ((L, B, H, W, 1), tuple(mask.shape)))
    pc = None
    if not nn_helpers.is_standard_grid(pixel_coords_src, B, H, W):
        pc = _b200.dev_f32(pixel_coords_src.as_subclass(torch.Tensor), 'pixel_coords_src')
        if tuple(pc.shape) != (B, H, W, 3):
            raise RuntimeError('lsi_b200: pixel_coords_src must be %s, got %s' % ((B, H, W, 3), tuple(pc.shape)))
    from lsi.geometry import projection
    pose_entry, all_rect = None, None
    if _variant == 0 and _pose_key is not None:
        pose_entry, all_rect = _POSE_CACHE.lookup_key(('key', _pose_key))
    elif _variant == 0 and all(isinstance(c, torch.Tensor) and c.is_cuda for c in (k_s, k_t, rot, t)):
        pose_entry, all_rect = _POSE_CACHE.lookup((k_s, k_t, rot, t))
    k_s, k_t, rot, t = projection._cam(k_s, k_t, rot, t)
    if k_s.shape[0] != B:
        raise RuntimeError('lsi_b200: camera batch %d != LDI batch %d' % (k_s.shape[0], B))
    focal = None
    if focal_disps is not None:
        focal = _b200.dev_f32(focal_disps, 'focal_disps').reshape(-1)
        if focal.numel() != B:
            raise RuntimeError('lsi_b200: focal_disps must have B=%d elements' % B)
    h_t, w_t = H * trg_downsampling, W * trg_downsampling       # ldi.py:113-114 (floats accepted if integral)
    if h_t != int(h_t) or w_t != int(w_t) or int(h_t) < 1 or int(w_t) < 1:
        raise RuntimeError('lsi_b200: trg_downsampling=%r does not give an integral target size for %dx%d'
                           % (trg_downsampling, H, W))
    cfg = (int(h_t), int(w_t), float(trg_downsampling), bool(compose_layers), bool(compute_trg_disp),
           float(bg_layer_disp), float(max_disp), float(zbuf_scale), {'all': 5, 'none': 6}.get(all_rect, int(_variant)))
    return _ForwardSplat.apply(tex, mask, disp, pc, k_s, k_t, rot, t, focal, cfg, pose_entry)
