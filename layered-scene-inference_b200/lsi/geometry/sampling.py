"""Mirror of lsi/geometry/sampling.py (reference tree): bilinear forward splat and bilinear gather sampler."""
import torch

from lsi import _b200


class _Splat(torch.autograd.Function):
    @staticmethod
    def forward(ctx, src, coords, init):
        b, h_s, w_s, c = src.shape
        _, h_t, w_t, _ = init.shape
        out = torch.empty_like(init)
        _b200.call('lsi_b200_splat', _b200.ptr(src), _b200.ptr(coords), _b200.ptr(init), _b200.ptr(out),
                   b, h_s, w_s, h_t, w_t, c, _b200.stream())
        ctx.save_for_backward(src, coords)
        ctx.dims = (b, h_s, w_s, h_t, w_t, c)
        return out

    @staticmethod
    def backward(ctx, g):
        src, coords = ctx.saved_tensors
        g = g.contiguous()
        d_src, d_coords = torch.empty_like(src), torch.empty_like(coords)
        _b200.call('lsi_b200_splat_backward', _b200.ptr(src), _b200.ptr(coords), _b200.ptr(g), _b200.ptr(d_src),
                   _b200.ptr(d_coords), *ctx.dims, _b200.stream())
        return d_src, d_coords, g


def splat(src_image, tgt_coords, init_trg_image):
    """sampling.py:171-254.  src_image [B,Hs,Ws,C], tgt_coords [B,Hs,Ws,2], init_trg_image [B,Ht,Wt,C] -> new
    target image; functional (the init is not modified)."""
    src = _b200.dev_f32(src_image, 'src_image')
    coords = _b200.dev_f32(tgt_coords, 'tgt_coords')
    init = _b200.dev_f32(init_trg_image, 'init_trg_image')
    if src.dim() != 4 or coords.shape != src.shape[:3] + (2,) or init.dim() != 4 or init.shape[0] != src.shape[0] \
            or init.shape[3] != src.shape[3]:
        raise RuntimeError('lsi_b200: splat shape mismatch: src %s coords %s init %s'
                           % (tuple(src.shape), tuple(coords.shape), tuple(init.shape)))
    return _Splat.apply(src, coords, init)


class _Bilinear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, imgs, coords):
        b, h_s, w_s, c = imgs.shape
        _, h_t, w_t, _ = coords.shape
        out = torch.empty(b, h_t, w_t, c, dtype=torch.float32, device=imgs.device)
        _b200.call('lsi_b200_bilinear', _b200.ptr(imgs), _b200.ptr(coords), _b200.ptr(out), b, h_s, w_s, h_t, w_t, c,
                   _b200.stream())
        ctx.save_for_backward(imgs, coords)
        ctx.dims = (b, h_s, w_s, h_t, w_t, c)
        return out

    @staticmethod
    def backward(ctx, g):
        imgs, coords = ctx.saved_tensors
        g = g.contiguous()
        d_imgs, d_coords = torch.zeros_like(imgs), torch.empty_like(coords)
        _b200.call('lsi_b200_bilinear_backward', _b200.ptr(imgs), _b200.ptr(coords), _b200.ptr(g), _b200.ptr(d_imgs),
                   _b200.ptr(d_coords), *ctx.dims, _b200.stream())
        return d_imgs, d_coords


def bilinear(imgs, coords, compose=True):
    """sampling.py:41-132.  imgs [B,Hs,Ws,C], coords [B,Ht,Wt,2] -> [B,Ht,Wt,C]; zero outside the image.
    compose=False (sampling.py:117-131, the reference's data generator): returns (out_ims, out_wts), two lists of four
    tensors -- the corner samples masked by validity [B,Ht,Wt,C] and their raw bilinear weights [B,Ht,Wt,1] -- in the order
    (x0,y0), (x0,y1), (x1,y0), (x1,y1); forward only (no gradient, as the generator needs none)."""
    imgs = _b200.dev_f32(imgs, 'imgs')
    coords = _b200.dev_f32(coords, 'coords')
    if imgs.dim() != 4 or coords.dim() != 4 or coords.shape[0] != imgs.shape[0] or coords.shape[3] != 2:
        raise RuntimeError('lsi_b200: bilinear shape mismatch: imgs %s coords %s' % (tuple(imgs.shape), tuple(coords.shape)))
    if not compose:
        b, h_s, w_s, c = imgs.shape
        _, h_t, w_t, _ = coords.shape
        ims = torch.empty(4, b, h_t, w_t, c, dtype=torch.float32, device=imgs.device)
        wts = torch.empty(4, b, h_t, w_t, 1, dtype=torch.float32, device=imgs.device)
        _b200.call('lsi_b200_bilinear_corners', _b200.ptr(imgs.detach()), _b200.ptr(coords.detach()), _b200.ptr(ims), _b200.ptr(wts),
                   b, h_s, w_s, h_t, w_t, c, _b200.stream())
        return list(ims.unbind(0)), list(wts.unbind(0))
    return _Bilinear.apply(imgs, coords)


def bilinear_wrapper(imgs, coords, compose=True):
    """sampling.py:135-168 -- arbitrary leading dims."""
    init_dims = list(imgs.shape[:-3])
    out = bilinear(imgs.reshape([-1] + list(imgs.shape[-3:])), coords.reshape([-1] + list(coords.shape[-3:])), compose)
    if not compose:                                                   # sampling.py:160-166: reshape every image / weight of the two lists
        return tuple([o.reshape(init_dims + list(o.shape[-3:])) for o in lst] for lst in out)
    return out.reshape(init_dims + list(out.shape[-3:]))


def scatter_add_tensor(init, indices, updates):
    """sampling.py:257-284 -- init + scatter_nd(indices, updates); indices [N,1] into dim 0."""
    return init.index_add(0, indices.reshape(-1).long(), updates)


def batch_scatter_add_tensor(init, indices, updates):
    """sampling.py:287-313 -- per-batch scatter-add; init [B,P], indices/updates [B,U]."""
    return init.scatter_add(1, indices.long(), updates)
