"""Mirror of lsi/loss/loss.py (reference tree) plus the photometric splat loss that the reference keeps inline in
ldi_enc_dec.py:337-357 (`splat_photo_loss`) and the loss mix of ldi_enc_dec.py:265-410 (`view_synthesis_loss`)."""
import torch

from lsi import _b200


class _DecreasingDisp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, disp):
        L = disp.shape[0]
        n = disp.numel() // L
        out = torch.empty((), dtype=torch.float32, device=disp.device)
        _b200.call('lsi_b200_decreasing_disp_loss', _b200.ptr(disp), L, n, _b200.ptr(out),
                   _b200.ptr(_b200.partials(disp.device)), _b200.stream())
        ctx.save_for_backward(disp)
        return out

    @staticmethod
    def backward(ctx, g):
        disp, = ctx.saved_tensors
        L = disp.shape[0]
        d = torch.empty_like(disp)
        _b200.call('lsi_b200_decreasing_disp_loss_backward', _b200.ptr(disp), L, disp.numel() // L,
                   _b200.ptr(g.contiguous().float()), _b200.ptr(d), _b200.stream())
        return d


def event_prob(layer_masks):
    """loss.py:27-45 -- per-pixel layer assignment probabilities by ordered multiplication: p_l = m_l * prod_{k<l} (1 - m_k) with
    the masks clipped to [1e-6, 1 - 1e-6]; returns (layer_probs [L,...,1], escape_probs [1,...,1]).  Unused by either reference
    script (kept for API parity; plain elementwise torch ops on the caller's device)."""
    eps = 1e-6
    m = torch.clamp(layer_masks, eps, 1 - eps)
    log_inv = torch.log(1 - m)
    layer_probs = torch.exp(torch.cumsum(log_inv, dim=0) - log_inv + torch.log(m))
    return layer_probs, 1 - layer_probs.sum(dim=0, keepdim=True)


def decreasing_disp_loss(layer_disps):
    """loss.py:48-63 -- mean relu(d[l+1] - stop_gradient(d[l])); 0 for a single layer."""
    return _DecreasingDisp.apply(_b200.dev_f32(layer_disps, 'layer_disps'))


class _ZbufComposition(torch.autograd.Function):
    @staticmethod
    def forward(ctx, tex, mask, disp, trg, bg, max_disp, scale):
        L = tex.shape[0]
        n = tex.numel() // (3 * L)
        out = torch.empty((), dtype=torch.float32, device=tex.device)
        _b200.call('lsi_b200_zbuf_composition_loss', _b200.ptr(tex), _b200.ptr(mask), _b200.ptr(disp), _b200.ptr(trg),
                   L, n, bg, max_disp, scale, _b200.ptr(out), _b200.ptr(_b200.partials(tex.device)), _b200.stream())
        ctx.save_for_backward(tex, mask, disp, trg)
        ctx.cfg = (L, n, bg, max_disp, scale)
        return out

    @staticmethod
    def backward(ctx, g):
        tex, mask, disp, trg = ctx.saved_tensors
        L, n, bg, max_disp, scale = ctx.cfg
        d_tex, d_disp = torch.empty_like(tex), torch.empty_like(disp)
        d_mask = torch.empty_like(mask) if (mask is not None and ctx.needs_input_grad[1]) else None
        _b200.call('lsi_b200_zbuf_composition_loss_backward', _b200.ptr(tex), _b200.ptr(mask), _b200.ptr(disp),
                   _b200.ptr(trg), L, n, bg, max_disp, scale, _b200.ptr(g.contiguous().float()), _b200.ptr(d_tex),
                   _b200.ptr(d_mask), _b200.ptr(d_disp), _b200.stream())
        return d_tex, d_mask, d_disp, None, None, None, None


def zbuffer_composition_loss(layer_imgs, layer_masks, layer_disps, trg_imgs, bg_layer_disp=0, max_disp=1,
                             zbuf_scale=10):
    """loss.py:66-115.  layer_imgs [L,...,C=3], layer_masks/layer_disps [L,...,1], trg_imgs [...,3] -> scalar."""
    tex = _b200.dev_f32(layer_imgs, 'layer_imgs')
    disp = _b200.dev_f32(layer_disps, 'layer_disps')
    trg = _b200.dev_f32(trg_imgs, 'trg_imgs')
    if tex.shape[-1] != 3 or tuple(trg.shape) != tuple(tex.shape[1:]) or tuple(disp.shape) != tuple(tex.shape[:-1]) + (1,):
        raise RuntimeError('lsi_b200: zbuffer_composition_loss shape mismatch: imgs %s disps %s trg %s'
                           % (tuple(tex.shape), tuple(disp.shape), tuple(trg.shape)))
    mask = None
    if layer_masks is not None and not getattr(layer_masks, '_lsi_all_ones', False):
        mask = _b200.dev_f32(layer_masks, 'layer_masks')
    return _ZbufComposition.apply(tex, mask, disp, trg, float(bg_layer_disp), float(max_disp), float(zbuf_scale))


class _SplatPhoto(torch.autograd.Function):
    @staticmethod
    def forward(ctx, render, gt, bdry):
        nl, b, h_t, w_t, _ = render.shape
        _, h, w, _ = gt.shape
        out = torch.empty((), dtype=torch.float32, device=render.device)
        _b200.call('lsi_b200_photo_loss', _b200.ptr(render), _b200.ptr(gt), nl, b, h, w, h_t, w_t, bdry, _b200.ptr(out),
                   _b200.ptr(_b200.partials(render.device)), _b200.stream())
        ctx.save_for_backward(render, gt)
        ctx.cfg = (nl, b, h, w, h_t, w_t, bdry)
        return out

    @staticmethod
    def backward(ctx, g):
        render, gt = ctx.saved_tensors
        d = torch.empty_like(render)
        _b200.call('lsi_b200_photo_loss_backward', _b200.ptr(render), _b200.ptr(gt), *ctx.cfg,
                   _b200.ptr(g.contiguous().float()), _b200.ptr(d), _b200.stream())
        return d, None, None


def splat_photo_loss(recons_splat, to_recons_img, splat_bdry_ignore):
    """ldi_enc_dec.py:337-357 -- AREA-downsample the GT image to the splat size, mean_c |GT - render|, min over the
    layer axis, crop round(size*splat_bdry_ignore) pixels at every border, mean."""
    render = _b200.dev_f32(recons_splat, 'recons_splat')
    gt = _b200.dev_f32(to_recons_img, 'to_recons_img')
    if render.dim() != 5 or render.shape[4] != 3 or gt.dim() != 4 or gt.shape[3] != 3 or gt.shape[0] != render.shape[1]:
        raise RuntimeError('lsi_b200: splat_photo_loss shape mismatch: render %s gt %s' % (tuple(render.shape), tuple(gt.shape)))
    return _SplatPhoto.apply(render, gt, float(splat_bdry_ignore))


def view_synthesis_loss(ldi_src, ldi_trg, imgs_src, imgs_trg, pixel_coords, k_s, k_t, rot_mat, trans_mat, opts):
    """ldi_enc_dec.py:265-410 (Trainer.define_loss_graph) as a function: self-consistency + 4 forward splats
    ({indep, compose} x {src->trg, trg->src}) + smoothness + layer ordering, mixed with the reference's weights.
    `opts` needs the reference flag names (self_cons_wt, indep_splat_wt, compose_splat_wt, splat_bdry_ignore,
    zbuf_scale, trg_splat_downsampling, disp_smoothness_wt, incr_depth_wt, bg_layer_disp, max_disp, l0_self_cons).
    Returns (total_loss, dict of the component losses)."""
    from lsi.geometry import ldi as ldi_utils
    inv_rot = rot_mat.transpose(-1, -2).contiguous()            # ldi_enc_dec.py:193-194
    inv_trans = -torch.matmul(inv_rot, trans_mat)
    kw = dict(zbuf_scale=opts.zbuf_scale, bg_layer_disp=opts.bg_layer_disp, max_disp=opts.max_disp)
    if opts.l0_self_cons:
        sc = (imgs_src - ldi_src[0][0]).abs().mean() + (imgs_trg - ldi_trg[0][0]).abs().mean()
    else:
        sc = (zbuffer_composition_loss(ldi_src[0], ldi_src[1], ldi_src[2], imgs_src, **kw)
              + zbuffer_composition_loss(ldi_trg[0], ldi_trg[1], ldi_trg[2], imgs_trg, **kw))
    parts = {'indep': 0, 'compose': 0}
    for name, compose in (('indep', False), ('compose', True)):
        r_trg, _ = ldi_utils.forward_splat(ldi_src, pixel_coords, k_s, k_t, rot_mat, trans_mat, compose_layers=compose,
                                           trg_downsampling=opts.trg_splat_downsampling, **kw)
        r_src, _ = ldi_utils.forward_splat(ldi_trg, pixel_coords, k_t, k_s, inv_rot, inv_trans, compose_layers=compose,
                                           trg_downsampling=opts.trg_splat_downsampling, **kw)
        parts[name] = (splat_photo_loss(r_trg, imgs_trg, opts.splat_bdry_ignore)
                       + splat_photo_loss(r_src, imgs_src, opts.splat_bdry_ignore))
    smooth = ldi_utils.disp_smoothness_loss(ldi_src[2]) + ldi_utils.disp_smoothness_loss(ldi_trg[2])
    incr = decreasing_disp_loss(ldi_src[2]) + decreasing_disp_loss(ldi_trg[2])
    total = 0.0
    if opts.self_cons_wt > 0:
        total = total + opts.self_cons_wt * sc
    if opts.compose_splat_wt > 0:
        total = total + opts.compose_splat_wt * parts['compose']
    if opts.indep_splat_wt > 0:
        total = total + opts.indep_splat_wt * parts['indep']
    if opts.incr_depth_wt > 0:
        total = total + (opts.incr_depth_wt / opts.max_disp) * incr
    if opts.disp_smoothness_wt > 0:
        total = total + (opts.disp_smoothness_wt / (opts.max_disp * opts.max_disp)) * smooth
    return total, dict(self_cons=sc, indep_splat=parts['indep'], compose_splat=parts['compose'], incr_depth=incr,
                       disp_smoothness=smooth)
