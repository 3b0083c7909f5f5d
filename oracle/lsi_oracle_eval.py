"""CPU restatement of the evaluation metrics of the reference (TEST INFRASTRUCTURE: only tests/ may import this).

Follows `ldi_pred_eval.Tester.define_metrics` (ldi_pred_eval.py:297-548) op for op on torch-CPU, on top of the oracle's
`forward_splat` / `disocclusion_mask` (oracle/lsi_oracle.py).  Pinned by tests/golden/eval_{synth,kitti}.npz, produced by
running the reference's own, unmodified define_metrics over the TF-1 shim (oracle/gen_golden_eval.py).

One quirk is restated on purpose: at ldi_pred_eval.py:402-405 the `> 0.95` threshold is assigned to the FULL-resolution
mask variable, so the mask actually used (`to_recons_disocc_mask_ds`) is the un-thresholded AREA average -- fractional weights.
"""
import torch

from oracle import lsi_oracle as O


def _area(x, h_t, w_t):
    """[TF1.4] tf.image.resize_images(AREA) by an integer factor = box mean."""
    b, h, w, c = x.shape
    return x.reshape(b, h_t, h // h_t, w_t, w // w_t, c).mean(dim=(2, 4))


def define_metrics(opts, ldi_src, ldi_trg, imgs_src, imgs_trg, k_s, k_t, rot, trans, src_gt_disp=None, trg_gt_disp=None,
                   src_gt_disp_bg=None, trg_gt_disp_bg=None, src_gt_tex_bg=None, trg_gt_tex_bg=None):
    """-> (metrics, metrics_norm) dicts of 0-d tensors; ldi_* = (tex [L,B,H,W,3], mask, disp [L,B,H,W,1])."""
    B, H, W, _ = imgs_src.shape
    pc = O.pixel_coords(B, H, W, dtype=imgs_src.dtype)
    synthetic = opts.dataset == 'synthetic'
    disocc = synthetic or (opts.dataset == 'kitti' and opts.kitti_dl_disparities)            # :304-309
    inv_rot = rot.transpose(-1, -2)                                                           # :180-181
    inv_trans = -torch.matmul(inv_rot, trans)
    if synthetic:                                                                             # :152-161
        dm_src = O.disocclusion_mask(src_gt_disp, trg_gt_disp, pc, O.forward_projection_matrix(k_s, k_t, rot, trans))
        dm_trg = O.disocclusion_mask(trg_gt_disp, src_gt_disp, pc, O.inverse_projection_matrix(k_s, k_t, rot, trans))
    elif disocc:                                                                              # :172-173
        dm_src, dm_trg = (src_gt_disp == 0).to(imgs_src.dtype), (trg_gt_disp == 0).to(imgs_src.dtype)
    z = imgs_src.new_zeros(())
    acc = dict(compose=z.clone(), valid=z.clone(), compose_d=z.clone(), valid_d=z.clone(), depth=z.clone(), depth_d=z.clone())
    kw = dict(compose_layers=True, compute_trg_disp=True, trg_downsampling=opts.trg_splat_downsampling,
              zbuf_scale=opts.zbuf_scale, bg_layer_disp=opts.bg_layer_disp, max_disp=opts.max_disp)
    for name in ('trg', 'src'):                                                               # :335-470
        if name == 'trg':
            img, gt, dm = imgs_trg, trg_gt_disp, (dm_trg if disocc else None)
            recons, _, rdisp = O.forward_splat(ldi_src, pc, k_s, k_t, rot, trans, **kw)
        else:
            img, gt, dm = imgs_src, src_gt_disp, (dm_src if disocc else None)
            recons, _, rdisp = O.forward_splat(ldi_trg, pc, k_t, k_s, inv_rot, inv_trans, **kw)
        valid = (gt > opts.bg_layer_disp).to(img.dtype) if synthetic else torch.ones(B, H, W, 1, dtype=img.dtype)
        h_t, w_t = recons.shape[2], recons.shape[3]
        img_ds = _area(img, h_t, w_t)
        valid = (_area(valid, h_t, w_t) > 0.95).to(img.dtype)[..., 0]                          # :395-400
        pw = (img_ds - recons).abs().mean(dim=4).min(dim=0).values                            # :417-421
        x_min, y_min = int(round(w_t * opts.splat_bdry_ignore)), int(round(h_t * opts.splat_bdry_ignore))
        centre = torch.zeros(B, h_t, w_t, dtype=img.dtype)
        centre[:, y_min:h_t - y_min, x_min:w_t - x_min] = 1
        centre = centre * valid
        pw = pw * centre
        acc['compose'] += pw.sum(); acc['valid'] += centre.sum()
        if disocc:
            dm_ds = _area(dm, h_t, w_t)[..., 0]                                                # fractional: see module docstring
            acc['compose_d'] += (pw * dm_ds).sum(); acc['valid_d'] += (centre * dm_ds).sum()
        if synthetic:
            pd = (_area(gt, h_t, w_t) - rdisp).abs().mean(dim=4).min(dim=0).values * centre    # :407-415,437-438
            acc['depth'] += pd.sum()
            acc['depth_d'] += (pd * dm_ds).sum()
    metrics, norm = {'compose_loss': acc['compose']}, {'compose_loss': acc['valid']}
    if synthetic:                                                                             # :477-533
        nl = opts.n_layers
        for key, layer, gt_tex, gt_disp, mask_gt in (
                ('bg', nl - 1, (src_gt_tex_bg, trg_gt_tex_bg), (src_gt_disp_bg, trg_gt_disp_bg), (src_gt_disp_bg, trg_gt_disp_bg)),
                ('fg', 0, (imgs_src, imgs_trg), (src_gt_disp, trg_gt_disp), (opts.bg_layer_disp, opts.bg_layer_disp))):
            vs = (src_gt_disp > mask_gt[0]).to(imgs_src.dtype)
            vt = (trg_gt_disp > mask_gt[1]).to(imgs_src.dtype)
            tex = ((ldi_src[0][layer] - gt_tex[0]).abs() * vs).sum() / 3 + ((ldi_trg[0][layer] - gt_tex[1]).abs() * vt).sum() / 3
            dsp = ((ldi_src[2][layer] - gt_disp[0]).abs() * vs).sum() + ((ldi_trg[2][layer] - gt_disp[1]).abs() * vt).sum()
            cnt = (vt + vs).sum()
            metrics[key + '_tex_error'], metrics[key + '_disp_error'] = tex, dsp
            norm[key + '_tex_error'] = norm[key + '_disp_error'] = cnt
    if disocc:
        metrics['compose_loss_disocc'], norm['compose_loss_disocc'] = acc['compose_d'], acc['valid_d']
    if synthetic:
        metrics['depth_loss'], norm['depth_loss'] = acc['depth'], acc['valid']
        metrics['depth_loss_disocc'], norm['depth_loss_disocc'] = acc['depth_d'], acc['valid_d']
    return metrics, norm
