"""Mirror of the evaluation metrics of the reference: `ldi_pred_eval.Tester.define_metrics` (ldi_pred_eval.py:297-548) and the
sum / sum-of-normalisers accumulation of `test_utils.Tester.test` (test_utils.py:225-262).

The renders come from the CUDA path (`lsi.geometry.ldi.forward_splat(compute_trg_disp=True)`,
`lsi.geometry.projection.disocclusion_mask`); what is left -- AREA down-sampling of the ground truth, masks, border crop and the
masked sums over a few hundred KB per batch -- is evaluation-time glue written with torch ops on whatever device the renders
live on (`metrics_from_renders`; not a hot path, and deliberately not a fallback of any kernel).

Restated quirk (ldi_pred_eval.py:402-405): the `> 0.95` threshold lands on the full-resolution mask variable, so the
disocclusion weights actually used are the un-thresholded AREA averages.
"""
import numpy as np
import math

import torch

from lsi.geometry import ldi as ldi_utils
from lsi.geometry import projection
from lsi.nnutils import helpers as nn_helpers


def _area(x, h_t, w_t):
    """tf.image.resize_images(..., AREA) by an integer factor (the only case the reference uses) = box mean."""
    b, h, w, c = x.shape
    if h % h_t or w % w_t:
        raise RuntimeError('lsi_b200: AREA resize needs an integer factor, got %dx%d -> %dx%d' % (h, w, h_t, w_t))
    return x.reshape(b, h_t, h // h_t, w_t, w // w_t, c).mean(dim=(2, 4))


def metrics_from_renders(opts, renders, imgs, gt_disps=None, disocc_masks=None):
    """The per-direction loop of define_metrics (ldi_pred_eval.py:335-470) given the renders.

    renders: {'trg': (recons [1,B,Ht,Wt,3], recons_disp [1,B,Ht,Wt,1]), 'src': (...)}; imgs: {'trg': imgs_trg, 'src': imgs_src};
    gt_disps / disocc_masks: same keys, [B,H,W,1] (None where the dataset has none).  -> dict of six 0-d sums."""
    synthetic = opts.dataset == 'synthetic'
    any_img = imgs['trg']
    z = any_img.new_zeros(())
    acc = dict(compose=z.clone(), valid=z.clone(), compose_d=z.clone(), valid_d=z.clone(), depth=z.clone(), depth_d=z.clone())
    for name in ('trg', 'src'):
        img = imgs[name]
        recons, rdisp = renders[name]
        B, H, W, _ = img.shape
        h_t, w_t = recons.shape[2], recons.shape[3]
        if synthetic:
            valid = (gt_disps[name] > opts.bg_layer_disp).to(img.dtype)
        else:
            valid = torch.ones(B, H, W, 1, dtype=img.dtype, device=img.device)
        valid = (_area(valid, h_t, w_t) > 0.95).to(img.dtype)[..., 0]                  # ignore pixels that might have aliasing
        pw = (_area(img, h_t, w_t) - recons).abs().mean(dim=4).min(dim=0).values
        # Python-2 round() (the reference's interpreter) rounds halves away from zero, as does the CUDA photo-loss kernel's floor(x + 0.5);
        # Python-3 round() would round halves to even
        x_min, y_min = int(math.floor(w_t * opts.splat_bdry_ignore + 0.5)), int(math.floor(h_t * opts.splat_bdry_ignore + 0.5))
        centre = torch.zeros(B, h_t, w_t, dtype=img.dtype, device=img.device)
        centre[:, y_min:h_t - y_min, x_min:w_t - x_min] = 1
        centre = centre * valid
        pw = pw * centre
        acc['compose'] += pw.sum(); acc['valid'] += centre.sum()
        dm_ds = None
        if disocc_masks is not None:
            dm_ds = _area(disocc_masks[name].to(img.dtype), h_t, w_t)[..., 0]
            acc['compose_d'] += (pw * dm_ds).sum(); acc['valid_d'] += (centre * dm_ds).sum()
        if synthetic:
            pd = (_area(gt_disps[name], h_t, w_t) - rdisp).abs().mean(dim=4).min(dim=0).values * centre
            acc['depth'] += pd.sum()
            if dm_ds is not None:
                acc['depth_d'] += (pd * dm_ds).sum()
    return acc


def layer_errors(opts, ldi_src, ldi_trg, imgs_src, imgs_trg, src_gt_disp, trg_gt_disp, src_gt_disp_bg, trg_gt_disp_bg,
                 src_gt_tex_bg, trg_gt_tex_bg):
    """Foreground / background texture and disparity errors of the first / last layer (ldi_pred_eval.py:477-533)."""
    out, norm = {}, {}
    nl = opts.n_layers
    for key, layer, gt_tex, gt_disp, thr in (('bg', nl - 1, (src_gt_tex_bg, trg_gt_tex_bg), (src_gt_disp_bg, trg_gt_disp_bg),
                                              (src_gt_disp_bg, trg_gt_disp_bg)),
                                             ('fg', 0, (imgs_src, imgs_trg), (src_gt_disp, trg_gt_disp),
                                              (opts.bg_layer_disp, opts.bg_layer_disp))):
        vs = (src_gt_disp > thr[0]).to(imgs_src.dtype)
        vt = (trg_gt_disp > thr[1]).to(imgs_src.dtype)
        out[key + '_tex_error'] = (((ldi_src[0][layer] - gt_tex[0]).abs() * vs).sum() / 3
                                   + ((ldi_trg[0][layer] - gt_tex[1]).abs() * vt).sum() / 3)
        out[key + '_disp_error'] = (((ldi_src[2][layer] - gt_disp[0]).abs() * vs).sum()
                                    + ((ldi_trg[2][layer] - gt_disp[1]).abs() * vt).sum())
        norm[key + '_tex_error'] = norm[key + '_disp_error'] = (vt + vs).sum()
    return out, norm


def define_metrics(opts, ldi_src, ldi_trg, imgs_src, imgs_trg, k_s, k_t, rot_mat, trans_mat, src_gt_disp=None,
                   trg_gt_disp=None, src_gt_disp_bg=None, trg_gt_disp_bg=None, src_gt_tex_bg=None, trg_gt_tex_bg=None):
    """ldi_pred_eval.py:297-548 on the B200 path.  LDIs as returned by nets.ldi_predictor (disparities already scaled by
    max_disp), images [B,H,W,3], cameras as in forward_splat.  -> (metrics, metrics_norm): {name: 0-d tensor}."""
    B, H, W, _ = imgs_src.shape
    pc = nn_helpers.pixel_coords(B, H, W, _device=imgs_src.device)
    synthetic = opts.dataset == 'synthetic'
    disocc = synthetic or (opts.dataset == 'kitti' and getattr(opts, 'kitti_dl_disparities', False))
    inv_rot = nn_helpers.transpose(rot_mat)                                         # ldi_pred_eval.py:180-181
    inv_trans = -torch.matmul(inv_rot, trans_mat)
    masks = None
    if synthetic:                                                                   # ldi_pred_eval.py:152-161
        src2trg = projection.forward_projection_matrix(k_s, k_t, rot_mat, trans_mat)
        trg2src = projection.inverse_projection_matrix(k_s, k_t, rot_mat, trans_mat)
        masks = {'src': projection.disocclusion_mask(src_gt_disp, trg_gt_disp, pc, src2trg),
                 'trg': projection.disocclusion_mask(trg_gt_disp, src_gt_disp, pc, trg2src)}
    elif disocc:                                                                    # ldi_pred_eval.py:172-173
        masks = {'src': (src_gt_disp == 0), 'trg': (trg_gt_disp == 0)}
    kw = dict(compose_layers=True, compute_trg_disp=True, trg_downsampling=opts.trg_splat_downsampling,
              zbuf_scale=opts.zbuf_scale, bg_layer_disp=opts.bg_layer_disp, max_disp=opts.max_disp)
    with torch.no_grad():
        r_trg = ldi_utils.forward_splat(tuple(ldi_src), pc, k_s, k_t, rot_mat, trans_mat, **kw)
        r_src = ldi_utils.forward_splat(tuple(ldi_trg), pc, k_t, k_s, inv_rot, inv_trans, **kw)
        acc = metrics_from_renders(opts, {'trg': (r_trg[0], r_trg[2]), 'src': (r_src[0], r_src[2])},
                                   {'trg': imgs_trg, 'src': imgs_src},
                                   {'trg': trg_gt_disp, 'src': src_gt_disp} if src_gt_disp is not None else None, masks)
        metrics, norm = {'compose_loss': acc['compose']}, {'compose_loss': acc['valid']}
        if synthetic:
            e, n = layer_errors(opts, ldi_src, ldi_trg, imgs_src, imgs_trg, src_gt_disp, trg_gt_disp, src_gt_disp_bg,
                                trg_gt_disp_bg, src_gt_tex_bg, trg_gt_tex_bg)
            metrics.update(e); norm.update(n)
        if disocc:
            metrics['compose_loss_disocc'], norm['compose_loss_disocc'] = acc['compose_d'], acc['valid_d']
        if synthetic:
            metrics['depth_loss'], norm['depth_loss'] = acc['depth'], acc['valid']
            metrics['depth_loss_disocc'], norm['depth_loss_disocc'] = acc['depth_d'], acc['valid_d']
    return metrics, norm


class MetricsAccumulator(object):
    """test_utils.py:225-262: per-iteration metric sums and normalisers are appended, the reported value of a metric is
    sum(metric) / sum(normaliser) over the evaluation set."""

    def __init__(self):
        self.metrics_data, self.metrics_norm_data = {}, {}

    def add(self, metrics, metrics_norm):
        for k in metrics:
            self.metrics_data.setdefault(k, []).append(float(metrics[k]))
            self.metrics_norm_data.setdefault(k, []).append(float(metrics_norm[k]))

    def means(self):
        return {k: float(np.sum(self.metrics_data[k]) / np.sum(self.metrics_norm_data[k])) for k in self.metrics_data}

    def write(self, path):
        """results.txt of test_utils.py:255-262."""
        with open(path, 'w') as f:
            for k, v in self.means().items():
                f.write('Mean {}: {}\n'.format(k, v))
