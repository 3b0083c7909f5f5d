"""Mirror of the training wiring: lsi/nnutils/train_utils.py (Trainer: Adam, seeds, step loop) and the model/loss
wiring of ldi_enc_dec.py (Trainer.define_pred_graph :175-228, define_loss_graph :265-410), re-hosted on torch +
the B200 kernels.  The TF session / Supervisor / Saver / summary plumbing of the reference is out of scope; what is
kept is the arithmetic and the hyper-parameters:

  * two U-Net towers with shared variables (src and trg image), heads, disp *= max_disp      ldi_enc_dec.py:196-221
  * view-synthesis loss                                                                      ldi_enc_dec.py:265-410
  * Adam(learning_rate=1e-4, beta1=0.9, beta2=0.999, eps=1e-8) on every variable with a gradient   train_utils.py:56-57,107-117
  * seed 0                                                                                   train_utils.py:158-160

Data parallelism (not in the reference): one process per GPU, each rank runs the step on its shard of the
batch (batch-norm statistics are per replica), then ONE all-reduce (sum) of the flat gradient buffer over NCCL and a
fused Adam step on the flat parameter buffer with the gradients scaled by 1/world_size.
"""
import types

import torch

from lsi import _b200
from lsi.loss import loss as loss_mod
from lsi.nnutils import helpers as nn_helpers
from lsi.nnutils import nets


def default_opts(**kw):
    """Flag defaults of ldi_enc_dec.py:39-123 / train_utils.py:31-57 with the synthetic-dataset constants of main()
    (ldi_enc_dec.py:415-420); pass dataset='kitti' for the KITTI ones (:421-425)."""
    o = types.SimpleNamespace(
        n_layers=2, batch_size=2, img_height=256, img_width=256, learning_rate=1e-4, beta1=0.9,
        self_cons_wt=1.0, indep_splat_wt=1.0, compose_splat_wt=1.0, splat_bdry_ignore=0.1, zbuf_scale=50.0,
        trg_splat_downsampling=0.5, disp_smoothness_wt=0.1, incr_depth_wt=10.0, l0_self_cons=False,
        use_unet=True, n_layerwise_steps=3, pred_ldi_masks=False, bg_layer_disp=0.2, max_disp=1.0, dataset='synthetic')
    if kw.get('dataset') == 'kitti':
        o.bg_layer_disp, o.max_disp = 1e-3, 0.4
    o.__dict__.update(kw)
    return o


def shard_batch(batch, rank, world_size):
    """Contiguous, disjoint, complete split of the leading (batch) dimension across ranks."""
    out = {}
    for k, v in batch.items():
        n = v.shape[0]
        if n % world_size:
            raise ValueError('batch size %d is not divisible by world size %d' % (n, world_size))
        per = n // world_size
        out[k] = v[rank * per:(rank + 1) * per]
    return out


def allreduce_sum_(flat, group=None):
    """The one collective of the training step: sum the flat gradient buffer across ranks (NCCL on GPUs; gloo in the
    CPU tests).  No-op when torch.distributed is not initialised."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        return dist.get_world_size(group)
    return 1


def predict_ldi(img, opts, store, reuse):
    """ldi_enc_dec.py:196-213 for one tower: U-Net trunk -> heads -> disp *= max_disp.  Images whose size the reference
    U-Net cannot take (not a multiple of 128) are zero-padded and the prediction cropped (nets.pad_to_legal)."""
    padded, (h, w) = nets.pad_to_legal(img)
    _, feat_dec, skip_feat, _ = nets.encoder_decoder_unet(padded, nl_diff_enc_dec=opts.n_layerwise_steps, reuse=reuse,
                                                          _store=store)
    # the crop back to (h, w) is fused into the prediction conv (it only evaluates the top-left window)
    tex, masks, disps = nets.ldi_predictor(feat_dec, n_layers=opts.n_layers, reuse=reuse,
                                           n_layerwise_steps=opts.n_layerwise_steps, skip_feat=skip_feat,
                                           pred_masks=opts.pred_ldi_masks, _store=store, _out_hw=(h, w))
    if opts.max_disp == 1:
        return [tex, masks, disps]
    if not torch.is_grad_enabled():
        # inference: scale the disparity channel of the packed [L,B,H,W,4] head output in place, so that textures and
        # disparities stay views of one packed tensor and the renderer reads 16 bytes per pixel-layer
        disps.mul_(opts.max_disp)
        return [tex, masks, disps]
    return [tex, masks, disps * opts.max_disp]


class Trainer(object):
    """One training step of ldi_enc_dec.py on the B200 path."""

    def __init__(self, opts, store=None, group=None):
        self.opts = opts
        self.store = store if store is not None else nets.ParamStore(seed=0)
        self.group = group
        self.step_count = 0
        self.m = self.v = None
        self._built = False

    def define_pred_graph(self, imgs_src, imgs_trg):
        """ldi_enc_dec.py:175-228: both towers share their variables (reuse=True for the second)."""
        ldi_src = predict_ldi(imgs_src, self.opts, self.store, reuse=self._built)
        self._built = True
        ldi_trg = predict_ldi(imgs_trg, self.opts, self.store, reuse=True)
        return ldi_src, ldi_trg

    def define_loss_graph(self, ldi_src, ldi_trg, batch):
        """ldi_enc_dec.py:265-410."""
        b, h, w, _ = batch['imgs_src'].shape
        pc = nn_helpers.pixel_coords(b, h, w, device=batch['imgs_src'].device)
        return loss_mod.view_synthesis_loss(ldi_src, ldi_trg, batch['imgs_src'], batch['imgs_trg'], pc, batch['k_s'],
                                            batch['k_t'], batch['rot_mat'], batch['trans_mat'], self.opts)

    def train_step(self, batch):
        """forward -> loss -> backward -> all-reduce(sum) of the flat gradients -> Adam.  Returns (total_loss, parts)."""
        if self.store.flat is None:
            with torch.no_grad():                    # create every variable once, then flatten
                self.define_pred_graph(batch['imgs_src'][:1], batch['imgs_trg'][:1])
            flat, _ = self.store.flatten()
            self.m, self.v = torch.zeros_like(flat), torch.zeros_like(flat)
        self.store.zero_grad()
        ldi_src, ldi_trg = self.define_pred_graph(batch['imgs_src'], batch['imgs_trg'])
        total, parts = self.define_loss_graph(ldi_src, ldi_trg, batch)
        total.backward()
        world = allreduce_sum_(self.store.flat_grad, self.group)
        self.step_count += 1
        o = self.opts
        _b200.call('lsi_b200_adam_step', _b200.ptr(self.store.flat), _b200.ptr(self.store.flat_grad), _b200.ptr(self.m),
                   _b200.ptr(self.v), self.store.flat.numel(), o.learning_rate, o.beta1, 0.999, 1e-8, self.step_count,
                   1.0 / world, _b200.stream())
        return total.detach(), {k: (v.detach() if torch.is_tensor(v) else v) for k, v in parts.items()}
