"""Mirror of the evaluation wiring: lsi/nnutils/test_utils.py (Tester: checkpoint selection, evaluation loop, sum / sum-of-
normalisers reporting, :182-262) and the model / metric wiring of ldi_pred_eval.py (define_pred_graph :117-224, define_metrics
:297-548, feed :226-295), re-hosted on torch + the B200 kernels.  Out of scope, as in DESIGN.md section 0: the HTML summary page,
PNG / .mat dumps of visuals and predictions (save_visuals, save_preds, write_summary_page).
"""
import os
import types

import numpy as np
import torch

from lsi.nnutils import checkpoint as ckpt
from lsi.nnutils import eval_metrics
from lsi.nnutils import nets
from lsi.nnutils import train_utils


def default_opts(**kw):
    """Flag defaults of ldi_pred_eval.py:39-115 / test_utils.py:38-72 with the dataset constants of main() (:556-564)."""
    o = train_utils.default_opts(**kw)
    base = dict(n_layers=3, zbuf_scale=10.0, trg_splat_downsampling=1.0, splat_bdry_ignore=0.0, batch_norm_training=True,
                data_split='val', synth_dl_eval_data=True, train_iter=0, num_eval_iter=100, visuals_freq=10,
                save_pred_results=False, results_eval_dir='/code/lsi/cachedir/evaluation/')
    for k, v in base.items():
        if k not in kw:
            setattr(o, k, v)
    return o


class Tester(object):
    """test_utils.py:75-262 on the B200 path."""

    def __init__(self, opts, store=None):
        self.opts = opts
        self.store = store if store is not None else nets.ParamStore(seed=0)
        self._built = False

    def define_data_loader(self):
        """ldi_pred_eval.py:117-126."""
        opts = self.opts
        if opts.dataset == 'synthetic':
            from lsi.data.syntheticPlanes import data as synthetic_planes
            self.data_loader = synthetic_planes.DataLoader(opts)
        elif opts.dataset == 'kitti':
            from lsi.data.kitti import data as kitti_data
            self.data_loader = kitti_data.DataLoader(opts)
            self.data_loader.define_queues()
            self.data_loader.preload_calib_files()
        else:
            raise ValueError('unknown dataset %r' % (opts.dataset,))

    def feed(self):
        """ldi_pred_eval.py:226-295: one batch, keyed like the reference's placeholders."""
        opts = self.opts
        data = self.data_loader.forward(opts.batch_size)
        dev = self.store.device
        f = lambda x: torch.as_tensor(x, dtype=torch.float32).to(dev)
        names = ['imgs_src', 'imgs_trg', 'k_s', 'k_t', 'rot_mat', 'trans_mat']
        if opts.dataset == 'synthetic' and getattr(opts, 'synth_dl_eval_data', False):
            names += ['n_hat', 'a', 'src_gt_disp', 'src_gt_disp_bg', 'trg_gt_disp', 'trg_gt_disp_bg', 'src_gt_tex_bg', 'trg_gt_tex_bg']
        elif opts.dataset == 'kitti' and getattr(opts, 'kitti_dl_disparities', False):
            names += ['src_gt_disp', 'trg_gt_disp']
        return {k: f(v) for k, v in zip(names, data)}

    def define_pred_graph(self, batch):
        """ldi_pred_eval.py:175-224: both towers with shared variables, batch-statistics batch norm (batch_norm_training)."""
        with torch.no_grad():
            ldi_src = train_utils.predict_ldi(batch['imgs_src'], self.opts, self.store, reuse=self._built)
            self._built = True
            ldi_trg = train_utils.predict_ldi(batch['imgs_trg'], self.opts, self.store, reuse=True)
        f32 = lambda ldi: [nets.to_float(t) if not isinstance(t, torch.Tensor) else t for t in ldi]
        return f32(ldi_src), f32(ldi_trg)

    def define_metrics(self, ldi_src, ldi_trg, batch):
        """ldi_pred_eval.py:297-548."""
        extra = {k: batch[k] for k in ('src_gt_disp', 'trg_gt_disp', 'src_gt_disp_bg', 'trg_gt_disp_bg', 'src_gt_tex_bg', 'trg_gt_tex_bg')
                 if k in batch}
        return eval_metrics.define_metrics(self.opts, ldi_src, ldi_trg, batch['imgs_src'], batch['imgs_trg'], batch['k_s'], batch['k_t'],
                                           batch['rot_mat'], batch['trans_mat'], **extra)

    def load_checkpoint(self):
        """test_utils.py:202-221: `model-<train_iter>` when train_iter > 0, else the latest checkpoint of checkpoint_dir (if any)."""
        opts = self.opts
        path = None
        if getattr(opts, 'train_iter', 0) > 0:
            path = ckpt.checkpoint_path(opts.checkpoint_dir, opts.train_iter)
        else:
            path = ckpt.latest_checkpoint(opts.checkpoint_dir)
        if path is not None:
            saved = ckpt.read_checkpoint(path)
            self.store.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in saved.items()
                                        if k not in ckpt.NON_VARIABLE_KEYS and not k.endswith('/Adam') and not k.endswith('/Adam_1')})
            self._built = True
        return path

    def test(self, on_step=None):
        """test_utils.py:182-262 -- the evaluation routine: seed 0, data loader, checkpoint, opts.num_eval_iter batches; per batch the
        metric sums and their normalisers are accumulated, the reported value of a metric is sum / sum of normalisers; results.txt
        is written to opts.results_eval_dir.  Returns {metric: mean}."""
        opts = self.opts
        torch.manual_seed(0)
        np.random.seed(0)
        self.define_data_loader()
        self.checkpoint = self.load_checkpoint()
        acc = eval_metrics.MetricsAccumulator()
        for step in range(1, opts.num_eval_iter + 1):
            batch = self.feed()
            ldi_src, ldi_trg = self.define_pred_graph(batch)
            metrics, norm = self.define_metrics(ldi_src, ldi_trg, batch)
            acc.add(metrics, norm)
            if on_step is not None:
                on_step(step, metrics, norm)
        self.metrics_data, self.metrics_norm_data = acc.metrics_data, acc.metrics_norm_data
        if getattr(opts, 'results_eval_dir', None):
            os.makedirs(opts.results_eval_dir, exist_ok=True)
            acc.write(os.path.join(opts.results_eval_dir, 'results.txt'))
        return acc.means()
