"""GPU: the training and evaluation ROUTINES (lsi/nnutils/train_utils.py:149-222, lsi/nnutils/test_utils.py:182-262 of the reference):
Trainer.train() iterates the synthetic planar-room loader with the reference's logging / snapshot cadence and resumes from its own
snapshots; Tester.test() loads the snapshot, accumulates metric sums over normaliser sums and writes results.txt."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_train_loop_then_test_loop(tmp_path):
    from lsi.nnutils import checkpoint as ck
    from lsi.nnutils import nets, test_utils, train_utils
    run = str(tmp_path / 'snapshots')
    common = dict(dataset='synthetic', n_layers=1, batch_size=2, img_height=128, img_width=128, n_obj_max=2, n_obj_min=1, checkpoint_dir=run)
    opts = train_utils.default_opts(num_iter=6, log_freq=2, save_latest_freq=3, checkpoint_freq=5, learning_rate=1e-3, **common)
    seen = []
    tr = train_utils.Trainer(opts, store=nets.ParamStore(seed=0))
    log = tr.train(on_log=lambda gs, total, parts: seen.append(gs))
    assert log[0] == ('init', 'fresh', None) and [r[0] for r in log[1:]] == [2, 4, 6] and seen == [2, 4, 6]
    assert all(np.isfinite(r[1]) for r in log[1:]) and set(log[1][2]) >= {'self_cons', 'compose_splat', 'indep_splat'}
    assert os.path.isfile(os.path.join(run, 'model.latest.npz')) and os.path.isfile(os.path.join(run, 'model-5.npz'))
    assert int(ck.read_checkpoint(os.path.join(run, 'model.latest.npz'))['global_step']) == 6
    assert int(ck.read_checkpoint(os.path.join(run, 'model-5.npz'))['global_step']) == 5
    assert ck.latest_checkpoint(run) == os.path.join(run, 'model.latest.npz')
    # a second run resumes (train_utils.py:190-195): fresh Trainer, variables come from the snapshot, global_step continues
    opts2 = train_utils.default_opts(num_iter=2, log_freq=1, save_latest_freq=100, checkpoint_freq=100, learning_rate=1e-3, **common)
    tr2 = train_utils.Trainer(opts2, store=nets.ParamStore(seed=5))
    log2 = tr2.train()
    assert log2[0][:2] == ('init', 'resumed') and [r[0] for r in log2[1:]] == [7, 8] and tr2.adam_t == 8
    # evaluation routine on the synthetic loader with ground truth (ldi_pred_eval.py), from the latest snapshot
    topts = test_utils.default_opts(num_eval_iter=3, results_eval_dir=str(tmp_path / 'eval'), **common)
    te = test_utils.Tester(topts, store=nets.ParamStore(seed=9))
    steps = []
    means = te.test(on_step=lambda step, m, n: steps.append(step))
    assert te.checkpoint == os.path.join(run, 'model.latest.npz') and steps == [1, 2, 3]
    assert {'compose_loss', 'compose_loss_disocc', 'depth_loss', 'depth_loss_disocc'} <= set(means)
    assert all(np.isfinite(v) for v in means.values()) and 0.0 < means['compose_loss'] < 1.0
    txt = open(os.path.join(str(tmp_path / 'eval'), 'results.txt')).read()
    assert 'Mean compose_loss: ' in txt and len(txt.strip().splitlines()) == len(means)
    # the evaluated weights are the snapshot's, not the Tester's own initialisation
    w = 'encoder_decoder_unet/cnv1/weights'
    assert torch.equal(te.store.vars[w].detach().cpu(), torch.from_numpy(ck.read_checkpoint(te.checkpoint)[w]))
