"""CPU: checkpoint I/O of the trainer (lsi.nnutils.checkpoint; reference: train_utils.py:172-200,224-232 and
helpers.optimistic_restorer, helpers.py:27-62) on a CPU ParamStore -- pure host logic, no kernels."""
import os

import numpy as np
import pytest
import torch

import _util  # noqa: F401  (puts the package on sys.path)


def _store(seed, names_shapes):
    from lsi.nnutils import nets
    st = nets.ParamStore(device='cpu', seed=seed)
    for n, shp in names_shapes:
        st.get(n, shp, reuse=False, kind='weights' if len(shp) == 4 else 'beta')
    return st


VARS = [('encoder_decoder_unet/cnv1/weights', [7, 7, 3, 32]), ('encoder_decoder_unet/cnv1/BatchNorm/beta', [32]),
        ('ldi_tex_disp/pixelwise_pred/upsample_0/pred_0/weights', [3, 3, 32, 4]),
        ('ldi_tex_disp/pixelwise_pred/upsample_0/pred_0/biases', [4])]


def test_save_restore_round_trip_and_index(tmp_path):
    from lsi.nnutils import checkpoint as ck
    a = _store(1, VARS)
    with torch.no_grad():
        a.vars[VARS[1][0]].add_(0.25)
    d = str(tmp_path / 'snap')
    assert ck.latest_checkpoint(d) is None
    p1 = ck.save_checkpoint(ck.checkpoint_path(d, 2000), a.vars, global_step=2000)
    assert os.path.basename(p1) == 'model-2000.npz' and ck.latest_checkpoint(d) == p1
    p2 = ck.save_checkpoint(ck.checkpoint_path(d, 'latest'), a.vars, global_step=2100)
    assert os.path.basename(p2) == 'model.latest.npz' and ck.latest_checkpoint(d) == p2      # index follows the last save
    saved = ck.read_checkpoint(p2)
    assert sorted(saved) == sorted([n for n, _ in VARS] + ['global_step']) and int(saved['global_step']) == 2100
    b = _store(7, VARS)                                         # different initialisation
    r = ck.optimistic_restorer(p2, b)
    assert r.var_names == sorted(n for n, _ in VARS) and not r.new_vars and not r.shape_mismatch
    assert r.restore(b) == 2100
    for n, _ in VARS:
        assert torch.equal(a.vars[n], b.vars[n])


def test_optimistic_restorer_skips_new_and_reshaped_variables(tmp_path):
    """helpers.py:27-62: only variables present in the file with the same shape are restored."""
    from lsi.nnutils import checkpoint as ck
    from lsi.nnutils import helpers
    # exported under the reference's module and signature too (helpers.py:27: optimistic_restorer(save_file, vars_all=None))
    assert helpers.optimistic_restorer.__doc__ and 'helpers.py:27-62' in helpers.optimistic_restorer.__doc__
    a = _store(1, VARS[:3])
    path = ck.save_checkpoint(ck.checkpoint_path(str(tmp_path), 5), a.vars, global_step=5)
    # same names, but pred weights now predict 5 channels (pred_ldi_masks) and there is a variable the file lacks
    b = _store(3, [VARS[0], VARS[1], (VARS[2][0], [3, 3, 32, 5]), VARS[3]])
    before = {k: v.detach().clone() for k, v in b.vars.items()}
    r = ck.optimistic_restorer(path, b)
    assert r.var_names == sorted([VARS[0][0], VARS[1][0]])
    assert r.new_vars == [VARS[3][0]] and r.shape_mismatch == [VARS[2][0]]
    r.restore(b)
    assert torch.equal(b.vars[VARS[0][0]], a.vars[VARS[0][0]]) and torch.equal(b.vars[VARS[1][0]], a.vars[VARS[1][0]])
    assert torch.equal(b.vars[VARS[2][0]], before[VARS[2][0]]) and torch.equal(b.vars[VARS[3][0]], before[VARS[3][0]])
    # restricting the variable list (vars_all) restores only those
    c = _store(9, VARS[:3])
    r = ck.optimistic_restorer(path, c, vars_all=[VARS[1][0]])
    assert r.var_names == [VARS[1][0]]


def test_trainer_resume_protocol(tmp_path):
    """train_utils.py:176-200: latest checkpoint of the run directory wins; else the pretrained net (optimistic); else fresh.
    Adam slots and global_step survive a save/restore (exact resume)."""
    from lsi.nnutils import checkpoint as ck
    from lsi.nnutils import train_utils
    opts = train_utils.default_opts()
    run = str(tmp_path / 'run')
    tr = train_utils.Trainer(opts, store=_store(1, VARS))
    assert tr.init_from_checkpoints(run) == ('fresh', None)
    flat, _ = tr.store.flatten()
    tr.m, tr.v = torch.rand_like(flat), torch.rand_like(flat)
    tr.step_count = 1234
    path = tr.save(run, tr.step_count)
    assert os.path.basename(path) == 'model-1234.npz'
    saved = ck.read_checkpoint(path)
    assert VARS[0][0] + '/Adam' in saved and VARS[0][0] + '/Adam_1' in saved          # TF slot names
    tr2 = train_utils.Trainer(opts, store=_store(5, VARS))
    flat2, _ = tr2.store.flatten()
    tr2.m, tr2.v = torch.zeros_like(flat2), torch.zeros_like(flat2)
    what, src = tr2.init_from_checkpoints(run, pretrain_name='other', pretrain_iter=7)
    assert (what, src) == ('resumed', path) and tr2.step_count == 1234
    assert torch.equal(tr2.store.flat, tr.store.flat) and torch.equal(tr2.m, tr.m) and torch.equal(tr2.v, tr.v)
    # pretrained-net path: a sibling directory, optimistic restore, used only when the run directory has no checkpoint
    pre = train_utils.Trainer(opts, store=_store(11, VARS[:2]))
    pre.step_count = 7
    pre.save(str(tmp_path / 'other'), 7)
    tr3 = train_utils.Trainer(opts, store=_store(13, VARS))
    keep = tr3.store.vars[VARS[2][0]].detach().clone()
    what, src = tr3.init_from_checkpoints(str(tmp_path / 'run3'), pretrain_name='other', pretrain_iter=7)
    assert what == 'pretrained' and os.path.basename(src) == 'model-7.npz' and tr3.step_count == 7
    assert torch.equal(tr3.store.vars[VARS[0][0]], pre.store.vars[VARS[0][0]])
    assert torch.equal(tr3.store.vars[VARS[2][0]], keep)
    # a checkpoint that does not match the model is an error for the strict restore
    with pytest.raises(RuntimeError):
        train_utils.Trainer(opts, store=_store(1, VARS)).restore(ck.checkpoint_path(str(tmp_path / 'other'), 7))


def test_resume_into_fresh_trainer_creates_variables_from_the_checkpoint(tmp_path):
    """A Trainer whose variables do not exist yet (they are created by the first forward) must come out of a strict resume with
    the SAVED weights -- not report success and later random-initialise them -- and Adam's slots / step count must survive."""
    from lsi.nnutils import checkpoint as ck
    from lsi.nnutils import train_utils
    opts = train_utils.default_opts()
    run = str(tmp_path / 'run')
    tr = train_utils.Trainer(opts, store=_store(1, VARS))
    flat, _ = tr.store.flatten()
    tr.m, tr.v = torch.rand_like(flat), torch.rand_like(flat)
    tr.step_count, tr.adam_t = 500, 321
    tr.save(run, 500)
    fresh = train_utils.Trainer(opts)                           # default ParamStore, no variables
    fresh.store = __import__('lsi.nnutils.nets', fromlist=['x']).ParamStore(device='cpu', seed=99)
    assert not fresh.store.vars
    assert fresh.init_from_checkpoints(run)[0] == 'resumed' and fresh.step_count == 500 and fresh.adam_t == 321
    assert sorted(fresh.store.vars) == sorted(n for n, _ in VARS)
    for n, _ in VARS:
        assert torch.equal(fresh.store.vars[n].detach(), tr.store.vars[n].detach())
    # the slots wait until train_step allocates m / v (here: done by hand), then land in them
    assert fresh.m is None and fresh._pending_slots is not None
    flat2, _ = fresh.store.flatten()
    fresh.m, fresh.v = torch.zeros_like(flat2), torch.zeros_like(flat2)
    fresh._apply_pending_slots()
    assert torch.equal(fresh.m, tr.m) and torch.equal(fresh.v, tr.v)
    # a checkpoint without slots (the reference's own snapshots): Adam restarts at t = 0 with zero slots
    ck.save_checkpoint(ck.checkpoint_path(str(tmp_path / 'ref'), 7), tr.store.vars, global_step=7)
    again = train_utils.Trainer(opts, store=_store(5, VARS))
    f3, _ = again.store.flatten()
    again.m, again.v, again.adam_t = torch.ones_like(f3), torch.ones_like(f3), 40
    again.restore(ck.checkpoint_path(str(tmp_path / 'ref'), 7))
    assert again.step_count == 7 and again.adam_t == 0 and float(again.m.abs().max()) == 0 and float(again.v.abs().max()) == 0


def test_deferred_pretrain_restore_and_tf_step_alias(tmp_path):
    """Optimistic (pretrained-net) restore into a Trainer without variables is applied once the variables exist; a converted TF-1
    snapshot names its counter 'train_op/global_step' (train_utils.py:107-116)."""
    import numpy as np
    from lsi.nnutils import checkpoint as ck
    from lsi.nnutils import nets, train_utils
    opts = train_utils.default_opts()
    pre = _store(11, VARS[:2])
    arrays = {k: v.detach().numpy() for k, v in pre.vars.items()}
    arrays['train_op/global_step'] = np.asarray(4000, dtype=np.int64)
    os.makedirs(str(tmp_path / 'pre'))
    np.savez(ck.checkpoint_path(str(tmp_path / 'pre'), 4000), **arrays)
    tr = train_utils.Trainer(opts, store=nets.ParamStore(device='cpu', seed=3))
    what, _ = tr.init_from_checkpoints(str(tmp_path / 'run'), pretrain_name='pre', pretrain_iter=4000)
    assert what == 'pretrained' and tr.step_count == 4000 and tr.adam_t == 0 and tr._deferred is not None

    def fake_forward(a, b):                                       # stands for the first forward, which creates the variables
        for n, shp in VARS:
            tr.store.get(n, shp, reuse=False, kind='weights' if len(shp) == 4 else 'beta')
    tr.define_pred_graph = fake_forward
    keep = None
    tr._build_variables({'imgs_src': torch.zeros(1, 1), 'imgs_trg': torch.zeros(1, 1)})
    assert tr._deferred is None and tr.store.flat is not None
    assert torch.equal(tr.store.vars[VARS[0][0]].detach(), pre.vars[VARS[0][0]].detach())      # restored
    assert tr.store.vars[VARS[2][0]].abs().sum() > 0                                            # kept its initialisation


def test_host_view_pipeline_rejects_depth_one():
    from lsi.nnutils import train_utils
    with pytest.raises(ValueError, match='depth >= 2'):
        train_utils.HostViewPipeline(None, None, {}, 1, 8, 8, 'cpu', depth=1)
