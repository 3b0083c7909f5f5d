"""Checkpoint I/O of the trainer: the saver / restorer plumbing of lsi/nnutils/train_utils.py:172-200,224-232 and
lsi/nnutils/helpers.py:27-62 (reference tree) re-hosted on numpy archives.

A checkpoint is one `.npz` file whose keys are the reference's TF variable names (`encoder_decoder_unet/cnv1/weights`,
`ldi_tex_disp/pixelwise_pred/upsample_0/decoder/upcnv3/BatchNorm/beta`, ...) plus `global_step`, and -- for exact resume --
the Adam slots under TF's slot names (`<var>/Adam`, `<var>/Adam_1`) with Adam's own step count `adam_t` (the reference's
Saver keeps model variables and the step counter only, so its Adam restarts after every restore; a file without slots does the
same here).  `tools/tf1_ckpt_to_npz.py` writes the same layout
from a TF-1 `model-<step>` checkpoint where TensorFlow is installed (it is not in this image), so reference snapshots load
through the same path.  File naming follows `Trainer.save` (train_utils.py:224-232): `model-<global_step>.npz` and
`model.latest.npz`; the text file `checkpoint` in the directory records the most recent one, as tf.train.Saver does for
tf.train.latest_checkpoint (train_utils.py:190).
"""
import os

import numpy as np
import torch

INDEX_FILE = 'checkpoint'
MODEL_NAME = 'model'
# keys of an archive that are not model variables.  The reference creates its step counter inside tf.name_scope('train_op')
# (train_utils.py:107-116), so a converted TF-1 snapshot carries it as 'train_op/global_step'.
GLOBAL_STEP_KEYS = ('global_step', 'train_op/global_step')
NON_VARIABLE_KEYS = GLOBAL_STEP_KEYS + ('adam_t',)


def global_step_key(saved):
    for k in GLOBAL_STEP_KEYS:
        if k in saved:
            return k
    return None


def checkpoint_path(checkpoint_dir, step):
    """train_utils.py:224-232: 'latest' -> model.latest, else model-<step>."""
    if step == 'latest':
        return os.path.join(checkpoint_dir, MODEL_NAME + '.latest.npz')
    return os.path.join(checkpoint_dir, '%s-%d.npz' % (MODEL_NAME, int(step)))


def save_checkpoint(path, variables, global_step=0, adam_m=None, adam_v=None, adam_t=None):
    """variables / adam_m / adam_v: {TF name: tensor}.  Writes atomically (tmp + rename) and updates the directory index."""
    arrays = {k: v.detach().cpu().numpy() for k, v in variables.items()}
    arrays['global_step'] = np.asarray(int(global_step), dtype=np.int64)
    if adam_t is not None and adam_m is not None:
        arrays['adam_t'] = np.asarray(int(adam_t), dtype=np.int64)
    for slot, d in (('Adam', adam_m), ('Adam_1', adam_v)):
        if d is not None:
            for k, v in d.items():
                arrays['%s/%s' % (k, slot)] = v.detach().cpu().numpy()
    d = os.path.dirname(os.path.abspath(path))
    os.makedirs(d, exist_ok=True)
    tmp = path + '.tmp.npz'
    np.savez(tmp, **arrays)
    os.replace(tmp, path)
    with open(os.path.join(d, INDEX_FILE), 'w') as f:
        f.write('model_checkpoint_path: "%s"\n' % os.path.basename(path))
    return path


def latest_checkpoint(checkpoint_dir):
    """tf.train.latest_checkpoint (train_utils.py:190): the file the directory index names, or None."""
    idx = os.path.join(checkpoint_dir, INDEX_FILE)
    if not os.path.isfile(idx):
        return None
    for line in open(idx):
        if line.startswith('model_checkpoint_path:'):
            name = line.split(':', 1)[1].strip().strip('"')
            path = name if os.path.isabs(name) else os.path.join(checkpoint_dir, name)
            return path if os.path.isfile(path) else None
    return None


def read_checkpoint(path):
    """-> {name: np.ndarray} (the tf.train.NewCheckpointReader role)."""
    with np.load(path) as z:
        return {k: z[k] for k in z.files}


class Restorer(object):
    """What helpers.optimistic_restorer returns: the variables that will be restored, and `restore(...)`."""

    def __init__(self, save_file, saved, var_names, new_vars, shape_mismatch):
        self.save_file, self._saved = save_file, saved
        self.var_names, self.new_vars, self.shape_mismatch = var_names, new_vars, shape_mismatch

    def restore(self, store, save_file=None):
        saved = self._saved if save_file in (None, self.save_file) else read_checkpoint(save_file)
        with torch.no_grad():
            for k in self.var_names:
                store.vars[k].copy_(torch.from_numpy(np.asarray(saved[k])).to(store.vars[k].device, torch.float32))
        k = global_step_key(saved)
        return int(saved[k]) if k else 0


def optimistic_restorer(save_file, store, vars_all=None):
    """helpers.py:27-62 -- restore the variables of `store` (or the sub-list vars_all of their names) that are present in
    save_file with the same shape; variables missing from the file are reported as new, shape mismatches are skipped."""
    saved = read_checkpoint(save_file)
    names = sorted(store.vars if vars_all is None else vars_all)
    present = [k for k in names if k in saved]
    new_vars = [k for k in names if k not in saved]
    var_names, mismatch = [], []
    for k in present:
        (var_names if list(saved[k].shape) == list(store.vars[k].shape) else mismatch).append(k)
    return Restorer(save_file, saved, var_names, new_vars, mismatch)
