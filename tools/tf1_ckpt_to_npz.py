"""Convert a TF-1 checkpoint written by the reference (train_utils.py:224-232: `model-<step>` / `model.latest`) to the
`.npz` layout of lsi.nnutils.checkpoint.  Needs TensorFlow (any version with tf.train.load_checkpoint); it is NOT installed in
the build image, so this script is provided for maintainers and is not exercised by the tests.
    python tools/tf1_ckpt_to_npz.py /path/to/model-400000 out_dir/model-400000.npz"""
import sys

import numpy as np


def main(src, dst):
    import tensorflow as tf
    reader = tf.train.load_checkpoint(src)
    arrays = {}
    for name in reader.get_variable_to_shape_map():
        arrays[name] = reader.get_tensor(name)       # TF names are kept verbatim.  The reference's Saver holds the model variables and the
        # step counter only (train_utils.py:172-174) -- the latter as 'train_op/global_step', which lsi.nnutils.checkpoint accepts --
        # and no Adam slots: after loading such a file Adam restarts, exactly as in the reference
    np.savez(dst, **arrays)
    print('wrote %d arrays to %s' % (len(arrays), dst))


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2])
