"""Golden fixtures for the KITTI loader's host logic: the reference's own lsi/data/kitti/data.py functions
(`resize_instrinsic` :32-36, `raw_city_sequences` :39-75, `DataLoader.forward_instance` :303-342 and the train/val/test split
of `init_img_names_seq_list` :155-170) imported unmodified over the TF shim and run on synthetic calibration data.
    python oracle/gen_golden_kitti.py  ->  tests/golden/kitti_loader.npz   (test infrastructure; needs /root/reference)"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(HERE, 'tf1_shim'))
sys.path.insert(1, '/root/reference')
from lsi.data.kitti import data as kd  # noqa: E402


def main():
    rs = np.random.RandomState(7)
    blob = {}
    # calibration of a KITTI raw day (P_rect_02 / P_rect_03 as 12 floats), three variants
    for i in range(3):
        fx = 721.5377 + rs.uniform(-5, 5); fy = fx + rs.uniform(-1, 1)
        cx, cy = 609.5593 + rs.uniform(-3, 3), 172.854 + rs.uniform(-3, 3)
        p2 = np.array([fx, 0, cx, 44.85728 + rs.uniform(-1, 1), 0, fy, cy, 0.2163791, 0, 0, 1, 2.745884e-03])
        p3 = np.array([fx, 0, cx, -339.5242 + rs.uniform(-1, 1), 0, fy, cy, 2.199936, 0, 0, 1, 2.729905e-03])
        calib = {'P_rect_02': p2, 'P_rect_03': p3}
        me = types.SimpleNamespace(w=832 if i else 416, h=256 if i else 128)
        src_shape, trg_shape = (375 - i, 1242 + i, 3), (375, 1242, 3)
        out = kd.DataLoader.forward_instance(me, None, None, src_shape, trg_shape, calib)      # data.py:303-342, unmodified
        blob['c%d_p2' % i], blob['c%d_p3' % i] = p2, p3
        blob['c%d_hw' % i] = np.array([me.h, me.w]); blob['c%d_src_shape' % i] = np.array(src_shape); blob['c%d_trg_shape' % i] = np.array(trg_shape)
        for k, v in zip(('k_s', 'k_t', 'rot', 'trans'), out[2:]):
            blob['c%d_%s' % (i, k)] = np.asarray(v, dtype=np.float64)
    blob['resize_k'] = kd.resize_instrinsic(np.arange(9, dtype=np.float64).reshape(3, 3) + 1, 0.67, 0.6827)
    # train / val / test split of the raw_city sequences (data.py:155-170): RandomState(0).shuffle + 70/15/15
    names = kd.raw_city_sequences()
    seq = list(names)
    rng = np.random.RandomState(0)
    rng.shuffle(seq)
    n_all = len(seq); n_train = int(round(0.7 * n_all)); n_val = int(round(0.15 * n_all))
    blob['seq_all'] = np.array(names)
    blob['seq_train'] = np.array(seq[0:n_train]); blob['seq_val'] = np.array(seq[n_train:n_train + n_val])
    blob['seq_test'] = np.array(seq[n_train + n_val:n_all])
    path = os.path.join(ROOT, 'tests', 'golden', 'kitti_loader.npz')
    np.savez_compressed(path, **blob)
    print(path, os.path.getsize(path), 'bytes;', 'train/val/test =', n_train, n_val, n_all - n_train - n_val)


if __name__ == '__main__':
    main()
