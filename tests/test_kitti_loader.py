"""KITTI loader (reference: lsi/data/kitti/data.py).  CPU: calibration parsing, camera computation and the
sequence split against fixtures produced by the reference's own functions (oracle/gen_golden_kitti.py), the oracle's AREA
resize against its defining properties, file lists on a synthetic directory tree.  GPU: the AREA-resize kernel against the
oracle, and DataLoader.forward end to end on a synthetic tree of PNG pairs."""
import os
import types

import numpy as np
import pytest
import torch

from _util import load_golden


def _calib_text(p2, p3):
    fmt = lambda a: ' '.join('%.12e' % v for v in a)
    return ('calib_time: 09-Jan-2012 13:57:47\ncorner_dist: 9.950000e-02\nS_02: 1.392000e+03 5.120000e+02\n'
            'P_rect_02: %s\nP_rect_03: %s\n' % (fmt(p2), fmt(p3)))


def test_camera_computation_matches_reference_function(tmp_path):
    from lsi.data.kitti import data as kd
    from oracle import lsi_oracle_data as OD
    g = load_golden('kitti_loader')
    for i in range(3):
        path = tmp_path / ('calib%d.txt' % i)
        path.write_text(_calib_text(g['c%d_p2' % i], g['c%d_p3' % i]))
        calib = kd.read_calib_file(str(path))
        assert isinstance(calib['calib_time'], str) and calib['S_02'].shape == (2,) and calib['P_rect_02'].shape == (12,)
        h, w = [int(v) for v in g['c%d_hw' % i]]
        got = kd.stereo_cameras(calib, tuple(g['c%d_src_shape' % i]), tuple(g['c%d_trg_shape' % i]), h, w)
        ora = OD.stereo_cameras(g['c%d_p2' % i], g['c%d_p3' % i], tuple(g['c%d_src_shape' % i]), tuple(g['c%d_trg_shape' % i]), h, w)
        for a, b, k in zip(got, ora, ('k_s', 'k_t', 'rot', 'trans')):
            ref = g['c%d_%s' % (i, k)]
            assert np.allclose(a, ref, rtol=1e-10, atol=1e-12), k
            assert np.allclose(b, ref, rtol=1e-12, atol=1e-14), k
    assert np.array_equal(kd.resize_instrinsic(np.arange(9, dtype=np.float64).reshape(3, 3) + 1, 0.67, 0.6827), g['resize_k'])


def test_sequence_list_and_split_match_reference():
    from lsi.data.kitti import data as kd
    g = load_golden('kitti_loader')
    assert kd.raw_city_sequences() == [str(s) for s in g['seq_all']]
    for split in ('train', 'val', 'test'):
        assert kd.split_sequences(split) == [str(s) for s in g['seq_' + split]]


def test_oracle_area_resize_properties():
    from oracle import lsi_oracle_data as OD
    rs = np.random.RandomState(0)
    img = rs.randint(0, 256, (12, 18, 3)).astype(np.uint8)
    box = img.reshape(6, 2, 6, 3, 3).mean(axis=(1, 3)) / 255.0                      # integer factors: the box mean
    assert np.allclose(OD.area_resize(img, 6, 6), box, atol=1e-12)
    const = np.full((375, 1242, 3), 200, np.uint8)
    assert np.allclose(OD.area_resize(const, 256, 832), 200 / 255.0, atol=1e-12)    # weights sum to one
    big = rs.randint(0, 256, (375, 1242, 3)).astype(np.uint8)                       # every input pixel is distributed exactly once:
    out = OD.area_resize(big, 256, 832)                                              # the image mean is preserved at any factor
    assert abs(out.mean() - big.mean() / 255.0) < 1e-12
    assert np.allclose(OD.area_resize(big.transpose(1, 0, 2), 832, 256), out.transpose(1, 0, 2), atol=1e-12)   # separable / symmetric
    two = np.zeros((3, 5, 1), np.uint8); two[:, 2] = 255                             # hand-computed: 5 -> 2 columns, scale 2.5
    assert np.allclose(OD.area_resize(two, 3, 2, nc=1)[0, :, 0], [0.5 / 2.5, 0.5 / 2.5], atol=1e-12)


def _fake_tree(root, n_imgs=3):
    from PIL import Image
    from lsi.data.kitti import data as kd
    rs = np.random.RandomState(5)
    g = load_golden('kitti_loader')
    for seq in kd.raw_city_sequences():
        day = seq[:10]
        os.makedirs(os.path.join(root, 'kitti_raw', day), exist_ok=True)
        with open(os.path.join(root, 'kitti_raw', day, 'calib_cam_to_cam.txt'), 'w') as f:
            f.write(_calib_text(g['c1_p2'], g['c1_p3']))
        for cam in ('image_02', 'image_03'):
            d = os.path.join(root, 'kitti_raw', day, seq + '_sync', cam, 'data')
            os.makedirs(d, exist_ok=True)
            for i in range(n_imgs):
                Image.fromarray(rs.randint(0, 256, (37, 123, 3)).astype(np.uint8)).save(os.path.join(d, '%010d.png' % i))


def _opts(root, split='val', bs=2):
    return types.SimpleNamespace(batch_size=bs, kitti_dataset_variant='raw_city', kitti_data_root=root, data_split=split,
                                 img_height=16, img_width=48)


def test_file_lists_on_a_synthetic_tree(tmp_path):
    from lsi.data.kitti import data as kd
    _fake_tree(str(tmp_path))
    dl = kd.DataLoader(_opts(str(tmp_path), 'val'))
    val = kd.split_sequences('val')
    assert len(dl.img_list_src) == 3 * len(val) and len(dl.img_list_trg) == len(dl.img_list_src)
    assert all('image_02' in s and t == s.replace('image_02', 'image_03') for s, t in zip(dl.img_list_src, dl.img_list_trg))
    assert sorted(set(dl.seq_id_list)) == sorted(set(s[:10] for s in val))
    dl.preload_calib_files()
    assert set(dl.cam_calibration) == {'2011_09_26', '2011_09_28', '2011_09_29'}
    n_train = len(kd.DataLoader(_opts(str(tmp_path), 'train')).img_list_src)
    n_test = len(kd.DataLoader(_opts(str(tmp_path), 'test')).img_list_src)
    # the excluded frame (data.py:153) is number 74 of drive 0117: absent from this 3-frame tree, so nothing is dropped
    assert n_train + n_test + len(dl.img_list_src) == 3 * 28


@pytest.mark.gpu
@pytest.mark.parametrize('H,W,C,h,w,nc', [(375, 1242, 3, 256, 832, 3), (37, 123, 4, 16, 48, 3), (64, 64, 1, 32, 16, 1),
                                          (20, 30, 3, 20, 30, 3), (9, 7, 3, 4, 3, 3)])
def test_area_resize_kernel_matches_oracle(H, W, C, h, w, nc):
    from lsi.data.kitti import data as kd
    from oracle import lsi_oracle_data as OD
    img = np.random.RandomState(H + W).randint(0, 256, (H, W, C)).astype(np.uint8)
    got = kd.area_resize(img, h, w, nc).cpu().numpy()
    ref = OD.area_resize(img, h, w, nc)
    assert got.shape == (h, w, nc) and np.abs(got - ref).max() < 1e-5


@pytest.mark.gpu
def test_loader_forward_end_to_end(tmp_path):
    from PIL import Image
    from lsi.data.kitti import data as kd
    from oracle import lsi_oracle_data as OD
    _fake_tree(str(tmp_path))
    dl = kd.DataLoader(_opts(str(tmp_path), 'val', bs=3))
    dl.define_queues(); dl.preload_calib_files()
    img_s, img_t, k_s, k_t, rot, trans = dl.forward(3)
    assert img_s.shape == (3, 16, 48, 3) and img_s.is_cuda and k_s.shape == (3, 3, 3) and trans.shape == (3, 3, 1)
    g = load_golden('kitti_loader')
    for b, name in enumerate(dl.src_image_names):
        ref = OD.area_resize(np.asarray(Image.open(name)), 16, 48)
        assert np.abs(img_s[b].cpu().numpy() - ref).max() < 1e-5
        ref_t = OD.area_resize(np.asarray(Image.open(name.replace('image_02', 'image_03'))), 16, 48)
        assert np.abs(img_t[b].cpu().numpy() - ref_t).max() < 1e-5
        ks, kt, r, t = OD.stereo_cameras(g['c1_p2'], g['c1_p3'], (37, 123, 3), (37, 123, 3), 16, 48)
        assert np.allclose(k_s[b].cpu().numpy(), ks, rtol=1e-6) and np.allclose(trans[b].cpu().numpy(), t, rtol=1e-6, atol=1e-9)
    again = dl.forward(3)                                            # the epoch order moves on and wraps
    assert dl.src_image_names and again[0].shape == (3, 16, 48, 3)


@pytest.mark.gpu
def test_prefetching_loader_equals_synchronous_loader(tmp_path):
    """define_queues() starts a decoder pool that works batches ahead; the batches (and their order, also across a batch-size
    change) must be exactly those of the synchronous loader."""
    from lsi.data.kitti import data as kd
    _fake_tree(str(tmp_path))
    sync = kd.DataLoader(_opts(str(tmp_path), 'val', bs=2)); sync.preload_calib_files()
    pre = kd.DataLoader(_opts(str(tmp_path), 'val', bs=2)); pre.define_queues(_threads=3, _prefetch=2); pre.preload_calib_files()
    for bs in (2, 2, 3, 2):
        a, b = sync.forward(bs), pre.forward(bs)
        assert sync.src_image_names == pre.src_image_names
        assert all(torch.equal(x, y) for x, y in zip(a, b))
