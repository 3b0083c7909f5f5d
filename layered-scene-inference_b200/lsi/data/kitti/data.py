"""Mirror of lsi/data/kitti/data.py (reference tree): the KITTI stereo-pair loader.

Same class, option names, sequence lists, train/val/test split, calibration parsing and camera computation as the reference.
What differs is the transport: the reference feeds TF filename queues (string_input_producer / WholeFileReader /
tf.train.batch, data.py:247-301); here the PNGs are decoded on the host (PIL) and every image is scaled to [0, 1] and
AREA-resized to the network resolution on the GPU (lsi_b200_area_resize_u8, [TF1.4] ResizeArea semantics), and the sample
order is a seeded permutation (the TF queue shuffles cannot be reproduced outside TensorFlow; seed 0 as in data.py:250,279).
`forward(bs)` returns the reference's tuple as CUDA tensors ready for lsi.nnutils.nets / lsi.geometry.ldi.
"""
import fnmatch
import os
import re

import numpy as np
import torch

from lsi import _b200


def resize_instrinsic(intrinsic, scale_x, scale_y):
    """data.py:32-36 (name as in the reference): intrinsics of an image resized by (scale_x, scale_y) = diag(sx, sy, 1) @ K."""
    return np.diag([float(scale_x), float(scale_y), 1.0]) @ np.asarray(intrinsic, dtype=np.float64)


def raw_city_sequences():
    """data.py:39-75 -- the 28 city sequences of KITTI raw."""
    day26 = [1, 2, 5, 9, 11, 13, 14, 17, 18, 48, 51, 56, 57, 59, 60, 84, 91, 93, 95, 96, 104, 106, 113, 117]
    return (['2011_09_26_drive_%04d' % i for i in day26] + ['2011_09_28_drive_0001', '2011_09_28_drive_0002',
                                                             '2011_09_29_drive_0026', '2011_09_29_drive_0071'])


_NUMERIC_LINE = re.compile(r'^[0-9eE.+\- ]+$')


def read_calib_file(file_path):
    """data.py:227-245 -- a KITTI calibration file as {key: value}: `key: v0 v1 ...` lines whose value is purely numeric become
    float64 arrays, anything else (e.g. `calib_time: 09-Jan-2012 13:57:47`) stays the stripped string."""
    entries = {}
    for line in open(file_path).read().splitlines():
        key, sep, text = line.partition(':')
        if not sep:
            continue
        text = text.strip()
        entries[key] = text
        if text and _NUMERIC_LINE.match(text):
            try:
                entries[key] = np.array(text.split(' '), dtype=np.float64)
            except ValueError:
                pass                                         # dates such as 09-01 2012 pass the character test but are not numbers
    return entries


def _camera_from_projection(p_rect):
    """P_rect = K [I | c] with the offset given in homogeneous image coordinates: returns (K [3,3], c [3]) with the camera
    offset converted to metric 3-D, c = K^-1 P[:, 3] restricted to the pinhole form K = [[fx,0,cx],[0,fy,cy],[0,0,1]]."""
    p = np.asarray(p_rect, dtype=np.float64).reshape(3, 4)
    k, col = p[:, :3].copy(), p[:, 3]
    c = np.array([(col[0] - k[0, 2] * col[2]) / k[0, 0], (col[1] - k[1, 2] * col[2]) / k[1, 1], col[2]])
    return k, c


def stereo_cameras(calib_data, src_shape, trg_shape, h, w):
    """data.py:303-342 (forward_instance without the images): the colour cameras 2 (source) and 3 (target) of the rectified rig:
    intrinsics rescaled from the image sizes to the network resolution (h, w), identity rotation (rectified pair), translation =
    difference of the two camera offsets.  -> (k_s, k_t, rot, trans [3,1])."""
    (k_src, c_src), (k_trg, c_trg) = (_camera_from_projection(calib_data[key]) for key in ('P_rect_02', 'P_rect_03'))
    k_s = resize_instrinsic(k_src, w / src_shape[1], h / src_shape[0])
    k_t = resize_instrinsic(k_trg, w / trg_shape[1], h / trg_shape[0])
    return k_s, k_t, np.eye(3), (c_trg - c_src).reshape(3, 1)


def split_sequences(data_split):
    """data.py:155-170 -- RandomState(0) shuffle of the city sequences, 70 / 15 / 15 % train / val / test."""
    seq_names = raw_city_sequences()
    rng = np.random.RandomState(0)
    rng.shuffle(seq_names)
    n_all = len(seq_names)
    n_train = int(round(0.7 * n_all))
    n_val = int(round(0.15 * n_all))
    return {'train': seq_names[0:n_train], 'val': seq_names[n_train:n_train + n_val],
            'test': seq_names[n_train + n_val:n_all]}[data_split]


def area_resize(img_u8, h, w, nc=3, device='cuda', out=None):
    """Decoded image (numpy / torch uint8 [H,W,C]) -> CUDA float [h,w,nc] in [0,1], AREA-resized (data.py:255-264); `out`: write
    into this [h,w,nc] slice of a batch tensor instead of allocating."""
    t = torch.as_tensor(np.ascontiguousarray(img_u8)) if not torch.is_tensor(img_u8) else img_u8
    if t.dim() == 2:
        t = t.unsqueeze(-1)
    if t.dtype != torch.uint8 or t.dim() != 3:
        raise RuntimeError('lsi_b200: area_resize expects a uint8 [H,W,C] image, got %s %s' % (t.dtype, tuple(t.shape)))
    t = t.contiguous().to(device)
    if not t.is_cuda:
        raise RuntimeError('lsi_b200: area_resize runs on CUDA only (no CPU fallback)')
    if out is None:
        out = torch.empty(h, w, nc, dtype=torch.float32, device=t.device)
    elif tuple(out.shape) != (h, w, nc) or not out.is_contiguous() or out.dtype != torch.float32 or out.device != t.device:
        raise RuntimeError('lsi_b200: area_resize output slice must be a contiguous float32 [%d,%d,%d] tensor on %s' % (h, w, nc, t.device))
    _b200.call('lsi_b200_area_resize_u8', _b200.ptr(t), t.shape[0], t.shape[1], t.shape[2], _b200.ptr(out), h, w, nc, _b200.stream())
    return out


class DataLoader(object):
    """data.py:78-390.  opts: batch_size, kitti_dataset_variant ('mview' | 'odom' | 'raw_city'), kitti_data_root,
    data_split, img_height, img_width[, kitti_dl_disparities]."""

    def __init__(self, opts):
        self.opts = opts
        self.batch_size = opts.batch_size
        self.dataset_variant = opts.kitti_dataset_variant
        self.output_disparities = (self.dataset_variant == 'raw_city' and getattr(opts, 'kitti_dl_disparities', False)
                                   and opts.data_split != 'train')
        self.root_dir = opts.kitti_data_root
        if self.dataset_variant == 'odom':
            self.root_dir = os.path.join(self.root_dir, 'odometry', 'dataset', 'sequences')
        elif self.dataset_variant == 'mview':
            self.root_dir = os.path.join(self.root_dir, 'stereo_multiview_2015')
            self.root_dir += '/training' if opts.data_split == 'train' else '/testing'
        elif self.dataset_variant == 'raw_city':
            self.root_dir = os.path.join(self.root_dir, 'kitti_raw')
        self.h, self.w = opts.img_height, opts.img_width
        self.init_img_names_seq_list()
        self._order = np.random.RandomState(0).permutation(len(self.img_list_src)) if self.img_list_src else np.zeros(0, int)
        self._cursor = 0

    @staticmethod
    def _pngs(top):
        out = []
        for root, _, filenames in os.walk(top):
            for filename in fnmatch.filter(filenames, '*.png'):
                out.append(os.path.join(root, filename))
        return sorted(out)             # (os.walk order is file-system dependent in the reference; sorted here)

    def init_img_names_seq_list(self):
        """data.py:120-196."""
        opts = self.opts
        self.img_list_src, self.img_list_trg, self.seq_id_list = [], [], []
        if self.dataset_variant == 'mview':
            self.img_list_src = self._pngs(os.path.join(self.root_dir, 'image_2'))
            self.seq_id_list = [int(n.split('/')[-1].split('_')[0]) for n in self.img_list_src]
        elif self.dataset_variant == 'odom':
            data_seq = {'train': list(range(0, 7)) + list(range(12, 21)), 'val': list(range(7, 9)), 'test': list(range(9, 11))}[opts.data_split]
            for seq_id in data_seq:
                for name in self._pngs(os.path.join(self.root_dir, '{:02d}'.format(seq_id), 'image_2')):
                    self.img_list_src.append(name)
                    self.seq_id_list.append(seq_id)
        elif self.dataset_variant == 'raw_city':
            exclude_img = '2011_09_26_drive_0117_sync/image_02/data/0000000074.png'
            for seq_id in split_sequences(opts.data_split):
                seq_date = seq_id[0:10]
                seq_dir = os.path.join(self.root_dir, seq_date, '{}_sync'.format(seq_id))
                for name in self._pngs(os.path.join(seq_dir, 'image_02')):
                    if exclude_img not in name:
                        self.img_list_src.append(name)
                        self.seq_id_list.append(seq_date)
        if self.dataset_variant == 'raw_city':
            self.img_list_trg = [f.replace('image_02', 'image_03') for f in self.img_list_src]
            if self.output_disparities:
                self.img_list_disp_src = []
                for im_name in self.img_list_src:
                    parts = im_name.split('/')
                    self.img_list_disp_src.append(os.path.join(self.root_dir, 'spss_stereo_results', parts[-4],
                                                               parts[-1][:-4] + '_left_initial_disparity.png'))
                self.img_list_disp_trg = [f.replace('left', 'right') for f in self.img_list_disp_src]
        else:
            self.img_list_trg = [f.replace('image_2', 'image_3') for f in self.img_list_src]

    def preload_calib_files(self):
        """data.py:198-225."""
        self.cam_calibration = {}
        if self.dataset_variant == 'mview':
            for root, _, filenames in os.walk(os.path.join(self.root_dir, 'calib_cam_to_cam')):
                for filename in fnmatch.filter(filenames, '*.txt'):
                    self.cam_calibration[int(filename.split('.txt')[0])] = read_calib_file(os.path.join(root, filename))
        elif self.dataset_variant == 'odom':
            for seq_id in range(22):
                cal = read_calib_file(os.path.join(self.root_dir, '{:02d}'.format(seq_id), 'calib.txt'))
                for key in ['P_rect_00', 'P_rect_01', 'P_rect_02', 'P_rect_03']:
                    cal[key] = np.copy(cal[key.replace('_rect_0', '')])
                self.cam_calibration[seq_id] = cal
        elif self.dataset_variant == 'raw_city':
            for seq_id in raw_city_sequences():
                seq_date = seq_id[0:10]
                self.cam_calibration[seq_date] = read_calib_file(os.path.join(self.root_dir, seq_date, 'calib_cam_to_cam.txt'))

    read_calib_file = staticmethod(read_calib_file)

    def define_queues(self, _threads=8, _prefetch=2):
        """data.py:268-301 -- the reference starts TF filename queues, reader ops and a shuffle batch of capacity; here a pool of
        decoder threads (PIL releases the GIL while inflating) that works `_prefetch` batches ahead of forward(), so decoding the
        next batches overlaps the GPU work on the current one.  Without this call forward() decodes synchronously."""
        import collections
        from concurrent.futures import ThreadPoolExecutor
        self._pool = ThreadPoolExecutor(max_workers=int(_threads))
        self._prefetch = int(_prefetch)
        self._queued = collections.deque()

    def forward_instance(self, img_src, img_trg, src_shape, trg_shape, calib_data):
        """data.py:303-342."""
        k_s, k_t, rot, trans = stereo_cameras(calib_data, src_shape, trg_shape, self.h, self.w)
        return (img_src, img_trg, k_s, k_t, rot, trans)

    @staticmethod
    def _decode(path):
        """PNG -> uint8 [H,W,C] (host)."""
        from PIL import Image
        img = np.asarray(Image.open(path))
        if img.ndim == 2:
            img = img[:, :, None]
        if img.dtype != np.uint8:
            # limitation: the SPS-stereo disparity PNGs of --kitti_dl_disparities are 16-bit; the resize kernel takes 8-bit data only
            raise RuntimeError('lsi_b200: %s is not an 8-bit image (16-bit disparity PNGs are not supported)' % path)
        return np.ascontiguousarray(img)

    def _schedule(self, bs):
        """Next bs samples of the epoch order: (ids, [decoded-image futures or arrays per sample])."""
        ids = [int(self._order[(self._cursor + b) % len(self._order)]) for b in range(bs)]
        self._cursor = (self._cursor + bs) % len(self._order)
        pool = getattr(self, '_pool', None)
        run = (lambda p: pool.submit(self._decode, p)) if pool is not None else (lambda p: self._decode(p))
        jobs = []
        for i in ids:
            paths = [self.img_list_src[i], self.img_list_trg[i]]
            if self.output_disparities:
                paths += [self.img_list_disp_src[i], self.img_list_disp_trg[i]]
            jobs.append([run(p) for p in paths])
        return ids, jobs

    def forward(self, bs):
        """data.py:344-390 -- (img_s, img_t, k_s, k_t, rot, trans[, disp_s, disp_t]) for the next bs samples of the epoch
        order, as CUDA tensors (images [bs,h,w,3] float32 in [0,1], cameras float32)."""
        if len(self.img_list_src) == 0:
            raise RuntimeError('lsi_b200: no KITTI images under %s' % self.root_dir)
        queued = getattr(self, '_queued', None)
        if queued is None:
            ids, jobs = self._schedule(bs)
        else:
            if queued and len(queued[0][0]) != bs:          # batch size changed: drop what was decoded ahead (and rewind the order)
                self._cursor = (self._cursor - sum(len(q[0]) for q in queued)) % len(self._order)
                queued.clear()
            while len(queued) < 1 + self._prefetch:
                queued.append(self._schedule(bs))
            ids, jobs = queued.popleft()
        self.src_image_names = [self.img_list_src[i] for i in ids]
        dev = torch.device('cuda')
        get = lambda j: j.result() if hasattr(j, 'result') else j
        imgs = [torch.empty(bs, self.h, self.w, 3, dtype=torch.float32, device=dev) for _ in range(2)]
        disps = [torch.empty(bs, self.h, self.w, 1, dtype=torch.float32, device=dev) for _ in range(2)] if self.output_disparities else []
        cams = [[] for _ in range(4)]
        for b, (i, job) in enumerate(zip(ids, jobs)):
            dec = [get(j) for j in job]
            area_resize(dec[0], self.h, self.w, 3, out=imgs[0][b])          # one launch per image (KITTI frames differ in size), written
            area_resize(dec[1], self.h, self.w, 3, out=imgs[1][b])          # straight into the batch tensor
            inst = self.forward_instance(None, None, dec[0].shape, dec[1].shape, self.cam_calibration[self.seq_id_list[i]])
            for c, v in zip(cams, inst[2:]):
                c.append(v)
            for k in range(len(disps)):
                area_resize(dec[2 + k], self.h, self.w, 1, out=disps[k][b])
        out = imgs + [torch.tensor(np.stack(c), dtype=torch.float32, device=dev) for c in cams] + disps
        return out
