"""CPU restatement of the KITTI loader's arithmetic (TEST INFRASTRUCTURE: only tests/ may import this).
  * area_resize: [TF1.4] tf.image.resize_images(method=AREA) (core/kernels/resize_area_op.cc), restated -- TensorFlow itself
    is not installable here, so this op is pinned only for integer factors (box mean, the case the TF shim covers) and by
    its defining properties (partition of unity, exactness on constants and on linear ramps away from the border);
  * stereo_cameras: lsi/data/kitti/data.py:303-342, pinned by tests/golden/kitti_loader.npz (the reference's own function).
"""
import numpy as np


def _axis_weights(n_in, n_out):
    """[n_out, n_in] matrix of the covered fractions of every input cell, indices clamped to the image."""
    scale = n_in / float(n_out)
    w = np.zeros((n_out, n_in), dtype=np.float64)
    for y in range(n_out):
        lo, hi = y * scale, (y + 1) * scale
        for i in range(int(np.floor(lo)), int(np.ceil(hi))):
            if i < lo:
                f = (hi - lo) if i + 1 > hi else i + 1 - lo
            else:
                f = (hi - i) if i + 1 > hi else 1.0
            w[y, min(max(i, 0), n_in - 1)] += f
    return w / scale


def area_resize(img_u8, h, w, nc=3):
    """uint8 [H,W,C] -> float64 [h,w,nc] in [0,1]."""
    img = np.asarray(img_u8, dtype=np.float64)[:, :, :nc] / 255.0
    wy, wx = _axis_weights(img.shape[0], h), _axis_weights(img.shape[1], w)
    rows = np.tensordot(wy, img, axes=([1], [0]))            # [h, W, C]   (two separable passes)
    return np.einsum('xj,yjc->yxc', wx, rows)


def stereo_cameras(p_rect_02, p_rect_03, src_shape, trg_shape, h, w):
    """data.py:303-342."""
    p2, p3 = np.asarray(p_rect_02, dtype=np.float64).reshape(3, 4), np.asarray(p_rect_03, dtype=np.float64).reshape(3, 4)
    k_s, k_t = p2[:3, :3].copy(), p3[:3, :3].copy()
    ts, tt = p2[:, 3].copy(), p3[:, 3].copy()
    ts[0] = (ts[0] - k_s[0, 2] * ts[2]) / k_s[0, 0]; ts[1] = (ts[1] - k_s[1, 2] * ts[2]) / k_s[1, 1]
    tt[0] = (tt[0] - k_t[0, 2] * tt[2]) / k_t[0, 0]; tt[1] = (tt[1] - k_t[1, 2] * tt[2]) / k_t[1, 1]
    k_s[0] *= w / src_shape[1]; k_s[1] *= h / src_shape[0]
    k_t[0] *= w / trg_shape[1]; k_t[1] *= h / trg_shape[0]
    return k_s, k_t, np.eye(3), (tt - ts).reshape(3, 1)
