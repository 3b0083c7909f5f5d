"""ncu driver: forward_splat on the LDI a random-init CNN predicts (noisy disparities: the condition the splat kernel
meets inside bench.py), at a reduced batch.    ncu ... python tools/profile_render_noisy.py [--batch 14]"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'layered-scene-inference_b200')); sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench
from oracle import gen_inputs
from lsi.geometry import ldi
from lsi.nnutils import helpers, nets, train_utils
ap = argparse.ArgumentParser(); ap.add_argument('--batch', type=int, default=14); ap.add_argument('--iters', type=int, default=3)
a = ap.parse_args()
B = a.batch
host = bench.make_inputs(B, 0)
rs = np.random.RandomState(100)
img = torch.tensor(np.stack([gen_inputs.band_limited(rs, (bench.H, bench.W), 3) for _ in range(B)]).astype(np.float32), device='cuda')
opts = train_utils.default_opts(dataset='kitti', n_layers=bench.L, batch_size=B, img_height=bench.H, img_width=bench.W, zbuf_scale=bench.ZBUF_SCALE)
store = nets.ParamStore(seed=0)
with torch.no_grad():
    train_utils.predict_ldi(img[:1], opts, store, reuse=False)
    ldi_pred = train_utils.predict_ldi(img, opts, store, reuse=True)
cam = [torch.tensor(host[k], device='cuda') for k in ('k_s', 'k_t', 'rot', 't')]
pc = helpers.pixel_coords(B, bench.H, bench.W)
d = ldi_pred[2]
print('disp stats: mean %.4f std %.4f; mean |dx| %.5f (in target px: %.3f)' % (float(d.mean()), float(d.std()), float((d[:, :, :, 1:] - d[:, :, :, :-1]).abs().mean()),
      float((d[:, :, :, 1:] - d[:, :, :, :-1]).abs().mean()) * 257.0))
import ctypes
from lsi import _b200
lib = _b200.lib()
lib.lsi_b200_kernel_timing_enable(1)
for _ in range(a.iters):
    with torch.no_grad():
        ldi.forward_splat(tuple(ldi_pred), pc, *cam, compose_layers=True, trg_downsampling=1.0, bg_layer_disp=bench.BG_DISP,
                          max_disp=bench.MAX_DISP, zbuf_scale=bench.ZBUF_SCALE)
torch.cuda.synchronize()
kms, kn = (ctypes.c_double * 8)(), (ctypes.c_int * 8)()
_b200.call('lsi_b200_kernel_timing_collect', ctypes.cast(kms, ctypes.c_void_p), ctypes.cast(kn, ctypes.c_void_p))
lib.lsi_b200_kernel_timing_enable(0)
gb = 4.0 * 4 * bench.L * bench.H * bench.W * B / 1e9
print('noisy LDI, B=%d: splat %.4f ms/call (%.0f GB/s), normalize %.4f ms/call' % (B, kms[0] / a.iters, gb / (kms[0] / a.iters * 1e-3), kms[1] / a.iters))
print('done')
