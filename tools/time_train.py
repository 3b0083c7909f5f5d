"""Times one training step (two towers + view-synthesis loss + backward + Adam) and prints the per-kernel-kind split."""
import argparse, ctypes, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'layered-scene-inference_b200')); sys.path.insert(0, ROOT)
import numpy as np, torch
from lsi import _b200
from lsi.nnutils import nets, train_utils
from oracle import gen_inputs
ap = argparse.ArgumentParser()
ap.add_argument('--batch', type=int, default=4); ap.add_argument('--h', type=int, default=256); ap.add_argument('--w', type=int, default=768)
ap.add_argument('--layers', type=int, default=2); ap.add_argument('--iters', type=int, default=3)
a = ap.parse_args()
B, H, W, L = a.batch, a.h, a.w, a.layers
opts = train_utils.default_opts(dataset='kitti', n_layers=L, batch_size=B, img_height=H, img_width=W)
s = gen_inputs.scene(1, B, H, W, 'kitti', 0, 0.4)
rs = np.random.RandomState(0)
batch = dict(imgs_src=rs.uniform(0, 1, (B, H, W, 3)).astype(np.float32), imgs_trg=rs.uniform(0, 1, (B, H, W, 3)).astype(np.float32),
             k_s=s['k_s'], k_t=s['k_t'], rot_mat=s['rot'], trans_mat=s['t'])
gb = {k: torch.tensor(v, device='cuda') for k, v in batch.items()}
tr = train_utils.Trainer(opts, store=nets.ParamStore(seed=0))
for _ in range(2): tr.train_step(gb)
torch.cuda.synchronize()
lib = _b200.lib(); lib.lsi_b200_kernel_timing_enable(1)
t0 = time.perf_counter()
for _ in range(a.iters): loss, _ = tr.train_step(gb)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / a.iters
kms, kn = (ctypes.c_double * 8)(), (ctypes.c_int * 8)()
_b200.call('lsi_b200_kernel_timing_collect', ctypes.cast(kms, ctypes.c_void_p), ctypes.cast(kn, ctypes.c_void_p))
names = ['splat_fwd', 'normalize', 'splat_bwd_target', 'splat_bwd_source', 'conv_tc', 'conv_fp32', 'wgrad', 'other']
print('train step: %.1f ms (B=%d %dx%d L=%d) = %.1f image pairs/s, loss %.4f, peak mem %.1f GB' % (dt * 1e3, B, H, W, L, B / dt, loss.item(), torch.cuda.max_memory_allocated() / 1e9))
for n, m, c in zip(names, kms, kn):
    if c: print('  %-18s %8.2f ms/step  (%d launches/step)' % (n, m / a.iters, c // a.iters))
