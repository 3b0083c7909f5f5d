"""Per-call timing of one training step (BASELINE config 4's per-GPU shard by default): every C-ABI call is bracketed by CUDA
events on the launching stream and aggregated by (entry point, layer shape); the gap to the step's wall time is torch glue
(allocations, autograd bookkeeping, the few eager ops of the loss)."""
import argparse, collections, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'layered-scene-inference_b200')); sys.path.insert(0, ROOT)
import numpy as np, torch
from lsi import _b200
from lsi.nnutils import nets, train_utils
from oracle import gen_inputs

ap = argparse.ArgumentParser()
ap.add_argument('--batch', type=int, default=8); ap.add_argument('--h', type=int, default=256); ap.add_argument('--w', type=int, default=832)
ap.add_argument('--layers', type=int, default=4); ap.add_argument('--iters', type=int, default=2); ap.add_argument('--mode', default='tf32')
a = ap.parse_args()
nets.set_conv_mode(a.mode)
B, H, W, L = a.batch, a.h, a.w, a.layers
opts = train_utils.default_opts(dataset='kitti', n_layers=L, batch_size=B, img_height=H, img_width=W)
s = gen_inputs.scene(1, B, H, W, 'kitti', 0, 0.4)
rs = np.random.RandomState(0)
gb = {k: torch.tensor(v, device='cuda') for k, v in dict(
    imgs_src=rs.uniform(0, 1, (B, H, W, 3)).astype(np.float32), imgs_trg=rs.uniform(0, 1, (B, H, W, 3)).astype(np.float32),
    k_s=s['k_s'], k_t=s['k_t'], rot_mat=s['rot'], trans_mat=s['t']).items()}
tr = train_utils.Trainer(opts, store=nets.ParamStore(seed=0))
for _ in range(2):
    tr.train_step(gb)
torch.cuda.synchronize()
records = []
orig = _b200.call


def timed(name, *args):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); orig(name, *args); e1.record()
    key = name.replace('lsi_b200_', '')
    if args and isinstance(args[0], _b200.ConvDesc):
        d = args[0]
        key = '%s %dx%d %d->%d k%d s%d m%d' % (key, d.h_in, d.w_in, d.c_in, d.c_out, d.kh, d.stride, d.mode)
    records.append((key, e0, e1))


_b200.call = timed
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0.record()
for _ in range(a.iters):
    tr.train_step(gb)
t1.record()
torch.cuda.synchronize()
agg = collections.OrderedDict()
for key, e0, e1 in records:
    v = agg.setdefault(key, [0.0, 0]); v[0] += e0.elapsed_time(e1); v[1] += 1
wall = t0.elapsed_time(t1) / a.iters
tot = sum(v[0] for v in agg.values()) / a.iters
print('train step B=%d %dx%d L=%d mode=%s: %.2f ms/step wall, %.2f ms in C-ABI calls (%d calls/step), peak mem %.1f GB'
      % (B, H, W, L, a.mode, wall, tot, len(records) // a.iters, torch.cuda.max_memory_allocated() / 1e9))
by_entry = collections.OrderedDict()
for key, v in agg.items():
    e = key.split(' ')[0]
    w = by_entry.setdefault(e, [0.0, 0]); w[0] += v[0]; w[1] += v[1]
print('--- by entry point')
for e, v in sorted(by_entry.items(), key=lambda kv: -kv[1][0]):
    print('%-34s %5d %9.3f ms/step  %5.1f %%' % (e, v[1] // a.iters, v[0] / a.iters, 100 * v[0] / a.iters / wall))
print('--- top calls')
for key, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:40]:
    print('%-70s %5d %9.3f' % (key[:70], v[1] // a.iters, v[0] / a.iters))
