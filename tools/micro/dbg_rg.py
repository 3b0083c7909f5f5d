import sys, os
sys.path.insert(0, '/root/repo/layered-scene-inference_b200'); sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import torch
import test_gpu_rowgather as T
from lsi.geometry import ldi
from lsi.nnutils import helpers
for ci in (4, 5):
    case = T.CASES[ci]
    L, B, H, W, ds, compose, use_mask, packed, noisy = case
    sc = T._scene(L, B, H, W, seed=hash(case) % 1000, noisy=noisy)
    kw = dict(compose_layers=compose, trg_downsampling=ds, bg_layer_disp=1e-3, max_disp=0.4, zbuf_scale=50)
    ref = T._run(ldi, helpers, sc, use_mask, packed, 1, **kw)
    for rep in range(3):
        got = T._run(ldi, helpers, sc, use_mask, packed, 0, **kw)
        for name, a, b in zip(('img', 'wts'), got, ref):
            d = ((a - b).abs() / b.abs().clamp_min(1e-3)).amax(dim=-1)   # [nl, B, H, W]
            bad = (d > 1e-3).nonzero()
            print('case', ci, 'rep', rep, name, 'bad', bad.shape[0], 'of', d.numel(), bad[:12].tolist())
