"""compute-sanitizer driver for the row-gather splat: a few small renders (packed / planar, mask, ds, wide rows, noisy lists).
    compute-sanitizer --tool memcheck|racecheck python tools/micro/rg_sanitize.py"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, 'layered-scene-inference_b200')); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
import test_gpu_rowgather as T
from lsi.geometry import ldi
from lsi.nnutils import helpers
for ci in (0, 1, 2, 3, 4, 7):
    L, B, H, W, ds, compose, use_mask, packed, noisy = T.CASES[ci]
    H = min(H, 4)
    sc = T._scene(L, 1, H, W, seed=ci, noisy=noisy)
    kw = dict(compose_layers=compose, trg_downsampling=ds, bg_layer_disp=1e-3, max_disp=0.4, zbuf_scale=50)
    got = T._run(ldi, helpers, sc, use_mask, packed, 0, **kw)
    ref = T._run(ldi, helpers, sc, use_mask, packed, 1, **kw)
    torch.cuda.synchronize()
    print('case', ci, 'max rel diff', max(float((a - b).abs().max() / b.abs().max()) for a, b in zip(got, ref)))
print('done')
