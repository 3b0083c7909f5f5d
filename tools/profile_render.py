"""Short driver for ncu captures: a few forward (and optionally backward) renders at the headline configuration.
    ncu ... python tools/profile_render.py [--batch 16] [--iters 3] [--backward] [--ds 1.0]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'layered-scene-inference_b200'))
sys.path.insert(0, ROOT)
import torch

import bench
from lsi.geometry import ldi
from lsi.nnutils import helpers

ap = argparse.ArgumentParser()
ap.add_argument('--batch', type=int, default=16)
ap.add_argument('--iters', type=int, default=3)
ap.add_argument('--backward', action='store_true')
ap.add_argument('--ds', type=float, default=1.0)
ap.add_argument('--variant', type=int, default=0)
ap.add_argument('--packed', action='store_true')
a = ap.parse_args()
B = a.batch
host = bench.make_inputs(B, 0)
tex = torch.tensor(host['tex'], device='cuda').requires_grad_(a.backward)
disp = torch.tensor(host['disp'], device='cuda').requires_grad_(a.backward)
if a.packed:
    pk = torch.cat([tex, disp], dim=-1)
    tex, disp = pk[..., :3], pk[..., 3:]
masks = torch.ones(bench.L, B, bench.H, bench.W, 1, device='cuda')
masks._lsi_all_ones = True
cam = [torch.tensor(host[k], device='cuda') for k in ('k_s', 'k_t', 'rot', 't')]
pc = helpers.pixel_coords(B, bench.H, bench.W)
for _ in range(a.iters):
    img, wts = ldi.forward_splat((tex, masks, disp), pc, *cam, compose_layers=True, trg_downsampling=a.ds,
                                 bg_layer_disp=bench.BG_DISP, max_disp=bench.MAX_DISP, zbuf_scale=bench.ZBUF_SCALE,
                                 _variant=a.variant)
    if a.backward:
        img.sum().backward()
torch.cuda.synchronize()
print('done')
