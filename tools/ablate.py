"""Ablation timings of the forward splat kernel (measurement aid): variant 0 = real kernel, 1 = plain atomic kernel,
3 = ring-less reduction kernel, 101 = no reductions, 102 = loads + one coalesced reduction per pixel (both on the
ring-less kernel).  ABLATE_PACKED=1 uses the packed head-output layout."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'layered-scene-inference_b200')); sys.path.insert(0, ROOT)
import torch
import bench
from lsi import _b200
from lsi.geometry import ldi
from lsi.nnutils import helpers
B = 64
host = bench.make_inputs(B, 0)
tex = torch.tensor(host['tex'], device='cuda'); disp = torch.tensor(host['disp'], device='cuda')
masks = torch.ones(bench.L, B, bench.H, bench.W, 1, device='cuda'); masks._lsi_all_ones = True
cam = [torch.tensor(host[k], device='cuda') for k in ('k_s', 'k_t', 'rot', 't')]
pc = helpers.pixel_coords(B, bench.H, bench.W)
lib = _b200.lib()
variants = [int(v) for v in sys.argv[1:]] or [0, 3, 1, 101, 102]
if os.environ.get('ABLATE_PACKED'):   # the head-output layout [L,B,H,W,4] = (r,g,b,disp)
    pk = torch.cat([tex, disp], dim=-1)
    tex, disp = pk[..., :3], pk[..., 3:]
for ds in (1.0, 0.5):
    for v in variants:
        def step():
            with torch.no_grad():
                ldi.forward_splat((tex, masks, disp), pc, *cam, compose_layers=True, trg_downsampling=ds, bg_layer_disp=bench.BG_DISP,
                                  max_disp=bench.MAX_DISP, zbuf_scale=bench.ZBUF_SCALE, _variant=v)
        for _ in range(3): step()
        torch.cuda.synchronize()
        lib.lsi_b200_kernel_timing_enable(1)
        for _ in range(10): step()
        torch.cuda.synchronize()
        kms, kn = (ctypes.c_double * 8)(), (ctypes.c_int * 8)()
        _b200.call('lsi_b200_kernel_timing_collect', ctypes.cast(kms, ctypes.c_void_p), ctypes.cast(kn, ctypes.c_void_p))
        lib.lsi_b200_kernel_timing_enable(0)
        gb = 4.0 * 4 * bench.L * bench.H * bench.W * B / 1e9
        print('ds=%.2f variant=%3d splat %.4f ms/step (%.0f GB/s)  normalize %.4f ms/step' % (ds, v, kms[0] / 10, gb / (kms[0] / 10 * 1e-3), kms[1] / 10))
