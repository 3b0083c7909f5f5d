"""Times the CNN slice (U-Net trunk + L heads) forward and forward+backward on one GPU."""
import argparse, ctypes, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'layered-scene-inference_b200')); sys.path.insert(0, ROOT)
import torch
from lsi.nnutils import nets
ap = argparse.ArgumentParser()
ap.add_argument('--batch', type=int, default=8); ap.add_argument('--h', type=int, default=256); ap.add_argument('--w', type=int, default=896)
ap.add_argument('--layers', type=int, default=4); ap.add_argument('--iters', type=int, default=3); ap.add_argument('--backward', action='store_true')
a = ap.parse_args()
store = nets.ParamStore()
img = torch.rand(a.batch, a.h, a.w, 3, device='cuda')
def fwd(reuse):
    _, fd, sk, _ = nets.encoder_decoder_unet(img, nl_diff_enc_dec=3, reuse=reuse, _store=store)
    return nets.ldi_predictor(fd, n_layers=a.layers, reuse=reuse, n_layerwise_steps=3, skip_feat=sk, _store=store)
with torch.no_grad():
    fwd(False)
torch.cuda.synchronize()
def flops_per_image(h, w, L):   # forward 2*MAC, SURVEY.md appendix B scaled from 256x768
    return (21.8e9 + 23.0e9 * L) * (h * w) / (256.0 * 768.0)
from lsi import _b200
for mode in (['fwd', 'fwd+bwd'] if a.backward else ['fwd']):
    _b200.lib().lsi_b200_kernel_timing_enable(1)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(a.iters):
        if mode == 'fwd':
            with torch.no_grad():
                fwd(True)
        else:
            tex, m, d = fwd(True)
            (tex.sum() + d.sum()).backward()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / a.iters
    kms, kn = (ctypes.c_double * 8)(), (ctypes.c_int * 8)()
    _b200.call('lsi_b200_kernel_timing_collect', ctypes.cast(kms, ctypes.c_void_p), ctypes.cast(kn, ctypes.c_void_p))
    _b200.lib().lsi_b200_kernel_timing_enable(0)
    print('   conv_tc %.2f ms, conv_fp32 %.2f ms, wgrad %.2f ms per step' % (kms[4] / a.iters, kms[5] / a.iters, kms[6] / a.iters))
    fl = flops_per_image(a.h, a.w, a.layers) * a.batch * (3 if mode != 'fwd' else 1)
    print('%s: %.1f ms/step, %.1f views/s, %.2f TFLOP/s (B=%d %dx%d L=%d), peak mem %.1f GB' % (mode, dt * 1e3, a.batch / dt, fl / dt / 1e12, a.batch, a.h, a.w, a.layers, torch.cuda.max_memory_allocated() / 1e9))
