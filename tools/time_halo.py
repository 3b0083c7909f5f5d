"""Times the three full-resolution head layers (upcnv1, upcnv1b, pred; nets.py:87-159) through lsi_b200_conv2d_halo at the
bench shape (B=64, 256x896), exactly as the inference pipeline calls them: producer's batch norm applied on load,
own batch statistics reduced in the epilogue.  Prints ms, algorithmic GB/s (input once + output once) and TFLOP/s."""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'layered-scene-inference_b200')); sys.path.insert(0, ROOT)
import torch
from lsi import _b200

ap = argparse.ArgumentParser()
ap.add_argument('--batch', type=int, default=64); ap.add_argument('--h', type=int, default=256); ap.add_argument('--w', type=int, default=896)
ap.add_argument('--iters', type=int, default=10); ap.add_argument('--no_bn', action='store_true'); ap.add_argument('--no_stats', action='store_true')
ap.add_argument('--only', type=str, default='')
a = ap.parse_args()
lib = _b200.lib()
B, H, W = a.batch, a.h, a.w
layers = [
    ('upcnv2  128->64 4x4/2 up', dict(h_in=H // 4, w_in=W // 4, c_in=128, h_out=H // 2, w_out=W // 2, c_out=64, kh=4, kw=4, stride=2, pad_top=1, pad_left=1, mode=1,
                                      w_tap_stride=128 * 64, w_ci_stride=1, w_co_stride=128, in_c_stride=128, out_c_stride=64, epilogue=0), True),
    ('upcnv1  64->32 4x4/2 up', dict(h_in=H // 2, w_in=W // 2, c_in=64, h_out=H, w_out=W, c_out=32, kh=4, kw=4, stride=2, pad_top=1, pad_left=1, mode=1,
                                     w_tap_stride=64 * 32, w_ci_stride=1, w_co_stride=64, in_c_stride=64, out_c_stride=32, epilogue=0), True),
    ('upcnv1b 32->32 3x3', dict(h_in=H, w_in=W, c_in=32, h_out=H, w_out=W, c_out=32, kh=3, kw=3, stride=1, pad_top=1, pad_left=1, mode=0,
                                w_tap_stride=32 * 32, w_ci_stride=32, w_co_stride=1, in_c_stride=32, out_c_stride=32, epilogue=0), True),
    ('pred    32->4  3x3 sigmoid', dict(h_in=H, w_in=W, c_in=32, h_out=H, w_out=832, c_out=4, kh=3, kw=3, stride=1, pad_top=1, pad_left=1, mode=0,
                                        w_tap_stride=32 * 4, w_ci_stride=4, w_co_stride=1, in_c_stride=32, out_c_stride=4, epilogue=2), False),
]
for name, kw, has_stats in layers:
    if a.only and a.only not in name:
        continue
    d = _b200.ConvDesc(batch=B, accumulate=0, **kw)
    assert lib.lsi_b200_conv2d_halo_supported(d) == 1
    x = torch.randn(B, d.h_in, d.w_in, d.c_in, device='cuda')
    w = torch.randn(d.kh * d.kw * d.c_in * d.c_out, device='cuda') * 0.05
    bias = torch.zeros(64, device='cuda')
    out = torch.empty(B, d.h_out, d.w_out, d.c_out, device='cuda')
    in_stats = None if a.no_bn else torch.stack([torch.zeros(d.c_in), torch.ones(d.c_in)], dim=1).contiguous().cuda()
    beta = None if a.no_bn else torch.zeros(d.c_in, device='cuda')
    st = torch.empty(d.c_out, 2, device='cuda') if (has_stats and not a.no_stats) else None
    nws = lib.lsi_b200_conv2d_halo_workspace_bytes(d)
    ws = torch.empty(nws, dtype=torch.uint8, device='cuda')

    def run():
        _b200.call('lsi_b200_conv2d_halo', d, _b200.ptr(x), _b200.ptr(in_stats), _b200.ptr(beta), _b200.ptr(w), _b200.ptr(bias),
                   None, _b200.ptr(out), _b200.ptr(st), 1e-3, _b200.ptr(ws), nws, _b200.stream())
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    _b200.lib().lsi_b200_kernel_timing_enable(1)
    for _ in range(a.iters):
        run()
    torch.cuda.synchronize()
    import ctypes
    kms, kn = (ctypes.c_double * 8)(), (ctypes.c_int * 8)()
    _b200.call('lsi_b200_kernel_timing_collect', ctypes.cast(kms, ctypes.c_void_p), ctypes.cast(kn, ctypes.c_void_p))
    _b200.lib().lsi_b200_kernel_timing_enable(0)
    ms = kms[4] / a.iters
    by = 4.0 * B * (d.h_in * d.w_in * d.c_in + d.h_out * d.w_out * d.c_out)
    taps = d.kh * d.kw / (d.stride * d.stride if d.mode == 1 else 1)
    fl = 2.0 * B * d.h_out * d.w_out * d.c_out * d.c_in * taps
    print('%-28s bn_in=%d stats=%d  %.3f ms  %6.0f GB/s (alg)  %6.1f TFLOP/s   [CTAS=%s STAGES=%s]' % (
        name, in_stats is not None, st is not None, ms, by / ms / 1e6, fl / ms / 1e9,
        os.environ.get('LSI_B200_HALO_CTAS', '-'), os.environ.get('LSI_B200_HALO_STAGES', '-')))
    del x, out
