"""Times lsi_b200_conv2d_wgrad_tc on the weight-gradient shapes of one training step (B = 8 per tower), halo-tile kernel
against the per-tap kernel (LSI_B200_WGRAD_HALO=0).  Usage: python tools/time_wgrad.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'layered-scene-inference_b200'))
from lsi import _b200  # noqa: E402
from lsi.nnutils.nets import same_pad  # noqa: E402

# (H, W of the strided-gather side, Ca, Cb, k, s, calls per step)
SHAPES = [
    (256, 896, 32, 32, 3, 1, 8), (256, 896, 32, 4, 3, 1, 8), (64, 224, 192, 128, 3, 1, 8), (128, 448, 96, 64, 3, 1, 8),
    (128 * 2, 448 * 2, 32, 64, 4, 2, 8), (128, 448, 64, 128, 4, 2, 8), (64, 224, 128, 128, 4, 2, 8),
    (128, 448, 32, 32, 7, 1, 2), (64, 224, 64, 64, 5, 1, 2), (32, 112, 256, 128, 3, 1, 2), (32, 112, 128, 128, 3, 1, 2),
    (16, 56, 512, 256, 3, 1, 2), (16, 56, 256, 256, 3, 1, 2), (8, 28, 1024, 512, 3, 1, 2), (8, 28, 512, 512, 3, 1, 2),
    (4, 14, 1024, 512, 3, 1, 2), (4, 14, 512, 512, 3, 1, 2), (2, 7, 512, 512, 3, 1, 2),
    (128, 448, 32, 64, 5, 2, 2), (64, 224, 64, 128, 3, 2, 2), (32, 112, 128, 256, 3, 2, 2), (16, 56, 256, 512, 3, 2, 2),
    (8, 28, 512, 512, 3, 2, 2), (4, 14, 512, 512, 3, 2, 2),
    (32, 112, 256, 128, 4, 2, 2), (16, 56, 512, 256, 4, 2, 2), (8, 28, 512, 512, 4, 2, 2), (4, 14, 512, 512, 4, 2, 2),
]


def run(B=8, iters=10):
    tot = {'1': 0.0, '0': 0.0}
    for (H, W, Ca, Cb, k, s, n) in SHAPES:
        Ho, Wo = -(-H // s), -(-W // s)
        pt, pl = (1, 1) if (k, s) == (4, 2) else (same_pad(H, k, s)[0], same_pad(W, k, s)[0])
        big = torch.randn(B, H, W, Ca, device='cuda')
        small = torch.randn(B, Ho, Wo, Cb, device='cuda')
        dw = torch.empty(k, k, Ca, Cb, device='cuda')
        d = _b200.ConvDesc(batch=B, h_in=H, w_in=W, c_in=Ca, h_out=Ho, w_out=Wo, c_out=Cb, kh=k, kw=k, stride=s, pad_top=pt,
                           pad_left=pl, mode=0, w_tap_stride=0, w_ci_stride=0, w_co_stride=0, in_c_stride=Ca, out_c_stride=Cb,
                           epilogue=0, accumulate=0)
        ms = {}
        for halo in ('1', '0'):
            os.environ['LSI_B200_WGRAD_HALO'] = halo
            for _ in range(2):
                _b200.call('lsi_b200_conv2d_wgrad_tc', d, _b200.ptr(big), _b200.ptr(small), _b200.ptr(dw), _b200.stream())
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                _b200.call('lsi_b200_conv2d_wgrad_tc', d, _b200.ptr(big), _b200.ptr(small), _b200.ptr(dw), _b200.stream())
            e1.record()
            torch.cuda.synchronize()
            ms[halo] = e0.elapsed_time(e1) / iters
            tot[halo] += ms[halo] * n
        flops = 2.0 * B * Ho * Wo * k * k * Ca * Cb
        print(f'{H}x{W} {Ca}->{Cb} k{k} s{s} x{n}: halo {ms["1"]*1e3:8.1f} us ({flops / ms["1"] / 1e9:6.1f} TF/s)   '
              f'per-tap {ms["0"]*1e3:8.1f} us   per step {ms["1"]*n:.3f} / {ms["0"]*n:.3f} ms', flush=True)
    print(f'total per step: halo path {tot["1"]:.2f} ms, per-tap {tot["0"]:.2f} ms')


if __name__ == '__main__':
    run()
