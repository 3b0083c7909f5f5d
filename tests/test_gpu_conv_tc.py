"""GPU: the tcgen05/TMA convolution (lsi_b200_conv2d_tc) against the fp32 CUDA-core kernel (lsi_b200_conv2d) that is
itself checked against the CPU oracle in tests/test_gpu_nets.py.  TF32 inputs, fp32 accumulation: tolerance 2e-3 of
the output scale."""
import ctypes

import pytest
import torch

from _util import rel_err

pytestmark = pytest.mark.gpu
TOL = 2e-3


def _run_both(B, Hi, Wi, Cin, Cout, k, s, mode, transposed_weights, split=None, epilogue=0, accumulate=0, seed=0, h16=False,
              out_f16=False):
    from lsi import _b200
    from lsi.nnutils.nets import same_pad
    lib = _b200.lib()
    torch.manual_seed(seed)
    dev = 'cuda'
    if mode == 0:
        Ho, Wo = -(-Hi // s), -(-Wi // s)
        pt, pl = same_pad(Hi, k, s)[0], same_pad(Wi, k, s)[0]
    else:
        Ho, Wo = Hi * s, Wi * s
        pt = pl = 1 if s == 2 else same_pad(Ho, k, 1)[0]
    x = torch.randn(B, Hi, Wi, Cin, device=dev)
    if h16:                    # fp16-stored activations: the reference sees the same rounded values
        x = x.half().float()
    if transposed_weights:     # [kh,kw,cout,cin]
        w = torch.randn(k, k, Cout, Cin, device=dev) / (k * k * Cin) ** 0.5
        ws = dict(w_tap_stride=Cin * Cout, w_ci_stride=1, w_co_stride=Cin)
    else:                      # [kh,kw,cin,cout]
        w = torch.randn(k, k, Cin, Cout, device=dev) / (k * k * Cin) ** 0.5
        ws = dict(w_tap_stride=Cin * Cout, w_ci_stride=Cout, w_co_stride=1)
    bias = torch.randn(Cout, device=dev)
    kw = dict(batch=B, h_in=Hi, w_in=Wi, c_in=Cin, h_out=Ho, w_out=Wo, c_out=Cout, kh=k, kw=k, stride=s, pad_top=pt,
              pad_left=pl, mode=mode, in_c_stride=Cin, out_c_stride=Cout, epilogue=epilogue, accumulate=accumulate, **ws)
    d = _b200.ConvDesc(**kw)
    init = torch.randn(B, Ho, Wo, Cout, device=dev)
    ref = init.clone()
    _b200.call('lsi_b200_conv2d', d, _b200.ptr(x), _b200.ptr(w), _b200.ptr(bias), _b200.ptr(ref), _b200.stream())
    out = init.clone()
    ca = Cin if split is None else split
    assert lib.lsi_b200_conv2d_tc_supported(d, ca) == 1
    nws = lib.lsi_b200_conv2d_tc_workspace_bytes(d)
    wsb = torch.empty(nws, dtype=torch.uint8, device=dev)
    if h16:
        stats = torch.zeros(Cout, 2, device=dev) if (epilogue == 0 and not accumulate) else None
        outh = torch.full((B, Ho, Wo, Cout), 7.0, device=dev, dtype=torch.float16) if out_f16 else out
        if split is None:
            _b200.call('lsi_b200_conv2d_tc_h', d, _b200.ptr(x.half()), Cin, None, 0, _b200.ptr(w), _b200.ptr(bias), _b200.ptr(outh),
                       int(out_f16), _b200.ptr(stats), 1e-3, _b200.ptr(wsb), nws, _b200.stream())
        else:
            xa, xb = x[..., :split].half().contiguous(), x[..., split:].half().contiguous()
            d2 = _b200.ConvDesc(**dict(kw, in_c_stride=split))
            _b200.call('lsi_b200_conv2d_tc_h', d2, _b200.ptr(xa), split, _b200.ptr(xb), Cin - split, _b200.ptr(w), _b200.ptr(bias),
                       _b200.ptr(outh), int(out_f16), _b200.ptr(stats), 1e-3, _b200.ptr(wsb), nws, _b200.stream())
        torch.cuda.synchronize()
        if stats is not None:
            mean = ref.mean(dim=(0, 1, 2)); var = ref.var(dim=(0, 1, 2), unbiased=False)
            assert rel_err(stats[:, 0].cpu(), mean.cpu()) < TOL
            assert rel_err(stats[:, 1].cpu(), torch.rsqrt(var + 1e-3).cpu()) < TOL
        return outh.float(), ref
    if split is None:
        _b200.call('lsi_b200_conv2d_tc', d, _b200.ptr(x), Cin, None, 0, _b200.ptr(w), _b200.ptr(bias), _b200.ptr(out),
                   _b200.ptr(wsb), nws, _b200.stream())
    else:                      # concat on the fly: two separately allocated sources
        xa, xb = x[..., :split].contiguous(), x[..., split:].contiguous()
        d2 = _b200.ConvDesc(**dict(kw, in_c_stride=split))
        _b200.call('lsi_b200_conv2d_tc', d2, _b200.ptr(xa), split, _b200.ptr(xb), Cin - split, _b200.ptr(w), _b200.ptr(bias),
                   _b200.ptr(out), _b200.ptr(wsb), nws, _b200.stream())
    torch.cuda.synchronize()
    return out, ref


@pytest.mark.parametrize('B,H,W,Cin,Cout,k', [
    (1, 8, 16, 32, 32, 3),        # exactly one tile
    (2, 16, 32, 64, 64, 3),
    (1, 11, 21, 32, 128, 3),      # ragged spatial size: partial tiles, zero-filled halo
    (2, 8, 16, 96, 64, 3),        # head layer upcnv2b
    (1, 8, 16, 192, 128, 3),      # head layer upcnv3b
    (1, 8, 16, 32, 4, 3),         # prediction conv: Cout padded to 16
    (1, 8, 16, 32, 32, 7),        # cnv1b
    (1, 8, 16, 64, 64, 5),        # cnv2b
    (1, 4, 6, 1024, 512, 3),      # icnv7: four N tiles
])
def test_conv_stride1(B, H, W, Cin, Cout, k):
    out, ref = _run_both(B, H, W, Cin, Cout, k, 1, 0, False)
    assert rel_err(out.cpu(), ref.cpu()) < TOL


def test_conv_concat_on_the_fly_bias_sigmoid_accumulate():
    out, ref = _run_both(2, 8, 16, 96, 64, 3, 1, 0, False, split=64)
    assert rel_err(out.cpu(), ref.cpu()) < TOL
    out, ref = _run_both(1, 8, 16, 32, 4, 3, 1, 0, False, epilogue=2)
    assert rel_err(out.cpu(), ref.cpu()) < TOL
    out, ref = _run_both(1, 8, 16, 64, 32, 3, 1, 0, False, accumulate=1)
    assert rel_err(out.cpu(), ref.cpu()) < TOL


@pytest.mark.parametrize('B,H,W,Cin,Cout,k', [
    (1, 32, 24, 32, 32, 3),       # x-merged halo path (output height >= 16): one 10-pixel halo row per ky feeds 3 taps
    (2, 21, 37, 96, 64, 3),       # ragged: partial 16x8 tiles
    (1, 16, 16, 192, 128, 3),     # 6 channel chunks, N = 128
    (1, 32, 16, 32, 4, 3),        # N padded to 16
    (1, 16, 40, 32, 32, 7),       # 7 taps per halo row (14-pixel halo)
    (1, 48, 16, 64, 64, 5),
])
def test_conv_stride1_xmerge(B, H, W, Cin, Cout, k):
    out, ref = _run_both(B, H, W, Cin, Cout, k, 1, 0, False)
    assert rel_err(out.cpu(), ref.cpu()) < TOL
    out, ref = _run_both(B, H, W, Cin, Cout, k, 1, 0, False, split=(32 if Cin >= 64 else None))
    assert rel_err(out.cpu(), ref.cpu()) < TOL


@pytest.mark.parametrize('B,H,W,Cin,Cout', [(1, 4, 8, 128, 64), (2, 8, 16, 64, 32), (1, 5, 9, 128, 128), (1, 16, 24, 64, 32),
                                            (2, 19, 11, 128, 64)])
def test_upconv_phase_decomposed(B, H, W, Cin, Cout):
    out, ref = _run_both(B, H, W, Cin, Cout, 4, 2, 1, True)
    assert rel_err(out.cpu(), ref.cpu()) < TOL


@pytest.mark.parametrize('B,H,W,Cin,Cout,k', [(1, 16, 32, 32, 64, 5), (2, 16, 32, 64, 128, 3), (1, 32, 64, 32, 32, 7)])
def test_conv_stride2_element_strides(B, H, W, Cin, Cout, k):
    out, ref = _run_both(B, H, W, Cin, Cout, k, 2, 0, False)
    assert rel_err(out.cpu(), ref.cpu()) < TOL


@pytest.mark.parametrize('B,H,W,Cin,Cout,k,s,mode,split,out_f16', [
    (1, 8, 16, 32, 32, 3, 1, 0, None, True),         # one tile, dense 64-byte rows
    (2, 21, 37, 96, 64, 3, 1, 0, 64, True),          # x-merged halo rows, ragged, concat on the fly (upcnv2b)
    (1, 16, 16, 192, 128, 3, 1, 0, 128, True),       # upcnv3b
    (1, 16, 40, 32, 32, 7, 1, 0, None, True),        # cnv1b: 7 taps per halo row
    (1, 48, 16, 64, 64, 5, 1, 0, None, False),       # fp32 output
    (1, 4, 6, 1024, 512, 3, 1, 0, 512, True),        # icnv7: four N tiles
    (2, 16, 32, 64, 128, 3, 2, 0, None, True),       # stride 2 via element strides
    (1, 32, 64, 32, 64, 5, 2, 0, None, True),
    (2, 19, 11, 128, 64, 4, 2, 1, None, True),       # phase-decomposed up-conv
    (1, 5, 9, 128, 128, 4, 2, 1, None, True),
    (1, 8, 16, 32, 4, 3, 1, 0, None, False),         # N padded to 16
])
def test_conv_fp16_operands(B, H, W, Cin, Cout, k, s, mode, split, out_f16):
    """lsi_b200_conv2d_tc_h (fp16 activations and weights, fp32 accumulation, statistics from the accumulators) against the
    fp32 kernel on the same fp16-rounded activations."""
    out, ref = _run_both(B, H, W, Cin, Cout, k, s, mode, mode == 1, split=split, h16=True, out_f16=out_f16)
    assert rel_err(out.cpu(), ref.cpu()) < (3e-3 if out_f16 else TOL)


@pytest.mark.parametrize('B,H,W,Ca,Cb,k,s', [
    (1, 8, 16, 32, 32, 3, 1),        # one pixel tile, 9 taps -> 3 M tiles (last one partial)
    (2, 16, 32, 64, 64, 3, 1),       # two B blocks
    (1, 11, 21, 96, 64, 3, 1),       # ragged spatial size
    (1, 8, 16, 32, 4, 3, 1),         # prediction conv: 4 output channels
    (2, 8, 16, 192, 128, 3, 1),      # two N tiles
    (1, 8, 16, 32, 32, 7, 1),
    (2, 16, 32, 32, 64, 5, 2),       # stride-2 conv: element strides on the big side
    (1, 16, 32, 64, 128, 3, 2),
    (2, 8, 16, 32, 64, 4, 2),        # transposed conv: big = dout (2H x 2W, 32 ch), small = input (H x W, 64 ch)
    (1, 8, 16, 40, 24, 3, 1),        # channel counts that are not multiples of 32
    (2, 40, 70, 32, 32, 3, 1),       # several pixel tiles per split, ragged
    (1, 24, 48, 64, 64, 5, 1),       # 5 taps per row: two M tiles per filter row, two accumulator groups
    (1, 16, 32, 32, 64, 7, 1),       # 14 M tiles in groups of 7
    (1, 16, 32, 64, 256, 3, 1),      # four B blocks per N tile, two N tiles
    (1, 4, 14, 128, 128, 3, 1),      # image narrower than the tile
])
@pytest.mark.parametrize('halo', [True, False])
def test_wgrad_tensor_core(B, H, W, Ca, Cb, k, s, halo, monkeypatch):
    """lsi_b200_conv2d_wgrad_tc against the fp32 weight-gradient kernel; sizes are those of the strided-gather side
    (H x W, Ca channels) like the descriptor's *_in fields.  halo: the stride-1 halo-tile kernel (default) / the per-tap kernel."""
    from lsi import _b200
    if not halo and s != 1:
        pytest.skip('stride-2 shapes always take the per-tap kernel')
    monkeypatch.setenv('LSI_B200_WGRAD_HALO', '1' if halo else '0')
    from lsi.nnutils.nets import same_pad
    torch.manual_seed(Ca + Cb + k)
    Ho, Wo = -(-H // s), -(-W // s)
    pt, pl = (1, 1) if (k, s) == (4, 2) else (same_pad(H, k, s)[0], same_pad(W, k, s)[0])
    big = torch.randn(B, H, W, Ca, device='cuda')
    small = torch.randn(B, Ho, Wo, Cb, device='cuda')
    d = _b200.ConvDesc(batch=B, h_in=H, w_in=W, c_in=Ca, h_out=Ho, w_out=Wo, c_out=Cb, kh=k, kw=k, stride=s, pad_top=pt,
                       pad_left=pl, mode=0, w_tap_stride=0, w_ci_stride=0, w_co_stride=0, in_c_stride=Ca, out_c_stride=Cb,
                       epilogue=0, accumulate=0)
    ref = torch.empty(k, k, Ca, Cb, device='cuda')
    out = torch.full((k, k, Ca, Cb), 7.0, device='cuda')
    _b200.call('lsi_b200_conv2d_wgrad', d, _b200.ptr(big), _b200.ptr(small), _b200.ptr(ref), _b200.stream())
    assert _b200.lib().lsi_b200_conv2d_wgrad_tc_supported(d) == 1
    _b200.call('lsi_b200_conv2d_wgrad_tc', d, _b200.ptr(big), _b200.ptr(small), _b200.ptr(out), _b200.stream())
    torch.cuda.synchronize()
    assert rel_err(out.cpu(), ref.cpu()) < TOL


@pytest.mark.parametrize('B,H,W,out_f16', [(1, 4, 128, True), (2, 37, 150, True), (1, 64, 256, False), (3, 128, 128, True)])
def test_stem_tensor_core(B, H, W, out_f16):
    """lsi_b200_conv2d_stem_tc (7x7 stride-2 conv 3 -> 32, im2col built in shared memory, kind::f16 MMAs) against the fp32
    direct kernel, statistics included; ragged sizes exercise the zero-filled staging and partial tiles."""
    from lsi import _b200
    from lsi.nnutils.nets import same_pad
    torch.manual_seed(B * 1000 + H + W)
    dev = 'cuda'
    Ho, Wo = -(-H // 2), -(-W // 2)
    x = torch.rand(B, H, W, 3, device=dev)
    w = torch.randn(7, 7, 3, 32, device=dev) / 147 ** 0.5
    d = _b200.ConvDesc(batch=B, h_in=H, w_in=W, c_in=3, h_out=Ho, w_out=Wo, c_out=32, kh=7, kw=7, stride=2, pad_top=same_pad(H, 7, 2)[0],
                       pad_left=same_pad(W, 7, 2)[0], mode=0, w_tap_stride=96, w_ci_stride=32, w_co_stride=1, in_c_stride=3,
                       out_c_stride=32, epilogue=0, accumulate=0)
    ref = torch.zeros(B, Ho, Wo, 32, device=dev)
    _b200.call('lsi_b200_conv2d', d, _b200.ptr(x), _b200.ptr(w), None, _b200.ptr(ref), _b200.stream())
    assert _b200.lib().lsi_b200_conv2d_stem_tc_supported(d) == 1
    out = torch.full((B, Ho, Wo, 32), 7.0, device=dev, dtype=torch.float16 if out_f16 else torch.float32)
    st = torch.zeros(32, 2, device=dev)
    nws = _b200.lib().lsi_b200_conv2d_stem_tc_workspace_bytes()
    wsb = torch.empty(nws, dtype=torch.uint8, device=dev)
    for _ in range(2):     # twice: persistent state (barrier phases, TMEM) is per launch
        _b200.call('lsi_b200_conv2d_stem_tc', d, _b200.ptr(x), _b200.ptr(w), _b200.ptr(out), int(out_f16), _b200.ptr(st), 1e-3,
                   _b200.ptr(wsb), nws, _b200.stream())
    torch.cuda.synchronize()
    assert rel_err(out.float().cpu(), ref.cpu()) < (3e-3 if out_f16 else TOL)
    mean = ref.mean(dim=(0, 1, 2)); var = ref.var(dim=(0, 1, 2), unbiased=False)
    assert rel_err(st[:, 0].cpu(), mean.cpu()) < TOL
    assert rel_err(st[:, 1].cpu(), torch.rsqrt(var + 1e-3).cpu()) < TOL


@pytest.mark.parametrize('B,H,W,Cin,Cout,k,s,mode', [
    (5, 2, 7, 64, 128, 3, 1, 0),       # 2x7 images: 8 per tile, ragged batch
    (11, 2, 7, 512, 512, 3, 1, 0),     # cnv7b
    (3, 4, 14, 128, 64, 3, 1, 0),      # 4x14: 2 images per tile, odd batch
    (5, 8, 28, 64, 64, 3, 2, 0),       # stride 2 down to 4x14 (element strides + batched tile)
    (3, 4, 14, 96, 128, 3, 2, 0),      # stride 2 down to 2x7
    (3, 2, 7, 128, 64, 4, 2, 1),       # up-conv 2x7 -> 4x14: phase space 2x7
    (9, 1, 4, 64, 32, 4, 2, 1),
    (4, 3, 5, 32, 32, 3, 1, 0),        # height not a power of two
])
def test_small_images_share_a_tile(B, H, W, Cin, Cout, k, s, mode):
    """Images of at most 16 pixels across are batched into one 128-row tile (TMA box over several images): TF32 and fp16
    operand modes against the fp32 kernel."""
    out, ref = _run_both(B, H, W, Cin, Cout, k, s, mode, mode == 1)
    assert rel_err(out.cpu(), ref.cpu()) < TOL
    out, ref = _run_both(B, H, W, Cin, Cout, k, s, mode, mode == 1, h16=True, out_f16=True)
    assert rel_err(out.cpu(), ref.cpu()) < 3e-3
