"""GPU: the halo-tile tensor-core convolution (lsi_b200_conv2d_halo: resident weights, one TMA halo box per tile,
producer's batch norm + ReLU applied on load, batch statistics + coalesced stores in the epilogue) against the fp32
CUDA-core kernel (lsi_b200_conv2d) fed with the explicitly normalised input.  TF32 inputs, fp32 accumulation:
tolerance 2e-3 of the output scale (same bar as tests/test_gpu_conv_tc.py)."""
import pytest
import torch

from _util import rel_err

pytestmark = pytest.mark.gpu
TOL = 2e-3
EPS = 1e-3


def _run(B, Hi, Wi, Cin, Cout, k, mode, bn_in, stats_out, epilogue=0, out_hw=None, seed=0, in_f16=False, out_f16=False):
    from lsi import _b200
    from lsi.nnutils.nets import same_pad
    lib = _b200.lib()
    torch.manual_seed(seed)
    dev = 'cuda'
    if mode == 0:
        Ho, Wo = (Hi, Wi) if out_hw is None else out_hw
        s, pt, pl = 1, same_pad(Hi, k, 1)[0], same_pad(Wi, k, 1)[0]
        w = torch.randn(k, k, Cin, Cout, device=dev) / (k * k * Cin) ** 0.5
        ws = dict(w_tap_stride=Cin * Cout, w_ci_stride=Cout, w_co_stride=1)
    else:                      # 4x4 stride-2 up-convolution, weights [kh,kw,cout,cin]
        assert k == 4
        Ho, Wo, s, pt, pl = 2 * Hi, 2 * Wi, 2, 1, 1
        w = torch.randn(k, k, Cout, Cin, device=dev) / (4 * Cin) ** 0.5
        ws = dict(w_tap_stride=Cin * Cout, w_ci_stride=1, w_co_stride=Cin)
    x = torch.randn(B, Hi, Wi, Cin, device=dev) * 1.7 + 0.3
    bias = torch.randn(max(Cout, 4), device=dev)
    kw = dict(batch=B, h_in=Hi, w_in=Wi, c_in=Cin, h_out=Ho, w_out=Wo, c_out=Cout, kh=k, kw=k, stride=s, pad_top=pt,
              pad_left=pl, mode=mode, in_c_stride=Cin, out_c_stride=Cout, epilogue=epilogue, accumulate=0, **ws)
    d = _b200.ConvDesc(**kw)
    assert lib.lsi_b200_conv2d_halo_supported(d) == 1
    if in_f16:                 # the producer stored its raw output as fp16: the reference sees the same rounded values
        x = x.half().float()
    if bn_in:
        mean = x.mean(dim=(0, 1, 2)); var = x.var(dim=(0, 1, 2), unbiased=False)
        in_stats = torch.stack([mean, torch.rsqrt(var + EPS)], dim=1).contiguous()
        beta = torch.randn(Cin, device=dev) * 0.5
        xn = torch.relu((x - mean) * in_stats[:, 1] + beta).contiguous()
    else:
        in_stats = beta = None
        xn = x
    ref = torch.zeros(B, Ho, Wo, Cout, device=dev)
    _b200.call('lsi_b200_conv2d', d, _b200.ptr(xn), _b200.ptr(w), _b200.ptr(bias), _b200.ptr(ref), _b200.stream())
    out = torch.full((B, Ho, Wo, Cout), 7.0, device=dev, dtype=torch.float16 if out_f16 else torch.float32)
    st = torch.zeros(Cout, 2, device=dev) if stats_out else None
    nws = lib.lsi_b200_conv2d_halo_workspace_bytes(d)
    wsb = torch.empty(nws, dtype=torch.uint8, device=dev)
    if in_f16 or out_f16:
        xin = x.half() if in_f16 else x
        _b200.call('lsi_b200_conv2d_halo_h', d, _b200.ptr(xin), None, Cin, 0, int(in_f16), _b200.ptr(in_stats), _b200.ptr(beta),
                   _b200.ptr(w), _b200.ptr(bias), None, _b200.ptr(out), int(out_f16), _b200.ptr(st), EPS, _b200.ptr(wsb), nws,
                   _b200.stream())
    else:
        _b200.call('lsi_b200_conv2d_halo', d, _b200.ptr(x), _b200.ptr(in_stats), _b200.ptr(beta), _b200.ptr(w), _b200.ptr(bias),
                   None, _b200.ptr(out), _b200.ptr(st), EPS, _b200.ptr(wsb), nws, _b200.stream())
    torch.cuda.synchronize()
    return out.float(), ref, st


def _check(out, ref, st):
    assert rel_err(out.cpu(), ref.cpu()) < TOL
    if st is not None:
        mean = ref.mean(dim=(0, 1, 2)); var = ref.var(dim=(0, 1, 2), unbiased=False)
        assert rel_err(st[:, 0].cpu(), mean.cpu()) < TOL
        assert rel_err(st[:, 1].cpu(), torch.rsqrt(var + EPS).cpu()) < TOL


@pytest.mark.parametrize('B,H,W,bn_in', [
    (1, 16, 8, False),        # exactly one tile
    (1, 16, 8, True),
    (2, 21, 37, True),        # ragged: partial tiles on both edges
    (1, 48, 40, False),
    (4, 128, 160, True),      # 640 tiles: several tiles per persistent CTA, the ring wraps
])
def test_conv3x3_32to32(B, H, W, bn_in):
    _check(*_run(B, H, W, 32, 32, 3, 0, bn_in, True))


@pytest.mark.parametrize('B,H,W,Cin,Cout,bn_in', [
    (1, 8, 4, 64, 32, False),     # one tile per phase
    (2, 16, 24, 64, 32, True),    # head layer upcnv1
    (1, 19, 11, 64, 32, True),    # ragged
    (3, 64, 112, 64, 32, True),   # many tiles per CTA
    (1, 16, 16, 32, 32, True),
    (1, 24, 16, 64, 64, True),
    (2, 32, 24, 128, 64, True),   # head layer upcnv2: four chunk stages per tile, 128 KB of resident weights
    (1, 16, 8, 128, 64, False),
])
def test_upconv_4x4_s2(B, H, W, Cin, Cout, bn_in):
    _check(*_run(B, H, W, Cin, Cout, 4, 1, bn_in, True))


@pytest.mark.parametrize('B,H,W,out_hw,bn_in', [
    (1, 16, 8, None, False),
    (2, 32, 48, (32, 40), True),      # crop fused into the prediction conv
    (2, 128, 128, (128, 104), True),
])
def test_prediction_conv_bias_sigmoid(B, H, W, out_hw, bn_in):
    out, ref, _ = _run(B, H, W, 32, 4, 3, 0, bn_in, False, epilogue=2, out_hw=out_hw)
    assert rel_err(out.cpu(), ref.cpu()) < TOL


@pytest.mark.parametrize('B,H,W,Cin,Cout,k,mode,in_f16,out_f16', [
    (1, 16, 8, 32, 32, 3, 0, True, True),       # upcnv1b between two fp16-stored tensors, one tile
    (2, 21, 37, 32, 32, 3, 0, True, True),      # ragged
    (4, 128, 160, 32, 32, 3, 0, True, False),   # ring wraps
    (2, 16, 24, 64, 32, 4, 1, False, True),     # upcnv1 writes fp16
    (1, 19, 11, 64, 32, 4, 1, True, True),      # up-conv reading fp16 (two chunks per tile)
    (2, 32, 24, 128, 64, 4, 1, True, False),    # four chunks, wide variant
])
def test_fp16_stored_activations(B, H, W, Cin, Cout, k, mode, in_f16, out_f16):
    """RAW activations stored as fp16 between halo layers (lsi_b200_conv2d_halo_h): same result as the fp32-stored path up to
    the fp16 rounding of the stored values (2^-11 relative)."""
    out, ref, st = _run(B, H, W, Cin, Cout, k, mode, True, True, in_f16=in_f16, out_f16=out_f16)
    assert rel_err(out.cpu(), ref.cpu()) < (3e-3 if out_f16 else TOL)
    _check(ref, ref, None)
    mean = ref.mean(dim=(0, 1, 2)); var = ref.var(dim=(0, 1, 2), unbiased=False)
    assert rel_err(st[:, 0].cpu(), mean.cpu()) < TOL
    assert rel_err(st[:, 1].cpu(), torch.rsqrt(var + EPS).cpu()) < TOL


@pytest.mark.parametrize('B,H,W,Ca,Cb,Cout', [(1, 16, 8, 64, 32, 64), (2, 21, 37, 64, 32, 64), (2, 32, 24, 32, 32, 32), (1, 48, 40, 64, 64, 64)])
def test_two_source_concat_on_the_fly(B, H, W, Ca, Cb, Cout):
    """upcnv2b's input is tf.concat([upcnv2 output (batch norm pending), skip (normalised)]): the halo kernel reads both
    sources, normalising only the first; reference = fp32 kernel on the explicitly normalised, concatenated input."""
    from lsi import _b200
    from lsi.nnutils.nets import same_pad
    lib = _b200.lib()
    torch.manual_seed(Ca + Cb + H)
    dev = 'cuda'
    xa = (torch.randn(B, H, W, Ca, device=dev) * 1.7 + 0.3).half()
    xb = torch.relu(torch.randn(B, H, W, Cb, device=dev)).half()
    Cin = Ca + Cb
    w = torch.randn(3, 3, Cin, Cout, device=dev) / (9 * Cin) ** 0.5
    kw = dict(batch=B, h_in=H, w_in=W, c_in=Cin, h_out=H, w_out=W, c_out=Cout, kh=3, kw=3, stride=1, pad_top=same_pad(H, 3, 1)[0],
              pad_left=same_pad(W, 3, 1)[0], mode=0, w_tap_stride=Cin * Cout, w_ci_stride=Cout, w_co_stride=1, out_c_stride=Cout,
              epilogue=0, accumulate=0)
    xaf = xa.float()
    mean = xaf.mean(dim=(0, 1, 2)); var = xaf.var(dim=(0, 1, 2), unbiased=False)
    in_stats = torch.stack([mean, torch.rsqrt(var + EPS)], dim=1).contiguous()
    beta = torch.randn(Ca, device=dev) * 0.5
    xn = torch.cat([torch.relu((xaf - mean) * in_stats[:, 1] + beta), xb.float()], dim=3).contiguous()
    ref = torch.zeros(B, H, W, Cout, device=dev)
    _b200.call('lsi_b200_conv2d', _b200.ConvDesc(in_c_stride=Cin, **kw), _b200.ptr(xn), _b200.ptr(w), None, _b200.ptr(ref), _b200.stream())
    d = _b200.ConvDesc(in_c_stride=Ca, **kw)
    assert lib.lsi_b200_conv2d_halo_h_supported(d) == 1
    out = torch.full((B, H, W, Cout), 7.0, device=dev, dtype=torch.float16)
    st = torch.zeros(Cout, 2, device=dev)
    nws = lib.lsi_b200_conv2d_halo_workspace_bytes(d)
    wsb = torch.empty(nws, dtype=torch.uint8, device=dev)
    _b200.call('lsi_b200_conv2d_halo_h', d, _b200.ptr(xa), _b200.ptr(xb), Ca, Cb, 1, _b200.ptr(in_stats), _b200.ptr(beta), _b200.ptr(w),
               None, None, _b200.ptr(out), 1, _b200.ptr(st), EPS, _b200.ptr(wsb), nws, _b200.stream())
    torch.cuda.synchronize()
    assert rel_err(out.float().cpu(), ref.cpu()) < 3e-3
    mean_o = ref.mean(dim=(0, 1, 2)); var_o = ref.var(dim=(0, 1, 2), unbiased=False)
    assert rel_err(st[:, 0].cpu(), mean_o.cpu()) < TOL
    assert rel_err(st[:, 1].cpu(), torch.rsqrt(var_o + EPS).cpu()) < TOL


def test_fp16_stored_prediction_input():
    out, ref, _ = _run(2, 32, 48, 32, 4, 3, 0, True, False, epilogue=2, out_hw=(32, 40), in_f16=True)
    assert rel_err(out.cpu(), ref.cpu()) < TOL


def test_conv_bias_epilogue_64_channels():
    out, ref, _ = _run(2, 32, 24, 32, 64, 3, 0, True, False, epilogue=1)
    assert rel_err(out.cpu(), ref.cpu()) < TOL


def test_heads_match_generic_path():
    """The whole inference pipeline with the halo kernel on vs off (both TF32): same LDI within TF32 noise."""
    from lsi.nnutils import nets
    torch.manual_seed(0)
    store = nets.ParamStore()
    img = torch.rand(2, 128, 128, 3, device='cuda')
    outs = []
    for on in (False, True):
        nets.set_halo_mode(on)
        with torch.no_grad():
            _, fd, sk, _ = nets.encoder_decoder_unet(img, nl_diff_enc_dec=3, reuse=on, _store=store)
            tex, _, disp = nets.ldi_predictor(fd, n_layers=2, reuse=on, n_layerwise_steps=3, skip_feat=sk, _store=store)
        outs.append((tex.clone(), disp.clone()))
    nets.set_halo_mode(True)
    for a, b in zip(outs[0], outs[1]):
        d = (a - b).abs()
        assert float(d.max()) < 2e-2 and float(d.mean()) < 1e-3, (float(d.max()), float(d.mean()))


def test_f16_inference_mode_error_class():
    """nets.set_conv_mode('f16'): every activation between the stem and the prediction conv stored as fp16, kind::f16 MMAs.
    Measured against the fp32 CUDA-core mode, its error must be of the TF32 mode's class (one extra 2^-11 rounding per
    layer; the 2-sample batch norms of this fixture amplify either)."""
    from lsi.nnutils import nets
    torch.manual_seed(0)
    store = nets.ParamStore()
    img = torch.rand(2, 128, 128, 3, device='cuda')
    outs = {}
    try:
        for i, mode in enumerate(('fp32', 'tf32', 'f16')):
            nets.set_conv_mode(mode)
            with torch.no_grad():
                _, fd, sk, _ = nets.encoder_decoder_unet(img, nl_diff_enc_dec=3, reuse=i > 0, _store=store)
                tex, _, disp = nets.ldi_predictor(fd, n_layers=2, reuse=i > 0, n_layerwise_steps=3, skip_feat=sk, _store=store)
            assert tex.dtype == torch.float32 and disp.dtype == torch.float32
            outs[mode] = torch.cat([tex, disp], dim=-1).clone()
    finally:
        nets.set_conv_mode('tf32')
    e_tf32 = (outs['tf32'] - outs['fp32']).abs()
    e_f16 = (outs['f16'] - outs['fp32']).abs()
    print('tf32 vs fp32: max %.3e mean %.3e; f16 vs fp32: max %.3e mean %.3e' % (float(e_tf32.max()), float(e_tf32.mean()),
                                                                                float(e_f16.max()), float(e_f16.mean())))
    assert float(e_f16.mean()) < 2.5 * float(e_tf32.mean()) + 1e-4
    assert float(e_f16.max()) < 2.5 * float(e_tf32.max()) + 1e-3


def test_f16_mode_at_bench_resolution():
    """The bench pipeline's exact layer shapes (256x832 padded to 896, L=4; batch cut to 2): predict_ldi in the f16 mode --
    tensor-core stem, fp16 conv_tc, all-phase upcnv1, two-source upcnv2b, fp16 halo chain, fused crop -- against the TF32
    and fp32 modes, and the rendered view through forward_splat."""
    from lsi.geometry import ldi
    from lsi.nnutils import helpers, nets, train_utils
    torch.manual_seed(3)
    B, H, W, L = 2, 256, 832, 4
    img = torch.rand(B, H, W, 3, device='cuda')
    opts = train_utils.default_opts(dataset='kitti', n_layers=L, batch_size=B, img_height=H, img_width=W, zbuf_scale=50.0)
    store = nets.ParamStore(seed=0)
    k = torch.tensor([[721.54 * W / 1242.0, 0, 609.56 * W / 1242.0], [0, 721.54 * H / 375.0, 172.85 * H / 375.0], [0, 0, 1.0]],
                     device='cuda').expand(B, 3, 3).contiguous()
    rot = torch.eye(3, device='cuda').expand(B, 3, 3).contiguous()
    t = torch.tensor([[-0.5327], [0.0], [0.0]], device='cuda').expand(B, 3, 1).contiguous()
    pc = helpers.pixel_coords(B, H, W)
    outs = {}
    try:
        for i, mode in enumerate(('fp32', 'tf32', 'f16')):
            nets.set_conv_mode(mode)
            with torch.no_grad():
                pred = train_utils.predict_ldi(img, opts, store, reuse=i > 0)
                view, _ = ldi.forward_splat(tuple(pred), pc, k, k, rot, t, compose_layers=True, bg_layer_disp=1e-3,
                                            max_disp=0.4, zbuf_scale=50.0)
            assert pred[0].shape == (L, B, H, W, 3) and pred[2].shape == (L, B, H, W, 1)
            assert pred[0].dtype == torch.float32 and torch.isfinite(view).all()
            outs[mode] = (torch.cat([pred[0], pred[2] / 0.4], dim=-1).clone(), view.clone())
    finally:
        nets.set_conv_mode('tf32')
    # error class: measured against the exact fp32 mode, the f16 mode must stay within 2.5x of the TF32 mode's deviation
    # (batch-stat BN over 28 samples per channel at the bottleneck amplifies either)
    e = {m: ((outs[m][0] - outs['fp32'][0]).abs(), (outs[m][1] - outs['fp32'][1]).abs()) for m in ('tf32', 'f16')}
    print('LDI  mean|d| tf32 %.3e f16 %.3e; view mean|d| tf32 %.3e f16 %.3e' % (float(e['tf32'][0].mean()), float(e['f16'][0].mean()),
                                                                              float(e['tf32'][1].mean()), float(e['f16'][1].mean())))
    assert float(e['f16'][0].mean()) < 2.5 * float(e['tf32'][0].mean()) + 1e-4
    assert float(e['f16'][1].mean()) < 2.5 * float(e['tf32'][1].mean()) + 1e-4


def test_unsupported_shapes_are_refused():
    from lsi import _b200
    lib = _b200.lib()
    base = dict(batch=1, h_in=16, w_in=16, c_in=32, h_out=16, w_out=16, c_out=32, kh=3, kw=3, stride=1, pad_top=1, pad_left=1,
                mode=0, w_tap_stride=1024, w_ci_stride=32, w_co_stride=1, in_c_stride=32, out_c_stride=32, epilogue=0,
                accumulate=0)
    assert lib.lsi_b200_conv2d_halo_supported(_b200.ConvDesc(**base)) == 1
    for bad in (dict(c_in=3, in_c_stride=4), dict(c_in=256, in_c_stride=256), dict(stride=2, h_out=8, w_out=8),
                dict(c_out=128, out_c_stride=128), dict(kh=7, kw=7, pad_top=3, pad_left=3), dict(accumulate=1)):
        assert lib.lsi_b200_conv2d_halo_supported(_b200.ConvDesc(**dict(base, **bad))) == 0, bad
    d = _b200.ConvDesc(**dict(base, c_out=128, out_c_stride=128))
    x = torch.zeros(1, 16, 16, 32, device='cuda')
    with pytest.raises(RuntimeError):
        _b200.call('lsi_b200_conv2d_halo', d, _b200.ptr(x), None, None, _b200.ptr(x), None, None, _b200.ptr(x), None, EPS,
                   _b200.ptr(x), 4, _b200.stream())
