"""GPU parity tests of the 'split' conv mode (pytest -m gpu): the tcgen05 tensor-core inference path whose activations and
weights are fp16 (hi, lo) pairs -- 22 mantissa bits, csrc/split.cuh -- held to the SAME bars as the fp32 CUDA-core mode
(1e-4 of the tensor scale per layer; whole network within max(1e-4, 2 x the reference arithmetic's own fp32-vs-fp64 noise)),
always against the CPU oracle (oracle/lsi_oracle_nets.py) or the fixture generated from the reference's own nets.py."""
import numpy as np
import pytest
import torch

from _util import load_golden, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def nets_split():
    assert torch.cuda.is_available()
    from lsi.nnutils import nets
    nets.set_conv_mode('split')
    yield nets
    nets.set_conv_mode('tf32')


def test_split_pack_roundtrip(nets_split):
    """v -> (hi, lo) -> v: error <= 2^-21 |v| + 2^-36, exact zeros, saturation at the fp16 range."""
    torch.manual_seed(0)
    x = torch.randn(2, 5, 7, 64, device='cuda') * torch.logspace(-6, 3, 64, device='cuda')
    x[0, 0, 0, :8] = 0.0
    s = nets_split._SplitAct.pack(x)
    assert s.t.dtype == torch.float16 and s.t.numel() == 2 * x.numel() and tuple(s.shape) == (2, 5, 7, 64)
    y = s.float()
    excess = ((y - x).abs() - (2.0 ** -21 * x.abs() + 2.0 ** -36)).max().item()
    assert excess <= 0, excess
    assert torch.all(y[0, 0, 0, :8] == 0)
    # the layout: per pixel and 32-channel chunk, 32 hi values then 32 lo values
    hi = s.t[..., 0, :].reshape(2, 5, 7, 64).float()
    assert torch.equal(hi, x.half().float())
    big = torch.full((1, 1, 1, 32), 1e6, device='cuda')
    assert torch.isfinite(nets_split._SplitAct.pack(big).float()).all()


LAYERS = [
    (7, 2, 3, 32, 16, 24, 2),      # stem (fp32 CUDA-core conv, packed afterwards)
    (7, 1, 32, 32, 8, 12, 2),
    (7, 1, 32, 32, 32, 40, 1),     # x-merged halo tiles (>= 16 rows)
    (5, 2, 32, 64, 12, 8, 1),      # stride 2 through the tensor map's element strides, SAME padding (1,2)
    (5, 1, 64, 64, 32, 16, 1),
    (3, 2, 64, 128, 8, 8, 2),
    (3, 1, 128, 128, 16, 24, 2),   # N = 128: 256-column MMAs, the whole TMEM
    (3, 1, 64, 256, 16, 16, 1),    # two Cout tiles
    (3, 1, 512, 512, 2, 7, 4),     # K = 4608, several images per tile
    (3, 1, 96, 64, 20, 36, 1),
]


@pytest.mark.parametrize('k,s,cin,cout,H,W,B', LAYERS)
def test_conv_bn_relu_layer_split(nets_split, k, s, cin, cout, H, W, B):
    from oracle import lsi_oracle_nets as N
    torch.manual_seed(k * 100 + cin + H)
    x = torch.randn(B, H, W, cin)
    w = torch.randn(k, k, cin, cout) * (1.0 / (k * k * cin) ** 0.5)
    beta = torch.randn(cout) * 0.3
    ref = N.bn_relu(N.conv2d(x.double(), w.double(), s), beta.double())
    store = nets_split.ParamStore()
    store.load_state_dict({'t/weights': w, 't/BatchNorm/beta': beta})
    with torch.no_grad():
        out = nets_split._conv_layer(store, 't', x.cuda(), cout, k, s, reuse=True)
    assert isinstance(out, nets_split._SplitAct)
    e = rel_err(out.float().cpu(), ref)
    print('split conv k%d s%d %d->%d: rel err %.3g' % (k, s, cin, cout, e))
    assert e < 1e-4, e


def test_two_source_conv_split(nets_split):
    """tf.concat([a, b], axis=3) -> 3x3 conv (nets.py:108-109) read from two split tensors."""
    from oracle import lsi_oracle_nets as N
    torch.manual_seed(5)
    a, b = torch.randn(2, 16, 24, 128), torch.randn(2, 16, 24, 64)
    w = torch.randn(3, 3, 192, 128) / (9 * 192) ** 0.5
    beta = torch.randn(128) * 0.3
    ref = N.bn_relu(N.conv2d(torch.cat([a, b], 3).double(), w.double(), 1), beta.double())
    store = nets_split.ParamStore()
    store.load_state_dict({'t/weights': w, 't/BatchNorm/beta': beta})
    with torch.no_grad():
        out = nets_split._conv_layer(store, 't', (a.cuda(), b.cuda()), 128, 3, 1, reuse=True)
    assert rel_err(out.float().cpu(), ref) < 1e-4


@pytest.mark.parametrize('cin,cout,H,W,B', [(128, 64, 4, 6, 2), (512, 512, 1, 2, 2), (32, 32, 8, 8, 1), (64, 32, 32, 16, 1),
                                            (128, 128, 16, 20, 2)])
def test_upconv_bn_relu_layer_split(nets_split, cin, cout, H, W, B):
    from oracle import lsi_oracle_nets as N
    torch.manual_seed(cin + cout)
    x = torch.randn(B, H, W, cin)
    w = torch.randn(4, 4, cout, cin) * (1.0 / (4 * cin) ** 0.5)
    beta = torch.randn(cout) * 0.3
    ref = N.bn_relu(N.conv2d_transpose(x.double(), w.double()), beta.double())
    store = nets_split.ParamStore()
    store.load_state_dict({'t/weights': w, 't/BatchNorm/beta': beta})
    with torch.no_grad():
        out = nets_split._conv_layer(store, 't', x.cuda(), cout, 4, 2, reuse=True, transposed=True)
    e = rel_err(out.float().cpu(), ref)
    print('split upconv %d->%d: rel err %.3g' % (cin, cout, e))
    assert e < 1e-4, e


@pytest.mark.parametrize('halo', [True, False])
def test_head_chain_split(nets_split, halo):
    """Tail of one LDI head (nets.py:87-114, 139-155): 4x4/2 up-conv 64->32 -> 3x3 32->32 -> 3x3 32->4 + bias + sigmoid, every batch
    norm + ReLU left pending and applied on load by the halo-tile kernel (halo=True) or in place by the generic path."""
    from oracle import lsi_oracle_nets as N
    torch.manual_seed(9)
    B, H, W = 2, 24, 20
    x = torch.relu(torch.randn(B, H, W, 64))
    w1, b1 = torch.randn(4, 4, 32, 64) / 16.0, torch.randn(32) * 0.3
    w2, b2 = torch.randn(3, 3, 32, 32) / 17.0, torch.randn(32) * 0.3
    w3, b3 = torch.randn(3, 3, 32, 4) / 17.0, torch.randn(4) * 0.3
    d = lambda t: t.double()
    f = N.bn_relu(N.conv2d_transpose(d(x), d(w1)), d(b1))
    f = N.bn_relu(N.conv2d(f, d(w2), 1), d(b2))
    ref_feat = f
    ref = torch.sigmoid(N.conv2d(f, d(w3), 1) + d(b3))[:, :2 * H - 3, :2 * W - 5] * torch.tensor([1.0, 1.0, 1.0, 0.4], dtype=torch.float64)
    store = nets_split.ParamStore()
    store.load_state_dict({'h/upcnv1/weights': w1, 'h/upcnv1/BatchNorm/beta': b1, 'h/upcnv1b/weights': w2, 'h/upcnv1b/BatchNorm/beta': b2})
    old = nets_split._HALO
    nets_split.set_halo_mode(halo)
    try:
        with torch.no_grad():
            xs = nets_split._SplitAct.pack(x.cuda())
            f1 = nets_split._conv_layer(store, 'h/upcnv1', xs, 32, 4, 2, reuse=True, transposed=True, defer=True)
            assert isinstance(f1, nets_split._Pending)
            f2 = nets_split._conv_layer(store, 'h/upcnv1b', f1, 32, 3, 1, reuse=True, defer=True)
            assert (f1._done is None) == halo           # halo path: upcnv1's normalised output never existed in HBM
            geo = nets_split._Geometry(False, B, 2 * H, 2 * W, 32, 4, 3, 1, out_hw=(2 * H - 3, 2 * W - 5))
            dp = nets_split._b200.ConvDesc(**dict(geo.fwd, epilogue=2))
            y = torch.empty(B, geo.Ho, geo.Wo, 4, device='cuda')
            scale = torch.tensor([1.0, 1.0, 1.0, 0.4], device='cuda')
            lib = nets_split._b200.lib()
            w3g, b3g = w3.cuda(), b3.cuda()
            if halo:
                assert lib.lsi_b200_conv2d_halo_s_supported(dp) == 1
                ws = nets_split._tc_workspace(y.device, int(lib.lsi_b200_conv2d_halo_workspace_bytes(dp)))
                nets_split._b200.call('lsi_b200_conv2d_halo_s', dp, nets_split._b200.ptr(f2.z.t), nets_split._b200.ptr(f2.stats),
                                      nets_split._b200.ptr(f2.beta), nets_split._b200.ptr(w3g), nets_split._b200.ptr(b3g),
                                      nets_split._b200.ptr(scale), nets_split._b200.ptr(y), 0, None, 1e-3, nets_split._b200.ptr(ws),
                                      ws.numel(), nets_split._b200.stream())
            else:
                fm = f2.materialize()
                ws = nets_split._tc_workspace(y.device, int(lib.lsi_b200_conv2d_tc_workspace_bytes(dp)))
                nets_split._b200.call('lsi_b200_conv2d_tc_s', dp, nets_split._b200.ptr(fm.t), 32, None, 0, nets_split._b200.ptr(w3g),
                                      nets_split._b200.ptr(b3g), nets_split._b200.ptr(scale), nets_split._b200.ptr(y), 0, None, 1e-3,
                                      nets_split._b200.ptr(ws), ws.numel(), nets_split._b200.stream())
            feat = nets_split.to_float(f2)
    finally:
        nets_split.set_halo_mode(old)
    e_f, e_y = rel_err(feat.cpu(), ref_feat), rel_err(y.cpu(), ref)
    print('split head chain (halo=%s): features %.3g, prediction %.3g' % (halo, e_f, e_y))
    assert e_f < 1e-4 and e_y < 1e-4


def test_halo_split_plain_input_and_wide_output(nets_split):
    """Halo-tile kernel on an already normalised split input (no transform warps involved) and with 64 output channels."""
    from oracle import lsi_oracle_nets as N
    torch.manual_seed(4)
    x = torch.relu(torch.randn(1, 40, 24, 32))
    w, beta = torch.randn(3, 3, 32, 64) / 17.0, torch.randn(64) * 0.3
    ref = N.bn_relu(N.conv2d(x.double(), w.double(), 1), beta.double())
    store = nets_split.ParamStore()
    store.load_state_dict({'t/weights': w, 't/BatchNorm/beta': beta})
    with torch.no_grad():
        out = nets_split._conv_layer(store, 't', nets_split._SplitAct.pack(x.cuda()), 64, 3, 1, reuse=True)
    assert rel_err(out.float().cpu(), ref) < 1e-4


def _predict(nets, params, img, L, steps, max_disp, out_hw=None):
    store = nets.ParamStore()
    store.load_state_dict(params)
    with torch.no_grad():
        _, feat_dec, skip_feat, _ = nets.encoder_decoder_unet(img, nl_diff_enc_dec=steps, reuse=True, _store=store)
        tex, masks, disps = nets.ldi_predictor(feat_dec, n_layers=L, reuse=True, n_layerwise_steps=steps, skip_feat=skip_feat,
                                               _store=store, _out_hw=out_hw, _disp_scale=max_disp)
    return nets.to_float(feat_dec), torch.cat([tex, disps], dim=-1)


def test_unet_and_heads_golden_split(nets_split):
    """Whole network at 128x128, L=2, B=2 against the fixture from the reference's own wiring (fp64 evaluation as truth),
    with the bars of the fp32 mode (tests/test_gpu_nets.py::test_unet_and_heads_golden)."""
    from oracle import lsi_oracle_nets as N
    g = load_golden('nets_unet_l2')
    L, B, H, W, steps = (int(v) for v in g['meta'])
    params = N.init_params(L, seed=int(g['param_seed']), n_layerwise_steps=steps, random_beta=True)
    feat_dec, pred = _predict(nets_split, params, torch.tensor(g['in_img'], device='cuda'), L, steps, float(g['max_disp']))
    e_feat, e_pred = rel_err(feat_dec.cpu()[:, ::2, ::2, ::4], g['feat_dec_f64']), rel_err(pred.cpu()[:, :, ::4, ::4, :], g['pred_f64'])
    n_feat, n_pred = rel_err(g['feat_dec_f32'], g['feat_dec_f64']), rel_err(g['pred_f32'], g['pred_f64'])
    print('split whole net 128x128: feat_dec %.3g (fp32 oracle noise %.3g), pred %.3g (noise %.3g)' % (e_feat, n_feat, e_pred, n_pred))
    assert e_feat < max(1e-3, 2 * n_feat)
    assert e_pred < max(1e-4, 2 * n_pred)


@pytest.mark.parametrize('H,W,B,L,crop_w', [(256, 256, 4, 2, None), (256, 896, 2, 4, 832)])
def test_whole_net_vs_oracle_split(nets_split, H, W, B, L, crop_w):
    """Well-conditioned legal sizes (256x256 B=4; the bench's 256x832 padded to 896, B=2, L=4, crop fused into the prediction
    conv): the split tensor-core path against the oracle's fp64 evaluation; bar = max(1e-4, 2 x |oracle fp32 - oracle fp64|)."""
    from oracle import lsi_oracle_nets as N
    from oracle import gen_inputs
    rs = np.random.RandomState(11)
    img = np.stack([gen_inputs.band_limited(rs, (H, W), 3) for _ in range(B)]).astype(np.float32)
    if crop_w is not None:
        img[:, :, crop_w:] = 0.0                      # nets.pad_to_legal: zero padding to the next multiple of 128
    params = N.init_params(L, seed=3, random_beta=True)
    max_disp = 0.4
    refs = {}
    for dt in (torch.float32, torch.float64):
        p = {k: v.to(dt) for k, v in params.items()}
        with torch.no_grad():
            tex, _, disps = N.predict_ldi(p, torch.tensor(img, dtype=dt), L, max_disp)
        refs[dt] = torch.cat([tex, disps], dim=-1)[:, :, :, :(crop_w or W)].numpy()
    out_hw = None if crop_w is None else (H, crop_w)
    _, pred = _predict(nets_split, params, torch.tensor(img, device='cuda'), L, 3, max_disp, out_hw=out_hw)
    assert tuple(pred.shape) == refs[torch.float64].shape
    noise = rel_err(refs[torch.float32], refs[torch.float64])
    e = rel_err(pred.cpu(), refs[torch.float64])
    mean_abs = float(np.abs(pred.cpu().numpy().astype(np.float64) - refs[torch.float64]).mean())
    print('split whole net %dx%d B=%d L=%d: max rel err %.3g, mean |d| %.3g (fp32 oracle noise %.3g)' % (H, W, B, L, e, mean_abs, noise))
    assert e < max(1e-4, 2 * noise), (e, noise)


def test_split_mode_with_autograd_runs_exact_kernels(nets_split):
    """With autograd enabled the 'split' mode runs the fp32 CUDA-core kernels (the mode the gradient parity tests use)."""
    torch.manual_seed(1)
    x = torch.randn(1, 8, 8, 32, device='cuda', requires_grad=True)
    store = nets_split.ParamStore()
    out = nets_split._conv_layer(store, 't', x, 32, 3, 1, reuse=False)
    assert isinstance(out, torch.Tensor) and out.dtype == torch.float32 and out.requires_grad


@pytest.mark.parametrize('cin,cout,H,W', [(64, 64, 16, 16), (32, 32, 16, 24)])      # generic kernel / halo-tile kernel
def test_filter_memo_follows_the_weights(nets_split, cin, cout, H, W):
    """The inference path memoises the re-laid-out (split) filter bank per (weight tensor, store epoch, tensor version)
    (lsi_b200_set_weight_version).  It must never serve a stale filter: in-place updates (version counter), a new store whose weight
    tensor lands on the same address, and raw-pointer updates followed by ParamStore.touch() (what Trainer.train_step does after the
    fused Adam kernel) all have to show up in the next forward pass -- bit for bit what a fresh store gives."""
    nets = nets_split
    torch.manual_seed(5)
    x = torch.randn(2, H, W, cin).cuda()
    ws = [torch.randn(3, 3, cin, cout) / (9 * cin) ** 0.5 for _ in range(4)]
    beta = torch.randn(cout) * 0.3

    def fresh(w):
        st = nets.ParamStore()
        st.load_state_dict({'t/weights': w, 't/BatchNorm/beta': beta})
        with torch.no_grad():
            return nets._conv_layer(st, 't', x, cout, 3, 1, reuse=True).float()

    want = [fresh(w) for w in ws]
    assert not torch.equal(want[0], want[1])
    store = nets.ParamStore()
    store.load_state_dict({'t/weights': ws[0], 't/BatchNorm/beta': beta})
    run = lambda: nets._conv_layer(store, 't', x, cout, 3, 1, reuse=True).float()
    with torch.no_grad():
        assert torch.equal(run(), want[0])
        assert torch.equal(run(), want[0])                       # memo hit
        store.vars['t/weights'].copy_(ws[1].cuda())              # in-place update: version counter
        assert torch.equal(run(), want[1])
        store.load_state_dict({'t/weights': ws[2]})              # epoch
        assert torch.equal(run(), want[2])
        store.vars['t/weights'].data.copy_(ws[3].cuda())         # a write the version counter does not see ...
        store.touch()                                            # ... announced the way Trainer.train_step announces the Adam kernel
        assert torch.equal(run(), want[3])
