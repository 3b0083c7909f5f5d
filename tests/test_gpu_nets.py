"""GPU parity tests of the CNN slice (pytest -m gpu): layers and the whole U-Net + heads against the CPU oracle
(oracle/lsi_oracle_nets.py) and the fixture generated from the reference's own nets.py wiring."""
import numpy as np
import pytest
import torch

from _util import load_golden, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module', params=['fp32', 'tf32'])
def nets_mod(request):
    """Every test runs in both arithmetic modes: 'fp32' (CUDA-core kernels, 1e-4 class parity) and 'tf32' (tcgen05
    tensor-core kernels; TF32 has a 10-bit mantissa, so the bars are multiplied by TF32_SLACK)."""
    assert torch.cuda.is_available()
    from lsi.nnutils import nets
    nets.set_conv_mode(request.param)
    yield nets
    nets.set_conv_mode('tf32')


def _slack(nets_mod):
    return 1.0 if nets_mod.get_conv_mode() == 'fp32' else 100.0


def _check(nets_mod, got, ref, tol, what):
    """fp32 mode: max error within tol of the tensor scale.  tf32 mode: a TF32-sized perturbation flips ReLU masks
    (y near 0) and with them whole gradient contributions, so the max error of a gradient is O(1) at isolated elements;
    the check is on the bulk instead: 97 % of the elements within 100*tol, relative L2 error below 3e-2."""
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    if nets_mod.get_conv_mode() == 'fp32':
        assert rel_err(got, ref) < tol, what
        return
    scale = max(np.abs(ref).max(), 1e-30)
    frac = float((np.abs(got - ref) / scale > 100 * tol).mean())
    l2 = float(np.sqrt(((got - ref) ** 2).sum()) / max(np.sqrt((ref ** 2).sum()), 1e-30))
    assert frac <= 0.03 and l2 < 3e-2, '%s: %.3g of elements beyond %g, relative L2 error %.3g' % (what, frac, 100 * tol, l2)


@pytest.mark.parametrize('k,s,cin,cout,H,W,B', [
    (7, 2, 3, 32, 16, 24, 2),      # stem: 3 input channels, asymmetric SAME padding (2,3)
    (7, 1, 32, 32, 8, 12, 2),
    (5, 2, 32, 64, 12, 8, 1),      # SAME padding (1,2)
    (3, 2, 64, 128, 8, 8, 2),      # SAME padding (0,1)
    (3, 1, 192, 128, 6, 10, 2),    # head concat layer
    (3, 1, 40, 24, 5, 7, 3),       # ragged channel counts / sizes
])
def test_conv_bn_relu_layer(nets_mod, k, s, cin, cout, H, W, B):
    from oracle import lsi_oracle_nets as N
    torch.manual_seed(k * 100 + cin)
    x = torch.randn(B, H, W, cin)
    w = torch.randn(k, k, cin, cout) * (1.0 / (k * k * cin) ** 0.5)
    beta = torch.randn(cout) * 0.3
    g = torch.randn(B, -(-H // s), -(-W // s), cout)
    leaves = [t.clone().requires_grad_(True) for t in (x, w, beta)]
    ref = N.bn_relu(N.conv2d(leaves[0], leaves[1], s), leaves[2])
    ref_g = torch.autograd.grad((ref * g).sum(), leaves)
    store = nets_mod.ParamStore()
    store.load_state_dict({'t/weights': w, 't/BatchNorm/beta': beta})
    xg = x.cuda().requires_grad_(True)
    out = nets_mod._conv_layer(store, 't', xg, cout, k, s, reuse=True)
    got_g = torch.autograd.grad((out * g.cuda()).sum(), [xg, store.vars['t/weights'], store.vars['t/BatchNorm/beta']])
    _check(nets_mod, out.detach().cpu(), ref.detach(), 1e-4, 'y')
    for a, b, nme in zip(got_g, ref_g, ('dx', 'dw', 'dbeta')):
        _check(nets_mod, a.cpu(), b, 2e-4, nme)


@pytest.mark.parametrize('cin,cout,H,W,B', [(128, 64, 4, 6, 2), (512, 512, 1, 2, 2), (32, 32, 8, 8, 1), (24, 40, 3, 5, 2)])
def test_upconv_bn_relu_layer(nets_mod, cin, cout, H, W, B):
    from oracle import lsi_oracle_nets as N
    torch.manual_seed(cin + cout)
    x = torch.randn(B, H, W, cin)
    w = torch.randn(4, 4, cout, cin) * (1.0 / (4 * cin) ** 0.5)
    beta = torch.randn(cout) * 0.3
    g = torch.randn(B, 2 * H, 2 * W, cout)
    leaves = [t.clone().requires_grad_(True) for t in (x, w, beta)]
    ref = N.bn_relu(N.conv2d_transpose(leaves[0], leaves[1]), leaves[2])
    ref_g = torch.autograd.grad((ref * g).sum(), leaves)
    store = nets_mod.ParamStore()
    store.load_state_dict({'t/weights': w, 't/BatchNorm/beta': beta})
    xg = x.cuda().requires_grad_(True)
    out = nets_mod._conv_layer(store, 't', xg, cout, 4, 2, reuse=True, transposed=True)
    got_g = torch.autograd.grad((out * g.cuda()).sum(), [xg, store.vars['t/weights'], store.vars['t/BatchNorm/beta']])
    _check(nets_mod, out.detach().cpu(), ref.detach(), 1e-4, 'y')
    for a, b, nme in zip(got_g, ref_g, ('dx', 'dw', 'dbeta')):
        _check(nets_mod, a.cpu(), b, 2e-4, nme)


def test_unet_and_heads_golden(nets_mod):
    """Whole network at 128x128, L=2, B=2 against the fixture from the reference's wiring (fp64 evaluation as truth;
    the reference's own fp32 evaluation is allowed the same distance).  In tf32 mode this is an integration check only:
    20 TF32 layers with batch-stat BN over as few as 2 samples (the 1x1 bottleneck at this size) drift by ~2e-2 at the
    sigmoid outputs, so the bars are 100x wider there."""
    from oracle import lsi_oracle_nets as N
    g = load_golden('nets_unet_l2')
    L, B, H, W, steps = (int(v) for v in g['meta'])
    params = N.init_params(L, seed=int(g['param_seed']), n_layerwise_steps=steps, random_beta=True)
    store = nets_mod.ParamStore()
    store.load_state_dict(params)
    img = torch.tensor(g['in_img'], device='cuda').requires_grad_(True)
    _, feat_dec, skip_feat, _ = nets_mod.encoder_decoder_unet(img, nl_diff_enc_dec=steps, reuse=True, _store=store)
    tex, masks, disps = nets_mod.ldi_predictor(feat_dec, n_layers=L, reuse=True, n_layerwise_steps=steps, skip_feat=skip_feat,
                                               _store=store)
    assert getattr(masks, '_lsi_all_ones', False) and float(masks.min()) == 1.0
    pred = torch.cat([tex, disps * float(g['max_disp'])], dim=-1)
    ref_noise = rel_err(g['pred_f32'], g['pred_f64'])
    k = _slack(nets_mod)
    _check(nets_mod, feat_dec.detach().cpu()[:, ::2, ::2, ::4], g['feat_dec_f64'], max(1e-3, 2 * rel_err(g['feat_dec_f32'], g['feat_dec_f64'])), 'feat_dec')
    _check(nets_mod, pred.detach().cpu()[:, :, ::4, ::4, :], g['pred_f64'], max(1e-4, 2 * ref_noise), 'pred')
    g_pred = torch.tensor(np.random.RandomState(int(g['g_seed'])).normal(0, 1, (L, B, H, W, 4)).astype(np.float32), device='cuda')
    names = sorted(store.vars)
    assert names == [str(n) for n in g['grad_names']]
    grads = torch.autograd.grad((pred * g_pred).sum(), [store.vars[n] for n in names] + [img])
    if nets_mod.get_conv_mode() == 'fp32':
        assert rel_err(grads[-1].cpu()[:, ::4, ::4], g['d_img_f64']) < max(1e-3, 2 * rel_err(g['d_img_f32'], g['d_img_f64']))
    for key in g:
        if key.startswith('grad:') and key.endswith('_f64'):
            nme = key[5:-4]
            bar = max(1e-3, 2 * rel_err(g['grad:' + nme + '_f32'], g[key]))
            if nets_mod.get_conv_mode() == 'fp32':
                assert rel_err(grads[names.index(nme)].cpu(), g[key]) < bar, nme
    l2 = np.array([float(x.double().pow(2).sum().sqrt()) for x in grads[:-1]])
    # tf32 mode: the gradient norms of all 72 variables stay within 30 % (ReLU-mask flips through ~20 layers); fp32: 0.5 %
    assert np.all(np.abs(l2 - g['grad_l2_f64']) <= (5e-3 if k == 1.0 else 0.5) * np.maximum(g['grad_l2_f64'], 1e-6))


def test_prediction_conv_with_fused_crop(nets_mod):
    """The padded-input policy crops inside the prediction conv: evaluating the top-left window only must equal
    conv-then-crop, forward and backward (sigmoid head, bias)."""
    from oracle import lsi_oracle_nets as N
    torch.manual_seed(3)
    B, H, W, cin, nc, h, w = 2, 16, 32, 32, 4, 16, 23
    x, wt, b = torch.randn(B, H, W, cin), torch.randn(3, 3, cin, nc) * 0.1, torch.randn(nc) * 0.1
    g = torch.randn(B, h, w, nc)
    leaves = [t.clone().requires_grad_(True) for t in (x, wt, b)]
    ref = torch.sigmoid(N.conv2d(leaves[0], leaves[1], 1) + leaves[2])[:, :h, :w]
    ref_g = torch.autograd.grad((ref * g).sum(), leaves)
    geo = nets_mod._Geometry(False, B, H, W, cin, nc, 3, 1, out_hw=(h, w))
    gl = [t.cuda().requires_grad_(True) for t in (x, wt, b)]
    out = nets_mod._ConvBiasSigmoid.apply(gl[0], gl[1], gl[2], geo)
    assert out.shape == (B, h, w, nc)
    got_g = torch.autograd.grad((out * g.cuda()).sum(), gl)
    _check(nets_mod, out.detach().cpu(), ref.detach(), 1e-4, 'y')
    for a, bb, nme in zip(got_g, ref_g, ('dx', 'dw', 'db')):
        _check(nets_mod, a.cpu(), bb, 2e-4, nme)


def test_unet_rejects_illegal_sizes_and_pads(nets_mod):
    store = nets_mod.ParamStore()
    with pytest.raises(ValueError, match='multiples of 128'):
        nets_mod.encoder_decoder_unet(torch.rand(1, 64, 64, 3, device='cuda'), _store=store)
    with pytest.raises(NotImplementedError):
        nets_mod.encoder_decoder_unet(torch.rand(1, 128, 128, 3, device='cuda'), is_training=False, _store=store)
    padded, (h, w) = nets_mod.pad_to_legal(torch.rand(1, 64, 200, 3, device='cuda'))
    assert padded.shape == (1, 128, 256, 3) and (h, w) == (64, 200)


def test_param_store_flatten_and_adam(nets_mod):
    """Flat parameter/gradient buffers and the fused Adam step against torch.optim.Adam semantics restated for TF
    (epsilon outside the square root, lr_t = lr*sqrt(1-b2^t)/(1-b1^t))."""
    from lsi import _b200
    store = nets_mod.ParamStore()
    a = store.get('a/weights', [3, 3, 4, 8], False, 'weights')
    b = store.get('a/BatchNorm/beta', [8], False, 'beta')
    flat, grad = store.flatten()
    assert flat.numel() == 3 * 3 * 4 * 8 + 8
    (store.vars['a/weights'].sum() * 2 + (store.vars['a/BatchNorm/beta'] * 3).sum()).backward()
    assert torch.all(grad[:8] == 3) and torch.all(grad[8:] == 2)        # sorted names: beta first
    p0 = flat.clone()
    m, v = torch.zeros_like(flat), torch.zeros_like(flat)
    lr, b1, b2, eps = 1e-4, 0.9, 0.999, 1e-8
    for step in (1, 2, 3):
        _b200.call('lsi_b200_adam_step', _b200.ptr(flat), _b200.ptr(grad), _b200.ptr(m), _b200.ptr(v), flat.numel(), lr, b1, b2,
                   eps, step, 1.0, _b200.stream())
    gm, gv, ref = torch.zeros_like(p0), torch.zeros_like(p0), p0.clone()
    for step in (1, 2, 3):
        gm = b1 * gm + (1 - b1) * grad
        gv = b2 * gv + (1 - b2) * grad * grad
        ref -= lr * (1 - b2 ** step) ** 0.5 / (1 - b1 ** step) * gm / (gv.sqrt() + eps)
    assert rel_err(flat.cpu(), ref.cpu()) < 1e-6


@pytest.mark.parametrize('mode', ['fp32', 'split', 'tf32'])
def test_simple_encoder_decoder_golden(mode):
    """--use_unet=false (nets.py:29-70, 211-241): encoder_simple + fully connected stack + decoder_simple + heads against the fixture
    produced by the reference's own wiring.  fp32 / split: parity bars; tf32: integration check."""
    from lsi.nnutils import nets
    from oracle import lsi_oracle_nets as N
    g = load_golden('nets_simple_l1')
    L, B, H, W, steps, nz = (int(v) for v in g['meta'])
    params = N.init_params_simple(L, (H, W), seed=int(g['param_seed']), random_beta=True, nz=nz)
    nets.set_conv_mode(mode)
    try:
        store = nets.ParamStore()
        store.load_state_dict(params)
        with torch.no_grad():
            feat, feat_dec, skip, ep = nets.encoder_decoder_simple(torch.tensor(g['in_img'], device='cuda'), nz=nz, nl_diff_enc_dec=steps,
                                                                  reuse=True, _store=store)
            tex, masks, disps = nets.ldi_predictor(feat_dec, n_layers=L, reuse=True, n_layerwise_steps=steps, skip_feat=skip, _store=store)
        assert skip is None and tuple(feat.shape) == (B, nz) and sorted(store.vars) == [str(n) for n in g['var_names']]
        pred = torch.cat([tex, disps * float(g['max_disp'])], dim=-1)
        e_feat = rel_err(nets.to_float(feat).cpu(), g['feat_f64'])
        e_pred = rel_err(pred.cpu()[:, :, ::8, ::8, :], g['pred_f64'])
        n_feat, n_pred = rel_err(g['feat_f32'], g['feat_f64']), rel_err(g['pred_f32'], g['pred_f64'])
        print('simple enc-dec, mode %s: feat %.3g (oracle fp32 noise %.3g), pred %.3g (noise %.3g)' % (mode, e_feat, n_feat, e_pred, n_pred))
        if mode == 'tf32':        # integration check only: batch norm over 4 samples after the fully connected layers amplifies TF32 rounding
            assert np.isfinite(pred.cpu().numpy()).all() and e_pred < 1.0
        else:
            assert e_feat < max(1e-3, 3 * n_feat) and e_pred < max(1e-4, 3 * n_pred)
    finally:
        nets.set_conv_mode('tf32')


@pytest.mark.parametrize('mode', ['fwd', 'dgrad'])
@pytest.mark.parametrize('H,W', [(9, 13), (16, 64)])
def test_thin_conv_quad_kernel(mode, H, W):
    """lsi_b200_conv2d_thin on the shape the training step sends it (data gradient of the 3x3 prediction conv, nets.py:139-155:
    4 -> 32 channels; four pixels x eight channels per thread) and the same shape as a forward gather, against the generic fp32
    kernel lsi_b200_conv2d on the same descriptor; widths that are not multiples of four, image borders, accumulate."""
    from lsi import _b200
    from lsi.nnutils import nets
    torch.manual_seed(H + W)
    B = 2
    if mode == 'fwd':
        geo = nets._Geometry(False, B, H, W, 4, 32, 3, 1)
        desc, w = geo.fwd, torch.randn(3, 3, 4, 32, device='cuda')
    else:
        geo = nets._Geometry(False, B, H, W, 32, 4, 3, 1)
        desc, w = geo.dgrad, torch.randn(3, 3, 32, 4, device='cuda')
    x = torch.randn(B, H, W, 4, device='cuda')
    for acc in (0, 1):
        d = _b200.ConvDesc(**dict(desc, accumulate=acc))
        assert _b200.lib().lsi_b200_conv2d_thin_supported(d) == 1
        base = torch.randn(B, H, W, 32, device='cuda')
        got, ref = base.clone(), base.clone()
        _b200.call('lsi_b200_conv2d_thin', d, _b200.ptr(x), _b200.ptr(w), _b200.ptr(got), _b200.stream())
        _b200.call('lsi_b200_conv2d', d, _b200.ptr(x), _b200.ptr(w), None, _b200.ptr(ref), _b200.stream())
        assert rel_err(got.cpu(), ref.cpu()) < 1e-6
