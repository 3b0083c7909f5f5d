"""CPU: the oracle CNN (oracle/lsi_oracle_nets.py) against the fixture produced by the reference's own nets.py wiring
(oracle/gen_golden.py gen_nets), plus TF-SAME padding and U-Net size-legality unit tests (KAT 9)."""
import numpy as np
import pytest
import torch

from oracle import lsi_oracle_nets as N
from _util import load_golden, rel_err


def _run(g, dtype=torch.float32):
    L, B, H, W, steps = (int(v) for v in g['meta'])
    params = N.init_params(L, seed=int(g['param_seed']), n_layerwise_steps=steps, random_beta=True, dtype=dtype)
    leaves = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    x = torch.tensor(g['in_img'], dtype=dtype, requires_grad=True)
    tex, masks, disps = N.predict_ldi(leaves, x, L, float(g['max_disp']), n_layerwise_steps=steps)
    pred = torch.cat([tex, disps], dim=-1)
    g_pred = torch.tensor(np.random.RandomState(int(g['g_seed'])).normal(0, 1, (L, B, H, W, 4)).astype(np.float32), dtype=dtype)
    names = sorted(leaves)
    grads = torch.autograd.grad((pred * g_pred).sum(), [leaves[n] for n in names] + [x])
    return pred.detach(), dict(zip(names, grads[:-1])), grads[-1]


def test_unet_matches_reference_wiring():
    g = load_golden('nets_unet_l2')
    pred, grads, d_img = _run(g)
    assert rel_err(pred[:, :, ::4, ::4, :], g['pred_f32']) < 1e-5
    assert abs(float(pred.double().sum()) - g['pred_stats_f32'][0]) < 1e-4 * abs(g['pred_stats_f32'][0])
    assert rel_err(d_img[:, ::4, ::4], g['d_img_f32']) < 1e-3
    for key in g:
        if key.startswith('grad:') and key.endswith('_f32'):
            assert rel_err(grads[key[5:-4]], g[key]) < 1e-3, key
    names = [str(n) for n in g['grad_names']]
    assert names == sorted(grads)
    l2 = np.array([float(grads[n].double().pow(2).sum().sqrt()) for n in names])
    assert np.all(np.abs(l2 - g['grad_l2_f32']) <= 2e-3 * np.maximum(g['grad_l2_f32'], 1e-6))


def test_fp32_fixture_vs_fp64_fixture():
    """How far the reference's own fp32 CNN is from its fp64 evaluation: the bar the CUDA path is held to."""
    g = load_golden('nets_unet_l2')
    assert rel_err(g['pred_f32'], g['pred_f64']) < 1e-4
    assert rel_err(g['feat_dec_f32'], g['feat_dec_f64']) < 1e-3


def test_tf_same_padding_is_asymmetric_at_stride_2():
    assert N.same_pad(256, 7, 2) == (2, 3)
    assert N.same_pad(256, 5, 2) == (1, 2)
    assert N.same_pad(256, 3, 2) == (0, 1)
    assert N.same_pad(256, 7, 1) == (3, 3) and N.same_pad(256, 3, 1) == (1, 1)
    x = torch.arange(16, dtype=torch.float32).reshape(1, 4, 4, 1)
    w = torch.ones(3, 3, 1, 1)
    y = N.conv2d(x, w, 2)                      # 3x3 s2 on a 4x4 ramp: windows start at rows/cols 0 and 2, pad on the far side
    assert y.shape == (1, 2, 2, 1)
    assert y[0, 0, 0, 0].item() == float(x[0, 0:3, 0:3, 0].sum())
    assert y[0, 1, 1, 0].item() == float(x[0, 2:4, 2:4, 0].sum())


def test_conv_transpose_is_adjoint_of_same_stride2_conv():
    torch.manual_seed(0)
    x = torch.randn(1, 4, 6, 3, dtype=torch.float64)
    w = torch.randn(4, 4, 5, 3, dtype=torch.float64)          # [kh,kw,cout,cin] of the transposed conv
    y = N.conv2d_transpose(x, w)                              # [1,8,12,5]
    z = torch.randn_like(y)
    # <convT(x), z> == <x, conv_s2(z)> with the forward-conv weight layout [kh,kw,cin=5,cout=3]
    lhs = (y * z).sum()
    rhs = (x * N.conv2d(z, w, 2)).sum()
    assert abs(lhs.item() - rhs.item()) < 1e-9 * abs(lhs.item())


def test_unet_rejects_sizes_that_are_not_multiples_of_128():
    params = N.init_params(1, seed=0)
    with pytest.raises(ValueError, match='multiples of 128'):
        N.encoder_decoder_unet(params, torch.rand(1, 64, 64, 3), nl_diff_enc_dec=3)


def test_parameter_inventory():
    shapes = N.param_shapes(4)
    n = sum(int(np.prod(s)) for s in shapes.values())
    assert abs(n - 39.07e6) < 0.3e6            # 14.29 M enc + 21.92 M dec + 0.713 M x L
    assert shapes['ldi_tex_disp/pixelwise_pred/upsample_0/decoder/upcnv3b/weights'] == [3, 3, 192, 128]
    assert shapes['ldi_tex_disp/pixelwise_pred/upsample_3/pred_3/weights'] == [3, 3, 32, 4]
    assert shapes['encoder_decoder_unet/icnv7/weights'] == [3, 3, 1024, 512]
