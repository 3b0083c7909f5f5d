"""ORACLE -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement (torch) of the reference CNN: lsi/nnutils/nets.py:244-348 (encoder_decoder_unet) and
:73-208 (decoder_simple, pixelwise_predictor, ldi_predictor) with the [TF1.4 slim] layer semantics spelled out:
NHWC, SAME padding (asymmetric at stride 2: pad_before = total // 2), conv -> batch-stat BN (beta only, eps 1e-3,
biased variance) -> ReLU, no conv bias under BN; prediction conv has bias + sigmoid and no BN; 4x4 stride-2
transposed conv == gradient of a SAME stride-2 conv (torch conv_transpose2d padding=1); weights
[kh,kw,cin,cout] (conv) / [kh,kw,cout,cin] (transposed conv); variables named as TF names them.

Parity status: the WIRING is pinned by fixtures produced by running the reference's own nets.py over the slim
stand-in (oracle/gen_golden.py -> tests/golden/nets_*.npz); the per-layer TF semantics above are restated, not pinned.
The FC stack on the bottleneck (nets.py:289-291) is built by the reference but never executed by either script (its
output is discarded, ldi_enc_dec.py:198) and is omitted here.
"""
import math
import zlib

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-3     # [TF1.4] slim.batch_norm default epsilon

# (name, kernel, stride, cout) of the encoder, nets.py:273-286
ENC = [('cnv1', 7, 2, 32), ('cnv1b', 7, 1, 32), ('cnv2', 5, 2, 64), ('cnv2b', 5, 1, 64), ('cnv3', 3, 2, 128),
       ('cnv3b', 3, 1, 128), ('cnv4', 3, 2, 256), ('cnv4b', 3, 1, 256), ('cnv5', 3, 2, 512), ('cnv5b', 3, 1, 512),
       ('cnv6', 3, 2, 512), ('cnv6b', 3, 1, 512), ('cnv7', 3, 2, 512), ('cnv7b', 3, 1, 512)]
# decoder level k: upcnv<k> (convT) -> concat skip -> icnv<k>; (k, cout, skip name), nets.py:296-345
DEC = [(7, 512, 'cnv6b'), (6, 512, 'cnv5b'), (5, 256, 'cnv4b'), (4, 128, 'cnv3b'), (3, 64, 'cnv2b'), (2, 32, 'cnv1b'),
       (1, 32, None)]
HEAD_FILTERS = [32, 64, 128, 256]     # nets.py:87


def param_shapes(n_layers, n_layerwise_steps=3, pred_masks=False, trunk_levels=None):
    """TF variable name -> shape for every variable the executed graph touches (trunk down to the level that feeds
    the heads, plus the L heads)."""
    shapes = {}
    cin = 3
    for name, k, _, cout in ENC:
        shapes['encoder_decoder_unet/%s/weights' % name] = [k, k, cin, cout]
        shapes['encoder_decoder_unet/%s/BatchNorm/beta' % name] = [cout]
        cin = cout
    enc_ch = {name: cout for name, _, _, cout in ENC}
    feat_c = 512
    n_dec = 7 - n_layerwise_steps if trunk_levels is None else trunk_levels
    for k, cout, skip in DEC[:n_dec]:
        shapes['encoder_decoder_unet/upcnv%d/weights' % k] = [4, 4, cout, feat_c]
        shapes['encoder_decoder_unet/upcnv%d/BatchNorm/beta' % k] = [cout]
        cat = cout + (enc_ch[skip] if skip else 0)
        shapes['encoder_decoder_unet/icnv%d/weights' % k] = [3, 3, cat, cout]
        shapes['encoder_decoder_unet/icnv%d/BatchNorm/beta' % k] = [cout]
        feat_c = cout
    nc = 4 + (1 if pred_masks else 0)
    skip_ch = [512, 512, 256, 128, 64, 32]        # skip_feat list order, nets.py:302-339
    for l in range(n_layers):
        base = 'ldi_tex_disp/pixelwise_pred/upsample_%d/' % l
        c = feat_c
        for step in range(n_layerwise_steps, 0, -1):
            cout = HEAD_FILTERS[step - 1]
            shapes[base + 'decoder/upcnv%d/weights' % step] = [4, 4, cout, c]
            shapes[base + 'decoder/upcnv%d/BatchNorm/beta' % step] = [cout]
            cat = cout + (skip_ch[len(skip_ch) - step + 1] if step > 1 else 0)      # skip_feat[-step+1], nets.py:108-109
            shapes[base + 'decoder/upcnv%db/weights' % step] = [3, 3, cat, cout]
            shapes[base + 'decoder/upcnv%db/BatchNorm/beta' % step] = [cout]
            c = cout
        shapes[base + 'pred_%d/weights' % l] = [3, 3, c, nc]
        shapes[base + 'pred_%d/biases' % l] = [nc]
    return shapes


def init_params(n_layers, seed=0, n_layerwise_steps=3, pred_masks=False, random_beta=False, dtype=torch.float32):
    """Xavier-uniform weights ([TF1.4] slim default), zero biases / betas (or small random ones for tests).  Each
    variable has its own RandomState keyed by (seed, name) so that any subset can be regenerated."""
    params = {}
    for name, shp in sorted(param_shapes(n_layers, n_layerwise_steps, pred_masks).items()):
        rs = np.random.RandomState((zlib.crc32(name.encode()) + seed) % (2 ** 31))
        if name.endswith('weights'):
            kh, kw, a, b = shp
            limit = math.sqrt(6.0 / (kh * kw * a + kh * kw * b))
            params[name] = torch.tensor(rs.uniform(-limit, limit, shp), dtype=dtype)
        elif random_beta:
            params[name] = torch.tensor(rs.uniform(-0.2, 0.2, shp), dtype=dtype)
        else:
            params[name] = torch.zeros(shp, dtype=dtype)
    return params


def same_pad(size, k, s):
    """[TF1.4] SAME: out = ceil(size/s); total = max((out-1)*s + k - size, 0); before = total // 2."""
    out = -(-size // s)
    total = max((out - 1) * s + k - size, 0)
    return total // 2, total - total // 2


def conv2d(x, w, stride):
    """NHWC conv with TF SAME padding; w [kh,kw,cin,cout]."""
    kh, kw = w.shape[0], w.shape[1]
    pt, pb = same_pad(x.shape[1], kh, stride)
    pl, pr = same_pad(x.shape[2], kw, stride)
    xn = F.pad(x.permute(0, 3, 1, 2), (pl, pr, pt, pb))
    return F.conv2d(xn, w.permute(3, 2, 0, 1), stride=stride).permute(0, 2, 3, 1)


def conv2d_transpose(x, w):
    """4x4 stride-2 SAME transposed conv; w [kh,kw,cout,cin]."""
    return F.conv_transpose2d(x.permute(0, 3, 1, 2), w.permute(3, 2, 0, 1), stride=2, padding=1).permute(0, 2, 3, 1)


def bn_relu(x, beta):
    """slim.batch_norm(is_training=True, center=True, scale=False, eps=1e-3) then ReLU (nets.py:263-272)."""
    dims = (0, 1, 2)
    mean = x.mean(dim=dims, keepdim=True)
    var = ((x - mean) ** 2).mean(dim=dims, keepdim=True)
    return torch.relu((x - mean) / torch.sqrt(var + BN_EPS) + beta)


def encoder_decoder_unet(params, inp_img, nl_diff_enc_dec=0):
    """nets.py:244-348.  Returns (feat_dec, skip_feat, end_points)."""
    P = lambda n: params['encoder_decoder_unet/' + n]
    ep = {}
    x = inp_img
    for name, _, stride, _ in ENC:
        x = bn_relu(conv2d(x, P(name + '/weights'), stride), P(name + '/BatchNorm/beta'))
        ep[name] = x
    skip_feat = [ep['cnv6b'], ep['cnv5b'], ep['cnv4b'], ep['cnv3b'], ep['cnv2b'], ep['cnv1b']]
    feats_dec = []
    feat = ep['cnv7b']
    for k, _, skip in DEC[:7 - nl_diff_enc_dec]:
        up = bn_relu(conv2d_transpose(feat, P('upcnv%d/weights' % k)), P('upcnv%d/BatchNorm/beta' % k))
        if skip is not None:
            if up.shape[1:3] != ep[skip].shape[1:3]:
                raise ValueError('U-Net needs H and W to be multiples of 128 (nets.py:298-300 concat %s vs %s)'
                                 % (tuple(up.shape), tuple(ep[skip].shape)))
            up = torch.cat([up, ep[skip]], dim=3)
        feat = bn_relu(conv2d(up, P('icnv%d/weights' % k), 1), P('icnv%d/BatchNorm/beta' % k))
        ep['icnv%d' % k] = feat
        feats_dec.append(feat)
    return feats_dec[-1], skip_feat, ep


def ldi_predictor(params, feat, n_layers, n_layerwise_steps, skip_feat, pred_masks=False):
    """nets.py:164-208 (+ pixelwise_predictor :117-161, decoder_simple :73-114).  Returns [tex, masks, disps]."""
    preds = []
    for l in range(n_layers):
        base = 'ldi_tex_disp/pixelwise_pred/upsample_%d/' % l
        f = feat
        for step in range(n_layerwise_steps, 0, -1):
            f = bn_relu(conv2d_transpose(f, params[base + 'decoder/upcnv%d/weights' % step]),
                        params[base + 'decoder/upcnv%d/BatchNorm/beta' % step])
            if step > 1 and skip_feat is not None:
                f = torch.cat([f, skip_feat[-step + 1]], dim=3)
            f = bn_relu(conv2d(f, params[base + 'decoder/upcnv%db/weights' % step], 1),
                        params[base + 'decoder/upcnv%db/BatchNorm/beta' % step])
        pred = torch.sigmoid(conv2d(f, params[base + 'pred_%d/weights' % l], 1) + params[base + 'pred_%d/biases' % l])
        preds.append(pred)
    preds = torch.stack(preds, dim=0)
    if pred_masks:
        tex, masks, disps = preds[..., 0:3], preds[..., 3:4], preds[..., 4:5]
        masks = torch.sigmoid(masks).clone()       # sigmoid applied twice, as the reference does (nets.py:143,202)
        masks[-1] = masks[-1] * 0 + 1
    else:
        tex, disps = preds[..., 0:3], preds[..., 3:4]
        masks = torch.ones_like(disps)
    return [tex, masks, disps]


def predict_ldi(params, img, n_layers, max_disp, n_layerwise_steps=3, pred_masks=False):
    """ldi_enc_dec.py:196-213: U-Net trunk -> heads -> disp *= max_disp."""
    feat_dec, skip_feat, _ = encoder_decoder_unet(params, img, nl_diff_enc_dec=n_layerwise_steps)
    tex, masks, disps = ldi_predictor(params, feat_dec, n_layers, n_layerwise_steps, skip_feat, pred_masks)
    return [tex, masks, disps * max_disp]


# ---------------------------------------------------------------------------------------------------------------------
# the non-U-Net variant (--use_unet=false): nets.py:29-70 (encoder_simple), :211-241 (encoder_decoder_simple)
# ---------------------------------------------------------------------------------------------------------------------
DEC_SIMPLE_FILTERS = [32, 64, 128, 256, 512, 512, 512, 512]      # nets.py:87-90 with nconv = 8


def param_shapes_simple(n_layers, img_hw, nz=1000, nupconv=8, nl_diff_enc_dec=3, n_layerwise_steps=3, pred_masks=False):
    """TF variable name -> shape for encoder_decoder_simple + the L heads (no skip connections)."""
    shapes = {}
    cin = 3
    for name, k, _, cout in ENC:
        shapes['encoder/%s/weights' % name] = [k, k, cin, cout]
        shapes['encoder/%s/BatchNorm/beta' % name] = [cout]
        cin = cout
    h7, w7 = img_hw[0] // 128, img_hw[1] // 128
    k_in = h7 * w7 * 512
    for i, n_out in enumerate([2 * nz, nz, nz]):                  # slim.stack scopes fc/fc_1..3
        shapes['encoder/fc/fc_%d/weights' % (i + 1)] = [k_in, n_out]
        shapes['encoder/fc/fc_%d/BatchNorm/beta' % (i + 1)] = [n_out]
        k_in = n_out
    c = nz
    for nc in range(nupconv - nl_diff_enc_dec, 0, -1):
        cout = DEC_SIMPLE_FILTERS[nc - 1]
        shapes['decoder/upcnv%d/weights' % nc] = [4, 4, cout, c]
        shapes['decoder/upcnv%d/BatchNorm/beta' % nc] = [cout]
        shapes['decoder/upcnv%db/weights' % nc] = [3, 3, cout, cout]
        shapes['decoder/upcnv%db/BatchNorm/beta' % nc] = [cout]
        c = cout
    nc_out = 4 + (1 if pred_masks else 0)
    for l in range(n_layers):
        base = 'ldi_tex_disp/pixelwise_pred/upsample_%d/' % l
        cc = c
        for step in range(n_layerwise_steps, 0, -1):
            cout = HEAD_FILTERS[step - 1]
            shapes[base + 'decoder/upcnv%d/weights' % step] = [4, 4, cout, cc]
            shapes[base + 'decoder/upcnv%d/BatchNorm/beta' % step] = [cout]
            shapes[base + 'decoder/upcnv%db/weights' % step] = [3, 3, cout, cout]
            shapes[base + 'decoder/upcnv%db/BatchNorm/beta' % step] = [cout]
            cc = cout
        shapes[base + 'pred_%d/weights' % l] = [3, 3, cc, nc_out]
        shapes[base + 'pred_%d/biases' % l] = [nc_out]
    return shapes


def init_params_simple(n_layers, img_hw, seed=0, random_beta=False, dtype=torch.float32, **kw):
    params = {}
    for name, shp in sorted(param_shapes_simple(n_layers, img_hw, **kw).items()):
        rs = np.random.RandomState((zlib.crc32(name.encode()) + seed) % (2 ** 31))
        if name.endswith('weights'):
            fan = (shp[0] + shp[1]) if len(shp) == 2 else (shp[0] * shp[1] * (shp[2] + shp[3]))
            limit = math.sqrt(6.0 / fan)
            params[name] = torch.tensor(rs.uniform(-limit, limit, shp), dtype=dtype)
        elif random_beta:
            params[name] = torch.tensor(rs.uniform(-0.2, 0.2, shp), dtype=dtype)
        else:
            params[name] = torch.zeros(shp, dtype=dtype)
    return params


def encoder_simple(params, inp_img):
    """nets.py:29-70: the 14 convolutions of the U-Net encoder under scope `encoder`, flatten (NHWC order), three fully
    connected layers (2nz, nz, nz) each with batch-stat BN (over the batch) + ReLU and no bias.  Returns (enc [B,nz], end_points)."""
    ep = {}
    x = inp_img
    for name, _, stride, _ in ENC:
        x = bn_relu(conv2d(x, params['encoder/%s/weights' % name], stride), params['encoder/%s/BatchNorm/beta' % name])
        ep[name] = x
    x = x.reshape(x.shape[0], -1)
    for i in (1, 2, 3):
        z = x @ params['encoder/fc/fc_%d/weights' % i]
        mean = z.mean(dim=0, keepdim=True)
        var = ((z - mean) ** 2).mean(dim=0, keepdim=True)
        x = torch.relu((z - mean) / torch.sqrt(var + BN_EPS) + params['encoder/fc/fc_%d/BatchNorm/beta' % i])
        ep['fc_%d' % i] = x
    return x, ep


def decoder_simple_plain(params, feat, nconv, scope='decoder/'):
    """nets.py:73-114 without skip features: nconv x [4x4 s2 up-conv -> 3x3 conv]; a [B,nz] input becomes [B,1,1,nz]."""
    if feat.dim() == 2:
        feat = feat[:, None, None, :]
    for nc in range(nconv, 0, -1):
        feat = bn_relu(conv2d_transpose(feat, params[scope + 'upcnv%d/weights' % nc]), params[scope + 'upcnv%d/BatchNorm/beta' % nc])
        feat = bn_relu(conv2d(feat, params[scope + 'upcnv%db/weights' % nc], 1), params[scope + 'upcnv%db/BatchNorm/beta' % nc])
    return feat


def encoder_decoder_simple(params, inp_img, nupconv=8, nl_diff_enc_dec=0):
    """nets.py:211-241.  Returns (feat, feat_dec, skip_feat=None, end_points)."""
    feat, ep = encoder_simple(params, inp_img)
    feat_dec = decoder_simple_plain(params, feat, nupconv - nl_diff_enc_dec)
    return feat, feat_dec, None, ep
