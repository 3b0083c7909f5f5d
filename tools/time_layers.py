"""Per-launch timing of the inference CNN (U-Net trunk + L heads) on one GPU: every C-ABI call is bracketed by CUDA events
on the launching stream and aggregated by (entry point, layer shape).  Prints ms per step, effective TFLOP/s and the
algorithmic HBM GB/s (input read once + output written once) per layer, i.e. which roofline each layer sits under."""
import argparse, collections, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'layered-scene-inference_b200')); sys.path.insert(0, ROOT)
import torch
from lsi import _b200
from lsi.nnutils import nets

ap = argparse.ArgumentParser()
ap.add_argument('--batch', type=int, default=64); ap.add_argument('--h', type=int, default=256); ap.add_argument('--w', type=int, default=896)
ap.add_argument('--layers', type=int, default=4); ap.add_argument('--iters', type=int, default=3)
ap.add_argument('--out_w', type=int, default=832); ap.add_argument('--mode', default=None)
a = ap.parse_args()
if a.mode:
    nets.set_conv_mode(a.mode)

records = []
orig_call = _b200.call


def timed_call(name, *args):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    orig_call(name, *args)
    e1.record()
    key, fl, by = name, 0.0, 0.0
    if args and isinstance(args[0], _b200.ConvDesc):
        d = args[0]
        key = '%s %dx%d %d->%d k%d s%d m%d @%dx%d' % (name.replace('lsi_b200_', ''), d.h_in, d.w_in, d.c_in, d.c_out, d.kh, d.stride, d.mode, d.h_out, d.w_out)
        taps = d.kh * d.kw / (d.stride * d.stride if d.mode == 1 else 1)
        fl = 2.0 * d.batch * d.h_out * d.w_out * d.c_out * d.c_in * taps
        by = 4.0 * d.batch * (d.h_in * d.w_in * d.c_in + d.h_out * d.w_out * d.c_out)
    elif name == 'lsi_b200_bn_relu_apply_h':
        P, C = args[5], args[6]
        key = 'bn_relu_apply_h P=%d C=%d' % (P, C)
        by = (6.0 if not args[1] else 4.0) * P * C
    elif name == 'lsi_b200_split_convert':
        P, C = args[6], args[7]
        key = 'split_convert P=%d C=%d in_split=%d bn=%d' % (P, C, args[1], int(args[2] is not None))
        by = 8.0 * P * C
    elif name == 'lsi_b200_bn_relu_forward':
        P, C = args[4], args[5]
        key = 'bn_relu_forward P=%d C=%d' % (P, C)
        by = 8.0 * P * C
    records.append((key, e0, e1, fl, by))


_b200.call = timed_call
store = nets.ParamStore()
img = torch.rand(a.batch, a.h, a.w, 3, device='cuda')


def fwd(reuse):
    _, fd, sk, _ = nets.encoder_decoder_unet(img, nl_diff_enc_dec=3, reuse=reuse, _store=store)
    return nets.ldi_predictor(fd, n_layers=a.layers, reuse=reuse, n_layerwise_steps=3, skip_feat=sk, _store=store,
                              _out_hw=(a.h, a.out_w))


with torch.no_grad():
    fwd(False); fwd(True)
    torch.cuda.synchronize()
    records.clear()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(a.iters):
        fwd(True)
    t1.record()
torch.cuda.synchronize()
agg = collections.OrderedDict()
for key, e0, e1, fl, by in records:
    v = agg.setdefault(key, [0.0, 0, 0.0, 0.0])
    v[0] += e0.elapsed_time(e1); v[1] += 1; v[2] += fl; v[3] += by
tot = sum(v[0] for v in agg.values()) / a.iters
print('mode=%s halo=%s  B=%d %dx%d L=%d: %.2f ms/step wall (events), %.2f ms summed over launches' % (nets.get_conv_mode(), nets._HALO, a.batch, a.h, a.w, a.layers, t0.elapsed_time(t1) / a.iters, tot))
print('%-72s %5s %9s %8s %9s' % ('call', 'n', 'ms/step', 'TFLOP/s', 'alg GB/s'))
for key, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    ms = v[0] / a.iters
    print('%-72s %5d %9.3f %8.1f %9.0f' % (key[:72], v[1] // a.iters, ms, v[2] / a.iters / ms / 1e9 if ms > 0 else 0, v[3] / a.iters / ms / 1e6 if ms > 0 else 0))
