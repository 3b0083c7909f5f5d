"""Mirror of lsi/nnutils/nets.py (reference tree): the encoder-decoder U-Net and the per-layer LDI prediction heads.

Same function names, arguments and return values as the reference; the layers run as CUDA kernels behind the C ABI
(lsi_b200_conv2d / _conv2d_wgrad / _bn_relu_forward / _bn_relu_backward / ... in include/lsi_b200.h).  Semantics
follow TF-1.4 slim as the reference uses it: NHWC, SAME padding (asymmetric at stride 2), conv -> batch-stat BN (beta
only, eps 1e-3, biased variance) -> ReLU with no conv bias; the prediction conv has bias + sigmoid and no BN; 4x4
stride-2 transposed conv.  Variables are named exactly as TF names them (`encoder_decoder_unet/cnv1/weights`,
`.../BatchNorm/beta`, `ldi_tex_disp/pixelwise_pred/upsample_<l>/decoder/upcnv3/weights`, `.../pred_<l>/biases`), kept in
a `ParamStore` that plays the role of the TF variable scope (`reuse=True` looks variables up instead of creating them).

Not provided (never executed by either reference script): the FC stack on the bottleneck (nets.py:289-291, its output
`feat` is discarded at ldi_enc_dec.py:198 -- `None` is returned in its place), the decoder levels below the one that
feeds the heads, and inference-mode batch norm (the reference never updates the moving averages, train_utils.py:107-117, and evaluates with
batch statistics, ldi_pred_eval.py:45).
"""
import math
import os
import zlib

import numpy as np
import torch

from lsi import _b200
from lsi.nnutils import helpers as nn_helpers

BN_EPS = 1e-3     # [TF1.4] slim.batch_norm default epsilon

ENC = [('cnv1', 7, 2, 32), ('cnv1b', 7, 1, 32), ('cnv2', 5, 2, 64), ('cnv2b', 5, 1, 64), ('cnv3', 3, 2, 128),
       ('cnv3b', 3, 1, 128), ('cnv4', 3, 2, 256), ('cnv4b', 3, 1, 256), ('cnv5', 3, 2, 512), ('cnv5b', 3, 1, 512),
       ('cnv6', 3, 2, 512), ('cnv6b', 3, 1, 512), ('cnv7', 3, 2, 512), ('cnv7b', 3, 1, 512)]          # nets.py:273-286
DEC = [(7, 512, 'cnv6b'), (6, 512, 'cnv5b'), (5, 256, 'cnv4b'), (4, 128, 'cnv3b'), (3, 64, 'cnv2b'), (2, 32, 'cnv1b'),
       (1, 32, None)]                                                                                   # nets.py:296-345
HEAD_FILTERS = [32, 64, 128, 256]                                                                       # nets.py:87


# ---------------------------------------------------------------------------------------------------------------------
# variables
# ---------------------------------------------------------------------------------------------------------------------
class ParamStore(object):
    """TF-variable-scope stand-in: name -> leaf tensor.  `flatten()` re-packs every variable (and its gradient) into
    one contiguous buffer each, which is what the NCCL all-reduce and the fused Adam kernel operate on."""

    _epochs = [0]        # process-wide: no two stores, and no two parameter states of one store, ever share an epoch

    def __init__(self, device='cuda', seed=0):
        self.device = torch.device(device)
        self.seed = seed
        self.vars = {}
        self.flat = None
        self.flat_grad = None
        self.touch()

    def touch(self):
        """Call after changing parameter VALUES by anything that does not move the tensors' version counters (a kernel writing through
        raw pointers: the fused Adam step).  The inference path memoises re-laid-out filters per (weight tensor, epoch, version)."""
        ParamStore._epochs[0] += 1
        self.epoch = ParamStore._epochs[0]

    def weight_version(self, w):
        """Non-zero 64-bit tag that changes whenever `w` may hold different values: this store's epoch and the tensor's version."""
        return ((self.epoch & 0xffffffff) << 32) | ((w._version + 1) & 0xffffffff)

    def get(self, name, shape, reuse, kind):
        if name in self.vars:
            v = self.vars[name]
            if list(v.shape) != list(shape):
                raise RuntimeError('variable %s has shape %s, requested %s' % (name, list(v.shape), list(shape)))
            return v
        if reuse:
            raise RuntimeError('variable %s does not exist (reuse=True)' % name)
        if self.flat is not None:
            raise RuntimeError('ParamStore is already flattened; create every variable first')
        rs = np.random.RandomState((zlib.crc32(name.encode()) + self.seed) % (2 ** 31))
        if kind == 'weights':          # [TF1.4] slim default: Xavier-uniform
            fan = (shape[0] + shape[1]) if len(shape) == 2 else shape[0] * shape[1] * (shape[2] + shape[3])   # fully connected: [in, out]
            limit = math.sqrt(6.0 / fan)
            v = torch.tensor(rs.uniform(-limit, limit, shape), dtype=torch.float32)
        else:                          # biases / BatchNorm beta: zeros
            v = torch.zeros(shape, dtype=torch.float32)
        v = v.to(self.device).requires_grad_(True)
        self.vars[name] = v
        return v

    def load_state_dict(self, state):
        """Set variables from {TF name: tensor}; creates missing ones."""
        for k, t in state.items():
            t = t.detach().to(self.device, torch.float32)
            if k in self.vars:
                with torch.no_grad():
                    self.vars[k].copy_(t)
            else:
                if self.flat is not None:
                    raise RuntimeError('ParamStore is already flattened')
                self.vars[k] = t.clone().requires_grad_(True)
        self.touch()

    def state_dict(self):
        return {k: v.detach().clone() for k, v in self.vars.items()}

    def flatten(self):
        """One flat parameter buffer + one flat gradient buffer; variables become views (TF name order)."""
        if self.flat is not None:
            return self.flat, self.flat_grad
        names = sorted(self.vars)
        n = sum(self.vars[k].numel() for k in names)
        flat = torch.empty(n, dtype=torch.float32, device=self.device)
        grad = torch.zeros(n, dtype=torch.float32, device=self.device)
        off = 0
        for k in names:
            v = self.vars[k]
            m = v.numel()
            flat[off:off + m].copy_(v.detach().reshape(-1))
            nv = flat[off:off + m].view(v.shape).detach().requires_grad_(True)
            nv.grad = grad[off:off + m].view(v.shape)
            self.vars[k] = nv
            off += m
        self.flat, self.flat_grad = flat, grad
        self.touch()
        return flat, grad

    def zero_grad(self):
        if self.flat_grad is not None:
            self.flat_grad.zero_()
        else:
            for v in self.vars.values():
                v.grad = None


_DEFAULT_STORE = None


def get_default_store():
    global _DEFAULT_STORE
    if _DEFAULT_STORE is None:
        _DEFAULT_STORE = ParamStore()
    return _DEFAULT_STORE


def set_default_store(store):
    global _DEFAULT_STORE
    _DEFAULT_STORE = store


# ---------------------------------------------------------------------------------------------------------------------
# layer kernels
# ---------------------------------------------------------------------------------------------------------------------
def same_pad(size, k, s):
    """[TF1.4] SAME: out = ceil(size/s); total = max((out-1)*s + k - size, 0); before = total // 2."""
    out = -(-size // s)
    total = max((out - 1) * s + k - size, 0)
    return total // 2, total - total // 2


_bn_ws = {}


def _bn_workspace(device, channels):
    key = (device.index, torch.cuda.current_stream().cuda_stream)
    need = int(_b200.lib().lsi_b200_bn_workspace_bytes(max(channels, 1024)))
    if key not in _bn_ws or _bn_ws[key].numel() < need:
        _bn_ws[key] = torch.empty(need, dtype=torch.uint8, device=device)
    return _bn_ws[key]


_CONV_MODE = os.environ.get('LSI_B200_CONV_MODE', 'tf32')   # 'tf32' | 'fp32' | 'f16' | 'split' (see set_conv_mode)


def set_conv_mode(mode):
    """'tf32' (default): tcgen05 tensor-core kernels (TF32 inputs, fp32 accumulation) wherever the layer shape allows,
    fp32 CUDA-core kernels elsewhere (the 3-channel stem, weight gradients).  'fp32': CUDA-core kernels everywhere --
    the bit-for-bit-reproducible-arithmetic mode the 1e-4 parity tests of the CNN run in.  'split': the tensor-core mode
    that meets the fp32 parity bars (inference path: every activation and weight is a pair of fp16 numbers carrying 22
    mantissa bits, three exact fp16 MMAs per fp32 product, fp32 accumulation -- csrc/split.cuh; with autograd enabled
    the layers run the 'fp32' kernels).  'f16': inference-only speed mode with plain fp16 activations."""
    global _CONV_MODE
    if mode not in ('tf32', 'fp32', 'f16', 'split'):
        raise ValueError(mode)
    _CONV_MODE = mode


def _tc_mode():
    return _CONV_MODE in ('tf32', 'f16')


def _f16_infer():
    """'f16': inference-only (no_grad) variant of 'tf32' in which every activation between the stem and the prediction conv
    lives in HBM as fp16 (same 10-bit mantissa as a TF32 operand), the tensor-core kernels run kind::f16 MMAs with fp32
    accumulation and batch statistics come from the fp32 accumulators.  With autograd enabled it behaves like 'tf32'."""
    return _CONV_MODE == 'f16' and not torch.is_grad_enabled()


def _split_infer():
    """'split': no-grad inference path on split fp16-pair activations (see set_conv_mode)."""
    return _CONV_MODE == 'split' and not torch.is_grad_enabled()


class _SplitAct(object):
    """An activation [B,H,W,C] (C % 32 == 0) in the split layout of csrc/split.cuh: per pixel and 32-channel chunk 32 fp16
    `hi` values then 32 fp16 `lo` values (v = hi + lo / 2048) -- the byte size and chunk addresses of the fp32 tensor.
    Exists only on the no-grad inference path of the 'split' conv mode; `.float()` converts back."""

    def __init__(self, t, shape):
        self.t, self.shape, self.device = t, torch.Size(shape), t.device

    @staticmethod
    def empty(B, H, W, C, device):
        if C % 32:
            raise RuntimeError('lsi_b200: split activations need a multiple of 32 channels, got %d' % C)
        return _SplitAct(torch.empty(B, H, W, C // 32, 2, 32, dtype=torch.float16, device=device), (B, H, W, C))

    @staticmethod
    def pack(x):
        x = _b200.dev_f32(x, 'activation')
        B, H, W, C = x.shape
        out = _SplitAct.empty(B, H, W, C, x.device)
        _b200.call('lsi_b200_split_convert', _b200.ptr(x), 0, None, None, _b200.ptr(out.t), 1, B * H * W, C, _b200.stream())
        return out

    def float(self):
        B, H, W, C = self.shape
        out = torch.empty(B, H, W, C, dtype=torch.float32, device=self.device)
        _b200.call('lsi_b200_split_convert', _b200.ptr(self.t), 1, None, None, _b200.ptr(out), 0, B * H * W, C, _b200.stream())
        return out

    def dim(self):
        return 4


def to_float(x):
    """fp32 tensor of an activation returned by the inference paths ('split' / 'f16' modes hand their internal formats on)."""
    x = _materialize(x)
    return x.float() if isinstance(x, _SplitAct) or x.dtype != torch.float32 else x


def _dev_act(t, name):
    """A hot-path activation: CUDA, contiguous, fp32 (or fp16 on the 'f16' inference path)."""
    if isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float16 and _f16_infer():
        return t if t.is_contiguous() else t.contiguous()
    return _b200.dev_f32(t, name)


def get_conv_mode():
    return _CONV_MODE


_tc_ws = {}


def _tc_workspace(device, nbytes):
    key = (device.index, torch.cuda.current_stream().cuda_stream)
    if key not in _tc_ws or _tc_ws[key].numel() < nbytes:
        _tc_ws[key] = torch.empty(max(int(nbytes), 1 << 20), dtype=torch.uint8, device=device)
    return _tc_ws[key]


_WEIGHT_MEMO = os.environ.get('LSI_B200_WEIGHT_MEMO', '1') != '0'


def _vouch(store, w):
    """Inference path (no_grad): tell the library that `w` is unchanged since the last call that passed the same tag, so that the
    K-major / fp16 / split re-layout of the filter (a launch in front of every tensor-core convolution) is reused instead of redone
    (lsi_b200_set_weight_version, include/lsi_b200.h).  Must be followed immediately by the conv call, which consumes the tag."""
    if _WEIGHT_MEMO and store is not None and not torch.is_grad_enabled():
        _b200.lib().lsi_b200_set_weight_version(store.weight_version(w))


def _tc_ok(d, c_in_a, *tensors):
    return (_tc_mode() and _b200.lib().lsi_b200_conv2d_tc_supported(d, c_in_a) == 1
            and all(t is None or t.data_ptr() % 16 == 0 for t in tensors))


def _conv(desc_kw, inp, w, out, bias=None, bn_stats=None, inp_b=None, c_in_a=None, store=None):
    """One convolution launch.  inp_b: optional second source (channels [c_in_a, c_in)), i.e. tf.concat on the fly
    (tensor-core path only).  bn_stats: [C,2] tensor to receive (mean, rsqrt(var+eps)) of the output, reduced in the conv
    epilogue (tensor-core path only).  Returns True when bn_stats was filled."""
    d = _b200.ConvDesc(**desc_kw)
    lib = _b200.lib()
    ca = d.c_in if c_in_a is None else c_in_a
    if (_HALO and _HALO_GENERIC and _tc_mode() and inp_b is None and inp.dtype == torch.float32 and out.dtype == torch.float32
            and inp.data_ptr() % 16 == 0 and out.data_ptr() % 16 == 0 and lib.lsi_b200_conv2d_halo_supported(d) == 1):
        # full-resolution 32/64-channel head layers, also inside the training step (forward convs, the data gradient of
        # upcnv1b): resident filter bank + one halo box per tile instead of one box per tap
        _conv_halo(d, inp, w, out, bias=bias, out_stats=bn_stats, store=store)
        return bn_stats is not None
    if _tc_ok(d, ca, inp, inp_b):
        nws = int(lib.lsi_b200_conv2d_tc_workspace_bytes(d))
        ws = _tc_workspace(inp.device, nws)
        cb_stride = 0 if inp_b is None else inp_b.shape[-1]
        _vouch(store, w)      # (the two towers of a training step share their filters: the second one reuses the re-laid-out banks)
        if bn_stats is not None and d.epilogue == 0 and d.accumulate == 0:
            _b200.call('lsi_b200_conv2d_tc_bnstats', d, _b200.ptr(inp), ca, _b200.ptr(inp_b), cb_stride, _b200.ptr(w),
                       _b200.ptr(out), _b200.ptr(bn_stats), BN_EPS, _b200.ptr(ws), ws.numel(), _b200.stream())
            return True
        _b200.call('lsi_b200_conv2d_tc', d, _b200.ptr(inp), ca, _b200.ptr(inp_b), cb_stride, _b200.ptr(w), _b200.ptr(bias),
                   _b200.ptr(out), _b200.ptr(ws), ws.numel(), _b200.stream())
        return False
    if inp_b is not None:
        raise RuntimeError('lsi_b200: two-source convolution needs the tensor-core path')
    if bias is None and lib.lsi_b200_conv2d_thin_supported(d) == 1:      # e.g. the data gradient of the 32 -> 4 prediction conv
        _b200.call('lsi_b200_conv2d_thin', d, _b200.ptr(inp), _b200.ptr(w), _b200.ptr(out), _b200.stream())
        return False
    _b200.call('lsi_b200_conv2d', d, _b200.ptr(inp), _b200.ptr(w), _b200.ptr(bias), _b200.ptr(out), _b200.stream())
    return False


def _wgrad(desc_kw, big, small, dw):
    d = _b200.ConvDesc(**desc_kw)
    if (_tc_mode() and _b200.lib().lsi_b200_conv2d_wgrad_tc_supported(d) == 1 and big.data_ptr() % 16 == 0
            and small.data_ptr() % 16 == 0):
        _b200.call('lsi_b200_conv2d_wgrad_tc', d, _b200.ptr(big), _b200.ptr(small), _b200.ptr(dw), _b200.stream())
        return
    if _b200.lib().lsi_b200_conv2d_stem_wgrad_supported(d) == 1 and dw.is_contiguous():      # the 3-channel stem
        _b200.call('lsi_b200_conv2d_stem_wgrad', d, _b200.ptr(big), _b200.ptr(small), _b200.ptr(dw), _b200.stream())
        return
    _b200.call('lsi_b200_conv2d_wgrad', d, _b200.ptr(big), _b200.ptr(small), _b200.ptr(dw), _b200.stream())


_HALO = os.environ.get('LSI_B200_CONV_HALO', '1') != '0'
_HALO_F16_STORE = os.environ.get('LSI_B200_HALO_F16_STORE', '1') != '0'
_STEM_TC = os.environ.get('LSI_B200_STEM_TC', '1') != '0'
_HALO_CONCAT = os.environ.get('LSI_B200_HALO_CONCAT', '1') != '0'
_HALO_GENERIC = os.environ.get('LSI_B200_HALO_GENERIC', '1') != '0'      # halo kernel from _conv (training step)
_OUT_SCALE_CACHE = {}
_BN_BWD_FAST = os.environ.get('LSI_B200_BN_BWD_FAST', '1') != '0'
_SPLIT_UPCONV_MATERIALIZE = os.environ.get('LSI_B200_SPLIT_UPCONV_MATERIALIZE', '1') != '0'


def set_halo_mode(on):
    """Inference path only: route the full-resolution few-channel head layers through the halo-tile kernel
    (lsi_b200_conv2d_halo: resident weights, one TMA halo box per tile, the producer's batch norm + ReLU applied on
    load) instead of the generic tensor-core kernel + a separate normalise pass.  Default on (LSI_B200_CONV_HALO=0
    disables)."""
    global _HALO
    _HALO = bool(on)


class _Pending(object):
    """RAW output of a slim.conv2d whose batch_norm + ReLU (nets.py:263-272) has not been applied yet: z [B,H,W,C],
    stats [C,2] = (mean, rstd), beta [C].  Exists only on the no-grad inference path, between a producing conv and a
    consumer that normalises on load; `materialize()` applies it in place for every other consumer."""

    def __init__(self, z, stats, beta):
        self.z, self.stats, self.beta = z, stats, beta
        self.shape = z.shape
        self.device = z.device
        self._done = None

    def materialize(self):
        if self._done is None:
            self._done = self._apply()
        return self._done

    def _apply(self):
        z = self.z
        B, H, W, C = z.shape
        if isinstance(z, _SplitAct):     # normalise + ReLU in place on the split pairs
            _b200.call('lsi_b200_split_convert', _b200.ptr(z.t), 1, _b200.ptr(self.beta), _b200.ptr(self.stats), _b200.ptr(z.t), 1,
                       B * H * W, C, _b200.stream())
            return z
        if z.dtype == torch.float16 and _f16_infer():
            _b200.call('lsi_b200_bn_relu_apply_h', _b200.ptr(z), 1, _b200.ptr(self.beta), _b200.ptr(self.stats), _b200.ptr(z),
                       B * H * W, C, _b200.stream())
            return z
        if z.dtype != torch.float32:     # fp16-stored raw output met a consumer that cannot normalise on load
            z = z.float()
        _b200.call('lsi_b200_bn_relu_forward', _b200.ptr(z), _b200.ptr(self.beta), _b200.ptr(z), _b200.ptr(self.stats),
                   B * H * W, C, C, C, BN_EPS, 1, 1, _b200.ptr(_bn_workspace(z.device, C)), _b200.stream())
        return z


def _materialize(x):
    return x.materialize() if isinstance(x, _Pending) else x


def _halo_ok(d, *tensors):
    return (_HALO and _tc_mode() and not torch.is_grad_enabled()
            and _b200.lib().lsi_b200_conv2d_halo_supported(d) == 1
            and all(t is None or t.data_ptr() % 16 == 0 for t in tensors))


def _conv_halo(d, x, w, out, bias=None, out_stats=None, out_scale=None, x_b=None, store=None):
    """One halo-tile launch; x: tensor or _Pending (normalised on load); x_b: optional second source (already normalised),
    i.e. tf.concat([x, x_b], axis=3) on the fly."""
    lib = _b200.lib()
    pend = isinstance(x, _Pending)
    xin = x.z if pend else x
    nws = int(lib.lsi_b200_conv2d_halo_workspace_bytes(d))
    ws = _tc_workspace(xin.device, nws)
    _vouch(store, w)          # (inference call sites pass their store; the training path does not)
    _b200.call('lsi_b200_conv2d_halo_h', d, _b200.ptr(xin), _b200.ptr(x_b), xin.shape[3], 0 if x_b is None else x_b.shape[3],
               int(xin.dtype == torch.float16),
               _b200.ptr(x.stats) if pend else None, _b200.ptr(x.beta) if pend else None, _b200.ptr(w), _b200.ptr(bias),
               _b200.ptr(out_scale), _b200.ptr(out), int(out.dtype == torch.float16), _b200.ptr(out_stats), BN_EPS,
               _b200.ptr(ws), ws.numel(), _b200.stream())


class _Geometry(object):
    """Descriptors of one layer: forward, data gradient, weight gradient."""

    def __init__(self, transposed, B, Hi, Wi, Cin, Cout, k, s, out_hw=None):
        self.transposed, self.B, self.Hi, self.Wi, self.Cin, self.Cout, self.k, self.s = transposed, B, Hi, Wi, Cin, Cout, k, s
        if not transposed:
            self.Ho, self.Wo = -(-Hi // s), -(-Wi // s)
            if out_hw is not None:      # stride-1 conv evaluated on the top-left out_hw window only (crop fused into the conv)
                assert s == 1 and out_hw[0] <= self.Ho and out_hw[1] <= self.Wo
                self.Ho, self.Wo = out_hw
            pt, pl = same_pad(Hi, k, s)[0], same_pad(Wi, k, s)[0]
            if s > 1 and (Hi % s or Wi % s):
                raise ValueError('stride-%d conv needs even input sizes, got %dx%d' % (s, Hi, Wi))
            com = dict(batch=B, kh=k, kw=k, stride=s, pad_top=pt, pad_left=pl, epilogue=0, accumulate=0)
            # weights [kh,kw,cin,cout]
            self.fwd = dict(com, h_in=Hi, w_in=Wi, c_in=Cin, h_out=self.Ho, w_out=self.Wo, c_out=Cout, mode=0,
                            w_tap_stride=Cin * Cout, w_ci_stride=Cout, w_co_stride=1, in_c_stride=Cin, out_c_stride=Cout)
            self.dgrad = dict(com, h_in=self.Ho, w_in=self.Wo, c_in=Cout, h_out=Hi, w_out=Wi, c_out=Cin, mode=1,
                              w_tap_stride=Cin * Cout, w_ci_stride=1, w_co_stride=Cout, in_c_stride=Cout, out_c_stride=Cin)
            self.wgrad = dict(com, h_in=Hi, w_in=Wi, c_in=Cin, h_out=self.Ho, w_out=self.Wo, c_out=Cout, mode=0,
                              w_tap_stride=0, w_ci_stride=0, w_co_stride=0, in_c_stride=Cin, out_c_stride=Cout)
            self.w_shape = [k, k, Cin, Cout]
        else:
            assert (k, s) == (4, 2), 'only the 4x4 stride-2 up-convolution exists in the reference'
            self.Ho, self.Wo = 2 * Hi, 2 * Wi
            com = dict(batch=B, kh=4, kw=4, stride=2, pad_top=1, pad_left=1, epilogue=0, accumulate=0)
            # weights [kh,kw,cout,cin]; forward = gradient of a SAME stride-2 conv from the 2x-sized side
            self.fwd = dict(com, h_in=Hi, w_in=Wi, c_in=Cin, h_out=self.Ho, w_out=self.Wo, c_out=Cout, mode=1,
                            w_tap_stride=Cin * Cout, w_ci_stride=1, w_co_stride=Cin, in_c_stride=Cin, out_c_stride=Cout)
            self.dgrad = dict(com, h_in=self.Ho, w_in=self.Wo, c_in=Cout, h_out=Hi, w_out=Wi, c_out=Cin, mode=0,
                              w_tap_stride=Cin * Cout, w_ci_stride=Cin, w_co_stride=1, in_c_stride=Cout, out_c_stride=Cin)
            self.wgrad = dict(com, h_in=self.Ho, w_in=self.Wo, c_in=Cout, h_out=Hi, w_out=Wi, c_out=Cin, mode=0,
                              w_tap_stride=0, w_ci_stride=0, w_co_stride=0, in_c_stride=Cout, out_c_stride=Cin)
            self.w_shape = [4, 4, Cout, Cin]


_SYNC_BN = None      # (process group or None for the default group) when batch-norm statistics span the data-parallel ranks


def set_sync_bn(enabled, group=None):
    """Synchronised batch norm for data-parallel training: the reference normalises over the whole batch on one device
    (nets.py:263-272, ldi_enc_dec.py:198-201); with the batch sharded over ranks, per-replica statistics are a different
    function.  When enabled, every training-path BN all-reduces its per-channel (sum, sum of squares) in the forward pass and
    its (sum dz, sum dz*xhat) in the backward pass -- two [C,2] collectives per layer -- so that an N-rank step equals the
    single-device step on the global batch.  No effect when torch.distributed is not initialised or world size is 1."""
    global _SYNC_BN
    _SYNC_BN = (group,) if enabled else None


def _sync_bn_world():
    if _SYNC_BN is None:
        return None, 1
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return None, 1
    return _SYNC_BN[0], dist.get_world_size(_SYNC_BN[0])


_GRAD_SINK = [False]


class grad_sink(object):
    """Inside this context the backward passes of the conv layers add their weight gradients straight into `weights.grad` (the
    weight-gradient kernels take an accumulate flag) and hand autograd `None` for them -- no temporary, no memset, no separate
    accumulation launch per use of a variable (a training step uses every filter twice: two towers).  Only for leaves whose
    `.grad` already exists as a dense tensor, i.e. after `ParamStore.flatten()` + `zero_grad()`: what Trainer.train_step does."""

    def __enter__(self):
        self.prev = _GRAD_SINK[0]
        _GRAD_SINK[0] = True

    def __exit__(self, *exc):
        _GRAD_SINK[0] = self.prev
        return False


def _sink_of(w):
    return w if (_GRAD_SINK[0] and w.is_leaf and w.grad is not None and w.grad.is_contiguous() and w.grad.dtype == torch.float32) else None


def _dbeta_of(sums):
    """dbeta = the first column of the (sum dz, sum dz*xhat) pairs: inside the trainer's gradient sink a strided view (autograd adds it to
    the flat gradient buffer as it is -- no copy launch); a dense tensor for every other caller (e.g. one that all-reduces it)."""
    return sums[:, 0] if _GRAD_SINK[0] else sums[:, 0].contiguous()


def _wgrad_into(ctx_leaf, geo, big, small, like):
    """Weight gradient of one use of a filter: into the leaf's .grad (accumulating) when a sink is active, else a fresh tensor."""
    if ctx_leaf is not None and ctx_leaf.grad is not None:
        _wgrad(dict(geo.wgrad, accumulate=1), big, small, ctx_leaf.grad)
        return None
    dw = torch.empty_like(like)
    _wgrad(geo.wgrad, big, small, dw)
    return dw


class _ConvBNReLU(torch.autograd.Function):
    """slim.conv2d / slim.conv2d_transpose with normalizer_fn=batch_norm and activation relu (nets.py:263-272)."""

    @staticmethod
    def forward(ctx, x, w, beta, geo):
        dev = x.device
        z = torch.empty(geo.B, geo.Ho, geo.Wo, geo.Cout, dtype=torch.float32, device=dev)
        stats = torch.empty(geo.Cout, 2, dtype=torch.float32, device=dev)
        have_stats = _conv(geo.fwd, x, w, z, bn_stats=stats, store=getattr(geo, 'store', None))        # tensor-core path reduces the BN statistics in its epilogue
        y = torch.empty_like(z)
        P = geo.B * geo.Ho * geo.Wo
        group, world = _sync_bn_world()
        ctx.sync = (group, world)
        if world > 1:            # statistics over the global batch: all-reduce (sum z, sum z^2), finish in fp64
            import torch.distributed as dist
            sums = torch.empty(geo.Cout, 2, dtype=torch.float64, device=dev)
            _b200.call('lsi_b200_channel_sums_f64', _b200.ptr(z), _b200.ptr(sums), P, geo.Cout, geo.Cout,
                       _b200.ptr(_bn_workspace(dev, geo.Cout)), _b200.stream())
            dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
            mean = sums[:, 0] / (P * world)
            var = (sums[:, 1] / (P * world) - mean * mean).clamp_min(0.0)
            stats.copy_(torch.stack([mean, 1.0 / torch.sqrt(var + BN_EPS)], dim=1).float())
            have_stats = True
        _b200.call('lsi_b200_bn_relu_forward', _b200.ptr(z), _b200.ptr(beta), _b200.ptr(y), _b200.ptr(stats), P, geo.Cout,
                   geo.Cout, geo.Cout, BN_EPS, 1, 1 if have_stats else 0, _b200.ptr(_bn_workspace(dev, geo.Cout)),
                   _b200.stream())
        ctx.save_for_backward(x, w, z, y, stats, beta)
        ctx.geo = geo
        ctx.w_leaf = _sink_of(w)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, z, y, stats, beta = ctx.saved_tensors
        geo = ctx.geo
        dev = x.device
        dy = dy.contiguous()
        P = geo.B * geo.Ho * geo.Wo
        dz = torch.empty_like(z)
        sums = torch.empty(geo.Cout, 2, dtype=torch.float32, device=dev)
        group, world = ctx.sync
        if world > 1:            # (sum dz, sum dz*xhat) over the global batch; dbeta stays this rank's share (the gradient all-reduce sums it)
            import torch.distributed as dist
            args = (_b200.ptr(z), _b200.ptr(y), _b200.ptr(dy), _b200.ptr(stats), _b200.ptr(dz), _b200.ptr(sums), P, P * world, geo.Cout,
                    geo.Cout, geo.Cout, geo.Cout, geo.Cout, 1, 0)
            _b200.call('lsi_b200_bn_relu_backward_staged', *args, 1, _b200.ptr(_bn_workspace(dev, geo.Cout)), _b200.stream())
            dbeta = sums[:, 0].contiguous()
            dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
            _b200.call('lsi_b200_bn_relu_backward_staged', *args, 2, _b200.ptr(_bn_workspace(dev, geo.Cout)), _b200.stream())
        elif geo.Cout % 4 == 0 and geo.Cout <= 1024 and _BN_BWD_FAST:
            # dense fast path: reads z and dy only (ReLU mask recomputed from z exactly as the forward kernel evaluates it)
            _b200.call('lsi_b200_bn_relu_backward_z', _b200.ptr(z), _b200.ptr(beta), _b200.ptr(dy), _b200.ptr(stats), _b200.ptr(dz),
                       _b200.ptr(sums), P, geo.Cout, _b200.ptr(_bn_workspace(dev, geo.Cout)), _b200.stream())
            dbeta = _dbeta_of(sums)
        else:
            _b200.call('lsi_b200_bn_relu_backward', _b200.ptr(z), _b200.ptr(y), _b200.ptr(dy), _b200.ptr(stats), _b200.ptr(dz),
                       _b200.ptr(sums), P, geo.Cout, geo.Cout, geo.Cout, geo.Cout, geo.Cout, 1, 0,
                       _b200.ptr(_bn_workspace(dev, geo.Cout)), _b200.stream())
            dbeta = _dbeta_of(sums)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            _conv(geo.dgrad, dz, w, dx, store=getattr(geo, 'store', None))
        dw = _wgrad_into(ctx.w_leaf, geo, dz, x, w) if geo.transposed else _wgrad_into(ctx.w_leaf, geo, x, dz, w)
        return dx, dw, dbeta, None


class _ConvBNReLUCat(torch.autograd.Function):
    """_ConvBNReLU followed by tf.concat([., skip], axis=3) (nets.py:108-109, 300): the normalise + ReLU pass writes straight into the
    first channels of the concat buffer, the skip is copied behind them (one copy instead of two), and the backward reads dy as a
    channel slice of the concat gradient where it lies and returns the skip's share as a view (no copies).  Per-replica batch
    statistics only (the caller falls back to the unfused layers under synchronised batch norm)."""

    @staticmethod
    def forward(ctx, x, w, beta, skip, geo):
        dev = x.device
        z = torch.empty(geo.B, geo.Ho, geo.Wo, geo.Cout, dtype=torch.float32, device=dev)
        stats = torch.empty(geo.Cout, 2, dtype=torch.float32, device=dev)
        have_stats = _conv(geo.fwd, x, w, z, bn_stats=stats, store=getattr(geo, 'store', None))
        cb = skip.shape[3]
        cat = torch.empty(geo.B, geo.Ho, geo.Wo, geo.Cout + cb, dtype=torch.float32, device=dev)
        P = geo.B * geo.Ho * geo.Wo
        _b200.call('lsi_b200_bn_relu_forward', _b200.ptr(z), _b200.ptr(beta), _b200.ptr(cat), _b200.ptr(stats), P, geo.Cout,
                   geo.Cout, geo.Cout + cb, BN_EPS, 1, 1 if have_stats else 0, _b200.ptr(_bn_workspace(dev, geo.Cout)), _b200.stream())
        _b200.call('lsi_b200_copy_channels', _b200.ptr(skip), ctypes_offset(cat, geo.Cout), P, cb, cb, geo.Cout + cb, 0, _b200.stream())
        ctx.save_for_backward(x, w, z, stats, beta)
        ctx.geo, ctx.cb = geo, cb
        ctx.w_leaf = _sink_of(w)
        return cat

    @staticmethod
    def backward(ctx, g):
        x, w, z, stats, beta = ctx.saved_tensors
        geo, cb = ctx.geo, ctx.cb
        dev = x.device
        g = g.contiguous()
        P = geo.B * geo.Ho * geo.Wo
        dz = torch.empty_like(z)
        sums = torch.empty(geo.Cout, 2, dtype=torch.float32, device=dev)
        _b200.call('lsi_b200_bn_relu_backward_zs', _b200.ptr(z), _b200.ptr(beta), _b200.ptr(g), geo.Cout + cb, _b200.ptr(stats),
                   _b200.ptr(dz), _b200.ptr(sums), P, geo.Cout, _b200.ptr(_bn_workspace(dev, geo.Cout)), _b200.stream())
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            _conv(geo.dgrad, dz, w, dx, store=getattr(geo, 'store', None))
        dw = _wgrad_into(ctx.w_leaf, geo, dz, x, w) if geo.transposed else _wgrad_into(ctx.w_leaf, geo, x, dz, w)
        return dx, dw, _dbeta_of(sums), g[..., geo.Cout:], None


class _ConvBiasSigmoid(torch.autograd.Function):
    """The prediction conv: normalizer_fn=None, biases, activation sigmoid (nets.py:139-155)."""

    @staticmethod
    def forward(ctx, x, w, bias, geo):
        y = torch.empty(geo.B, geo.Ho, geo.Wo, geo.Cout, dtype=torch.float32, device=x.device)
        _conv(dict(geo.fwd, epilogue=2), x, w, y, bias, store=getattr(geo, 'store', None))
        ctx.save_for_backward(x, w, y)
        ctx.geo = geo
        ctx.w_leaf = _sink_of(w)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, y = ctx.saved_tensors
        geo = ctx.geo
        dev = x.device
        dy = dy.contiguous()
        dz = torch.empty_like(y)
        _b200.call('lsi_b200_sigmoid_backward', _b200.ptr(y), _b200.ptr(dy), _b200.ptr(dz), y.numel(), _b200.stream())
        sums = torch.empty(geo.Cout, 2, dtype=torch.float32, device=dev)
        _b200.call('lsi_b200_channel_sums', _b200.ptr(dz), _b200.ptr(sums), geo.B * geo.Ho * geo.Wo, geo.Cout, geo.Cout,
                   _b200.ptr(_bn_workspace(dev, geo.Cout)), _b200.stream())
        dx = torch.empty_like(x)
        _conv(geo.dgrad, dz, w, dx, store=getattr(geo, 'store', None))
        dw = _wgrad_into(ctx.w_leaf, geo, x, dz, w)
        return dx, dw, _dbeta_of(sums), None


class _ConcatChannels(torch.autograd.Function):
    """tf.concat([a, b], axis=3) (nets.py:300,...) through the strided channel-copy kernel."""

    @staticmethod
    def forward(ctx, a, b):
        B, H, W, Ca = a.shape
        Cb = b.shape[3]
        out = torch.empty(B, H, W, Ca + Cb, dtype=torch.float32, device=a.device)
        P = B * H * W
        _b200.call('lsi_b200_copy_channels', _b200.ptr(a), _b200.ptr(out), P, Ca, Ca, Ca + Cb, 0, _b200.stream())
        _b200.call('lsi_b200_copy_channels', _b200.ptr(b), ctypes_offset(out, Ca), P, Cb, Cb, Ca + Cb, 0, _b200.stream())
        ctx.dims = (B, H, W, Ca, Cb)
        return out

    @staticmethod
    def backward(ctx, g):
        B, H, W, Ca, Cb = ctx.dims
        g = g.contiguous()
        P = B * H * W
        ga = torch.empty(B, H, W, Ca, dtype=torch.float32, device=g.device)
        gb = torch.empty(B, H, W, Cb, dtype=torch.float32, device=g.device)
        _b200.call('lsi_b200_copy_channels', _b200.ptr(g), _b200.ptr(ga), P, Ca, Ca + Cb, Ca, 0, _b200.stream())
        _b200.call('lsi_b200_copy_channels', ctypes_offset(g, Ca), _b200.ptr(gb), P, Cb, Ca + Cb, Cb, 0, _b200.stream())
        return ga, gb


def ctypes_offset(t, n_floats):
    import ctypes
    return ctypes.c_void_p(t.data_ptr() + 4 * n_floats)


def _conv_layer_split(store, scope, x, cout, k, stride, reuse, transposed, defer):
    """_conv_layer on the no-grad path of the 'split' mode: split fp16-pair activations in and out, lsi_b200_conv2d_tc_s
    (RAW output + batch statistics from its epilogue), then the normalise + ReLU pass in place.  The 3-channel stem runs the
    fp32 CUDA-core kernels (TMA cannot address 12-byte pixels) and is packed afterwards."""
    pair = isinstance(x, (tuple, list))
    lib = _b200.lib()
    if transposed and _SPLIT_UPCONV_MATERIALIZE and isinstance(x, _Pending):
        # a 4x4/2 up-conv runs one CTA group per output phase: normalising on load would repeat the (ALU-heavy) split transform of
        # every halo tile four times; one in-place pass over the input in HBM is cheaper
        x = x.materialize()
    if not pair and (isinstance(x, _SplitAct) or (isinstance(x, _Pending) and isinstance(x.z, _SplitAct) and x._done is None)) \
            and _HALO and cout % 32 == 0:
        # full-resolution few-channel head layers: halo-tile kernel (resident filter bank, one TMA halo box per tile, a pending
        # producer's batch norm + ReLU applied on load -- the normalised tensor never exists in HBM)
        pend = isinstance(x, _Pending)
        a = x.z if pend else x
        B, H, W, ca = a.shape
        geo = _Geometry(transposed, B, H, W, ca, cout, k, stride)
        d = _b200.ConvDesc(**dict(geo.fwd, in_c_stride=ca))
        if lib.lsi_b200_conv2d_halo_s_supported(d) == 1:
            w = store.get(scope + '/weights', geo.w_shape, reuse, 'weights')
            beta = store.get(scope + '/BatchNorm/beta', [cout], reuse, 'beta')
            z = _SplitAct.empty(B, geo.Ho, geo.Wo, cout, a.device)
            stats = torch.empty(cout, 2, dtype=torch.float32, device=a.device)
            ws = _tc_workspace(a.device, int(lib.lsi_b200_conv2d_halo_workspace_bytes(d)))
            _vouch(store, w)
            _b200.call('lsi_b200_conv2d_halo_s', d, _b200.ptr(a.t), _b200.ptr(x.stats) if pend else None,
                       _b200.ptr(x.beta) if pend else None, _b200.ptr(w), None, None, _b200.ptr(z.t), 2, _b200.ptr(stats), BN_EPS,
                       _b200.ptr(ws), ws.numel(), _b200.stream())
            out = _Pending(z, stats, beta)
            return out if defer else out.materialize()
    srcs = [_materialize(t) for t in (x if pair else [x])]
    if not pair and _STEM_TC and isinstance(srcs[0], torch.Tensor) and srcs[0].dtype == torch.float32 and not transposed:
        # the 3-channel stem on tensor cores: im2col built in shared memory as (hi, lo) pairs, raw split output + statistics
        xin = _b200.dev_f32(srcs[0], scope + ' input')
        B, H, W, cin = xin.shape
        geo = _Geometry(False, B, H, W, cin, cout, k, stride)
        d = _b200.ConvDesc(**dict(geo.fwd, in_c_stride=cin))
        if lib.lsi_b200_conv2d_stem_tc_supported(d) == 1:
            w = store.get(scope + '/weights', geo.w_shape, reuse, 'weights')
            beta = store.get(scope + '/BatchNorm/beta', [cout], reuse, 'beta')
            z = _SplitAct.empty(B, geo.Ho, geo.Wo, cout, xin.device)
            stats = torch.empty(cout, 2, dtype=torch.float32, device=xin.device)
            ws = _tc_workspace(xin.device, int(lib.lsi_b200_conv2d_stem_tc_workspace_bytes()))
            _b200.call('lsi_b200_conv2d_stem_tc_s', d, _b200.ptr(xin), _b200.ptr(w), _b200.ptr(z.t), 2, _b200.ptr(stats), BN_EPS,
                       _b200.ptr(ws), ws.numel(), _b200.stream())
            out = _Pending(z, stats, beta)
            return out if defer else out.materialize()
    if not pair and (srcs[0].shape[3] % 32 or cout % 32):
        # channel counts the split layout cannot hold (the 3-channel stem; nz = 1000 of the non-U-Net variant): exact fp32 kernels
        xin = _b200.dev_f32(to_float(srcs[0]), scope + ' input')
        B, H, W, cin = xin.shape
        geo = _Geometry(transposed, B, H, W, cin, cout, k, stride)
        w = store.get(scope + '/weights', geo.w_shape, reuse, 'weights')
        beta = store.get(scope + '/BatchNorm/beta', [cout], reuse, 'beta')
        y = _ConvBNReLU.apply(xin, w, beta, geo)
        return y if cout % 32 else _SplitAct.pack(y)
    srcs = [t if isinstance(t, _SplitAct) else _SplitAct.pack(t) for t in srcs]
    a, b = srcs[0], (srcs[1] if pair else None)
    B, H, W, ca = a.shape
    cin = ca + (b.shape[3] if pair else 0)
    geo = _Geometry(transposed, B, H, W, cin, cout, k, stride)
    w = store.get(scope + '/weights', geo.w_shape, reuse, 'weights')
    beta = store.get(scope + '/BatchNorm/beta', [cout], reuse, 'beta')
    d = _b200.ConvDesc(**dict(geo.fwd, in_c_stride=ca))
    if lib.lsi_b200_conv2d_tc_supported(d, ca) != 1 or cout % 32:
        raise RuntimeError('lsi_b200: layer %s (%d -> %d channels) is not supported by the split tensor-core path' % (scope, cin, cout))
    z = _SplitAct.empty(B, geo.Ho, geo.Wo, cout, a.device)
    stats = torch.empty(cout, 2, dtype=torch.float32, device=a.device)
    ws = _tc_workspace(a.device, int(lib.lsi_b200_conv2d_tc_workspace_bytes(d)))
    _vouch(store, w)
    _b200.call('lsi_b200_conv2d_tc_s', d, _b200.ptr(a.t), ca, None if b is None else _b200.ptr(b.t), 0 if b is None else b.shape[3],
               _b200.ptr(w), None, None, _b200.ptr(z.t), 2, _b200.ptr(stats), BN_EPS, _b200.ptr(ws), ws.numel(), _b200.stream())
    out = _Pending(z, stats, beta)
    return out if defer else out.materialize()


def _conv_layer(store, scope, x, cout, k, stride, reuse, transposed=False, defer=False, cat_with=None):
    """conv / up-conv + batch-stat BN + ReLU.  `x` may be a pair (a, b) standing for tf.concat([a, b], axis=3): under
    no_grad the tensor-core kernel reads the two sources directly; with autograd the concat is materialised.
    Inference path (no_grad, tensor-core mode): the conv writes its RAW output and reduces the batch statistics in its
    epilogue; with defer=True the normalise + ReLU pass is left to the consumer (`_Pending`), otherwise it runs in
    place.  A `_Pending` input is normalised on load when the halo-tile kernel supports the layer."""
    if _split_infer():
        return _conv_layer_split(store, scope, x, cout, k, stride, reuse, transposed, defer)
    if not torch.is_grad_enabled() and _tc_mode():
        h = _f16_infer()
        pair = isinstance(x, (tuple, list))
        halo_pair = None
        if pair and h and _HALO_CONCAT and isinstance(x[0], _Pending) and x[0].z.dtype == torch.float16 and not transposed:
            # tf.concat([pending up-conv output, skip]) -> 3x3 conv (upcnv2b): the halo kernel reads both sources, normalising
            # the first on load, if the layer fits it (filter bank resident in shared memory)
            pb = _dev_act(_materialize(x[1]), scope + ' input')
            cin_t = x[0].shape[3] + pb.shape[3]
            g2 = _Geometry(False, x[0].shape[0], x[0].shape[1], x[0].shape[2], cin_t, cout, k, stride)
            d2 = _b200.ConvDesc(**dict(g2.fwd, in_c_stride=x[0].shape[3]))
            if (pb.dtype == torch.float16 and _HALO and _b200.lib().lsi_b200_conv2d_halo_h_supported(d2) == 1
                    and x[0].z.data_ptr() % 16 == 0 and pb.data_ptr() % 16 == 0):
                halo_pair = (x[0], pb, g2, d2)
        if halo_pair is not None:
            pa, pb, geo, d = halo_pair
            w = store.get(scope + '/weights', geo.w_shape, reuse, 'weights')
            beta = store.get(scope + '/BatchNorm/beta', [cout], reuse, 'beta')
            z = torch.empty(geo.B, geo.Ho, geo.Wo, cout, dtype=torch.float16, device=pb.device)
            stats = torch.empty(cout, 2, dtype=torch.float32, device=pb.device)
            _conv_halo(d, pa, w, z, out_stats=stats, x_b=pb, store=store)
            out = _Pending(z, stats, beta)
            return out if defer else out.materialize()
        if pair:
            a, b = (_dev_act(_materialize(t), scope + ' input') for t in x)
            ca, cin = a.shape[3], a.shape[3] + b.shape[3]
        else:
            a = x if isinstance(x, _Pending) else _dev_act(x, scope + ' input')
            b, ca, cin = None, a.shape[3], a.shape[3]
        B, H, W = a.shape[0], a.shape[1], a.shape[2]
        geo = _Geometry(transposed, B, H, W, cin, cout, k, stride)
        w = store.get(scope + '/weights', geo.w_shape, reuse, 'weights')
        beta = store.get(scope + '/BatchNorm/beta', [cout], reuse, 'beta')
        d = _b200.ConvDesc(**dict(geo.fwd, in_c_stride=ca))
        dev = a.device
        done = False
        a_raw = a.z if isinstance(a, _Pending) else a
        if not pair and _halo_ok(d, a_raw) and (a_raw.dtype == torch.float32 or isinstance(a, _Pending)):
            # the raw output of a 32-channel head layer is read by exactly one consumer, a halo-kernel conv that
            # normalises it on load and feeds fp16 MMAs: store it as fp16 (half the bytes of these byte-bound layers)
            z_dt = torch.float16 if (h or (defer and _HALO_F16_STORE and cout == 32)) else torch.float32
            z = torch.empty(B, geo.Ho, geo.Wo, cout, dtype=z_dt, device=dev)
            stats = torch.empty(cout, 2, dtype=torch.float32, device=dev)
            _conv_halo(d, a, w, z, out_stats=stats, store=store)
            done = True
        else:
            a = _materialize(a)
            if h and _tc_ok(d, ca, a, b):
                a, b = a.half() if a.dtype != torch.float16 else a, (b.half() if (b is not None and b.dtype != torch.float16) else b)
                z = torch.empty(B, geo.Ho, geo.Wo, cout, dtype=torch.float16, device=dev)
                stats = torch.empty(cout, 2, dtype=torch.float32, device=dev)
                nws = int(_b200.lib().lsi_b200_conv2d_tc_workspace_bytes(d))
                ws = _tc_workspace(dev, nws)
                _vouch(store, w)
                _b200.call('lsi_b200_conv2d_tc_h', d, _b200.ptr(a), ca, _b200.ptr(b), 0 if b is None else b.shape[-1], _b200.ptr(w),
                           None, _b200.ptr(z), 1, _b200.ptr(stats), BN_EPS, _b200.ptr(ws), ws.numel(), _b200.stream())
                done = True
            elif (h and not pair and a.dtype == torch.float32 and _STEM_TC
                  and _b200.lib().lsi_b200_conv2d_stem_tc_supported(d) == 1):
                # the 3-channel stem on tensor cores (im2col built in shared memory), raw fp16 output + statistics
                z = torch.empty(B, geo.Ho, geo.Wo, cout, dtype=torch.float16, device=dev)
                stats = torch.empty(cout, 2, dtype=torch.float32, device=dev)
                nws = int(_b200.lib().lsi_b200_conv2d_stem_tc_workspace_bytes())
                ws = _tc_workspace(dev, nws)
                _b200.call('lsi_b200_conv2d_stem_tc', d, _b200.ptr(a), _b200.ptr(w), _b200.ptr(z), 1, _b200.ptr(stats), BN_EPS,
                           _b200.ptr(ws), ws.numel(), _b200.stream())
                done = True
            elif _tc_ok(d, ca, a, b):
                z = torch.empty(B, geo.Ho, geo.Wo, cout, dtype=torch.float32, device=dev)
                stats = torch.empty(cout, 2, dtype=torch.float32, device=dev)
                have = _conv(dict(geo.fwd, in_c_stride=ca), a, w, z, bn_stats=stats, inp_b=b, c_in_a=ca)
                if not have:       # cannot happen on the tensor-core path; keep the contract explicit
                    raise RuntimeError('lsi_b200: conv epilogue did not produce batch statistics')
                done = True
            x = (a, b) if pair else a
        if done:
            out = _Pending(z, stats, beta)
            return out if defer else out.materialize()
    if isinstance(x, (tuple, list)):
        a, b = (_b200.dev_f32(t, scope + ' input') for t in x)
        x = _ConcatChannels.apply(a, b)
    if isinstance(x, torch.Tensor) and x.dtype == torch.float16:
        x = x.float()
    x = _b200.dev_f32(x, scope + ' input')
    B, H, W, cin = x.shape
    geo = _Geometry(transposed, B, H, W, cin, cout, k, stride)
    w = store.get(scope + '/weights', geo.w_shape, reuse, 'weights')
    beta = store.get(scope + '/BatchNorm/beta', [cout], reuse, 'beta')
    geo.store = store
    if (cat_with is not None and torch.is_grad_enabled() and _BN_BWD_FAST and _sync_bn_world()[1] == 1 and isinstance(cat_with, torch.Tensor)
            and cat_with.is_cuda and cat_with.dtype == torch.float32 and cat_with.is_contiguous() and cat_with.dim() == 4
            and tuple(cat_with.shape[:3]) == (B, geo.Ho, geo.Wo) and cout % 4 == 0 and cout <= 1024 and cat_with.shape[3] % 4 == 0):
        # training path: this layer's output is only ever read as the first part of tf.concat([., cat_with], axis=3): build the concat
        # directly (the caller recognises the tag and does not concatenate again)
        out = _ConvBNReLUCat.apply(x, w, beta, cat_with, geo)
        out._lsi_cat_done = True
        return out
    y = _ConvBNReLU.apply(x, w, beta, geo)
    return y.half() if _f16_infer() else y       # (the 3-channel stem: fp32 CUDA-core conv, handed on as fp16)


# ---------------------------------------------------------------------------------------------------------------------
# the reference's public functions
# ---------------------------------------------------------------------------------------------------------------------
def decoder_simple(feat, nconv=7, is_training=True, skip_feat=None, reuse=False, _scope='decoder', _store=None,
                   _defer=False):
    """nets.py:73-114 -- nconv x [4x4 s2 up-conv -> concat skip -> 3x3 conv].  Returns (feat, end_points)."""
    _require_training(is_training)
    store = _store or get_default_store()
    n_filters = [32, 64, 128, 256] + [512] * max(nconv - 4, 0)
    end_points = {}
    if isinstance(feat, torch.Tensor) and feat.dim() == 2:        # B x nz bottleneck code (nets.py:103-104)
        feat = feat[:, None, None, :]
    for nc in range(nconv, 0, -1):
        n_filt = n_filters[nc - 1]
        skip = skip_feat[-nc + 1] if (nc > 1 and skip_feat is not None) else None
        feat = _conv_layer(store, '%s/upcnv%d' % (_scope, nc), feat, n_filt, 4, 2, reuse, transposed=True, defer=_defer, cat_with=skip)
        if skip is not None and not getattr(feat, '_lsi_cat_done', False):
            feat = (feat, skip)                                  # tf.concat([feat, skip], axis=3), nets.py:108-109
        feat = _conv_layer(store, '%s/upcnv%db' % (_scope, nc), feat, n_filt, 3, 1, reuse, defer=_defer)
        end_points['%s/upcnv%db' % (_scope, nc)] = feat
    return feat, end_points


def pixelwise_predictor(feat, nc=3, n_layers=1, n_layerwise_steps=0, skip_feat=None, reuse=False, is_training=True,
                        _scope='pixelwise_pred', _store=None, _out_hw=None, _out_scale=None):
    """nets.py:117-161 -- per layer its own decoder_simple, then a 3x3 conv + bias + sigmoid.  Returns
    (preds [L,B,H,W,nc], end_points)."""
    _require_training(is_training)
    store = _store or get_default_store()
    preds = []
    packed = None         # inference path: every head writes its slice of one [L,B,H,W,nc] tensor (no tf.stack copy)
    for l in range(n_layers):
        base = '%s/upsample_%d' % (_scope, l)
        # _defer: inside this function nothing but the next conv sees the decoder features, so (inference path) their
        # batch norm + ReLU may stay pending and be applied on load by the consumer
        feat_l, _ = decoder_simple(feat, nconv=n_layerwise_steps, skip_feat=skip_feat, reuse=reuse, is_training=is_training,
                                   _scope=base + '/decoder', _store=store, _defer=True)
        B, H, W, cin = feat_l.shape
        geo = _Geometry(False, B, H, W, cin, nc, 3, 1, out_hw=_out_hw)
        w = store.get('%s/pred_%d/weights' % (base, l), geo.w_shape, reuse, 'weights')
        b = store.get('%s/pred_%d/biases' % (base, l), [nc], reuse, 'biases')
        dp = _b200.ConvDesc(**dict(geo.fwd, epilogue=2))
        if _split_infer() and isinstance(feat_l.z if isinstance(feat_l, _Pending) else feat_l, _SplitAct):
            # 'split' mode: bias + sigmoid + per-channel output factor in the epilogue of the split tensor-core conv, fp32 output
            if packed is None and l == 0:
                packed = torch.empty(n_layers, B, geo.Ho, geo.Wo, nc, dtype=torch.float32, device=feat_l.device)
            y = packed[l] if packed is not None else torch.empty(B, geo.Ho, geo.Wo, nc, dtype=torch.float32, device=feat_l.device)
            lib = _b200.lib()
            pend = isinstance(feat_l, _Pending) and feat_l._done is None
            if _HALO and nc <= 4 and lib.lsi_b200_conv2d_halo_s_supported(dp) == 1:
                fa = feat_l.z if isinstance(feat_l, _Pending) else feat_l
                ws = _tc_workspace(fa.device, int(lib.lsi_b200_conv2d_halo_workspace_bytes(dp)))
                _vouch(store, w)
                _b200.call('lsi_b200_conv2d_halo_s', dp, _b200.ptr(fa.t), _b200.ptr(feat_l.stats) if pend else None,
                           _b200.ptr(feat_l.beta) if pend else None, _b200.ptr(w), _b200.ptr(b), _b200.ptr(_out_scale), _b200.ptr(y), 0,
                           None, BN_EPS, _b200.ptr(ws), ws.numel(), _b200.stream())
            else:
                fm = _materialize(feat_l)
                ws = _tc_workspace(fm.device, int(lib.lsi_b200_conv2d_tc_workspace_bytes(dp)))
                _vouch(store, w)
                _b200.call('lsi_b200_conv2d_tc_s', dp, _b200.ptr(fm.t), cin, None, 0, _b200.ptr(w), _b200.ptr(b), _b200.ptr(_out_scale),
                           _b200.ptr(y), 0, None, BN_EPS, _b200.ptr(ws), ws.numel(), _b200.stream())
            preds.append(y)
        elif _halo_ok(dp, feat_l.z if isinstance(feat_l, _Pending) else feat_l):
            if packed is None and l == 0:
                packed = torch.empty(n_layers, B, geo.Ho, geo.Wo, nc, dtype=torch.float32, device=feat_l.device)
            y = packed[l] if packed is not None else torch.empty(B, geo.Ho, geo.Wo, nc, dtype=torch.float32, device=feat_l.device)
            _conv_halo(dp, feat_l, w, y, bias=b, out_scale=_out_scale, store=store)
            preds.append(y)
        else:
            fm = _materialize(feat_l)
            geo.store = store
            y = _ConvBiasSigmoid.apply(fm.float() if fm.dtype != torch.float32 else fm, w, b, geo)
            preds.append(y if _out_scale is None else y * _out_scale)
    if packed is not None and len(preds) == n_layers and all(p_.data_ptr() == packed[i].data_ptr() for i, p_ in enumerate(preds)):
        return packed, {}
    return torch.stack(preds, dim=0), {}


def ldi_predictor(feat, n_layers=1, reuse=False, n_layerwise_steps=0, skip_feat=None, pred_masks=False, is_training=True,
                  _store=None, _out_hw=None, _disp_scale=None):
    """nets.py:164-208.  Returns ldi = [textures [L,B,H,W,3], masks [L,B,H,W,1], disps [L,B,H,W,1]].  The textures and
    disparities are channel views of the packed [L,B,H,W,nc] head output (no copy); with pred_masks=False the masks
    are all ones and tagged so (the renderer and the losses then never read them)."""
    nc = 3 + 1 + (1 if pred_masks else 0)
    # _disp_scale (inference, no predicted masks): `disps *= max_disp` (ldi_enc_dec.py:213) as a per-channel output factor
    # (1, 1, 1, max_disp) of the prediction conv instead of a separate pass over the packed head output
    out_scale = None
    if _disp_scale is not None and not pred_masks and not torch.is_grad_enabled():
        key = (str(feat.device), float(_disp_scale))
        if key not in _OUT_SCALE_CACHE:           # created once: a per-call host->device copy would stall the launch queue
            _OUT_SCALE_CACHE[key] = torch.tensor([1.0, 1.0, 1.0, float(_disp_scale)], dtype=torch.float32, device=feat.device)
        out_scale = _OUT_SCALE_CACHE[key]
    pred, _ = pixelwise_predictor(feat, nc=nc, n_layers=n_layers, n_layerwise_steps=n_layerwise_steps, skip_feat=skip_feat,
                                  reuse=reuse, is_training=is_training, _scope='ldi_tex_disp/pixelwise_pred', _store=_store,
                                  _out_hw=_out_hw, _out_scale=out_scale)
    if pred_masks:
        tex, masks, disps = pred[..., 0:3], pred[..., 3:4], pred[..., 4:5]
        masks = nn_helpers.enforce_bg_occupied(torch.sigmoid(masks))      # sigmoid applied twice, as nets.py:143,202
    else:
        tex, disps = pred[..., 0:3], pred[..., 3:4]
        masks = torch.ones(disps.shape, dtype=torch.float32, device=pred.device)
        masks._lsi_all_ones = True
    return [tex, masks, disps]


def _fc_layer(store, scope, x, cout, reuse):
    """slim.fully_connected with batch_norm + ReLU (nets.py:43-49,67-68): x [B,K] @ weights [K,cout] (the TF variable shape),
    batch statistics over the batch, beta only.  Runs as a 1x1 convolution on a [B,1,1,K] tensor."""
    x = _b200.dev_f32(to_float(x), scope + ' input')
    B, K = x.shape
    w = store.get(scope + '/weights', [K, cout], reuse, 'weights')
    beta = store.get(scope + '/BatchNorm/beta', [cout], reuse, 'beta')
    geo = _Geometry(False, B, 1, 1, K, cout, 1, 1)
    return _ConvBNReLU.apply(x.view(B, 1, 1, K), w.view(1, 1, K, cout), beta, geo).view(B, cout)


def encoder_simple(inp_img, nz=1000, is_training=True, reuse=False, _store=None):
    """nets.py:29-70 -- the 14-convolution encoder under scope `encoder`, flatten, fully connected stack (2nz, nz, nz).
    Returns (enc [B,nz], end_points)."""
    _require_training(is_training)
    store = _store or get_default_store()
    x = _b200.dev_f32(inp_img, 'inp_img')
    if x.dim() != 4 or x.shape[3] != 3:
        raise RuntimeError('lsi_b200: inp_img must be [B,H,W,3], got %s' % (tuple(x.shape),))
    if x.shape[1] % 128 or x.shape[2] % 128:
        raise ValueError('the encoder needs H and W to be multiples of 128 (seven stride-2 convolutions on even sizes), got %dx%d'
                         % (x.shape[1], x.shape[2]))
    ep = {}
    for name, k, stride, cout in ENC:
        x = _conv_layer(store, 'encoder/%s' % name, x, cout, k, stride, reuse)
        ep[name] = x
    x = to_float(x)
    x = x.reshape(x.shape[0], -1)                                  # slim.flatten: NHWC order
    for i, n_out in enumerate([2 * nz, nz, nz]):                   # slim.stack scopes fc/fc_1 .. fc_3
        x = _fc_layer(store, 'encoder/fc/fc_%d' % (i + 1), x, n_out, reuse)
        ep['fc_%d' % (i + 1)] = x
    return x, ep


def encoder_decoder_simple(inp_img, nz=1000, nupconv=8, is_training=True, reuse=False, nl_diff_enc_dec=0, _store=None):
    """nets.py:211-241 (`--use_unet=false`, ldi_enc_dec.py:202-205).  Returns (feat [B,nz], feat_dec, skip_feat=None, end_points)."""
    store = _store or get_default_store()
    feat, enc_ep = encoder_simple(inp_img, nz=nz, is_training=is_training, reuse=reuse, _store=store)
    feat_dec, dec_ep = decoder_simple(feat, nconv=nupconv - nl_diff_enc_dec, is_training=is_training, reuse=reuse, _store=store)
    return feat, feat_dec, None, dict(enc_ep, **dec_ep)


def encoder_decoder_unet(inp_img, nz=1000, is_training=True, reuse=False, nl_diff_enc_dec=0, _store=None):
    """nets.py:244-348.  Returns (feat, feat_dec, skip_feat, end_points) like the reference; `feat` (the FC features
    the reference builds but never runs) is None.  H and W must be multiples of 128 (nets.py:298-300)."""
    _require_training(is_training)
    store = _store or get_default_store()
    x = _b200.dev_f32(inp_img, 'inp_img')
    if x.dim() != 4 or x.shape[3] != 3:
        raise RuntimeError('lsi_b200: inp_img must be [B,H,W,3], got %s' % (tuple(x.shape),))
    if x.shape[1] % 128 or x.shape[2] % 128:
        raise ValueError('U-Net needs H and W to be multiples of 128 (nets.py:298-300), got %dx%d; see '
                         'nets.pad_to_legal()' % (x.shape[1], x.shape[2]))
    ep = {}
    sc = 'encoder_decoder_unet'
    for name, k, stride, cout in ENC:
        x = _conv_layer(store, '%s/%s' % (sc, name), x, cout, k, stride, reuse)
        ep[name] = x
    skip_feat = [ep['cnv6b'], ep['cnv5b'], ep['cnv4b'], ep['cnv3b'], ep['cnv2b'], ep['cnv1b']]
    feats_dec = []
    feat = ep['cnv7b']
    for k, cout, skip in DEC[:7 - nl_diff_enc_dec]:
        up = _conv_layer(store, '%s/upcnv%d' % (sc, k), feat, cout, 4, 2, reuse, transposed=True,
                         cat_with=ep[skip] if skip is not None else None)
        if skip is not None and not getattr(up, '_lsi_cat_done', False):
            up = (up, ep[skip])                                  # tf.concat([upcnv, skip], axis=3), nets.py:300,...
        feat = _conv_layer(store, '%s/icnv%d' % (sc, k), up, cout, 3, 1, reuse)
        ep['icnv%d' % k] = feat
        feats_dec.append(feat)
    return None, feats_dec[-1], skip_feat, ep


def pad_to_legal(img):
    """Policy for sizes the reference U-Net cannot run (64x64, 128x416, 256x832: BASELINE configs): zero-pad bottom/right
    to the next multiple of 128 and crop the prediction back.  Returns (padded image, (H, W))."""
    B, H, W, C = img.shape
    Hp, Wp = -(-H // 128) * 128, -(-W // 128) * 128
    if (Hp, Wp) == (H, W):
        return img, (H, W)
    out = torch.zeros(B, Hp, Wp, C, dtype=img.dtype, device=img.device)
    out[:, :H, :W] = img
    return out, (H, W)


def _require_training(is_training):
    if not is_training:
        raise NotImplementedError('inference-mode batch norm: the reference never updates its moving averages '
                                  '(train_utils.py:107-117) and evaluates with batch statistics (ldi_pred_eval.py:45)')
