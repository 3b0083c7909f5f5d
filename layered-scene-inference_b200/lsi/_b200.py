"""ctypes binding of the C-ABI library (include/lsi_b200.h).  Everything the Python layer does on the device
goes through `call()`; torch is used only for memory, streams and autograd bookkeeping."""
import ctypes
import os

import torch

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), '_lib', 'liblsi_b200.so')
_lib = None

c_float_p = ctypes.c_void_p   # device pointers travel as void*


class SplatDesc(ctypes.Structure):
    """struct lsi_b200_splat_desc (include/lsi_b200.h)."""
    _fields_ = [('n_layers', ctypes.c_int), ('batch', ctypes.c_int), ('h_s', ctypes.c_int), ('w_s', ctypes.c_int),
                ('h_t', ctypes.c_int), ('w_t', ctypes.c_int), ('trg_downsampling', ctypes.c_float),
                ('bg_layer_disp', ctypes.c_float), ('max_disp', ctypes.c_float), ('zbuf_scale', ctypes.c_float),
                ('compose_layers', ctypes.c_int), ('compute_trg_disp', ctypes.c_int),
                ('tex_px_stride', ctypes.c_int), ('disp_px_stride', ctypes.c_int), ('mask_px_stride', ctypes.c_int),
                ('variant', ctypes.c_int)]


class ConvDesc(ctypes.Structure):
    """struct lsi_b200_conv_desc (include/lsi_b200.h)."""
    _fields_ = [(n, ctypes.c_int) for n in
                ('batch', 'h_in', 'w_in', 'c_in', 'h_out', 'w_out', 'c_out', 'kh', 'kw', 'stride', 'pad_top', 'pad_left',
                 'mode', 'w_tap_stride', 'w_ci_stride', 'w_co_stride', 'in_c_stride', 'out_c_stride', 'epilogue',
                 'accumulate')]


# name -> (restype, argtypes); mirrors include/lsi_b200.h one to one (tests/test_capi_symbols.py checks)
_P, _I, _F, _LL, _SZ = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_longlong, ctypes.c_size_t
_DP = ctypes.POINTER(SplatDesc)
_CP = ctypes.POINTER(ConvDesc)
SIGNATURES = {
    'lsi_b200_version': (_I, []),
    'lsi_b200_last_error': (ctypes.c_char_p, []),
    'lsi_b200_launch_count': (ctypes.c_ulonglong, []),
    'lsi_b200_kernel_timing_enable': (_I, [_I]),
    'lsi_b200_set_weight_version': (None, [ctypes.c_ulonglong]),
    'lsi_b200_weight_cache_clear': (None, []),
    'lsi_b200_kernel_timing_collect': (_I, [_P, _P]),
    'lsi_b200_projection_matrix': (_I, [_P, _P, _P, _P, _I, _I, _P, _P]),
    'lsi_b200_forward_splat_workspace_bytes': (_SZ, [_DP]),
    'lsi_b200_forward_splat_backward_workspace_bytes': (_SZ, [_DP]),
    'lsi_b200_forward_splat': (_I, [_DP] + [_P] * 13 + [_P, _SZ, _P]),
    'lsi_b200_forward_splat_backward': (_I, [_DP] + [_P] * 18 + [_P, _SZ, _P]),
    'lsi_b200_forward_splat_host': (_I, [_DP] + [_P] * 10),
    'lsi_b200_splat': (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    'lsi_b200_splat_backward': (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    'lsi_b200_bilinear': (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    'lsi_b200_bilinear_backward': (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    'lsi_b200_bilinear_corners': (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    'lsi_b200_loss_partials_count': (_SZ, []),
    'lsi_b200_zbuf_composition_loss': (_I, [_P, _P, _P, _P, _I, _LL, _F, _F, _F, _P, _P, _P]),
    'lsi_b200_zbuf_composition_loss_backward': (_I, [_P, _P, _P, _P, _I, _LL, _F, _F, _F, _P, _P, _P, _P, _P]),
    'lsi_b200_photo_loss': (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _F, _P, _P, _P]),
    'lsi_b200_photo_loss_backward': (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _F, _P, _P, _P]),
    'lsi_b200_disp_smoothness_loss': (_I, [_P, _I, _I, _I, _P, _P, _P]),
    'lsi_b200_disp_smoothness_loss_backward': (_I, [_P, _I, _I, _I, _P, _P, _P]),
    'lsi_b200_decreasing_disp_loss': (_I, [_P, _I, _LL, _P, _P, _P]),
    'lsi_b200_decreasing_disp_loss_backward': (_I, [_P, _I, _LL, _P, _P, _P]),
    'lsi_b200_conv2d': (_I, [_CP, _P, _P, _P, _P, _P]),
    'lsi_b200_conv2d_wgrad': (_I, [_CP, _P, _P, _P, _P]),
    'lsi_b200_conv2d_thin_supported': (_I, [_CP]),
    'lsi_b200_conv2d_thin': (_I, [_CP, _P, _P, _P, _P]),
    'lsi_b200_conv2d_stem_wgrad_supported': (_I, [_CP]),
    'lsi_b200_conv2d_stem_wgrad': (_I, [_CP, _P, _P, _P, _P]),
    'lsi_b200_bn_relu_backward_z': (_I, [_P, _P, _P, _P, _P, _P, _LL, _I, _P, _P]),
    'lsi_b200_bn_relu_backward_zs': (_I, [_P, _P, _P, _I, _P, _P, _P, _LL, _I, _P, _P]),
    'lsi_b200_conv2d_wgrad_tc_supported': (_I, [_CP]),
    'lsi_b200_conv2d_wgrad_tc': (_I, [_CP, _P, _P, _P, _P]),
    'lsi_b200_conv2d_tc_supported': (_I, [_CP, _I]),
    'lsi_b200_conv2d_tc_workspace_bytes': (_SZ, [_CP]),
    'lsi_b200_conv2d_tc': (_I, [_CP, _P, _I, _P, _I, _P, _P, _P, _P, _SZ, _P]),
    'lsi_b200_conv2d_halo_supported': (_I, [_CP]),
    'lsi_b200_conv2d_halo_workspace_bytes': (_SZ, [_CP]),
    'lsi_b200_conv2d_halo': (_I, [_CP, _P, _P, _P, _P, _P, _P, _P, _P, _F, _P, _SZ, _P]),
    'lsi_b200_conv2d_halo_h_supported': (_I, [_CP]),
    'lsi_b200_conv2d_halo_h': (_I, [_CP, _P, _P, _I, _I, _I, _P, _P, _P, _P, _P, _P, _I, _P, _F, _P, _SZ, _P]),
    'lsi_b200_conv2d_tc_h': (_I, [_CP, _P, _I, _P, _I, _P, _P, _P, _I, _P, _F, _P, _SZ, _P]),
    'lsi_b200_conv2d_tc_s': (_I, [_CP, _P, _I, _P, _I, _P, _P, _P, _P, _I, _P, _F, _P, _SZ, _P]),
    'lsi_b200_conv2d_halo_s_supported': (_I, [_CP]),
    'lsi_b200_conv2d_halo_s': (_I, [_CP, _P, _P, _P, _P, _P, _P, _P, _I, _P, _F, _P, _SZ, _P]),
    'lsi_b200_split_convert': (_I, [_P, _I, _P, _P, _P, _I, _LL, _I, _P]),
    'lsi_b200_conv2d_stem_tc_supported': (_I, [_CP]),
    'lsi_b200_conv2d_stem_tc_workspace_bytes': (_SZ, []),
    'lsi_b200_conv2d_stem_tc': (_I, [_CP, _P, _P, _P, _I, _P, _F, _P, _SZ, _P]),
    'lsi_b200_conv2d_stem_tc_s': (_I, [_CP, _P, _P, _P, _I, _P, _F, _P, _SZ, _P]),
    'lsi_b200_bn_relu_apply_h': (_I, [_P, _I, _P, _P, _P, _LL, _I, _P]),
    'lsi_b200_bn_workspace_bytes': (_SZ, [_I]),
    'lsi_b200_conv2d_tc_bnstats': (_I, [_CP, _P, _I, _P, _I, _P, _P, _P, _F, _P, _SZ, _P]),
    'lsi_b200_bn_relu_forward': (_I, [_P, _P, _P, _P, _LL, _I, _I, _I, _F, _I, _I, _P, _P]),
    'lsi_b200_bn_relu_backward': (_I, [_P, _P, _P, _P, _P, _P, _LL, _I, _I, _I, _I, _I, _I, _I, _P, _P]),
    'lsi_b200_bn_relu_backward_staged': (_I, [_P, _P, _P, _P, _P, _P, _LL, _LL, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P]),
    'lsi_b200_channel_sums': (_I, [_P, _P, _LL, _I, _I, _P, _P]),
    'lsi_b200_channel_sums_f64': (_I, [_P, _P, _LL, _I, _I, _P, _P]),
    'lsi_b200_copy_channels': (_I, [_P, _P, _LL, _I, _I, _I, _I, _P]),
    'lsi_b200_sigmoid_backward': (_I, [_P, _P, _P, _LL, _P]),
    'lsi_b200_render_planes': (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _F, _F, _P, _P, _P, _P, _P]),
    'lsi_b200_procedural_texture': (_I, [_P, _P, _I, _I, _I, _I, _P, _P, _P]),
    'lsi_b200_box_downsample': (_I, [_P, _P, _I, _I, _I, _I, _I, _P]),
    'lsi_b200_area_resize_u8': (_I, [_P, _I, _I, _I, _P, _I, _I, _I, _P]),
    'lsi_b200_adam_step': (_I, [_P, _P, _P, _P, _LL, _F, _F, _F, _F, _LL, _F, _P]),
}


def lib():
    """Load the C-ABI library (once).  Fails loudly when it has not been built -- there is no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError('lsi_b200: %s is missing -- build it with `python -c "import __graft_entry__ as g; '
                               'g.build()"` (or `make -C layered-scene-inference_b200/csrc`); there is no CPU/torch '
                               'fallback for the hot path' % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def call(name, *args):
    rc = getattr(lib(), name)(*args)
    if rc != 0:
        raise RuntimeError('%s failed (%d): %s' % (name, rc, lib().lsi_b200_last_error().decode()))


def launch_count():
    return int(lib().lsi_b200_launch_count())


def stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def dev_f32(t, name, contiguous=True):
    """Validate a hot-path tensor: CUDA, fp32 (and contiguous unless the caller handles strides)."""
    if not isinstance(t, torch.Tensor):
        raise RuntimeError('lsi_b200: %s must be a torch.Tensor, got %r' % (name, type(t)))
    if not t.is_cuda:
        raise RuntimeError('lsi_b200: %s is on %s -- the hot path runs on CUDA only (no CPU fallback)' % (name, t.device))
    if t.dtype != torch.float32:
        raise RuntimeError('lsi_b200: %s must be float32, got %s' % (name, t.dtype))
    if contiguous and not t.is_contiguous():
        t = t.contiguous()
    return t


_partials = {}


def partials(device):
    """Per-device scratch for the two-stage loss reductions."""
    key = (device.index, torch.cuda.current_stream().cuda_stream)
    if key not in _partials:
        _partials[key] = torch.empty(int(lib().lsi_b200_loss_partials_count()), dtype=torch.float32, device=device)
    return _partials[key]
