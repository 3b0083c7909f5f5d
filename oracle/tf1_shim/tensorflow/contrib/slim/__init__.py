"""TEST INFRASTRUCTURE ONLY -- a minimal eager tf.contrib.slim for the reference's lsi/nnutils/nets.py, so that
oracle/gen_golden.py can execute the reference's OWN network wiring (layer order, channel counts, strides, skip
indices, scopes) on torch CPU tensors.

Restated [TF1.4 slim] semantics (NOT pinned by the reference itself):
  conv2d / conv2d_transpose: padding='SAME' (asymmetric for stride 2 on even sizes: pad_before = total//2), NHWC,
      weights [kh,kw,cin,cout] / [kh,kw,cout,cin]; when normalizer_fn is given there is NO bias and the normalizer is
      applied before the activation; otherwise biases (zeros init) are added.
  batch_norm: center=True (beta), scale=False (no gamma), epsilon=0.001, is_training=True -> batch mean and BIASED
      batch variance over (N,H,W); the moving averages are never consulted in training mode.
  fully_connected: weights [in,out], same normalizer/bias rule.  flatten: [B,-1].  stack: repeated layer with scopes
      name/name_1..n (slim.stack naming).
Variables live in `tensorflow._VARS` (name -> torch tensor); missing ones are created with Xavier-uniform / zeros.
"""
import contextlib
import functools
import math
import zlib

import numpy as np
import torch
import torch.nn.functional as F

import tensorflow as tf

_ARG_STACK = [{}]


def l2_regularizer(scale):
    return ('l2', scale)


@contextlib.contextmanager
def arg_scope(funcs, **kwargs):
    top = dict(_ARG_STACK[-1])
    for f in funcs:
        key = getattr(f, '_slim_name', f.__name__)
        merged = dict(top.get(key, {}))
        merged.update(kwargs)
        top[key] = merged
    _ARG_STACK.append(top)
    try:
        yield top
    finally:
        _ARG_STACK.pop()


def _with_arg_scope(fn):
    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        merged = dict(_ARG_STACK[-1].get(fn.__name__, {}))
        merged.update(kwargs)
        return fn(*args, **merged)
    wrapper._slim_name = fn.__name__
    return wrapper


def _xavier(shape, fan_in, fan_out, name):
    limit = math.sqrt(6.0 / (fan_in + fan_out))
    rs = np.random.RandomState(zlib.crc32(name.encode()) % (2 ** 31))
    return torch.tensor(rs.uniform(-limit, limit, shape), dtype=tf._FLOAT)


def _same_pad(size, k, s):
    out = -(-size // s)
    total = max((out - 1) * s + k - size, 0)
    return total // 2, total - total // 2


@_with_arg_scope
def batch_norm(inputs, is_training=True, scope=None, **_ignored):
    x = tf._raw(inputs)
    with tf.variable_scope(scope or 'BatchNorm'):
        beta = tf.get_variable('beta', [x.shape[-1]], init=lambda shp: torch.zeros(shp, dtype=tf._FLOAT))
    assert is_training, 'the reference always runs BN in training mode (ldi_pred_eval.py:45)'
    dims = list(range(x.dim() - 1))
    mean = x.mean(dim=dims, keepdim=True)
    var = ((x - mean) ** 2).mean(dim=dims, keepdim=True)          # biased
    return tf.Tensor((x - mean) / torch.sqrt(var + 0.001) + beta)


def _finish(y, cout, normalizer_fn, normalizer_params, activation_fn):
    if normalizer_fn is not None:
        y = tf._raw(normalizer_fn(tf.Tensor(y), **(normalizer_params or {})))
    else:
        b = tf.get_variable('biases', [cout], init=lambda shp: torch.zeros(shp, dtype=tf._FLOAT))
        y = y + b
    if activation_fn is not None:
        y = tf._raw(activation_fn(tf.Tensor(y)))
    return tf.Tensor(y)


@_with_arg_scope
def conv2d(inputs, num_outputs, kernel_size, stride=1, scope=None, normalizer_fn=None, normalizer_params=None,
           activation_fn=tf.nn.relu, weights_regularizer=None, outputs_collections=None, **_ignored):
    x = tf._raw(inputs)
    kh, kw = kernel_size
    cin = x.shape[-1]
    with tf.variable_scope(scope):
        w = tf.get_variable('weights', [kh, kw, cin, num_outputs],
                            init=lambda shp: _xavier(shp, kh * kw * cin, kh * kw * num_outputs, tf.current_scope() + '/weights'))
        pt, pb = _same_pad(x.shape[1], kh, stride)
        pl, pr = _same_pad(x.shape[2], kw, stride)
        xn = F.pad(x.permute(0, 3, 1, 2), (pl, pr, pt, pb))
        y = F.conv2d(xn, w.permute(3, 2, 0, 1), stride=stride).permute(0, 2, 3, 1)
        return _finish(y, num_outputs, normalizer_fn, normalizer_params, activation_fn)


@_with_arg_scope
def conv2d_transpose(inputs, num_outputs, kernel_size, stride=1, scope=None, normalizer_fn=None,
                     normalizer_params=None, activation_fn=tf.nn.relu, weights_regularizer=None,
                     outputs_collections=None, **_ignored):
    x = tf._raw(inputs)
    kh, kw = kernel_size
    cin = x.shape[-1]
    assert (kh, kw, stride) == (4, 4, 2), 'shim covers the 4x4 stride-2 SAME up-convolution only'
    with tf.variable_scope(scope):
        w = tf.get_variable('weights', [kh, kw, num_outputs, cin],
                            init=lambda shp: _xavier(shp, kh * kw * cin, kh * kw * num_outputs, tf.current_scope() + '/weights'))
        # gradient of conv(k=4, s=2, SAME): SAME pads (1,1) on the 2n-sized side  ==  conv_transpose2d(padding=1)
        y = F.conv_transpose2d(x.permute(0, 3, 1, 2), w.permute(3, 2, 0, 1), stride=2, padding=1).permute(0, 2, 3, 1)
        return _finish(y, num_outputs, normalizer_fn, normalizer_params, activation_fn)


@_with_arg_scope
def fully_connected(inputs, num_outputs, scope=None, normalizer_fn=None, normalizer_params=None,
                    activation_fn=tf.nn.relu, weights_regularizer=None, outputs_collections=None, **_ignored):
    x = tf._raw(inputs)
    cin = x.shape[-1]
    with tf.variable_scope(scope):
        w = tf.get_variable('weights', [cin, num_outputs],
                            init=lambda shp: _xavier(shp, cin, num_outputs, tf.current_scope() + '/weights'))
        return _finish(x @ w, num_outputs, normalizer_fn, normalizer_params, activation_fn)


def flatten(inputs, scope=None):
    x = tf._raw(inputs)
    return tf.Tensor(x.reshape(x.shape[0], -1))


def stack(inputs, layer, stack_args, scope=None, **kwargs):
    out = inputs
    with tf.variable_scope(scope):
        for i, a in enumerate(stack_args):
            out = layer(out, a, scope='%s_%d' % (scope, i + 1), **kwargs)
    return out
