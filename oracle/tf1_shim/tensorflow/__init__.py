"""TEST INFRASTRUCTURE ONLY -- a tiny eager stand-in for the TensorFlow-1.4 API surface that the
reference's hot-path modules touch (lsi/geometry/{ldi,sampling,projection}.py, lsi/nnutils/helpers.py,
lsi/loss/loss.py and the loss glue in ldi_enc_dec.py:265-410).

TensorFlow is not installable in this image, so `oracle/gen_golden.py` puts this
directory on sys.path *in front of* /root/reference and imports the reference's OWN, UNMODIFIED Python
sources; every `tf.*` call they make lands here and is executed eagerly on torch CPU tensors.  That
gives (a) forward values produced by the reference's own wiring and (b) gradients from torch autograd
through that same wiring -- both are committed as fixtures under tests/golden/.

What is restated here (and therefore NOT pinned by the reference itself) is the semantics of each TF
library op [TF1.4]:  scatter_nd accumulates duplicates; clip_by_value == maximum(minimum(x,hi),lo)
and so passes gradient on the closed interval; floor/equal/greater/cast-of-bool carry no gradient;
reduce_max/min route gradient to the arg-max/min; resize_images(AREA) on an integer factor is a box
mean.  Nothing under layered-scene-inference_b200/ may import this module.
"""
import builtins
import contextlib
import math
import types

import numpy as np
import torch

_FLOAT = torch.float32  # switched to float64 by gen_golden.py for finite-difference checks


def _set_float(dtype):
    global _FLOAT
    _FLOAT = dtype


float32 = 'float32'
int32 = 'int32'
int64 = 'int64'
uint8 = 'uint8'
bool = 'bool'  # noqa: A001  (mirrors tf.bool)


def _dt(d):
    if isinstance(d, torch.dtype):
        return d
    d = str(d)
    if 'float' in d:
        return _FLOAT
    if d == 'int32':
        return torch.int32
    if d == 'int64':
        return torch.int64
    if d == 'uint8':
        return torch.uint8
    if d == 'bool':
        return torch.bool
    raise TypeError(d)


class TensorShape(list):
    def as_list(self):
        return list(self)

    @property
    def ndims(self):
        return len(self)


class Tensor(object):
    """Immutable-looking wrapper: every operator returns a NEW Tensor (TF semantics), so the
    reference's `coords -= 0.5` never mutates the caller's data."""
    __array_priority__ = 1000

    def __init__(self, t):
        self.t = t

    def get_shape(self):
        return TensorShape(int(s) for s in self.t.shape)

    @property
    def shape(self):
        return self.get_shape()

    @property
    def dtype(self):
        return self.t.dtype

    def __len__(self):
        return self.t.shape[0]

    def __iter__(self):
        for i in builtins.range(self.t.shape[0]):
            yield Tensor(self.t[i])

    def __getitem__(self, idx):
        return Tensor(self.t[idx])

    def _bin(self, other, fn, rev=False):
        o = _raw(other, like=self.t)
        return Tensor(fn(o, self.t) if rev else fn(self.t, o))

    def __add__(self, o): return self._bin(o, torch.add)
    def __radd__(self, o): return self._bin(o, torch.add, True)
    def __sub__(self, o): return self._bin(o, torch.sub)
    def __rsub__(self, o): return self._bin(o, torch.sub, True)
    def __mul__(self, o): return self._bin(o, torch.mul)
    def __rmul__(self, o): return self._bin(o, torch.mul, True)
    def __truediv__(self, o): return self._bin(o, torch.div)
    def __rtruediv__(self, o): return self._bin(o, torch.div, True)
    __div__ = __truediv__
    __rdiv__ = __rtruediv__
    def __neg__(self): return Tensor(-self.t)
    def __gt__(self, o): return self._bin(o, torch.gt)
    def __lt__(self, o): return self._bin(o, torch.lt)

    def numpy(self):
        return self.t.detach().cpu().numpy()


def _raw(x, like=None):
    """Python scalar / list / ndarray / Tensor -> torch tensor (float data -> _FLOAT)."""
    if isinstance(x, Tensor):
        return x.t
    if isinstance(x, torch.Tensor):
        return x
    if isinstance(x, (int, float)) and like is not None and like.dtype.is_floating_point:
        return torch.tensor(x, dtype=like.dtype)
    a = np.asarray(x)
    if a.dtype.kind == 'f':
        return torch.as_tensor(a.astype(np.float64)).to(_FLOAT)
    if a.dtype.kind == 'b':
        return torch.as_tensor(a)
    if like is not None and like.dtype.is_floating_point:
        return torch.as_tensor(a).to(like.dtype)
    return torch.as_tensor(a)


def _ishape(shape):
    out = []
    for s in shape:
        if isinstance(s, float):
            assert s == math.floor(s), 'non-integral shape entry %r' % s  # TF accepts 128.0
        out.append(int(s))
    return out


def convert_to_tensor(x, dtype=None):
    t = _raw(x)
    if dtype is not None:
        t = t.to(_dt(dtype))
    return Tensor(t)


def constant(value, dtype=None, shape=None):
    t = _raw(value)
    if dtype is not None:
        t = t.to(_dt(dtype))
    if shape is not None:
        t = t.reshape(_ishape(shape))
    return Tensor(t)


@contextlib.contextmanager
def name_scope(*_a, **_k):
    yield


@contextlib.contextmanager
def control_dependencies(deps):
    yield


def assert_equal(a, b):
    assert torch.equal(_raw(a).to(torch.int64), _raw(b).to(torch.int64)), (a, b)
    return None


def Print(x, *_a, **_k):  # noqa: N802
    return x


def cast(x, dtype):
    return Tensor(_raw(x).to(_dt(dtype)))


def reshape(x, shape):
    if isinstance(shape, Tensor):
        shape = shape.t.tolist()
    return Tensor(_raw(x).reshape(_ishape([s.t.item() if isinstance(s, Tensor) else s for s in shape])))


def shape(x, out_type='int32'):
    return Tensor(torch.tensor(list(_raw(x).shape), dtype=_dt(out_type)))


def ones(shape, dtype='float32'):
    if isinstance(shape, Tensor):
        shape = shape.t.tolist()
    return Tensor(torch.ones(_ishape(shape), dtype=_dt(dtype)))


def zeros(shape, dtype='float32'):
    if isinstance(shape, Tensor):
        shape = shape.t.tolist()
    return Tensor(torch.zeros(_ishape(shape), dtype=_dt(dtype)))


def range(*a):  # noqa: A001
    return Tensor(torch.arange(*[int(v) for v in a], dtype=torch.int32))


def concat(values, axis):
    return Tensor(torch.cat([_raw(v) for v in values], dim=axis))


def stack(values, axis=0):
    vals = [_raw(v) for v in values]
    if all(v.dim() == 0 for v in vals) and not any(v.dtype.is_floating_point for v in vals):
        return Tensor(torch.stack([v.to(torch.int64) for v in vals]))
    return Tensor(torch.stack(vals, dim=axis))


def split(value, num_or_size_splits, axis=0):
    t = _raw(value)
    if isinstance(num_or_size_splits, int):
        sizes = t.shape[axis] // num_or_size_splits
    else:
        sizes = list(num_or_size_splits)
    return [Tensor(p) for p in torch.split(t, sizes, dim=axis)]


def expand_dims(x, axis):
    return Tensor(_raw(x).unsqueeze(axis))


def tile(x, multiples):
    return Tensor(_raw(x).repeat(*_ishape(multiples)))


def transpose(x, perm=None):
    t = _raw(x)
    if perm is None:
        perm = list(builtins.range(t.dim()))[::-1]
    return Tensor(t.permute(*perm))


def matmul(a, b, name=None):
    return Tensor(torch.matmul(_raw(a), _raw(b)))


def matrix_inverse(x, name=None):
    return Tensor(torch.linalg.inv(_raw(x)))


def add(a, b):
    return Tensor(_raw(a) + _raw(b))


def add_n(xs):
    out = _raw(xs[0])
    for v in xs[1:]:
        out = out + _raw(v)
    return Tensor(out)


def divide(a, b, name=None):
    return Tensor(_raw(a) / _raw(b))


def floor(x):
    return Tensor(torch.floor(_raw(x)))


def exp(x, name=None):
    return Tensor(torch.exp(_raw(x)))


def log(x):
    return Tensor(torch.log(_raw(x)))


def abs(x):  # noqa: A001
    return Tensor(torch.abs(_raw(x)))


def square(x):
    t = _raw(x)
    return Tensor(t * t)


def clip_by_value(x, lo, hi):
    t = _raw(x)
    if not t.dtype.is_floating_point:
        t = t.to(_FLOAT)
    lo_t, hi_t = _raw(lo, like=t), _raw(hi, like=t)
    # [TF1.4] clip_by_value = maximum(minimum(t, hi), lo)
    return Tensor(torch.maximum(torch.minimum(t, hi_t), lo_t))


def pad(tensor, paddings, mode='CONSTANT', constant_values=0):
    """[TF1.4] tf.pad, CONSTANT mode: paddings [rank, 2] = (before, after) per dimension."""
    t = _raw(tensor)
    pads = _raw(paddings).to(torch.int64).tolist()
    assert mode == 'CONSTANT' and len(pads) == t.dim()
    flat = []
    for before, after in reversed(pads):          # torch pads the last dimension first
        flat += [int(before), int(after)]
    return Tensor(torch.nn.functional.pad(t, flat, value=constant_values))


def equal(a, b):
    ta = _raw(a)
    return Tensor(torch.eq(ta, _raw(b, like=ta)))


def greater(a, b):
    ta = _raw(a)
    if not ta.dtype.is_floating_point and isinstance(b, float):
        ta = ta.to(_FLOAT)
    return Tensor(torch.gt(ta, _raw(b, like=ta)))


def less(a, b):
    ta = _raw(a)
    return Tensor(torch.lt(ta, _raw(b, like=ta)))


def _reduce(fn, x, axis, keep_dims):
    t = _raw(x)
    if axis is None:
        return Tensor(fn(t))
    out = fn(t, dim=axis, keepdim=keep_dims)
    if isinstance(out, tuple):
        out = out[0]
    return Tensor(out)


def reduce_sum(x, axis=None, keep_dims=False):
    return _reduce(torch.sum, x, axis, keep_dims)


def reduce_mean(x, axis=None, keep_dims=False):
    return _reduce(torch.mean, x, axis, keep_dims)


def reduce_max(x, axis=None, keep_dims=False):
    t = _raw(x)
    if axis is None:
        return Tensor(t.max())
    return Tensor(torch.amax(t, dim=axis, keepdim=keep_dims))


def reduce_min(x, axis=None, keep_dims=False):
    t = _raw(x)
    if axis is None:
        return Tensor(t.min())
    return Tensor(torch.amin(t, dim=axis, keepdim=keep_dims))


def argmin(x, axis=None):      # [TF1.4] axis=None means axis 0 (math_ops.argmin), not a flattened arg-min
    return Tensor(torch.argmin(_raw(x), dim=0 if axis is None else axis))


def argmax(x, axis=None):      # [TF1.4] axis=None means axis 0 (layers.py:72 relies on it)
    return Tensor(torch.argmax(_raw(x), dim=0 if axis is None else axis))


def cumsum(x, axis=0):
    return Tensor(torch.cumsum(_raw(x), dim=axis))


def one_hot(indices, depth, axis=-1):
    oh = torch.nn.functional.one_hot(_raw(indices).to(torch.int64), int(depth)).to(_FLOAT)
    if axis not in (-1, oh.dim() - 1):
        oh = oh.movedim(-1, axis)
    return Tensor(oh)


def stop_gradient(x):
    return Tensor(_raw(x).detach())


def gather(params, indices):
    p = _raw(params)
    idx = _raw(indices).to(torch.int64)
    return Tensor(p[idx.reshape(-1)].reshape(list(idx.shape) + list(p.shape[1:])))


def scatter_nd(indices, updates, shape):
    """[TF1.4] zeros(shape) with updates summed at indices (duplicates accumulate)."""
    idx = _raw(indices).to(torch.int64)
    upd = _raw(updates)
    shp = _ishape(_raw(shape).tolist())
    assert idx.dim() == 2 and idx.shape[1] == 1 and len(shp) == 1, 'shim covers the 1-D use only'
    out = torch.zeros(shp, dtype=upd.dtype)
    return Tensor(out.index_add(0, idx[:, 0], upd))


def _resize_area(images, size):
    t = _raw(images)
    b, h, w, c = t.shape
    ho, wo = _ishape(size)
    assert h % ho == 0 and w % wo == 0, 'shim covers integer-factor AREA resize only'
    fh, fw = h // ho, w // wo
    return Tensor(t.reshape(b, ho, fh, wo, fw, c).mean(dim=(2, 4)))


class _Obj(types.SimpleNamespace):
    pass


nn = _Obj(relu=lambda x: Tensor(torch.relu(_raw(x))),
          sigmoid=lambda x: Tensor(torch.sigmoid(_raw(x))))
image = _Obj(resize_images=lambda images, size, method=None: _resize_area(images, size),
             ResizeMethod=_Obj(AREA='area'))
summary = _Obj(scalar=lambda *a, **k: None, image=lambda *a, **k: None,
               histogram=lambda *a, **k: None)
train = _Obj()


# ---------------------------------------------------------------------------------------------------
# variable scopes + a flat variable store (used by the slim stand-in for lsi/nnutils/nets.py)
# ---------------------------------------------------------------------------------------------------
_VARS = {}
_SCOPES = []


class _Scope(object):
    def __init__(self, name):
        self.name = name
        self.original_name_scope = name + '/'


@contextlib.contextmanager
def variable_scope(name, reuse=None):
    _SCOPES.append(name if name is not None else 'default')
    try:
        yield _Scope('/'.join(_SCOPES))
    finally:
        _SCOPES.pop()


def current_scope():
    return '/'.join(_SCOPES)


def get_variable(name, shape, init):
    full = current_scope() + '/' + name
    if full not in _VARS:
        _VARS[full] = init(list(shape))
    v = _VARS[full]
    assert list(v.shape) == list(shape), (full, list(v.shape), list(shape))
    return v
