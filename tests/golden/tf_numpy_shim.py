"""A NumPy-backed stand-in for the ``tensorflow`` module, just large enough to EXECUTE the
reference's own, unmodified ``src/postprocess.py``, ``src/anchors.py``, ``src/utils_box.py``
and ``src/utils_extra.py`` in a container where TensorFlow 2.10 cannot be installed.

Only ``tests/golden/make_golden.py`` uses it (to produce the committed fixtures); nothing in
the product or in the GPU tests imports it.  The control flow, indexing, operation order and
dtype casts that end up in the fixtures are therefore the reference's own source text; the
primitive ops (exp, top_k tie order, reduce_std, gather ...) are this file's restatement of
TF semantics and are listed as such in DESIGN.md:

  * no implicit dtype promotion (NumPy >= 2 weak Python scalars behave like TF here);
  * ``reduce_mean`` / ``reduce_std`` over axis 0: sequential fp32 sum in sample order / count,
    two-pass population std;
  * ``math.top_k``: value descending, index ascending (sorted=False order is unspecified in
    TF; the canonical order is used);
  * ``sigmoid(x)`` = fp32(1 / (1 + exp(-fp64(x))));
  * graph-mode shape semantics needed by ``per_class_nms``: the un-padded outputs of
    ``NonMaxSuppressionV5`` have an unknown leading dimension (``shape[0] is None``), and
    ``gather`` follows the TF-GPU rule for out-of-range rows (zeros);
  * ``raw_ops.NonMaxSuppressionV5`` -> ``oracle/nms_v5.c`` (restated TF kernel, unpinned).
"""
import sys
import types
from unittest import mock

import numpy as np


class _Shape(tuple):
    def as_list(self):
        return list(self)


class Tensor(np.ndarray):
    """ndarray with a TF-like ``shape`` (``as_list``; optional unknown leading dim)."""

    _dyn0 = False

    def __array_finalize__(self, obj):
        self._dyn0 = False

    @property
    def shape(self):
        s = np.ndarray.shape.__get__(self)
        if self._dyn0 and len(s) > 0:
            return _Shape((None,) + tuple(s[1:]))
        return _Shape(s)

    def numpy(self):
        return np.asarray(self)


def _t(x, dyn0=False):
    out = np.asarray(x).view(Tensor)
    out._dyn0 = dyn0
    return out


def _a(x):
    return np.asarray(x)


def _is_dyn(x):
    return isinstance(x, Tensor) and x._dyn0


float32, float64, int32, int64 = np.float32, np.float64, np.int32, np.int64


def cast(x, dtype=None, **_):
    return _t(_a(x).astype(dtype))


def convert_to_tensor(x, dtype=None, **_):
    return _t(np.asarray(x, dtype=dtype))


def identity(x, **_):
    return _t(_a(x).copy())


def shape(x):
    return _t(np.asarray(np.ndarray.shape.__get__(np.asarray(x)), np.int32))


def reshape(x, shp):
    return _t(_a(x).reshape([int(s) for s in shp]))


def transpose(x, perm=None):
    return _t(np.transpose(_a(x), perm))


def concat(values, axis, **_):
    return _t(np.concatenate([_a(v) for v in values], axis=axis))


def stack(values, axis=0, **_):
    return _t(np.stack([_a(v) for v in values], axis=axis))


def unstack(x, num=None, axis=0):
    x = _a(x)
    return [_t(np.take(x, i, axis=axis)) for i in range(x.shape[axis])]


def expand_dims(x, axis):
    return _t(np.expand_dims(_a(x), axis))


def tile(x, multiples):
    return _t(np.tile(_a(x), [int(m) for m in multiples]))


def range_(n):
    return _t(np.arange(int(n), dtype=np.int32))


def zeros_like(x):
    return _t(np.zeros_like(_a(x)))


def ones_like(x):
    return _t(np.ones_like(_a(x)))


def zeros(shp, dtype=np.float32):
    return _t(np.zeros(shp, dtype))


def equal(a, b):
    return _t(_a(a) == b)


def not_equal(a, b):
    return _t(_a(a) != b)


def where(cond):
    return _t(np.argwhere(_a(cond)).astype(np.int64))


def minimum(a, b):
    return _t(np.minimum(_a(a), _a(b)))


def abs_(x):
    return _t(np.abs(_a(x)))


def clip_by_value(x, lo, hi):
    x = _a(x)
    return _t(np.minimum(np.maximum(x, np.asarray(lo, x.dtype)), np.asarray(hi, x.dtype)))


def pad(x, paddings):
    return _t(np.pad(_a(x), paddings))


def slice_(x, begin, size):
    x = _a(x)
    idx = tuple(slice(b, b + s) for b, s in zip(begin, size))
    return _t(x[idx])


def _seq_sum0(x):
    acc = x[0].copy()
    for i in range(1, x.shape[0]):
        acc = acc + x[i]
    return acc


def reduce_mean(x, axis=None):
    x = _a(x)
    assert axis == 0, "shim implements the MC-axis reduction only"
    return _t(_seq_sum0(x) / x.dtype.type(x.shape[0]))


def reduce_std(x, axis=None):
    x = _a(x)
    assert axis == 0
    m = _seq_sum0(x) / x.dtype.type(x.shape[0])
    d = x - m[None]
    return _t(np.sqrt(_seq_sum0(d * d) / x.dtype.type(x.shape[0])))


def reduce_max(x, axis=None):
    return _t(_a(x).max(axis=axis))


def reduce_sum(x, axis=None):
    return _t(_a(x).sum(axis=axis))


def argmax(x, axis=None, output_type=np.int64):
    return _t(np.argmax(_a(x), axis=axis).astype(output_type))


def sigmoid(x):
    x = _a(x)
    return _t((1.0 / (1.0 + np.exp(-x.astype(np.float64)))).astype(x.dtype))


def top_k(values, k, sorted=True):  # noqa: A002 - TF keyword
    values = _a(values)
    order = np.argsort(-values, axis=-1, kind="stable")[..., :k].astype(np.int32)
    return _t(np.take_along_axis(values, order, axis=-1)), _t(order)


def gather(params, indices, **_):
    params, idx = _a(params), _a(indices)
    flat = idx.reshape(-1)
    out = np.zeros((flat.shape[0],) + params.shape[1:], params.dtype)
    ok = (flat >= 0) & (flat < params.shape[0])
    out[ok] = params[flat[ok]]
    return _t(out.reshape(idx.shape + params.shape[1:]), dyn0=_is_dyn(indices))


def gather_nd(params, indices, batch_dims=0):
    params, idx = _a(params), _a(indices)
    if batch_dims == 0:
        depth = idx.shape[-1]
        lead = idx.reshape(-1, depth)
        out = np.zeros((lead.shape[0],) + params.shape[depth:], params.dtype)
        ok = np.ones(lead.shape[0], bool)
        for d in range(depth):
            ok &= (lead[:, d] >= 0) & (lead[:, d] < params.shape[d])
        out[ok] = params[tuple(lead[ok].T)]
        return _t(out.reshape(idx.shape[:-1] + params.shape[depth:]))
    assert batch_dims == 1
    return _t(np.stack([params[b][tuple(np.moveaxis(idx[b], -1, 0))] for b in range(params.shape[0])]))


def _nms_v5(boxes, scores, max_output_size, iou_threshold, score_threshold, soft_nms_sigma,
            pad_to_max_output_size):
    from oracle import nms_ref

    idx, sc, valid = nms_ref.non_max_suppression_v5(
        _a(boxes), _a(scores), max_output_size, iou_threshold, score_threshold, soft_nms_sigma,
        pad_to_max_output_size, variant=VARIANT[0])
    dyn = not pad_to_max_output_size
    return _t(idx, dyn0=dyn), _t(sc, dyn0=dyn), _t(np.int32(valid))


VARIANT = ["new"]


def _moments(x, axes):
    x = _a(x)
    m = x.mean(axis=tuple(axes))
    return _t(m), _t(np.mean(np.square(x - m), axis=tuple(axes)))


class _Fallback(types.ModuleType):
    """module whose unknown attributes are MagicMocks (lets import-time class bodies run)."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        m = mock.MagicMock(name=self.__name__ + "." + name)
        setattr(self, name, m)
        return m


def install():
    """Register the stand-in modules in sys.modules; returns the fake ``tensorflow``."""
    tf = _Fallback("tensorflow")
    tf.Tensor = Tensor
    for name, fn in dict(
        float32=float32, float64=float64, int32=int32, int64=int64, cast=cast,
        convert_to_tensor=convert_to_tensor, identity=identity, shape=shape, reshape=reshape,
        transpose=transpose, concat=concat, stack=stack, unstack=unstack, expand_dims=expand_dims,
        tile=tile, range=range_, zeros_like=zeros_like, ones_like=ones_like, zeros=zeros,
        equal=equal, not_equal=not_equal, where=where, minimum=minimum, abs=abs_,
        clip_by_value=clip_by_value, pad=pad, slice=slice_, reduce_mean=reduce_mean,
        reduce_max=reduce_max, reduce_sum=reduce_sum, gather=gather, gather_nd=gather_nd,
    ).items():
        setattr(tf, name, fn)
    math = _Fallback("tensorflow.math")
    math.top_k = top_k
    math.argmax = argmax
    math.sigmoid = sigmoid
    math.reduce_std = reduce_std
    math.exp = lambda x: _t(np.exp(_a(x)))
    math.sqrt = lambda x: _t(np.sqrt(_a(x)))
    math.square = lambda x: _t(np.square(_a(x)))
    tf.math = math
    nn = _Fallback("tensorflow.nn")
    nn.moments = _moments
    tf.nn = nn
    raw = _Fallback("tensorflow.raw_ops")
    raw.NonMaxSuppressionV5 = _nms_v5
    tf.raw_ops = raw
    compat = _Fallback("tensorflow.compat")
    compat.v1 = tf
    compat.v2 = tf
    tf.compat = compat
    mods = {
        "tensorflow": tf, "tensorflow.math": math, "tensorflow.compat": compat,
        "tensorflow.compat.v1": tf, "tensorflow.compat.v2": tf,
    }
    for name in (
        "tensorflow.python", "tensorflow.python.tpu", "tensorflow_probability",
        "uncertainty_toolbox", "uncertainty_toolbox.viz", "matplotlib", "matplotlib.pyplot",
        "object_detection_efficientdet", "tensorflow_addons", "cv2", "PIL",
    ):
        mods[name] = _Fallback(name)
    for name, m in mods.items():
        sys.modules.setdefault(name, m)
    return tf
