"""Mirror of ``src/utils_box.py`` ``decode_uncert`` (utils_box.py:105-276) on the device.

The stand-alone entry point wraps the same fused kernel the post-processing uses
(csrc/decode_moments.cu) with a one-level, one-anchor-per-location geometry, so the numbers are
the ones the pipeline produces.
"""
import ctypes
import os

import numpy as np

from . import _lib
from . import engine as _engine

_GENERIC = dict(
    image_size=(1, 1), min_level=0, max_level=0, num_scales=1, aspect_ratios=[1.0], anchor_scale=1.0,
    num_classes=1, mc_dropout=False, mc_dropoutrate=0.0, mc_classheadrate=0.0, mc_boxheadrate=0.0,
    mc_dropoutsamp=1, fpn_num_filters=4, box_class_repeats=1,
    nms_configs=dict(method="hard", iou_thresh=None, score_thresh=None, sigma=None,
                     max_nms_inputs=0, max_output_size=1),
)


_engines = {}  # (n, loss attenuation, method, device) -> cached one-level engine (context + anchor buffer)


def _generic_engine(n, la, method, device_id):
    """One-level, one-anchor-per-location engine for stand-alone calls on ``n`` rows; cached per shape so that repeated
    calls (the reference calls decode_uncert once per MC sample) reuse one CUDA context instead of creating one each."""
    key = (n, la, method, device_id)
    eng = _engines.get(key)
    if eng is None:
        if len(_engines) >= 8:  # bounded: drop the oldest shape
            _engines.pop(next(iter(_engines))).ctx.close()
        params = dict(_GENERIC, loss_attenuation=la, uncert_adjust_method=method if la else "l-norm")
        eng = _engines[key] = _engine.Engine(params, device_id, level_hw=[(1, n)], anchors=np.zeros((n, 4), np.float32))
    return eng


def _decode(pred_boxes, box_uncert, anchor_boxes, method, device_id=0):
    pred = np.asarray(pred_boxes, np.float32)
    lead = pred.shape[:-1]
    n = int(np.prod(lead, dtype=np.int64)) if lead else 1
    if pred.shape[-1] != 4:
        raise ValueError("pred_boxes must have a trailing dimension of 4")
    anc = np.ascontiguousarray(np.broadcast_to(np.asarray(anchor_boxes, np.float32), pred.shape).reshape(n, 4))
    la = box_uncert is not None
    eng = _generic_engine(n, la, method, device_id)
    _lib.check(eng.lib.udal_set_anchors(eng.ctx.handle, anc.ctypes.data, n))
    t = pred.reshape(1, 1, n, 4)
    box = np.concatenate([t, np.asarray(box_uncert, np.float32).reshape(1, 1, n, 4)], -1) if la else t
    out = eng.decode_moments([eng.ctx.zeros((1, 1, n, 1))], [eng.ctx.to_device(box)], 1, want=("boxes", "albox"))
    coords = out["boxes"].numpy().reshape(pred.shape)
    if la:
        return coords, out["albox"].numpy().reshape(pred.shape)
    return coords


def philox_normals(n_samples, n, seed):
    """NumPy twin of the in-kernel standard normals of the 'sample' method (csrc/decode_sample.cu): Philox4x32-10 with
    counter s * n + i and key ``seed``; the four words give two Box-Muller pairs (z_y, z_x | z_h, z_w) from
    u = ((x >> 8) + 0.5) * 2^-24.  Returns float64 [n_samples, 4, n]."""
    g = np.arange(n_samples * n, dtype=np.uint64)
    c = [g & np.uint64(0xFFFFFFFF), g >> np.uint64(32), np.zeros_like(g), np.zeros_like(g)]
    k0, k1 = np.uint64(seed & 0xFFFFFFFF), np.uint64((seed >> 32) & 0xFFFFFFFF)
    m0, m1, mask = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0, p1 = m0 * c[0], m1 * c[2]
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & mask, p1 >> np.uint64(32), p1 & mask
        c = [hi1 ^ c[1] ^ k0, lo1, hi0 ^ c[3] ^ k1, lo0]
        k0 = (k0 + np.uint64(0x9E3779B9)) & mask
        k1 = (k1 + np.uint64(0xBB67AE85)) & mask
    u = [((w >> np.uint64(8)).astype(np.float64) + 0.5) / 16777216.0 for w in c]
    r0, a0 = np.sqrt(-2.0 * np.log(u[0])), 2.0 * np.pi * u[1]
    r1, a1 = np.sqrt(-2.0 * np.log(u[2])), 2.0 * np.pi * u[3]
    z = np.stack([r0 * np.cos(a0), r0 * np.sin(a0), r1 * np.cos(a1), r1 * np.sin(a1)], 0)  # [4, S*n]
    return np.ascontiguousarray(z.reshape(4, n_samples, n).transpose(1, 0, 2))


def _decode_sample(pred_boxes, box_uncert, anchor_boxes, n_samples, normals, seed, device_id=0):
    pred = np.asarray(pred_boxes, np.float32)
    if pred.shape[-1] != 4:
        raise ValueError("pred_boxes must have a trailing dimension of 4")
    n = pred.size // 4
    eng = _generic_engine(n, True, "l-norm", device_id)
    ctx = eng.ctx
    t = ctx.to_device(pred.reshape(n, 4))
    sg = ctx.to_device(np.ascontiguousarray(np.broadcast_to(np.asarray(box_uncert, np.float32), pred.shape)).reshape(n, 4))
    anc = ctx.to_device(np.ascontiguousarray(np.broadcast_to(np.asarray(anchor_boxes, np.float32), pred.shape)).reshape(n, 4))
    zptr, z = None, None
    if normals is not None:
        z = np.asarray(normals, np.float32)
        if z.shape[:2] != (n_samples, 4) or z.size != n_samples * 4 * n:
            raise ValueError("normals must be [n_samples, 4, ...] matching pred_boxes")
        z = ctx.to_device(z.reshape(n_samples, 4, n))
        zptr = z.ptr
    coords, stds = ctx.empty((n, 4)), ctx.empty((n, 4))
    _lib.check(eng.lib.udal_decode_sample(ctx.handle, t.ptr, sg.ptr, anc.ptr, n, int(n_samples), zptr,
                                          ctypes.c_uint64(int(seed) & ((1 << 64) - 1)), coords.ptr, stds.ptr))
    return coords.numpy().reshape(pred.shape), stds.numpy().reshape(pred.shape)


def decode_uncert(pred_boxes, box_uncert, anchor_boxes, method="l-norm", n_samples=30, device_id=0,
                  normals=None, seed=None):
    """utils_box.py:105-276.  Methods 'l-norm', 'n-flow', 'falsedec' (fp64 on the device, rounded to fp32 like the
    reference) and 'sample' (utils_box.py:162-184: moments of ``n_samples`` decoded draws).  The reference draws the
    'sample' normals from tfp (stateful, not reproducible); here they are either injected - ``normals``
    [n_samples, 4, ...] standard normal, the parity mode - or drawn in-kernel from Philox4x32-10 keyed by ``seed``
    (None: a fresh seed per call; ``philox_normals`` is the NumPy twin of that stream).
    Host arrays in, host arrays out (device tensors go through postprocess.*, which fuses the decode)."""
    if method == "sample":
        if box_uncert is None:
            raise ValueError("decode method 'sample' needs box_uncert")
        if seed is None:
            seed = int.from_bytes(os.urandom(8), "little")
        return _decode_sample(pred_boxes, box_uncert, anchor_boxes, n_samples, normals, seed, device_id)
    if method not in _lib.DECODE_METHODS:
        raise ValueError("unknown decode method {}".format(method))
    return _decode(pred_boxes, box_uncert, anchor_boxes, method, device_id)
