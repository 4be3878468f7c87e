"""Calibrated-uncertainty application + auto-label threshold pass on the device.

Mirror of the per-image logic of the reference's ``InferImages`` loop, as one kernel over the
detections of ``postprocess_global`` / ``HeadSampler.detect``:

  src/infer_model.py:585-595     entropy of ``stable_softmax(logits)``            (src/utils_class.py:36-41)
  src/utils_box.py:404-524       ``CalibrateBoxUncert.calibrate_boxuncert`` on the aleatoric box std
  src/utils_box.py:279-292       ``relativize_uncert``
  src/infer_model.py:688-691, 742-764   ``opt_uncert = sum(opt_param * uncert)`` over the uncertainties named in
                                 ``thr_sel_uncert``; the image is labelled automatically when
                                 ``all(opt_uncert[scores > min_score] < mean(opt_thrs))``

Calibrators are tables, not pickles: an sklearn ``IsotonicRegression(increasing=True,
out_of_bounds="clip")`` is exported with ``IsotonicTable.from_sklearn`` (its ``X_thresholds_`` /
``y_thresholds_``); temperature scaling is a scalar.
"""
import ctypes

import numpy as np

from . import _lib, device
from . import postprocess as _post


class IsotonicTable:
    """Knots of a fitted isotonic regressor: predict(x) = interp(clip(x, x[0], x[-1]), x, y)."""

    def __init__(self, x, y):
        self.x = np.ascontiguousarray(x, np.float32).reshape(-1)
        self.y = np.ascontiguousarray(y, np.float32).reshape(-1)
        if self.x.size != self.y.size or self.x.size == 0:
            raise ValueError("isotonic table needs matching, non-empty knot arrays")
        if np.any(np.diff(self.x) < 0):
            raise ValueError("isotonic knots must be sorted")

    @classmethod
    def from_sklearn(cls, model):
        return cls(model.X_thresholds_, model.y_thresholds_)


def stable_softmax(logits):
    """utils_class.py:36-41 (host helper; the device pass computes it itself)."""
    x = np.asarray(logits)
    e = np.exp(x - x.max(axis=-1, keepdims=True))
    return e / e.sum(axis=-1, keepdims=True)


class AutoLabeler:
    """``decide(detections)`` -> dict(entropy [B,M], calib_albox [B,M,4], rel_albox [B,M,4],
    opt_uncert [B,M], auto_label [B] bool).

    params: the model_params keys of the reference that matter here - ``num_classes``,
    ``thr_sel_uncert`` (subset of ["ENT", "ALBOX"]), ``calib_method_box`` (None | ts_all | ts_percoo |
    iso_all | iso_percoo | iso_perclscoo | rel_iso_perclscoo), ``min_score``.
    opt_params / opt_thrs: as read by infer_model.py:70-165 (one weight per selected uncertainty, in
    the order ENT, ALBOX; the threshold is their mean).
    tables: list of IsotonicTable (1, 4 or num_classes*4, class-major); temps: scalar or 4 values;
    class_temp: temperature of CalibrateClass' ``ts_all`` (logits / T), 1 = uncalibrated.
    strict_reference: reproduce infer_model.py:688-691, where after calibration the relative std of
    every detection is computed from the FIRST detection's calibrated std.
    """

    def __init__(self, params, opt_params, opt_thrs, tables=None, temps=None, class_temp=1.0,
                 strict_reference=True, device_id=0):
        self.C = int(params["num_classes"])
        sel = list(params.get("thr_sel_uncert", ["ENT", "ALBOX"]))
        for s in sel:
            if s not in ("ENT", "ALBOX"):
                raise ValueError("thr_sel_uncert entry %r is not offered on the device (ENT, ALBOX)" % (s,))
        weights = list(opt_params)
        if len(weights) < len([s for s in ("ENT", "ALBOX") if s in sel]):
            raise ValueError("opt_params needs one weight per selected uncertainty")
        it = iter(weights)
        self.w_entropy = float(next(it)) if "ENT" in sel else 0.0   # zip(opt_params, [entropy, albox]) order
        self.w_albox = float(next(it)) if "ALBOX" in sel else 0.0
        self.threshold = float(np.mean(opt_thrs))
        self.min_score = float(params.get("min_score", 0.1))
        method = params.get("calib_method_box")
        if method not in _lib.CALIB_METHODS:
            raise ValueError("Unknown calibration method {}".format(method))
        self.method = _lib.CALIB_METHODS[method]
        self.temps = np.ones(4, np.float32)
        if temps is not None:
            self.temps[:] = np.broadcast_to(np.asarray(temps, np.float32).reshape(-1), (4,)) if np.size(temps) in (1, 4) else 1
        self.class_temp = float(class_temp)
        self.strict = bool(strict_reference)
        self.tables = list(tables or [])
        need = {3: 1, 4: 4, 5: 4 * self.C, 6: 4 * self.C}.get(self.method, 0)
        if need and len(self.tables) != need:
            raise ValueError("calibration method %s needs %d isotonic tables, got %d" % (method, need, len(self.tables)))
        self._dev = None
        self.device_id = device_id

    def _upload(self, ctx):
        if self._dev is None or self._dev[0] is not ctx:
            if self.tables:
                off = np.zeros(len(self.tables) + 1, np.int32)
                off[1:] = np.cumsum([t.x.size for t in self.tables])
                tx = ctx.to_device(np.concatenate([t.x for t in self.tables]))
                ty = ctx.to_device(np.concatenate([t.y for t in self.tables]))
                to = ctx.to_device(off)
            else:
                tx = ty = to = None
            self._dev = (ctx, tx, ty, to)
        return self._dev[1:]

    def decide(self, detections):
        """detections: the 5-tuple of postprocess_global - (boxes|albox|mcbox [B,M,4k], scores [B,M],
        class|mcclass [B,M,1+C] or [B,M], valid [B], logits [B,M,C]) as NumPy or device arrays."""
        boxes, scores, classes, _valid, logits = detections
        host = not _post._is_dev(boxes)
        eng = _post._any_engine()
        # device inputs: work in THEIR context (same streams as the producer - no cross-context ordering, and a timer on
        # that context sees the pass); host inputs: any cached context
        ctx = boxes.ctx if isinstance(boxes, device.DeviceArray) else eng.ctx
        bx, _ = device.as_device(ctx, boxes, np.float32)
        sc, _ = device.as_device(ctx, scores, np.float32)
        cl, _ = device.as_device(ctx, classes, np.float32)
        lg, _ = device.as_device(ctx, logits, np.float32)
        batch, m = sc.shape
        box_stride = bx.shape[-1]
        class_stride = cl.shape[-1] if cl.ndim == 3 else 1
        if lg.shape != (batch, m, self.C):
            raise ValueError("logits must be [B,M,%d], got %s" % (self.C, lg.shape))
        albox_col = 4 if box_stride >= 8 else -1   # extract_uncertainties order: boxes | albox | mcbox
        tx, ty, to = self._upload(ctx)
        prm = _lib.AutolabelParams()
        prm.calib_method_box = self.method
        prm.num_tables = len(self.tables)
        prm.table_x, prm.table_y, prm.table_off = (tx.ptr, ty.ptr, to.ptr) if tx is not None else (None, None, None)
        for j in range(4):
            prm.temps[j] = float(self.temps[j])
        prm.class_temp = self.class_temp
        prm.w_entropy, prm.w_albox = self.w_entropy, self.w_albox
        prm.threshold, prm.min_score = self.threshold, self.min_score
        prm.strict_reference = int(self.strict)
        out = dict(entropy=ctx.empty((batch, m)), calib_albox=ctx.empty((batch, m, 4)),
                   rel_albox=ctx.empty((batch, m, 4)), opt_uncert=ctx.empty((batch, m)),
                   auto_label=ctx.empty((batch,), np.int32))
        _lib.check(eng.lib.udal_autolabel(ctx.handle, bx.ptr, box_stride, albox_col, sc.ptr, cl.ptr, class_stride,
                                          lg.ptr, self.C, batch, m, ctypes.byref(prm), out["entropy"].ptr,
                                          out["calib_albox"].ptr, out["rel_albox"].ptr, out["opt_uncert"].ptr,
                                          out["auto_label"].ptr))
        if host:
            out = {k: v.numpy() for k, v in out.items()}
            out["auto_label"] = out["auto_label"].astype(bool)
        return out
