"""Mirror of the inference part of the reference's ``src/utils_class.py``: ``stable_softmax`` (36-41) and
``CalibrateClass`` (44-272) - the classification calibrators applied per detection in the ``InferImages`` /
``Validate`` loops (``src/infer_model.py:694-730``): temperature scaling (one temperature, or one per class) and
isotonic regression on the softmax probabilities (one table, or one per class), each giving calibrated
probabilities and their entropy.  The arithmetic runs on the device (``udal_calibrate_class``, csrc/autolabel.cu).

Differences from the reference, on purpose:
  * calibrators are handed over as data (``{"ts_all": T, "ts_percls": [T_c], "iso_all": table, "iso_percls":
    [table_c]}``, tables = ``autolabel.IsotonicTable`` / fitted sklearn models) instead of being unpickled from
    ``results/calibration/<model>/classification/`` - the file layout is the reference's control plane;
  * the branch with the MC class uncertainty draws ten tfp samples per logit (stochastic, 127-138): it is not offered
    (``uncert`` must be None);
  * ``strict_reference=True`` (default) keeps the reference's selection quirk: the keys tested at 208-233 lack the
    underscore (``"classification" + method``), so ``select_entropy`` is ALWAYS the empty array and the caller keeps the
    uncalibrated entropy; ``strict_reference=False`` returns the selected method's entropy.
"""
import numpy as np

from . import _lib, autolabel, device
from . import postprocess as _post

AVAILABLE_CALIB = ["ts_all", "ts_percls", "iso_all", "iso_percls"]


def stable_softmax(logits):
    """utils_class.py:36-41"""
    return autolabel.stable_softmax(logits)


def _tables(models):
    out = []
    for m in models:
        out.append(m if isinstance(m, autolabel.IsotonicTable) else autolabel.IsotonicTable.from_sklearn(m))
    return out


class CalibrateClass:
    def __init__(self, logits, calibrators, calib_method="ts_all", uncert=None, y_true=None, strict_reference=True):
        if uncert is not None:
            raise NotImplementedError("CalibrateClass with the MC class uncertainty draws tfp samples (utils_class.py:127-138): "
                                      "not offered on the device")
        self.logits = logits
        self.calibrators = dict(calibrators)
        self.calib_method = calib_method
        self.y_true = y_true
        self.strict_reference = strict_reference
        self.available_calib = list(AVAILABLE_CALIB)

    def _perform_class_calib(self, calib_method):
        """-> (entropy [M], probab [M,C]); NumPy for host logits, device arrays for device logits"""
        if calib_method not in _lib.CLASSCAL:
            raise ValueError("Unknown calibration method")
        eng = _post._any_engine()
        host = not _post._is_dev(self.logits)
        ctx = self.logits.ctx if isinstance(self.logits, device.DeviceArray) else eng.ctx
        lg, _ = device.as_device(ctx, self.logits if not host else np.ascontiguousarray(self.logits, np.float32), np.float32)
        rows, c = lg.shape[0], lg.shape[-1]
        cal = self.calibrators[calib_method]
        temps = tx = ty = off = None
        if calib_method.startswith("ts"):
            t = np.reshape(np.asarray(cal, np.float32), -1)
            if t.size != (c if calib_method == "ts_percls" else 1):
                raise ValueError("%s needs %d temperature(s)" % (calib_method, c if calib_method == "ts_percls" else 1))
            temps = ctx.to_device(t)
        else:
            tabs = _tables(cal if calib_method == "iso_percls" else [cal])
            if len(tabs) != (c if calib_method == "iso_percls" else 1):
                raise ValueError("%s needs one table per class" % calib_method)
            tx = ctx.to_device(np.concatenate([t.x for t in tabs]))
            ty = ctx.to_device(np.concatenate([t.y for t in tabs]))
            off = ctx.to_device(np.concatenate([[0], np.cumsum([t.x.size for t in tabs])]).astype(np.int32))
        prob = ctx.empty((rows, c))
        ent = ctx.empty((rows,))
        _lib.check(eng.lib.udal_calibrate_class(ctx.handle, lg.ptr, rows, c, _lib.CLASSCAL[calib_method],
                                                temps.ptr if temps is not None else None, tx.ptr if tx is not None else None,
                                                ty.ptr if ty is not None else None, off.ptr if off is not None else None,
                                                prob.ptr, ent.ptr))
        if host:
            return ent.numpy(), prob.numpy()
        return ent, prob

    def calibrate_class(self):
        """utils_class.py:189-272 without the uncertainty branch: (select_entropy, ts_all probab, ts_all entropy, ts_percls
        probab, entropy, iso_all probab, entropy, iso_percls probab, entropy); a missing calibrator gives empty arrays."""
        res = []
        for m in self.available_calib:
            if m in self.calibrators:
                ent, prob = self._perform_class_calib(m)
                res.append((ent, prob))
            else:
                res.append((np.array([]), np.array([])))
        select_entropy = np.array([])
        if not self.strict_reference and self.calib_method in self.calibrators and self.calib_method in self.available_calib:
            select_entropy = res[self.available_calib.index(self.calib_method)][0]
        out = [select_entropy]
        for ent, prob in res:
            out += [prob, ent]
        return tuple(out)
