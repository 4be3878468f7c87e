"""CPU oracle (test infrastructure only) of the calibrated-uncertainty application + auto-label threshold
pass - NumPy restatement of the per-image logic of the reference's InferImages loop:

  src/infer_model.py:585-595   probab = stable_softmax(logits); entropy = -sum(p * nan_to_num(log2(max(p, 1e-7))))
  src/utils_class.py:36-41     stable_softmax
  src/utils_box.py:404-524     CalibrateBoxUncert.calibrate_boxuncert (temperature / isotonic tables)
  src/utils_box.py:279-292     relativize_uncert
  src/infer_model.py:688-691   relativize_uncert(boxes[0], select_albox[0])  (after calibration select_albox is
                               2-D, so [0] is the FIRST detection's std - reproduced under strict_reference)
  src/infer_model.py:742-764   opt_uncert = sum(opt_param * uncert), decision over scores > min_score

Pinned by tests/golden/autolabel.npz: produced by the reference's own relativize_uncert and
calibrate_boxuncert (sklearn IsotonicRegression calibrators, executed from /root/reference/src through the
NumPy tensorflow stand-in) - see tests/golden/make_golden_autolabel.py.  The isotonic predict below is the
published inference rule of sklearn.isotonic.IsotonicRegression (scikit-learn is not vendored by the
reference: requirements pin scikit-learn; out_of_bounds="clip" -> np.clip then scipy interp1d linear).
"""
import numpy as np


def stable_softmax(logits):
    out = []
    for x in logits:
        out.append(np.exp(x - max(x)) / np.sum(np.exp(x - max(x))))
    return np.asarray(out)


def entropy_of_logits(logits, class_temp=1.0):
    p = stable_softmax(np.asarray(logits, np.float32) / np.float32(class_temp))
    return -np.sum(p * np.nan_to_num(np.log2(np.maximum(p, 10**-7))), axis=1)


def iso_predict(table_x, table_y, x):
    tx, ty = np.asarray(table_x, np.float64), np.asarray(table_y, np.float64)
    x = np.clip(np.asarray(x, np.float64), tx[0], tx[-1])
    if tx.size == 1:
        return np.full(x.shape, ty[0])
    hi = np.clip(np.searchsorted(tx, x, side="right"), 1, tx.size - 1)
    lo = hi - 1
    slope = (ty[hi] - ty[lo]) / (tx[hi] - tx[lo])
    return slope * (x - tx[lo]) + ty[lo]


def relativize_uncert(pred_boxes, box_uncert):
    pred_boxes, box_uncert = np.asarray(pred_boxes), np.asarray(box_uncert)
    width = pred_boxes[:, 3] - pred_boxes[:, 1]
    height = pred_boxes[:, 2] - pred_boxes[:, 0]
    return box_uncert / np.swapaxes([height, width, height, width], 0, 1)


def calibrate_boxuncert(method, uncert, classes, boxes, num_classes, tables=None, temps=None):
    """-> calibrated [M,4] float32 (the select_uncert of utils_box.py:496-513)."""
    if method in (None, "none"):
        return np.asarray(uncert, np.float32)  # no calibrator: the raw std, NaNs included (infer_model.py:640-686)
    uncert = np.nan_to_num(np.asarray(uncert, np.float32))
    if method == "ts_all":
        return uncert / np.float32(np.reshape(temps, -1)[0])
    if method == "ts_percoo":
        return uncert / np.asarray(temps, np.float32).reshape(1, 4)
    if method == "iso_all":
        t = tables[0]
        return iso_predict(t[0], t[1], uncert.flatten()).reshape([-1, 4])  # float64, as sklearn returns it
    if method == "iso_percoo":
        return np.stack([iso_predict(tables[j][0], tables[j][1], uncert[:, j]) for j in range(4)], 1)  # float64
    out = np.zeros_like(uncert)
    cls = np.asarray(classes).astype(int)
    if method == "iso_perclscoo":
        for ci in range(1, num_classes + 1):
            if np.any(cls == ci):
                for j in range(4):
                    t = tables[(ci - 1) * 4 + j]
                    out[:, j][cls == ci] = iso_predict(t[0], t[1], uncert[:, j][cls == ci])
        return out
    if method == "rel_iso_perclscoo":
        width = np.asarray(boxes[:, 3] - boxes[:, 1])
        height = np.asarray(boxes[:, 2] - boxes[:, 0])
        norm = np.swapaxes([height, width, height, width], 0, 1)
        rel = np.divide(uncert, norm, out=np.zeros_like(uncert), where=norm != 0, dtype=np.float16)
        for ci in range(1, num_classes + 1):
            if np.any(cls == ci):
                for j in range(4):
                    t = tables[(ci - 1) * 4 + j]
                    out[:, j][cls == ci] = iso_predict(t[0], t[1], rel[:, j][cls == ci])
        return out * norm
    raise ValueError("Unknown calibration method")


def autolabel_image(boxes, albox, scores, classes, logits, num_classes, w_entropy, w_albox, threshold, min_score,
                    method=None, tables=None, temps=None, class_temp=1.0, strict_reference=True):
    """One image: boxes [M,4], albox [M,4], scores [M], classes [M] (1-based), logits [M,C]."""
    entropy = entropy_of_logits(logits, class_temp).astype(np.float32)
    calib = calibrate_boxuncert(method, albox, classes, np.asarray(boxes, np.float32), num_classes, tables, temps)
    calibrated = method not in (None, "none")
    sel = calib[0] if (strict_reference and calibrated) else calib
    rel = relativize_uncert(boxes, sel).astype(np.float32)
    if rel.ndim == 1:
        rel = np.broadcast_to(rel, calib.shape)
    opt = np.float32(w_entropy) * entropy + np.float32(w_albox) * np.mean(rel, axis=-1, dtype=np.float32)
    decision = bool(np.all(opt[np.asarray(scores) > min_score] < threshold))
    return dict(entropy=entropy, calib_albox=calib.astype(np.float32), rel_albox=rel.astype(np.float32),
                opt_uncert=opt.astype(np.float32), auto_label=decision)


def calibrate_class(logits, method, temps=None, tables=None):
    """CalibrateClass._perform_class_calib without the MC class uncertainty (src/utils_class.py:116-187):
    -> (entropy [M], probab [M,C]).  ts_*: softmax(logits / T) in the logits' dtype; iso_*: isotonic predict (float64) of
    the softmax probabilities, renormalised."""
    logits = np.asarray(logits)
    if method == "ts_all":
        prob = stable_softmax(logits / np.float32(np.reshape(temps, -1)[0]))
    elif method == "ts_percls":
        prob = stable_softmax(logits / np.asarray(temps, np.float32))
    elif method in ("iso_all", "iso_percls"):
        sm = stable_softmax(logits)
        if method == "iso_all":
            post = iso_predict(tables[0][0], tables[0][1], sm.flatten()).reshape(sm.shape)
        else:
            post = np.stack([iso_predict(tables[i][0], tables[i][1], sm[:, i]) for i in range(logits.shape[-1])], axis=1)
        prob = post / np.stack([np.sum(post, axis=-1)] * logits.shape[-1], axis=-1)
    else:
        raise ValueError("Unknown calibration method")
    ent = -np.sum(prob * np.nan_to_num(np.log2(np.maximum(prob, 10**-7))), axis=1)
    return ent, prob
