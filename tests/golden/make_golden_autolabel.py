"""Generates tests/golden/autolabel.npz (BUILD container only: needs /root/reference and scikit-learn).

Executes the reference's own ``utils_box.relativize_uncert`` and
``utils_box.CalibrateBoxUncert.calibrate_boxuncert`` (unmodified source, imported through the NumPy
``tensorflow`` stand-in) with sklearn IsotonicRegression calibrators fitted on synthetic data, and stores
inputs, calibrator knots and outputs.  The entropy / decision lines live inline in
``InferImages`` (src/infer_model.py:585-595, 742-764) and cannot be imported; they are executed here as
written there (same expressions) on the same synthetic detections.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import tf_numpy_shim  # noqa: E402

tf_numpy_shim.install()
sys.path.insert(0, "/root/reference/src")
import utils_box as ref_utils_box  # noqa: E402
from sklearn.isotonic import IsotonicRegression  # noqa: E402

rng = np.random.default_rng(31)
M, C = 100, 7
out = {}
# synthetic detections in the postprocess_global layout
y0 = rng.uniform(0, 300, M); x0 = rng.uniform(0, 1100, M)
boxes = np.stack([y0, x0, y0 + rng.uniform(5, 80, M), x0 + rng.uniform(5, 160, M)], 1).astype(np.float32)
albox = np.abs(rng.normal(0, 4, (M, 4))).astype(np.float32)
albox[3, 1] = np.nan                      # nan_to_num path
classes = rng.integers(1, C + 1, M).astype(np.float32)
classes[classes == 5] = 6                 # a class without detections
logits = rng.normal(-3, 2.5, (M, C)).astype(np.float32)
scores = np.sort(rng.uniform(0, 1, M))[::-1].astype(np.float32)
scores[60:] = 0
out.update(boxes=boxes, albox=albox, classes=classes, logits=logits, scores=scores)


def fit(seed, scale):
    r = np.random.default_rng(seed)
    x = np.abs(r.normal(0, scale, 300))
    y = 0.6 * x + np.abs(r.normal(0, 0.2 * scale, 300))
    return IsotonicRegression(increasing=True, out_of_bounds="clip").fit(x, y)


abs_models = [fit(100 + i, 4.0) for i in range(C * 4)]
rel_models = [fit(300 + i, 0.1) for i in range(C * 4)]
for name, models in (("abs", abs_models), ("rel", rel_models)):
    out[name + "_tx"] = np.concatenate([m.X_thresholds_ for m in models]).astype(np.float64)
    out[name + "_ty"] = np.concatenate([m.y_thresholds_ for m in models]).astype(np.float64)
    out[name + "_off"] = np.concatenate([[0], np.cumsum([m.X_thresholds_.size for m in models])]).astype(np.int32)

cal = object.__new__(ref_utils_box.CalibrateBoxUncert)
cal.model_params = {"num_classes": C, "calib_method_box": "iso_perclscoo"}
cal.iso_calib_all = abs_models[0]
cal.temp_regres_all = 1.7
cal.ymin_calib, cal.xmin_calib, cal.ymax_calib, cal.xmax_calib = abs_models[0:4]
cal.ymin_calib_temp, cal.xmin_calib_temp, cal.ymax_calib_temp, cal.xmax_calib_temp = 1.3, 1.5, 0.9, 2.1
cal.iso_perclscoo = abs_models
cal.iso_perclscoo_rel = rel_models
sel, iso_all, ts_all, ts_percoo, iso_percoo, iso_perclscoo, rel_iso = cal.calibrate_boxuncert(albox, classes, boxes)
out.update(iso_all=iso_all, ts_all=ts_all, ts_percoo=ts_percoo, iso_percoo=iso_percoo, iso_perclscoo=iso_perclscoo,
           rel_iso_perclscoo=rel_iso, temps_percoo=np.float32([1.3, 1.5, 0.9, 2.1]), temp_all=np.float32(1.7))
assert np.array_equal(sel, iso_perclscoo)
out["rel_plain"] = ref_utils_box.relativize_uncert(boxes, albox)
out["rel_first_row"] = ref_utils_box.relativize_uncert(boxes, iso_perclscoo[0])   # infer_model.py:688-691 after calibration

# infer_model.py:585-595 and 742-764, expressions as written there
probab_logits = []
for x in logits:
    probab_logits.append(np.exp(x - max(x)) / np.sum(np.exp(x - max(x))))   # utils_class.stable_softmax
probab_logits = [np.asarray(probab_logits)]
entropy = -np.sum(probab_logits[0] * np.nan_to_num(np.log2(np.maximum(probab_logits[0], 10**-7))), axis=1)
out["entropy"] = entropy
opt_params, opt_thrs, min_score = [0.5, 0.5], [0.5, 2.5], 0.4
for tag, rel in (("plain", out["rel_plain"]), ("strict", out["rel_first_row"])):
    thr_uncerts = [entropy, np.mean(rel, axis=-1)]
    opt_uncert = sum(opt_param * uncert for opt_param, uncert in zip(opt_params, thr_uncerts))
    out["opt_" + tag] = opt_uncert
    out["decision_" + tag] = np.all(opt_uncert[scores > min_score] < np.mean(opt_thrs))
out.update(opt_params=np.float64(opt_params), opt_thrs=np.float64(opt_thrs), min_score=np.float64(min_score))
path = os.path.join(HERE, "autolabel.npz")
np.savez_compressed(path, **out)
print("wrote", path, os.path.getsize(path), "bytes; decisions", out["decision_plain"], out["decision_strict"])
