"""Generates tests/golden/classcalib.npz (BUILD container only: needs /root/reference and scikit-learn).

Executes the reference's own ``utils_class.CalibrateClass._perform_class_calib`` (unmodified source, imported through the
NumPy ``tensorflow`` stand-in) for the four classification calibrators - ts_all, ts_percls, iso_all, iso_percls - without
the MC class uncertainty (that branch draws tfp samples), with sklearn IsotonicRegression calibrators fitted on synthetic
data, and stores inputs, calibrator knots and outputs (entropy, probabilities).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import tf_numpy_shim  # noqa: E402

tf_numpy_shim.install()
sys.path.insert(0, "/root/reference/src")
import utils_class as ref  # noqa: E402
from sklearn.isotonic import IsotonicRegression  # noqa: E402

rng = np.random.default_rng(41)
M, C = 100, 7
logits = rng.normal(-3, 2.5, (M, C)).astype(np.float32)
logits[5] = 0.0            # a uniform row
logits[6, 2] = 40.0        # a saturated row


def fit(seed):
    r = np.random.default_rng(seed)
    x = r.uniform(0, 1, 400)
    y = np.clip(x ** 1.5 + r.normal(0, 0.05, 400), 0, 1)
    return IsotonicRegression(increasing=True, out_of_bounds="clip").fit(x, y)


iso_all = fit(500)
iso_pc = [fit(510 + i) for i in range(C)]
temps_pc = rng.uniform(0.6, 2.5, C).astype(np.float32)
cal = object.__new__(ref.CalibrateClass)
cal.logits = logits
cal.uncert = None
cal.y_true = None
cal.calibrators = {"classification_ts_all": np.float32(1.7), "classification_ts_percls": temps_pc,
                   "classification_iso_all": iso_all, "classification_iso_percls": iso_pc}
out = dict(logits=logits, temp_all=np.float32(1.7), temps_pc=temps_pc)
models = [iso_all] + iso_pc
out["tx"] = np.concatenate([m.X_thresholds_ for m in models]).astype(np.float64)
out["ty"] = np.concatenate([m.y_thresholds_ for m in models]).astype(np.float64)
out["off"] = np.concatenate([[0], np.cumsum([m.X_thresholds_.size for m in models])]).astype(np.int32)
for method in ("ts_all", "ts_percls", "iso_all", "iso_percls"):
    ent, prob = cal._perform_class_calib(method)
    out["entropy_" + method] = np.asarray(ent)
    out["probab_" + method] = np.asarray(prob)
np.savez_compressed(os.path.join(HERE, "classcalib.npz"), **out)
print("wrote classcalib.npz", {k: (v.shape, str(v.dtype)) for k, v in out.items() if k.startswith(("entropy", "probab"))})
