"""Generates the committed golden fixtures in tests/golden/*.npz.

Run in the BUILD container only (needs /root/reference):

    python tests/golden/make_golden.py

What executes:
  * ``src/nms_np.py``  - the reference's own module, imported as is (NumPy only).
  * ``src/postprocess.py``, ``src/anchors.py``, ``src/utils_box.py``, ``src/utils_extra.py``,
    ``src/utils.py`` - the reference's own, unmodified source, imported with the NumPy-backed
    ``tensorflow`` stand-in of ``tf_numpy_shim.py`` (TensorFlow 2.10 is not installable here).
The fixtures travel to the GPU box; /root/reference does not.
"""
import copy
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF_SRC = "/root/reference/src"
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import tf_numpy_shim  # noqa: E402

tf_numpy_shim.install()
sys.path.insert(0, REF_SRC)
import anchors as ref_anchors  # noqa: E402
import nms_np as ref_nms_np  # noqa: E402
import postprocess as ref_post  # noqa: E402
import utils_box as ref_utils_box  # noqa: E402
import utils_extra as ref_utils_extra  # noqa: E402

from oracle import ref_np  # noqa: E402  (only for default_params / level_shapes helpers)


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **{k: np.asarray(v) for k, v in arrays.items()})
    print("wrote", os.path.relpath(path, ROOT), os.path.getsize(path), "bytes")


def synth_head_outputs(params, batch, nsamp, seed, mc=True, la=True):
    """Synthetic raw head outputs with SURVEY 8(d) config-4 distributions."""
    rng = np.random.default_rng(seed)
    a = ref_np.num_anchors_per_location(params)
    c = params["num_classes"]
    lead = (nsamp, batch) if mc else (batch,)
    cls, box = [], []
    for h, w in ref_np.level_shapes(params):
        cls.append(rng.normal(-4.6, 2.0, lead + (h, w, a * c)).astype(np.float32))
        t = rng.normal(0, 1, lead + (h, w, a, 4))
        t[..., :2] *= 0.5
        t[..., 2:] *= 0.25
        parts = [t.reshape(lead + (h, w, a * 4))]
        if la:
            s = np.clip(np.abs(rng.normal(0, 0.3, lead + (h, w, a * 4))), 0.01, 2.0)
            parts.append(s)
        box.append(np.concatenate(parts, -1).astype(np.float32))
    return cls, box


def golden_anchors():
    out = {}
    for tag, size in (("512", 512), ("384x1280", (384, 1280)), ("720x1280", (720, 1280)),
                      ("768", 768), ("str1024x512", "1024x512"), ("64x96", (64, 96))):
        b = np.asarray(ref_anchors.Anchors(3, 7, 3, [1.0, 2.0, 0.5], 4.0, size).boxes)
        n = b.shape[0]
        rows = np.unique(np.concatenate([np.arange(0, 20), np.linspace(0, n - 1, 64).astype(int),
                                         np.arange(n - 20, n)]))
        out[tag + "_n"] = n
        out[tag + "_rows"] = rows
        out[tag + "_vals"] = b[rows]
        out[tag + "_sum64"] = b.astype(np.float64).sum(0)
        out[tag + "_abs64"] = np.abs(b.astype(np.float64)).sum(0)
    b = np.asarray(ref_anchors.Anchors(3, 7, 3, [1.0, 2.0, 0.5], 4.0, (64, 96)).boxes)
    out["full_64x96"] = b
    b = np.asarray(ref_anchors.Anchors(3, 5, 2, [1.0, [1.4, 0.7]], [4.0, 3.0, 5.0], 128).boxes)
    out["custom_128"] = b
    save("anchors", **out)


def golden_decode():
    rng = np.random.default_rng(99)
    anc = np.asarray(ref_anchors.Anchors(3, 7, 3, [1.0, 2.0, 0.5], 4.0, 512).boxes)
    rows = rng.integers(0, anc.shape[0], 512)
    a = anc[rows]
    t = rng.normal(0, 1, (3, 512, 4)).astype(np.float32) * np.float32([0.5, 0.5, 0.25, 0.25])
    s = np.clip(np.abs(rng.normal(0, 0.3, (3, 512, 4))), 0.01, 2.0).astype(np.float32)
    out = dict(anchors=a, t=t, sigma=s)
    for method in ("l-norm", "falsedec"):
        c, u = ref_utils_box.decode_uncert(t, s, a, method=method)
        out["box_" + method], out["std_" + method] = c, u
    out["plain"] = ref_anchors.decode_box_outputs(tf_numpy_shim.cast(t, np.float32), a)
    # the three known answers recorded in SURVEY 8(c)
    ka_t = np.float32([[0.1, -0.2, 0.3, -0.4], [0, 0, 0, 0], [1.0, 0.5, -0.5, 0.25]])
    ka_s = np.float32([[0.05, 0.1, 0.2, 0.3], [0.01] * 4, [0.5, 0.4, 0.3, 0.2]])
    ka_a = anc[[0, 100, 40000]]
    c, u = ref_utils_box.decode_uncert(ka_t, ka_s, ka_a, method="l-norm")
    out.update(ka_t=ka_t, ka_s=ka_s, ka_a=ka_a, ka_box=c, ka_std=u)
    save("decode", **out)


def golden_mcuncert():
    rng = np.random.default_rng(5)
    levels = [rng.normal(-4, 2, (6, 2, h, h, 27)).astype(np.float32) for h in (8, 4, 2, 1, 1)]
    mean, std = ref_utils_extra.get_mcuncert(levels)
    out = {}
    for i in range(5):
        out["in%d" % i], out["mean%d" % i], out["std%d" % i] = levels[i], mean[i], std[i]
    save("mcuncert", **out)


def _params(**kw):
    return ref_np.default_params(**kw)


def golden_postprocess():
    cases = {
        # variant A (serving): max-reduce + global NMS + all uncertainties
        "A_mcla_gauss": dict(params=_params(image_size=(64, 96), num_classes=3, mc_dropoutsamp=4),
                             batch=2, mc=True, la=True),
        "A_mcla_hard": dict(params=_params(image_size=(64, 96), num_classes=3, mc_dropoutsamp=4,
                                           nms_configs=dict(method="hard", iou_thresh=0.4,
                                                            score_thresh=0.02)),
                            batch=2, mc=True, la=True),
        "A_la_only": dict(params=_params(image_size=64, num_classes=4, mc_dropout=False,
                                         mc_classheadrate=0.0, mc_boxheadrate=0.0),
                          batch=3, mc=False, la=True),
        "A_mc_only": dict(params=_params(image_size=64, num_classes=4, loss_attenuation=False,
                                         mc_dropoutsamp=3),
                          batch=2, mc=True, la=False),
        "A_plain": dict(params=_params(image_size=64, num_classes=4, loss_attenuation=False,
                                       mc_dropout=False, mc_classheadrate=0.0, mc_boxheadrate=0.0),
                        batch=2, mc=False, la=False),
        "A_mcla_falsedec": dict(params=_params(image_size=64, num_classes=3, mc_dropoutsamp=3,
                                               uncert_adjust_method="falsedec"),
                                batch=1, mc=True, la=True),
        # variant B (eval): top-k + per-class NMS
        "B_mcla_gauss": dict(params=_params(image_size=(64, 96), num_classes=3, mc_dropoutsamp=4,
                                            nms_configs=dict(max_nms_inputs=300)),
                             batch=2, mc=True, la=True, per_class=True),
        "B_mcla_hard": dict(params=_params(image_size=(64, 96), num_classes=3, mc_dropoutsamp=4,
                                           nms_configs=dict(max_nms_inputs=300, method="hard",
                                                            iou_thresh=0.5, score_thresh=0.0)),
                            batch=2, mc=True, la=True, per_class=True),
        "B_plain_hard": dict(params=_params(image_size=64, num_classes=4, loss_attenuation=False,
                                            mc_dropout=False, mc_classheadrate=0.0,
                                            mc_boxheadrate=0.0,
                                            nms_configs=dict(max_nms_inputs=200, method="hard")),
                             batch=2, mc=False, la=False, per_class=True),
    }
    for seed, (name, case) in enumerate(cases.items()):
        params = case["params"]
        cls, box = synth_head_outputs(params, case["batch"], params["mc_dropoutsamp"], 100 + seed,
                                      mc=case["mc"], la=case["la"])
        scales = np.linspace(1.0, 1.5, case["batch"]).astype(np.float32)
        out = {"scales": scales}
        for i in range(5):
            out["cls%d" % i], out["box%d" % i] = cls[i], box[i]
        # pre_nms (through extract_uncertainties, as every caller does)
        res = ref_post.extract_uncertainties(copy.deepcopy(params), [c.copy() for c in cls],
                                             [b.copy() for b in box])
        boxes, uncerts, scores, classes, multi = res
        out.update(pre_boxes=boxes, pre_scores=scores, pre_classes=classes, pre_multi=multi)
        if uncerts is not None:
            for i, u in enumerate(uncerts):
                if u is not None:
                    out["pre_unc%d" % i] = u
        if case.get("per_class"):
            r = ref_post.postprocess_per_class(copy.deepcopy(params), [c.copy() for c in cls],
                                               [b.copy() for b in box], scales)
            for i, v in enumerate(r):
                out["out%d" % i] = v
            ids = np.arange(case["batch"]).astype(np.float32) + 7
            det = ref_post.generate_detections(copy.deepcopy(params), [c.copy() for c in cls],
                                               [b.copy() for b in box], scales, ids, flip=False)
            out["det"] = det
            out["det_flip"] = ref_post.generate_detections(
                copy.deepcopy(params), [c.copy() for c in cls], [b.copy() for b in box], scales,
                ids, flip=True)
            out["det_xywh"] = ref_post.transform_detections(det)
        else:
            r = ref_post.postprocess_global(copy.deepcopy(params), [c.copy() for c in cls],
                                            [b.copy() for b in box], scales)
            for i, v in enumerate(r):
                out["out%d" % i] = v
        out["params_repr"] = np.array(repr(params))
        save("post_" + name, **out)


def golden_nms_np():
    rng = np.random.default_rng(11)
    n = 400
    ctr = rng.uniform(20, 300, (n, 2))
    wh = rng.uniform(8, 80, (n, 2))
    boxes = np.concatenate([ctr - wh / 2, ctr + wh / 2], 1).astype(np.float32)  # y1 x1 y2 x2
    scores = rng.uniform(0.01, 1.0, n).astype(np.float32)
    classes = rng.integers(0, 5, n).astype(np.int32)
    out = dict(boxes=boxes, scores=scores, classes=classes)
    cfgs = {
        "hard": dict(method="hard", iou_thresh=None, score_thresh=None, sigma=None),
        "hard03": dict(method="hard", iou_thresh=0.3, score_thresh=None, sigma=None),
        "gaussian": dict(method="gaussian", iou_thresh=None, score_thresh=None, sigma=None),
        "gaussian_s03": dict(method="gaussian", iou_thresh=None, score_thresh=0.05, sigma=0.3),
        "linear": dict(method="linear", iou_thresh=None, score_thresh=None, sigma=None),
        "diou": dict(method="diou", iou_thresh=None, score_thresh=None, sigma=None),
    }
    for name, cfg in cfgs.items():
        det = ref_nms_np.per_class_nms(boxes.copy(), scores.copy(), classes.copy(),
                                       np.float32([3.0]), np.float32([1.25]), 5, 100, cfg)
        out["det_" + name] = det
        dets = np.column_stack((boxes[:, [1, 0, 3, 2]], scores))
        out["raw_" + name] = ref_nms_np.nms(dets.copy(), cfg)
    save("nms_np", **out)


if __name__ == "__main__":
    golden_anchors()
    golden_decode()
    golden_mcuncert()
    golden_postprocess()
    golden_nms_np()
