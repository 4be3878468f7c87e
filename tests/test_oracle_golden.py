"""The CPU oracle (oracle/) against the committed golden fixtures (tests/golden/*.npz), which
were produced by executing the reference's own source (see tests/golden/make_golden.py)."""
import ast
import copy
import os

import numpy as np
import pytest

from oracle import nms_np_ref, nms_ref, ref_np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(GOLD, name + ".npz"), allow_pickle=False)


def params_of(g):
    return ast.literal_eval(str(g["params_repr"]).replace("-inf", "-1e999"))


def test_anchor_known_answers():
    # SURVEY 8(c) known answers
    b = ref_np.anchor_boxes(3, 7, 3, [1.0, 2.0, 0.5], 4.0, (512, 512))
    assert b.shape == (49104, 4)
    np.testing.assert_allclose(b[0], [-12, -12, 20, 20])
    np.testing.assert_allclose(b[1], [-7.3137083, -18.627417, 15.313708, 26.627417], rtol=1e-7)
    np.testing.assert_allclose(b[8], [-31.918785, -13.959393, 39.918785, 21.959393], rtol=1e-7)
    np.testing.assert_allclose(b[-1], [-126.70057, 160.64972, 1022.70056, 735.3503], rtol=1e-7)
    assert ref_np.anchor_boxes(3, 7, 3, [1.0, 2.0, 0.5], 4.0, (384, 1280)).shape[0] == 92070
    assert ref_np.anchor_boxes(3, 7, 3, [1.0, 2.0, 0.5], 4.0, (768, 768)).shape[0] == 110484


@pytest.mark.parametrize("tag,size", [("512", 512), ("384x1280", (384, 1280)),
                                      ("720x1280", (720, 1280)), ("768", 768),
                                      ("str1024x512", "1024x512"), ("64x96", (64, 96))])
def test_anchors_vs_reference(tag, size):
    g = load("anchors")
    b = ref_np.anchor_boxes(3, 7, 3, [1.0, 2.0, 0.5], 4.0, size)
    assert b.shape[0] == int(g[tag + "_n"])
    np.testing.assert_array_equal(b[g[tag + "_rows"]], g[tag + "_vals"])
    np.testing.assert_array_equal(b.astype(np.float64).sum(0), g[tag + "_sum64"])
    np.testing.assert_array_equal(np.abs(b.astype(np.float64)).sum(0), g[tag + "_abs64"])


def test_anchors_full_and_custom():
    g = load("anchors")
    np.testing.assert_array_equal(
        ref_np.anchor_boxes(3, 7, 3, [1.0, 2.0, 0.5], 4.0, (64, 96)), g["full_64x96"])
    np.testing.assert_array_equal(
        ref_np.anchor_boxes(3, 5, 2, [1.0, [1.4, 0.7]], [4.0, 3.0, 5.0], 128), g["custom_128"])


@pytest.mark.parametrize("method", ["l-norm", "falsedec"])
def test_decode_uncert_vs_reference(method):
    g = load("decode")
    c, u = ref_np.decode_uncert(g["t"], g["sigma"], g["anchors"], method=method)
    np.testing.assert_array_equal(c, g["box_" + method])
    np.testing.assert_array_equal(u, g["std_" + method])


def test_decode_nflow_equals_lnorm_analytically():
    g = load("decode")
    c, u = ref_np.decode_uncert(g["t"], g["sigma"], g["anchors"], method="n-flow")
    np.testing.assert_allclose(c, g["box_l-norm"], rtol=1e-6)
    np.testing.assert_allclose(u, g["std_l-norm"], rtol=1e-6)


def test_decode_known_answers_and_plain():
    g = load("decode")
    c, u = ref_np.decode_uncert(g["ka_t"], g["ka_s"], g["ka_a"])
    np.testing.assert_array_equal(c, g["ka_box"])
    np.testing.assert_array_equal(u, g["ka_std"])
    np.testing.assert_allclose(c[0], [-14.834044, -13.618775, 29.234045, 8.818775], rtol=1e-6)
    np.testing.assert_allclose(u[0], [4.7300735, 4.7003045, 4.7300735, 4.7003045], rtol=1e-6)
    np.testing.assert_allclose(c[2], [206.93019, 438.3266, 243.10484, 587.70844], rtol=1e-6)
    np.testing.assert_array_equal(ref_np.decode_box_outputs(g["t"], g["anchors"]), g["plain"])


def test_decode_sample_method_converges():
    g = load("decode")
    rng = np.random.default_rng(0)
    t, s, a = g["t"][0, :8], g["sigma"][0, :8], g["anchors"][:8]
    z = rng.standard_normal((20000, 4, 8))
    c, u = ref_np.decode_uncert(t, s, a, method="sample", n_samples=20000, normals=z)
    c0, u0 = ref_np.decode_uncert(t, s, a)
    np.testing.assert_allclose(c, c0, rtol=0.05, atol=1.0)
    np.testing.assert_allclose(u, u0, rtol=0.08)


def test_get_mcuncert_vs_reference():
    g = load("mcuncert")
    mean, std = ref_np.get_mcuncert([g["in%d" % i] for i in range(5)])
    for i in range(5):
        np.testing.assert_array_equal(mean[i], g["mean%d" % i])
        np.testing.assert_array_equal(std[i], g["std%d" % i])


A_CASES = ["A_mcla_gauss", "A_mcla_hard", "A_la_only", "A_mc_only", "A_plain", "A_mcla_falsedec"]
B_CASES = ["B_mcla_gauss", "B_mcla_hard", "B_plain_hard"]


@pytest.mark.parametrize("name", A_CASES + B_CASES)
def test_pre_nms_vs_reference(name):
    g = load("post_" + name)
    params = params_of(g)
    cls = [g["cls%d" % i] for i in range(5)]
    box = [g["box%d" % i] for i in range(5)]
    boxes, uncerts, scores, classes, multi = ref_np.extract_uncertainties(
        copy.deepcopy(params), cls, box)
    np.testing.assert_array_equal(boxes, g["pre_boxes"])
    np.testing.assert_array_equal(scores, g["pre_scores"])
    np.testing.assert_array_equal(classes, g["pre_classes"])
    np.testing.assert_array_equal(multi, g["pre_multi"])
    for i in range(3):
        key = "pre_unc%d" % i
        if key in g.files:
            np.testing.assert_array_equal(uncerts[i], g[key])
        else:
            assert uncerts is None or uncerts[i] is None


@pytest.mark.parametrize("name", A_CASES)
def test_postprocess_global_vs_reference(name):
    g = load("post_" + name)
    params = params_of(g)
    cls = [g["cls%d" % i] for i in range(5)]
    box = [g["box%d" % i] for i in range(5)]
    out = ref_np.postprocess_global(copy.deepcopy(params), cls, box, g["scales"])
    n_out = len([k for k in g.files if k.startswith("out")])
    assert len(out) == n_out
    for i in range(n_out):
        np.testing.assert_array_equal(out[i], g["out%d" % i])


@pytest.mark.parametrize("name", B_CASES)
def test_postprocess_per_class_vs_reference(name):
    g = load("post_" + name)
    params = params_of(g)
    cls = [g["cls%d" % i] for i in range(5)]
    box = [g["box%d" % i] for i in range(5)]
    out = ref_np.postprocess_per_class(copy.deepcopy(params), cls, box, g["scales"],
                                       strict_reference=True)
    n_out = len([k for k in g.files if k.startswith("out")])
    assert len(out) == n_out
    for i in range(n_out):
        np.testing.assert_array_equal(out[i], g["out%d" % i])
    ids = np.arange(cls[0].shape[-4]).astype(np.float32) + 7
    det = ref_np.generate_detections(copy.deepcopy(params), cls, box, g["scales"], ids,
                                     strict_reference=True)
    np.testing.assert_array_equal(det, g["det"])
    det_f = ref_np.generate_detections(copy.deepcopy(params), cls, box, g["scales"], ids, flip=True,
                                       strict_reference=True)
    np.testing.assert_array_equal(det_f, g["det_flip"])
    np.testing.assert_array_equal(ref_np.transform_detections(det), g["det_xywh"])


NMS_NP_CFGS = {
    "hard": dict(method="hard", iou_thresh=None, score_thresh=None, sigma=None),
    "hard03": dict(method="hard", iou_thresh=0.3, score_thresh=None, sigma=None),
    "gaussian": dict(method="gaussian", iou_thresh=None, score_thresh=None, sigma=None),
    "gaussian_s03": dict(method="gaussian", iou_thresh=None, score_thresh=0.05, sigma=0.3),
    "linear": dict(method="linear", iou_thresh=None, score_thresh=None, sigma=None),
    "diou": dict(method="diou", iou_thresh=None, score_thresh=None, sigma=None),
}


@pytest.mark.parametrize("name", sorted(NMS_NP_CFGS))
def test_nms_np_restatement_vs_reference(name):
    g = load("nms_np")
    cfg = NMS_NP_CFGS[name]
    det = nms_np_ref.per_class_nms(g["boxes"].copy(), g["scores"].copy(), g["classes"].copy(),
                                   np.float32([3.0]), np.float32([1.25]), 5, 100, cfg)
    np.testing.assert_array_equal(det, g["det_" + name])
    dets = np.column_stack((g["boxes"][:, [1, 0, 3, 2]], g["scores"]))
    np.testing.assert_array_equal(nms_np_ref.nms(dets.copy(), cfg), g["raw_" + name])


def test_nms_np_bad_method():
    with pytest.raises(ValueError):
        nms_np_ref.nms(np.zeros((1, 5), np.float32), dict(method="bogus"))


@pytest.mark.parametrize("sigma,iou_thr,score_thr,variant", [
    (0.25, 0.5, 0.001, "new"), (0.25, 0.5, 0.001, "old"), (0.0, 0.5, float("-inf"), "new"),
    (0.0, 0.3, 0.05, "new"), (0.0, 0.5, 0.0, "old"), (0.5, 1.0, 0.01, "old"),
])
def test_nms_v5_c_matches_python_transcription(sigma, iou_thr, score_thr, variant):
    rng = np.random.default_rng(3)
    for n in (0, 1, 7, 300):
        ctr = rng.uniform(0, 100, (n, 2))
        wh = rng.uniform(5, 50, (n, 2))
        boxes = np.concatenate([ctr - wh / 2, ctr + wh / 2], 1).astype(np.float32)
        if n > 4:
            boxes[3] = boxes[2]  # exact duplicate
            boxes[4, 2:] = boxes[4, :2]  # zero-area box
        scores = rng.uniform(0, 1, n).astype(np.float32)
        if n > 6:
            scores[5] = scores[6]  # exact tie
        for padded in (True, False):
            a = nms_ref.non_max_suppression_v5(boxes, scores, 20, iou_thr, score_thr, sigma, padded, variant)
            b = nms_ref.non_max_suppression_v5_py(boxes, scores, 20, iou_thr, score_thr, sigma, padded, variant)
            np.testing.assert_array_equal(a[0], b[0])
            np.testing.assert_array_equal(a[1], b[1])
            assert a[2] == b[2]


def test_nms_v5_hard_semantics_small():
    # three boxes: 0 and 1 overlap heavily, 2 is disjoint
    boxes = np.float32([[0, 0, 10, 10], [0, 1, 10, 11], [20, 20, 30, 30]])
    scores = np.float32([0.9, 0.8, 0.7])
    idx, sc, valid = nms_ref.non_max_suppression_v5(boxes, scores, 10, 0.5, float("-inf"), 0.0, False)
    assert idx.tolist() == [0, 2] and valid == 2
    idx, sc, valid = nms_ref.non_max_suppression_v5(boxes, scores, 3, 0.5, 0.001, 0.25, True)
    assert valid == 3 and idx.tolist() == [0, 2, 1]
    assert sc[2] < 0.8 and sc[0] == np.float32(0.9)
