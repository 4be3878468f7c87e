"""GPU parity: the CUDA path (through the C ABI / the Python mirror) against the CPU oracle and
the committed golden fixtures.  Bar: bit-exact for index / integer outputs, rtol 1e-4 for fp32
(plus an atol of a few box ulps where the reference itself cancels, SURVEY hard part 1)."""
import copy
import ctypes

import numpy as np
import pytest

from oracle import nms_ref, ref_np
from tests.helpers import (golden_inputs, golden_params, load_golden, random_boxes, synth_head_outputs)

pytestmark = pytest.mark.gpu

RTOL = 1e-4
BOX_ATOL = 2e-4  # ~ 4 ulp of a 512 px coordinate


@pytest.fixture(scope="module")
def u():
    import udal_b200
    assert udal_b200._lib.device_count() > 0
    return udal_b200


A_CASES = ["A_mcla_gauss", "A_mcla_hard", "A_la_only", "A_mc_only", "A_plain", "A_mcla_falsedec"]
B_CASES = ["B_mcla_gauss", "B_mcla_hard", "B_plain_hard"]


# ---------------------------------------------------------------------------------------------
# golden fixtures (made by executing the reference's own source, tests/golden/make_golden.py)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", A_CASES + B_CASES)
def test_extract_uncertainties_vs_golden(u, name):
    g = load_golden("post_" + name)
    params = golden_params(g)
    cls, box = golden_inputs(g)
    boxes, uncerts, scores, classes, multi = u.postprocess.extract_uncertainties(copy.deepcopy(params), cls, box)
    np.testing.assert_allclose(boxes, g["pre_boxes"], rtol=RTOL, atol=BOX_ATOL)
    np.testing.assert_allclose(scores, g["pre_scores"], rtol=1e-6)
    np.testing.assert_array_equal(classes, g["pre_classes"])
    np.testing.assert_array_equal(multi, g["pre_multi"])
    for i in range(3):
        key = "pre_unc%d" % i
        if key in g.files:
            # std of T fp32 boxes: the reference's own cancellation floor is a few box ulps
            np.testing.assert_allclose(uncerts[i], g[key], rtol=RTOL, atol=BOX_ATOL if i == 2 else 1e-6)
        else:
            assert uncerts is None or uncerts[i] is None
    # the fp64 decode is expected to be bit-exact in all but a vanishing fraction of values
    # (the plain fp32 decode of the no-loss-attenuation modes depends on the last ulp of expf)
    if params["loss_attenuation"]:
        assert np.mean(boxes != g["pre_boxes"]) < 1e-3


@pytest.mark.parametrize("name", A_CASES)
def test_postprocess_global_vs_golden(u, name):
    g = load_golden("post_" + name)
    params = golden_params(g)
    cls, box = golden_inputs(g)
    out = u.postprocess.postprocess_global(copy.deepcopy(params), cls, box, g["scales"])
    n_out = len([k for k in g.files if k.startswith("out")])
    assert len(out) == n_out
    np.testing.assert_array_equal(out[3], g["out3"])                      # valid_len
    cls_col = out[2][..., 0] if out[2].ndim == 3 else out[2]
    ref_cls = g["out2"][..., 0] if g["out2"].ndim == 3 else g["out2"]
    np.testing.assert_array_equal(cls_col, ref_cls)                        # classes => same keep indices
    np.testing.assert_allclose(out[1], g["out1"], rtol=1e-6, atol=1e-7)    # (soft-decayed) scores
    np.testing.assert_allclose(out[0], g["out0"], rtol=RTOL, atol=BOX_ATOL)
    np.testing.assert_allclose(out[2], g["out2"], rtol=RTOL, atol=1e-6)
    np.testing.assert_array_equal(out[4], g["out4"])                       # logits rows (gathered means)


@pytest.mark.parametrize("name", B_CASES)
def test_postprocess_per_class_vs_golden(u, name):
    g = load_golden("post_" + name)
    params = golden_params(g)
    cls, box = golden_inputs(g)
    out = u.postprocess.postprocess_per_class(copy.deepcopy(params), cls, box, g["scales"])
    np.testing.assert_array_equal(out[3], g["out3"])
    np.testing.assert_array_equal(out[2], g["out2"])
    np.testing.assert_allclose(out[1], g["out1"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(out[0], g["out0"], rtol=RTOL, atol=BOX_ATOL)
    np.testing.assert_array_equal(out[4], g["out4"])  # strict reference logits chain
    ids = np.arange(cls[0].shape[-4]).astype(np.float32) + 7
    det = u.postprocess.generate_detections(copy.deepcopy(params), cls, box, g["scales"], ids)
    np.testing.assert_allclose(det, g["det"], rtol=RTOL, atol=BOX_ATOL)
    det_f = u.postprocess.generate_detections(copy.deepcopy(params), cls, box, g["scales"], ids, flip=True)
    np.testing.assert_allclose(det_f, g["det_flip"], rtol=RTOL, atol=BOX_ATOL)
    np.testing.assert_array_equal(u.postprocess.transform_detections(g["det"]), g["det_xywh"])


def test_per_class_correct_logits_flag(u):
    g = load_golden("post_B_mcla_gauss")
    params = dict(golden_params(g), strict_reference=False)
    cls, box = golden_inputs(g)
    out = u.postprocess.postprocess_per_class(copy.deepcopy(params), cls, box, g["scales"])
    # non-strict: logits of a detection = mean logits of ITS anchor; its class column must then
    # hold the (sigmoid^-1 of the) selected score for hard NMS, or at least be the anchor's row
    pre = ref_np.extract_uncertainties(copy.deepcopy(dict(params, strict_reference=True)), cls, box)
    multi = pre[4]
    for b in range(out[0].shape[0]):
        for j in range(int(out[3][b])):
            row = out[4][b, j]
            # the row must be one of the image's mean-logit rows
            assert np.any(np.all(multi[b] == row[None], axis=1))
            c = int(out[2][b, j]) - 1
            if params["nms_configs"]["method"] == "hard":
                assert ref_np.sigmoid(row[c:c + 1])[0] == out[1][b, j]


@pytest.mark.parametrize("method", ["l-norm", "n-flow", "falsedec"])
def test_decode_uncert_vs_golden(u, method):
    g = load_golden("decode")
    c, s = u.utils_box.decode_uncert(g["t"], g["sigma"], g["anchors"], method=method)
    ref_c, ref_s = ref_np.decode_uncert(g["t"], g["sigma"], g["anchors"], method=method)
    key = method if method != "n-flow" else "l-norm"
    np.testing.assert_allclose(c, g["box_" + key], rtol=1e-6, atol=1e-5)
    np.testing.assert_allclose(s, g["std_" + key], rtol=1e-6)
    assert np.mean(c != ref_c) < 1e-3 and np.mean(s != ref_s) < 1e-3
    c, s = u.utils_box.decode_uncert(g["ka_t"], g["ka_s"], g["ka_a"])
    np.testing.assert_array_equal(c, g["ka_box"])
    np.testing.assert_array_equal(s, g["ka_std"])
    np.testing.assert_allclose(u.anchors.decode_box_outputs(g["t"], g["anchors"]), g["plain"], rtol=1e-6, atol=1e-5)
    with pytest.raises(ValueError):
        u.utils_box.decode_uncert(g["t"], g["sigma"], g["anchors"], method="no-such-method")


def test_get_mcuncert_vs_golden(u):
    g = load_golden("mcuncert")
    mean, std = u.utils_extra.get_mcuncert([g["in%d" % i] for i in range(5)])
    for i in range(5):
        np.testing.assert_array_equal(mean[i], g["mean%d" % i])
        np.testing.assert_array_equal(std[i], g["std%d" % i])


# ---------------------------------------------------------------------------------------------
# top-k
# ---------------------------------------------------------------------------------------------
def _engine(u, **kw):
    p = u.hparams_config.get_detection_config("efficientdet-d0", image_size=64, num_classes=3,
                                              enable_softmax=True, **kw)
    return u.engine.get_engine(p)


@pytest.mark.parametrize("m,k,batch", [(1000, 1, 1), (1000, 1000, 2), (5000, 300, 3), (343728, 5000, 2),
                                       (1729800, 5000, 1), (70000, 8192, 1)])
def test_topk_vs_oracle(u, m, k, batch):
    rng = np.random.default_rng(m + k)
    v = rng.normal(-4.6, 2.0, (batch, m)).astype(np.float32)
    v[0, :: max(m // 37, 1)] = v[0, 0]  # exact ties across the row
    if m > 10:
        v[0, 5] = 0.0
        v[0, 6] = -0.0
    eng = _engine(u)
    val, idx = eng.topk(v, k)
    rv, ri = ref_np.top_k(v, k)
    np.testing.assert_array_equal(idx.numpy(), ri)
    np.testing.assert_array_equal(val.numpy(), rv + np.float32(0.0))


def test_topk_degenerate_inputs_use_the_fallback(u):
    eng = _engine(u)
    cap = ctypes.c_int.in_dll(eng.lib, "udal_topk_cand_cap_override")
    rng = np.random.default_rng(0)
    try:
        cap.value = 64  # force the candidate buffer to overflow
        for v in (np.zeros((2, 20000), np.float32),
                  np.repeat(rng.normal(size=(1, 40)).astype(np.float32), 500, axis=1),
                  rng.normal(size=(2, 30000)).astype(np.float32)):
            val, idx = eng.topk(v, 777)
            rv, ri = ref_np.top_k(v, 777)
            np.testing.assert_array_equal(idx.numpy(), ri)
            np.testing.assert_array_equal(val.numpy(), rv)
    finally:
        cap.value = 0
    v = np.full((1, 300000), 1.5, np.float32)  # one huge tie with the default buffer
    val, idx = eng.topk(v, 5000)
    np.testing.assert_array_equal(idx.numpy(), np.arange(5000, dtype=np.int32)[None])


def test_topk_argument_errors(u):
    eng = _engine(u)
    with pytest.raises(ValueError):
        eng.topk(np.zeros((1, 10), np.float32), 11)
    with pytest.raises(ValueError):
        eng.topk(np.zeros((1, 100000), np.float32), 9000)


# ---------------------------------------------------------------------------------------------
# NMS (TF NonMaxSuppressionV5 semantics)
# ---------------------------------------------------------------------------------------------
def _nms_engine(u, method, variant="new", prefilter=0, score_thresh=0.0, iou=None, sigma=None, max_out=100):
    p = u.hparams_config.get_detection_config(
        "efficientdet-d0", image_size=64, num_classes=3, enable_softmax=True, tf_nms_variant=variant,
        nms_prefilter_k=prefilter,
        nms_configs=dict(method=method, score_thresh=score_thresh, iou_thresh=iou, sigma=sigma,
                         max_output_size=max_out))
    return u.engine.get_engine(p), p


@pytest.mark.parametrize("method", ["hard", "gaussian"])
@pytest.mark.parametrize("variant", ["new", "old"])
@pytest.mark.parametrize("n,extent,prefilter", [(0, 300, 0), (1, 300, 0), (50, 100, 0), (700, 200, 0),
                                                (5000, 400, 0), (49104, 512, 0), (3000, 60, 64),
                                                (3000, 300, 16)])
def test_nms_v5_vs_oracle(u, method, variant, n, extent, prefilter):
    eng, p = _nms_engine(u, method, variant, prefilter)
    rng = np.random.default_rng(n + prefilter)
    S = 3
    boxes = np.stack([random_boxes(rng, n, extent) for _ in range(S)]) if n else np.zeros((S, 0, 4), np.float32)
    scores = rng.uniform(0, 1, (S, n)).astype(np.float32)
    if n >= 50:
        boxes[0, 3] = boxes[0, 2]
        boxes[0, 4, 2:] = boxes[0, 4, :2]
        scores[0, 5] = scores[0, 6]
        scores[1, :25] = 0.75  # block of ties
    sigma_tf, iou_thr, thr, max_out = ref_np.nms_thresholds(p["nms_configs"])
    idx, sc, valid = eng.nms_v5(boxes, scores) if n else (None, None, None)
    if n == 0:
        idx, sc, valid = eng.nms_v5(np.zeros((S, 0, 4), np.float32), np.zeros((S, 0), np.float32))
    idx, sc, valid = idx.numpy(), sc.numpy(), valid.numpy()
    for s in range(S):
        ri, rs, rv = nms_ref.non_max_suppression_v5(boxes[s], scores[s], max_out, iou_thr, thr, sigma_tf,
                                                    True, variant)
        assert valid[s] == rv
        np.testing.assert_array_equal(idx[s], ri)
        np.testing.assert_array_equal(sc[s], rs)


@pytest.mark.parametrize("method,variant", [("gaussian", "new"), ("gaussian", "old"), ("hard", "new")])
@pytest.mark.parametrize("n,spread,prefilter", [(2500, 4.0, 0), (2500, 4.0, 256), (6000, 25.0, 0), (1500, 0.0, 128)])
def test_nms_v5_dense_clusters_vs_oracle(u, method, variant, n, spread, prefilter):
    """worst case of the lazy soft-NMS: every candidate overlaps every selection (one cluster of jittered boxes, or n
    copies of ONE box with distinct / equal scores), with and without a truncating pre-filter (exact redo of the flagged
    images), through the cooperative CTA kernels and through the round-1 one-warp kernels"""
    import ctypes
    eng, p = _nms_engine(u, method, variant, prefilter)
    rng = np.random.default_rng(n + prefilter)
    S = 2
    ctr = np.float32([200.0, 300.0]) + rng.normal(0, spread, (S, n, 2)).astype(np.float32)
    hw = rng.uniform(60, 90, (S, n, 2)).astype(np.float32) if spread > 0 else np.full((S, n, 2), 70.0, np.float32)
    boxes = np.concatenate([ctr - hw / 2, ctr + hw / 2], -1).astype(np.float32)
    scores = rng.uniform(0.05, 1.0, (S, n)).astype(np.float32)
    scores[1, ::3] = 0.5      # many exact ties: decayed scores collide, the heap's index rule decides
    sigma_tf, iou_thr, thr, max_out = ref_np.nms_thresholds(p["nms_configs"])
    ref = [nms_ref.non_max_suppression_v5(boxes[s], scores[s], max_out, iou_thr, thr, sigma_tf, True, variant) for s in range(S)]
    switch = ctypes.c_int.in_dll(eng.lib, "udal_nms_cta")
    try:
        for cta in (1, 0):
            switch.value = cta
            idx, sc, valid = (a.numpy() for a in eng.nms_v5(boxes, scores))
            for s in range(S):
                ri, rs, rv = ref[s]
                assert valid[s] == rv, (cta, s)
                np.testing.assert_array_equal(idx[s], ri)
                np.testing.assert_array_equal(sc[s], rs)
    finally:
        switch.value = 1


def test_nms_mirror_padded_and_unpadded(u):
    eng, p = _nms_engine(u, "gaussian")
    rng = np.random.default_rng(5)
    n = 400
    boxes, scores = random_boxes(rng, n, 150), rng.uniform(0, 1, n).astype(np.float32)
    classes = rng.integers(0, 3, n).astype(np.int32)
    multi = rng.normal(size=(n, 3)).astype(np.float32)
    u1 = rng.normal(size=(n, 3)).astype(np.float32)
    u2 = rng.normal(size=(n, 4)).astype(np.float32)
    for padded in (True, False):
        got = u.postprocess.nms(p, boxes, scores, classes, padded, multiclass=multi, uncerts1=u1,
                                uncerts2=u2, uncerts3=u2)
        ref = ref_np.nms(p, boxes, scores, classes, padded, multiclass=multi, uncerts1=u1, uncerts2=u2,
                         uncerts3=u2)
        assert len(got) == len(ref) == 8
        for a, b in zip(got, ref):
            np.testing.assert_array_equal(np.asarray(a), np.asarray(b))
    with pytest.raises(ValueError, match="invalid nms method"):
        bad = copy.deepcopy(p)
        bad["nms_configs"]["method"] = "linear"
        u.postprocess.nms(bad, boxes, scores, classes, True)


def test_per_class_nms_mirror_vs_oracle(u):
    for method in ("hard", "gaussian"):
        eng, p = _nms_engine(u, method)
        rng = np.random.default_rng(8)
        B, K = 2, 600
        boxes = np.stack([random_boxes(rng, K, 120) for _ in range(B)])
        scores = -np.sort(-rng.uniform(0, 1, (B, K)).astype(np.float32), axis=1)
        classes = rng.integers(0, 3, (B, K)).astype(np.int32)
        classes[1][classes[1] == 1] = 2  # an empty class
        scales = np.float32([1.0, 2.0])
        logits = rng.normal(size=(B, 774, 3)).astype(np.float32)
        got = u.postprocess.per_class_nms(p, boxes, scores, classes, scales, logits)
        ref = ref_np.per_class_nms(p, boxes, scores, classes, scales, logits, strict_reference=True)
        for a, b in zip(got, ref):
            np.testing.assert_array_equal(a, b)
        got = u.postprocess.per_class_nms(p, boxes, scores, classes, None, None)
        ref = ref_np.per_class_nms(p, boxes, scores, classes, None, None)
        assert len(got) == len(ref) == 4
        for a, b in zip(got, ref):
            np.testing.assert_array_equal(a, b)


# ---------------------------------------------------------------------------------------------
# random configurations against the oracle (not only the golden ones)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("C,T,batch,size,method", [(7, 10, 2, (64, 96), "gaussian"), (8, 1, 1, 64, "hard"),
                                                   (10, 30, 1, 64, "gaussian"), (90, 4, 2, 64, "hard"),
                                                   (3, 40, 1, 64, "gaussian"),
                                                   # 17..32 samples: chunked loads of the decode kernel (chunk edges)
                                                   (8, 20, 2, (64, 96), "gaussian"), (7, 17, 1, 64, "hard"),
                                                   (8, 32, 1, 64, "gaussian"), (10, 24, 3, (64, 96), "gaussian")])
def test_postprocess_global_vs_oracle_random(u, C, T, batch, size, method):
    p = u.hparams_config.get_detection_config(
        "efficientdet-d0", image_size=size, num_classes=C, enable_softmax=True, loss_attenuation=True,
        mc_dropout=True, mc_classheadrate=0.05, mc_boxheadrate=0.05, mc_dropoutsamp=T,
        nms_configs=dict(method=method))
    cls, box = synth_head_outputs(p, batch, seed=C * 100 + T)
    scales = np.linspace(1, 2, batch).astype(np.float32)
    got = u.postprocess.postprocess_global(p, cls, box, scales)
    ref = ref_np.postprocess_global(copy.deepcopy(p), cls, box, scales)
    np.testing.assert_array_equal(got[3], ref[3])
    np.testing.assert_array_equal(got[2][..., 0], ref[2][..., 0])
    np.testing.assert_allclose(got[1], ref[1], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(got[0], ref[0], rtol=RTOL, atol=BOX_ATOL)
    np.testing.assert_allclose(got[2], ref[2], rtol=RTOL, atol=1e-6)
    np.testing.assert_array_equal(got[4], ref[4])


@pytest.mark.parametrize("C,T,batch,k,method", [(7, 10, 2, 500, "gaussian"), (10, 3, 1, 5000, "hard"),
                                                (90, 2, 1, 2000, "gaussian")])
@pytest.mark.parametrize("seg", [1, 0])   # soft NMS of a few segments: one CTA per segment (nms_cta.cu) | one warp (nms.cu)
def test_postprocess_per_class_vs_oracle_random(u, C, T, batch, k, method, seg):
    switch = ctypes.c_int.in_dll(u._lib.load(), "udal_nms_seg")
    switch.value = seg
    try:
        _per_class_vs_oracle_random(u, C, T, batch, k, method)
    finally:
        switch.value = 1


def _per_class_vs_oracle_random(u, C, T, batch, k, method):
    p = u.hparams_config.get_detection_config(
        "efficientdet-d0", image_size=(64, 96), num_classes=C, enable_softmax=True, loss_attenuation=True,
        mc_dropout=True, mc_classheadrate=0.05, mc_boxheadrate=0.05, mc_dropoutsamp=T,
        nms_configs=dict(method=method, max_nms_inputs=k))
    cls, box = synth_head_outputs(p, batch, seed=C + T + k)
    scales = np.linspace(1, 2, batch).astype(np.float32)
    got = u.postprocess.postprocess_per_class(p, cls, box, scales)
    ref = ref_np.postprocess_per_class(copy.deepcopy(p), cls, box, scales, strict_reference=True)
    np.testing.assert_array_equal(got[3], ref[3])
    np.testing.assert_array_equal(got[2], ref[2])
    np.testing.assert_allclose(got[1], ref[1], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(got[0], ref[0], rtol=RTOL, atol=BOX_ATOL)
    np.testing.assert_array_equal(got[4], ref[4])


# BASELINE.json geometries at full size (VERDICT r1 "what's weak" #3): configs[3] = 49 104 anchors x 10 classes with
# T in {1, 10, 30} (variant A global and variant B top-k 5000 + per-class NMS), configs[1] 92 070, configs[2] 172 980 with
# the non-integer strides of 720 -> 23 rows, configs[4] D2 768^2 = 110 484.  One or two images: the NumPy oracle decodes
# T x N x 4 values per image.
BIG_CASES = [
    # model, size, C, T, batch, method, max_nms_inputs
    ("efficientdet-d0", 512, 10, 1, 1, "gaussian", 0),
    ("efficientdet-d0", 512, 10, 10, 2, "gaussian", 0),
    ("efficientdet-d0", 512, 10, 30, 1, "gaussian", 0),
    ("efficientdet-d0", 512, 10, 30, 1, "hard", 0),
    ("efficientdet-d0", 512, 10, 1, 1, "hard", 5000),
    ("efficientdet-d0", 512, 10, 10, 2, "gaussian", 5000),
    ("efficientdet-d0", 512, 10, 30, 1, "hard", 5000),
    ("efficientdet-d0", (384, 1280), 8, 10, 1, "gaussian", 0),
    ("efficientdet-d0", (720, 1280), 10, 20, 1, "gaussian", 0),
    ("efficientdet-d0", (720, 1280), 10, 20, 1, "hard", 5000),
    ("efficientdet-d2", 768, 10, 30, 1, "gaussian", 0),
]


@pytest.mark.parametrize("model,size,C,T,batch,method,k", BIG_CASES)
def test_postprocess_vs_oracle_baseline_geometries(u, model, size, C, T, batch, method, k):
    p = u.hparams_config.get_detection_config(
        model, image_size=size, num_classes=C, enable_softmax=True, loss_attenuation=True,
        mc_dropout=True, mc_classheadrate=0.05, mc_boxheadrate=0.05, mc_dropoutsamp=T,
        nms_configs=dict(method=method, max_nms_inputs=k))
    cls, box = synth_head_outputs(p, batch, seed=99)
    scales = np.linspace(1, 2, batch).astype(np.float32)
    if k == 0:
        got = u.postprocess.postprocess_global(p, cls, box, scales)
        ref = ref_np.postprocess_global(copy.deepcopy(p), cls, box, scales)
        np.testing.assert_array_equal(got[3], ref[3])
        np.testing.assert_array_equal(got[2][..., 0], ref[2][..., 0])
        np.testing.assert_allclose(got[1], ref[1], rtol=1e-6, atol=1e-7)
        # boxes up to 1280 px: the reference's own cancellation floor (SURVEY hard part 1) is a few ulp of the coordinate
        np.testing.assert_allclose(got[0], ref[0], rtol=RTOL, atol=3 * BOX_ATOL)
        np.testing.assert_allclose(got[2], ref[2], rtol=RTOL, atol=1e-6)
        np.testing.assert_array_equal(got[4], ref[4])
        # per anchor, before NMS
        pre_g = u.postprocess.extract_uncertainties(copy.deepcopy(p), cls, box)
        pre_r = ref_np.extract_uncertainties(copy.deepcopy(p), cls, box)
        np.testing.assert_allclose(pre_g[0], pre_r[0], rtol=RTOL, atol=3 * BOX_ATOL)
        np.testing.assert_array_equal(pre_g[3], pre_r[3])
        np.testing.assert_allclose(pre_g[2], pre_r[2], rtol=1e-6)
        np.testing.assert_array_equal(pre_g[4], pre_r[4])
        for a, b, atol in zip(pre_g[1], pre_r[1], (1e-6, 1e-6, 3 * BOX_ATOL)):
            np.testing.assert_allclose(a, b, rtol=RTOL, atol=atol)
        assert np.mean(pre_g[0] != pre_r[0]) < 1e-3
    else:
        got = u.postprocess.postprocess_per_class(p, cls, box, scales)
        ref = ref_np.postprocess_per_class(copy.deepcopy(p), cls, box, scales, strict_reference=True)
        np.testing.assert_array_equal(got[3], ref[3])
        np.testing.assert_array_equal(got[2], ref[2])
        np.testing.assert_allclose(got[1], ref[1], rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(got[0], ref[0], rtol=RTOL, atol=3 * BOX_ATOL)
        np.testing.assert_array_equal(got[4], ref[4])


# decode_precision = "fp32": the closed form in fp32 (the arithmetic of udal_run's fused kernels) in the stand-alone K2 kernel.
# Contract = BASELINE.json's: decoded boxes, variances and scores within 1e-4 relative (boxes: relative to the box side, the
# scale the error lives on - a corner is a difference of centre and half size); mean logits / classes stay bit-exact.
F32_CASES = [
    ("efficientdet-d0", (64, 96), 7, 4, 2, "l-norm", True),
    ("efficientdet-d0", 512, 10, 1, 1, "l-norm", True),
    ("efficientdet-d0", 512, 10, 10, 2, "l-norm", True),
    ("efficientdet-d0", 512, 10, 30, 1, "n-flow", True),
    ("efficientdet-d0", 512, 10, 10, 1, "falsedec", True),
    ("efficientdet-d0", 512, 7, 10, 1, "l-norm", False),       # plain decode, 63 logits per pixel (unaligned tiles)
    ("efficientdet-d0", (384, 1280), 8, 10, 1, "l-norm", True),
    ("efficientdet-d0", (720, 1280), 10, 20, 1, "l-norm", True),
    ("efficientdet-d2", 768, 10, 30, 1, "l-norm", True),
]


@pytest.mark.parametrize("stream", [1, 0])   # persistent TMA-staged kernel | per-tile kernel (udal_decode_stream)
@pytest.mark.parametrize("model,size,C,T,batch,method,la", F32_CASES)
def test_decode_precision_fp32_vs_oracle(u, model, size, C, T, batch, method, la, stream):
    switch = ctypes.c_int.in_dll(u._lib.load(), "udal_decode_stream")
    switch.value = stream
    try:
        _check_decode_fp32(u, model, size, C, T, batch, method, la)
    finally:
        switch.value = 1


@pytest.mark.parametrize("stream", [1, 0])
@pytest.mark.parametrize("rate_cls,rate_box", [(0.05, 0.0), (0.0, 0.05)])
def test_decode_precision_fp32_one_head_with_mc(u, rate_cls, rate_box, stream):
    """MC dropout on one head only: the sample axes of the two inputs differ (T and 1)."""
    switch = ctypes.c_int.in_dll(u._lib.load(), "udal_decode_stream")
    switch.value = stream
    try:
        _check_decode_fp32(u, "efficientdet-d0", (64, 96), 8, 5, 2, "l-norm", True, rate_cls, rate_box)
    finally:
        switch.value = 1


def _check_decode_fp32(u, model, size, C, T, batch, method, la, rate_cls=0.05, rate_box=0.05):
    p = u.hparams_config.get_detection_config(
        model, image_size=size, num_classes=C, enable_softmax=True, loss_attenuation=la,
        mc_dropout=True, mc_classheadrate=rate_cls, mc_boxheadrate=rate_box, mc_dropoutsamp=T, uncert_adjust_method=method,
        nms_configs=dict(method="gaussian", max_nms_inputs=0), decode_precision="fp32")
    cls, box = synth_head_outputs(p, batch, seed=7, la=la, mc_cls=rate_cls > 0, mc_box=rate_box > 0)
    boxes, unc, scores, classes, multi = u.postprocess.extract_uncertainties(copy.deepcopy(p), cls, box)
    pr = copy.deepcopy(p)
    pr.pop("decode_precision")
    rb, runc, rs, rc, rm = ref_np.extract_uncertainties(pr, cls, box)
    np.testing.assert_array_equal(multi, rm)       # sequential fp32 sum / T: unchanged
    np.testing.assert_array_equal(classes, rc)
    np.testing.assert_allclose(scores, rs, rtol=2e-6)
    side = np.maximum(np.maximum(rb[..., 2] - rb[..., 0], rb[..., 3] - rb[..., 1]), 1.0)[..., None]
    err = np.abs(boxes.astype(np.float64) - rb) / side
    assert err.max() < 1e-4, err.max()
    names = ("mcclass", "albox", "mcbox")
    for i, (a, b) in enumerate(zip(unc, runc)):
        if b is None:
            assert a is None
            continue
        if names[i] == "mcbox":  # std of T corners that differ by a fraction of the side
            e = np.abs(a.astype(np.float64) - b) / np.maximum(b, 1e-4 * side)
        elif names[i] == "albox" and method == "falsedec":
            # utils_box.py:186-266 takes sqrt(|dc - dhalf|) of two decoded "variances": no relative bound where they cancel
            e = np.abs(a.astype(np.float64) - b) / np.maximum(b, 1e-2 * np.sqrt(side))
        else:
            e = np.abs(a.astype(np.float64) - b) / np.maximum(np.abs(b), 1e-3)
        assert e.max() < 1e-4, (names[i], e.max())
    # through NMS: the detections are those of the oracle up to ties the 1e-4 perturbation can flip
    scales = np.linspace(1, 2, batch).astype(np.float32)
    got = u.postprocess.postprocess_global(copy.deepcopy(p), cls, box, scales)
    ref = ref_np.postprocess_global(copy.deepcopy(pr), cls, box, scales)
    assert np.all(np.abs(got[3].astype(int) - ref[3].astype(int)) <= 1)
    gcls = got[2][..., 0] if got[2].ndim == 3 else got[2]     # (no MC on the class head: no logit-std columns)
    rcls = ref[2][..., 0] if ref[2].ndim == 3 else ref[2]
    matched = total = 0
    for b in range(batch):
        n = int(ref[3][b])
        total += n
        for j in range(n):
            d = np.abs(got[0][b, :int(got[3][b])] - ref[0][b, j]).max(-1)
            k = int(np.argmin(d)) if d.size else -1
            s = max(ref[0][b, j, 2] - ref[0][b, j, 0], ref[0][b, j, 3] - ref[0][b, j, 1], 1.0)
            if k >= 0 and d[k] < 2e-4 * s * scales[b] + 1e-4 and gcls[b, k] == rcls[b, j] \
                    and abs(got[1][b, k] - ref[1][b, j]) < 1e-5:
                matched += 1
    assert matched >= 0.99 * total, (matched, total)


@pytest.mark.parametrize("method,T", [("hard", 10), ("gaussian", 4)])
def test_decode_precision_fp32_per_class_variant(u, method, T):
    """top-k + per-class NMS with the fp32 decode of the gathered rows: same selections (mean logits, hence the top-k and its
    scores, are bit-exact), boxes and stds at the 1e-4 contract"""
    p = u.hparams_config.get_detection_config(
        "efficientdet-d0", image_size=(128, 192), num_classes=10, enable_softmax=True, loss_attenuation=True,
        mc_dropout=True, mc_classheadrate=0.05, mc_boxheadrate=0.05, mc_dropoutsamp=T,
        nms_configs=dict(method=method, max_nms_inputs=1000), decode_precision="fp32")
    cls, box = synth_head_outputs(p, 2, seed=11)
    scales = np.float32([1.0, 1.5])
    got = u.postprocess.postprocess_per_class(copy.deepcopy(p), cls, box, scales)
    pr = copy.deepcopy(p)
    pr.pop("decode_precision")
    ref = ref_np.postprocess_per_class(pr, cls, box, scales, strict_reference=True)
    assert np.all(np.abs(got[3].astype(int) - ref[3].astype(int)) <= 1)
    gcls = got[2][..., 0] if got[2].ndim == 3 else got[2]
    rcls = ref[2][..., 0] if ref[2].ndim == 3 else ref[2]
    matched = total = 0
    for b in range(2):
        n = int(ref[3][b])
        total += n
        for j in range(n):
            d = np.abs(got[0][b, :int(got[3][b]), :4] - ref[0][b, j, :4]).max(-1)
            k = int(np.argmin(d)) if d.size else -1
            side = max(ref[0][b, j, 2] - ref[0][b, j, 0], ref[0][b, j, 3] - ref[0][b, j, 1], 1.0)
            if k >= 0 and d[k] < 2e-4 * side + 1e-4 and gcls[b, k] == rcls[b, j] and abs(got[1][b, k] - ref[1][b, j]) < 1e-5:
                np.testing.assert_allclose(got[0][b, k, 4:], ref[0][b, j, 4:], rtol=2e-4, atol=2e-4 * side)   # albox | mcbox
                matched += 1
    assert matched >= 0.98 * total, (matched, total)


def test_device_arrays_in_device_arrays_out_and_dlpack(u):
    import torch
    g = load_golden("post_A_mcla_gauss")
    params = golden_params(g)
    cls, box = golden_inputs(g)
    # the inputs are WRITTEN by kernels on torch's stream right before the call and nobody synchronises: the import
    # path itself must order the context's (non-blocking) stream behind the producer (device.py)
    big = torch.zeros(64 << 20, device="cuda")
    for _ in range(4):
        big = big * 1.0001 + 1.0   # keeps torch's stream busy while the tensors below are produced behind it
    tcls = [(torch.from_numpy(c).cuda() * 2.0) / 2.0 for c in cls]
    tbox = [(torch.from_numpy(b).cuda() * 2.0) / 2.0 for b in box]
    out = u.postprocess.postprocess_global(copy.deepcopy(params), tcls, tbox, g["scales"])
    assert all(isinstance(o, u.device.DeviceArray) for o in out)
    back = torch.from_dlpack(out[0])  # zero-copy export
    assert back.is_cuda and tuple(back.shape) == g["out0"].shape
    np.testing.assert_allclose(back.cpu().numpy(), g["out0"], rtol=RTOL, atol=BOX_ATOL)
    t2 = torch.as_tensor(out[1], device="cuda")  # __cuda_array_interface__
    np.testing.assert_allclose(t2.cpu().numpy(), g["out1"], rtol=1e-6, atol=1e-7)


def test_minor_mirror_entry_points_vs_oracle(u):
    """merge_class_box_level_outputs / topk_class_boxes (both branches) / clip_boxes / batch_map_fn /
    pre_nms(topk=False): names the reference exports (postprocess.py:53-141, 276-282), host and device inputs."""
    p = u.hparams_config.get_detection_config(
        "efficientdet-d0", image_size=(64, 96), num_classes=7, enable_softmax=True, loss_attenuation=True,
        mc_dropout=True, mc_classheadrate=0.05, mc_boxheadrate=0.05, mc_dropoutsamp=3)
    rng = np.random.default_rng(12)
    levels = ref_np.level_shapes(p)
    cls = [rng.normal(-4, 2, (2, h, w, 63)).astype(np.float32) for h, w in levels]
    box = [rng.normal(0, 1, (2, h, w, 36)).astype(np.float32) for h, w in levels]
    mc, mb = u.postprocess.merge_class_box_level_outputs(p, cls, box)
    rc, rb = ref_np.merge_class_box_level_outputs(p, cls, box)
    np.testing.assert_array_equal(mc, rc)
    np.testing.assert_array_equal(mb, rb)
    eng = u.engine.get_engine(p)
    dc, db = u.postprocess.merge_class_box_level_outputs(p, [eng.ctx.to_device(c) for c in cls], [eng.ctx.to_device(b) for b in box])
    np.testing.assert_array_equal(dc.numpy(), rc)
    np.testing.assert_array_equal(db.numpy(), rb)
    # topk_class_boxes: max-reduce branch and top-k branch (canonical order = the oracle's)
    unc = [np.abs(rng.normal(size=rc.shape)).astype(np.float32), np.abs(rng.normal(size=rb.shape)).astype(np.float32), None]
    for k in (0, 300):
        pk = copy.deepcopy(p)
        pk["nms_configs"]["max_nms_inputs"] = k
        got = u.postprocess.topk_class_boxes(pk, rc, rb, [x if x is None else x.copy() for x in unc])
        ref = ref_np.topk_class_boxes(pk, rc, rb, [x if x is None else x.copy() for x in unc])
        for a, b in zip(got[:4], ref[:4]):
            np.testing.assert_array_equal(np.asarray(a), np.asarray(b))
        for a, b in zip(got[4], ref[4]):
            assert (a is None) == (b is None)
            if a is not None:
                np.testing.assert_array_equal(a, b)
        got4 = u.postprocess.topk_class_boxes(pk, rc, rb)
        assert len(got4) == 4
    # clip_boxes
    bx = rng.uniform(-50, 150, (2, 30, 4)).astype(np.float32)
    np.testing.assert_array_equal(u.postprocess.clip_boxes(bx, (64, 96)), ref_np.clip_boxes(bx, (64, 96)))
    # batch_map_fn
    outs = u.postprocess.batch_map_fn(lambda e: [e[0] * 2, e[1].sum()], [bx, bx[..., 0]])
    np.testing.assert_array_equal(outs[0], bx * 2)
    np.testing.assert_array_equal(outs[1], bx[..., 0].sum(1))
    # pre_nms(topk=False): all anchors, all classes, no class ids
    T = 3
    tbox = [rng.normal(0, 0.4, (T, 2, h, w, 36)).astype(np.float32) for h, w in levels]
    sig = [np.abs(rng.normal(0, 0.3, (T, 2, h, w, 36))).astype(np.float32) + 0.01 for h, w in levels]
    std = [np.abs(rng.normal(0, 0.3, c.shape)).astype(np.float32) for c in cls]
    got = u.postprocess.pre_nms(p, cls, tbox, topk=False, uncerts=[[x.copy() for x in std], [x.copy() for x in sig], None])
    ref = ref_np.pre_nms(copy.deepcopy(p), cls, tbox, topk=False, uncerts=[[x.copy() for x in std], [x.copy() for x in sig], None])
    np.testing.assert_allclose(got[0], ref[0], rtol=RTOL, atol=BOX_ATOL)
    np.testing.assert_allclose(got[2], ref[2], rtol=1e-6)
    assert got[3] is None and ref[3] is None and got[2].shape == rc.shape
    np.testing.assert_array_equal(got[4], ref[4])
    for a, b in zip(got[1], ref[1]):
        np.testing.assert_allclose(a, b, rtol=RTOL, atol=BOX_ATOL)


def test_nms_mirror_one_dimensional_uncertainties(u):
    """Reference top-k + class-MC gives uncerts1 of shape [k] (postprocess.py:116-121): a 1-D source must be gathered
    with width 1 (it used to be gathered with width N - an out-of-bounds device write)."""
    eng, p = _nms_engine(u, "hard")
    rng = np.random.default_rng(15)
    n = 300
    boxes, scores = random_boxes(rng, n, 150), rng.uniform(0, 1, n).astype(np.float32)
    classes = rng.integers(0, 3, n).astype(np.int32)
    u1 = rng.normal(size=n).astype(np.float32)
    u2 = rng.normal(size=(n, 4)).astype(np.float32)
    for padded in (True, False):
        got = u.postprocess.nms(p, boxes, scores, classes, padded, multiclass=u1, uncerts1=u1, uncerts2=u2, uncerts3=u2)
        ref = ref_np.nms(p, boxes, scores, classes, padded, multiclass=u1, uncerts1=u1, uncerts2=u2, uncerts3=u2)
        for a, b in zip(got, ref):
            np.testing.assert_array_equal(np.asarray(a), np.asarray(b))
    with pytest.raises(ValueError):
        u.postprocess.nms(p, boxes, scores, classes, True, uncerts1=u1[:-1], uncerts2=u2, uncerts3=u2)


def test_per_class_nms_ignores_out_of_range_class_ids(u):
    """class ids outside [0, num_classes) belong to no class segment: the reference loops over range(num_classes)
    (postprocess.py:655-657).  They used to corrupt the partition kernel's shared counters."""
    eng, p = _nms_engine(u, "hard")
    rng = np.random.default_rng(21)
    B, K = 2, 200
    boxes = np.stack([random_boxes(rng, K, 120) for _ in range(B)])
    scores = -np.sort(-rng.uniform(0, 1, (B, K)).astype(np.float32), axis=1)
    classes = rng.integers(0, p["num_classes"], (B, K)).astype(np.int32)
    bad = classes.copy()
    bad[:, ::7] = -3
    bad[:, 1::11] = p["num_classes"] + 5
    got = u.postprocess.per_class_nms(p, boxes, scores, bad, None, None)
    ref = ref_np.per_class_nms(p, boxes, scores, bad, None, None)
    for a, b in zip(got, ref):
        np.testing.assert_array_equal(a, b)


def test_decode_uncert_sample_method(u):
    """utils_box.py:162-184 with injected normals (the oracle consumes the same draws), and the in-kernel Philox
    stream against its NumPy twin."""
    rng = np.random.default_rng(31)
    n, S = 500, 40
    anchors = ref_np.anchor_boxes(3, 7, 3, [1.0, 2.0, 0.5], 4.0, (64, 96))[:n]
    t = rng.normal(0, 0.4, (n, 4)).astype(np.float32)
    sg = (np.abs(rng.normal(0, 0.3, (n, 4))) + 0.01).astype(np.float32)
    z = rng.standard_normal((S, 4, n)).astype(np.float32)
    got = u.utils_box.decode_uncert(t, sg, anchors, method="sample", n_samples=S, normals=z)
    ref = ref_np.decode_uncert(t, sg, anchors, method="sample", n_samples=S, normals=z)
    np.testing.assert_allclose(got[0], ref[0], rtol=1e-6, atol=1e-4)
    np.testing.assert_allclose(got[1], ref[1], rtol=1e-5, atol=1e-5)
    seed = 0x1234ABCD5678
    got = u.utils_box.decode_uncert(t, sg, anchors, method="sample", n_samples=S, seed=seed)
    twin = u.utils_box.philox_normals(S, n, seed)
    ref = ref_np.decode_uncert(t, sg, anchors, method="sample", n_samples=S, normals=twin)
    np.testing.assert_allclose(got[0], ref[0], rtol=1e-6, atol=1e-4)
    np.testing.assert_allclose(got[1], ref[1], rtol=1e-5, atol=1e-5)
    assert abs(float(twin.mean())) < 0.02 and abs(float(twin.std()) - 1.0) < 0.02
    # the moments converge to the closed form of the l-norm method
    many = u.utils_box.decode_uncert(t, sg, anchors, method="sample", n_samples=4000, seed=7)
    exact = u.utils_box.decode_uncert(t, sg, anchors, method="l-norm")
    assert np.median(np.abs(many[1] - exact[1]) / exact[1]) < 0.03
    a, b = (u.utils_box.decode_uncert(t, sg, anchors, method="sample", n_samples=S) for _ in range(2))
    assert not np.array_equal(a[0], b[0])   # seed=None: fresh draws per call


def test_error_behaviour_matches_reference(u):
    g = load_golden("post_A_mcla_gauss")
    params = golden_params(g)
    cls, box = golden_inputs(g)
    bad = copy.deepcopy(params)
    bad["nms_configs"]["method"] = "bogus"
    with pytest.raises(ValueError, match="invalid nms method"):
        u.postprocess.postprocess_global(bad, cls, box)
    no_soft = dict(copy.deepcopy(params), enable_softmax=False)
    assert u.postprocess.extract_uncertainties(no_soft, cls, box) is None
    with pytest.raises(TypeError):
        u.postprocess.postprocess_global(no_soft, cls, box)
    with pytest.raises(ValueError):
        u.postprocess.postprocess_global(params, cls[:4], box[:4])
    one = dict(copy.deepcopy(params), mc_dropoutsamp=1)
    with pytest.raises(ValueError):  # T == 1 is only defined for batch 1 (postprocess.py:180-203)
        u.postprocess.postprocess_global(one, [c[:1] for c in cls], [b[:1] for b in box])


# ---------------------------------------------------------------------------------------------
# nms_np family (device kernels) against the fixtures produced by the reference's own nms_np.py
# ---------------------------------------------------------------------------------------------
NMS_NP_CFGS = {
    "hard": dict(method="hard", iou_thresh=None, score_thresh=None, sigma=None),
    "hard03": dict(method="hard", iou_thresh=0.3, score_thresh=None, sigma=None),
    "gaussian": dict(method="gaussian", iou_thresh=None, score_thresh=None, sigma=None),
    "gaussian_s03": dict(method="gaussian", iou_thresh=None, score_thresh=0.05, sigma=0.3),
    "linear": dict(method="linear", iou_thresh=None, score_thresh=None, sigma=None),
    "diou": dict(method="diou", iou_thresh=None, score_thresh=None, sigma=None),
}


@pytest.mark.parametrize("name", sorted(NMS_NP_CFGS))
def test_nms_np_device_vs_reference_fixtures(u, name):
    g = load_golden("nms_np")
    cfg = NMS_NP_CFGS[name]
    dets = np.column_stack((g["boxes"][:, [1, 0, 3, 2]], g["scores"]))
    raw = u.nms_np.nms(dets.copy(), cfg)
    ref = g["raw_" + name]
    assert raw.shape == ref.shape
    np.testing.assert_array_equal(raw[:, :4], ref[:, :4])           # same boxes in the same order
    # gaussian decay goes through fp32 exp (NumPy SIMD vs CUDA expf): a few ulp on the scores
    np.testing.assert_allclose(raw[:, 4], ref[:, 4], rtol=2e-6 if "gaussian" in name else 0, atol=0)
    det = u.nms_np.per_class_nms(g["boxes"].copy(), g["scores"].copy(), g["classes"].copy(),
                                 np.float32([3.0]), np.float32([1.25]), 5, 100, cfg)
    np.testing.assert_allclose(det, g["det_" + name], rtol=2e-6 if "gaussian" in name else 0, atol=0)


def test_nms_np_edge_cases(u):
    from oracle import nms_np_ref
    cfg = dict(method="hard", iou_thresh=0.5, score_thresh=None, sigma=None)
    one = np.float32([[0, 0, 10, 10, 0.5]])
    np.testing.assert_array_equal(u.nms_np.nms(one, cfg), nms_np_ref.nms(one.copy(), cfg))
    assert u.nms_np.hard_nms(np.zeros((0, 5), np.float32)).shape == (0, 5)
    with pytest.raises(ValueError, match="Unknown NMS method"):
        u.nms_np.nms(one, dict(method="bogus"))
    # no size limit (the reference has none): > 16384 boxes sort in global scratch instead of shared memory
    rng = np.random.default_rng(3)
    big = 20000
    dets = np.column_stack((random_boxes(rng, big, 3000)[:, [1, 0, 3, 2]], rng.permutation(big).astype(np.float32) / big))
    for c in (cfg, dict(method="linear", iou_thresh=None, score_thresh=0.5, sigma=None)):
        np.testing.assert_array_equal(u.nms_np.nms(dets.copy(), c), nms_np_ref.nms(dets.copy(), c))
    rng = np.random.default_rng(2)
    n = 5000
    dets = np.column_stack((random_boxes(rng, n, 400)[:, [1, 0, 3, 2]], rng.permutation(n).astype(np.float32) / n))
    for c in (cfg, dict(method="diou", iou_thresh=None, score_thresh=None, sigma=None),
              dict(method="linear", iou_thresh=None, score_thresh=0.3, sigma=None)):
        np.testing.assert_array_equal(u.nms_np.nms(dets.copy(), c), nms_np_ref.nms(dets.copy(), c))
    dummy = u.nms_np.per_class_nms(np.zeros((0, 4), np.float32), np.zeros(0, np.float32), np.zeros(0, np.int32),
                                   np.float32([9.0]), np.float32([1.0]), 3, 10, cfg)
    np.testing.assert_array_equal(dummy, nms_np_ref.per_class_nms(np.zeros((0, 4), np.float32), np.zeros(0, np.float32),
                                  np.zeros(0, np.int32), np.float32([9.0]), np.float32([1.0]), 3, 10, cfg))
