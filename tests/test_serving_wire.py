"""weight import from checkpoint variable names, serving shim, wire formats."""
import os

import numpy as np
import pytest

from oracle import heads_ref


def _variables(w, min_level=3):
    """Names and shapes as the reference's Keras layers create them (efficientdet_keras.py:421-446, 583-626)."""
    v = {}
    for net, prefix, t in (("class_net", "class", w["class"]), ("box_net", "box", w["box"])):
        for i in range(len(t["dw"])):
            f = t["dw"][i].shape[-1]
            s = "efficientdet-d0/%s/%s-%d/" % (net, prefix, i)
            v[s + "depthwise_kernel:0"] = np.asarray(t["dw"][i]).reshape(3, 3, f, 1)
            v[s + "pointwise_kernel:0"] = np.asarray(t["pw"][i]).reshape(1, 1, f, f)
            v[s + "bias:0"] = np.asarray(t["b"][i])
            for l, bn in enumerate(t["bn"][i]):
                b = "efficientdet-d0/%s/%s-%d-bn-%d/" % (net, prefix, i, min_level + l)
                v[b + "gamma:0"], v[b + "beta:0"] = bn["gamma"], bn["beta"]
                v[b + "moving_mean:0"], v[b + "moving_variance:0"] = bn["mean"], bn["var"]
        s = "efficientdet-d0/%s/%s-predict/" % (net, prefix)
        v[s + "depthwise_kernel:0"] = np.asarray(t["dwp"]).reshape(3, 3, -1, 1)
        v[s + "pointwise_kernel:0"] = np.asarray(t["pwp"]).reshape(1, 1, t["pwp"].shape[0], -1)
        v[s + "bias:0"] = np.asarray(t["bp"])
    return v


def test_weights_from_variables_roundtrip():
    import udal_b200 as u
    p = u.hparams_config.get_detection_config("efficientdet-d0", image_size=64, num_classes=7, loss_attenuation=True)
    w = heads_ref.init_head_weights(64, 3, 5, 9, 7, True, seed=3, randomize_bn=True)
    got = u.serving.weights_from_variables(_variables(w), p)
    for head in ("class", "box"):
        for k in ("dwp", "pwp", "bp"):
            np.testing.assert_array_equal(got[head][k], np.asarray(w[head][k], np.float32).reshape(got[head][k].shape))
        for i in range(3):
            np.testing.assert_array_equal(got[head]["dw"][i], np.asarray(w[head]["dw"][i]).reshape(3, 3, 64))
            np.testing.assert_array_equal(got[head]["pw"][i], w[head]["pw"][i])
            for l in range(5):
                for k in ("gamma", "beta", "mean", "var"):
                    np.testing.assert_array_equal(got[head]["bn"][i][l][k], w[head]["bn"][i][l][k])
    bad = dict(p, num_classes=8)
    with pytest.raises(ValueError):
        u.serving.weights_from_variables(_variables(w), bad)
    with pytest.raises(KeyError):
        u.serving.weights_from_variables({}, p)


def test_prediction_data_wire_format(tmp_path):
    import udal_b200 as u
    rng = np.random.default_rng(0)
    m, c = 12, 7
    boxes = rng.uniform(0, 300, (m, 12)).astype(np.float32)
    boxes[2, 5] = np.inf
    scores = np.sort(rng.uniform(0, 1, m))[::-1].astype(np.float32)
    classes = np.concatenate([rng.integers(1, c + 1, (m, 1)), rng.uniform(0, 1, (m, c))], 1).astype(np.float32)
    logits = rng.normal(-2, 2, (m, c)).astype(np.float32)
    recs = u.wire.prediction_records("img_000.jpg", (boxes, scores, classes, logits), min_score=0.4)
    assert len(recs) == int((scores > 0.4).sum())
    keys = list(recs[0].keys())
    assert keys == ["image_name", "score_thresh", "top_5scores", "det_score", "bbox", "class", "logits", "entropy",
                    "probab", "uncalib_mcclass", "uncalib_albox", "uncalib_mcbox"]
    path = os.path.join(tmp_path, "prediction_data.txt")
    u.wire.write_prediction_data(path, recs)
    back = u.wire.read_prediction_data(path)          # the consumers' ast.literal_eval reader
    assert len(back) == len(recs)
    for a, b in zip(recs, back):
        assert b["image_name"] == "img_000.jpg" and len(b["bbox"]) == 4 and len(b["uncalib_albox"]) == 4
        np.testing.assert_allclose(b["det_score"], a["det_score"], rtol=1e-6)
        np.testing.assert_allclose(b["logits"], np.around(a["logits"], 4), atol=1e-6)
        assert abs(sum(b["probab"]) - 1.0) < 1e-5
    # rounding / nan_to_num of add_array_dict (utils_extra.py:67-81)
    sel = int(np.where(scores > 0.4)[0][min(2, len(recs) - 1)])
    if sel == 2:
        assert back[2]["uncalib_albox"][1] == pytest.approx(3.4028235e38)


@pytest.mark.gpu
def test_feature_serving_driver_matches_detect():
    import udal_b200 as u
    p = u.hparams_config.get_detection_config(
        "efficientdet-d0", image_size=(64, 96), num_classes=7, enable_softmax=True, loss_attenuation=True,
        mc_dropout=True, mc_classheadrate=0.05, mc_boxheadrate=0.05, mc_dropoutsamp=3)
    eng = u.engine.get_engine(p)
    w = heads_ref.init_head_weights(eng.F, eng.R, len(eng.level_hw), eng.A, 7, True, seed=3, randomize_bn=True)
    feats = heads_ref.make_features(eng.level_hw, 2, eng.F, seed=2)
    scales = np.float32([1.0, 2.0])
    drv = u.serving.FeatureServingDriver(p, u.serving.weights_from_variables(_variables(w), p),
                                         lambda images: (feats, scales))
    det = drv.serve([None, None])
    ref = u.heads.HeadSampler(p, w).detect(feats, scales, seed=1)
    assert len(det) == 5 and det[0].shape == (2, 100, 12) and det[2].shape == (2, 100, 8)
    for a, b in zip(det, ref):
        np.testing.assert_array_equal(a, b)


@pytest.mark.gpu
def test_feature_serving_driver_with_bifpn_on_the_device():
    """backbone-level maps -> FPNCells -> head sampler inside the driver = the two stages chained by hand"""
    import udal_b200 as u
    size, cin, batch = (128, 192), [40, 112, 320, 64, 64], 2
    p = u.hparams_config.get_detection_config(
        "efficientdet-d0", image_size=size, num_classes=7, enable_softmax=True, loss_attenuation=True,
        mc_dropout=True, mc_classheadrate=0.05, mc_boxheadrate=0.05, mc_dropoutsamp=3)
    nodes = u.fpn_configs.bifpn_config(3, 7)["nodes"]
    wf = u.synthetic.init_bifpn_weights(64, 3, cin, nodes, seed=5)
    wh = heads_ref.init_head_weights(64, 3, 5, 9, 7, True, seed=3, randomize_bn=True)
    eng = u.engine.get_engine(p)
    sizes = [(16, 24), (8, 12), (4, 6), (2, 3), (1, 2)]
    assert [tuple(x) for x in eng.level_hw] == sizes
    rng = np.random.default_rng(2)
    feats = [rng.normal(size=(batch, h, w, c)).astype(np.float32) for (h, w), c in zip(sizes, cin)]
    scales = np.float32([1.0, 1.5])
    drv = u.serving.FeatureServingDriver(p, wh, lambda images: (feats, scales), bifpn_weights=wf)
    det = drv.serve([None] * batch)
    fpn = u.bifpn.FPNCells(p, wf)(feats)
    ref = u.heads.HeadSampler(p, wh).detect(fpn, scales, seed=1)
    for a, b in zip(det, ref):
        np.testing.assert_array_equal(a, b)
