"""calibrated-uncertainty application + auto-label threshold pass.
CPU: the oracle against golden vectors produced by the reference's own calibrate_boxuncert /
relativize_uncert (tests/golden/make_golden_autolabel.py).  GPU: udal_autolabel against the oracle."""
import os

import numpy as np
import pytest

from oracle import autolabel_ref as ar

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "autolabel.npz"))
C = 7


def _tables(prefix):
    off = G[prefix + "_off"]
    return [(G[prefix + "_tx"][off[i]:off[i + 1]], G[prefix + "_ty"][off[i]:off[i + 1]]) for i in range(len(off) - 1)]


METHODS = {
    "ts_all": dict(temps=[float(G["temp_all"])]),
    "ts_percoo": dict(temps=G["temps_percoo"]),
    "iso_all": dict(tables=_tables("abs")[:1]),
    "iso_percoo": dict(tables=_tables("abs")[:4]),
    "iso_perclscoo": dict(tables=_tables("abs")),
    "rel_iso_perclscoo": dict(tables=_tables("rel")),
}


@pytest.mark.parametrize("method", sorted(METHODS))
def test_oracle_calibration_matches_reference(method):
    got = ar.calibrate_boxuncert(method, G["albox"], G["classes"], G["boxes"], C, **METHODS[method])
    ref = G[method]
    assert got.shape == ref.shape
    assert got.dtype == ref.dtype
    np.testing.assert_array_equal(got, ref)


def test_oracle_entropy_relativize_decision_match_reference():
    np.testing.assert_array_equal(ar.entropy_of_logits(G["logits"]), G["entropy"])
    np.testing.assert_array_equal(ar.relativize_uncert(G["boxes"], G["albox"]), G["rel_plain"])
    w = G["opt_params"]
    thr = float(np.mean(G["opt_thrs"]))
    plain = ar.autolabel_image(G["boxes"], G["albox"], G["scores"], G["classes"], G["logits"], C, w[0], w[1], thr,
                               float(G["min_score"]), method=None)
    np.testing.assert_allclose(plain["opt_uncert"], G["opt_plain"], rtol=1e-6, equal_nan=True)
    assert plain["auto_label"] == bool(G["decision_plain"])
    strict = ar.autolabel_image(G["boxes"], G["albox"], G["scores"], G["classes"], G["logits"], C, w[0], w[1], thr,
                                float(G["min_score"]), method="iso_perclscoo", tables=_tables("abs"), strict_reference=True)
    np.testing.assert_allclose(strict["rel_albox"], G["rel_first_row"], rtol=1e-6)
    np.testing.assert_allclose(strict["opt_uncert"], G["opt_strict"], rtol=1e-6)
    assert strict["auto_label"] == bool(G["decision_strict"])


def _detections(batch, seed):
    """The golden image plus perturbed copies, in the postprocess_global output layout."""
    rng = np.random.default_rng(seed)
    M = G["boxes"].shape[0]
    boxes = np.zeros((batch, M, 12), np.float32)
    scores = np.zeros((batch, M), np.float32)
    classes = np.zeros((batch, M, 1 + C), np.float32)
    logits = np.zeros((batch, M, C), np.float32)
    for b in range(batch):
        jitter = 1.0 if b == 0 else rng.uniform(0.7, 1.3)
        boxes[b, :, 0:4] = G["boxes"]
        boxes[b, :, 4:8] = np.nan_to_num(G["albox"]) * jitter if b else G["albox"]
        boxes[b, :, 8:12] = rng.uniform(0, 2, (M, 4))
        scores[b] = G["scores"] if b == 0 else rng.permutation(G["scores"])
        classes[b, :, 0] = G["classes"] if b == 0 else rng.integers(1, C + 1, M)
        logits[b] = G["logits"] * jitter
    return boxes, scores, classes, np.full(batch, M, np.int32), logits


@pytest.mark.gpu
@pytest.mark.parametrize("method,strict", [(None, True), ("ts_all", True), ("ts_percoo", False), ("iso_all", True),
                                           ("iso_percoo", False), ("iso_perclscoo", True), ("iso_perclscoo", False),
                                           ("rel_iso_perclscoo", False)])
def test_device_autolabel_matches_oracle(method, strict):
    import udal_b200 as u
    batch = 5
    det = _detections(batch, 3)
    kw = METHODS.get(method, {})
    tables = [u.autolabel.IsotonicTable(x, y) for x, y in kw.get("tables", [])]
    params = dict(num_classes=C, thr_sel_uncert=["ENT", "ALBOX"], calib_method_box=method, min_score=float(G["min_score"]))
    labeler = u.autolabel.AutoLabeler(params, G["opt_params"], G["opt_thrs"], tables=tables, temps=kw.get("temps"),
                                      strict_reference=strict)
    got = labeler.decide(det)
    thr = float(np.mean(G["opt_thrs"]))
    for b in range(batch):
        ref = ar.autolabel_image(det[0][b, :, 0:4], det[0][b, :, 4:8], det[1][b], det[2][b, :, 0], det[4][b], C,
                                 G["opt_params"][0], G["opt_params"][1], thr, float(G["min_score"]), method=method,
                                 tables=kw.get("tables"), temps=kw.get("temps"), strict_reference=strict)
        np.testing.assert_allclose(got["entropy"][b], ref["entropy"], rtol=2e-5, atol=1e-6)
        np.testing.assert_allclose(got["calib_albox"][b], ref["calib_albox"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(got["rel_albox"][b], ref["rel_albox"], rtol=1e-5, atol=1e-7)
        np.testing.assert_allclose(got["opt_uncert"][b], ref["opt_uncert"], rtol=2e-5, atol=1e-6)
        margin = np.abs(ref["opt_uncert"][det[1][b] > float(G["min_score"])] - thr).min()
        if margin > 1e-4:  # not a knife-edge case
            assert bool(got["auto_label"][b]) == ref["auto_label"]
    # the golden image itself reproduces the reference's decisions
    if method is None:
        assert bool(got["auto_label"][0]) == bool(G["decision_plain"])
    if method == "iso_perclscoo" and strict:
        assert bool(got["auto_label"][0]) == bool(G["decision_strict"])


@pytest.mark.gpu
def test_autolabel_on_device_detections_and_errors():
    import udal_b200 as u
    det = _detections(3, 9)
    eng = u.postprocess._any_engine()
    dev = tuple(eng.ctx.to_device(x) for x in det)
    labeler = u.autolabel.AutoLabeler(dict(num_classes=C, thr_sel_uncert=["ENT"], calib_method_box=None, min_score=0.4),
                                      [1.0], [1.5])
    host = labeler.decide(det)
    on_dev = labeler.decide(dev)
    np.testing.assert_array_equal(on_dev["opt_uncert"].numpy(), host["opt_uncert"])
    np.testing.assert_array_equal(on_dev["opt_uncert"].numpy(), host["entropy"])      # weight 1 on ENT only
    with pytest.raises(ValueError):
        u.autolabel.AutoLabeler(dict(num_classes=C, calib_method_box="bogus"), [0.5, 0.5], [0.5])
    with pytest.raises(ValueError):
        u.autolabel.AutoLabeler(dict(num_classes=C, calib_method_box="iso_percoo"), [0.5, 0.5], [0.5], tables=[])


@pytest.mark.gpu
def test_autolabel_behind_a_sampler_of_another_context():
    """detections left on the device by one context's udal_run (tail on its post stream) feed the auto-label pass: the
    pass must see finished data (round 2 found it reading the buffers before the producer's tail had run) - decisions
    equal the pass on host copies, repeatedly, with runs queued back to back"""
    import udal_b200 as u
    from oracle import heads_ref
    C = 8
    p = u.hparams_config.get_detection_config("efficientdet-d0", image_size=(256, 384), num_classes=C, enable_softmax=True,
                                              loss_attenuation=True, mc_dropout=True, mc_classheadrate=0.05,
                                              mc_boxheadrate=0.05, mc_dropoutsamp=6, heads_mode="fp16")
    w = heads_ref.init_head_weights(64, 3, 5, 9, C, True)
    w["class"]["bp"][...] = -1.0
    sampler = u.heads.HeadSampler(p, w)
    eng = sampler.engine
    batch = 8
    labeler = u.autolabel.AutoLabeler(dict(num_classes=C, thr_sel_uncert=["ENT", "ALBOX"], calib_method_box=None, min_score=0.3),
                                      opt_params=[0.5, 0.5], opt_thrs=[0.5])
    other = u.engine.get_engine(p)          # a second context: device arrays of `eng` handed to it are ordered behind eng
    assert other.ctx is not eng.ctx
    for trial in range(4):
        feats = [eng.ctx.to_device(f * np.linspace(0.5, 1.5, batch, dtype=np.float32)[:, None, None, None])
                 for f in heads_ref.make_features(eng.level_hw, batch, 64, seed=trial)]
        det = eng.run(feats, None, None, seed=trial)
        tup = (det["boxes"], det["scores"], det["classes"], det["valid"], det["logits"])
        dev = labeler.decide(tup)                              # in the producer's context
        sig = other.ctx.empty(det["logits"].shape)
        u._lib.check(other.lib.udal_sigmoid(other.ctx.handle, u.device.as_device(other.ctx, det["logits"], np.float32)[0].ptr,
                                            det["logits"].size, sig.ptr))          # cross-context consumer
        host = labeler.decide(tuple(x.numpy() for x in tup))
        np.testing.assert_array_equal(dev["auto_label"].numpy(), host["auto_label"])
        np.testing.assert_allclose(dev["opt_uncert"].numpy(), host["opt_uncert"], rtol=1e-6)
        np.testing.assert_allclose(sig.numpy(), 1.0 / (1.0 + np.exp(-det["logits"].numpy().astype(np.float64))), rtol=1e-5)


# ---------------------------------------------------------------------------------------------
# CalibrateClass (utils_class.py:44-272): the four classification calibrators
# ---------------------------------------------------------------------------------------------
def _classcal_golden():
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "classcalib.npz"))


def _classcal_tables(g, first, count):
    off = g["off"]
    return [(g["tx"][off[t]:off[t + 1]], g["ty"][off[t]:off[t + 1]]) for t in range(first, first + count)]


@pytest.mark.parametrize("method", ["ts_all", "ts_percls", "iso_all", "iso_percls"])
def test_oracle_class_calibration_matches_reference(method):
    """the oracle against outputs of the reference's own CalibrateClass._perform_class_calib (make_golden_classcalib.py)"""
    g = _classcal_golden()
    c = g["logits"].shape[1]
    temps = g["temp_all"] if method == "ts_all" else g["temps_pc"]
    tables = _classcal_tables(g, 0, 1) if method == "iso_all" else _classcal_tables(g, 1, c)
    ent, prob = ar.calibrate_class(g["logits"], method, temps=temps, tables=tables)
    np.testing.assert_allclose(prob, g["probab_" + method], rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(ent, g["entropy_" + method], rtol=1e-6, atol=1e-7)


@pytest.mark.gpu
def test_device_class_calibration_matches_reference():
    import udal_b200 as u
    g = _classcal_golden()
    c = g["logits"].shape[1]
    tabs = [u.autolabel.IsotonicTable(x, y) for x, y in _classcal_tables(g, 0, 1 + c)]
    cals = {"ts_all": g["temp_all"], "ts_percls": g["temps_pc"], "iso_all": tabs[0], "iso_percls": tabs[1:]}
    cc = u.utils_class.CalibrateClass(g["logits"], cals, "iso_percls")
    for method in u.utils_class.AVAILABLE_CALIB:
        ent, prob = cc._perform_class_calib(method)
        np.testing.assert_allclose(prob, g["probab_" + method], rtol=2e-5, atol=1e-7)
        np.testing.assert_allclose(ent, g["entropy_" + method], rtol=2e-5, atol=2e-6)
    out = cc.calibrate_class()
    assert len(out) == 9 and out[0].size == 0          # the reference's selection quirk (strict)
    np.testing.assert_allclose(out[7], g["probab_iso_percls"], rtol=2e-5, atol=1e-7)
    loose = u.utils_class.CalibrateClass(g["logits"], cals, "iso_percls", strict_reference=False).calibrate_class()
    np.testing.assert_allclose(loose[0], g["entropy_iso_percls"], rtol=2e-5, atol=2e-6)
    # device logits in -> device arrays out; a missing calibrator -> empty arrays
    eng = u.postprocess._any_engine()
    dev = u.utils_class.CalibrateClass(eng.ctx.to_device(g["logits"]), {"ts_all": g["temp_all"]}, "ts_all").calibrate_class()
    np.testing.assert_allclose(dev[1].numpy(), g["probab_ts_all"], rtol=2e-5, atol=1e-7)
    assert dev[3].size == 0 and dev[5].size == 0
    with pytest.raises(ValueError):
        cc._perform_class_calib("bogus")
    with pytest.raises(NotImplementedError):
        u.utils_class.CalibrateClass(g["logits"], cals, "ts_all", uncert=[np.ones_like(g["logits"])])
