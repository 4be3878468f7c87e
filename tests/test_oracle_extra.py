"""Cross-checks that tighten the parts of the oracle the reference cannot pin (VERDICT r1 "what's weak" #2).

* the torch-CPU heads restatement (oracle/heads_ref.py) against an independent float64 NumPy twin
  (oracle/heads_np64.py);
* a differential on the one arithmetic choice of the NMS-V5 restatement that a TF host could round differently:
  soft weight = fp32(exp(fp64(x))) (oracle and CUDA kernels) versus the C library's expf(x) (what TF's
  std::exp(float) executes) - how often the weight differs, and how often a keep-set changes because of it;
* the real reference, imported when TensorFlow is importable (skipped here: TF 2.10 is not installable offline).
"""
import ctypes
import os
import sys

import numpy as np
import pytest

from oracle import heads_np64, heads_ref, nms_ref, ref_np


@pytest.mark.parametrize("size,C,T,batch,la,rc,rb,bn", [
    ((32, 48), 7, 3, 2, True, 0.05, 0.05, True),
    ((24, 40), 10, 2, 1, False, 0.3, 0.0, True),
    ((16, 16), 3, 1, 2, True, 0.0, 0.0, False),
])
def test_heads_torch_oracle_vs_float64_twin(size, C, T, batch, la, rc, rb, bn):
    p = ref_np.default_params(image_size=size, num_classes=C, mc_dropoutsamp=T, loss_attenuation=la)
    levels = ref_np.level_shapes(p)
    A = ref_np.num_anchors_per_location(p)
    w = heads_ref.init_head_weights(64, 3, len(levels), A, C, la, seed=5, randomize_bn=bn)
    feats = heads_ref.make_features(levels, batch, 64, seed=3)
    masks = heads_ref.make_masks(T, len(levels), 3, batch, 64, rc, rb, seed=9)
    a_cls, a_box = heads_ref.heads_sample(feats, w, masks, rc, rb, T)
    b_cls, b_box = heads_np64.heads_sample(feats, w, masks, rc, rb, T)
    worst = 0.0
    for x, y in zip(a_cls + a_box, b_cls + b_box):
        assert x.shape == y.shape
        worst = max(worst, float(np.abs(x - y).max()))
        # fp32 conv (any summation order) against float64: a few ulp of the O(1..5) outputs
        np.testing.assert_allclose(x, y, rtol=2e-5, atol=2e-5)
    print("torch fp32 oracle vs float64 twin: max abs diff %.3g" % worst)


def _clustered_boxes(rng, n):
    """heavily overlapping random boxes (so that most candidates are decayed several times)"""
    centres = rng.uniform(50, 250, size=(max(n // 12, 1), 2))
    c = centres[rng.integers(0, len(centres), n)] + rng.normal(0, 6, size=(n, 2))
    hw = rng.uniform(20, 60, size=(n, 2))
    return np.concatenate([c - hw / 2, c + hw / 2], 1).astype(np.float32)


def test_soft_nms_expf_rounding_differential():
    """fp32(exp(fp64)) vs expf: the oracle header estimates ~0.4 % of the arguments round differently by 1 ulp.  Measured
    here on the box's own libm: the rate of differing weights and - what matters - how many of 10^4 random soft-NMS
    problems (n = 120 clustered boxes, 30 outputs) end with a different selection.  The unpinned risk gets a number."""
    lib = nms_ref.lib()
    mode = ctypes.c_int.in_dll(lib, "udal_oracle_weight_mode")
    rng = np.random.default_rng(42)
    # (a) weights
    iou = rng.uniform(0, 1, 200000).astype(np.float32)
    scale = np.float32(-0.5 / 0.25)
    try:
        mode.value = 0
        w0 = np.array([lib.udal_oracle_soft_weight(scale, float(u)) for u in iou[:50000]], np.float32)
        mode.value = 1
        w1 = np.array([lib.udal_oracle_soft_weight(scale, float(u)) for u in iou[:50000]], np.float32)
    finally:
        mode.value = 0
    differ = float((w0 != w1).mean())
    ulp = np.abs(w0.view(np.int32).astype(np.int64) - w1.view(np.int32).astype(np.int64)).max()
    assert ulp <= 1, "expf further than 1 ulp from the correctly rounded value"
    # (b) keep sets
    cases, flips, score_flips = 10000, 0, 0
    for i in range(cases):
        n = 120
        boxes = _clustered_boxes(rng, n)
        scores = rng.uniform(0.01, 1.0, n).astype(np.float32)
        try:
            mode.value = 0
            a = nms_ref.non_max_suppression_v5(boxes, scores, 30, 0.5, 0.001, 0.25, True)
            mode.value = 1
            b = nms_ref.non_max_suppression_v5(boxes, scores, 30, 0.5, 0.001, 0.25, True)
        finally:
            mode.value = 0
        if not np.array_equal(a[0], b[0]) or a[2] != b[2]:
            flips += 1
        elif not np.array_equal(a[1], b[1]):
            score_flips += 1
    print("soft-NMS weight: expf != fp32(exp(fp64)) for %.4f %% of 50000 arguments (max 1 ulp); "
          "%d of %d problems select different indices, %d more differ only in a returned score (1 ulp)"
          % (100 * differ, flips, cases, score_flips))
    # glibc's expf is correctly rounded for all but a few 1e-4 of its arguments: selections must be (almost) unaffected
    assert differ < 0.01
    assert flips <= cases // 100


def test_real_reference_when_tensorflow_is_importable():
    """on a machine with TensorFlow (the pinned 2.10) and the reference checkout, the unmodified
    reference post-processing and TF's own NonMaxSuppressionV5 are compared with the oracle directly.  Skipped in this
    image (no TF wheel, no network); /root/reference is read only when it exists (never on the GPU box)."""
    tf = pytest.importorskip("tensorflow")
    ref_src = os.environ.get("UDAL_REFERENCE_SRC", "/root/reference/src")
    if not os.path.isdir(ref_src):
        pytest.skip("reference checkout not present")
    sys.path.insert(0, ref_src)
    try:
        import postprocess as ref_post  # noqa: the reference's own module
    finally:
        sys.path.remove(ref_src)
    rng = np.random.default_rng(0)
    # (1) the raw op, both modes, against nms_v5.c
    for sigma in (0.0, 0.25):
        boxes = _clustered_boxes(rng, 400)
        scores = rng.uniform(0.0, 1.0, 400).astype(np.float32)
        idx, sc, valid = tf.raw_ops.NonMaxSuppressionV5(
            boxes=boxes, scores=scores, max_output_size=100, iou_threshold=0.5, score_threshold=0.001,
            soft_nms_sigma=sigma, pad_to_max_output_size=True)
        o_idx, o_sc, o_valid = nms_ref.non_max_suppression_v5(boxes, scores, 100, 0.5, 0.001, sigma, True)
        assert int(valid) == o_valid
        np.testing.assert_array_equal(idx.numpy(), o_idx)
        np.testing.assert_array_equal(sc.numpy(), o_sc)
    # (2) postprocess_global end to end on random head outputs
    p = ref_np.default_params(image_size=(64, 96), num_classes=7, mc_dropoutsamp=4)
    levels = ref_np.level_shapes(p)
    cls = [rng.normal(-4.6, 2.0, (4, 2, h, w, 63)).astype(np.float32) for h, w in levels]
    box = [np.concatenate([rng.normal(0, 0.4, (4, 2, h, w, 36)), np.abs(rng.normal(0, 0.3, (4, 2, h, w, 36))) + 0.01], -1).astype(np.float32)
           for h, w in levels]
    scales = np.float32([1.0, 1.5])
    got = ref_post.postprocess_global(dict(p), [tf.constant(c) for c in cls], [tf.constant(b) for b in box], scales)
    want = ref_np.postprocess_global(dict(p), cls, box, scales)
    for a, b in zip(got, want):
        np.testing.assert_allclose(np.asarray(a), b, rtol=1e-6, atol=1e-6)


# ---------------------------------------------------------------------------------------------
# the "epoch" formulation of the lazy soft-NMS heap used by csrc/nms_cta.cu, transcribed to NumPy and checked against
# the heap of oracle/nms_v5.c (the CUDA kernel is checked against the same oracle in the -m gpu tests)
# ---------------------------------------------------------------------------------------------
def _epoch_nms(boxes, scores, max_out, iou_thr, thr, sigma, old):
    f32 = np.float32
    n = len(scores)
    soft = sigma > 0
    scale = f32(-0.5 / sigma) if soft else f32(0)

    def iou(a, b):
        a, b = np.ascontiguousarray(a, f32), np.ascontiguousarray(b, f32)
        return f32(nms_ref.lib().udal_oracle_iou(a.ctypes.data, b.ctypes.data))

    def step(s, u):
        w = f32(np.exp(np.float64(f32(f32(scale * u) * u))))
        if old:
            if not (u <= iou_thr):
                w = f32(0)
            hard = u >= iou_thr
        else:
            if not (soft or u <= iou_thr):
                w = f32(0)
            hard = (not soft) and u > iou_thr
        return f32(s * w), hard

    def chain(s, j, frm, to, sel):
        for q in range(frm, to - 1, -1):
            s, hard = step(s, iou(boxes[j], boxes[sel[q]]))
            if hard:
                return s, True
            if s <= thr:
                break
        return s, not (s > thr)

    cur = np.where(scores > thr, scores, -np.inf).astype(f32)
    beg = np.zeros(n, int)
    sel, out = [], []
    key = lambda s, i: (s, -i)
    for e in range(max_out):
        ev, best = {}, None
        for j in range(n):                       # round 1: empty or one-step chains
            if cur[j] == -np.inf or beg[j] < e - 1:
                continue
            v, dead = (cur[j], False) if beg[j] == e else chain(cur[j], j, e - 1, e - 1, sel)
            ev[j] = -np.inf if dead else v
            if not dead and (best is None or key(v, j) > best):
                best = key(v, j)
        m1 = best
        for j in range(n):                       # round 2: pending chains that reach the round-1 maximum
            if cur[j] == -np.inf or j in ev or (m1 is not None and key(cur[j], j) < m1):
                continue
            v, dead = chain(cur[j], j, e - 1, beg[j], sel)
            ev[j] = -np.inf if dead else v
            if not dead and (best is None or key(v, j) > best):
                best = key(v, j)
        if best is None:
            break
        win = -best[1]
        for j, v in ev.items():                  # commit
            if j == win:
                sel.append(j)
                out.append(v)
                cur[j] = -np.inf
            elif key(cur[j], j) > best:
                cur[j] = v
                beg[j] = e
    return sel, np.asarray(out, f32)


def test_epoch_formulation_of_the_lazy_heap_matches_nms_v5_c():
    rng = np.random.default_rng(0)
    for trial in range(25):
        n = int(rng.integers(5, 90))
        spread = rng.choice([0.0, 3.0, 20.0, 80.0])
        ctr = np.float32([200, 300]) + rng.normal(0, spread, (n, 2)).astype(np.float32)
        hw = rng.uniform(20, 90, (n, 2)).astype(np.float32) if rng.random() < 0.8 else np.full((n, 2), 70, np.float32)
        boxes = np.concatenate([ctr - hw / 2, ctr + hw / 2], -1).astype(np.float32)
        scores = rng.uniform(0.01, 1, n).astype(np.float32)
        if rng.random() < 0.5:
            scores[::3] = 0.5
        for sigma, old, iou_thr, thr in [(0.25, False, 0.5, 0.001), (0.25, True, 0.5, 0.001), (0.0, False, 0.5, -np.inf),
                                         (0.0, True, 0.3, 0.2), (0.1, False, 0.5, 0.2)]:
            mo = int(rng.integers(1, 30))
            ri, rs, _ = nms_ref.non_max_suppression_v5(boxes, scores, mo, iou_thr, thr, sigma, False, "old" if old else "new")
            si, so = _epoch_nms(boxes, scores, mo, np.float32(iou_thr), np.float32(thr), sigma, old)
            assert list(ri) == si, (trial, sigma, old)
            np.testing.assert_array_equal(so, rs)
