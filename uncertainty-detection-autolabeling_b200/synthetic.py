"""Synthetic inputs of the named shapes: random-init head weights with the reference's initialisers,
BiFPN feature maps and SpatialDropout2D keep masks.  Used by ``bench.py``, ``tools/`` and the examples - there is no
network for checkpoints or datasets.  Pure NumPy; ``oracle/heads_ref.py`` (test infrastructure) carries its own copy of
these generators, ``tests/test_abi_cpu.py`` checks that the two stay identical.

Weight dict layout = the argument of ``Engine.set_head_weights`` / ``HeadSampler``:
{"class": tower, "box": tower}, tower = dict(dw[R][3,3,F], pw[R][F,F], b[R][F], bn[R][L]{gamma,beta,mean,var},
dwp[3,3,F], pwp[F,Cout], bp[Cout]); keep masks [T, 2 (0 = class, 1 = box), L, R, B, F] uint8.
"""
import math

import numpy as np

HEAD_CLASS, HEAD_BOX = 0, 1


def _trunc_normal(rng, shape, std):
    # tf.initializers.variance_scaling(): truncated normal, stddev = sqrt(scale/fan_in)/.8796
    out = rng.standard_normal(shape)
    bad = np.abs(out) > 2
    while bad.any():
        out[bad] = rng.standard_normal(int(bad.sum()))
        bad = np.abs(out) > 2
    return (out * (std / 0.87962566103423978)).astype(np.float32)


def init_head_weights(num_filters, repeats, num_levels, num_anchors, num_classes, loss_attenuation,
                      seed=2024, randomize_bn=False):
    """Random weights with the reference's initialisers (efficientdet_keras.py:493-494, 510,
    587-588).  ``randomize_bn`` draws non-trivial BN statistics so per-level BN is exercised."""
    rng = np.random.default_rng(seed)
    f = num_filters

    def tower(cout, bias_value):
        w = {
            "dw": [_trunc_normal(rng, (3, 3, f), math.sqrt(1.0 / 9.0)) for _ in range(repeats)],
            "pw": [_trunc_normal(rng, (f, f), math.sqrt(1.0 / f)) for _ in range(repeats)],
            "b": [np.zeros(f, np.float32) for _ in range(repeats)],
            "bn": [],
            "dwp": _trunc_normal(rng, (3, 3, f), math.sqrt(1.0 / 9.0)),
            "pwp": _trunc_normal(rng, (f, cout), math.sqrt(1.0 / f)),
            "bp": np.full(cout, bias_value, np.float32),
        }
        for _ in range(repeats):
            per_level = []
            for _ in range(num_levels):
                if randomize_bn:
                    per_level.append({
                        "gamma": rng.uniform(0.5, 1.5, f).astype(np.float32),
                        "beta": rng.normal(0, 0.2, f).astype(np.float32),
                        "mean": rng.normal(0, 0.2, f).astype(np.float32),
                        "var": rng.uniform(0.5, 1.5, f).astype(np.float32),
                    })
                else:
                    per_level.append({
                        "gamma": np.ones(f, np.float32), "beta": np.zeros(f, np.float32),
                        "mean": np.zeros(f, np.float32), "var": np.ones(f, np.float32),
                    })
            w["bn"].append(per_level)
        if randomize_bn:
            w["b"] = [rng.normal(0, 0.1, f).astype(np.float32) for _ in range(repeats)]
        return w

    box_out = 4 * num_anchors * (2 if loss_attenuation else 1)
    return {
        "class": tower(num_anchors * num_classes, -math.log((1 - 0.01) / 0.01)),
        "box": tower(box_out, 0.0),
    }


def make_masks(num_samples, num_levels, repeats, batch, num_filters, rate_class, rate_box, seed=7):
    rng = np.random.default_rng(seed)
    u = rng.random((num_samples, 2, num_levels, repeats, batch, num_filters))
    keep = np.empty(u.shape, np.uint8)
    keep[:, HEAD_CLASS] = u[:, HEAD_CLASS] >= rate_class
    keep[:, HEAD_BOX] = u[:, HEAD_BOX] >= rate_box
    return keep


def make_features(level_shapes, batch, num_filters, seed=1234):
    rng = np.random.default_rng(seed)
    return [rng.standard_normal((batch, h, w, num_filters)).astype(np.float32) for h, w in level_shapes]


# ---- BASELINE configs[4] (auto-label threshold pass) --------------------------------------------------------------
# With the reference initialiser (class-predict bias -log(99)) every score of a random-init head sits near 0.01, far
# below the synthetic min_score 0.4 of SURVEY 8d: no detection is ever examined and the decision rule
# (src/infer_model.py:742-764) says "auto-label" for every image.  The measured pass therefore shifts the class-predict
# bias so that the median image's best score lands at min_score, and gives the images different feature amplitudes:
# roughly half of them then carry detections above min_score (high entropy -> "examine"), the rest do not.
AUTOLABEL_CLASS_BIAS = -1.3   # measured on B200: best score per image 0.26 .. 0.46 (median 0.35) at -1.5


def autolabel_variant(weights, class_bias=None):
    """in place: class-predict bias of the synthetic head for the auto-label pass"""
    weights["class"]["bp"][...] = AUTOLABEL_CLASS_BIAS if class_bias is None else class_bias
    return weights


def autolabel_amplitudes(batch):
    """per-image amplitude of the synthetic BiFPN features of the auto-label pass"""
    return np.linspace(0.6, 1.4, batch).astype(np.float32)


# ---- BiFPN --------------------------------------------------------------------------------------------
def init_bifpn_weights(num_filters, cell_repeats, in_channels, nodes, weight_method="fastattn", seed=77, randomize=True):
    """Random FPNCells weights in the layout of ``bifpn.FPNCells``.  ``in_channels``: channel count of every input level of the
    FIRST cell (e.g. EfficientNet-B0: [40, 112, 320, F, F]); ``nodes``: fpn_configs.bifpn_config(...)["nodes"].  The
    reference initialises the edge weights with ones and BN with identity statistics; ``randomize`` draws non-trivial
    values so that every term is exercised."""
    rng = np.random.default_rng(seed)
    f = num_filters
    per_channel = weight_method.startswith("channel_")

    def bn():
        if not randomize:
            return {"gamma": np.ones(f, np.float32), "beta": np.zeros(f, np.float32), "mean": np.zeros(f, np.float32),
                    "var": np.ones(f, np.float32)}
        return {"gamma": rng.uniform(0.5, 1.5, f).astype(np.float32), "beta": rng.normal(0, 0.2, f).astype(np.float32),
                "mean": rng.normal(0, 0.2, f).astype(np.float32), "var": rng.uniform(0.5, 1.5, f).astype(np.float32)}

    cells = []
    for c in range(cell_repeats):
        chans = list(in_channels) if c == 0 else [f] * len(in_channels)
        fnodes = []
        for node in nodes:
            res, wsm = [], []
            for off in node["inputs_offsets"]:
                cin = chans[off] if off < len(chans) else f
                if cin != f:
                    res.append({"w": _trunc_normal(rng, (cin, f), math.sqrt(1.0 / cin)),
                                "b": (rng.normal(0, 0.1, f) if randomize else np.zeros(f)).astype(np.float32), "bn": bn()})
                else:
                    res.append(None)
                shape = (f,) if per_channel else ()
                wsm.append((rng.uniform(-0.3, 1.5, shape) if randomize else np.ones(shape)).astype(np.float32))
            fnodes.append({"resample": res, "wsm": None if weight_method == "sum" else wsm,
                           "dw": _trunc_normal(rng, (3, 3, f), math.sqrt(1.0 / 9.0)), "pw": _trunc_normal(rng, (f, f), math.sqrt(1.0 / f)),
                           "b": (rng.normal(0, 0.1, f) if randomize else np.zeros(f)).astype(np.float32), "bn": bn()})
        cells.append({"fnodes": fnodes})
    return {"cells": cells}
