"""CPU restatement of the EfficientDet class / box(+sigma) head towers and the MC-dropout loop
(TEST INFRASTRUCTURE ONLY).

Follows (paths relative to the reference's ``src/``):
  efficientdet_keras.py:353-513   ClassNet (shared separable convs, per-level BN, swish,
                                  SpatialDropout2D(training=True), class-predict)
  efficientdet_keras.py:516-692   BoxNet (same tower; box-predict has 4*num_anchors filters,
                                  num_anchors doubled under loss attenuation, :936-945)
  efficientdet_keras.py:979-1050  MC loop: T forward passes, outputs stacked on a new axis 0
  utils.py:42-59                  swish = x * sigmoid(x)
  utils_keras.py:42-82            BatchNormalization(momentum=.99, epsilon=1e-3), inference mode

The convolution / BN / dropout primitives are TensorFlow kernels (tensorflow==2.10.0, not
vendored, not installable offline) => PARITY UNPINNED; restated with torch-CPU fp32 conv2d.
Dropout is reproduced by injected keep masks (noise shape [B,1,1,F]).

Layouts: features [B,H,W,F] fp32 (NHWC); depthwise kernels [3,3,F]; pointwise [Fin,Fout];
keep masks [T, 2(head: 0=class, 1=box), L, R, B, F] uint8.
"""
import math

import numpy as np
import torch
import torch.nn.functional as tnf

BN_EPS = 1e-3  # utils_keras.py:78
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


def _sepconv(x, dw, pw, bias):
    """x [B,F,H,W]; depthwise 3x3 SAME (zero pad, no bias) then 1x1 + bias."""
    f = x.shape[1]
    k = torch.from_numpy(np.ascontiguousarray(dw.transpose(2, 0, 1)))[:, None]  # [F,1,3,3]
    x = tnf.conv2d(x, k, padding=1, groups=f)
    p = torch.from_numpy(np.ascontiguousarray(pw.T))[:, :, None, None]  # [Fout,Fin,1,1]
    return tnf.conv2d(x, p, bias=torch.from_numpy(bias))


def head_forward(feat, w, level, rate, keep):
    """One pass of one head on one level.  feat [B,H,W,F] -> [B,H,W,Cout].
    keep: [R,B,F] uint8 or None (efficientdet_keras.py:448-483 / 628-664)."""
    x = torch.from_numpy(np.ascontiguousarray(feat)).permute(0, 3, 1, 2).contiguous()
    repeats = len(w["dw"])
    for i in range(repeats):
        x = _sepconv(x, w["dw"][i], w["pw"][i], w["b"][i])
        bn = w["bn"][i][level]
        inv = torch.from_numpy(bn["gamma"]) / torch.sqrt(torch.from_numpy(bn["var"]) + BN_EPS)
        shift = torch.from_numpy(bn["beta"]) - torch.from_numpy(bn["mean"]) * inv
        x = x * inv[None, :, None, None] + shift[None, :, None, None]
        x = x * torch.sigmoid(x)
        if rate:
            k = torch.from_numpy(keep[i].astype(np.float32))  # [B,F]
            x = (x * np.float32(1.0 / (1.0 - rate))) * k[:, :, None, None]
    y = _sepconv(x, w["dwp"], w["pwp"], w["bp"])
    return y.permute(0, 2, 3, 1).contiguous().numpy()


def heads_sample(feats, weights, masks, rate_class, rate_box, num_samples):
    """MC loop (efficientdet_keras.py:999-1050) starting at the BiFPN outputs.

    feats: list[L] of [B,H,W,F]; masks: [T,2,L,R,B,F] uint8 (ignored for a head whose rate is 0).
    Returns (cls_outputs, box_outputs): list[L] of [T,B,H,W,A*C], list[L] of [T,B,H,W,8A]."""
    cls_out, box_out = [], []
    with torch.no_grad():
        for lvl, feat in enumerate(feats):
            cs, bs = [], []
            for t in range(num_samples):
                kc = masks[t, HEAD_CLASS, lvl] if rate_class else None
                kb = masks[t, HEAD_BOX, lvl] if rate_box else None
                cs.append(head_forward(feat, weights["class"], lvl, rate_class, kc))
                bs.append(head_forward(feat, weights["box"], lvl, rate_box, kb))
            cls_out.append(np.stack(cs, 0))
            box_out.append(np.stack(bs, 0))
    return cls_out, box_out


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
