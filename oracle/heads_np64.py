"""Independent float64 twin of ``oracle/heads_ref.py`` (TEST INFRASTRUCTURE ONLY).

``heads_ref.head_forward`` restates the reference's ClassNet / BoxNet towers with torch-CPU fp32 ``conv2d`` - a
library kernel the reference does not pin (tensorflow==2.10.0 is not installable offline => PARITY UNPINNED).  This
file states the same arithmetic a second time with nothing but NumPy float64 loops over the 9 taps and a matrix
product, written from the layer definitions rather than from the torch code, so that the two restatements check each
other (tests/test_oracle_extra.py):

  efficientdet_keras.py:421-430, 444-446   SeparableConv2D: depthwise 3x3, padding "same" (zero), no bias, depth
                                           multiplier 1 -> pointwise 1x1 + bias; weights shared over the levels
  efficientdet_keras.py:448-483, 628-664   _conv_bn_act loop: conv -> BN(level) -> swish -> SpatialDropout2D(training=True)
  utils_keras.py:71-80                     BatchNormalization inference: (x - mean) * gamma / sqrt(var + 1e-3) + beta
  utils.py:42-59                           swish(x) = x * sigmoid(x)
  tf.keras SpatialDropout2D                noise shape [B,1,1,F]; y = x * (1 / (1 - rate)) * keep

Keras SeparableConv2D "same" padding with stride 1 and a 3x3 kernel pads one zero pixel on every side; the depthwise
kernel [3,3,F,1] is applied as a cross-correlation (no flip): out[y,x,c] = sum_{i,j} in[y+i-1, x+j-1, c] * k[i,j,c].
"""
import numpy as np

BN_EPS = 1e-3


def depthwise3x3_same(x, k):
    """x [B,H,W,F] float64, k [3,3,F] -> [B,H,W,F]"""
    b, h, w, f = x.shape
    pad = np.zeros((b, h + 2, w + 2, f), np.float64)
    pad[:, 1:-1, 1:-1] = x
    out = np.zeros_like(x)
    for i in range(3):
        for j in range(3):
            out += pad[:, i:i + h, j:j + w] * k[i, j][None, None, None, :]
    return out


def head_forward(feat, w, level, rate, keep):
    """One pass of one head on one level in float64.  feat [B,H,W,F] -> [B,H,W,Cout]; keep [R,B,F] or None."""
    x = np.asarray(feat, np.float64)
    for i in range(len(w["dw"])):
        x = depthwise3x3_same(x, np.asarray(w["dw"][i], np.float64))
        x = x @ np.asarray(w["pw"][i], np.float64) + np.asarray(w["b"][i], np.float64)
        bn = w["bn"][i][level]
        x = (x - np.float64(bn["mean"])) * (np.float64(bn["gamma"]) / np.sqrt(np.float64(bn["var"]) + BN_EPS)) + np.float64(bn["beta"])
        x = x / (1.0 + np.exp(-x))
        if rate:
            x = x * (1.0 / (1.0 - rate)) * np.asarray(keep[i], np.float64)[:, None, None, :]
    y = depthwise3x3_same(x, np.asarray(w["dwp"], np.float64))
    return y @ np.asarray(w["pwp"], np.float64) + np.asarray(w["bp"], np.float64)


def heads_sample(feats, weights, masks, rate_class, rate_box, num_samples):
    """float64 twin of heads_ref.heads_sample: (list[L] of [T,B,H,W,A*C], list[L] of [T,B,H,W,8A])"""
    cls_out, box_out = [], []
    for lvl, feat in enumerate(feats):
        cs, bs = [], []
        for t in range(num_samples):
            kc = masks[t, 0, lvl] if rate_class else None
            kb = masks[t, 1, lvl] if rate_box else None
            cs.append(head_forward(feat, weights["class"], lvl, rate_class, kc))
            bs.append(head_forward(feat, weights["box"], lvl, rate_box, kb))
        cls_out.append(np.stack(cs, 0))
        box_out.append(np.stack(bs, 0))
    return cls_out, box_out
