"""serving-driver shim and weight import.

``FeatureServingDriver.serve`` returns what the reference's ``ServingDriver.serve`` returns
(src/infer_lib.py:118-296: ``postprocess_global``'s tuple - boxes|albox|mcbox [B,100,12], scores, class|mcclass-std,
valid_len, logits), but computes it with the head sampler from BiFPN feature maps; the producer of those maps
(backbone + BiFPN of the reference model, run ONCE instead of T times - see INTEGRATION.md 2) is a callable.

``weights_from_variables`` maps the checkpoint variable names of the reference's ClassNet / BoxNet
(src/efficientdet_keras.py:371-445, 533-625: ``class_net/class-{i}/...``, ``class-{i}-bn-{level}``,
``class-predict``; ``box_net/box-*``) to the weight dict of ``HeadSampler`` / ``udal_set_head_weights``.
"""
import numpy as np

from . import device, heads


def _find(variables, *suffixes):
    """Value of the first variable whose name ends with one of the suffixes (TF appends ':0' and optimiser /
    EMA slots may prefix the scope)."""
    for name, val in variables.items():
        base = name.split(":")[0]
        for s in suffixes:
            if base == s or base.endswith("/" + s):
                return np.asarray(val)
    raise KeyError("checkpoint has no variable ending with %s" % (suffixes,))


def _tower(variables, net, prefix, repeats, min_level, max_level, filters):
    t = dict(dw=[], pw=[], b=[], bn=[])
    for i in range(repeats):
        scope = "%s/%s-%d" % (net, prefix, i)
        dw = _find(variables, scope + "/depthwise_kernel")          # [3,3,F,1]
        pw = _find(variables, scope + "/pointwise_kernel")          # [1,1,F,F]
        t["dw"].append(dw.reshape(3, 3, filters).astype(np.float32))
        t["pw"].append(pw.reshape(filters, filters).astype(np.float32))
        t["b"].append(_find(variables, scope + "/bias").reshape(filters).astype(np.float32))
        levels = []
        for level in range(min_level, max_level + 1):
            bn = "%s/%s-%d-bn-%d" % (net, prefix, i, level)
            levels.append(dict(gamma=_find(variables, bn + "/gamma").astype(np.float32),
                               beta=_find(variables, bn + "/beta").astype(np.float32),
                               mean=_find(variables, bn + "/moving_mean").astype(np.float32),
                               var=_find(variables, bn + "/moving_variance").astype(np.float32)))
        t["bn"].append(levels)
    scope = "%s/%s-predict" % (net, prefix)
    dwp = _find(variables, scope + "/depthwise_kernel")
    pwp = _find(variables, scope + "/pointwise_kernel")
    t["dwp"] = dwp.reshape(3, 3, filters).astype(np.float32)
    t["pwp"] = pwp.reshape(filters, -1).astype(np.float32)
    t["bp"] = _find(variables, scope + "/bias").reshape(-1).astype(np.float32)
    return t


def weights_from_variables(variables, params):
    """variables: mapping name -> array (e.g. ``{v.name: v.numpy() for v in model.variables}`` or a
    ``tf.train.load_checkpoint`` reader dumped to a dict).  Returns {"class": tower, "box": tower}."""
    f = params.get("fpn_num_filters", 64)
    r = params.get("box_class_repeats", 3)
    lo, hi = params["min_level"], params["max_level"]
    w = {"class": _tower(variables, "class_net", "class", r, lo, hi, f),
         "box": _tower(variables, "box_net", "box", r, lo, hi, f)}
    a = params["num_scales"] * len(params["aspect_ratios"])
    want_cls = a * params["num_classes"]
    want_box = 4 * a * (2 if params.get("loss_attenuation") else 1)
    if w["class"]["pwp"].shape[1] != want_cls or w["box"]["pwp"].shape[1] != want_box:
        raise ValueError("predict layers have %d / %d channels, the configuration needs %d / %d"
                         % (w["class"]["pwp"].shape[1], w["box"]["pwp"].shape[1], want_cls, want_box))
    return w


class FeatureServingDriver:
    """``serve(image_arrays)`` like infer_lib.ServingDriver: ``features_fn(image_arrays)`` must return
    ``(feats, image_scales)`` - the per-image scale factors of ``DetectionInputProcessor`` (infer_lib.py:238-254) and
    either the 5 BiFPN maps [B,H_l,W_l,F] or, when ``bifpn_weights`` is given (``bifpn.weights_from_variables``), the
    backbone-level maps the first FPN cell reads (levels 3..5 + the resample_p6 / p7 outputs): ``FPNCells`` then runs on the
    device and its outputs feed the head sampler without leaving the GPU (efficientdet_keras.py:979-1050 from ``fpn_cells`` on).
    Host or device arrays."""

    def __init__(self, params, weights, features_fn, device_id=None, heads_mode=None, bifpn_weights=None):
        self.params = params
        self.features_fn = features_fn
        self.sampler = heads.HeadSampler(params, weights, device_id, heads_mode)
        self.fpn = None
        if bifpn_weights is not None:
            from . import bifpn
            self.fpn = bifpn.FPNCells(params, bifpn_weights, device_id)
        self._seed = 0

    def serve(self, image_arrays):
        feats, scales = self.features_fn(image_arrays)
        if self.fpn is not None:
            feats = self.fpn([x if device.is_device_array(x) else self.fpn.ctx.to_device(np.ascontiguousarray(x, np.float32))
                              for x in feats])
        self._seed += 1
        return self.sampler.detect(feats, scales, seed=self._seed)
