"""Mirror of ``src/utils_box.py`` ``decode_uncert`` (utils_box.py:105-276) on the device.

The stand-alone entry point wraps the same fused kernel the post-processing uses
(csrc/decode_moments.cu) with a one-level, one-anchor-per-location geometry, so the numbers are
the ones the pipeline produces.
"""
import numpy as np

from . import _lib
from . import engine as _engine

_GENERIC = dict(
    image_size=(1, 1), min_level=0, max_level=0, num_scales=1, aspect_ratios=[1.0], anchor_scale=1.0,
    num_classes=1, mc_dropout=False, mc_dropoutrate=0.0, mc_classheadrate=0.0, mc_boxheadrate=0.0,
    mc_dropoutsamp=1, fpn_num_filters=4, box_class_repeats=1,
    nms_configs=dict(method="hard", iou_thresh=None, score_thresh=None, sigma=None,
                     max_nms_inputs=0, max_output_size=1),
)


def _decode(pred_boxes, box_uncert, anchor_boxes, method, device_id=0):
    pred = np.asarray(pred_boxes, np.float32)
    lead = pred.shape[:-1]
    n = int(np.prod(lead, dtype=np.int64)) if lead else 1
    if pred.shape[-1] != 4:
        raise ValueError("pred_boxes must have a trailing dimension of 4")
    anc = np.broadcast_to(np.asarray(anchor_boxes, np.float32), pred.shape).reshape(n, 4)
    la = box_uncert is not None
    params = dict(_GENERIC, loss_attenuation=la, uncert_adjust_method=method if la else "l-norm")
    eng = _engine.Engine(params, device_id, level_hw=[(1, n)], anchors=anc)
    try:
        t = pred.reshape(1, 1, n, 4)
        box = np.concatenate([t, np.asarray(box_uncert, np.float32).reshape(1, 1, n, 4)], -1) if la else t
        out = eng.decode_moments([eng.ctx.zeros((1, 1, n, 1))], [eng.ctx.to_device(box)], 1,
                                 want=("boxes", "albox"))
        coords = out["boxes"].numpy().reshape(pred.shape)
        if la:
            return coords, out["albox"].numpy().reshape(pred.shape)
        return coords
    finally:
        eng.ctx.close()


def decode_uncert(pred_boxes, box_uncert, anchor_boxes, method="l-norm", n_samples=30, device_id=0):
    """utils_box.py:105-276.  Methods 'l-norm', 'n-flow', 'falsedec' (fp64 on the device, rounded
    to fp32 like the reference); 'sample' draws from tfp and is not offered (ValueError).
    Host arrays in, host arrays out (device tensors go through postprocess.*, which fuses this)."""
    if method == "sample":
        raise ValueError("decode method 'sample' (tfp MultivariateNormalDiag sampling, "
                         "utils_box.py:162-184) is not offered on the device")
    if method not in _lib.DECODE_METHODS:
        raise ValueError("unknown decode method {}".format(method))
    return _decode(pred_boxes, box_uncert, anchor_boxes, method, device_id)
