"""Drop-in mirror of the reference's ``src/nms_np.py`` backed by device kernels (csrc/nms_np.cu).

nms_np.py:30-89 diou_nms | :92-129 hard_nms | :132-194 soft_nms | :197-220 nms | :223-278
per_class_nms.  Same signatures and return values (NumPy in / NumPy out, like the reference, which
calls these through ``tf.numpy_function``); the arithmetic - "+1" pixel convention, fp32, visiting
order - runs on the GPU.  Any number of boxes per ``nms`` call (up to 16384 sort in shared memory, more in global scratch; the reference feeds <= 5000).

Tie order: ``argsort()[::-1]`` of the reference uses NumPy's unstable default sort; the device
kernel orders equal scores by descending index (what a stable sort would give).
"""
import ctypes

import numpy as np

from . import _lib
from . import engine as _engine

MAX_DETECTION_POINTS = 5000
_DUMMY_DETECTION_SCORE = -1e5
_METHODS = {"hard": 0, "diou": 1, "linear": 2, "gaussian": 3, "soft-hard": 4}


def _ctx():
    from . import postprocess
    return postprocess._any_engine()


def _run(dets, method, iou_thresh, sigma, score_thresh):
    dets = np.ascontiguousarray(dets, dtype=np.float32)
    n = dets.shape[0]
    if dets.ndim != 2 or dets.shape[1] != 5:
        raise ValueError("dets must have shape (num, 5)")
    eng = _ctx()
    kept = np.empty((max(n, 1), 5), np.float32)
    cnt = ctypes.c_int32(0)
    _lib.check(eng.lib.udal_nms_np(eng.ctx.handle, dets.ctypes.data, n, _METHODS[method],
                                   ctypes.c_float(iou_thresh), ctypes.c_float(sigma),
                                   ctypes.c_float(score_thresh), kept.ctypes.data, ctypes.byref(cnt)))
    return kept[:cnt.value]


def hard_nms(dets, iou_thresh=None):
    """nms_np.py:92-129."""
    return _run(dets, "hard", iou_thresh or 0.5, 0.0, 0.0)


def diou_nms(dets, iou_thresh=None):
    """nms_np.py:30-89."""
    return _run(dets, "diou", iou_thresh or 0.5, 0.0, 0.0)


def soft_nms(dets, nms_configs):
    """nms_np.py:132-194."""
    method = nms_configs["method"]
    sigma = nms_configs["sigma"] or 0.5
    iou_thresh = nms_configs["iou_thresh"] or 0.3
    score_thresh = nms_configs["score_thresh"] or 0.001
    kind = method if method in ("linear", "gaussian") else "soft-hard"
    out = _run(dets, kind, iou_thresh, sigma, score_thresh)
    if out.shape[0] == 0:
        return np.vstack([])  # the reference raises on an empty stack as well
    return out


def nms(dets, nms_configs):
    """nms_np.py:197-220."""
    nms_configs = nms_configs or {}
    method = nms_configs["method"]
    if method == "hard" or not method:
        return hard_nms(dets, nms_configs["iou_thresh"])
    if method == "diou":
        return diou_nms(dets, nms_configs["iou_thresh"])
    if method in ("linear", "gaussian"):
        return soft_nms(dets, nms_configs)
    raise ValueError("Unknown NMS method: {}".format(method))


def per_class_nms(boxes, scores, classes, image_id, image_scale, num_classes, max_boxes_to_draw,
                  nms_configs):
    """nms_np.py:223-278 -> [max_boxes_to_draw, 7] float32 rows [id, x1, y1, x2, y2, score, class+1]."""
    boxes = np.asarray(boxes)[:, [1, 0, 3, 2]]
    scores = np.asarray(scores)
    classes = np.asarray(classes)
    detections = []
    for c in range(num_classes):
        indices = np.where(classes == c)[0]
        if indices.shape[0] == 0:
            continue
        top = nms(np.column_stack((boxes[indices, :], scores[indices])), nms_configs)
        detections.append(np.column_stack(
            (np.repeat(image_id, len(top)), top, np.repeat(c + 1, len(top)))))

    def dummy(number):
        d = np.zeros((number, 7), dtype=np.float32)
        d[:, 0] = image_id[0]
        d[:, 5] = _DUMMY_DETECTION_SCORE
        return d

    if detections:
        detections = np.vstack(detections)
        indices = np.argsort(-detections[:, -2])
        detections = np.array(detections[indices[0:max_boxes_to_draw]], dtype=np.float32)
        detections = np.vstack([detections, dummy(max(max_boxes_to_draw - len(detections), 0))])
    else:
        detections = dummy(max_boxes_to_draw)
    detections[:, 1:5] *= image_scale
    return detections
