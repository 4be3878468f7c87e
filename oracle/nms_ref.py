"""TF ``NonMaxSuppressionV5`` restated (TEST INFRASTRUCTURE ONLY) - ctypes front end of
``oracle/nms_v5.c`` plus an independent pure-Python transcription used to cross-check the C
file on small inputs.

Reference call site: src/postprocess.py:392-400.  Third-party kernel (tensorflow==2.10.0,
``core/kernels/image/non_max_suppression_op.cc``), not vendored and not installable offline:
PARITY UNPINNED (see oracle/__init__.py).
"""
import ctypes
import math

import numpy as np

from . import build as _build

_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(_build.build())
        _lib.udal_oracle_nms_v5.restype = ctypes.c_int
        _lib.udal_oracle_nms_v5.argtypes = [
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_float,
            ctypes.c_float, ctypes.c_float, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
        ]
        _lib.udal_oracle_nms_v5_batch.restype = None
        _lib.udal_oracle_nms_v5_batch.argtypes = [
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
            ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_int, ctypes.c_void_p,
            ctypes.c_void_p, ctypes.c_void_p,
        ]
        _lib.udal_oracle_iou.restype = ctypes.c_float
        _lib.udal_oracle_iou.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        _lib.udal_oracle_soft_weight.restype = ctypes.c_float
        _lib.udal_oracle_soft_weight.argtypes = [ctypes.c_float, ctypes.c_float]
    return _lib


def _score_thr(x):
    x = float(x)
    if math.isinf(x) and x < 0:
        return -np.inf
    return x


def non_max_suppression_v5(boxes, scores, max_output_size, iou_threshold, score_threshold,
                           soft_nms_sigma, pad_to_max_output_size, variant="new"):
    """-> (selected_indices int32, selected_scores float32, valid_outputs int).

    Unpadded results are sliced to ``valid`` entries, padded ones are zero filled to
    ``max_output_size`` (TF behaviour)."""
    boxes = np.ascontiguousarray(boxes, dtype=np.float32)
    scores = np.ascontiguousarray(scores, dtype=np.float32)
    n = boxes.shape[0]
    assert boxes.shape == (n, 4) and scores.shape == (n,)
    max_out = int(max_output_size)
    idx = np.zeros(max(max_out, 1), np.int32)
    sc = np.zeros(max(max_out, 1), np.float32)
    valid = lib().udal_oracle_nms_v5(
        boxes.ctypes.data, scores.ctypes.data, n, max_out, float(iou_threshold),
        _score_thr(score_threshold), float(soft_nms_sigma), 1 if variant == "old" else 0,
        idx.ctypes.data, sc.ctypes.data,
    )
    idx, sc = idx[:max_out], sc[:max_out]
    if not pad_to_max_output_size:
        idx, sc = idx[:valid], sc[:valid]
    return idx, sc, int(valid)


def non_max_suppression_v5_batch(boxes, scores, max_output_size, iou_threshold, score_threshold,
                                 soft_nms_sigma, variant="new"):
    """Padded batch form used by the CPU baseline: boxes [B,n,4], scores [B,n]."""
    boxes = np.ascontiguousarray(boxes, dtype=np.float32)
    scores = np.ascontiguousarray(scores, dtype=np.float32)
    b, n = scores.shape
    idx = np.zeros((b, max_output_size), np.int32)
    sc = np.zeros((b, max_output_size), np.float32)
    valid = np.zeros((b,), np.int32)
    lib().udal_oracle_nms_v5_batch(
        boxes.ctypes.data, scores.ctypes.data, b, n, int(max_output_size), float(iou_threshold),
        _score_thr(score_threshold), float(soft_nms_sigma), 1 if variant == "old" else 0,
        idx.ctypes.data, sc.ctypes.data, valid.ctypes.data,
    )
    return idx, sc, valid


# ---------------------------------------------------------------------------------------
# independent pure-Python transcription (small inputs only) - cross-checks nms_v5.c
# ---------------------------------------------------------------------------------------
def _iou_py(a, b):
    f = np.float32
    ay0, ax0 = min(a[0], a[2]), min(a[1], a[3])
    ay1, ax1 = max(a[0], a[2]), max(a[1], a[3])
    by0, bx0 = min(b[0], b[2]), min(b[1], b[3])
    by1, bx1 = max(b[0], b[2]), max(b[1], b[3])
    area_a = f(f(ay1 - ay0) * f(ax1 - ax0))
    area_b = f(f(by1 - by0) * f(bx1 - bx0))
    if area_a <= 0 or area_b <= 0:
        return f(0)
    iy0, ix0 = max(ay0, by0), max(ax0, bx0)
    iy1, ix1 = min(ay1, by1), min(ax1, bx1)
    ih = max(f(iy1 - iy0), f(0))
    iw = max(f(ix1 - ix0), f(0))
    inter = f(ih * iw)
    return f(inter / f(f(area_a + area_b) - inter))


def non_max_suppression_v5_py(boxes, scores, max_output_size, iou_threshold, score_threshold,
                              soft_nms_sigma, pad_to_max_output_size, variant="new"):
    import heapq

    f = np.float32
    boxes = np.asarray(boxes, np.float32)
    scores = np.asarray(scores, np.float32)
    thr = f(_score_thr(score_threshold))
    iou_thr = f(iou_threshold)
    sigma = f(soft_nms_sigma)
    soft = sigma > 0
    scale = f(f(-0.5) / sigma) if soft else f(0)
    heap = [(-float(s), i, 0) for i, s in enumerate(scores) if s > thr]
    heapq.heapify(heap)
    sel, sel_s = [], []
    while len(sel) < max_output_size and heap:
        neg, i, begin = heapq.heappop(heap)
        s = f(-neg)
        orig = s
        hard = False
        for j in range(len(sel) - 1, begin - 1, -1):
            u = _iou_py(boxes[i], boxes[sel[j]])
            w = f(math.exp(float(f(f(scale * u) * u))))
            if variant == "old":
                if not (u <= iou_thr):
                    w = f(0)
                s = f(s * w)
                if u >= iou_thr:
                    hard = True
                    break
            else:
                if not (soft or u <= iou_thr):
                    w = f(0)
                s = f(s * w)
                if (not soft) and u > iou_thr:
                    hard = True
                    break
            if s <= thr:
                break
        if not hard:
            if s == orig:
                sel.append(i)
                sel_s.append(s)
            elif s > thr:
                heapq.heappush(heap, (-float(s), i, len(sel)))
    valid = len(sel)
    idx = np.asarray(sel, np.int32)
    sc = np.asarray(sel_s, np.float32)
    if pad_to_max_output_size:
        idx = np.concatenate([idx, np.zeros(max_output_size - valid, np.int32)])
        sc = np.concatenate([sc, np.zeros(max_output_size - valid, np.float32)])
    return idx, sc, valid
