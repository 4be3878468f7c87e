"""Restatement of the reference's NumPy NMS family (TEST INFRASTRUCTURE ONLY).

Follows src/nms_np.py: diou_nms :30-89, hard_nms :92-129, soft_nms :132-194, nms :197-220,
per_class_nms :223-278.  PINNED: ``src/nms_np.py`` is NumPy-only and importable in the build
container; ``tests/golden/make_golden.py`` runs it and the fixtures ``tests/golden/nms_np_*.npz``
hold its outputs; ``tests/test_oracle_golden.py`` checks this file against them.

Conventions restated: boxes carry the "+1" pixel convention in areas and intersections,
candidates are visited in ``argsort()[::-1]`` order (ties: NumPy's default sort, unstable),
a candidate survives a round when ``overlap <= thresh``; soft-NMS decays eagerly in place with
``exp(-iou^2/sigma)`` (gaussian) or ``1-iou`` above the threshold (linear) and keeps
``score >= score_thresh``.
"""
import numpy as np

DUMMY_SCORE = -1e5


def _pair_terms(dets, i, rest):
    x1, y1, x2, y2 = dets[:, 0], dets[:, 1], dets[:, 2], dets[:, 3]
    iw = np.maximum(0.0, np.minimum(x2[i], x2[rest]) - np.maximum(x1[i], x1[rest]) + 1)
    ih = np.maximum(0.0, np.minimum(y2[i], y2[rest]) - np.maximum(y1[i], y1[rest]) + 1)
    return iw * ih


def _greedy(dets, iou_thresh, penalty=None):
    """Shared skeleton of hard_nms (:92-129) and diou_nms (:30-89)."""
    iou_thresh = iou_thresh or 0.5
    areas = (dets[:, 2] - dets[:, 0] + 1) * (dets[:, 3] - dets[:, 1] + 1)
    order = dets[:, 4].argsort()[::-1]
    keep = []
    while order.size > 0:
        i, rest = order[0], order[1:]
        keep.append(i)
        inter = _pair_terms(dets, i, rest)
        metric = inter / (areas[i] + areas[rest] - inter)
        if penalty is not None:
            metric = metric - penalty(dets, i, rest)
        order = rest[np.where(metric <= iou_thresh)[0]]
    return dets[keep]


def hard_nms(dets, iou_thresh=None):
    return _greedy(dets, iou_thresh)


def _diou_penalty(dets, i, rest):
    x1, y1, x2, y2 = dets[:, 0], dets[:, 1], dets[:, 2], dets[:, 3]
    cx, cy = (x1 + x2) / 2, (y1 + y2) / 2
    ex1, ex2 = np.minimum(x1[i], x1[rest]), np.maximum(x2[i], x2[rest])
    ey1, ey2 = np.minimum(y1[i], y1[rest]), np.maximum(y2[i], y2[rest])
    diag = (ex2 - ex1) ** 2 + (ey2 - ey1) ** 2
    dist = (cx[i] - cx[rest]) ** 2 + (cy[i] - cy[rest]) ** 2
    return dist / (diag + 1e-10)


def diou_nms(dets, iou_thresh=None):
    return _greedy(dets, iou_thresh, _diou_penalty)


def soft_nms(dets, nms_configs):
    """:132-194 - eager in-place decay; row 0 swap with the current argmax each round."""
    method = nms_configs["method"]
    sigma = nms_configs["sigma"] or 0.5
    iou_thresh = nms_configs["iou_thresh"] or 0.3
    score_thresh = nms_configs["score_thresh"] or 0.001
    areas = (dets[:, 2] - dets[:, 0] + 1) * (dets[:, 3] - dets[:, 1] + 1)
    work = np.concatenate((dets, areas[:, None]), axis=1)
    kept = []
    while work.size > 0:
        top = np.argmax(work[:, 4], axis=0)
        work[[0, top], :] = work[[top, 0], :]
        kept.append(work[0, :-1])
        iw = np.maximum(np.minimum(work[0, 2], work[1:, 2]) - np.maximum(work[0, 0], work[1:, 0]) + 1, 0.0)
        ih = np.maximum(np.minimum(work[0, 3], work[1:, 3]) - np.maximum(work[0, 1], work[1:, 1]) + 1, 0.0)
        inter = iw * ih
        iou = inter / (work[0, 5] + work[1:, 5] - inter)
        if method == "linear":
            weight = np.ones_like(iou)
            weight[iou > iou_thresh] -= iou[iou > iou_thresh]
        elif method == "gaussian":
            weight = np.exp(-(iou * iou) / sigma)
        else:
            weight = np.ones_like(iou)
            weight[iou > iou_thresh] = 0
        work[1:, 4] *= weight
        work = work[np.where(work[1:, 4] >= score_thresh)[0] + 1, :]
    return np.vstack(kept)


def nms(dets, nms_configs):
    """:197-220."""
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
    """:223-278 -> [max_boxes_to_draw, 7] float32 rows [id, x1, y1, x2, y2, score, class+1]."""
    xyxy = boxes[:, [1, 0, 3, 2]]
    chunks = []
    for c in range(num_classes):
        pos = np.where(classes == c)[0]
        if pos.shape[0] == 0:
            continue
        top = nms(np.column_stack((xyxy[pos, :], scores[pos])), nms_configs)
        chunks.append(np.column_stack((np.repeat(image_id, len(top)), top, np.repeat(c + 1, len(top)))))

    def dummy(n):
        d = np.zeros((n, 7), dtype=np.float32)
        d[:, 0] = image_id[0]
        d[:, 5] = DUMMY_SCORE
        return d

    if chunks:
        allc = np.vstack(chunks)
        order = np.argsort(-allc[:, -2])
        det = np.array(allc[order[0:max_boxes_to_draw]], dtype=np.float32)
        det = np.vstack([det, dummy(max(max_boxes_to_draw - len(det), 0))])
    else:
        det = dummy(max_boxes_to_draw)
    det[:, 1:5] *= image_scale
    return det
