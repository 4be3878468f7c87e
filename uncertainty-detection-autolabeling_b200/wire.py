"""the wire formats the reference's consumers read.

* ``prediction_data.txt`` / ``validate_results.txt``: one Python-dict literal per detection and line, written with
  ``str(dict)`` and re-read with ``ast.literal_eval(line.replace("inf", "2e308"))``
  (src/infer_model.py:836-960, src/active_learning_loop.py:532, src/SSL_stac.py:345).  Array values pass
  through ``add_array_dict`` (src/utils_extra.py:67-81): float32, rounded to 4 decimals, nan_to_num, lists for
  vectors.
* COCO detection rows ``[id, x, y, w, h, score, class]``: ``postprocess.transform_detections``.
"""
import ast

import numpy as np


def add_array_dict(data_dict, source_array, target_key, select_index):
    """utils_extra.py:67-81."""
    source_array = np.asarray(source_array)
    if source_array.size > 0:
        vals = np.nan_to_num(np.around(source_array[select_index].astype("float32"), 4))
        if source_array[select_index].size > 1:
            vals = list(vals)
        data_dict[target_key] = vals
    return data_dict


def prediction_records(image_name, detections, autolabel=None, min_score=0.1, calib_key=None):
    """Records of ONE image in the reference's key order (uncalibrated keys; ``calib_key`` names the calibrated
    aleatoric std taken from ``autolabel["calib_albox"]``, e.g. "iso_perclscoo_albox").

    detections: the postprocess_global tuple of one image WITHOUT the batch axis: boxes|albox|mcbox [M,4k],
    scores [M], class|mcclass [M,1+C] (or [M]), logits [M,C] or None.  autolabel: the per-image slices of
    ``AutoLabeler.decide`` (entropy [M], calib_albox [M,4]) or None."""
    boxes_all, scores, classes_all, logits = detections
    boxes_all, scores = np.asarray(boxes_all), np.asarray(scores)
    classes_all = np.asarray(classes_all)
    boxes = boxes_all[:, 0:4]
    albox = boxes_all[:, 4:8] if boxes_all.shape[1] >= 8 else np.array([])
    mcbox = boxes_all[:, 8:12] if boxes_all.shape[1] >= 12 else np.array([])
    classes = classes_all[:, 0] if classes_all.ndim == 2 else classes_all
    mcclass = classes_all[:, 1:] if classes_all.ndim == 2 and classes_all.shape[1] > 1 else np.array([])
    records = []
    for sel in np.where(scores > min_score)[0]:
        d = {"image_name": image_name, "score_thresh": min_score, "top_5scores": list(scores[:5]),
             "det_score": scores[sel], "bbox": list(boxes[sel]), "class": classes[sel]}
        if logits is not None:
            lg = np.asarray(logits)
            d = add_array_dict(d, lg, "logits", sel)
            x = lg[sel]
            probab = np.exp(x - max(x)) / np.sum(np.exp(x - max(x)))   # utils_class.stable_softmax
            entropy = (np.asarray(autolabel["entropy"]) if autolabel is not None else
                       -np.sum(probab * np.nan_to_num(np.log2(np.maximum(probab, 10**-7)))).reshape(1)[[0] * len(scores)])
            d = add_array_dict(d, entropy, "entropy", sel)
            d["probab"] = list(probab)
        d = add_array_dict(d, mcclass, "uncalib_mcclass", sel)
        d = add_array_dict(d, albox, "uncalib_albox", sel)
        if calib_key and autolabel is not None:
            d = add_array_dict(d, np.asarray(autolabel["calib_albox"]), calib_key, sel)
        d = add_array_dict(d, mcbox, "uncalib_mcbox", sel)
        records.append(d)
    return records


def write_prediction_data(path, records, mode="a"):
    """infer_model.py:959-960: ``f.write(str(uncert_data) + "\\n")``."""
    with open(path, mode) as f:
        for r in records:
            f.write(_literal(r) + "\n")


def _literal(record):
    # NumPy >= 2 prints scalars as np.float32(...), which ast.literal_eval cannot read back; the reference ran on
    # NumPy 1.23 where str() of a scalar is the bare number.  Write bare Python numbers.
    def py(v):
        if isinstance(v, (list, tuple)):
            return [py(x) for x in v]
        if isinstance(v, np.generic):
            return v.item()
        if isinstance(v, np.ndarray):
            return [py(x) for x in v.tolist()] if v.ndim else v.item()
        return v
    return str({k: py(v) for k, v in record.items()})


def read_prediction_data(path):
    """The consumers' reader (src/active_learning_loop.py:532)."""
    out = []
    with open(path) as f:
        for line in f:
            if line.strip():
                out.append(ast.literal_eval(line.replace("inf", "2e308")))
    return out
