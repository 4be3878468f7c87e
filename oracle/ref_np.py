"""NumPy restatement of the reference post-processing path (TEST INFRASTRUCTURE ONLY).

Follows (paths relative to the reference's ``src/``):
  anchors.py:41-75, 100-218      decode_box_outputs, Anchors
  utils.py:516-559               parse_image_size, get_feat_sizes
  utils_box.py:105-276           decode_uncert (l-norm / n-flow / falsedec / sample)
  utils_extra.py:201-244         stack_mcpred, get_mcuncert
  postprocess.py:44-887          to_list .. transform_detections

Arithmetic conventions of the restatement (the third-party TF ops are not available here, see
``oracle/__init__.py`` for the pin status):
  * reductions over the MC axis are sequential fp32 sums in sample order, divided by T;
    std is the two-pass population form sqrt(mean((x-mean)^2)) (tf.math.reduce_std).
  * sigmoid(x) = fp32( 1 / (1 + exp(-fp64(x))) ).
  * top_k order is canonical: value descending, flat index ascending (what TF yields with
    sorted=True; sorted=False order is unspecified in TF).
  * decode_uncert computes in float64 and rounds to the input dtype, as the reference does.
"""
import numpy as np

from . import nms_ref

CLASS_OFFSET = 1
MAX_DETECTION_POINTS = 5000


# --------------------------------------------------------------------------------------
# utils.py:516-559
# --------------------------------------------------------------------------------------
def parse_image_size(image_size):
    """utils.py:516-540 - int -> (s, s); 'WxH' string -> (H, W); tuple passes through."""
    if isinstance(image_size, int):
        return (image_size, image_size)
    if isinstance(image_size, str):
        w, h = image_size.lower().split("x")
        return (int(h), int(w))
    if isinstance(image_size, tuple):
        return image_size
    raise ValueError(
        "image_size must be an int, WxH string, or (height, width)tuple. Was %r" % (image_size,)
    )


def get_feat_sizes(image_size, max_level):
    """utils.py:543-559 - [(h, w)] for levels 0..max_level with (s-1)//2+1 halving."""
    h, w = parse_image_size(image_size)
    sizes = [(h, w)]
    for _ in range(max_level):
        h, w = (h - 1) // 2 + 1, (w - 1) // 2 + 1
        sizes.append((h, w))
    return sizes


# --------------------------------------------------------------------------------------
# anchors.py:100-218
# --------------------------------------------------------------------------------------
def anchor_boxes(min_level, max_level, num_scales, aspect_ratios, anchor_scale, image_size):
    """anchors.py:135-215 - float64 construction, final cast to float32, order = level,
    y, x, (octave major, aspect minor)."""
    img_h, img_w = parse_image_size(image_size)
    sizes = get_feat_sizes(image_size, max_level)
    if isinstance(anchor_scale, (list, tuple)):
        assert len(anchor_scale) == max_level - min_level + 1
        scales = list(anchor_scale)
    else:
        scales = [anchor_scale] * (max_level - min_level + 1)
    out = []
    for level in range(min_level, max_level + 1):
        stride_y = sizes[0][0] / float(sizes[level][0])
        stride_x = sizes[0][1] / float(sizes[level][1])
        ys = np.arange(stride_y / 2, img_h, stride_y)
        xs = np.arange(stride_x / 2, img_w, stride_x)
        yc = np.repeat(ys, len(xs))
        xc = np.tile(xs, len(ys))
        per_anchor = []
        for octave in range(num_scales):
            for aspect in aspect_ratios:
                oscale = octave / float(num_scales)
                base_x = scales[level - min_level] * stride_x * 2**oscale
                base_y = scales[level - min_level] * stride_y * 2**oscale
                if isinstance(aspect, list):
                    ax, ay = aspect
                else:
                    ax = np.sqrt(aspect)
                    ay = 1.0 / ax
                hx = base_x * ax / 2.0
                hy = base_y * ay / 2.0
                per_anchor.append(np.stack([yc - hy, xc - hx, yc + hy, xc + hx], axis=-1))
        out.append(np.stack(per_anchor, axis=1).reshape(-1, 4))
    return np.concatenate(out, axis=0).astype(np.float32)


class Anchors:
    """anchors.py:100-133, 217-218."""

    def __init__(self, min_level, max_level, num_scales, aspect_ratios, anchor_scale, image_size):
        self.min_level = min_level
        self.max_level = max_level
        self.num_scales = num_scales
        self.aspect_ratios = aspect_ratios
        self.image_size = parse_image_size(image_size)
        self.boxes = anchor_boxes(
            min_level, max_level, num_scales, aspect_ratios, anchor_scale, image_size
        )

    def get_anchors_per_location(self):
        return self.num_scales * len(self.aspect_ratios)


def decode_box_outputs(pred_boxes, anchor_boxes_):
    """anchors.py:41-75 - plain exp/offset decode in the dtype of pred_boxes."""
    pred_boxes = np.asarray(pred_boxes)
    a = np.asarray(anchor_boxes_).astype(pred_boxes.dtype)
    two = pred_boxes.dtype.type(2.0)
    yca = (a[..., 0] + a[..., 2]) / two
    xca = (a[..., 1] + a[..., 3]) / two
    ha = a[..., 2] - a[..., 0]
    wa = a[..., 3] - a[..., 1]
    ty, tx, th, tw = (pred_boxes[..., i] for i in range(4))
    w = np.exp(tw) * wa
    h = np.exp(th) * ha
    yc = ty * ha + yca
    xc = tx * wa + xca
    return np.stack([yc - h / two, xc - w / two, yc + h / two, xc + w / two], axis=-1)


# --------------------------------------------------------------------------------------
# utils_box.py:105-276
# --------------------------------------------------------------------------------------
def decode_uncert(pred_boxes, box_uncert, anchor_boxes_, method="l-norm", n_samples=30, normals=None):
    """utils_box.py:105-276.  ``normals`` ([n_samples, 4, ...] standard normal draws) replaces
    the tfp sampler of the 'sample' method so the method is reproducible."""
    pred_boxes = np.asarray(pred_boxes)
    orig = pred_boxes.dtype
    a = np.asarray(anchor_boxes_).astype(np.float64)
    yca = (a[..., 0] + a[..., 2]) / 2
    xca = (a[..., 1] + a[..., 3]) / 2
    ha = a[..., 2] - a[..., 0]
    wa = a[..., 3] - a[..., 1]
    t = pred_boxes.astype(np.float64)
    ty, tx, th, tw = (t[..., i] for i in range(4))
    var = np.square(np.asarray(box_uncert).astype(np.float64))
    vy, vx, vh, vw = (var[..., i] for i in range(4))

    if method == "l-norm":  # utils_box.py:140-160
        w = np.exp(tw + vw / 2) * wa
        h = np.exp(th + vh / 2) * ha
        yc = ty * ha + yca
        xc = tx * wa + xca
        ymin, xmin, ymax, xmax = yc - h / 2.0, xc - w / 2.0, yc + h / 2.0, xc + w / 2.0
        dw = (np.exp(vw) - 1) * np.exp(2 * tw + vw) * wa**2
        dh = (np.exp(vh) - 1) * np.exp(2 * th + vh) * ha**2
        dyc = vy * ha**2
        dxc = vx * wa**2
        dymin = dymax = dyc + dh / 4.0
        dxmin = dxmax = dxc + dw / 4.0
    elif method == "n-flow":  # utils_box.py:186-245, tfp closed forms
        sy, sx, sh, sw = np.sqrt(vy), np.sqrt(vx), np.sqrt(vh), np.sqrt(vw)
        # Normal -> Scale(ha) -> Shift(yca): mean = ty*ha + yca, stddev = |ha*sy|
        yc = ty * ha + yca
        xc = tx * wa + xca
        dyc = np.square(np.abs(ha * sy))
        dxc = np.square(np.abs(wa * sx))
        # LogNormal(loc, scale) -> Scale: mean = exp(loc + scale^2/2) * s
        h = np.exp(th + np.square(sh) / 2.0) * ha
        w = np.exp(tw + np.square(sw) / 2.0) * wa
        lh_var = (np.exp(np.square(sh)) - 1.0) * np.exp(2.0 * th + np.square(sh))
        lw_var = (np.exp(np.square(sw)) - 1.0) * np.exp(2.0 * tw + np.square(sw))
        dh = np.square(np.abs(ha * np.sqrt(lh_var)))
        dw = np.square(np.abs(wa * np.sqrt(lw_var)))
        ymin, xmin, ymax, xmax = yc - h / 2.0, xc - w / 2.0, yc + h / 2.0, xc + w / 2.0
        dymin = dymax = dyc + dh / 4.0
        dxmin = dxmax = dxc + dw / 4.0
    elif method == "falsedec":  # utils_box.py:247-266
        w = np.exp(tw) * wa
        h = np.exp(th) * ha
        yc = ty * ha + yca
        xc = tx * wa + xca
        ymin, xmin, ymax, xmax = yc - h / 2.0, xc - w / 2.0, yc + h / 2.0, xc + w / 2.0
        dw = np.exp(vw) * wa
        dh = np.exp(vh) * ha
        dyc = vy * ha + yca
        dxc = vx * wa + xca
        dymin = np.abs(dyc - dh / 2.0)
        dxmin = np.abs(dxc - dw / 2.0)
        dymax = dyc + dh / 2.0
        dxmax = dxc + dw / 2.0
    elif method == "sample":  # utils_box.py:162-184
        if normals is None:
            raise ValueError("oracle decode_uncert(method='sample') needs injected normals")
        z = np.asarray(normals, dtype=np.float64)  # [n, 4, ...]
        assert z.shape[0] == n_samples and z.shape[1] == 4
        s_y = ty + np.sqrt(vy) * z[:, 0]
        s_x = tx + np.sqrt(vx) * z[:, 1]
        s_h = th + np.sqrt(vh) * z[:, 2]
        s_w = tw + np.sqrt(vw) * z[:, 3]
        w = np.exp(s_w) * wa
        h = np.exp(s_h) * ha
        yc = s_y * ha + yca
        xc = s_x * wa + xca
        cs = [yc - h / 2.0, xc - w / 2.0, yc + h / 2.0, xc + w / 2.0]
        means = [c.mean(axis=0) for c in cs]
        vars_ = [np.mean(np.square(c - m), axis=0) for c, m in zip(cs, means)]
        ymin, xmin, ymax, xmax = means
        dymin, dxmin, dymax, dxmax = vars_
    else:
        raise ValueError("unknown decode method {}".format(method))

    coords = np.stack([ymin, xmin, ymax, xmax], axis=-1).astype(orig)
    stds = np.sqrt(np.stack([dymin, dxmin, dymax, dxmax], axis=-1)).astype(orig)
    return coords, stds


# --------------------------------------------------------------------------------------
# utils_extra.py:201-244
# --------------------------------------------------------------------------------------
def mean_over_samples(x):
    """tf.reduce_mean(x, axis=0) restated: sequential fp32 sum over the leading axis / T."""
    x = np.asarray(x)
    acc = x[0].copy()
    for t in range(1, x.shape[0]):
        acc = acc + x[t]
    return acc / x.dtype.type(x.shape[0])


def std_over_samples(x):
    """tf.math.reduce_std(x, axis=0): sqrt(mean((x - mean)^2)), ddof = 0, input dtype."""
    x = np.asarray(x)
    m = mean_over_samples(x)
    d = x - m[None]
    return np.sqrt(mean_over_samples(d * d))


def stack_mcpred(output):
    """utils_extra.py:201-217 - list[5] of list[T] -> list[5] of [T, ...]."""
    o0, o1, o2, o3, o4 = output
    return [np.stack(o, axis=0) for o in (o0, o1, o2, o3, o4)]


def get_mcuncert(output):
    """utils_extra.py:220-244 - per level mean and population std over the sample axis."""
    o0, o1, o2, o3, o4 = output
    mean = [mean_over_samples(o) for o in (o0, o1, o2, o3, o4)]
    std = [std_over_samples(o) for o in (o0, o1, o2, o3, o4)]
    return mean, std


def sigmoid(x):
    x = np.asarray(x)
    return (1.0 / (1.0 + np.exp(-x.astype(np.float64)))).astype(x.dtype)


def top_k(values, k):
    """tf.math.top_k along the last axis in canonical order (value desc, index asc)."""
    values = np.asarray(values)
    order = np.argsort(-values, axis=-1, kind="stable")[..., :k]
    return np.take_along_axis(values, order, axis=-1), order.astype(np.int32)


# --------------------------------------------------------------------------------------
# postprocess.py
# --------------------------------------------------------------------------------------
def to_list(inputs):
    """postprocess.py:44-50."""
    if isinstance(inputs, dict):
        return [inputs[k] for k in sorted(inputs.keys())]
    if isinstance(inputs, list):
        return inputs
    if isinstance(inputs, tuple):
        return list(inputs)
    return None


def clip_boxes(boxes, image_size):
    """postprocess.py:69-72."""
    h, w = parse_image_size(image_size)
    hi = np.asarray([h, w, h, w], dtype=boxes.dtype)
    return np.minimum(np.maximum(boxes, boxes.dtype.type(0)), hi)


def merge_class_box_level_outputs(params, cls_outputs, box_outputs):
    """postprocess.py:75-87 - [B,H,W,A*C] x L -> [B,N,C]; [B,H,W,4A] x L -> [B,N,4]."""
    cls_all, box_all = [], []
    batch = np.asarray(cls_outputs[0]).shape[0]
    for level in range(params["max_level"] - params["min_level"] + 1):
        c = np.asarray(cls_outputs[level])
        b = np.asarray(box_outputs[level])
        if params["data_format"] == "channels_first":
            c = np.transpose(c, [0, 2, 3, 1])
            b = np.transpose(b, [0, 2, 3, 1])
        cls_all.append(c.reshape(batch, -1, params["num_classes"]))
        box_all.append(b.reshape(batch, -1, 4))
    return np.concatenate(cls_all, 1), np.concatenate(box_all, 1)


def _rows(x, idx):
    """x[b, idx[b, j]] (gather_nd with batch_dims=1 on the anchor axis)."""
    return np.take_along_axis(x, idx.reshape(idx.shape + (1,) * (x.ndim - 2)), axis=1)


def topk_class_boxes(params, cls_outputs, box_outputs, uncerts=None):
    """postprocess.py:90-141.  Mutates ``uncerts`` in place in the top-k branch, like the
    reference."""
    cls_outputs = np.asarray(cls_outputs)
    box_outputs = np.asarray(box_outputs)
    batch = cls_outputs.shape[0]
    num_classes = params["num_classes"]
    max_nms_inputs = params["nms_configs"].get("max_nms_inputs", 0)
    if max_nms_inputs > 0:
        flat = cls_outputs.reshape(batch, -1)
        _, flat_idx = top_k(flat, max_nms_inputs)
        indices = flat_idx // num_classes
        classes = flat_idx % num_classes
        cls_topk = np.take_along_axis(flat, flat_idx, axis=1)
        box_topk = _rows(box_outputs, indices)
        if uncerts is not None:
            for i in range(len(uncerts)):
                if uncerts[i] is None:
                    continue
                u = np.asarray(uncerts[i])
                if i != 0:
                    uncerts[i] = _rows(u, indices)
                else:
                    uncerts[i] = np.take_along_axis(u.reshape(batch, -1), flat_idx, axis=1)
    else:
        classes = np.argmax(cls_outputs, axis=-1).astype(np.int32)
        n = cls_outputs.shape[1]
        indices = np.tile(np.arange(n, dtype=np.int32)[None], [batch, 1])
        cls_topk = cls_outputs.max(-1)
        box_topk = box_outputs
    if uncerts is not None:
        return cls_topk, box_topk, classes, indices, uncerts
    return cls_topk, box_topk, classes, indices


def pre_nms(params, cls_outputs, box_outputs, topk=True, uncerts=None):
    """postprocess.py:144-339.

    Returns [boxes, uncerts, scores, classes(, classes_multi)] with
    uncerts = [mcclass, albox, mcbox] (entries None when the mode does not produce them)."""
    box_mc = bool(params["mc_boxheadrate"] or params["mc_dropoutrate"])
    cls_mc = bool(params["mc_classheadrate"] or params["mc_dropoutrate"])
    la = bool(params["loss_attenuation"])
    anchors = anchor_boxes(
        params["min_level"], params["max_level"], params["num_scales"],
        params["aspect_ratios"], params["anchor_scale"], params["image_size"],
    )
    cls_outputs = [np.asarray(c) for c in cls_outputs]
    box_outputs = [np.asarray(b) for b in box_outputs]
    nlev = len(box_outputs)

    # --- level merge (postprocess.py:172-208) -----------------------------------------
    if la and not box_mc:
        uncerts[1] = merge_class_box_level_outputs(params, cls_outputs, uncerts[1])[1]
    if cls_mc:
        uncerts[0] = merge_class_box_level_outputs(params, uncerts[0], box_outputs)[0]
    if box_mc:
        nsamp = box_outputs[0].shape[0]
        if la:
            uncerts[1] = np.stack(
                [
                    merge_class_box_level_outputs(
                        params, cls_outputs, [np.asarray(uncerts[1][i])[j] for i in range(nlev)]
                    )[1]
                    for j in range(nsamp)
                ],
                axis=0,
            )
        merged = [
            merge_class_box_level_outputs(
                params, cls_outputs, [box_outputs[i][j] for i in range(nlev)]
            )
            for j in range(nsamp)
        ]
        box_all = np.stack([m[1] for m in merged], axis=0)  # [T,B,N,4]
        cls_all = merged[-1][0]
        if nsamp == 1 and cls_all.shape[0] != 1:
            # postprocess.py:188-203: with one sample the reference concatenates instead of
            # stacking and then indexes the *batch* axis as if it were the sample axis.
            raise ValueError("mc_dropoutsamp == 1 is only defined for batch size 1")
    else:
        cls_all, box_all = merge_class_box_level_outputs(params, cls_outputs, box_outputs)

    classes_multi = cls_all.copy() if params["enable_softmax"] else None

    # --- candidate selection (postprocess.py:212-282) -----------------------------------
    if topk:
        if uncerts is not None and box_mc:
            nsamp = params["mc_dropoutsamp"]
            if la:
                uncerts[1] = np.stack(
                    [
                        topk_class_boxes(
                            params, cls_all, box_all[i], [uncerts[0], uncerts[1][i], uncerts[2]]
                        )[-1][1]
                        for i in range(nsamp)
                    ],
                    axis=0,
                )
            picked = []
            for i in range(nsamp):
                tmp = [uncerts[0], None, uncerts[2]]
                t_cls, t_box, classes, indices, tmp = topk_class_boxes(
                    params, cls_all, box_all[i], tmp
                )
                picked.append(t_box)
            box_all = np.stack(picked, axis=0)
            cls_sel = t_cls
            uncerts[0] = tmp[0]
            uncerts[2] = tmp[2]
        elif uncerts is not None:
            cls_sel, box_all, classes, indices, uncerts = topk_class_boxes(
                params, cls_all, box_all, uncerts
            )
        else:
            cls_sel, box_all, classes, indices = topk_class_boxes(params, cls_all, box_all)
        anchor_sel = anchors[indices]
    else:
        cls_sel = cls_all
        anchor_sel = anchors
        classes = None

    scores = sigmoid(cls_sel)

    # --- decode (postprocess.py:286-334) ------------------------------------------------
    method = params["uncert_adjust_method"]
    nsmp = params["decode_nsamples"]
    if la and not box_mc:
        boxes, uncerts[1] = decode_uncert(box_all, uncerts[1], anchor_sel, method=method, n_samples=nsmp)
    elif box_mc:
        nsamp = params["mc_dropoutsamp"]
        if la:
            dec = [
                decode_uncert(box_all[i], uncerts[1][i], anchor_sel, method=method, n_samples=nsmp)
                for i in range(nsamp)
            ]
            box_dec = np.stack([d[0] for d in dec], axis=0)
            uncerts[1] = mean_over_samples(np.stack([d[1] for d in dec], axis=0))
        else:
            box_dec = np.stack(
                [decode_box_outputs(box_all[i], anchor_sel) for i in range(nsamp)], axis=0
            )
        boxes = mean_over_samples(box_dec)
        uncerts[2] = std_over_samples(box_dec)
    else:
        boxes = decode_box_outputs(box_all, anchor_sel)

    out = [boxes, uncerts, scores, classes]
    if params["enable_softmax"]:
        out.append(classes_multi)
    return out


def nms_thresholds(nms_configs):
    """postprocess.py:373-388 -> (sigma_tf, iou_thresh, score_thresh, max_output_size)."""
    method = nms_configs["method"]
    if method == "hard" or not method:
        sigma = 0.0
        iou_thresh = nms_configs["iou_thresh"] or 0.5
        score_thresh = nms_configs["score_thresh"] or float("-inf")
    elif method == "gaussian":
        sigma = nms_configs["sigma"] or 0.5
        iou_thresh = 0.5
        score_thresh = nms_configs["score_thresh"] or 0.001
    else:
        raise ValueError("Inference has invalid nms method {}".format(method))
    return sigma / 2, iou_thresh, score_thresh, nms_configs["max_output_size"]


def nms(params, boxes, scores, classes, padded, multiclass=None, uncerts1=None, uncerts2=None,
        uncerts3=None):
    """postprocess.py:342-420."""
    sigma_tf, iou_thresh, score_thresh, max_out = nms_thresholds(params["nms_configs"])
    boxes = np.asarray(boxes)
    idx, nms_scores, valid = nms_ref.non_max_suppression_v5(
        boxes, np.asarray(scores), max_out, iou_thresh, score_thresh, sigma_tf, padded,
        variant=params.get("tf_nms_variant", "new"),
    )
    out = [
        boxes[idx],
        nms_scores,
        (np.asarray(classes)[idx] + CLASS_OFFSET).astype(boxes.dtype),
        np.int32(valid),
    ]
    if multiclass is not None:
        out.append(np.asarray(multiclass)[idx])
    if uncerts1 is not None:
        out.extend(
            [np.asarray(u)[idx].astype(boxes.dtype) for u in (uncerts1, uncerts2, uncerts3)]
        )
    return out


def extract_uncertainties(params, cls_outputs, box_outputs):
    """postprocess.py:423-469."""
    cls_outputs = to_list(cls_outputs)
    box_outputs = to_list(box_outputs)
    uncerts = None
    if params["loss_attenuation"] or params["mc_dropout"]:
        uncerts = [None, None, None]
        if params["mc_classheadrate"] or params["mc_dropoutrate"]:
            cls_outputs, uncerts[0] = get_mcuncert(cls_outputs)
        if params["loss_attenuation"]:
            split = int(np.asarray(box_outputs[0]).shape[-1] / 2)
            uncerts[1] = [np.asarray(b)[..., split:] for b in box_outputs]
            box_outputs = [np.asarray(b)[..., :split] for b in box_outputs]
    if params["enable_softmax"]:
        return pre_nms(params, cls_outputs, box_outputs, uncerts=uncerts)
    # postprocess.py:467-469: ``return pre_nms_output.append(None)`` evaluates to None.
    pre_nms(params, cls_outputs, box_outputs, uncerts=uncerts)
    return None


def postprocess_global(params, cls_outputs, box_outputs, image_scales=None):
    """postprocess.py:472-621."""
    res = extract_uncertainties(params, cls_outputs, box_outputs)
    if res is None:
        raise TypeError("cannot unpack non-iterable NoneType object")
    return global_from_pre_nms(params, res, image_scales)


def global_from_pre_nms(params, res, image_scales=None):
    """postprocess.py:497-621: the part of postprocess_global after extract_uncertainties (per-image NMS, clip,
    image scales, output assembly), on its own so that the parity tests can apply it to per-anchor tensors
    produced elsewhere (NMS keep-indices must be bit-exact GIVEN identical decoded scores and boxes)."""
    boxes, uncerts, scores, classes, classes_multi = res
    has_unc = bool(params["loss_attenuation"] or params["mc_dropout"])
    per_image = []
    for b in range(boxes.shape[0]):
        kw = {}
        if classes_multi is not None:
            kw["multiclass"] = classes_multi[b]
        if has_unc:
            filled = [np.zeros_like(boxes[b]) if u is None else u[b] for u in uncerts]
            kw.update(uncerts1=filled[0], uncerts2=filled[1], uncerts3=filled[2])
        per_image.append(nms(params, boxes[b], scores[b], classes[b], True, **kw))
    cols = [np.stack(c) for c in zip(*per_image)]
    nms_boxes, nms_scores, nms_classes, nms_valid = cols[:4]
    rest = cols[4:]
    nms_multi = rest.pop(0) if classes_multi is not None else None
    sel_unc = [None, None, None]
    if has_unc:
        for i in range(3):
            if uncerts[i] is not None:
                sel_unc[i] = rest[i]
    nms_boxes = clip_boxes(nms_boxes, params["image_size"])
    if image_scales is not None:
        sc = np.asarray(image_scales).astype(nms_boxes.dtype)[:, None, None]
        nms_boxes = nms_boxes * sc
        for i in (1, 2):
            if sel_unc[i] is not None:
                sel_unc[i] = sel_unc[i] * sc
    out = [nms_boxes, nms_scores, nms_classes, nms_valid]
    if params["enable_softmax"]:
        out.append(nms_multi)
    if sel_unc[0] is not None:
        out[2] = np.concatenate([out[2][..., None], sel_unc[0]], -1)
    if sel_unc[1] is not None:
        out[0] = np.concatenate([out[0], sel_unc[1]], -1)
    if sel_unc[2] is not None:
        out[0] = np.concatenate([out[0], sel_unc[2]], -1)
    return tuple(out)


def _gather_rows_oob_zero(x, idx):
    """tf.gather on GPU: out-of-range rows read as zero (the CPU kernel raises instead)."""
    x = np.asarray(x)
    idx = np.asarray(idx).reshape(-1)
    out = np.zeros((idx.shape[0],) + x.shape[1:], dtype=x.dtype)
    ok = (idx >= 0) & (idx < x.shape[0])
    out[ok] = x[idx[ok]]
    return out


def per_class_nms(params, boxes, scores, classes, image_scales=None, logits=None,
                  strict_reference=False):
    """postprocess.py:624-716.

    ``strict_reference`` reproduces the reference's logits chain (postprocess.py:659-666): the
    loop variable ``logits`` is overwritten by each class's NMS result, so class c>first gathers
    from the previous class's output (graph mode skips the where-gather because the static
    leading dim is unknown; out-of-range rows follow the TF-GPU gather rule = zeros).  With
    strict_reference=False the logits rows are the mean logits of the selected candidates'
    positions - what the code evidently intends when ``logits`` is indexed like ``boxes``."""
    boxes = np.asarray(boxes)
    scores = np.asarray(scores)
    classes = np.asarray(classes)
    max_out = params["nms_configs"].get("max_output_size", 100)
    sigma_tf, iou_thresh, score_thresh, nms_max = nms_thresholds(params["nms_configs"])
    res_b, res_s, res_c, res_v, res_l = [], [], [], [], []
    for b in range(boxes.shape[0]):
        bb, ss, cc, vv, ll = [], [], [], [], []
        chain = None if logits is None else np.asarray(logits[b])
        first = True
        for c in range(params["num_classes"]):
            pos = np.nonzero(classes[b] == c)[0]
            if pos.shape[0] == 0:
                continue
            idx, s, valid = nms_ref.non_max_suppression_v5(
                boxes[b][pos], scores[b][pos], nms_max, iou_thresh, score_thresh, sigma_tf, False,
                variant=params.get("tf_nms_variant", "new"),
            )
            bb.append(boxes[b][pos][idx])
            ss.append(s)
            cc.append((classes[b][pos][idx] + CLASS_OFFSET).astype(boxes.dtype))
            vv.append(valid)
            if logits is not None:
                if strict_reference:
                    if first:
                        chain = _gather_rows_oob_zero(chain, pos)
                    chain = _gather_rows_oob_zero(chain, idx)
                    ll.append(chain)
                else:
                    ll.append(_gather_rows_oob_zero(np.asarray(logits[b]), pos[idx]))
            first = False
        nb = np.concatenate(bb + [np.zeros((max_out, 4), boxes.dtype)], 0)
        ns = np.concatenate(ss + [np.zeros((max_out,), scores.dtype)], 0)
        nc = np.concatenate(cc + [np.zeros((max_out,), boxes.dtype)], 0)
        _, order = top_k(ns, max_out)
        res_b.append(nb[order])
        res_s.append(ns[order])
        res_c.append(nc[order])
        res_v.append(np.int32(min(max_out, int(np.sum(vv)))))
        if logits is not None:
            nl = np.concatenate(ll + [np.zeros((max_out, ll[0].shape[-1]), ll[0].dtype)], 0)
            res_l.append(nl[order])
    nms_boxes = np.stack(res_b)
    if image_scales is not None:
        nms_boxes = nms_boxes * np.asarray(image_scales).astype(nms_boxes.dtype)[:, None, None]
    out = [nms_boxes, np.stack(res_s), np.stack(res_c), np.stack(res_v)]
    if logits is not None:
        out.append(np.stack(res_l))
    return tuple(out)


def postprocess_per_class(params, cls_outputs, box_outputs, image_scales=None,
                          strict_reference=False):
    """postprocess.py:719-740."""
    res = extract_uncertainties(params, cls_outputs, box_outputs)
    if res is None:
        raise TypeError("cannot unpack non-iterable NoneType object")
    boxes, _, scores, classes, classes_multi = res
    return per_class_nms(params, boxes, scores, classes, image_scales, classes_multi,
                         strict_reference=strict_reference)


def generate_detections_from_nms_output(nms_boxes_bs, nms_classes_bs, nms_scores_bs, image_ids,
                                        original_image_widths=None, flip=False,
                                        nms_multi_class_bs=None):
    """postprocess.py:743-785 -> [id, x1, y1, x2, y2, score, class, logits...]."""
    ids = np.asarray(image_ids).astype(nms_scores_bs.dtype)[:, None] * np.ones_like(nms_scores_bs)
    if flip:
        cols = [
            ids,
            original_image_widths - nms_boxes_bs[:, :, 3],
            nms_boxes_bs[:, :, 0],
            original_image_widths - nms_boxes_bs[:, :, 1],
            nms_boxes_bs[:, :, 2],
            nms_scores_bs,
            nms_classes_bs,
        ]
    else:
        cols = [
            ids,
            nms_boxes_bs[:, :, 1],
            nms_boxes_bs[:, :, 0],
            nms_boxes_bs[:, :, 3],
            nms_boxes_bs[:, :, 2],
            nms_scores_bs,
            nms_classes_bs,
        ]
    if nms_multi_class_bs is not None:
        for i in range(nms_multi_class_bs.shape[-1]):
            cols.append(nms_multi_class_bs[:, :, i])
    return np.stack(cols, axis=-1)


def generate_detections(params, cls_outputs, box_outputs, image_scales, image_ids, flip=False,
                        per_class_nms=True, strict_reference=False):
    """postprocess.py:788-871 (pyfunc=False branch; the pyfunc branch is dead in the reference,
    SURVEY 3.4)."""
    _, width = parse_image_size(params["image_size"])
    image_scales = np.asarray(image_scales)
    widths = image_scales[:, None] * width
    if params["nms_configs"].get("pyfunc", True):
        raise NotImplementedError("pyfunc branch is broken in the reference (postprocess.py:804-840)")
    if per_class_nms:
        res = postprocess_per_class(params, cls_outputs, box_outputs, image_scales,
                                    strict_reference=strict_reference)
    else:
        res = postprocess_global(params, cls_outputs, box_outputs, image_scales)
    if params["enable_softmax"]:
        nb, ns, nc, _, nm = res
        return generate_detections_from_nms_output(nb, nc, ns, image_ids, widths, flip, nm)
    nb, ns, nc, _ = res
    return generate_detections_from_nms_output(nb, nc, ns, image_ids, widths, flip)


def transform_detections(detections):
    """postprocess.py:874-887 -> [id, x, y, w, h, score, class]."""
    d = np.asarray(detections)
    return np.stack(
        [d[:, :, 0], d[:, :, 1], d[:, :, 2], d[:, :, 3] - d[:, :, 1], d[:, :, 4] - d[:, :, 2],
         d[:, :, 5], d[:, :, 6]],
        axis=-1,
    )


# --------------------------------------------------------------------------------------
# default parameter dict (hparams_config.py:183-370, 373-452 restated as plain data)
# --------------------------------------------------------------------------------------
def default_params(**overrides):
    p = dict(
        min_level=3, max_level=7, num_scales=3, aspect_ratios=[1.0, 2.0, 0.5], anchor_scale=4.0,
        image_size=512, num_classes=7, data_format="channels_last",
        loss_attenuation=True, mc_dropout=True, mc_dropoutrate=0.0, mc_classheadrate=0.05,
        mc_boxheadrate=0.05, mc_dropoutsamp=10, uncert_adjust_method="l-norm",
        decode_nsamples=100, enable_softmax=True, fpn_num_filters=64, box_class_repeats=3,
        nms_configs=dict(method="gaussian", iou_thresh=None, score_thresh=0.0, sigma=None,
                         pyfunc=False, max_nms_inputs=0, max_output_size=100),
    )
    nms_over = overrides.pop("nms_configs", None)
    p.update(overrides)
    if nms_over:
        p["nms_configs"] = dict(p["nms_configs"], **nms_over)
    return p


def level_shapes(params):
    sizes = get_feat_sizes(params["image_size"], params["max_level"])
    return [sizes[l] for l in range(params["min_level"], params["max_level"] + 1)]


def num_anchors_per_location(params):
    return params["num_scales"] * len(params["aspect_ratios"])
