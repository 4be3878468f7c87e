"""Drop-in mirror of the reference's ``src/postprocess.py`` backed by libudal (CUDA, sm_100a).

Same names, argument order, return structure and error behaviour as the reference; arrays are
NumPy on the host (copied in, results copied out) or device arrays (``DeviceArray``, torch/cupy
tensors via ``__cuda_array_interface__`` / DLPack - borrowed zero-copy, results stay on the device).

Reference map (src/postprocess.py):
  to_list :44-50 | batch_map_fn :53-66 | clip_boxes :69-72 | merge_class_box_level_outputs :75-87 | topk_class_boxes :90-141 |
  pre_nms :144-339 | nms :342-420 | extract_uncertainties :423-469 | postprocess_global :472-621 |
  per_class_nms :624-716 | postprocess_per_class :719-740 |
  generate_detections_from_nms_output :743-785 | generate_detections :788-871 |
  transform_detections :874-887

Reference quirks kept: class uncertainty is the std of logits; aleatoric output is
the mean of stds; ``enable_softmax=False`` makes ``extract_uncertainties`` return None; with
top-k + per-class NMS the returned logits follow the reference's gather chain when
``params.get("strict_reference", True)`` (set it False for the logits of the selected anchors).

Arithmetic of the dense decode + MC moments (``pre_nms`` without top-k, ``extract_uncertainties``, ``postprocess_global``):
``params["decode_precision"]`` = ``"fp64"`` - the reference's float64 decode (utils_box.py:105-276) value for value - or
``"fp32"`` - the closed form in fp32: boxes / variances / scores within 1e-4 relative, mean logits and classes unchanged bit
for bit, through a persistent TMA-staged kernel at 0.77-0.99 of the HBM bandwidth (fp64: 0.34-0.58).  Without the key it
follows ``strict_reference``: ``"fp64"`` by default, ``"fp32"`` when ``strict_reference`` is False.
"""
import ctypes

import numpy as np

from . import _lib, device, utils
from . import engine as _engine

CLASS_OFFSET = 1


def to_list(inputs):
    """postprocess.py:44-50."""
    if isinstance(inputs, dict):
        return [inputs[k] for k in sorted(inputs.keys())]
    if isinstance(inputs, list):
        return inputs
    if isinstance(inputs, tuple):
        return list(inputs)
    return None


def batch_map_fn(map_fn, inputs, *args):
    """postprocess.py:53-66: apply ``map_fn`` per image and stack the results (host or device arrays)."""
    first = inputs[0]
    batch_size = len(first) if isinstance(first, (list, tuple)) else first.shape[0]
    outputs = [map_fn([_row(x, i) for x in inputs]) for i in range(batch_size)]
    return [_stack(list(y)) for y in zip(*outputs)]


def _row(x, i):
    return x.slice0(i, i + 1).reshape(x.shape[1:]) if isinstance(x, device.DeviceArray) else x[i]


def _stack(items):
    if not isinstance(items[0], device.DeviceArray):
        return np.stack([np.asarray(v) for v in items])
    ctx, a0 = items[0].ctx, items[0]
    out = ctx.empty((len(items),) + a0.shape, a0.dtype)
    for i, a in enumerate(items):
        _lib.check(ctx.lib.udal_memcpy_d2d(ctx.handle, out.ptr + i * a0.nbytes, a.ptr, a0.nbytes))
    return out


def clip_boxes(boxes, image_size):
    """postprocess.py:69-72: clip [..., 4] boxes (ymin, xmin, ymax, xmax) to the image."""
    h, w = utils.parse_image_size(image_size)
    host = not _is_dev(boxes)
    eng = _any_engine()
    b, _ = device.as_device(eng.ctx, boxes, np.float32)
    if b.shape[-1] != 4:
        raise ValueError("clip_boxes expects a trailing dimension of 4")
    out = eng.ctx.empty(b.shape)
    _lib.check(eng.lib.udal_clip_boxes(eng.ctx.handle, b.ptr, b.size // 4, float(h), float(w), out.ptr))
    return out.numpy() if host else out


def merge_class_box_level_outputs(params, cls_outputs, box_outputs):
    """postprocess.py:75-87: per level [B,H,W,A*C] -> [B,H*W*A,C] and [B,H,W,4A] -> [B,H*W*A,4], levels concatenated
    (anchor order = level, y, x, anchor).  A pure re-layout: device-to-device copies, no arithmetic."""
    if params.get("data_format", "channels_last") == "channels_first":
        raise ValueError("channels_first head outputs are not supported (NHWC only)")
    nc = params["num_classes"]
    nlev = params["max_level"] - params["min_level"] + 1
    host = not any(_is_dev(x) for x in list(cls_outputs)[:nlev] + list(box_outputs)[:nlev])
    if host:
        cls = [np.asarray(cls_outputs[l]) for l in range(nlev)]
        box = [np.asarray(box_outputs[l]) for l in range(nlev)]
        b = cls[0].shape[0]
        return (np.concatenate([c.reshape(b, -1, nc) for c in cls], 1),
                np.concatenate([x.reshape(b, -1, 4) for x in box], 1))
    eng = _any_engine()
    outs = []
    for levels, width in ((cls_outputs, nc), (box_outputs, 4)):
        arrs = [device.as_device(eng.ctx, levels[l], np.float32)[0] for l in range(nlev)]
        b = arrs[0].shape[0]
        rows = [a.size // (b * width) for a in arrs]
        merged = eng.ctx.empty((b, sum(rows), width))
        for i in range(b):
            off = 0
            for a, r in zip(arrs, rows):
                nbytes = r * width * 4
                _lib.check(eng.lib.udal_memcpy_d2d(eng.ctx.handle, merged.ptr + (i * sum(rows) + off) * width * 4,
                                                   a.ptr + i * nbytes, nbytes))
                off += r
        outs.append(merged)
    return outs[0], outs[1]


def topk_class_boxes(params, cls_outputs, box_outputs, uncerts=None):
    """postprocess.py:90-141 on merged tensors: cls_outputs [B,N,C], box_outputs [B,N,4];
    uncerts = [class [B,N,C] | None, box [B,N,4] | None, box [B,N,4] | None] is gathered in place.
    top-k order is canonical (value descending, index ascending; the reference asks for sorted=False)."""
    host = not (_is_dev(cls_outputs) or _is_dev(box_outputs))
    eng = _any_engine()
    ctx, lib = eng.ctx, eng.lib
    cls, _ = device.as_device(ctx, cls_outputs, np.float32)
    box, _ = device.as_device(ctx, box_outputs, np.float32)
    b, n, c = cls.shape
    if c != params["num_classes"]:
        raise ValueError("cls_outputs must be [B,N,num_classes]")
    k = int(params["nms_configs"].get("max_nms_inputs", 0) or 0)
    if k > 0:
        flat = ctx.empty((b, k), np.int32)
        cls_topk = ctx.empty((b, k))
        _lib.check(lib.udal_topk(ctx.handle, cls.ptr, b, n * c, k, flat.ptr, cls_topk.ptr))
        indices, classes = ctx.empty((b, k), np.int32), ctx.empty((b, k), np.int32)
        _lib.check(lib.udal_divmod_i32(ctx.handle, flat.ptr, b * k, c, indices.ptr, classes.ptr))
        box_topk = ctx.empty((b, k, 4))
        _lib.check(lib.udal_gather_rows(ctx.handle, box.ptr, b, n, 4, indices.ptr, k, 0, box_topk.ptr))
        if uncerts is not None:
            for i in range(len(uncerts)):
                if uncerts[i] is None:
                    continue
                u, _ = device.as_device(ctx, uncerts[i], np.float32)
                if i == 0:  # class uncertainty: by (anchor, class)
                    g = ctx.empty((b, k))
                    _lib.check(lib.udal_gather_rows(ctx.handle, u.ptr, b, n * c, 1, flat.ptr, k, 0, g.ptr))
                else:       # box uncertainties: by anchor
                    g = ctx.empty((b, k, 4))
                    _lib.check(lib.udal_gather_rows(ctx.handle, u.ptr, b, n, 4, indices.ptr, k, 0, g.ptr))
                uncerts[i] = g.numpy() if host else g
    else:
        cls_topk, classes = ctx.empty((b, n)), ctx.empty((b, n), np.int32)
        _lib.check(lib.udal_max_reduce(ctx.handle, cls.ptr, b * n, c, cls_topk.ptr, classes.ptr))
        box_topk = box
        indices = np.tile(np.arange(n, dtype=np.int32)[None], (b, 1))
        if not host:
            indices = ctx.to_device(indices)
    res = [cls_topk, box_topk, classes, indices]
    if host:
        res = [r.numpy() if isinstance(r, device.DeviceArray) else r for r in res]
    if uncerts is not None:
        return res[0], res[1], res[2], res[3], uncerts
    return tuple(res)


def _is_dev(x):
    return isinstance(x, device.DeviceArray) or hasattr(x, "__cuda_array_interface__") or (
        hasattr(x, "__dlpack__") and not isinstance(x, np.ndarray))


def _finish(arrs, host):
    """DeviceArrays -> NumPy when the caller passed host arrays (one sync), else unchanged."""
    if not host:
        return arrs
    out = [a.copy_to_host(sync=False) if isinstance(a, device.DeviceArray) else a for a in arrs]
    for a in arrs:
        if isinstance(a, device.DeviceArray):
            a.ctx.sync()
            break
    return out


def _raw_inputs(params, cls_outputs, box_outputs):
    eng = _engine.get_engine(params)
    cls_outputs, box_outputs = to_list(cls_outputs), to_list(box_outputs)
    if eng.box_mc and eng.T == 1:
        b1 = np.shape(box_outputs[0])[1] if not _is_dev(box_outputs[0]) else box_outputs[0].shape[1]
        if b1 != 1:
            # postprocess.py:180-203: with one sample the reference concatenates instead of stacking
            raise ValueError("mc_dropoutsamp == 1 is only defined for batch size 1 in the reference")
    cls, batch, h1 = eng.level_inputs(cls_outputs, eng.A * eng.C, eng.cls_mc)
    box, batch_b, h2 = eng.level_inputs(box_outputs, eng.box_channels, eng.box_mc)
    if batch != batch_b:
        raise ValueError("class and box outputs disagree on the batch size")
    return eng, cls, box, batch, (h1 or h2)


def _check_unc_config(params):
    # extract_uncertainties (postprocess.py:437-462) only builds the uncertainty list under these
    # flags; pre_nms then indexes it - mirror the reference's failure mode as a clear error.
    has = bool(params["loss_attenuation"] or params["mc_dropout"])
    mc = bool(params["mc_classheadrate"] or params["mc_dropoutrate"] or params["mc_boxheadrate"])
    if mc and not has:
        raise TypeError("'NoneType' object does not support item assignment "
                        "(dropout rates set but mc_dropout and loss_attenuation are both False)")
    return has


def extract_uncertainties(params, cls_outputs, box_outputs):
    """postprocess.py:423-469 -> [boxes, uncerts, scores, classes, classes_multi] or None."""
    has_unc = _check_unc_config(params)
    eng, cls, box, batch, host = _raw_inputs(params, cls_outputs, box_outputs)
    if eng.k > 0:
        o = eng.prenms_topk(cls, box, batch)
        uncerts = [o.get("mcclass"), o.get("albox"), o.get("mcbox")] if has_unc else None
        res = [o["boxes"], uncerts, o["scores"], o["classes"], o["mean_logits"]]
    else:
        o = eng.decode_moments(cls, box, batch)
        uncerts = [o.get("std_logits"), o.get("albox"), o.get("mcbox")] if has_unc else None
        res = [o["boxes"], uncerts, o["scores"], o["classes"], o["mean_logits"]]
    if host:
        flat = [res[0]] + [u for u in (res[1] or []) if u is not None] + res[2:]
        conv = _finish(flat, True)
        it = iter(conv)
        res[0] = next(it)
        if res[1] is not None:
            res[1] = [next(it) if u is not None else None for u in res[1]]
        res[2], res[3], res[4] = next(it), next(it), next(it)
    if not params["enable_softmax"]:
        return None  # postprocess.py:467-469: ``return pre_nms_output.append(None)``
    return res


def pre_nms(params, cls_outputs, box_outputs, topk=True, uncerts=None):
    """postprocess.py:144-339 with the call pattern of its only live caller
    (extract_uncertainties): ``cls_outputs`` are the per-level MEAN logits, ``box_outputs`` the
    per-level box regressions (with a leading sample axis under box MC dropout) and ``uncerts`` =
    [per-level logit std | None, per-level sigma | None, None]."""
    cls_outputs, box_outputs = to_list(cls_outputs), to_list(box_outputs)
    if not topk:
        # postprocess.py:276-282: no candidate selection - every anchor, every class; scores = sigmoid of all
        # logits [B,N,C], classes = None.  Same decode as the max-reduce variant (which keeps all anchors too).
        params = dict(params, nms_configs=dict(params["nms_configs"], max_nms_inputs=0))
    eng = _engine.get_engine(params, cls_mc=False)
    la = eng.la
    host = not any(_is_dev(x) for x in cls_outputs + box_outputs)
    cls, batch, _ = eng.level_inputs(cls_outputs, eng.A * eng.C, False)
    half = 4 * eng.A
    box = []
    for l, b in enumerate(box_outputs):
        bt, _ = device.as_device(eng.ctx, b, np.float32)
        if la:
            if uncerts is None or uncerts[1] is None:
                raise TypeError("loss_attenuation needs uncerts[1] (the per-level sigma outputs)")
            sg, _ = device.as_device(eng.ctx, uncerts[1][l], np.float32)
            rows = bt.size // half
            merged = eng.ctx.empty(bt.shape[:-1] + (2 * half,))
            _lib.check(eng.lib.udal_concat_channels(eng.ctx.handle, bt.ptr, half, sg.ptr, half, rows, merged.ptr))
            bt = merged
        box.append(bt)
    box, _, _ = eng.level_inputs(box, eng.box_channels, eng.box_mc)
    std_merged = None
    if uncerts is not None and uncerts[0] is not None:
        # level merge of the logit std (postprocess.py:177-178) = the moments kernel with T = 1
        std_levels, _, _ = eng.level_inputs(to_list(uncerts[0]), eng.A * eng.C, False)
        # (the kernel reads the box head's full [T,B,...] extent even when only the logit moments are wanted)
        lead = (eng.T,) if eng.box_mc else ()
        dummy = [eng.ctx.zeros(lead + (batch, h, w, eng.box_channels)) for h, w in eng.level_hw]
        std_merged = eng.decode_moments(std_levels, dummy, batch, want=("mean_logits",))["mean_logits"]
    if eng.k > 0:
        o = eng.prenms_topk(cls, box, batch)
        mcclass = None
        if std_merged is not None:
            mcclass = eng.ctx.empty((batch, eng.k))
            _lib.check(eng.lib.udal_gather_rows(eng.ctx.handle, std_merged.ptr, batch, eng.N * eng.C, 1,
                                                o["topk_idx"].ptr, eng.k, 0, mcclass.ptr))
        unc = [mcclass, o.get("albox"), o.get("mcbox")]
    else:
        o = eng.decode_moments(cls, box, batch)
        unc = [std_merged, o.get("albox"), o.get("mcbox")]
    if uncerts is not None:
        for i in range(3):
            uncerts[i] = unc[i]
        if host:
            for i in range(3):
                if uncerts[i] is not None:
                    uncerts[i] = uncerts[i].numpy()
    boxes, scores, classes, multi = o["boxes"], o["scores"], o["classes"], o["mean_logits"]
    if not topk:
        scores = eng.ctx.empty(multi.shape)
        _lib.check(eng.lib.udal_sigmoid(eng.ctx.handle, multi.ptr, multi.size, scores.ptr))
        classes = None
    if host:
        boxes, scores, multi = boxes.numpy(), scores.numpy(), multi.numpy()
        classes = classes.numpy() if classes is not None else None
    out = [boxes, uncerts, scores, classes]
    if params["enable_softmax"]:
        out.append(multi)
    return out


def nms(params, boxes, scores, classes, padded, multiclass=None, uncerts1=None, uncerts2=None,
        uncerts3=None):
    """postprocess.py:342-420 - one image: boxes [N,4], scores [N], classes [N]."""
    eng = _engine.get_engine(params)  # raises ValueError for an invalid nms method
    host = not _is_dev(boxes)
    bx, _ = device.as_device(eng.ctx, boxes, np.float32)
    sc, _ = device.as_device(eng.ctx, scores, np.float32)
    n = bx.shape[0]
    idx, nms_scores, valid = eng.nms_v5(bx.reshape(1, n, 4), sc.reshape(1, n))
    mo = eng.max_out
    m = mo
    if not padded:
        m = int(valid.numpy()[0])  # un-padded outputs have a data-dependent length

    def gather(src, width, mode=0, dtype=np.float32):
        a, _ = device.as_device(eng.ctx, src, np.int32 if mode == 1 else np.float32)
        if a.ndim <= 1:
            width = 1  # [n] sources (top-k class uncertainties, 1-D multiclass): one value per row
        if a.size != n * width:
            raise ValueError("nms: a gathered tensor has %d values, expected %d rows x %d" % (a.size, n, width))
        out = eng.ctx.empty((m,) + ((width,) if a.ndim > 1 else ()), dtype)
        _lib.check(eng.lib.udal_gather_rows(eng.ctx.handle, a.ptr, 1, n, width, idx.ptr, m, mode, out.ptr))
        return out

    cls_in = classes if _is_dev(classes) else np.asarray(classes, np.int32)
    out = [gather(bx, 4), nms_scores.reshape(mo).slice0(0, m), gather(cls_in, 1, mode=1),
           valid.reshape(1).numpy()[0] if host else valid.reshape(1)]
    if multiclass is not None:
        out.append(gather(multiclass, int(np.shape(multiclass)[-1]) if not _is_dev(multiclass) else multiclass.shape[-1]))
    if uncerts1 is not None:
        for u in (uncerts1, uncerts2, uncerts3):
            w = int(np.shape(u)[-1]) if not _is_dev(u) else u.shape[-1]
            out.append(gather(u, w))
    if host:
        out = [o.numpy() if isinstance(o, device.DeviceArray) else o for o in out]
    return out


def postprocess_global(params, cls_outputs, box_outputs, image_scales=None):
    """postprocess.py:472-621 - serving variant (max-reduce + global NMS + all uncertainties).

    Returns (boxes|albox|mcbox [B,M,4..12], scores [B,M], class|mcclass-std [B,M,(1+C)],
    valid_len [B], logits [B,M,C]) - the logits entry only with ``enable_softmax``."""
    _check_unc_config(params)
    if not params["enable_softmax"]:
        # extract_uncertainties returns None -> the reference fails while unpacking
        raise TypeError("cannot unpack non-iterable NoneType object")
    eng, cls, box, batch, host = _raw_inputs(params, cls_outputs, box_outputs)
    sc, sptr = eng.scales_input(image_scales, batch)
    bufs = eng.postprocess_global(cls, box, batch, sptr)
    out = [bufs["boxes"], bufs["scores"], bufs["classes"], bufs["valid"], bufs["logits"]]
    return tuple(_finish(out, host))


def per_class_nms(params, boxes, scores, classes, image_scales=None, logits=None):
    """postprocess.py:624-716 - boxes [B,K,4], scores [B,K], classes [B,K]."""
    eng = _engine.get_engine(params)
    host = not _is_dev(boxes)
    bx, _ = device.as_device(eng.ctx, boxes, np.float32)
    sc, _ = device.as_device(eng.ctx, scores, np.float32)
    cl, _ = device.as_device(eng.ctx, classes if _is_dev(classes) else np.asarray(classes, np.int32), np.int32)
    batch, k = sc.shape
    ssc, sptr = eng.scales_input(image_scales, batch)
    bufs, st = eng.detections_buffers(batch, False)
    lptr, lrows = 0, 0
    if logits is not None:
        lg, _ = device.as_device(eng.ctx, logits, np.float32)
        lptr, lrows = lg.ptr, lg.shape[1]
    else:
        st.logits = None
    strict = 1 if params.get("strict_reference", True) else 0
    _lib.check(eng.lib.udal_per_class_nms(eng.ctx.handle, bx.ptr, sc.ptr, cl.ptr, batch, k, sptr, lptr,
                                          lrows, strict, ctypes.byref(st)))
    out = [bufs["boxes"], bufs["scores"], bufs["classes"], bufs["valid"]]
    if logits is not None:
        out.append(bufs["logits"])
    return tuple(_finish(out, host))


def postprocess_per_class(params, cls_outputs, box_outputs, image_scales=None):
    """postprocess.py:719-740 - eval variant (top-k + per-class NMS, no uncertainties out)."""
    _check_unc_config(params)
    if not params["enable_softmax"]:
        raise TypeError("cannot unpack non-iterable NoneType object")
    eng, cls, box, batch, host = _raw_inputs(params, cls_outputs, box_outputs)
    if eng.k <= 0:
        raise ValueError("postprocess_per_class needs nms_configs.max_nms_inputs > 0 (eval.py:75)")
    sc, sptr = eng.scales_input(image_scales, batch)
    bufs = eng.postprocess_per_class(cls, box, batch, sptr, params.get("strict_reference", True))
    out = [bufs["boxes"], bufs["scores"], bufs["classes"], bufs["valid"], bufs["logits"]]
    return tuple(_finish(out, host))


def generate_detections_from_nms_output(nms_boxes_bs, nms_classes_bs, nms_scores_bs, image_ids,
                                        original_image_widths=None, flip=False,
                                        nms_multi_class_bs=None, params=None):
    """postprocess.py:743-785 -> [B,M,7(+C)] rows [id, x1, y1, x2, y2, score, class, logits...]."""
    host = not _is_dev(nms_boxes_bs)
    eng = _engine.get_engine(params) if params is not None else _any_engine()
    bx, _ = device.as_device(eng.ctx, nms_boxes_bs, np.float32)
    sc, _ = device.as_device(eng.ctx, nms_scores_bs, np.float32)
    cl, _ = device.as_device(eng.ctx, nms_classes_bs, np.float32)
    batch, mo = sc.shape
    ids, _ = device.as_device(eng.ctx, np.asarray(image_ids, np.float32).reshape(batch)
                              if not _is_dev(image_ids) else image_ids, np.float32)
    wptr = 0
    if original_image_widths is not None:
        wd, _ = device.as_device(eng.ctx, np.asarray(original_image_widths, np.float32).reshape(batch)
                                 if not _is_dev(original_image_widths) else original_image_widths, np.float32)
        wptr = wd.ptr
    lptr, nl = 0, 0
    if nms_multi_class_bs is not None:
        lg, _ = device.as_device(eng.ctx, nms_multi_class_bs, np.float32)
        lptr, nl = lg.ptr, lg.shape[-1]
    out = eng.ctx.empty((batch, mo, 7 + nl))
    _lib.check(eng.lib.udal_format_detections(
        eng.ctx.handle, bx.ptr, bx.shape[-1], sc.ptr, cl.ptr, cl.size // (batch * mo), ids.ptr, wptr,
        1 if flip else 0, lptr, nl, batch, mo, out.ptr))
    return out.numpy() if host else out


def _any_engine():
    if _engine._engines:
        return next(iter(_engine._engines.values()))
    from . import hparams_config
    return _engine.get_engine(hparams_config.get_detection_config("efficientdet-d0", image_size=64))


def generate_detections(params, cls_outputs, box_outputs, image_scales, image_ids, flip=False,
                        per_class_nms=True):
    """postprocess.py:788-871 (the ``pyfunc=False`` branch; the pyfunc branch is dead in the
    reference - it unpacks 3 values from a 4/5-element list and reads ``enable_softnax``)."""
    _, width = utils.parse_image_size(params["image_size"])
    if params["nms_configs"].get("pyfunc", True):
        raise KeyError("enable_softnax")  # what the reference's pyfunc branch raises (postprocess.py:806)
    host = not any(_is_dev(x) for x in to_list(cls_outputs))
    scales_host = np.asarray(image_scales if not _is_dev(image_scales) else device.as_device(
        _engine.get_engine(params).ctx, image_scales)[0].numpy(), np.float32)
    widths = scales_host * np.float32(width)
    post = postprocess_per_class if per_class_nms else postprocess_global
    res = post(params, cls_outputs, box_outputs, image_scales)
    if params["enable_softmax"]:
        nb, ns, nc, _, nm = res
    else:
        nb, ns, nc, _ = res
        nm = None
    det = generate_detections_from_nms_output(nb, nc, ns, image_ids, widths, flip, nm, params=params)
    return det


def transform_detections(detections):
    """postprocess.py:874-887 -> [B,M,7] rows [id, x, y, w, h, score, class]."""
    host = not _is_dev(detections)
    eng = _any_engine()
    d, _ = device.as_device(eng.ctx, detections, np.float32)
    rows = d.size // d.shape[-1]
    out = eng.ctx.empty(d.shape[:-1] + (7,))
    _lib.check(eng.lib.udal_transform_detections(eng.ctx.handle, d.ptr, rows, d.shape[-1], out.ptr))
    return out.numpy() if host else out
