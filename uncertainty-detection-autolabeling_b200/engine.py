"""params dict -> libudal context.  One cached ``Engine`` per (configuration, device, batch-free).

The reference rebuilds anchors and re-traces its graph from ``params`` on every call
(postprocess.py:164-171); here the equivalent state (anchor table in HBM, kernel configuration,
scratch) lives in a context that is created once per distinct configuration and reused.
"""
import ctypes
import json
import math

import numpy as np

from . import _lib, device, utils


def nms_thresholds(nms_configs):
    """postprocess.py:373-388 -> (method id, sigma_tf, iou_thresh, score_thresh)."""
    method = nms_configs["method"]
    if method == "hard" or not method:
        return (_lib.NMS_HARD, 0.0, nms_configs["iou_thresh"] or 0.5,
                nms_configs["score_thresh"] or float("-inf"))
    if method == "gaussian":
        sigma = nms_configs["sigma"] or 0.5
        return (_lib.NMS_GAUSSIAN, sigma / 2, 0.5, nms_configs["score_thresh"] or 0.001)
    raise ValueError("Inference has invalid nms method {}".format(method))


def anchor_table(min_level, max_level, num_scales, aspect_ratios, anchor_scale, image_size):
    """anchors.py:135-215: float64 construction exactly as the reference's NumPy code (same
    expressions, same evaluation order), cast to float32 at the end.  Host-side by design - the
    reference builds it on the host too; the device only reads the table."""
    image_size_hw = utils.parse_image_size(image_size)
    feat_sizes = utils.get_feat_sizes(image_size, max_level)
    if isinstance(anchor_scale, (list, tuple)):
        assert len(anchor_scale) == max_level - min_level + 1
        anchor_scales = anchor_scale
    else:
        anchor_scales = [anchor_scale] * (max_level - min_level + 1)
    boxes_all = []
    for level in range(min_level, max_level + 1):
        stride = (feat_sizes[0]["height"] / float(feat_sizes[level]["height"]),
                  feat_sizes[0]["width"] / float(feat_sizes[level]["width"]))
        boxes_level = []
        for scale_octave in range(num_scales):
            for aspect in aspect_ratios:
                octave_scale = scale_octave / float(num_scales)
                scale = anchor_scales[level - min_level]
                base_x = scale * stride[1] * 2 ** octave_scale
                base_y = scale * stride[0] * 2 ** octave_scale
                if isinstance(aspect, list):
                    aspect_x, aspect_y = aspect
                else:
                    aspect_x = np.sqrt(aspect)
                    aspect_y = 1.0 / aspect_x
                half_x = base_x * aspect_x / 2.0
                half_y = base_y * aspect_y / 2.0
                x = np.arange(stride[1] / 2, image_size_hw[1], stride[1])
                y = np.arange(stride[0] / 2, image_size_hw[0], stride[0])
                xv, yv = np.meshgrid(x, y)
                xv, yv = xv.reshape(-1), yv.reshape(-1)
                b = np.vstack((yv - half_y, xv - half_x, yv + half_y, xv + half_x))
                boxes_level.append(np.expand_dims(np.swapaxes(b, 0, 1), axis=1))
        boxes_all.append(np.concatenate(boxes_level, axis=1).reshape([-1, 4]))
    # C-contiguous [N,4]: the raw pointer of this array is uploaded.  (astype keeps the memory order of its input, and
    # np.concatenate of the transposed per-aspect views above comes out Fortran-ordered for large levels - the device then
    # read a scrambled table at every geometry above ~100 px, silently: round 1 never compared decoded boxes at those
    # sizes with the oracle.  tests/test_abi_cpu.py pins the layout.)
    return np.ascontiguousarray(np.vstack(boxes_all), dtype=np.float32)


def decode_precision(params):
    """Arithmetic of the stand-alone decode + moments kernel: ``params["decode_precision"]`` = "fp64" (the reference's float64
    decode value for value) | "fp32" (closed form in fp32, 1e-4 contract, HBM bound).  Without the key it follows
    ``strict_reference`` (default True -> "fp64"; False -> "fp32")."""
    prec = params.get("decode_precision")
    if prec is None:
        prec = "fp64" if params.get("strict_reference", True) else "fp32"
    if prec not in ("fp64", "fp32"):
        raise ValueError("decode_precision must be 'fp64' or 'fp32', got %r" % (prec,))
    return prec


def _key(params, device_id, heads_mode):
    keys = ("image_size", "min_level", "max_level", "num_scales", "aspect_ratios", "anchor_scale",
            "num_classes", "loss_attenuation", "mc_dropout", "mc_dropoutrate", "mc_classheadrate",
            "mc_boxheadrate", "mc_dropoutsamp", "uncert_adjust_method", "fpn_num_filters",
            "box_class_repeats", "tf_nms_variant", "nms_prefilter_k", "decode_precision")
    d = {k: params.get(k) for k in keys}
    d["nms"] = {k: params["nms_configs"].get(k) for k in
                ("method", "iou_thresh", "score_thresh", "sigma", "max_nms_inputs", "max_output_size")}
    d["device"], d["heads_mode"] = device_id, heads_mode
    d["decode_precision"] = decode_precision(params)
    return json.dumps(d, sort_keys=True, default=str)


_engines = {}


def get_engine(params, device_id=None, heads_mode=None, cls_mc=None, box_mc=None, instance=0):
    """Cached Engine for this configuration.  ``cls_mc`` / ``box_mc`` override the flags derived
    from the dropout rates (used when a caller passes already reduced class outputs);
    ``instance`` > 0 gives additional independent contexts (own stream + scratch) on the same GPU,
    used to overlap the host<->device copies of one batch with the kernels of another."""
    if device_id is None:
        device_id = params.get("device", 0) or 0
    heads_mode = heads_mode or params.get("heads_mode", "fp32")
    k = _key(params, device_id, heads_mode) + repr((cls_mc, box_mc, instance))
    eng = _engines.get(k)
    if eng is None:
        eng = Engine(params, device_id, heads_mode, cls_mc=cls_mc, box_mc=box_mc)
        _engines[k] = eng
    return eng


def clear_engines():
    for e in _engines.values():
        e.ctx.close()
    _engines.clear()


class Engine:
    def __init__(self, params, device_id=0, heads_mode="fp32", cls_mc=None, box_mc=None,
                 level_hw=None, anchors=None):
        self.params = params
        h, w = utils.parse_image_size(params["image_size"])
        if level_hw is None:
            sizes = utils.get_feat_sizes(params["image_size"], params["max_level"])
            lv = list(range(params["min_level"], params["max_level"] + 1))
            self.level_hw = [(sizes[l]["height"], sizes[l]["width"]) for l in lv]
        else:  # generic geometry (stand-alone decode / moments entry points)
            self.level_hw = [tuple(x) for x in level_hw]
            lv = self.level_hw
        cfg = _lib.Config()
        cfg.abi_version = _lib.ABI_VERSION
        cfg.device = device_id
        cfg.image_h, cfg.image_w = h, w
        cfg.num_levels = len(lv)
        for i, (lh, lw) in enumerate(self.level_hw):
            cfg.level_h[i], cfg.level_w[i] = lh, lw
        self.A = params["num_scales"] * len(params["aspect_ratios"])
        self.C = params["num_classes"]
        self.F = params.get("fpn_num_filters", 64)
        self.R = params.get("box_class_repeats", 3)
        self.T = params["mc_dropoutsamp"]
        cfg.anchors_per_loc, cfg.num_classes = self.A, self.C
        cfg.num_filters, cfg.repeats, cfg.mc_samples = self.F, self.R, self.T
        self.la = bool(params["loss_attenuation"])
        self.cls_mc = bool(params["mc_classheadrate"] or params["mc_dropoutrate"]) if cls_mc is None else bool(cls_mc)
        self.box_mc = bool(params["mc_boxheadrate"] or params["mc_dropoutrate"]) if box_mc is None else bool(box_mc)
        cfg.loss_attenuation, cfg.cls_mc, cfg.box_mc = int(self.la), int(self.cls_mc), int(self.box_mc)
        # efficientdet_keras.py:907-916
        if params.get("mc_dropout"):
            self.rate_class = float(params["mc_classheadrate"] or params["mc_dropoutrate"] or 0.0)
            self.rate_box = float(params["mc_boxheadrate"] or params["mc_dropoutrate"] or 0.0)
        else:
            self.rate_class = self.rate_box = 0.0
        cfg.rate_class, cfg.rate_box = self.rate_class, self.rate_box
        cfg.inv_keep_class = 1.0 / (1.0 - self.rate_class)
        cfg.inv_keep_box = 1.0 / (1.0 - self.rate_box)
        method = params.get("uncert_adjust_method", "l-norm")
        if method not in _lib.DECODE_METHODS:
            if method == "sample":
                raise ValueError(
                    "uncert_adjust_method='sample' draws from tfp's MultivariateNormalDiag "
                    "(utils_box.py:162-184) and is not offered on the device; use 'l-norm'")
            raise ValueError("unknown uncert_adjust_method {}".format(method))
        cfg.decode_method = _lib.DECODE_METHODS[method]
        nms = params["nms_configs"]
        cfg.nms_method, sigma_tf, iou, thr = nms_thresholds(nms)
        cfg.nms_sigma_tf, cfg.nms_iou_thresh = sigma_tf, iou
        cfg.nms_score_thresh = thr if not math.isinf(thr) else -math.inf
        cfg.nms_variant_old = 1 if params.get("tf_nms_variant", "new") == "old" else 0
        cfg.max_nms_inputs = int(nms.get("max_nms_inputs", 0) or 0)
        cfg.max_output_size = int(nms.get("max_output_size", 100))
        modes = {"fp32": _lib.HEADS_FP32, "bf16": _lib.HEADS_BF16_TC, "bf16_tc": _lib.HEADS_BF16_TC,
                 "fp16": _lib.HEADS_FP16_TC, "fp16_tc": _lib.HEADS_FP16_TC, "fp32x3": _lib.HEADS_FP32X3_TC,
                 "fp32_tc": _lib.HEADS_FP32X3_TC}
        if heads_mode not in modes:
            raise ValueError("heads_mode must be one of fp32 | fp32x3 | fp16 | bf16, got %r" % (heads_mode,))
        cfg.heads_mode = modes[heads_mode]
        cfg.prefilter_k = int(params.get("nms_prefilter_k", 0) or 0)
        cfg.decode_precision = 1 if decode_precision(params) == "fp32" else 0
        self.cfg = cfg
        self.ctx = device.Context(cfg)
        self.lib = self.ctx.lib
        if anchors is None:
            self.anchors_host = anchor_table(params["min_level"], params["max_level"], params["num_scales"],
                                             params["aspect_ratios"], params["anchor_scale"],
                                             params["image_size"])
        else:
            self.anchors_host = np.ascontiguousarray(anchors, dtype=np.float32)
        self.N = self.anchors_host.shape[0]
        self.P = sum(a * b for a, b in self.level_hw)
        assert self.N == self.P * self.A
        self.anchors_host = np.ascontiguousarray(self.anchors_host, dtype=np.float32)
        assert self.anchors_host.flags["C_CONTIGUOUS"] and self.anchors_host.shape == (self.N, 4)
        _lib.check(self.lib.udal_set_anchors(self.ctx.handle, self.anchors_host.ctypes.data, self.N))
        self.box_channels = 4 * self.A * (2 if self.la else 1)
        self.max_out = cfg.max_output_size
        self.k = cfg.max_nms_inputs
        self.weights_set = False
        self._feat_format = _lib.FEAT_F32

    # ---- inputs ----------------------------------------------------------------------------
    def level_inputs(self, outputs, channels, mc):
        """list of per-level arrays (host or device) -> (list of DeviceArray, batch, any_host)."""
        arrs, any_host, batch = [], False, None
        if len(outputs) != len(self.level_hw):
            raise ValueError("expected %d levels, got %d" % (len(self.level_hw), len(outputs)))
        for l, x in enumerate(outputs):
            a, was_host = device.as_device(self.ctx, x, np.float32)
            any_host |= was_host
            lh, lw = self.level_hw[l]
            want_tail = (lh, lw, channels)
            if mc:
                if a.ndim != 5 or a.shape[0] != self.T or a.shape[2:] != want_tail:
                    raise ValueError("level %d: expected [T=%d,B,%d,%d,%d], got %s"
                                     % (l, self.T, lh, lw, channels, a.shape))
                b = a.shape[1]
            else:
                if a.ndim != 4 or a.shape[1:] != want_tail:
                    raise ValueError("level %d: expected [B,%d,%d,%d], got %s"
                                     % (l, lh, lw, channels, a.shape))
                b = a.shape[0]
            if batch is None:
                batch = b
            elif batch != b:
                raise ValueError("levels disagree on the batch size")
            arrs.append(a)
        return arrs, batch, any_host

    def scales_input(self, image_scales, batch):
        if image_scales is None:
            return None, 0
        # (isinstance first: reading a DeviceArray's __cuda_array_interface__ synchronises the context - the property
        # promises finished data to foreign consumers - and would serialise back-to-back udal_run calls)
        a, _ = device.as_device(self.ctx, image_scales if device.is_device_array(image_scales)
                                else np.asarray(image_scales, np.float32), np.float32)
        if a.size != batch:
            raise ValueError("image_scales must have one entry per image")
        return a, a.ptr

    # ---- kernels ---------------------------------------------------------------------------
    def decode_moments(self, cls, box, batch, want=("mean_logits", "std_logits", "boxes", "albox",
                                                    "mcbox", "scores", "classes"), out=None):
        """``out``: the dict a previous call returned - its arrays are written again (no allocation on this call)."""
        n, c = self.N, self.C
        reuse = out
        out = {}
        st = _lib.PreNmsOut()
        def mk(name, shape, dtype=np.float32, cond=True):
            if name in want and cond:
                out[name] = reuse[name] if reuse is not None else self.ctx.empty(shape, dtype)
                setattr(st, name, out[name].ptr)
        mk("mean_logits", (batch, n, c))
        mk("std_logits", (batch, n, c), cond=self.cls_mc)
        mk("boxes", (batch, n, 4))
        mk("albox", (batch, n, 4), cond=self.la)
        mk("mcbox", (batch, n, 4), cond=self.box_mc)
        mk("scores", (batch, n))
        mk("classes", (batch, n), np.int32)
        _lib.check(self.lib.udal_decode_moments(self.ctx.handle, device.ptr_array(cls),
                                                device.ptr_array(box), batch, ctypes.byref(st)))
        return out

    def prenms_topk(self, cls, box, batch, want_unc=True):
        n, c, k = self.N, self.C, self.k
        out = {}
        st = _lib.PreNmsTopkOut()
        def mk(name, shape, dtype=np.float32, cond=True):
            if cond:
                out[name] = self.ctx.empty(shape, dtype)
                setattr(st, name, out[name].ptr)
        mk("mean_logits", (batch, n, c))
        mk("topk_idx", (batch, k), np.int32)
        mk("boxes", (batch, k, 4))
        mk("albox", (batch, k, 4), cond=self.la and want_unc)
        mk("mcbox", (batch, k, 4), cond=self.box_mc and want_unc)
        mk("mcclass", (batch, k), cond=self.cls_mc and want_unc)
        mk("scores", (batch, k))
        mk("classes", (batch, k), np.int32)
        _lib.check(self.lib.udal_prenms_topk(self.ctx.handle, device.ptr_array(cls),
                                             device.ptr_array(box), batch, ctypes.byref(st)))
        return out

    def detections_buffers(self, batch, global_variant):
        mo, c = self.max_out, self.C
        if global_variant:
            nb = 1 + int(self.la) + int(self.box_mc)
            cw = 1 + (c if self.cls_mc else 0)
            shapes = dict(boxes=(batch, mo, 4 * nb), scores=(batch, mo),
                          classes=(batch, mo, cw) if self.cls_mc else (batch, mo),
                          valid=(batch,), logits=(batch, mo, c))
        else:
            shapes = dict(boxes=(batch, mo, 4), scores=(batch, mo), classes=(batch, mo),
                          valid=(batch,), logits=(batch, mo, c))
        bufs = {k: self.ctx.empty(s, np.int32 if k == "valid" else np.float32) for k, s in shapes.items()}
        st = _lib.Detections()
        for k, v in bufs.items():
            setattr(st, k, v.ptr)
        return bufs, st

    def postprocess_global(self, cls, box, batch, scales_ptr):
        bufs, st = self.detections_buffers(batch, True)
        _lib.check(self.lib.udal_postprocess_global(self.ctx.handle, device.ptr_array(cls),
                                                    device.ptr_array(box), batch, scales_ptr,
                                                    ctypes.byref(st)))
        return bufs

    def postprocess_per_class(self, cls, box, batch, scales_ptr, strict_reference):
        bufs, st = self.detections_buffers(batch, False)
        _lib.check(self.lib.udal_postprocess_per_class(self.ctx.handle, device.ptr_array(cls),
                                                       device.ptr_array(box), batch, scales_ptr,
                                                       1 if strict_reference else 0, ctypes.byref(st)))
        return bufs

    def topk(self, values, k):
        v, _ = device.as_device(self.ctx, values, np.float32)
        b, m = v.shape[0], v.size // v.shape[0]
        idx = self.ctx.empty((b, k), np.int32)
        val = self.ctx.empty((b, k), np.float32)
        _lib.check(self.lib.udal_topk(self.ctx.handle, v.ptr, b, m, k, idx.ptr, val.ptr))
        return val, idx

    def nms_v5(self, boxes, scores):
        bx, _ = device.as_device(self.ctx, boxes, np.float32)
        sc, _ = device.as_device(self.ctx, scores, np.float32)
        s, n = sc.shape
        idx = self.ctx.empty((s, self.max_out), np.int32)
        ss = self.ctx.empty((s, self.max_out), np.float32)
        valid = self.ctx.empty((s,), np.int32)
        _lib.check(self.lib.udal_nms_v5(self.ctx.handle, bx.ptr, sc.ptr, s, n, idx.ptr, ss.ptr, valid.ptr))
        return idx, ss, valid

    # ---- heads -----------------------------------------------------------------------------
    def set_head_weights(self, weights):
        """weights: {'class': tower, 'box': tower}; tower = dict(dw[R][3,3,F], pw[R][F,F], b[R][F],
        bn[R][L]{gamma,beta,mean,var}, dwp[3,3,F], pwp[F,Cout], bp[Cout])."""
        for head, name in ((_lib.HEAD_CLASS, "class"), (_lib.HEAD_BOX, "box")):
            w = weights[name]
            f32 = lambda x: np.ascontiguousarray(x, dtype=np.float32)
            dw = f32(np.stack([np.asarray(d).reshape(9, self.F) for d in w["dw"]]))
            pw = f32(np.stack(w["pw"]))
            bias = f32(np.stack(w["b"]))
            bn = {k: f32(np.stack([np.stack([lv[k] for lv in rep]) for rep in w["bn"]]))
                  for k in ("gamma", "beta", "mean", "var")}
            dwp = f32(np.asarray(w["dwp"]).reshape(9, self.F))
            pwp, bp = f32(w["pwp"]), f32(w["bp"])
            cout = self.A * self.C if head == _lib.HEAD_CLASS else self.box_channels
            if pwp.shape != (self.F, cout) or bp.shape != (cout,):
                raise ValueError("%s predict layer: expected pwp [%d,%d]" % (name, self.F, cout))
            if dw.shape != (self.R, 9, self.F) or pw.shape != (self.R, self.F, self.F):
                raise ValueError("%s tower: expected R=%d, F=%d" % (name, self.R, self.F))
            if bn["gamma"].shape != (self.R, len(self.level_hw), self.F):
                raise ValueError("%s tower: BN tables must be [R][L][F]" % name)
            p = lambda a: a.ctypes.data
            _lib.check(self.lib.udal_set_head_weights(
                self.ctx.handle, head, p(dw), p(pw), p(bias), p(bn["gamma"]), p(bn["beta"]),
                p(bn["mean"]), p(bn["var"]), p(dwp), p(pwp), p(bp)))
        self.weights_set = True

    def feats_input(self, feats):
        """BiFPN feature maps -> device arrays.  float32 (default) or, with heads_mode fp16 and 64 filters, float16 - all
        levels alike; the context is told which (udal_set_feature_format)."""
        arrs, any_host, batch = [], False, None
        first = feats[0]
        dt = np.dtype(getattr(first, "dtype", np.float32))
        if dt != np.float16:
            dt = np.dtype(np.float32)
        for l, x in enumerate(feats):
            if dt == np.float16 and np.dtype(getattr(x, "dtype", np.float32)) != np.float16:
                raise TypeError("feature levels must share one dtype (float16 or float32)")
            a, was_host = device.as_device(self.ctx, x, dt)
            lh, lw = self.level_hw[l]
            if a.ndim != 4 or a.shape[1:] != (lh, lw, self.F):
                raise ValueError("feature level %d: expected [B,%d,%d,%d], got %s" % (l, lh, lw, self.F, a.shape))
            batch = a.shape[0] if batch is None else batch
            if a.shape[0] != batch:
                raise ValueError("levels disagree on the batch size")
            any_host |= was_host
            arrs.append(a)
        fmt = _lib.FEAT_F16 if dt == np.float16 else _lib.FEAT_F32
        if fmt != self._feat_format:
            _lib.check(self.lib.udal_set_feature_format(self.ctx.handle, fmt))
            self._feat_format = fmt
        return arrs, batch, any_host

    def head_output_buffers(self, batch):
        cls = [self.ctx.empty(((self.T,) if self.cls_mc else ()) + (batch, h, w, self.A * self.C))
               for h, w in self.level_hw]
        box = [self.ctx.empty(((self.T,) if self.box_mc else ()) + (batch, h, w, self.box_channels))
               for h, w in self.level_hw]
        return cls, box

    def masks_input(self, masks, batch):
        if masks is None:
            return None, 0
        m, _ = device.as_device(self.ctx, masks if device.is_device_array(masks) else np.asarray(masks, np.uint8), np.uint8)
        want = (self.T, 2, len(self.level_hw), self.R, batch, self.F)
        if m.shape != want:
            raise ValueError("keep masks: expected %s, got %s" % (want, m.shape))
        return m, m.ptr

    def heads_sample(self, feats, masks=None, seed=0, out=None):
        if not self.weights_set:
            raise RuntimeError("head weights not set")
        f, batch, _ = self.feats_input(feats)
        m, mptr = self.masks_input(masks, batch)
        cls, box = out if out is not None else self.head_output_buffers(batch)
        _lib.check(self.lib.udal_heads_sample(self.ctx.handle, device.ptr_array(f), batch, mptr,
                                              ctypes.c_uint64(seed), device.ptr_array(cls),
                                              device.ptr_array(box)))
        return cls, box

    def run_prenms(self, feats, masks=None, seed=0):
        """features -> per-anchor tensors (dict of DeviceArray) through the kernels ``run`` launches
        (udal_run_prenms); max-reduce variant only."""
        if not self.weights_set:
            raise RuntimeError("head weights not set")
        f, batch, _ = self.feats_input(feats)
        m, mptr = self.masks_input(masks, batch)
        n, c = self.N, self.C
        out = {}
        st = _lib.PreNmsOut()
        for name, shape, dtype in (("mean_logits", (batch, n, c), np.float32), ("std_logits", (batch, n, c), np.float32),
                                   ("boxes", (batch, n, 4), np.float32), ("albox", (batch, n, 4), np.float32),
                                   ("mcbox", (batch, n, 4), np.float32), ("scores", (batch, n), np.float32),
                                   ("classes", (batch, n), np.int32)):
            if (name == "std_logits" and not self.cls_mc) or (name == "albox" and not self.la) or \
                    (name == "mcbox" and not self.box_mc):
                continue
            out[name] = self.ctx.empty(shape, dtype)
            setattr(st, name, out[name].ptr)
        _lib.check(self.lib.udal_run_prenms(self.ctx.handle, device.ptr_array(f), batch, mptr,
                                            ctypes.c_uint64(seed), ctypes.byref(st)))
        return out

    def run(self, feats, image_scales=None, masks=None, seed=0):
        if not self.weights_set:
            raise RuntimeError("head weights not set")
        f, batch, _ = self.feats_input(feats)
        m, mptr = self.masks_input(masks, batch)
        sc, sptr = self.scales_input(image_scales, batch)
        bufs, st = self.detections_buffers(batch, self.k == 0)
        _lib.check(self.lib.udal_run(self.ctx.handle, device.ptr_array(f), batch, mptr,
                                     ctypes.c_uint64(seed), sptr, ctypes.byref(st)))
        return bufs
