"""BiFPN on the device: mirror of the reference's ``FPNCells`` (src/efficientdet_keras.py:51-350, 766-847) - the
producer of the head sampler's input.

    cells = bifpn.FPNCells(params, weights)
    fpn_feats = cells(feats)          # feats: list[L] of [B,H_l,W_l,C_l] (backbone / resample_p6.. outputs)
    detections = heads.HeadSampler(params, head_weights).detect(fpn_feats, image_scales)

Structure as in the reference: ``FPNCells`` = ``fpn_cell_repeats`` x ``FPNCell`` = the nodes of
``fpn_configs.bifpn_config``; every ``FNode`` resamples its inputs to the node's level (1x1 conv + BN when the channel
count differs, max pooling with SAME padding down, nearest neighbour up), fuses them (fastattn | attn | sum and the
per-channel variants) and runs ``OpAfterCombine`` (swish -> separable 3x3 conv -> BN).  The arithmetic runs in three
device primitives (csrc/bifpn.cu: udal_conv1x1_bn, udal_bifpn_fuse, udal_sepconv_bn; 64-filter nodes - D0 - run their
separable conv on the tensor cores, fp32 accurate: udal_sepconv_tc, tables prepared once per node); fp32 in and out, within
2e-4 of the oracle; there is no CPU fallback.

Weights (plain dict, BN as {gamma, beta, mean, var}; ``weights_from_variables`` maps checkpoint variable names):
  {"cells": [{"fnodes": [{"resample": [None | {"w": [Cin,F], "b": [F], "bn": {...} | None}, ... per input],
                          "wsm": [scalar | [F], ... per input]            (absent for weight_method "sum"),
                          "dw": [3,3,F], "pw": [F,F], "b": [F], "bn": {...}}, ... per node]}, ... per cell]}
"""
import ctypes

import numpy as np

from . import _lib, device, engine, fpn_configs

BN_EPS = 1e-3  # utils_keras.py:78

_MODES = {"sum": (_lib.FUSE_SUM, 0), "fastattn": (_lib.FUSE_FASTATTN, 0), "attn": (_lib.FUSE_ATTN, 0),
          "channel_fastattn": (_lib.FUSE_FASTATTN, 1), "channel_attn": (_lib.FUSE_ATTN, 1)}


def _fold_bn(bn):
    scale = (np.asarray(bn["gamma"], np.float64) / np.sqrt(np.asarray(bn["var"], np.float64) + BN_EPS))
    shift = np.asarray(bn["beta"], np.float64) - np.asarray(bn["mean"], np.float64) * scale
    return scale.astype(np.float32), shift.astype(np.float32)


class FPNCells:
    """``FPNCells(params, weights)(feats) -> list[L] of [B,H_l,W_l,F]`` (device arrays for device inputs, NumPy for
    host inputs).  ``params``: the detection config dict (fpn_num_filters, fpn_cell_repeats, min/max_level,
    fpn_weight_method, apply_bn_for_resampling, conv_after_downsample, conv_bn_act_pattern, separable_conv)."""

    def __init__(self, params, weights, device_id=None):
        self.params = params
        self.F = int(params["fpn_num_filters"])
        self.min_level, self.max_level = int(params.get("min_level", 3)), int(params.get("max_level", 7))
        self.fpn_config = params.get("fpn_config") or fpn_configs.get_fpn_config(
            params.get("fpn_name"), self.min_level, self.max_level, params.get("fpn_weight_method"))
        method = self.fpn_config["weight_method"]
        if method not in _MODES:
            raise ValueError("unknown weight_method %s" % method)           # efficientdet_keras.py:123
        self.mode, self.per_channel = _MODES[method]
        if not params.get("separable_conv", True):
            raise ValueError("BiFPN on the device: separable_conv=True (the reference default) only")
        self.conv_bn_act = bool(params.get("conv_bn_act_pattern", False))
        self.conv_after_downsample = bool(params.get("conv_after_downsample", False))
        self.apply_bn = bool(params.get("apply_bn_for_resampling", True))
        self.pool_avg = int(params.get("pooling_type") == "avg")
        self.repeats = int(params.get("fpn_cell_repeats", len(weights["cells"])))
        if len(weights["cells"]) != self.repeats:
            raise ValueError("weights hold %d cells, fpn_cell_repeats is %d" % (len(weights["cells"]), self.repeats))
        self.engine = engine.get_engine(params, device_id=device_id)   # the weight-free cached engine: context only
        self.ctx, self.lib = self.engine.ctx, self.engine.lib
        self._cells = [self._upload_cell(c) for c in weights["cells"]]

    # ---- weights ------------------------------------------------------------------------------
    def _dev(self, a):
        return self.ctx.to_device(np.ascontiguousarray(a, np.float32))

    def _upload_cell(self, cell):
        nodes = self.fpn_config["nodes"]
        if len(cell["fnodes"]) != len(nodes):
            raise ValueError("a cell needs %d fnodes" % len(nodes))
        out = []
        for cfg, w in zip(nodes, cell["fnodes"]):
            n_in = len(cfg["inputs_offsets"])
            res = []
            for r in (w.get("resample") or [None] * n_in):
                if r is None:
                    res.append(None)
                    continue
                bn = _fold_bn(r["bn"]) if (self.apply_bn and r.get("bn") is not None) else None
                res.append(dict(w=self._dev(r["w"]), b=self._dev(r["b"]), cin=int(np.shape(r["w"])[0]),
                                scale=self._dev(bn[0]) if bn else None, shift=self._dev(bn[1]) if bn else None))
            wsm = None
            if self.mode != _lib.FUSE_SUM:
                wsm = [self._dev(np.reshape(np.asarray(v, np.float32), -1) if self.per_channel
                                 else np.reshape(np.asarray(v, np.float32), (1,))) for v in w["wsm"]]
                if len(wsm) != n_in:
                    raise ValueError("one edge weight per input")
            scale, shift = _fold_bn(w["bn"])
            node = dict(resample=res, wsm=wsm, dw=self._dev(np.reshape(w["dw"], (9, self.F))), pw=self._dev(w["pw"]),
                        b=self._dev(w["b"] if w.get("b") is not None else np.zeros(self.F)),
                        scale=self._dev(scale), shift=self._dev(shift), tc=None)
            if self.F == 64:
                # 64-channel nodes (D0): separable conv on the tensor cores, fp32 accurate (udal_sepconv_tc); tables built once
                table = ctypes.c_void_p()
                _lib.check(self.lib.udal_sepconv_tc_prepare(self.ctx.handle, node["pw"].ptr, node["b"].ptr, node["scale"].ptr,
                                                            node["shift"].ptr, ctypes.byref(table)))
                node["tc"] = table
            out.append(node)
        return out

    # ---- FNode (efficientdet_keras.py:166-173) --------------------------------------------------
    def _conv1x1(self, x, r):
        nb, h, w, cin = x.shape
        out = self.ctx.empty((nb, h, w, self.F))
        _lib.check(self.lib.udal_conv1x1_bn(self.ctx.handle, x.ptr, nb, h, w, cin, r["w"].ptr, r["b"].ptr,
                                            r["scale"].ptr if r["scale"] is not None else None,
                                            r["shift"].ptr if r["shift"] is not None else None, self.F, out.ptr))
        return out

    def _fuse(self, ins, wsm, nb, th, tw, act):
        n = len(ins)
        ptrs = (ctypes.c_void_p * 3)(*[x.ptr for x in ins])
        hs = (ctypes.c_int * 3)(*[x.shape[1] for x in ins])
        ws = (ctypes.c_int * 3)(*[x.shape[2] for x in ins])
        wp = (ctypes.c_void_p * 3)(*[v.ptr for v in wsm]) if wsm is not None else None
        out = self.ctx.empty((nb, th, tw, self.F))
        _lib.check(self.lib.udal_bifpn_fuse(self.ctx.handle, n, ptrs, hs, ws, wp, self.mode if wsm is not None else _lib.FUSE_SUM,
                                            self.per_channel, self.pool_avg, nb, th, tw, self.F, int(act), out.ptr))
        return out

    def _fnode(self, feats, cfg, w):
        level = cfg["feat_level"] - self.min_level
        nb, th, tw = feats[level].shape[0], feats[level].shape[1], feats[level].shape[2]
        ins = []
        for off, r in zip(cfg["inputs_offsets"], w["resample"]):
            x = feats[off]
            h, wd, ch = x.shape[1], x.shape[2], x.shape[3]
            down = h > th and wd > tw
            if not down and not (h <= th and wd <= tw):
                raise ValueError("Incompatible Resampling : feat shape {}x{} target_shape: {}x{}".format(h, wd, th, tw))
            if ch != self.F:
                if r is None:
                    raise ValueError("input %d has %d channels: its resample 1x1 conv is missing" % (off, ch))
                if down and self.conv_after_downsample:
                    pooled = self._fuse_raw_pool(x, nb, th, tw)      # pool first, 1x1 after (efficientdet_keras.py:333-338)
                    x = self._conv1x1(pooled, r)
                else:
                    x = self._conv1x1(x, r)
            ins.append(x)
        fused = self._fuse(ins, w["wsm"], nb, th, tw, act=not self.conv_bn_act)
        out = self.ctx.empty((nb, th, tw, self.F))
        act = _lib.ACT_BN_SWISH if self.conv_bn_act else _lib.ACT_BN
        if w["tc"] is not None:
            _lib.check(self.lib.udal_sepconv_tc(self.ctx.handle, fused.ptr, nb, th, tw, w["dw"].ptr, w["tc"], act, out.ptr))
        else:
            _lib.check(self.lib.udal_sepconv_bn(self.ctx.handle, fused.ptr, nb, th, tw, self.F, self.F, w["dw"].ptr, w["pw"].ptr,
                                                w["b"].ptr, w["scale"].ptr, w["shift"].ptr, act, out.ptr))
        return out

    def _fuse_raw_pool(self, x, nb, th, tw):
        # pooling of a map whose channel count is not F yet: the fuse kernel with one input and F = its channels
        ptrs = (ctypes.c_void_p * 3)(x.ptr, None, None)
        hs = (ctypes.c_int * 3)(x.shape[1], 0, 0)
        ws = (ctypes.c_int * 3)(x.shape[2], 0, 0)
        out = self.ctx.empty((nb, th, tw, x.shape[3]))
        _lib.check(self.lib.udal_bifpn_fuse(self.ctx.handle, 1, ptrs, hs, ws, None, _lib.FUSE_SUM, 0, self.pool_avg, nb, th, tw,
                                            x.shape[3], 0, out.ptr))
        return out

    # ---- FPNCells.call (efficientdet_keras.py:787-801) ------------------------------------------
    def __call__(self, feats):
        host = not any(device.is_device_array(x) for x in feats)
        cur = [device.as_device(self.ctx, x, np.float32)[0] for x in feats]
        nodes = self.fpn_config["nodes"]
        if len(cur) != self.max_level - self.min_level + 1:
            raise ValueError("expected %d feature levels" % (self.max_level - self.min_level + 1))
        for cell in self._cells:
            cell_feats = list(cur)
            for cfg, w in zip(nodes, cell):
                cell_feats.append(self._fnode(cell_feats, cfg, w))
            cur = []
            for level in range(self.min_level, self.max_level + 1):
                for i, cfg in enumerate(reversed(nodes)):
                    if cfg["feat_level"] == level:
                        cur.append(cell_feats[-1 - i])
                        break
        if host:
            cur = [c.numpy() for c in cur]
        return cur


def weights_from_variables(variables, params, prefix="fpn_cells"):
    """Checkpoint variables -> the weight dict above.  Names as the reference creates them (efficientdet_keras.py:
    FPNCells ``fpn_cells`` / FPNCell ``cell_%d`` / FNode ``fnode%d`` / ResampleFeatureMap
    ``resample_{i}_{offset}_{len(feats)}`` with ``conv2d`` + ``bn`` / WSM, WSM_1, .. / OpAfterCombine
    ``op_after_combine{len(feats)}`` with ``conv`` (depthwise_kernel, pointwise_kernel, bias) + ``bn``).
    ``variables``: mapping name -> array; a trailing ':0' is ignored."""
    var = {k[:-2] if k.endswith(":0") else k: np.asarray(v) for k, v in variables.items()}
    cfg = params.get("fpn_config") or fpn_configs.get_fpn_config(params.get("fpn_name"), int(params.get("min_level", 3)),
                                                                int(params.get("max_level", 7)), params.get("fpn_weight_method"))
    num_levels = int(params.get("max_level", 7)) - int(params.get("min_level", 3)) + 1

    def bn(base):
        return {"gamma": var[base + "/gamma"], "beta": var[base + "/beta"], "mean": var[base + "/moving_mean"],
                "var": var[base + "/moving_variance"]}

    cells = []
    for c in range(int(params["fpn_cell_repeats"])):
        fnodes = []
        for i, node in enumerate(cfg["nodes"]):
            base = "%s/cell_%d/fnode%d" % (prefix, c, i)
            n_feats = num_levels + i
            res, wsm = [], []
            for j, off in enumerate(node["inputs_offsets"]):
                rb = "%s/resample_%d_%d_%d" % (base, j, off, n_feats)
                if rb + "/conv2d/kernel" in var:
                    k = var[rb + "/conv2d/kernel"]
                    res.append({"w": k.reshape(k.shape[-2], k.shape[-1]), "b": var[rb + "/conv2d/bias"],
                                "bn": bn(rb + "/bn") if rb + "/bn/gamma" in var else None})
                else:
                    res.append(None)
                name = base + "/WSM" + ("" if j == 0 else "_%d" % j)
                if name in var:
                    wsm.append(var[name])
            ob = "%s/op_after_combine%d" % (base, n_feats)
            dw = var[ob + "/conv/depthwise_kernel"]
            pw = var[ob + "/conv/pointwise_kernel"]
            fnodes.append({"resample": res, "wsm": wsm or None, "dw": dw.reshape(3, 3, dw.shape[2]),
                           "pw": pw.reshape(pw.shape[-2], pw.shape[-1]), "b": var.get(ob + "/conv/bias"), "bn": bn(ob + "/bn")})
        cells.append({"fnodes": fnodes})
    return {"cells": cells}
