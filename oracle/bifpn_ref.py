"""CPU restatement of the BiFPN (FPNCells) of the reference (TEST INFRASTRUCTURE ONLY).

Follows (paths relative to the reference's ``src/``):
  efficientdet_keras.py:51-173    FNode: resample every input to the node's level, fuse_features (fastattn | attn | sum
                                  and the per-channel variants), op_after_combine
  efficientdet_keras.py:176-236   OpAfterCombine: swish -> SeparableConv2D 3x3 (bias unless conv_bn_act_pattern) -> BN
                                  (conv -> BN -> swish with conv_bn_act_pattern)
  efficientdet_keras.py:238-350   ResampleFeatureMap: 1x1 conv (+BN) when the channel count differs; MaxPooling2D /
                                  AveragePooling2D(pool = stride + 1, stride = (in-1)//out + 1, padding SAME) down,
                                  tf.compat.v1.image.resize_nearest_neighbor up
  efficientdet_keras.py:766-847   FPNCells / FPNCell: fpn_cell_repeats cells, outputs = the last node of every level
  fpn_configs.py:27-78            bifpn_config (restated in the package's fpn_configs.py, which this oracle does not import)
  utils_keras.py:42-82            BatchNormalization inference, epsilon 1e-3

The layers are TensorFlow / Keras kernels (tensorflow==2.10.0, not vendored, not installable offline) => PARITY
UNPINNED; restated with torch-CPU fp32 (conv2d, max_pool2d) and cross-checked against plain NumPy loops
(``pool_same_np`` / ``nearest_np`` below, tests/test_oracle_extra.py).

TF semantics encoded here:
  SAME pooling   out = ceil(in / s); pad_total = max((out-1)*s + k - in, 0), pad_before = pad_total // 2 (the extra cell
                 goes to the end); padded cells never win a max and are not counted by an average
  nearest        align_corners=False, half_pixel_centers=False: src = min(floor(dst * float32(in/out)), in - 1)
"""
import numpy as np
import torch
import torch.nn.functional as tnf

BN_EPS = 1e-3


def bifpn_nodes(min_level, max_level):
    """fpn_configs.py:27-78 (own restatement)"""
    n = max_level - min_level + 1
    ids = {min_level + i: [i] for i in range(n)}
    nxt, nodes = n, []
    for lvl in range(max_level - 1, min_level - 1, -1):
        nodes.append({"feat_level": lvl, "inputs_offsets": [ids[lvl][-1], ids[lvl + 1][-1]]})
        ids[lvl].append(nxt)
        nxt += 1
    for lvl in range(min_level + 1, max_level + 1):
        nodes.append({"feat_level": lvl, "inputs_offsets": list(ids[lvl]) + [ids[lvl - 1][-1]]})
        ids[lvl].append(nxt)
        nxt += 1
    return nodes


def swish(x):
    return x * torch.sigmoid(x)


def batch_norm(x, bn):
    """x [...,F]; inference BN in the order TF evaluates it: (x - mean) * (gamma * rsqrt(var + eps)) + beta"""
    g = torch.from_numpy(np.asarray(bn["gamma"], np.float32))
    b = torch.from_numpy(np.asarray(bn["beta"], np.float32))
    m = torch.from_numpy(np.asarray(bn["mean"], np.float32))
    v = torch.from_numpy(np.asarray(bn["var"], np.float32))
    return (x - m) * (g * torch.rsqrt(v + BN_EPS)) + b


def _same_pad(n, out, s, k):
    total = max((out - 1) * s + k - n, 0)
    return total // 2, total - total // 2


def pool_same(x, th, tw, avg=False):
    """x [B,H,W,C] torch -> [B,th,tw,C]"""
    b, h, w, c = x.shape
    sy, sx = (h - 1) // th + 1, (w - 1) // tw + 1
    ky, kx = sy + 1, sx + 1
    oy, ox = -(-h // sy), -(-w // sx)
    assert (oy, ox) == (th, tw), "SAME pooling does not give the target size"
    (pt, pb), (pl, pr) = _same_pad(h, oy, sy, ky), _same_pad(w, ox, sx, kx)
    t = x.permute(0, 3, 1, 2)
    if avg:
        ones = torch.ones_like(t)
        num = tnf.avg_pool2d(tnf.pad(t, (pl, pr, pt, pb)), (ky, kx), (sy, sx), divisor_override=1)
        den = tnf.avg_pool2d(tnf.pad(ones, (pl, pr, pt, pb)), (ky, kx), (sy, sx), divisor_override=1)
        out = num / den
    else:
        out = tnf.max_pool2d(tnf.pad(t, (pl, pr, pt, pb), value=float("-inf")), (ky, kx), (sy, sx))
    return out.permute(0, 2, 3, 1).contiguous()


def nearest(x, th, tw):
    b, h, w, c = x.shape
    sy, sx = np.float32(h) / np.float32(th), np.float32(w) / np.float32(tw)
    iy = np.minimum(np.floor(np.arange(th, dtype=np.float32) * sy).astype(np.int64), h - 1)
    ix = np.minimum(np.floor(np.arange(tw, dtype=np.float32) * sx).astype(np.int64), w - 1)
    return x[:, torch.from_numpy(iy)][:, :, torch.from_numpy(ix)].contiguous()


def resample(x, th, tw, res, apply_bn=True, conv_after_downsample=False, avg=False, f=None):
    """ResampleFeatureMap.call (efficientdet_keras.py:313-350); x torch [B,H,W,C]"""
    b, h, w, c = x.shape

    def conv(t):
        if c == f:
            return t
        t = t @ torch.from_numpy(np.asarray(res["w"], np.float32)) + torch.from_numpy(np.asarray(res["b"], np.float32))
        if apply_bn and res.get("bn") is not None:
            t = batch_norm(t, res["bn"])
        return t

    if h > th and w > tw:
        if not conv_after_downsample:
            x = conv(x)
        x = pool_same(x, th, tw, avg)
        if conv_after_downsample:
            x = conv(x)
    elif h <= th and w <= tw:
        x = conv(x)
        if h < th or w < tw:
            x = nearest(x, th, tw)
    else:
        raise ValueError("Incompatible Resampling : feat shape {}x{} target_shape: {}x{}".format(h, w, th, tw))
    return x


def fuse(nodes, wsm, method):
    """FNode.fuse_features (efficientdet_keras.py:86-125)"""
    if method in ("fastattn", "channel_fastattn"):
        ew = [torch.relu(torch.from_numpy(np.asarray(v, np.float32))) for v in wsm]
        ws = ew[0]
        for e in ew[1:]:
            ws = ws + e
        out = None
        for x, e in zip(nodes, ew):
            t = x * e / (ws + 0.0001)
            out = t if out is None else out + t
        return out
    if method in ("attn", "channel_attn"):
        ew = torch.stack([torch.from_numpy(np.asarray(v, np.float32)) for v in wsm], -1 if method == "channel_attn" else 0)
        nw = torch.softmax(ew, -1 if method == "channel_attn" else 0)
        st = torch.stack(nodes, -1)
        return torch.sum(st * nw, -1)
    if method == "sum":
        out = nodes[0]
        for x in nodes[1:]:
            out = out + x
        return out
    raise ValueError("unknown weight_method %s" % method)


def sepconv(x, dw, pw, bias):
    """SeparableConv2D 3x3, padding same, depth multiplier 1; x [B,H,W,F]"""
    f = x.shape[-1]
    t = x.permute(0, 3, 1, 2)
    wd = torch.from_numpy(np.asarray(dw, np.float32).reshape(3, 3, f)).permute(2, 0, 1).unsqueeze(1).contiguous()
    t = tnf.conv2d(t, wd, padding=1, groups=f)
    t = t.permute(0, 2, 3, 1) @ torch.from_numpy(np.asarray(pw, np.float32))
    if bias is not None:
        t = t + torch.from_numpy(np.asarray(bias, np.float32))
    return t


def fpn_cells(feats, weights, min_level=3, max_level=7, weight_method="fastattn", apply_bn_for_resampling=True,
              conv_after_downsample=False, conv_bn_act_pattern=False, pooling_type="max"):
    """feats: list[L] of NumPy [B,H_l,W_l,C_l] -> list[L] of NumPy [B,H_l,W_l,F] (FPNCells.call, 787-801)"""
    nodes = bifpn_nodes(min_level, max_level)
    cur = [torch.from_numpy(np.ascontiguousarray(f, np.float32)) for f in feats]
    with torch.no_grad():
        for cell in weights["cells"]:
            cf = list(cur)
            for cfg, w in zip(nodes, cell["fnodes"]):
                f = int(np.shape(w["pw"])[0])
                tgt = cf[cfg["feat_level"] - min_level]
                th, tw = tgt.shape[1], tgt.shape[2]
                ins = [resample(cf[off], th, tw, r, apply_bn_for_resampling, conv_after_downsample, pooling_type == "avg", f)
                       for off, r in zip(cfg["inputs_offsets"], w.get("resample") or [None] * len(cfg["inputs_offsets"]))]
                x = fuse(ins, w.get("wsm"), weight_method)
                if not conv_bn_act_pattern:
                    x = swish(x)
                x = batch_norm(sepconv(x, w["dw"], w["pw"], w.get("b")), w["bn"])
                if conv_bn_act_pattern:
                    x = swish(x)
                cf.append(x)
            cur = []
            for level in range(min_level, max_level + 1):
                for i, cfg in enumerate(reversed(nodes)):
                    if cfg["feat_level"] == level:
                        cur.append(cf[-1 - i])
                        break
    return [c.numpy() for c in cur]


# ---- independent NumPy loops (small inputs): cross-check of the two TF resampling rules -----------------------------
def pool_same_np(x, th, tw, avg=False):
    b, h, w, c = x.shape
    sy, sx = (h - 1) // th + 1, (w - 1) // tw + 1
    ky, kx = sy + 1, sx + 1
    pt = max((th - 1) * sy + ky - h, 0) // 2
    pl = max((tw - 1) * sx + kx - w, 0) // 2
    out = np.zeros((b, th, tw, c), x.dtype)
    for y in range(th):
        for xx in range(tw):
            ys = [v for v in range(y * sy - pt, y * sy - pt + ky) if 0 <= v < h]
            xs = [v for v in range(xx * sx - pl, xx * sx - pl + kx) if 0 <= v < w]
            win = x[:, ys][:, :, xs]
            out[:, y, xx] = win.mean((1, 2)) if avg else win.max((1, 2))
    return out


def nearest_np(x, th, tw):
    b, h, w, c = x.shape
    out = np.zeros((b, th, tw, c), x.dtype)
    for y in range(th):
        for xx in range(tw):
            out[:, y, xx] = x[:, min(int(np.floor(np.float32(y) * (np.float32(h) / np.float32(th)))), h - 1),
                              min(int(np.floor(np.float32(xx) * (np.float32(w) / np.float32(tw)))), w - 1)]
    return out
