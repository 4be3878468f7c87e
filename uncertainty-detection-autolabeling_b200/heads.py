"""Head sampler: the EfficientDet class / box(+sigma) towers x T MC-dropout samples, starting at
the BiFPN outputs (reference efficientdet_keras.py:353-692 ClassNet/BoxNet, 979-1050 MC loop).

Valid replacement for the reference's MC loop when dropout is head-only (mc_dropoutrate == 0, the
configuration every shipped inference YAML uses): the backbone/BiFPN are then deterministic and
their output is identical for all T passes.  With backbone dropout the features differ
per pass, which this entry point does not model - it raises.
"""
import os

import numpy as np

from . import device, utils, _lib
from . import engine as _engine


class HeadSampler:
    def __init__(self, params, weights, device_id=None, heads_mode=None, instance=0):
        if params.get("mc_dropout") and params.get("mc_dropoutrate"):
            raise ValueError("mc_dropoutrate > 0 (backbone MC dropout) changes the BiFPN features per "
                             "sample; the head sampler starts at the BiFPN outputs (head-only dropout)")
        self.params = params
        # every sampler owns its context: the head weights live in the context, so two samplers with the same
        # params but different checkpoints (ensembles, model comparison, reloads) must never share one.  The
        # config-keyed cache of engine.get_engine serves the weight-free postprocess.* entry points only.
        if device_id is None:
            device_id = params.get("device", 0) or 0
        self.engine = _engine.Engine(params, device_id, heads_mode or params.get("heads_mode", "fp32"))
        self.engine.set_head_weights(weights)
        # the reference draws fresh dropout masks on every call: seed=None continues a per-sampler random stream
        self._seed_base = int.from_bytes(os.urandom(8), "little")
        self._calls = 0

    def _seed(self, seed):
        if seed is not None:
            return int(seed) & ((1 << 64) - 1)
        self._calls += 1
        return utils.mix_seed(self._seed_base, self._calls)

    def close(self):
        self.engine.ctx.close()

    def __call__(self, fpn_feats, masks=None, seed=None):
        """fpn_feats: list[L] of [B,H_l,W_l,F].  Returns (cls_outputs, box_outputs) with the
        structure of EfficientDetNet.call: list[L] of [T,B,H,W,A*C] and [T,B,H,W,8A] (no leading T
        for a head without MC dropout).  ``masks`` [T,2,L,R,B,F] uint8 injects keep masks (parity
        runs); otherwise Philox4x32-10 masks are drawn on the device from ``seed``."""
        host = not any(isinstance(x, device.DeviceArray) or hasattr(x, "__cuda_array_interface__") for x in fpn_feats)
        cls, box = self.engine.heads_sample(list(fpn_feats), masks, self._seed(seed))
        if host:
            cls = [c.copy_to_host(sync=False) for c in cls]
            box = [b.copy_to_host(sync=False) for b in box]
            self.engine.ctx.sync()
        return cls, box

    def detect(self, fpn_feats, image_scales=None, masks=None, seed=None):
        """features -> detections in one device call (udal_run); tuple like postprocess_*."""
        host = not any(isinstance(x, device.DeviceArray) or hasattr(x, "__cuda_array_interface__") for x in fpn_feats)
        bufs = self.engine.run(list(fpn_feats), image_scales, masks, self._seed(seed))
        out = [bufs["boxes"], bufs["scores"], bufs["classes"], bufs["valid"], bufs["logits"]]
        if host:
            out = [o.copy_to_host(sync=False) for o in out]
            self.engine.ctx.sync()
        return tuple(out)


class PipelinedSampler:
    """Throughput front end for host-resident batches: ``depth`` independent contexts on one GPU
    (own stream, scratch and staging each), one host thread per context.  While context k runs the
    kernels of batch i, context k+1 is already copying batch i+1 host->device, so the PCIe copies
    disappear behind the compute.  ``map(batches)`` yields the detection tuples in input order."""

    def __init__(self, params, weights, device_id=None, heads_mode=None, depth=2):
        self.samplers = [HeadSampler(params, weights, device_id, heads_mode, instance=i) for i in range(depth)]
        self._seed_base = int.from_bytes(os.urandom(8), "little")
        self._maps = 0

    def map(self, batches, image_scales=None, seed=None):
        import concurrent.futures as cf

        depth = len(self.samplers)
        if seed is None:
            self._maps += 1
            seed = utils.mix_seed(self._seed_base, self._maps)
        with cf.ThreadPoolExecutor(max_workers=depth) as pool:
            pending = []
            for i, feats in enumerate(batches):
                sc = None if image_scales is None else image_scales[i]
                pending.append(pool.submit(self.samplers[i % depth].detect, feats, sc, None, batch_seed(seed, i)))
                if len(pending) >= depth:
                    yield pending.pop(0).result()
            for f in pending:
                yield f.result()


class StreamingSampler:
    """Throughput front end for host-resident batches on ONE context: the GPU runs exactly the sequence of back-to-back
    ``udal_run`` calls (heads of batch i+1 over the NMS tail of batch i), the feature maps of the next batches are uploaded
    on a copy stream into ``depth`` device buffer sets meanwhile, and the detections of a batch are fetched into pinned
    host memory right behind its own tail (udal_stage_* / udal_fetch_* of include/udal.h).  ``map(batches)`` yields the
    detection tuples (NumPy) in input order, ``depth - 1`` batches behind the submission.

    Host arrays should live in pinned memory (device.PinnedArray) for the uploads to be asynchronous; float32 maps or -
    fp16 heads - float16 maps (the reference's mixed_float16 exports)."""

    def __init__(self, params, weights, device_id=None, heads_mode=None, depth=3):
        if not 2 <= depth <= _lib.STAGE_SLOTS:
            raise ValueError("depth must be 2..%d" % _lib.STAGE_SLOTS)
        self.sampler = HeadSampler(params, weights, device_id, heads_mode)
        self.engine = self.sampler.engine
        self.depth = depth
        self._slots = [None] * depth      # per slot: dict(key=(batch, dtype), feats=[DeviceArray], scales=DeviceArray, host={name: PinnedArray})
        self._seed_base = int.from_bytes(os.urandom(8), "little")
        self._maps = 0

    def close(self):
        self._slots = [None] * self.depth
        self.sampler.close()

    def _slot(self, i, batch, dtype):
        eng = self.engine
        sl = self._slots[i]
        if sl is None or sl["key"] != (batch, dtype):
            sl = dict(key=(batch, dtype),
                      feats=[eng.ctx.empty((batch, h, w, eng.F), dtype) for h, w in eng.level_hw],
                      scales=eng.ctx.empty((batch,), np.float32), scales_host=device.PinnedArray((batch,), np.float32), host=None)
            self._slots[i] = sl
        return sl

    def map(self, batches, image_scales=None, seed=None):
        eng, lib, h = self.engine, self.engine.lib, self.engine.ctx.handle
        if seed is None:
            self._maps += 1
            seed = utils.mix_seed(self._seed_base, self._maps)
        pending = []

        def finish(slot):
            _lib.check(lib.udal_fetch_wait(h, slot))
            host = self._slots[slot]["host"]
            return tuple(np.array(host[k].array, copy=True) for k in ("boxes", "scores", "classes", "valid", "logits"))

        for i, feats in enumerate(batches):
            slot = i % self.depth
            if len(pending) >= self.depth:      # the slot's previous results must be out of its pinned buffers
                yield finish(pending.pop(0))
            feats = [np.ascontiguousarray(f) for f in feats]
            dt = np.dtype(np.float16) if feats[0].dtype == np.float16 else np.dtype(np.float32)
            batch = feats[0].shape[0]
            sl = self._slot(slot, batch, dt)
            # ---- upload on the copy stream ----
            _lib.check(lib.udal_stage_begin(h, slot))
            for dst, src in zip(sl["feats"], feats):
                if src.dtype != dt or src.shape != dst.shape:
                    raise ValueError("feature maps: expected %s %s, got %s %s" % (dt, dst.shape, src.dtype, src.shape))
                _lib.check(lib.udal_stage_h2d(h, dst.ptr, src.ctypes.data, dst.nbytes))
            sc = None
            if image_scales is not None:
                sl["scales_host"].array[...] = np.asarray(image_scales[i], np.float32)
                _lib.check(lib.udal_stage_h2d(h, sl["scales"].ptr, sl["scales_host"].ptr, sl["scales"].nbytes))
                sc = sl["scales"]
            _lib.check(lib.udal_stage_end(h, slot))
            # ---- the run, ordered behind its uploads; its inputs are free again once the heads are through ----
            _lib.check(lib.udal_stage_acquire(h, slot))
            bufs = eng.run(sl["feats"], sc, None, batch_seed(seed, i))
            _lib.check(lib.udal_stage_release(h, slot))
            # ---- results -> pinned host memory, behind this run's tail ----
            if sl["host"] is None or any(sl["host"][k].shape != bufs[k].shape for k in bufs):
                sl["host"] = {k: device.PinnedArray(v.shape, v.dtype) for k, v in bufs.items()}
            for k, v in bufs.items():
                _lib.check(lib.udal_fetch_d2h(h, sl["host"][k].ptr, v.ptr, v.nbytes))
            _lib.check(lib.udal_fetch_mark(h, slot))
            sl["bufs"] = bufs                   # keep the device buffers alive until the fetch has run
            pending.append(slot)
        while pending:
            yield finish(pending.pop(0))


def batch_seed(seed, i):
    """seed of batch / shard ``i`` of a call seeded with ``seed`` (see utils.mix_seed)"""
    return utils.mix_seed(seed, i)


def philox_keep_masks(shape, rate_class, rate_box, seed):
    """NumPy restatement of the in-kernel Philox4x32-10 keep masks (counter = flat index // 4,
    key = seed; u = (x >> 8) * 2^-24; keep = u >= rate).  shape = (T,2,L,R,B,F)."""
    total = int(np.prod(shape))
    g = np.arange((total + 3) // 4, dtype=np.uint64)
    c = [g & np.uint64(0xFFFFFFFF), g >> np.uint64(32), np.zeros_like(g), np.zeros_like(g)]
    k0 = np.uint64(seed & 0xFFFFFFFF)
    k1 = np.uint64((seed >> 32) & 0xFFFFFFFF)
    m0, m1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0, p1 = m0 * c[0], m1 * c[2]
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & mask, p1 >> np.uint64(32), p1 & mask
        c = [hi1 ^ c[1] ^ k0, lo1, hi0 ^ c[3] ^ k1, lo0]
        k0 = (k0 + np.uint64(0x9E3779B9)) & mask
        k1 = (k1 + np.uint64(0xBB67AE85)) & mask
    words = np.stack(c, axis=1).reshape(-1)[:total]
    u = (words >> np.uint64(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)
    u = u.reshape(shape)
    keep = np.empty(shape, np.uint8)
    keep[:, 0] = u[:, 0] >= np.float32(rate_class)
    keep[:, 1] = u[:, 1] >= np.float32(rate_box)
    return keep
