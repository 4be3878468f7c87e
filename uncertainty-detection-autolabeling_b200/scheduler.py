"""Image-parallel batch scheduler (north_star item 4).

The hot path is independent per image (per-image top-k and NMS; the head weights are ~0.2 MB and
replicated), so scaling is pure sharding of the image batch: contiguous chunks of
``ceil(B / G)`` images per GPU, **no collective on the data path**; detections
``[B_g,100,...]`` are copied device->host per shard and concatenated in input order.

Two front ends share the same shard arithmetic:
  * ``ImageScheduler``   one process, one host thread + one libudal context per visible GPU;
  * ``run_sharded``      one process per GPU (torchrun / ``torch.distributed``): every rank takes
                         its shard, results are optionally gathered to rank 0 with
                         ``gather_object`` (host-side, after the device work; works with gloo).
"""
import threading

import numpy as np


def shard_ranges(batch, world):
    """[(start, stop)] of length ``world``: contiguous chunks of ceil(batch/world) images;
    trailing shards may be shorter or empty."""
    if world < 1:
        raise ValueError("world must be >= 1")
    per = -(-batch // world) if batch > 0 else 0
    return [(min(r * per, batch), min((r + 1) * per, batch)) for r in range(world)]


def take_shard(levels, start, stop):
    """Slice every per-level array [B,...] to images [start, stop)."""
    return [x[start:stop] for x in levels]


def concat_results(parts):
    """list (per shard, in shard order) of tuples of arrays -> tuple of concatenated arrays;
    empty shards (None) are skipped."""
    parts = [p for p in parts if p is not None]
    if not parts:
        return None
    return tuple(np.concatenate([np.asarray(p[i]) for p in parts], axis=0) for i in range(len(parts[0])))


def run_sharded(fn, levels, batch, rank, world, extra=None, gather=True, group=None):
    """One-process-per-GPU front end.  ``fn(shard_levels, shard_extra)`` processes the local shard
    and returns a tuple of host arrays with the image axis first.  ``extra``: per-image host arrays
    (e.g. image_scales) sliced alongside.  With ``gather`` rank 0 returns the full result in input
    order (other ranks None); without it every rank returns its own shard."""
    start, stop = shard_ranges(batch, world)[rank]
    local = None
    if stop > start:
        ex = None if extra is None else [e[start:stop] for e in extra]
        local = fn(take_shard(levels, start, stop), ex)
    if not gather or world == 1:
        return local
    import torch.distributed as dist

    out = [None] * world if rank == 0 else None
    dist.gather_object(local, out, dst=0, group=group)
    return concat_results(out) if rank == 0 else None


class ImageScheduler:
    """Single-process scheduler: shards each call over ``devices`` (default: all visible GPUs),
    one host thread and one context (engine + head weights) per GPU."""

    def __init__(self, params, weights, devices=None, heads_mode=None):
        from . import _lib, heads

        n = _lib.device_count()
        if n == 0:
            raise RuntimeError("no CUDA device visible: the scheduler has no CPU fallback")
        self.devices = list(range(n)) if devices is None else list(devices)
        self.samplers = [heads.HeadSampler(params, weights, device_id=d, heads_mode=heads_mode)
                         for d in self.devices]

    def detect(self, fpn_feats, image_scales=None, masks=None, seed=None):
        """fpn_feats: list[L] of host arrays [B,H_l,W_l,F].  Returns the detection tuple of
        ``HeadSampler.detect`` for the whole batch, in input order."""
        from . import heads, utils

        batch = fpn_feats[0].shape[0]
        if seed is None:  # a fresh draw per call, like the reference's stateful dropout layers
            import os
            seed = int.from_bytes(os.urandom(8), "little")
        ranges = shard_ranges(batch, len(self.devices))
        results = [None] * len(self.devices)
        errors = []

        def work(i):
            start, stop = ranges[i]
            if stop <= start:
                return
            try:
                sc = None if image_scales is None else np.asarray(image_scales)[start:stop]
                mk = None if masks is None else np.ascontiguousarray(np.asarray(masks)[:, :, :, :, start:stop])
                results[i] = self.samplers[i].detect(take_shard(fpn_feats, start, stop), sc, masks=mk,
                                                     seed=heads.batch_seed(seed, i))
            except Exception as e:  # surfaced after the join
                errors.append(e)

        threads = [threading.Thread(target=work, args=(i,)) for i in range(len(self.devices))]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if errors:
            raise errors[0]
        return concat_results(results)
