"""Mirror of ``src/utils_extra.py`` hot-path symbols: stack_mcpred :201-217, get_mcuncert
:220-244, mc_eval :142-198 (the T-loop wrapper, here a single in-kernel T loop)."""
import numpy as np

from . import engine as _engine
from .utils_box import _GENERIC


def stack_mcpred(output):
    """utils_extra.py:201-217: list[5] of list[T] of host arrays -> list[5] of [T,...]."""
    o0, o1, o2, o3, o4 = output
    return [np.stack([np.asarray(x) for x in o], axis=0) for o in (o0, o1, o2, o3, o4)]


_engines = {}  # (T, level sizes, device) -> cached engine: one CUDA context per shape, not one per call


def get_mcuncert(output, device_id=0):
    """utils_extra.py:220-244: per level mean and population std over the leading (sample) axis,
    computed by the moments kernel (sequential fp32 sum / T, two-pass std).  Host arrays in/out
    (device tensors go through postprocess.*, which fuses this reduction)."""
    o0, o1, o2, o3, o4 = output
    levels = [np.asarray(x, np.float32) for x in (o0, o1, o2, o3, o4)]
    T, batch = levels[0].shape[0], levels[0].shape[1]
    sizes = [int(np.prod(x.shape[2:], dtype=np.int64)) for x in levels]
    key = (T, tuple(sizes), device_id)
    eng = _engines.get(key)
    if eng is None:
        if len(_engines) >= 8:  # bounded: drop the oldest shape
            _engines.pop(next(iter(_engines))).ctx.close()
        params = dict(_GENERIC, max_level=len(levels) - 1, loss_attenuation=False, mc_dropout=True,
                      mc_classheadrate=0.5, mc_dropoutsamp=T, uncert_adjust_method="l-norm")
        eng = _engines[key] = _engine.Engine(params, device_id, level_hw=[(1, n) for n in sizes],
                                             anchors=np.zeros((sum(sizes), 4), np.float32))
    cls = [eng.ctx.to_device(x.reshape(T, batch, 1, n, 1)) for x, n in zip(levels, sizes)]
    box = [eng.ctx.zeros((batch, 1, n, 4)) for n in sizes]
    out = eng.decode_moments(cls, box, batch, want=("mean_logits", "std_logits"))
    m, s = out["mean_logits"].numpy(), out["std_logits"].numpy()
    mean, std, off = [], [], 0
    for x, n in zip(levels, sizes):
        mean.append(m[:, off:off + n, 0].reshape(x.shape[1:]))
        std.append(s[:, off:off + n, 0].reshape(x.shape[1:]))
        off += n
    return mean, std


def mc_eval(mc_model, images, config=None):
    """utils_extra.py:142-198.  ``mc_model`` is a ``heads.HeadSampler`` (or any callable returning
    (cls_outputs, box_outputs) already stacked over T): the T forward passes of the reference
    collapse into one call whose kernel loops over the samples."""
    cls, box = mc_model(images)
    return [cls, box]
