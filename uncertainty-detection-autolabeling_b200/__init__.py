"""udal-b200: B200-native (sm_100a) uncertainty sampling + post-processing for EfficientDet
auto-labeling - the hot path of continental/uncertainty-detection-autolabeling behind the
reference's own Python entry points.

    import importlib; udal = importlib.import_module("uncertainty-detection-autolabeling_b200")
    (or ``import udal_b200`` - the alias module at the repository root)

Modules mirror the reference: ``postprocess``, ``anchors``, ``nms_np``, ``utils_box``,
``utils_extra``, ``hparams_config``, ``utils``; ``heads`` and ``scheduler`` are the new entry
points for the head sampler and the multi-GPU image scheduler, ``autolabel`` the calibrated-uncertainty /
auto-label threshold pass of the InferImages loop, ``bifpn`` / ``fpn_configs`` the BiFPN
(``FPNCells``) that produces the head sampler's input.  Everything computes on the GPU
through ``libudal.so``; importing fails loudly when the library is missing (no CPU fallback).
"""
from . import _lib

_lib.load()  # fail loudly at import time if the CUDA library is missing

from . import anchors, autolabel, bifpn, device, engine, fpn_configs, heads, hparams_config, nms_np, postprocess, scheduler, serving, synthetic, utils, utils_box, utils_class, utils_extra, wire  # noqa: E402,F401

__all__ = ["anchors", "autolabel", "bifpn", "fpn_configs", "device", "engine", "heads", "hparams_config", "nms_np", "postprocess", "synthetic",
           "scheduler", "serving", "utils", "utils_box", "utils_class", "utils_extra", "wire"]
