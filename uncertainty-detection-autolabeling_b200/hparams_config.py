"""The hot-path slice of the reference's ``src/hparams_config.py`` restated as plain data.

hparams_config.py:183-370 (default_detection_configs: uncertainty 193-209, anchors 277-282,
heads 322-328, nms_configs 332-340) and :373-452 (per-model table; fpn_num_filters /
box_class_repeats of efficientdet-d0..d7).  ``params`` everywhere in this package is the same
plain dict the reference passes around (``Config.as_dict()``).
"""
import copy

MODEL_TABLE = {  # name -> (image_size, fpn_num_filters, fpn_cell_repeats, box_class_repeats)
    "efficientdet-d0": (512, 64, 3, 3),
    "efficientdet-d1": (640, 88, 4, 3),
    "efficientdet-d2": (768, 112, 5, 3),
    "efficientdet-d3": (896, 160, 6, 4),
    "efficientdet-d4": (1024, 224, 7, 4),
    "efficientdet-d5": (1280, 288, 7, 4),
    "efficientdet-d6": (1280, 384, 8, 5),
    "efficientdet-d7": (1536, 384, 8, 5),
}


def default_detection_configs():
    return dict(
        name="efficientdet-d0", image_size=512, num_classes=90, data_format="channels_last",
        # uncertainty (hparams_config.py:193-209)
        enable_softmax=False, loss_attenuation=False, uncert_adjust_method="l-norm",
        decode_nsamples=100, mc_dropout=False, mc_dropoutrate=0.0, mc_classheadrate=0.0,
        mc_boxheadrate=0.0, mc_dropoutsamp=10,
        # anchors / heads
        min_level=3, max_level=7, num_scales=3, aspect_ratios=[1.0, 2.0, 0.5], anchor_scale=4.0,
        box_class_repeats=3, fpn_num_filters=64, separable_conv=True, act_type="swish",
        survival_prob=None,
        # BiFPN (hparams_config.py:322-328, 344-346)
        fpn_cell_repeats=3, apply_bn_for_resampling=True, conv_after_downsample=False, conv_bn_act_pattern=False,
        fpn_name=None, fpn_weight_method=None, fpn_config=None,
        # nms (hparams_config.py:332-340)
        nms_configs=dict(method="gaussian", iou_thresh=None, score_thresh=0.0, sigma=None,
                         pyfunc=False, max_nms_inputs=0, max_output_size=100),
        # B200 build additions (not in the reference)
        tf_nms_variant="new", heads_mode="fp32",
    )


def get_detection_config(model_name="efficientdet-d0", **overrides):
    """Config dict for a model name plus overrides (``nms_configs`` merges key-wise)."""
    p = default_detection_configs()
    size, filters, _, repeats = MODEL_TABLE[model_name]
    _, _, cells, _ = MODEL_TABLE[model_name]
    p.update(name=model_name, image_size=size, fpn_num_filters=filters, box_class_repeats=repeats, fpn_cell_repeats=cells)
    if model_name in ("efficientdet-d6", "efficientdet-d7"):
        p["fpn_weight_method"] = "sum"  # hparams_config.py:429,439
    overrides = copy.deepcopy(overrides)
    nms = overrides.pop("nms_configs", None)
    p.update(overrides)
    if nms:
        p["nms_configs"] = dict(p["nms_configs"], **nms)
    return p
