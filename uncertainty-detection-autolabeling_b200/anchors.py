"""Mirror of the reference's ``src/anchors.py`` hot-path symbols.

anchors.py:100-218 ``Anchors`` (table built on the host in float64 exactly like the reference,
see engine.anchor_table), :41-75 ``decode_box_outputs`` (device kernel), :38
``MAX_DETECTION_POINTS``.  ``AnchorLabeler`` (training target assignment) is out of scope.
"""
import numpy as np

from . import engine as _engine
from . import utils

MAX_DETECTION_POINTS = 5000


class Anchors:
    """Multi-scale anchors (anchors.py:100-218)."""

    def __init__(self, min_level, max_level, num_scales, aspect_ratios, anchor_scale, image_size):
        self.min_level = min_level
        self.max_level = max_level
        self.num_scales = num_scales
        self.aspect_ratios = aspect_ratios
        if isinstance(anchor_scale, (list, tuple)):
            assert len(anchor_scale) == max_level - min_level + 1
            self.anchor_scales = anchor_scale
        else:
            self.anchor_scales = [anchor_scale] * (max_level - min_level + 1)
        self.image_size = utils.parse_image_size(image_size)
        self.feat_sizes = utils.get_feat_sizes(image_size, max_level)
        self.boxes = _engine.anchor_table(min_level, max_level, num_scales, aspect_ratios,
                                          anchor_scale, image_size)

    def get_anchors_per_location(self):
        return self.num_scales * len(self.aspect_ratios)


def decode_box_outputs(pred_boxes, anchor_boxes, device_id=0):
    """anchors.py:41-75 - plain exp/offset decode in fp32 on the device.

    pred_boxes [..., 4] (ty,tx,th,tw), anchor_boxes [..., 4] broadcastable on the leading axes."""
    from . import utils_box
    return utils_box._decode(pred_boxes, None, anchor_boxes, "plain", device_id)
