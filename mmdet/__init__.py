"""Compatibility namespace: lets code written against the reference's import paths
(``from mmdet.models import build_detector``, ``from mmdet.ops import nms`` ...) run on the
B200-native implementation in ``iou_aware_single_stage_object_detector_b200``."""
__version__ = "0.6.0+b200"
