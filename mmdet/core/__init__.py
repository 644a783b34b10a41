from iou_aware_single_stage_object_detector_b200.api import (  # noqa: F401
    AnchorGenerator, delta2bbox, bbox2result, multi_apply, multiclass_nms, results2json, det2json, xyxy2xywh)
