from iou_aware_single_stage_object_detector_b200.api.ops import (  # noqa: F401
    nms, soft_nms, sigmoid_focal_loss, SigmoidFocalLoss)
