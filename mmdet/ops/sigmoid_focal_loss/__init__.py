from iou_aware_single_stage_object_detector_b200.api.ops import (  # noqa: F401
    SigmoidFocalLoss, sigmoid_focal_loss, sigmoid_focal_loss_cuda)
