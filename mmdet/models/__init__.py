from iou_aware_single_stage_object_detector_b200.api import (  # noqa: F401
    BACKBONES, NECKS, ROI_EXTRACTORS, SHARED_HEADS, HEADS, LOSSES, DETECTORS, build_backbone, build_neck,
    build_roi_extractor, build_shared_head, build_head, build_loss, build_detector, ResNet, ResNeXt, FPN,
    AnchorHead, IoUawareRetinaHead, BaseDetector, SingleStageDetector, RetinaNet, FocalLoss, SmoothL1Loss,
    ConvModule)
