from iou_aware_single_stage_object_detector_b200.api.ops import nms, soft_nms, nms_cuda, nms_cpu, soft_nms_cpu  # noqa: F401
