from iou_aware_single_stage_object_detector_b200.dist import init_dist  # noqa: F401
