from iou_aware_single_stage_object_detector_b200.api.datasets import (BboxTransform, DataContainer, bbox_flip,  # noqa: F401
                                                                       prepare_test_img, to_tensor)
from iou_aware_single_stage_object_detector_b200.api.transforms import ImageTransform  # noqa: F401
