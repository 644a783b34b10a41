"""Host-side mirror of the reference's model / core / ops interface for the hot path."""
from .registry import Registry, BACKBONES, NECKS, ROI_EXTRACTORS, SHARED_HEADS, HEADS, LOSSES, DETECTORS
from .builder import (build, build_backbone, build_neck, build_roi_extractor, build_shared_head,
                      build_head, build_loss, build_detector)
from .config import Config, ConfigDict
from .anchor_generator import AnchorGenerator
from .transforms import delta2bbox, bbox2result, multi_apply, ImageTransform
from .bbox_nms import multiclass_nms
from .results import batch_bbox2result, xyxy2xywh, det2json, results2json, dump_results
from .losses import FocalLoss, SmoothL1Loss, CrossEntropyLoss
from .conv_module import ConvModule, build_conv_layer, build_norm_layer
from .resnet import ResNet, ResNeXt, Bottleneck, make_res_layer
from .fpn import FPN
from .anchor_head import AnchorHead
from .iou_aware_retina_head import IoUawareRetinaHead
from .retina_head import RetinaHead
from .iou_aware_fcos_head import IoUawareFCOSHead
from .detectors import BaseDetector, SingleStageDetector, RetinaNet, FCOS, FusedPlan
from .ops import (nms, soft_nms, sigmoid_focal_loss, SigmoidFocalLoss, nms_cuda, nms_cpu, soft_nms_cpu,
                  sigmoid_focal_loss_cuda)
from .datasets import BboxTransform, DataContainer, bbox_flip, prepare_test_img, to_tensor
from .engine_cache import invalidate_plans
