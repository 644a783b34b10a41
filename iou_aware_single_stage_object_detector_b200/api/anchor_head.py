"""AnchorHead base with the reference's constructor (mmdet/models/anchor_heads/anchor_head.py:37-103)."""
import torch.nn as nn

from .anchor_generator import AnchorGenerator
from .builder import build_loss
from .registry import HEADS
from .transforms import multi_apply
from .weight_init import normal_init


@HEADS.register_module
class AnchorHead(nn.Module):
    def __init__(self, num_classes, in_channels, feat_channels=256, anchor_scales=[8, 16, 32],
                 anchor_ratios=[0.5, 1.0, 2.0], anchor_strides=[4, 8, 16, 32, 64], anchor_base_sizes=None,
                 target_means=(.0, .0, .0, .0), target_stds=(1.0, 1.0, 1.0, 1.0),
                 loss_cls=dict(type='CrossEntropyLoss', use_sigmoid=True, loss_weight=1.0),
                 loss_bbox=dict(type='SmoothL1Loss', beta=1.0 / 9.0, loss_weight=1.0)):
        super(AnchorHead, self).__init__()
        self.in_channels, self.num_classes, self.feat_channels = in_channels, num_classes, feat_channels
        self.anchor_scales, self.anchor_ratios, self.anchor_strides = anchor_scales, anchor_ratios, anchor_strides
        self.anchor_base_sizes = list(anchor_strides) if anchor_base_sizes is None else anchor_base_sizes
        self.target_means, self.target_stds = target_means, target_stds
        self.use_sigmoid_cls = loss_cls.get('use_sigmoid', False)
        self.sampling = loss_cls['type'] not in ['FocalLoss', 'GHMC', 'IOUbalancedSigmoidFocalLoss']
        self.cls_out_channels = num_classes - 1 if self.use_sigmoid_cls else num_classes
        self.loss_cls = build_loss(loss_cls)
        self.loss_bbox = build_loss(loss_bbox)
        self.anchor_generators = [AnchorGenerator(b, anchor_scales, anchor_ratios)
                                  for b in self.anchor_base_sizes]
        self.num_anchors = len(self.anchor_ratios) * len(self.anchor_scales)
        self._init_layers()

    def _init_layers(self):
        self.conv_cls = nn.Conv2d(self.feat_channels, self.num_anchors * self.cls_out_channels, 1)
        self.conv_reg = nn.Conv2d(self.feat_channels, self.num_anchors * 4, 1)

    def init_weights(self):
        normal_init(self.conv_cls, std=0.01)
        normal_init(self.conv_reg, std=0.01)

    def forward_single(self, x):
        raise NotImplementedError("the plain AnchorHead (RPN-style 1x1 head) is not on the accelerated path")

    def forward(self, feats):
        return multi_apply(self.forward_single, feats)
