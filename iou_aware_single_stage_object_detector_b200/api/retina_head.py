"""RetinaHead -- the plain sibling of IoUawareRetinaHead (mmdet/models/anchor_heads/retina_head.py:10-96,
get_bboxes at anchor_head.py:325-450): same towers, no IoU branch, score = sigmoid(cls).  It reuses the
same kernels with alpha = 1."""
import torch
import torch.nn as nn

from .. import postproc as PP
from .engine_cache import require_cuda
from .iou_aware_retina_head import IoUawareRetinaHead
from .registry import HEADS
from .weight_init import bias_init_with_prob, normal_init


@HEADS.register_module
class RetinaHead(IoUawareRetinaHead):
    def __init__(self, num_classes, in_channels, stacked_convs=4, octave_base_scale=4, scales_per_octave=3,
                 conv_cfg=None, norm_cfg=None, **kwargs):
        super(RetinaHead, self).__init__(num_classes, in_channels, stacked_convs=stacked_convs,
                                         octave_base_scale=octave_base_scale,
                                         scales_per_octave=scales_per_octave, conv_cfg=conv_cfg,
                                         norm_cfg=norm_cfg, **kwargs)
        self.alpha = 1.0

    def _init_layers(self):
        super(RetinaHead, self)._init_layers()
        del self.retina_iou                    # state_dict keys == the reference RetinaHead's

    def init_weights(self):
        for m in self.cls_convs:
            normal_init(m.conv, std=0.01)
        for m in self.reg_convs:
            normal_init(m.conv, std=0.01)
        normal_init(self.retina_cls, std=0.01, bias=bias_init_with_prob(0.01))
        normal_init(self.retina_reg, std=0.01)

    def plan_into(self, eng, sd, F, prefix=""):
        if self.norm_cfg is not None or not self.use_sigmoid_cls:
            raise NotImplementedError("only the sigmoid / no-norm RetinaHead is planned")
        return eng.add_head(sd, F, prefix=prefix, stacked=self.stacked_convs, num_anchors=self.num_anchors,
                            num_classes=self.cls_out_channels, with_iou=False)

    def forward(self, feats):
        cls, reg, _ = super(RetinaHead, self).forward(feats)
        return cls, reg

    def forward_single(self, x):
        c, r = self.forward((x,))
        return c[0], r[0]

    def get_bboxes_device(self, cls_scores, bbox_preds, img_metas, cfg, rescale=False, img_info=None):
        return super(RetinaHead, self).get_bboxes_device(cls_scores, bbox_preds, None, img_metas, cfg,
                                                         rescale, img_info)

    def get_bboxes(self, cls_scores, bbox_preds, gt_bboxes, gt_labels, img_metas, cfg, rescale=False):
        """Signature of this fork's AnchorHead.get_bboxes (anchor_head.py:325-362, gt_* added by WSK)."""
        for t in cls_scores:
            require_cuda(t, "RetinaHead.get_bboxes")
        if not self.in_kernel_envelope([tuple(t.shape[-2:]) for t in cls_scores], cfg):
            return self._get_bboxes_level_by_level(cls_scores, bbox_preds, None, img_metas, cfg, rescale)
        dets, labels, counts = self.get_bboxes_device(cls_scores, bbox_preds, img_metas, cfg, rescale)
        return [(d.clone(), l.clone()) for d, l in PP.split_results(dets, labels, counts)]
