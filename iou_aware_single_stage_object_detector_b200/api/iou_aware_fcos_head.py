"""IoUawareFCOSHead -- the anchor-free sibling (mmdet/models/anchor_heads/iou_aware_fcos_head.py:14-401),
SURVEY.md 8(f) rank 4.

What is on the accelerated path is its ``get_bboxes`` (:229-340): score = sigmoid(cls)^0.3 * sigmoid(iou)^0.7
(alpha hard-coded at :312), per-level top-k of the best class, ``distance2bbox`` from the level's points, clamp,
rescale, multiclass NMS -- the SAME kernels as the RetinaNet path with alpha = 0.3, one "anchor" per cell and the
distance decoder (iou_postproc_cfg.decode_mode = IOU_DECODE_DISTANCE).  The centerness maps are accepted and,
exactly as in the reference (:342-366, the centerness variants are commented out), do not influence the result.

The dense half (towers with GroupNorm, Scale, exp; :92-113) is NOT planned on the tap-GEMM engine: GroupNorm
needs per-sample statistics between the convolutions.  The layers are built so that configs and state_dicts
load; ``forward`` raises.
"""
import torch
import torch.nn as nn

from .. import lib as L
from .. import postproc as PP
from .conv_module import ConvModule
from .engine_cache import require_cuda
from .registry import HEADS
from .weight_init import bias_init_with_prob, normal_init

INF = 1e8


class Scale(nn.Module):
    """mmdet/models/utils/scale.py."""

    def __init__(self, scale=1.0):
        super(Scale, self).__init__()
        self.scale = nn.Parameter(torch.tensor(scale, dtype=torch.float))

    def forward(self, x):
        return x * self.scale


@HEADS.register_module
class IoUawareFCOSHead(nn.Module):
    alpha = 0.3                      # iou_aware_fcos_head.py:312

    def __init__(self, num_classes, in_channels, feat_channels=256, stacked_convs=4, strides=(4, 8, 16, 32, 64),
                 regress_ranges=((-1, 64), (64, 128), (128, 256), (256, 512), (512, INF)), conv_cfg=None,
                 norm_cfg=dict(type='GN', num_groups=32, requires_grad=True)):
        super(IoUawareFCOSHead, self).__init__()
        self.num_classes = num_classes
        self.cls_out_channels = num_classes - 1
        self.in_channels, self.feat_channels, self.stacked_convs = in_channels, feat_channels, stacked_convs
        self.strides, self.regress_ranges = strides, regress_ranges
        self.conv_cfg, self.norm_cfg = conv_cfg, norm_cfg
        self._post = {}
        self._init_layers()

    def _init_layers(self):
        self.cls_convs, self.reg_convs = nn.ModuleList(), nn.ModuleList()
        for i in range(self.stacked_convs):
            chn = self.in_channels if i == 0 else self.feat_channels
            for tower in (self.cls_convs, self.reg_convs):
                tower.append(ConvModule(chn, self.feat_channels, 3, stride=1, padding=1, conv_cfg=self.conv_cfg,
                                        norm_cfg=self.norm_cfg, bias=self.norm_cfg is None))
        self.fcos_cls = nn.Conv2d(self.feat_channels, self.cls_out_channels, 3, padding=1)
        self.fcos_centerness = nn.Conv2d(self.feat_channels, 1, 3, padding=1)
        self.fcos_reg = nn.Conv2d(self.feat_channels, 4, 3, padding=1)
        self.fcos_iou = nn.Conv2d(self.feat_channels, 1, 3, padding=1)
        self.scales = nn.ModuleList([Scale(1.0) for _ in self.strides])

    def init_weights(self):
        for m in list(self.cls_convs) + list(self.reg_convs):
            normal_init(m.conv, std=0.01)
        normal_init(self.fcos_cls, std=0.01, bias=bias_init_with_prob(0.01))
        normal_init(self.fcos_reg, std=0.01)
        normal_init(self.fcos_centerness, std=0.01)
        normal_init(self.fcos_iou, std=0.01)

    def forward(self, feats):
        raise NotImplementedError("IoUawareFCOSHead.forward (GroupNorm towers) is not planned on the tap-GEMM engine; "
                                  "only get_bboxes runs on libiou_b200 (SURVEY.md 8(f) rank 4)")

    # ---- get_bboxes --------------------------------------------------------------------------
    def postproc_workspace(self, featmap_sizes, n_img, cfg, device):
        nms_cfg = dict(cfg['nms'])
        key = (tuple(featmap_sizes), n_img, cfg.get('nms_pre', -1), cfg['score_thr'],
               tuple(sorted(nms_cfg.items())), cfg['max_per_img'], str(device))
        if key not in self._post:
            if nms_cfg.pop('type', 'nms') != 'nms':
                raise NotImplementedError("IoUawareFCOSHead: only nms type 'nms' is wired")
            pcfg = PP.make_cfg(featmap_sizes, self.strides, [torch.zeros(1, 4)] * len(featmap_sizes),
                               self.cls_out_channels, cfg.get('nms_pre', -1), cfg['max_per_img'], cfg['score_thr'],
                               nms_cfg.get('iou_thr', 0.5), alpha=self.alpha, decode_mode=L.DECODE_DISTANCE)
            self._post.clear()
            self._post[key] = PP.PostprocWorkspace(pcfg, n_img, device)
        return self._post[key]

    def get_bboxes_device(self, cls_scores, bbox_preds, ious, img_metas, cfg, rescale=False, img_info=None):
        """Padded device tensors (dets [n,K,5], labels [n,K], counts [n]); nothing synchronises the host."""
        assert len(cls_scores) == len(bbox_preds) == len(ious) == len(self.strides)
        n_img = len(img_metas)
        sizes = [tuple(t.shape[-2:]) for t in cls_scores]
        dev = cls_scores[0].device
        wsp = self.postproc_workspace(sizes, n_img, cfg, dev)
        if img_info is None:
            img_info = PP.make_img_info(img_metas, dev)
        with torch.cuda.device(dev):
            return PP.get_bboxes_device(wsp, cls_scores, bbox_preds, ious, img_info, bool(rescale))

    def get_bboxes(self, cls_scores, bbox_preds, centernesses, ious, gt_bboxes, gt_labels, img_metas, cfg,
                   rescale=None):
        """Same signature / return as the reference (:229-268): list[(Tensor(k,5), Tensor(k,) int64)]."""
        for t in cls_scores:
            require_cuda(t, "IoUawareFCOSHead.get_bboxes")
        dets, labels, counts = self.get_bboxes_device(cls_scores, bbox_preds, ious, img_metas, cfg, rescale)
        return [(d.clone(), l.clone()) for d, l in PP.split_results(dets, labels, counts)]
