"""IoUawareFCOSHead -- the anchor-free sibling (mmdet/models/anchor_heads/iou_aware_fcos_head.py:14-401),
SURVEY.md 8(f) rank 4.

What is on the accelerated path is its ``get_bboxes`` (:229-340): score = sigmoid(cls)^0.3 * sigmoid(iou)^0.7
(alpha hard-coded at :312), per-level top-k of the best class, ``distance2bbox`` from the level's points, clamp,
rescale, multiclass NMS -- the SAME kernels as the RetinaNet path with alpha = 0.3, one "anchor" per cell and the
distance decoder (iou_postproc_cfg.decode_mode = IOU_DECODE_DISTANCE).  The centerness maps are accepted and,
exactly as in the reference (:342-366, the centerness variants are commented out), do not influence the result.

The dense half (:92-113) runs on the conv engine too: every tower layer is a tap-GEMM (no bias) followed by an
in-place GroupNorm+ReLU pass over all levels (iou_group_norm_relu), fcos_cls + fcos_centerness share one GEMM with
a split store, fcos_reg + fcos_iou likewise, and bbox_pred = exp(scale_l * fcos_reg) is one small kernel per level.
"""
import torch
import torch.nn as nn

from .. import engine as E
from .. import lib as L
from .. import postproc as PP
from .conv_module import ConvModule
from .engine_cache import PlanCache, cuda_state_dict, param_stamp, require_cuda
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
        self._plans = PlanCache()
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

    # ---- forward -----------------------------------------------------------------------------
    def plan_into(self, eng, sd, F, prefix=""):
        if self.norm_cfg is None or self.norm_cfg.get('type') != 'GN':
            raise NotImplementedError("IoUawareFCOSHead is planned with its default GroupNorm towers only")
        if self.conv_cfg is not None and self.conv_cfg.get('type', 'Conv') != 'Conv':
            raise NotImplementedError("conv type %r is outside the accelerated path" % self.conv_cfg.get('type'))
        gn = self.cls_convs[0].gn
        return eng.add_fcos_head(sd, F, prefix=prefix, stacked=self.stacked_convs,
                                 num_classes=self.cls_out_channels, groups=gn.num_groups, eps=gn.eps)

    def forward(self, feats):
        """feats: tuple of (N, C, H, W) CUDA tensors -> (cls_scores, bbox_preds, centernesses, ious), lists per
        level with logical shapes (N, 80, H, W), (N, 4, H, W), (N, 1, H, W), (N, 1, H, W) (:89-113); NHWC storage."""
        for t in feats:
            require_cuda(t, "IoUawareFCOSHead.forward")
        assert len(feats) == len(self.strides)
        feats = [t.float().contiguous() for t in feats]
        dev = feats[0].device
        key = (tuple(tuple(t.shape) for t in feats), dev, param_stamp(self))

        def build():
            eng = E.Engine(dev)
            ins = [torch.empty_like(t) for t in feats]
            F = eng.new_map([(t.shape[0], t.shape[2], t.shape[3]) for t in ins], ins[0].shape[1])
            for s_, t in enumerate(ins):
                n, c, h, w = t.shape
                rs = F.segs[s_][0]
                lib, tp, fp = eng.lib, t.data_ptr(), F.ptr
                eng.ops.append(("pack", lambda st, tp=tp, n=n, c=c, h=h, w=w, rs=rs, fp=fp, lib=lib:
                                E.L.check(lib.iou_pack_nchw(tp, n, c, h, w, fp, rs, st))))
            outs = self.plan_into(eng, cuda_state_dict(self, dev), F)
            return eng, ins, outs
        eng, ins, outs = self._plans.get(key, build)
        for a_, b_ in zip(ins, feats):
            a_.copy_(b_)
        with torch.cuda.device(dev):
            eng.run()
        # fresh tensors (the plan's own maps are overwritten by the next call)
        return tuple([t.clone(memory_format=torch.preserve_format) for t in ts] for ts in outs)

    @staticmethod
    def postproc_inputs(outs):
        """(cls_scores, bbox_preds, ious) out of forward()'s (cls, bbox_pred, centerness, iou): the centerness maps
        do not enter get_bboxes (iou_aware_fcos_head.py:342-366)."""
        return outs[0], outs[1], outs[3]

    def forward_single(self, x, scale=None):
        if len(self.strides) != 1:
            raise NotImplementedError("forward_single needs the level's Scale: call forward(feats) with all levels")
        c, r, q, u = self.forward((x,))
        return c[0], r[0], q[0], u[0]

    # ---- get_bboxes --------------------------------------------------------------------------
    def postproc_cfg(self, featmap_sizes, cfg):
        """(iou_postproc_cfg, None) for a test_cfg, without allocating a workspace."""
        nms_cfg = dict(cfg['nms'])
        if nms_cfg.pop('type', 'nms') != 'nms':
            raise NotImplementedError("IoUawareFCOSHead: only nms type 'nms' is wired")
        pcfg = PP.make_cfg(featmap_sizes, self.strides, [torch.zeros(1, 4)] * len(featmap_sizes),
                           self.cls_out_channels, cfg.get('nms_pre', -1), cfg['max_per_img'], cfg['score_thr'],
                           nms_cfg.get('iou_thr', 0.5), alpha=self.alpha, decode_mode=L.DECODE_DISTANCE)
        return pcfg, None

    def postproc_workspace(self, featmap_sizes, n_img, cfg, device):
        key = (tuple(featmap_sizes), n_img, cfg.get('nms_pre', -1), cfg['score_thr'],
               tuple(sorted(dict(cfg['nms']).items())), cfg['max_per_img'], str(device))
        if key not in self._post:
            pcfg, _ = self.postproc_cfg(featmap_sizes, cfg)
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
