"""IoUawareRetinaHead with the reference's constructor, state_dict keys and forward/get_bboxes
signatures (mmdet/models/anchor_heads/iou_aware_retina_head.py:64-219,390-564).

forward   : one conv-engine plan for ALL levels (weights are shared across levels, so each of the
            8 tower convs + 2 output GEMMs is a single persistent tcgen05 launch over every level).
get_bboxes: five kernels for the whole batch (max-score, top-k, gather/decode, class NMS, final
            select) instead of per-image / per-level / per-class Python loops.
"""
import numpy as np
import torch
import torch.nn as nn

from .. import engine as E
from .. import postproc as PP
from .anchor_head import AnchorHead
from .conv_module import ConvModule
from .engine_cache import PlanCache, cuda_state_dict, param_stamp, require_cuda
from .registry import HEADS
from .weight_init import bias_init_with_prob, normal_init


@HEADS.register_module
class IoUawareRetinaHead(AnchorHead):
    def __init__(self, num_classes, in_channels, stacked_convs=4, octave_base_scale=4, scales_per_octave=3,
                 conv_cfg=None, norm_cfg=None,
                 loss_iou=dict(type='GHMIoU', bins=30, momentum=0.75, use_sigmoid=True, loss_weight=1.0),
                 **kwargs):
        self.stacked_convs = stacked_convs
        self.octave_base_scale, self.scales_per_octave = octave_base_scale, scales_per_octave
        self.conv_cfg, self.norm_cfg = conv_cfg, norm_cfg
        octave_scales = np.array([2 ** (i / scales_per_octave) for i in range(scales_per_octave)])
        super(IoUawareRetinaHead, self).__init__(num_classes, in_channels,
                                                 anchor_scales=octave_scales * octave_base_scale, **kwargs)
        self.alpha = 0.5                      # hard-coded at iou_aware_retina_head.py:510
        self._plans = PlanCache()
        self._post = {}

    def _init_layers(self):
        self.cls_convs, self.reg_convs = nn.ModuleList(), nn.ModuleList()
        for i in range(self.stacked_convs):
            chn = self.in_channels if i == 0 else self.feat_channels
            self.cls_convs.append(ConvModule(chn, self.feat_channels, 3, stride=1, padding=1,
                                             conv_cfg=self.conv_cfg, norm_cfg=self.norm_cfg))
            self.reg_convs.append(ConvModule(chn, self.feat_channels, 3, stride=1, padding=1,
                                             conv_cfg=self.conv_cfg, norm_cfg=self.norm_cfg))
        self.retina_cls = nn.Conv2d(self.feat_channels, self.num_anchors * self.cls_out_channels, 3, padding=1)
        self.retina_reg = nn.Conv2d(self.feat_channels, self.num_anchors * 4, 3, padding=1)
        self.shared_conv = 4                  # :122 -- the IoU branch reads the last reg-tower feature
        self.use_feature_alignment = False    # :139
        self.retina_iou = nn.Conv2d(self.feat_channels, self.num_anchors, 3, padding=1)

    def init_weights(self):
        for m in self.cls_convs:
            normal_init(m.conv, std=0.01)
        for m in self.reg_convs:
            normal_init(m.conv, std=0.01)
        normal_init(self.retina_cls, std=0.01, bias=bias_init_with_prob(0.01))
        normal_init(self.retina_reg, std=0.01)
        normal_init(self.retina_iou, std=0.01)

    # ---- forward -----------------------------------------------------------------------------
    def plan_into(self, eng, sd, F, prefix=""):
        if self.norm_cfg is not None:
            raise NotImplementedError("normalised head towers are not planned")
        if not self.use_sigmoid_cls:
            raise NotImplementedError("softmax classification is not on the IoU-aware RetinaNet path")
        return eng.add_head(sd, F, prefix=prefix, stacked=self.stacked_convs, num_anchors=self.num_anchors,
                            num_classes=self.cls_out_channels)

    def forward(self, feats):
        """feats: tuple of 5 (N,256,H,W) tensors -> (list5 cls, list5 reg, list5 iou) with logical shapes
        (N, A*C, H, W), (N, A*4, H, W), (N, A, H, W) (anchor_head.py:102-103); storage is NHWC."""
        for t in feats:
            require_cuda(t, "IoUawareRetinaHead.forward")
        feats = [t.float().contiguous() for t in feats]
        dev = feats[0].device
        key = (tuple(tuple(t.shape) for t in feats), dev, param_stamp(self))

        def build():
            eng = E.Engine(dev)
            ins = [torch.empty_like(t) for t in feats]
            F = eng.new_map([(t.shape[0], t.shape[2], t.shape[3]) for t in ins], ins[0].shape[1])
            for s, t in enumerate(ins):
                n, c, h, w = t.shape
                rs = F.segs[s][0]
                lib, tp, fp = eng.lib, t.data_ptr(), F.ptr
                eng.ops.append(("pack", lambda st, tp=tp, n=n, c=c, h=h, w=w, rs=rs, fp=fp, lib=lib:
                                E.L.check(lib.iou_pack_nchw(tp, n, c, h, w, fp, rs, st))))
            outs = self.plan_into(eng, cuda_state_dict(self, dev), F)
            return eng, ins, outs
        eng, ins, outs = self._plans.get(key, build)
        for a, b in zip(ins, feats):
            a.copy_(b)
        with torch.cuda.device(dev):
            eng.run()
        # fresh tensors, like the reference's forward (the plan's own maps are overwritten by the next call);
        # FusedPlan uses plan_into() directly and keeps the zero-copy buffers
        return tuple(None if ts is None else [t.clone(memory_format=torch.preserve_format) for t in ts] for ts in outs)

    def forward_single(self, x):
        c, r, q = self.forward((x,))
        return c[0], r[0], q[0]

    # ---- get_bboxes --------------------------------------------------------------------------
    def postproc_workspace(self, featmap_sizes, n_img, cfg, device):
        key = (tuple(featmap_sizes), n_img, cfg.get('nms_pre', -1), cfg['score_thr'],
               tuple(sorted(dict(cfg['nms']).items())), cfg['max_per_img'], str(device))
        if key not in self._post:
            pcfg, soft = self.postproc_cfg(featmap_sizes, cfg)
            self._post.clear()
            self._post[key] = PP.PostprocWorkspace(pcfg, n_img, device)
            self._post[key].soft = soft
        return self._post[key]

    def postproc_cfg(self, featmap_sizes, cfg):
        """(iou_postproc_cfg, soft-NMS arguments or None) for a test_cfg, without allocating a workspace."""
        nms_cfg = dict(cfg['nms'])
        nms_type = nms_cfg.pop('type', 'nms')
        if nms_type not in ('nms', 'soft_nms'):
            raise NotImplementedError("nms type '%s' is not on the accelerated path" % nms_type)
        soft = None
        if nms_type == 'soft_nms':             # keyword defaults of nms_wrapper.soft_nms (nms_wrapper.py:52)
            method = nms_cfg.get('method', 'linear')
            if method not in PP.SOFT_NMS_METHODS:
                raise ValueError('Invalid method for SoftNMS: {}'.format(method))
            soft = (PP.SOFT_NMS_METHODS[method], float(nms_cfg.get('sigma', 0.5)),
                    float(nms_cfg.get('min_score', 1e-3)))
        pcfg = PP.make_cfg(featmap_sizes, self.anchor_strides,
                           [g.base_anchors for g in self.anchor_generators], self.cls_out_channels,
                           cfg.get('nms_pre', -1), cfg['max_per_img'], cfg['score_thr'],
                           nms_cfg.get('iou_thr', 0.5), self.target_means, self.target_stds, self.alpha)
        return pcfg, soft

    def get_bboxes_device(self, cls_scores, bbox_preds, iou_preds, img_metas, cfg, rescale=False,
                          img_info=None):
        """Asynchronous form: returns padded device tensors (dets [n,K,5], labels [n,K], counts [n])."""
        assert len(cls_scores) == len(bbox_preds) and (iou_preds is None or len(iou_preds) == len(cls_scores))
        n_img = len(img_metas)
        assert cls_scores[0].shape[0] == n_img
        sizes = [tuple(t.shape[-2:]) for t in cls_scores]
        dev = cls_scores[0].device
        wsp = self.postproc_workspace(sizes, n_img, cfg, dev)
        if img_info is None:
            img_info = PP.make_img_info(img_metas, dev)
        with torch.cuda.device(dev):
            if getattr(wsp, 'soft', None) is not None:      # test_cfg.nms = dict(type='soft_nms', ...)
                boxes, scores_cm, _ = PP.decode_candidates(wsp, cls_scores, bbox_preds, iou_preds, img_info, rescale)
                return PP.batched_soft_nms(wsp, boxes, scores_cm, *wsp.soft)
            return PP.get_bboxes_device(wsp, cls_scores, bbox_preds, iou_preds, img_info, rescale)

    def in_kernel_envelope(self, featmap_sizes, cfg):
        """True when the five batched kernels cover this test_cfg (the limits iou_get_bboxes checks, csrc/postproc.cu):
        per-level top-k <= 2048, <= 6144 candidates per image, classes a multiple of 4 and <= 256,
        classes x (max_per_img + 1) <= 8192 kept-row slots."""
        nms_pre, k = cfg.get('nms_pre', -1), cfg['max_per_img']
        per_level = [h * w * self.num_anchors for (h, w) in featmap_sizes]
        m = sum(min(n, nms_pre) if nms_pre > 0 else n for n in per_level)
        c = self.cls_out_channels
        return (k >= 1 and (nms_pre <= 2048 or max(per_level) <= 2048) and m <= 6144 and c % 4 == 0 and c <= 256
                and c * (k + 1) <= 8192)

    def _get_bboxes_level_by_level(self, cls_scores, bbox_preds, iou_preds, img_metas, cfg, rescale):
        """test_cfg outside the batched kernels' envelope (nms_pre <= 0 or > 2048, max_per_img in the thousands, an
        odd class count ...): the reference's own schedule (:434-564) image by image and level by level on the device
        -- elementwise / top-k in torch, then mmdet.ops through multiclass_nms's class-by-class path.  Same outputs;
        none of the launch fusion."""
        from .bbox_nms import multiclass_nms
        from .transforms import delta2bbox
        nms_pre = cfg.get('nms_pre', -1)
        sizes = [tuple(t.shape[-2:]) for t in cls_scores]
        dev = cls_scores[0].device
        anchors = [g.grid_anchors(sz, st, device=dev) for g, sz, st in zip(self.anchor_generators, sizes, self.anchor_strides)]
        out = []
        for i, meta in enumerate(img_metas):
            boxes, scores = [], []
            for l, a in enumerate(anchors):
                s = cls_scores[l][i].permute(1, 2, 0).reshape(-1, self.cls_out_channels).sigmoid()
                d = bbox_preds[l][i].permute(1, 2, 0).reshape(-1, 4)
                if iou_preds is not None:
                    q = iou_preds[l][i].permute(1, 2, 0).reshape(-1, 1).sigmoid()
                    s = s.pow(self.alpha) * q.pow(1 - self.alpha)
                if nms_pre > 0 and s.shape[0] > nms_pre:
                    top = s.max(dim=1)[0].topk(nms_pre)[1]
                    a, d, s = a[top], d[top], s[top]
                boxes.append(delta2bbox(a, d, self.target_means, self.target_stds, meta['img_shape']))
                scores.append(s)
            boxes, scores = torch.cat(boxes), torch.cat(scores)
            if rescale:
                boxes = boxes / boxes.new_tensor(meta['scale_factor'])
            scores = torch.cat([scores.new_zeros(scores.shape[0], 1), scores], dim=1)     # background column
            out.append(multiclass_nms(boxes, scores, cfg['score_thr'], cfg['nms'], cfg['max_per_img']))
        return out

    def get_bboxes(self, cls_scores, bbox_preds, iou_preds, gt_bboxes, gt_labels, img_metas, cfg,
                   rescale=False):
        """Same signature/return as the reference (:390-461).  gt_* are accepted and ignored (the
        reference only feeds them to a discarded diagnostic, :517-524)."""
        for t in cls_scores:
            require_cuda(t, "IoUawareRetinaHead.get_bboxes")
        if not self.in_kernel_envelope([tuple(t.shape[-2:]) for t in cls_scores], cfg):
            return self._get_bboxes_level_by_level(cls_scores, bbox_preds, iou_preds, img_metas, cfg, rescale)
        dets, labels, counts = self.get_bboxes_device(cls_scores, bbox_preds, iou_preds, img_metas, cfg,
                                                      rescale)
        return [(d.clone(), l.clone()) for d, l in PP.split_results(dets, labels, counts)]


HEADS.register_module(IoUawareRetinaHead, name='IoUAwareRetinaHead')   # BASELINE.json spelling
