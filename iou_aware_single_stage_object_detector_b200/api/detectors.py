"""Detector shell with the reference's interface (mmdet/models/detectors/{base,single_stage,
retinanet}.py): BaseDetector.forward -> forward_test -> simple_test = extract_feat + bbox_head +
get_bboxes + bbox2result.  ``simple_test_batch`` is the batched entry point (the reference asserts
one image per GPU, base.py:97-98) and runs ONE fused plan: backbone, neck, head and get_bboxes as a
fixed launch sequence (optionally replayed as a CUDA graph) with a single device->host read at the end.
"""
import logging

import torch
import torch.nn as nn

from .. import engine as E
from .. import dist as D
from .. import postproc as PP
from . import builder
from .engine_cache import PlanCache, cuda_state_dict, param_stamp, require_cuda
from .registry import DETECTORS
from .transforms import bbox2result


class BaseDetector(nn.Module):
    def __init__(self):
        super(BaseDetector, self).__init__()

    @property
    def with_neck(self):
        return hasattr(self, 'neck') and self.neck is not None

    @property
    def with_bbox(self):
        return hasattr(self, 'bbox_head') and self.bbox_head is not None

    def init_weights(self, pretrained=None):
        if pretrained is not None:
            logging.getLogger().info('load model from: {}'.format(pretrained))

    def forward_test(self, imgs, img_metas, gt_bboxes, gt_labels, **kwargs):
        for var, name in [(imgs, 'imgs'), (img_metas, 'img_metas')]:
            if not isinstance(var, list):
                raise TypeError('{} must be a list, but got {}'.format(name, type(var)))
        num_augs = len(imgs)
        if num_augs != len(img_metas):
            raise ValueError('num of augmentations ({}) != num of image meta ({})'.format(
                len(imgs), len(img_metas)))
        imgs_per_gpu = imgs[0].size(0)
        assert imgs_per_gpu == 1
        if num_augs == 1:
            return self.simple_test(imgs[0], img_metas[0], gt_bboxes[0], gt_labels[0], **kwargs)
        return self.aug_test(imgs, img_metas, **kwargs)

    def forward(self, img, img_meta, return_loss=True, **kwargs):
        if return_loss:
            return self.forward_train(img, img_meta, **kwargs)
        return self.forward_test(img, img_meta, **kwargs)


class FusedPlan(object):
    """backbone + neck + head + get_bboxes for one (N,3,H,W) input shape."""

    def __init__(self, det, shape, device, rescale, use_graph=True, passes=3):
        n, _, h, w = shape
        self.device = torch.device(device)
        self.img = torch.empty(shape, dtype=torch.float32, device=self.device)
        self.img_info = torch.zeros(n, 8, dtype=torch.float32, device=self.device)
        self.rescale = rescale
        eng = E.Engine(self.device, passes=passes)
        sd = cuda_state_dict(det, self.device)
        if not det.with_neck:
            raise NotImplementedError("the fused plan needs a neck (FPN): the heads on this path read 5 FPN levels")
        if tuple(getattr(det.backbone, "out_indices", (0, 1, 2, 3))) != (0, 1, 2, 3):
            raise NotImplementedError("the fused plan feeds all four backbone stages to the neck: "
                                      "backbone.out_indices must be (0, 1, 2, 3), got %r" % (det.backbone.out_indices,))
        feats = det.backbone.plan_into(eng, sd, self.img, prefix="backbone.")
        F = det.neck.plan_into(eng, sd, feats, prefix="neck.")
        self.outs = det.bbox_head.plan_into(eng, sd, F, prefix="bbox_head.")
        if det.test_cfg is None:
            raise RuntimeError("SingleStageDetector was built without test_cfg (nms_pre, score_thr, nms, "
                               "max_per_img): pass test_cfg=cfg.test_cfg to build_detector")
        self.eng = eng
        # anchor heads: the retina_cls epilogue also writes each anchor's max class logit (Engine.add_head)
        self.cls_max2 = getattr(eng, "cls_max2", None) if hasattr(det.bbox_head, "num_anchors") else None
        # heads return their forward() tuple; get_bboxes' kernels take (cls, reg, iou-or-None)
        to_post = getattr(det.bbox_head, "postproc_inputs", None)
        self.post_in = to_post(self.outs) if to_post is not None else (self.outs[0], self.outs[1], self.outs[2])
        sizes = [tuple(t.shape[-2:]) for t in self.outs[0]]
        # a plan owns its post-processing scratch and outputs, so that two plans (detect_stream's pipeline slots)
        # can be in flight at once
        pcfg, soft = det.bbox_head.postproc_cfg(sizes, det.test_cfg)
        self.wsp = PP.PostprocWorkspace(pcfg, n, self.device)
        self.wsp.soft = soft
        self.graph = None
        self.use_graph = use_graph
        self.range_check = getattr(det, "range_check", True) and passes == 2
        self.range = None                     # range_report() of the first batch
        self.conv_flops = eng.flops
        self.launches = eng.num_launches() + 5

    def _launch(self):
        self.eng.run()
        if self.wsp.soft is not None:            # test_cfg.nms = dict(type='soft_nms', ...)
            boxes, scores_cm, _ = PP.decode_candidates(self.wsp, self.post_in[0], self.post_in[1], self.post_in[2],
                                                       self.img_info, self.rescale)
            PP.batched_soft_nms(self.wsp, boxes, scores_cm, *self.wsp.soft)
        else:
            PP.get_bboxes_device(self.wsp, self.post_in[0], self.post_in[1], self.post_in[2], self.img_info,
                                 self.rescale, cls_max2=self.cls_max2)

    def check_range(self):
        """The fp16 + e4m3 scheme holds |v| < 65504 (and full precision up to 448): look at every activation map the
        last run produced and raise when values were clamped at the fp16 limit -- the reference's fp32 would not have
        clamped them, so the detections would silently differ."""
        self.range = self.eng.range_report()
        bad = [r for r in self.range if r["saturated"] > 0]
        if bad:
            raise RuntimeError(
                "activation range beyond the fp16 + e4m3 conv scheme (passes = 2): %d map(s) hold values at or above "
                "65504, e.g. '%s' (%d of %d elements).  Use detector.passes = 3 (bf16 hi|lo, fp32 exponent range) for "
                "this checkpoint / input scaling, or set detector.range_check = False to accept the clamping."
                % (len(bad), bad[0]["label"], bad[0]["saturated"], bad[0]["elements"]))
        return self.range

    def run(self):
        """Enqueue one pass on the current stream (inputs: self.img, self.img_info)."""
        with torch.cuda.device(self.device):
            if self.range_check and self.range is None:
                self._launch()                       # first batch of this plan, eager: activation ranges are checked
                out = self.wsp.dets, self.wsp.labels, self.wsp.counts
                self.check_range()
                return out
            if not self.use_graph:
                self._launch()
            elif self.graph is None:
                self._launch()                       # warm-up: sets function attributes, loads modules
                torch.cuda.synchronize(self.device)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._launch()
                self.graph = g
                g.replay()
            else:
                self.graph.replay()
                E.L.launch_count += self.launches
        return self.wsp.dets, self.wsp.labels, self.wsp.counts


@DETECTORS.register_module
class SingleStageDetector(BaseDetector):
    def __init__(self, backbone, neck=None, bbox_head=None, train_cfg=None, test_cfg=None, pretrained=None):
        super(SingleStageDetector, self).__init__()
        self.backbone = builder.build_backbone(backbone)
        if neck is not None:
            self.neck = builder.build_neck(neck)
        self.bbox_head = builder.build_head(bbox_head)
        self.train_cfg, self.test_cfg = train_cfg, test_cfg
        self._fused = PlanCache(max_plans=4)
        self.use_cuda_graph = True
        # tensor-core scheme of the convs (engine.Engine): None = 2 (fp16 pass + e4m3 correction pass, the fastest
        # fp32-grade scheme); 3 = bf16 hi|lo x3; 1 (plain bf16) is an explicit opt-in only
        self.passes = None
        # passes == 2 only: the first batch of every plan runs eagerly and its activation ranges are checked
        # (FusedPlan.check_range): values at or beyond the fp16 limit raise instead of being clamped silently
        self.range_check = True
        self.init_weights(pretrained=pretrained)

    def init_weights(self, pretrained=None):
        super(SingleStageDetector, self).init_weights(pretrained)
        self.backbone.init_weights(pretrained=pretrained)
        if self.with_neck:
            if isinstance(self.neck, nn.Sequential):
                for m in self.neck:
                    m.init_weights()
            else:
                self.neck.init_weights()
        self.bbox_head.init_weights()

    def extract_feat(self, img):
        x = self.backbone(img)
        if self.with_neck:
            x = self.neck(x)
        return x

    def forward_train(self, img, img_metas, gt_bboxes, gt_labels, gt_bboxes_ignore=None):
        raise NotImplementedError("training is outside the accelerated inference path")

    def resolved_passes(self):
        return 2 if self.passes is None else self.passes

    def fused_plan(self, shape, device, rescale, slot=0):
        """`slot` distinguishes independent plans (own buffers, own CUDA graph) of the same shape."""
        passes = self.resolved_passes()
        key = (tuple(shape), str(device), bool(rescale), param_stamp(self), self.use_cuda_graph, passes, slot)
        return self._fused.get(key, lambda: FusedPlan(self, shape, device, rescale, self.use_cuda_graph, passes))

    def detect_device(self, img, img_metas, rescale=False, device=None, clone=True):
        """Batched, asynchronous: (dets [n,K,5], labels [n,K] int64, counts [n] int32) on the device.
        `img` may live on the GPU or in (pinned) host memory; a host batch is copied to `device`
        (default: the parameters' device) inside this call -- all arithmetic runs on the GPU.
        The results are fresh tensors; clone=False hands out the plan's own output buffers instead, which the next
        call with the same input shape overwrites (what detect_stream / simple_test_batch use internally)."""
        if device is None:
            device = img.device if img.is_cuda else next(self.parameters()).device
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("SingleStageDetector: move the model to a CUDA device first -- this path "
                               "has no CPU fallback")
        plan = self.fused_plan(img.shape, device, rescale)
        with torch.cuda.device(device):
            plan.img.copy_(img, non_blocking=True)
            plan.img_info.copy_(PP.make_img_info(img_metas, "cpu"), non_blocking=True)
            out = plan.run()
            return tuple(t.clone() for t in out) if clone else out

    def detect_stream(self, batches, rescale=False, device=None, gather=None, img_transform=None, depth=2):
        """Pipelined batched inference over an iterable of (img, img_metas) with HOST (ideally pinned) images;
        yields (dets, labels, counts) as CPU tensors per batch, in order.  Three overlaps:
          * the host->device copy of batch i+1 runs on a copy stream while batch i computes;
          * results are read back asynchronously into pinned buffers and batch i is handed out only after batch
            i+1 has been queued, so the GPU never waits for the host between batches;
          * with depth = 2 (default) consecutive batches use two independent launch plans (own buffers, own CUDA
            graph) on two compute streams, so the tail of each persistent conv kernel and the latency-bound
            post-processing of batch i are filled by kernels of batch i+1 (~3 % more throughput, one extra set
            of activation buffers).  depth = 1 keeps a single plan and stream.
        `gather(dets, labels, counts)` (e.g. dist.gather_detections) is applied on the device first.
        With `img_transform` (an api.ImageTransform) the batches are uint8 (n, h, w, 3) BGR frames: only
        the raw bytes cross PCIe and normalisation / padding / CHW run on the device."""
        device = torch.device(device) if device is not None else next(self.parameters()).device
        if device.type != "cuda":
            raise RuntimeError("SingleStageDetector: move the model to a CUDA device first -- this path "
                               "has no CPU fallback")
        assert depth in (1, 2)
        it = iter(batches)
        with torch.cuda.device(device):
            copy_stream = torch.cuda.Stream(device)
            main = torch.cuda.current_stream(device)
            compute = [main, torch.cuda.Stream(device) if depth == 2 else main]
            stage, ready, consumed = [None, None], [None, None], [None, None]
            host, pending = [None, None], None      # pinned result buffers per slot; (slot, event) not yet yielded
            busy = [None, None]                     # event behind the side-stream work that still reads a plan's outputs
            if depth == 2:
                start = torch.cuda.Event()
                start.record(main)
                compute[1].wait_event(start)        # work queued on the caller's stream before this call comes first

            def prefetch(slot, item):
                img, metas = item
                if stage[slot] is None or stage[slot].shape != img.shape:
                    stage[slot] = torch.empty(img.shape, dtype=img.dtype, device=device)
                with torch.cuda.stream(copy_stream):
                    if consumed[slot] is not None:
                        copy_stream.wait_event(consumed[slot])      # the previous use of this slot is over
                    stage[slot].copy_(img, non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(copy_stream)
                ready[slot] = (ev, metas)

            def read_back(slot, packed, stream, ranks):
                """Async device->host copy of this batch's PACKED results (dets | labels | counts of every rank in one
                buffer) into the slot's pinned buffer, enqueued on `stream` before anything overwrites the source."""
                if host[slot] is None or host[slot].numel() != packed.numel():
                    host[slot] = torch.empty(packed.numel(), dtype=torch.uint8, pin_memory=True)
                with torch.cuda.stream(stream):
                    host[slot].copy_(packed, non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(stream)
                return ev, ranks

            def hand_out(p):
                pslot, (pev, ranks), (b_, k_) = p
                pev.synchronize()
                return tuple(t.clone() for t in D.unpack_gathered(host[pslot], ranks, b_, k_))

            nxt = next(it, None)
            if nxt is None:
                return
            prefetch(0, nxt)
            slot = 0
            while True:
                ev, metas = ready[slot]
                nxt = next(it, None)
                if nxt is not None:
                    prefetch(slot ^ 1, nxt)                          # overlaps with this batch's compute
                cs = compute[slot]
                with torch.cuda.stream(cs):
                    if img_transform is not None:
                        n_, h_, w_, _ = stage[slot].shape
                        hp_, wp_ = img_transform.pad_shape(h_, w_)
                        plan = self.fused_plan((n_, 3, hp_, wp_), device, rescale, slot if depth == 2 else 0)
                        cs.wait_event(ev)
                        img_transform(stage[slot], out=plan.img)             # uint8 HWC -> normalised padded NCHW
                    else:
                        plan = self.fused_plan(stage[slot].shape, device, rescale, slot if depth == 2 else 0)
                        cs.wait_event(ev)
                        plan.img.copy_(stage[slot], non_blocking=True)   # device->device, ~70 us for 103 MB
                    done = torch.cuda.Event()
                    done.record(cs)
                    consumed[slot] = done
                    plan.img_info.copy_(PP.make_img_info(metas, "cpu"), non_blocking=True)
                    if busy[slot] is not None:
                        cs.wait_event(busy[slot])            # the gather / read-back of this plan's previous batch
                    dets, labels, counts = plan.run()
                    shape = (dets.shape[0], dets.shape[1])
                    if isinstance(gather, D.PackedGather):
                        # one all-gather of the packed results on the gather's SIDE stream, read-back behind it; this
                        # compute stream goes straight on to the next batch
                        buf, gdone = gather(plan.wsp.packed)
                        rb = read_back(slot, buf, gather.stream, gather.world)
                        busy[slot] = rb[0]
                    elif gather is not None:                 # legacy callable on (dets, labels, counts)
                        gd, gl, gc = gather(dets, labels, counts)
                        ranks = gd.shape[0] // dets.shape[0]
                        tmp = torch.empty(ranks * plan.wsp.packed.numel(), dtype=torch.uint8, device=device)
                        for r in range(ranks):
                            pv = D.packed_views(tmp[r * plan.wsp.packed.numel():], *shape)
                            b0 = r * shape[0]
                            pv[0].copy_(gd[b0:b0 + shape[0]]); pv[1].copy_(gl[b0:b0 + shape[0]]); pv[2].copy_(gc[b0:b0 + shape[0]])
                        rb = read_back(slot, tmp, cs, ranks)
                    else:
                        rb = read_back(slot, plan.wsp.packed, cs, 1)
                    rb_ev = (rb, shape)
                # the previous batch is handed out only now, after this batch's launches are queued
                if pending is not None:
                    yield hand_out(pending)
                pending = (slot, rb_ev[0], rb_ev[1])
                if nxt is None:
                    yield hand_out(pending)
                    if depth == 2:                                   # later work on the caller's stream comes after
                        tail = torch.cuda.Event()
                        tail.record(compute[1])
                        main.wait_event(tail)
                    return
                slot ^= 1

    def simple_test_batch(self, img, img_metas, gt_bboxes=None, gt_labels=None, rescale=False):
        """Batched simple_test: list (per image) of per-class ndarray lists."""
        dets, labels, counts = self.detect_device(img, img_metas, rescale, clone=False)
        return [bbox2result(d, l, self.bbox_head.num_classes)
                for d, l in PP.split_results(dets, labels, counts)]

    def simple_test(self, img, img_meta, gt_bboxes, gt_labels, rescale=False):
        """Reference signature (single_stage.py:64-96): one image in, bbox_results[0] out."""
        return self.simple_test_batch(img, img_meta, gt_bboxes, gt_labels, rescale)[0]

    def aug_test(self, imgs, img_metas, rescale=False):
        raise NotImplementedError


@DETECTORS.register_module
class FCOS(SingleStageDetector):
    """mmdet/models/detectors/fcos.py:6-16 (with IoUawareFCOSHead: configs/fcos/iou_aware_fcos_r50_caffe_fpn_gn_1x_4gpu.py)."""

    def __init__(self, backbone, neck, bbox_head, train_cfg=None, test_cfg=None, pretrained=None):
        super(FCOS, self).__init__(backbone, neck, bbox_head, train_cfg, test_cfg, pretrained)


@DETECTORS.register_module
class RetinaNet(SingleStageDetector):
    def __init__(self, backbone, neck, bbox_head, train_cfg=None, test_cfg=None, pretrained=None):
        super(RetinaNet, self).__init__(backbone, neck, bbox_head, train_cfg, test_cfg, pretrained)
