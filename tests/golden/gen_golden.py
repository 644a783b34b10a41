"""Generates tests/golden/*.npz by running the REAL reference (CPU, shimmed).

Run in the build container only (needs /root/reference):
    python tests/golden/gen_golden.py
The fixtures hold OUTPUTS of the reference's own code
(IoUawareRetinaHead.get_bboxes / get_bboxes_single, multiclass_nms, nms_cpu,
AnchorGenerator, delta2bbox) for inputs that ``tests/golden/cases.py`` rebuilds
from a frozen numpy RandomState, so the inputs themselves are not stored.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)

from oracle import ref_shim  # noqa: E402
import cases  # noqa: E402


def ref_head():
    ref_shim.load_reference()
    from mmdet.models.anchor_heads import IoUawareRetinaHead
    torch.manual_seed(0)
    return IoUawareRetinaHead(
        num_classes=81, in_channels=256, stacked_convs=4, feat_channels=256,
        octave_base_scale=4, scales_per_octave=3, anchor_ratios=[0.5, 1.0, 2.0],
        anchor_strides=[8, 16, 32, 64, 128], target_means=[.0, .0, .0, .0],
        target_stds=[1.0, 1.0, 1.0, 1.0],
        loss_cls=dict(type='FocalLoss', use_sigmoid=True, gamma=2.0, alpha=0.25, loss_weight=1.0),
        loss_bbox=dict(type='SmoothL1Loss', beta=0.11, loss_weight=1.0))


def run_postproc_case(head, case):
    """Real reference: candidates that enter multiclass_nms and final detections."""
    from mmdet.core import multiclass_nms, delta2bbox
    cfg = ref_shim._to_attr(case["cfg"])
    cls, reg, iou = case["cls"], case["reg"], case["iou"]     # lists of (N, ch, H, W)
    n_img = cls[0].shape[0]
    metas = case["img_metas"]
    gtb = [torch.zeros(0, 4)] * n_img
    gtl = [torch.zeros(0, dtype=torch.long)] * n_img
    out = {}
    with torch.no_grad():
        res = head.get_bboxes(cls, reg, iou, gtb, gtl, metas, cfg, rescale=case["rescale"])
    for i, (d, l) in enumerate(res):
        out["dets_%d" % i] = d.numpy()
        out["labels_%d" % i] = l.numpy()
    # per-level candidate selection, by re-running the reference's own statements
    # (iou_aware_retina_head.py:502-549) to expose the top-k indices it does not return
    anchors = [head.anchor_generators[i].grid_anchors(cls[i].shape[-2:], head.anchor_strides[i])
               for i in range(len(cls))]
    for i in range(n_img):
        idxs, boxes, scores = [], [], []
        for l in range(len(cls)):
            s = cls[l][i].permute(1, 2, 0).reshape(-1, 80).sigmoid()
            q = iou[l][i].permute(1, 2, 0).reshape(-1).sigmoid()
            bp = reg[l][i].permute(1, 2, 0).reshape(-1, 4)
            s = s.pow(0.5) * q.view(-1, 1).expand(-1, 80).pow(0.5)
            a = anchors[l]
            if cfg.nms_pre > 0 and s.shape[0] > cfg.nms_pre:
                _, ti = s.max(dim=1)[0].topk(cfg.nms_pre)
            else:
                ti = torch.arange(s.shape[0])
            idxs.append(ti)
            boxes.append(delta2bbox(a[ti], bp[ti], head.target_means, head.target_stds,
                                    metas[i]["img_shape"]))
            scores.append(s[ti])
        b = torch.cat(boxes)
        if case["rescale"]:
            b /= b.new_tensor(metas[i]["scale_factor"])
        sc = torch.cat(scores)
        out["cand_idx_%d" % i] = torch.cat(idxs).numpy().astype(np.int32)
        out["cand_boxes_%d" % i] = b.numpy()
        out["cand_scores_%d" % i] = sc.numpy().astype(np.float32)
        # cross-check: the exposed candidates reproduce the head's own result
        pad = torch.cat([sc.new_zeros(sc.shape[0], 1), sc], dim=1)
        d2, l2 = multiclass_nms(b, pad, cfg.score_thr, cfg.nms, cfg.max_per_img)
        assert torch.equal(d2, res[i][0]) and torch.equal(l2, res[i][1])
    return out


def main():
    head = ref_head()
    nw = sys.modules["mmdet.ops.nms.nms_wrapper"]
    from mmdet.core import delta2bbox
    ref_nms_cpu = sys.modules["mmdet.ops.nms.nms_cpu"]

    # ---- anchors / codec known answers ------------------------------------------------
    kat = {}
    for i, s in enumerate(head.anchor_strides):
        kat["base_anchors_%d" % s] = head.anchor_generators[i].base_anchors.numpy()
    kat["grid_s8_3x5"] = head.anchor_generators[0].grid_anchors((3, 5), 8).numpy()
    rois, deltas = cases.codec_inputs()
    kat["delta2bbox"] = delta2bbox(torch.from_numpy(rois), torch.from_numpy(deltas),
                                   [0., 0., 0., 0.], [1., 1., 1., 1.], (800, 1333, 3)).numpy()
    kat["delta2bbox_std"] = delta2bbox(torch.from_numpy(rois), torch.from_numpy(deltas),
                                       [0.1, 0., -0.1, 0.], [0.1, 0.1, 0.2, 0.2], (300, 400, 3)).numpy()
    np.savez_compressed(os.path.join(HERE, "kat_anchor_codec.npz"), **kat)

    # ---- plain nms --------------------------------------------------------------------
    nm = {}
    for name, dets in cases.nms_inputs().items():
        cases.assert_no_threshold_ties(dets, 0.5)
        keep = ref_nms_cpu.nms(torch.from_numpy(dets), 0.5)
        d2, k2 = nw.nms(torch.from_numpy(dets), 0.5)
        assert torch.equal(keep, k2)
        nm[name] = keep.numpy()
    np.savez_compressed(os.path.join(HERE, "nms_keep.npz"), **nm)

    # ---- get_bboxes cases ---------------------------------------------------------------
    for name in cases.POSTPROC_CASES:
        case = cases.postproc_case(name)
        out = run_postproc_case(head, case)
        np.savez_compressed(os.path.join(HERE, "postproc_%s.npz" % name), **out)
        print(name, {k: v.shape for k, v in out.items() if k.startswith("dets")})
    gen_plain_retina()
    gen_boundary()
    gen_test_items()
    gen_soft_nms()
    gen_results_json()
    gen_fcos()
    gen_resize()


def gen_boundary():
    """Drop-in ops beyond the batched kernels' envelope: nms on more than 6144 boxes (the reference's nms_cpu;
    no IoU == thr ties, so '>' and '>=' agree) and multiclass_nms with max_num <= 0 / score_factors / per-class
    boxes / soft_nms (the reference's own multiclass_nms, bbox_nms.py:6-67)."""
    ref_shim.load_reference()
    from mmdet.core import multiclass_nms
    ref_nms_cpu = sys.modules["mmdet.ops.nms.nms_cpu"]
    out = {}
    for name, dets in cases.nms_large_inputs().items():
        cases.assert_no_threshold_ties(dets, cases.NMS_LARGE_THR)
        out["nms_" + name] = ref_nms_cpu.nms(torch.from_numpy(dets), cases.NMS_LARGE_THR).numpy()
    for name, (b, sc, thr, nms_cfg, max_num, fac) in cases.multiclass_inputs().items():
        d, l = multiclass_nms(torch.from_numpy(b), torch.from_numpy(sc), thr, ref_shim._to_attr(dict(nms_cfg)), max_num,
                              None if fac is None else torch.from_numpy(fac))
        out["mc_%s_dets" % name], out["mc_%s_labels" % name] = d.numpy(), l.numpy()
    np.savez_compressed(os.path.join(HERE, "boundary_ops.npz"), **out)
    print("boundary", {k: v.shape for k, v in out.items()})


def gen_test_items():
    """the reference's OWN ImageTransform.__call__ (datasets/transforms.py:31-50), BboxTransform
    (:68-104) and CustomDataset.prepare_test_img (custom.py:283-359, incl. the fork's gt_bboxes / gt_labels) on
    synthetic frames.  mmcv is absent from the reference tree: the shim restates its 0.2.8 image functions on cv2 /
    numpy (oracle/ref_shim._install_mmcv_image), the call sequence and the item layout are the reference's code."""
    ref_shim.load_reference()
    from mmdet.datasets.custom import CustomDataset
    from mmdet.datasets.transforms import ImageTransform, BboxTransform
    from gen_golden_fixtures import test_item_cases, IMG_NORM
    out = {}
    for name, c in test_item_cases().items():
        ds = object.__new__(CustomDataset)          # no annotation file: fill in what prepare_test_img reads
        ds.img_infos, ds.img_prefix, ds.proposals = [c["img_info"]], "/synthetic", None
        ds.img_scales, ds.flip_ratio, ds.resize_keep_ratio = c["img_scales"], c["flip_ratio"], c["resize_keep_ratio"]
        ds.img_transform = ImageTransform(size_divisor=32, **IMG_NORM)
        ds.bbox_transform = BboxTransform()
        ds.get_ann_info = lambda idx, a=c["ann"]: a
        ref_shim.IMREAD_REGISTRY["/synthetic/" + c["img_info"]["filename"]] = c["frame"]
        item = ds.prepare_test_img(0)
        assert sorted(item.keys()) == ["gt_bboxes", "gt_labels", "img", "img_meta"]
        out[name + "_n_img"] = np.int64(len(item["img"]))
        out[name + "_n_gt"] = np.int64(len(item["gt_bboxes"]))
        for i, (im, meta) in enumerate(zip(item["img"], item["img_meta"])):
            m = meta.data
            assert meta.cpu_only and not meta.stack
            out["%s_img_%d" % (name, i)] = im.numpy()
            out["%s_meta_%d" % (name, i)] = np.array(list(m["ori_shape"]) + list(m["img_shape"]) + list(m["pad_shape"]) +
                                                     [int(m["flip"])], dtype=np.int64)
            out["%s_sf_%d" % (name, i)] = np.asarray(m["scale_factor"], dtype=np.float64).reshape(-1)
        for i, (gb, gl) in enumerate(zip(item["gt_bboxes"], item["gt_labels"])):
            out["%s_gtb_%d" % (name, i)], out["%s_gtl_%d" % (name, i)] = gb.data.numpy(), gl.data.numpy()
    np.savez_compressed(os.path.join(HERE, "test_items.npz"), **out)
    print("test items", {k: v.shape for k, v in out.items() if "_img_" in k})


def gen_soft_nms():
    """the reference's own soft_nms (nms_wrapper.soft_nms -> soft_nms_cpu.pyx compiled with
    Cython) on cases.soft_nms_inputs(), and its multiclass_nms with nms_cfg type 'soft_nms' on the candidates
    of the 'small' get_bboxes case."""
    ref_shim.load_reference()
    nw = sys.modules["mmdet.ops.nms.nms_wrapper"]
    from mmdet.core import multiclass_nms
    out = {}
    for name, (d, thr, method, sigma, min_score) in cases.soft_nms_inputs().items():
        nd, inds = nw.soft_nms(d, thr, method=method, sigma=sigma, min_score=min_score)
        out[name + "_dets"], out[name + "_inds"] = nd, inds
    g = np.load(os.path.join(HERE, "postproc_small.npz"))
    n_img = sum(1 for k in g.files if k.startswith("cand_boxes_"))
    for i in range(n_img):
        boxes = torch.from_numpy(g["cand_boxes_%d" % i])
        scores = torch.from_numpy(g["cand_scores_%d" % i])
        padded = torch.cat([scores.new_zeros(scores.shape[0], 1), scores], dim=1)
        d, l = multiclass_nms(boxes, padded, 0.05, ref_shim._to_attr(dict(cases.SOFT_MULTICLASS)), 100)
        out["mc_dets_%d" % i], out["mc_labels_%d" % i] = d.numpy(), l.numpy()
    # the head's own get_bboxes with test_cfg.nms = dict(type='soft_nms', ...) on the 'small' case maps
    head = ref_head()
    case = cases.postproc_case("small")
    cfgd = dict(case["cfg"])
    cfgd["nms"] = dict(cases.SOFT_MULTICLASS)
    n_img = case["cls"][0].shape[0]
    with torch.no_grad():
        res = head.get_bboxes(case["cls"], case["reg"], case["iou"], [torch.zeros(0, 4)] * n_img,
                              [torch.zeros(0, dtype=torch.long)] * n_img, case["img_metas"],
                              ref_shim._to_attr(cfgd), rescale=case["rescale"])
    for i, (d, l) in enumerate(res):
        out["gb_dets_%d" % i], out["gb_labels_%d" % i] = d.numpy(), l.numpy()
    np.savez_compressed(os.path.join(HERE, "soft_nms.npz"), **out)
    print("soft nms", {k: v.shape for k, v in out.items() if k.endswith("_dets") or k.startswith("mc_dets")})


from gen_golden_fixtures import results_fixture  # noqa: E402


def gen_fcos():
    """the reference IoUawareFCOSHead.get_bboxes on cases.fcos_case(), plus the candidates
    that enter multiclass_nms (re-running the reference's own statements, iou_aware_fcos_head.py:303-339)."""
    ref_shim.load_reference()
    from mmdet.models.anchor_heads import IoUawareFCOSHead
    from mmdet.core import distance2bbox
    torch.manual_seed(0)
    head = IoUawareFCOSHead(num_classes=81, in_channels=256, stacked_convs=4, feat_channels=256,
                            strides=list(cases.FCOS_STRIDES))
    case = cases.fcos_case()
    cfg = ref_shim._to_attr(case["cfg"])
    n_img = case["cls"][0].shape[0]
    with torch.no_grad():
        res = head.get_bboxes(case["cls"], case["reg"], case["cen"], case["iou"], [torch.zeros(0, 4)] * n_img,
                              [torch.zeros(0, dtype=torch.long)] * n_img, case["img_metas"], cfg,
                              rescale=case["rescale"])
    out = {}
    for i, (d, l) in enumerate(res):
        out["dets_%d" % i], out["labels_%d" % i] = d.numpy(), l.numpy()
    points = head.get_points([t.shape[-2:] for t in case["cls"]], torch.float32, torch.device("cpu"))
    for i in range(n_img):
        bs, ss, ii = [], [], []
        for l in range(len(case["cls"])):
            scores = case["cls"][l][i].permute(1, 2, 0).reshape(-1, 80).sigmoid()
            iou = case["iou"][l][i].permute(1, 2, 0).reshape(-1).sigmoid()
            scores = scores.pow(0.3) * iou.view(-1, 1).expand(-1, 80).pow(1 - 0.3)
            bp = case["reg"][l][i].permute(1, 2, 0).reshape(-1, 4)
            idx = torch.arange(scores.shape[0])
            if scores.shape[0] > cfg.nms_pre:
                idx = scores.max(dim=1)[0].topk(cfg.nms_pre)[1]
            bs.append(distance2bbox(points[l][idx], bp[idx], max_shape=case["img_metas"][i]["img_shape"]))
            ss.append(scores[idx]), ii.append(idx)
        boxes = torch.cat(bs)
        boxes /= boxes.new_tensor(case["img_metas"][i]["scale_factor"])
        out["cand_boxes_%d" % i], out["cand_scores_%d" % i] = boxes.numpy(), torch.cat(ss).numpy()
        out["cand_idx_%d" % i] = torch.cat(ii).numpy()
    np.savez_compressed(os.path.join(HERE, "postproc_fcos.npz"), **out)
    print("fcos", {k: v.shape for k, v in out.items() if k.startswith("dets")})


def gen_resize():
    """cv2.resize(..., INTER_LINEAR) -- the call inside mmcv.imrescale / imresize that
    ImageTransform makes (transforms.py:33-40) -- on the frames of gen_golden_fixtures.resize_cases()."""
    import cv2
    from gen_golden_fixtures import resize_cases
    out = {"cv2_version": np.array(cv2.__version__)}
    for name, (img, (dw, dh)) in resize_cases().items():
        out[name] = cv2.resize(img, (dw, dh), interpolation=cv2.INTER_LINEAR)
    np.savez_compressed(os.path.join(HERE, "resize_cv2.npz"), **out)
    print("resize", {k: v.shape for k, v in out.items() if k != "cv2_version"})


def gen_results_json():
    """the reference's det2json / xyxy2xywh (core/evaluation/coco_utils.py:78-117)."""
    ref_shim.load_reference()
    import json
    from mmdet.core.evaluation.coco_utils import det2json
    ds, results = results_fixture()
    with open(os.path.join(HERE, "results_det2json.json"), "w") as f:
        json.dump(det2json(ds, results), f)
    print("results json written")


def gen_plain_retina():
    """Sibling head: reference RetinaHead.get_bboxes on the 'small' case maps."""
    ref_shim.load_reference()
    from mmdet.models.anchor_heads import RetinaHead
    torch.manual_seed(0)
    head = RetinaHead(num_classes=81, in_channels=256, stacked_convs=4, feat_channels=256, octave_base_scale=4,
                      scales_per_octave=3, anchor_ratios=[0.5, 1.0, 2.0], anchor_strides=[8, 16, 32, 64, 128],
                      target_means=[.0, .0, .0, .0], target_stds=[1.0, 1.0, 1.0, 1.0],
                      loss_cls=dict(type='FocalLoss', use_sigmoid=True, gamma=2.0, alpha=0.25, loss_weight=1.0),
                      loss_bbox=dict(type='SmoothL1Loss', beta=0.11, loss_weight=1.0))
    case = cases.postproc_case("small")
    cfg = ref_shim._to_attr(case["cfg"])
    n_img = case["cls"][0].shape[0]
    with torch.no_grad():
        res = head.get_bboxes(case["cls"], case["reg"], [torch.zeros(0, 4)] * n_img,
                              [torch.zeros(0, dtype=torch.long)] * n_img, case["img_metas"], cfg,
                              rescale=case["rescale"])
    out = {}
    for i, (d, l) in enumerate(res):
        out["dets_%d" % i], out["labels_%d" % i] = d.numpy(), l.numpy()
    np.savez_compressed(os.path.join(HERE, "postproc_plain_retina_small.npz"), **out)
    print("plain retina", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    if "--plain-retina-only" in sys.argv:
        gen_plain_retina()
        sys.exit(0)
    if "--resize-only" in sys.argv:
        gen_resize()
        sys.exit(0)
    if "--fcos-only" in sys.argv:
        gen_fcos()
        sys.exit(0)
    if "--results-only" in sys.argv:
        gen_results_json()
        sys.exit(0)
    if "--items-only" in sys.argv:
        gen_test_items()
        sys.exit(0)
    if "--boundary-only" in sys.argv:
        gen_boundary()
        sys.exit(0)
    if "--soft-nms-only" in sys.argv:
        gen_soft_nms()
        sys.exit(0)
    main()
