"""Generates tests/golden/detector_<case>.npz by running the LIVE shimmed reference detector (CPU).

Run in the build container only (needs /root/reference):
    python tests/golden/gen_golden_detector.py [case ...]
For every case of detector_cases.CASES the UNMODIFIED reference (oracle/ref_shim.build_reference_detector,
i.e. mmdet.models.build_detector on the reference's own config) is loaded with the case's weights and run
one image per call, the way tools/test.py:18-34 drives it:
    feats = model.extract_feat(img); outs = model.bbox_head(feats)        (single_stage.py:39-43,86-87)
    model.bbox_head.get_bboxes(*outs, gt_bboxes, gt_labels, img_meta, test_cfg, rescale=True)
and, as a cross-check, through the stock entry model(return_loss=False, rescale=True, ...) (base.py:105-123).
Stored: final detections, the candidates entering multiclass_nms (re-running the reference's own statements,
iou_aware_retina_head.py:502-549, to expose the top-k indices it does not return), and a fixed random sample
of every head map (the maps themselves are 137 MB at 800x1344).
"""
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from oracle import ref_shim, model as om  # noqa: E402
import detector_cases as DC  # noqa: E402

CFG_DIR = os.path.join(ROOT, "configs", "iou_aware_single_stage_detector")


def reference_model(name, sd):
    """The reference's own detector for this case's config, with the case's weights."""
    c = DC.CASES[name]
    ref_cfg_path = os.path.join(ref_shim.REFERENCE_ROOT, "configs", "iou_aware_single_stage_detector", c["cfg"])
    if os.path.isfile(ref_cfg_path):
        cfg = ref_shim.load_config(ref_cfg_path)
    else:
        # no IoU-aware 64x4d config exists in the reference: derive it from the 32x4d file with groups=64
        # (cf. configs/retinanet_x101_64x4d_fpn_1x.py:5-13), SURVEY section 7
        cfg = ref_shim.reference_config("iou_aware_retinanet_x101_32x4d_fpn_1x_4gpu.py")
        cfg.model.backbone.groups = c["groups"]
    model, cfg = ref_shim.build_reference_detector(cfg)
    missing = model.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    return model, cfg


def run_case(name):
    import iou_aware_single_stage_object_detector_b200 as P
    t0 = time.time()
    _, sd, _ = DC.case_state_dict(name, P, om, CFG_DIR)
    model, cfg = reference_model(name, sd)
    from mmdet.core import multiclass_nms, delta2bbox, bbox2result
    head = model.bbox_head
    img, metas = DC.case_inputs(name)
    out = {}
    tcfg = cfg.test_cfg
    for i in range(img.shape[0]):
        x = img[i:i + 1]
        gtb, gtl = [torch.zeros(0, 4)], [torch.zeros(0, dtype=torch.long)]
        with torch.no_grad():
            feats = model.extract_feat(x)
            cls, reg, iou = head(feats)
            res = head.get_bboxes(cls, reg, iou, gtb, gtl, [metas[i]], tcfg, rescale=True)
            stock = model(return_loss=False, rescale=True, img=[x], img_meta=[[metas[i]]], gt_bboxes=[gtb],
                          gt_labels=[gtl])
        d, l = res[0]
        ref_lists = bbox2result(d, l, head.num_classes)
        assert all(np.array_equal(a, b) for a, b in zip(stock, ref_lists)), "stock path != get_bboxes path"
        out["dets_%d" % i], out["labels_%d" % i] = d.numpy(), l.numpy()
        # ---- head-map samples
        for kind, maps in (("cls", cls), ("reg", reg), ("iou", iou)):
            for lv, m in enumerate(maps):
                flat = m[0].contiguous().reshape(-1)
                idx = DC.sample_index(name, "%s%d" % (kind, i), lv, flat.numel())
                out["%s_l%d_%d" % (kind, lv, i)] = flat[torch.from_numpy(idx)].numpy()
                out["%s_l%d_%d_absmax" % (kind, lv, i)] = np.float32(flat.abs().max().item())
        # ---- candidates entering multiclass_nms (the reference's own statements, :502-549)
        anchors = [head.anchor_generators[k].grid_anchors(cls[k].shape[-2:], head.anchor_strides[k])
                   for k in range(len(cls))]
        idxs, boxes, scores = [], [], []
        for lv in range(len(cls)):
            s = cls[lv][0].permute(1, 2, 0).reshape(-1, 80).sigmoid()
            q = iou[lv][0].permute(1, 2, 0).reshape(-1).sigmoid()
            bp = reg[lv][0].permute(1, 2, 0).reshape(-1, 4)
            s = s.pow(0.5) * q.view(-1, 1).expand(-1, 80).pow(0.5)
            if tcfg.nms_pre > 0 and s.shape[0] > tcfg.nms_pre:
                _, ti = s.max(dim=1)[0].topk(tcfg.nms_pre)
            else:
                ti = torch.arange(s.shape[0])
            idxs.append(ti)
            boxes.append(delta2bbox(anchors[lv][ti], bp[ti], head.target_means, head.target_stds,
                                    metas[i]["img_shape"]))
            scores.append(s[ti])
        b = torch.cat(boxes)
        b /= b.new_tensor(metas[i]["scale_factor"])
        sc = torch.cat(scores)
        pad = torch.cat([sc.new_zeros(sc.shape[0], 1), sc], dim=1)
        d2, l2 = multiclass_nms(b, pad, tcfg.score_thr, tcfg.nms, tcfg.max_per_img)
        assert torch.equal(d2, d) and torch.equal(l2, l), "exposed candidates do not reproduce get_bboxes"
        out["cand_idx_%d" % i] = torch.cat(idxs).numpy().astype(np.int32)
        out["cand_boxes_%d" % i] = b.numpy()
        out["cand_max_%d" % i] = sc.max(dim=1)[0].numpy()
        out["cand_argmax_%d" % i] = sc.argmax(dim=1).numpy().astype(np.int16)
        rows = DC.sample_index(name, "candrows%d" % i, 0, sc.shape[0])[:DC.CAND_ROWS]  # full 80-score rows of a sample
        out["cand_rows_%d" % i] = sc[torch.from_numpy(rows)].numpy()
        out["cand_pairs_%d" % i] = np.int64((sc > tcfg.score_thr).sum().item())     # (candidate, class) pairs NMS sees
        print("  %s img %d: %d dets, %d candidates, %d (cand, class) pairs above thr, top score %.4f, 100th %.4f"
              % (name, i, d.shape[0], b.shape[0], int(out["cand_pairs_%d" % i]), float(d[:, 4].max()) if d.numel() else 0.0,
                 float(d[:, 4].min()) if d.numel() else 0.0))
    path = os.path.join(HERE, "detector_%s.npz" % name)
    np.savez_compressed(path, **out)
    print("%s: %.1f KB, %.0f s" % (os.path.basename(path), os.path.getsize(path) / 1e3, time.time() - t0))


if __name__ == "__main__":
    names = sys.argv[1:] or list(DC.CASES)
    for n in names:
        run_case(n)
