"""Which (class, image) makes class_nms_kernel's straggler: per class, the candidates above score_thr, the boxes a full greedy
NMS keeps, and how deep into the score order the scan has to go before max_per_img + 1 boxes are kept (the kernel stops there).
    python tools/nms_load_probe.py [spread|reference-init]"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench as B  # noqa: E402
from iou_aware_single_stage_object_detector_b200 import postproc as PP  # noqa: E402
from iou_aware_single_stage_object_detector_b200 import synthetic  # noqa: E402


def depth_to_keep(boxes, scores, thr, iou_thr, stop):
    """Greedy NMS on the candidates above thr in score order: (n, kept_total, candidates scanned until `stop` are kept)."""
    idx = np.nonzero(scores > thr)[0]
    order = idx[np.argsort(-scores[idx], kind="stable")]
    b = boxes[order]
    area = (b[:, 2] - b[:, 0] + 1) * (b[:, 3] - b[:, 1] + 1)
    kept = []
    depth = len(order)
    for i in range(len(order)):
        if kept:
            k = np.asarray(kept)
            xx1 = np.maximum(b[i, 0], b[k, 0]); yy1 = np.maximum(b[i, 1], b[k, 1])
            xx2 = np.minimum(b[i, 2], b[k, 2]); yy2 = np.minimum(b[i, 3], b[k, 3])
            inter = np.maximum(xx2 - xx1 + 1, 0) * np.maximum(yy2 - yy1 + 1, 0)
            if (inter / (area[i] + area[k] - inter) > iou_thr).any():
                continue
        kept.append(i)
        if len(kept) == stop:
            depth = i + 1
            break
    return len(order), len(kept), depth


def main():
    dev = torch.device("cuda:0")
    weights = sys.argv[1] if len(sys.argv) > 1 else "spread"
    det, cfg = B.build_detector(dev, weights)
    img, metas = synthetic.synthetic_batch(2, 800, 1344, seed=0)
    plan = det.fused_plan(img.shape, dev, rescale=True)
    plan.img.copy_(img.to(dev))
    plan.img_info.copy_(PP.make_img_info(metas, "cpu"))
    plan.run()
    torch.cuda.synchronize()
    boxes, scores_cm, _ = PP.decode_candidates(plan.wsp, plan.post_in[0], plan.post_in[1], plan.post_in[2], plan.img_info, True)
    torch.cuda.synchronize()
    boxes, scores_cm = boxes.cpu().numpy(), scores_cm.cpu().numpy()
    rows = []
    for i in range(boxes.shape[0]):
        for c in range(scores_cm.shape[1]):
            n, kept, depth = depth_to_keep(boxes[i], scores_cm[i, c], 0.05, 0.5, 101)
            rows.append((i, c, n, kept, depth))
    a = np.asarray(rows)
    print(json.dumps({"weights": weights, "classes": len(rows), "candidates_above_thr_mean": float(a[:, 2].mean()),
                      "kept_min": int(a[:, 3].min()), "scan_depth_median": float(np.median(a[:, 4])),
                      "scan_depth_p90": float(np.percentile(a[:, 4], 90)), "scan_depth_max": int(a[:, 4].max()),
                      "classes_scanning_everything": int((a[:, 4] >= a[:, 2]).sum())}))
    worst = a[np.argsort(-a[:, 4])[:8]]
    print("worst (img, class, n, kept, depth):", worst.tolist())


if __name__ == "__main__":
    main()
