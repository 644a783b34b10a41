"""BASELINE config 5: kernel-only microbench of decode + IoU-reweight + batched NMS.
5 levels (92x90 .. 6x6) x 9 anchors = 99 531 anchors/img, 80 classes, bs=64, seeded logits
cls~N(-3,2), iou~N(0,1.5), reg~N(0,0.5).  Prints one JSON line with the HBM roofline
of the streaming pass (max_score_kernel) and the times of the other stages."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity_util as U  # noqa: E402

PP = U.PP


def main():
    dev = torch.device("cuda:0")
    bs = int(os.environ.get("BS", 64))
    sizes = [(92, 90), (46, 45), (23, 23), (12, 12), (6, 6)]
    g = torch.Generator(device="cuda").manual_seed(0)
    cls = [torch.empty(bs, h, w, 720, device=dev).normal_(-3, 2, generator=g).permute(0, 3, 1, 2) for h, w in sizes]
    reg = [torch.empty(bs, h, w, 36, device=dev).normal_(0, 0.5, generator=g).permute(0, 3, 1, 2) for h, w in sizes]
    iou = [torch.empty(bs, h, w, 9, device=dev).normal_(0, 1.5, generator=g).permute(0, 3, 1, 2) for h, w in sizes]
    head = U.get_head()
    cfg = U.P.ConfigDict(dict(nms_pre=1000, min_bbox_size=0, score_thr=0.05, nms=dict(type='nms', iou_thr=0.5),
                              max_per_img=100))
    metas = [dict(img_shape=(736, 720, 3), scale_factor=1.0)] * bs
    wsp = head.postproc_workspace(sizes, bs, cfg, dev)
    info = PP.make_img_info(metas, dev)

    def timed(fn, iters=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(iters):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / iters

    t_all = timed(lambda: PP.get_bboxes_device(wsp, cls, reg, iou, info, False))
    boxes, scores, idx = PP.decode_candidates(wsp, cls, reg, iou, info, False)
    t_dec = timed(lambda: PP.decode_candidates(wsp, cls, reg, iou, info, False))
    t_nms = timed(lambda: PP.batched_nms(wsp, boxes, scores))
    n_anchor = sum(h * w * 9 for h, w in sizes)
    logits_bytes = bs * n_anchor * 85 * 4
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.isfile(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
    dets, labels, counts = PP.get_bboxes_device(wsp, cls, reg, iou, info, False)
    torch.cuda.synchronize()
    print(json.dumps({"workload": "config 5: decode+reweight+batched NMS, bs=%d, %d anchors/img, 80 classes" % (bs, n_anchor),
                      "ms_get_bboxes": round(t_all, 3), "ms_decode_stage": round(t_dec, 3), "ms_nms_stage": round(t_nms, 3),
                      "images_per_s": round(bs / t_all * 1e3, 1),
                      "algorithmic_read_bytes": logits_bytes,
                      "whole_pipeline_gbs": round(logits_bytes / t_all / 1e6, 1),
                      "decode_stage_gbs": round(logits_bytes / t_dec / 1e6, 1),
                      "hbm_peak_gbs": peaks["hbm_gbs"],
                      "decode_stage_frac_of_hbm": round(logits_bytes / t_dec / 1e6 / peaks["hbm_gbs"], 3),
                      "candidates_per_img": wsp.M, "dets_per_img_mean": float(counts.float().mean())}))


if __name__ == "__main__":
    main()
