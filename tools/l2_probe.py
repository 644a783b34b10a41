"""L2-residency probe: time launch ranges of the detector plan (e.g. all of layer1) as CUDA graphs at different
batch sizes.  If a stage's maps fit the 126 MB L2 at a small batch, its per-image time there shows what a
depth-first (per image chunk) launch order would buy over the layer-by-layer order at bs=8.

    python tools/l2_probe.py --batch 1 2 8 [--once]      (--once: a single eager pass per batch, for ncu)
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

RANGES = ["stem.", "backbone.layer1.", "backbone.layer2.", "backbone.layer3.", "backbone.layer4.", "neck.", "bbox_head."]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, nargs="+", default=[1, 2, 8])
    ap.add_argument("--once", action="store_true")
    ap.add_argument("--reps", type=int, default=20)
    args = ap.parse_args()
    from iou_aware_single_stage_object_detector_b200 import lib as L
    from iou_aware_single_stage_object_detector_b200 import synthetic
    dev = torch.device("cuda", 0)
    det, cfg = bench.build_detector(dev, "spread")
    det.use_cuda_graph = False
    out = {}
    for b in args.batch:
        img, metas = synthetic.synthetic_batch(b, 800, 1344, seed=0)
        plan = det.fused_plan((b, 3, 800, 1344), dev, rescale=True)
        plan.img.copy_(img.to(dev))
        plan.run()
        torch.cuda.synchronize()
        if args.once:
            torch.cuda.profiler.start()
            plan.eng.run()
            torch.cuda.synchronize()
            torch.cuda.profiler.stop()
            continue
        ops = plan.eng.ops
        res = {}
        for pref in RANGES:
            sel = [fn for name, fn in ops if name.startswith(pref)]
            if not sel:
                continue
            s = torch.cuda.Stream(dev)
            with torch.cuda.stream(s):
                st = L.stream_ptr()
                for fn in sel:
                    fn(st)
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=s):
                    st = L.stream_ptr()
                    for fn in sel:
                        fn(st)
                for _ in range(3):
                    g.replay()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(s)
                for _ in range(args.reps):
                    g.replay()
                e1.record(s)
                torch.cuda.synchronize()
                res[pref] = e0.elapsed_time(e1) / args.reps / b * 1e3      # us per image
        out[b] = res
        print("batch %d: " % b + "  ".join("%s %.1f" % (k, v) for k, v in res.items()) + "   (us per image)")
        del plan
        det._fused = type(det._fused)(max_plans=4)
        torch.cuda.empty_cache()
    if not args.once:
        print(json.dumps(out))


if __name__ == "__main__":
    main()
