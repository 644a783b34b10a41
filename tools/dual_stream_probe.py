"""Experiment: two independent launch plans (own buffers, own CUDA graph) replayed alternately on two streams, so
that the tail of every persistent conv kernel and the latency-bound post-processing of batch i can be filled by
kernels of batch i+1.  Prints single-plan and dual-plan throughput measured in the same process."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from iou_aware_single_stage_object_detector_b200 import postproc as PP, synthetic  # noqa: E402
from iou_aware_single_stage_object_detector_b200.api.detectors import FusedPlan  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    bench.CFG = os.path.join(bench.CFG_DIR, bench.MODELS["r50"][0])
    det, cfg = bench.build_detector(dev, "spread")
    img_host, metas = synthetic.synthetic_batch(8, 800, 1344, seed=0, pin=True)
    img = img_host.to(dev)
    plans = []
    for k in range(2):
        p = FusedPlan(det, img.shape, dev, True)
        pcfg = p.wsp.cfg
        p.wsp = PP.PostprocWorkspace(pcfg, 8, dev)          # the head caches ONE workspace: give each plan its own
        p.img.copy_(img)
        p.img_info.copy_(PP.make_img_info(metas, "cpu"))
        plans.append(p)
    streams = [torch.cuda.Stream(dev), torch.cuda.Stream(dev)]
    for k in range(2):
        with torch.cuda.stream(streams[k]):
            for _ in range(3):
                plans[k].run()
    torch.cuda.synchronize()
    K = 40

    def timed(fn):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        return time.perf_counter() - t0

    def single():
        with torch.cuda.stream(streams[0]):
            for _ in range(K):
                plans[0].run()

    def dual():
        for i in range(K):
            with torch.cuda.stream(streams[i & 1]):
                plans[i & 1].run()
    for name, fn in (("single", single), ("dual", dual), ("single", single), ("dual", dual)):
        t = timed(fn)
        print("%s: %.3f ms/step, %.1f img/s" % (name, 1e3 * t / K, 8 * K / t))
    a = [x.clone() for x in plans[0].run()]
    b = [x.clone() for x in plans[1].run()]
    torch.cuda.synchronize()
    print("plans agree:", all(torch.equal(x, y) for x, y in zip(a, b)))


if __name__ == "__main__":
    main()
