"""GPU diagnostic: where does the residual error of the split-bf16 tensor-core conv come from?
(1) inputs exactly representable in bf16 (lo planes zero): any error vs fp64 is ACCUMULATION error of
    the tensor core's fp32 accumulator; compare with torch fp32 conv error on the same data.
(2) generic fp32 inputs: adds the 16-bit hi+lo REPRESENTATION error."""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from iou_aware_single_stage_object_detector_b200 import engine as E  # noqa: E402

DEV = "cuda:0"


def run(x, w, passes):
    eng = E.Engine(DEV, passes=passes)
    m = eng.pack_input(x.to(DEV).contiguous())
    out = eng.conv("t", [m], E.TAPS_3X3, E.pack_weight(w, w.shape[0]), x.shape[1], w.shape[0])
    y = eng.unpack_output(out)
    eng.run()
    torch.cuda.synchronize()
    return y.cpu().double()


def main():
    g = torch.Generator().manual_seed(0)
    for cin in (64, 256, 1024):
        x = torch.randn(2, cin, 24, 32, generator=g)
        w = torch.randn(256, cin, 3, 3, generator=g) * (2.0 / (cin * 9)) ** 0.5
        xb, wb = x.bfloat16().float(), w.bfloat16().float()
        for name, xi, wi in (("bf16-exact inputs", xb, wb), ("generic fp32 inputs", x, w)):
            ref = F.conv2d(xi.double(), wi.double(), padding=1)
            scale = ref.abs().max().item()
            e_t = (F.conv2d(xi, wi, padding=1).double() - ref).abs()
            line = "K=%5d %-20s torch-cpu-fp32 max %.2e rms %.2e |" % (cin * 9, name, e_t.max() / scale, e_t.pow(2).mean().sqrt() / scale)
            for p in (3, 4):
                e = (run(xi, wi, p) - ref).abs()
                line += " tc passes=%d max %.2e rms %.2e |" % (p, e.max() / scale, e.pow(2).mean().sqrt() / scale)
            # output storage alone: fp64 result rounded to hi+lo bf16
            hi = ref.float().bfloat16().float()
            lo = (ref.float() - hi).bfloat16().float()
            e_s = ((hi + lo).double() - ref).abs()
            line += " hi+lo storage of exact result max %.2e rms %.2e" % (e_s.max() / scale, e_s.pow(2).mean().sqrt() / scale)
            print(line)


if __name__ == "__main__":
    main()
