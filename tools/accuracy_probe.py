"""GPU diagnostic: how far are the head maps / detections of the CUDA path from the CPU oracle, for
each MMA-pass mode, next to the fp32-vs-fp32 noise floor (cuDNN fp32 GPU forward vs the CPU oracle)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity_util as U  # noqa: E402
from oracle import model as om  # noqa: E402
from oracle import postproc as op  # noqa: E402
import cases  # noqa: E402


def stats(mine, ref):
    out = []
    for name, a_l, b_l in (("cls", mine[0], ref[0]), ("reg", mine[1], ref[1]), ("iou", mine[2], ref[2])):
        err = max((a.cpu().contiguous() - b).abs().max().item() for a, b in zip(a_l, b_l))
        rms = float(np.sqrt(np.mean([((a.cpu().contiguous() - b) ** 2).mean().item() for a, b in zip(a_l, b_l)])))
        out.append("%s max %.3g rms %.3g" % (name, err, rms))
    return "; ".join(out)


def det_stats(results, sd_maps, metas, cfg):
    bases = U.oracle_bases()
    ds, db, n_un = 0.0, 0.0, 0
    for i in range(len(results)):
        d_ref, l_ref = op.get_bboxes_single([c[i] for c in sd_maps[0]], [r[i] for r in sd_maps[1]],
                                            [q[i] for q in sd_maps[2]], cases.STRIDES, bases,
                                            metas[i]["img_shape"], 1.0, dict(cfg.test_cfg), rescale=False)
        d_my, l_my = U.results_to_arrays(results[i])
        gd, gl = d_ref.numpy(), l_ref.numpy()
        used = np.zeros(len(gd), bool)
        for k in range(len(d_my)):
            cand = np.where((gl == l_my[k]) & ~used)[0]
            if cand.size == 0:
                n_un += 1
                continue
            err = np.abs(gd[cand, :4] - d_my[k, :4]).max(1)
            j = cand[err.argmin()]
            if err.min() > 1.0:
                n_un += 1
                continue
            used[j] = True
            db = max(db, float(err.min()))
            ds = max(ds, float(abs(gd[j, 4] - d_my[k, 4])))
    return "dets: max|dscore| %.3g max|dbox| %.3g px unmatched %d" % (ds, db, n_un)


def main():
    h, w, n = int(os.environ.get("PH", 256)), int(os.environ.get("PW", 320)), 2
    det, cfg = U.small_detector()
    sd = {k: v.clone() for k, v in det.state_dict().items()}
    dev = torch.device("cuda:0")
    det = det.to(dev)
    img = torch.randn(n, 3, h, w, generator=torch.Generator().manual_seed(3))
    metas = [dict(ori_shape=(h, w, 3), img_shape=(h, w, 3), pad_shape=(h, w, 3), scale_factor=1.0, flip=False)] * n
    ref = om.detector_forward(sd, img)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    sd_gpu = {k: v.to(dev) for k, v in sd.items()}
    cud = om.detector_forward(sd_gpu, img.to(dev))
    print("cuDNN fp32 (GPU) vs CPU oracle  :", stats(cud, ref))
    for passes in (3, 2, 4, 1):
        det.passes, det.use_cuda_graph = passes, False
        res = det.simple_test_batch(img.to(dev), metas, rescale=False)
        plan = det.fused_plan(img.shape, dev, False)
        torch.cuda.synchronize()
        print("tcgen05 passes=%d vs CPU oracle   :" % passes, stats(plan.outs, ref), "|", det_stats(res, ref, metas, cfg))


if __name__ == "__main__":
    main()
