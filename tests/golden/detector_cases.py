"""Whole-detector golden cases (shared by gen_golden_detector.py and the tests).

A case = (config file, input size, images, seeds).  Weights and inputs are rebuilt from seeds on both
sides (this repo's build_detector under torch.manual_seed + oracle.model.spread_weights_; images from a
seeded torch.Generator), so the fixtures only hold the reference's OUTPUTS: final detections, the
candidates entering multiclass_nms, and a fixed random sample of every head map.
"""
import numpy as np
import torch

STRIDES = [8, 16, 32, 64, 128]
SAMPLES_PER_MAP = 4096
CAND_ROWS = 256          # candidates whose full 80-class score rows are stored
NEAR_TIE = 2e-6       # score gaps below this leave the reference's own order undefined

# name -> dict(cfg, depth, groups, n, h, w, real_w, img_seed)
CASES = {
    # BASELINE config 2 geometry (R50, 800x1344 padded, img_shape 800x1333), 2 images
    "r50_full": dict(cfg="iou_aware_retinanet_r50_fpn_1x_4gpu.py", depth=50, groups=1, n=2, h=800, w=1344,
                     real_w=1333, img_seed=11),
    # BASELINE configs 3-4 backbones at a reduced size
    "r101_small": dict(cfg="iou_aware_retinanet_r101_fpn_1x_4gpu.py", depth=101, groups=1, n=2, h=256, w=320,
                       real_w=317, img_seed=12),
    "x101_32x4d_small": dict(cfg="iou_aware_retinanet_x101_32x4d_fpn_1x_4gpu.py", depth=101, groups=32, n=2,
                             h=256, w=320, real_w=317, img_seed=13),
    "x101_64x4d_small": dict(cfg="iou_aware_retinanet_x101_64x4d_fpn_1x.py", depth=101, groups=64, n=2,
                             h=256, w=320, real_w=317, img_seed=14),
}


def case_inputs(name):
    c = CASES[name]
    g = torch.Generator().manual_seed(c["img_seed"])
    img = torch.randn(c["n"], 3, c["h"], c["w"], generator=g)
    meta = dict(ori_shape=(c["h"], c["real_w"], 3), img_shape=(c["h"], c["real_w"], 3),
                pad_shape=(c["h"], c["w"], 3), scale_factor=1.0, flip=False)
    return img, [dict(meta) for _ in range(c["n"])]


def case_state_dict(name, P, om, cfg_dir):
    """(detector module of this repo, its spread state_dict).  P = the package, om = oracle.model."""
    import os
    c = CASES[name]
    cfg = P.Config.fromfile(os.path.join(cfg_dir, c["cfg"]))
    cfg.model.pretrained = None
    torch.manual_seed(0)
    det = P.build_detector(cfg.model, train_cfg=None, test_cfg=cfg.test_cfg)
    det.eval()
    sd = {k: v.clone() for k, v in det.state_dict().items()}
    om.spread_weights_(sd, seed=1, depth=c["depth"], groups=c["groups"])
    det.load_state_dict(sd)
    return det, sd, cfg


def sample_index(name, kind, level, numel):
    """Fixed flat indices (into the contiguous (N, C, H, W) map) of the stored head-map sample."""
    rs = np.random.RandomState(sum(map(ord, name + kind)) * 31 + level)
    k = min(SAMPLES_PER_MAP, numel)
    return np.sort(rs.choice(numel, size=k, replace=False)).astype(np.int64)
