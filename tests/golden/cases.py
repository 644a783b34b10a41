"""Deterministic inputs for the golden fixtures (shared by gen_golden.py and the tests).

Inputs are rebuilt from numpy's frozen legacy ``RandomState`` so only the
reference's OUTPUTS need to be stored in the .npz fixtures.
"""
import numpy as np
import torch

STRIDES = [8, 16, 32, 64, 128]
TEST_CFG = dict(nms_pre=1000, min_bbox_size=0, score_thr=0.05,
                nms=dict(type='nms', iou_thr=0.5), max_per_img=100)


def level_sizes(pad_h, pad_w):
    """FPN level sizes for a padded input: /8, /16, /32 then two stride-2 3x3 convs."""
    sizes = [(pad_h // 8, pad_w // 8), (pad_h // 16, pad_w // 16), (pad_h // 32, pad_w // 32)]
    for _ in range(2):
        h, w = sizes[-1]
        sizes.append(((h + 1) // 2, (w + 1) // 2))
    return sizes


def random_maps(rs, n_img, sizes, cls_mu=-3.0, cls_sd=2.0, iou_sd=1.5, reg_sd=0.5, A=9, C=80):
    cls, reg, iou = [], [], []
    for (h, w) in sizes:
        cls.append(torch.from_numpy((rs.randn(n_img, A * C, h, w) * cls_sd + cls_mu).astype(np.float32)))
        reg.append(torch.from_numpy((rs.randn(n_img, A * 4, h, w) * reg_sd).astype(np.float32)))
        iou.append(torch.from_numpy((rs.randn(n_img, A, h, w) * iou_sd).astype(np.float32)))
    return cls, reg, iou


POSTPROC_CASES = ["small", "few", "empty", "full"]


def postproc_case(name):
    cfg = dict(TEST_CFG)
    if name == "small":
        rs = np.random.RandomState(1234)
        sizes = level_sizes(224, 288)
        cls, reg, iou = random_maps(rs, 2, sizes)
        cfg["nms_pre"] = 300
        metas = [dict(ori_shape=(125, 163, 3), img_shape=(200, 261, 3), pad_shape=(224, 288, 3),
                      scale_factor=1.6, flip=False),
                 dict(ori_shape=(238, 313, 3), img_shape=(190, 250, 3), pad_shape=(224, 288, 3),
                      scale_factor=0.8, flip=False)]
        rescale = True
    elif name == "few":
        rs = np.random.RandomState(99)
        sizes = level_sizes(224, 288)
        cls, reg, iou = random_maps(rs, 1, sizes, cls_mu=-11.5, cls_sd=1.6)
        cfg["nms_pre"] = 300
        metas = [dict(ori_shape=(200, 261, 3), img_shape=(200, 261, 3), pad_shape=(224, 288, 3),
                      scale_factor=1.0, flip=False)]
        rescale = False
    elif name == "empty":
        rs = np.random.RandomState(7)
        sizes = level_sizes(224, 288)
        cls, reg, iou = random_maps(rs, 1, sizes, cls_mu=-14.0, cls_sd=0.5)
        cfg["nms_pre"] = 300
        metas = [dict(ori_shape=(200, 261, 3), img_shape=(200, 261, 3), pad_shape=(224, 288, 3),
                      scale_factor=1.0, flip=False)]
        rescale = False
    elif name == "full":
        rs = np.random.RandomState(2024)
        sizes = level_sizes(800, 1344)
        cls, reg, iou = random_maps(rs, 1, sizes)
        metas = [dict(ori_shape=(800, 1333, 3), img_shape=(800, 1333, 3), pad_shape=(800, 1344, 3),
                      scale_factor=1.0, flip=False)]
        rescale = True
    else:
        raise KeyError(name)
    return dict(cls=cls, reg=reg, iou=iou, img_metas=metas, cfg=cfg, rescale=rescale, sizes=sizes)


def codec_inputs():
    rs = np.random.RandomState(5)
    rois = np.array([[0, 0, 31, 31], [10, 20, 100, 60], [-19, -7, 26, 14], [500, 300, 1400, 900],
                     [-298, -117, 425, 244]], dtype=np.float32)
    rois = np.concatenate([rois, (rs.rand(59, 4) * 600).astype(np.float32)])
    rois[5:, 2:] += rois[5:, :2]
    deltas = (rs.randn(64, 4) * 0.8).astype(np.float32)
    deltas[0] = [0.1, -0.2, 0.3, 5.0]       # SURVEY a8 known answer: [0, 0, 39.7977, 799]
    deltas[1] = [0.0, 0.0, -6.0, 6.0]       # both clamps of dw/dh
    return rois, deltas


def random_dets(rs, n, extent=600.0, size=120.0):
    xy = rs.rand(n, 2) * extent
    wh = rs.rand(n, 2) * size + 4.0
    sc = rs.rand(n, 1)
    return np.concatenate([xy, xy + wh, sc], axis=1).astype(np.float32)


def nms_inputs():
    rs = np.random.RandomState(11)
    out = {"n1": random_dets(rs, 1), "n3": np.array([[0, 0, 10, 10, 0.9], [1, 1, 10, 10, 0.8],
                                                    [20, 20, 30, 30, 0.7]], dtype=np.float32),
           "n64": random_dets(rs, 64, 200, 80), "n65": random_dets(rs, 65, 200, 80),
           "n700": random_dets(rs, 700, 500, 150), "n2000": random_dets(rs, 2000),
           "n4693_dense": random_dets(rs, 4693, 300, 200)}
    return out


def pairwise_iou_f32(b):
    """IoU matrix with the reference's fp32 operation order (nms_kernel.cu:13-21)."""
    b = b.astype(np.float32)
    one = np.float32(1)
    area = (b[:, 2] - b[:, 0] + one) * (b[:, 3] - b[:, 1] + one)
    left = np.maximum(b[:, None, 0], b[None, :, 0]); right = np.minimum(b[:, None, 2], b[None, :, 2])
    top = np.maximum(b[:, None, 1], b[None, :, 1]); bot = np.minimum(b[:, None, 3], b[None, :, 3])
    w = np.maximum(right - left + one, np.float32(0)); h = np.maximum(bot - top + one, np.float32(0))
    inter = w * h
    return inter / (area[:, None] + area[None, :] - inter)


def assert_no_threshold_ties(dets, thr):
    """nms_cpu suppresses at >= thr, nms_cuda at > thr: the golden is only valid for both
    when no pair sits exactly on the threshold."""
    n = dets.shape[0]
    for s in range(0, n, 1024):
        iou = pairwise_iou_f32(dets[:, :4])[s:s + 1024] if n <= 1024 else \
            _rows_iou(dets[:, :4], s, min(s + 1024, n))
        assert not np.any(iou == np.float32(thr)), "IoU == thr tie in golden input"


def _rows_iou(b, s, e):
    b = b.astype(np.float32)
    one = np.float32(1)
    area = (b[:, 2] - b[:, 0] + one) * (b[:, 3] - b[:, 1] + one)
    r = b[s:e]
    left = np.maximum(r[:, None, 0], b[None, :, 0]); right = np.minimum(r[:, None, 2], b[None, :, 2])
    top = np.maximum(r[:, None, 1], b[None, :, 1]); bot = np.minimum(r[:, None, 3], b[None, :, 3])
    w = np.maximum(right - left + one, np.float32(0)); h = np.maximum(bot - top + one, np.float32(0))
    inter = w * h
    return inter / (area[s:e, None] + area[None, :] - inter)


def soft_nms_inputs():
    """name -> (dets (n,5) float32, iou_thr, method, sigma, min_score) for the Soft-NMS goldens: dense and
    sparse scenes, duplicated scores (tie order), integer coordinates, every method and removal regime."""
    rs = np.random.RandomState(20240607)
    out = {}
    specs = [("linear_64", 64, 200.0, 60.0, 0.5, 'linear', 0.5, 1e-3),
             ("linear_dense_300", 300, 120.0, 80.0, 0.3, 'linear', 0.5, 0.05),
             ("gauss_200", 200, 300.0, 100.0, 0.5, 'gaussian', 0.5, 1e-3),
             ("gauss_dense_500", 500, 150.0, 90.0, 0.5, 'gaussian', 0.3, 0.05),
             ("linear_ties_257", 257, 200.0, 70.0, 0.5, 'linear', 0.5, 0.02),
             ("gauss_ties_int_129", 129, 100.0, 50.0, 0.7, 'gaussian', 1.0, 0.2),
             ("linear_1", 1, 50.0, 20.0, 0.5, 'linear', 0.5, 1e-3),
             ("linear_2", 2, 10.0, 20.0, 0.5, 'linear', 0.5, 0.3),
             ("gauss_1500", 1500, 600.0, 120.0, 0.5, 'gaussian', 0.5, 1e-3)]
    for name, n, extent, size, thr, method, sigma, min_score in specs:
        d = random_dets(rs, n, extent, size)
        if "ties" in name:
            d[:, 4] = np.round(d[:, 4] * 16) / 16
        if "int" in name:
            d[:, :4] = np.round(d[:, :4])
        out[name] = (d.astype(np.float32), thr, method, sigma, min_score)
    return out


SOFT_MULTICLASS = dict(type='soft_nms', iou_thr=0.5, method='linear', sigma=0.5, min_score=0.05)


FCOS_STRIDES = (8, 16, 32, 64, 128)


def fcos_case():
    """Maps for IoUawareFCOSHead.get_bboxes: one location per cell, 80 class logits,
    positive (l, t, r, b) distances (the head's forward already applied exp, iou_aware_fcos_head.py:108),
    centerness and IoU logits; 2 images of different test scales."""
    rs = np.random.RandomState(4321)
    sizes = level_sizes(224, 288)
    cls, reg, cen, iou = [], [], [], []
    for (h, w), s in zip(sizes, FCOS_STRIDES):
        cls.append(torch.from_numpy((rs.randn(2, 80, h, w) * 2.0 - 3.0).astype(np.float32)))
        reg.append(torch.from_numpy(np.exp(rs.randn(2, 4, h, w) * 0.6 + np.log(1.5 * s)).astype(np.float32)))
        cen.append(torch.from_numpy(rs.randn(2, 1, h, w).astype(np.float32)))
        iou.append(torch.from_numpy((rs.randn(2, 1, h, w) * 1.5).astype(np.float32)))
    cfg = dict(TEST_CFG)
    cfg["nms_pre"] = 300
    metas = [dict(ori_shape=(125, 163, 3), img_shape=(200, 261, 3), pad_shape=(224, 288, 3), scale_factor=1.6,
                  flip=False),
             dict(ori_shape=(238, 313, 3), img_shape=(190, 250, 3), pad_shape=(224, 288, 3), scale_factor=0.8,
                  flip=False)]
    return dict(cls=cls, reg=reg, cen=cen, iou=iou, img_metas=metas, cfg=cfg, rescale=True, sizes=sizes)


def nms_large_inputs():
    """More boxes than one block's shared memory holds (IOU_MAX_NMS_BOXES = 6144): the mask-tile path of iou_nms.
    The reference has no size limit (nms_kernel.cu:70-131)."""
    rs = np.random.RandomState(77)
    return {"n6145": random_dets(rs, 6145, 900, 150), "n9000_dense": random_dets(rs, 9000, 400, 200),
            "n20000": random_dets(rs, 20000, 2000, 120)}


NMS_LARGE_THR = 0.45      # no pair of the inputs above has IoU == 0.45 exactly (asserted by the generator)


def multiclass_inputs():
    """multiclass_nms calls outside the batched kernels' envelope (bbox_nms.py:6-11 defaults and options):
    name -> (multi_bboxes, multi_scores, score_thr, nms_cfg, max_num, score_factors)."""
    rs = np.random.RandomState(31)
    n, C = 400, 7
    d = random_dets(rs, n, 300, 90)
    boxes = d[:, :4]
    scores = np.concatenate([np.zeros((n, 1), np.float32), rs.rand(n, C).astype(np.float32) ** 3], axis=1)
    per_class = np.concatenate([boxes + rs.rand(n, 4).astype(np.float32) * 3.0 * c for c in range(C + 1)], axis=1)
    factors = (rs.rand(n).astype(np.float32) * 0.5 + 0.5)
    nms = dict(type='nms', iou_thr=0.5)
    return {
        "keep_all_default": (boxes, scores, 0.05, nms, -1, None),          # the reference default max_num = -1
        "keep_all_zero": (boxes, scores, 0.3, nms, 0, None),
        "score_factors": (boxes, scores, 0.05, nms, 50, factors),
        "per_class_boxes": (per_class, scores, 0.05, nms, 60, None),
        "soft_keep_all": (boxes, scores, 0.2, dict(SOFT_MULTICLASS), -1, None),
    }
