"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the post-processing half of the path.

Restates, on CPU, what the reference computes between the head's raw maps and
the final detections.  Float arithmetic is done with torch-CPU (ATen) fp32 ops
because ATen *is* the reference's arithmetic on this path;
index / ordering logic is restated explicitly; greedy NMS is the plain-C
``nms_oracle.c``.  Every function cites the reference lines it follows.

Layout differences from the reference are deliberate (this is a restatement,
not a copy): maps are consumed as flattened ``(H*W*A, C)`` rows and anchors are
produced analytically per row.
"""
import ctypes
import math
import os
import subprocess

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
_BUILD = os.path.join(HERE, "_build")
_LIB = None


def build_c_oracle():
    """gcc-compile nms_oracle.c -> oracle/_build/liboracle.so (idempotent)."""
    src = os.path.join(HERE, "nms_oracle.c")
    out = os.path.join(_BUILD, "liboracle.so")
    if os.path.isfile(out) and os.path.getmtime(out) >= os.path.getmtime(src):
        return out
    os.makedirs(_BUILD, exist_ok=True)
    subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-ffp-contract=off",
                           "-o", out, src])
    return out


def _lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build_c_oracle())
        _LIB.oracle_nms.restype = ctypes.c_int64
        _LIB.oracle_nms.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_float,
                                    ctypes.c_int, ctypes.c_void_p]
        _LIB.oracle_iou.restype = ctypes.c_float
        _LIB.oracle_iou.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    return _LIB


# --------------------------------------------------------------------------- anchors
def base_anchors(base_size, scales, ratios):
    """mmdet/core/anchor/anchor_generator.py:18-43 (scale_major=True, ctr=None).

    Row a = ratio_idx * len(scales) + scale_idx; xyxy inclusive; rounded.
    scales/ratios go through ``torch.Tensor(...)`` (fp32) exactly like :8-9.
    """
    scales = torch.Tensor(np.asarray(scales))
    ratios = torch.Tensor(np.asarray(ratios))
    ctr = 0.5 * (base_size - 1)
    h_r = torch.sqrt(ratios)
    w_r = 1 / h_r
    ws = (base_size * w_r[:, None] * scales[None, :]).reshape(-1)
    hs = (base_size * h_r[:, None] * scales[None, :]).reshape(-1)
    out = torch.stack([ctr - 0.5 * (ws - 1), ctr - 0.5 * (hs - 1),
                       ctr + 0.5 * (ws - 1), ctr + 0.5 * (hs - 1)], dim=-1)
    return out.round()


def grid_anchors(base, feat_h, feat_w, stride):
    """anchor_generator.py:53-70: anchor[(y*W+x)*A + a] = base[a] + (x*s, y*s, x*s, y*s)."""
    ys, xs = torch.meshgrid(torch.arange(feat_h), torch.arange(feat_w), indexing="ij")
    shift = torch.stack([xs, ys, xs, ys], dim=-1).reshape(-1, 1, 4) * stride
    return (base[None, :, :] + shift.to(base.dtype)).reshape(-1, 4)


def retina_anchor_scales(octave_base_scale=4, scales_per_octave=3):
    """iou_aware_retina_head.py:81-83 (numpy float64 -> fp32 inside AnchorGenerator)."""
    return np.array([2 ** (i / scales_per_octave) for i in range(scales_per_octave)]) * octave_base_scale


# --------------------------------------------------------------------------- box codec
MAX_RATIO = abs(math.log(16 / 1000))  # transforms.py:57, wh_ratio_clip=16/1000


def delta2bbox(rois, deltas, means=(0., 0., 0., 0.), stds=(1., 1., 1., 1.), max_shape=None):
    """mmdet/core/bbox/transforms.py:44-78 for (n,4) deltas."""
    means = deltas.new_tensor(means)
    stds = deltas.new_tensor(stds)
    d = deltas * stds + means
    dx, dy = d[:, 0], d[:, 1]
    dw = d[:, 2].clamp(min=-MAX_RATIO, max=MAX_RATIO)
    dh = d[:, 3].clamp(min=-MAX_RATIO, max=MAX_RATIO)
    px = (rois[:, 0] + rois[:, 2]) * 0.5
    py = (rois[:, 1] + rois[:, 3]) * 0.5
    pw = rois[:, 2] - rois[:, 0] + 1.0
    ph = rois[:, 3] - rois[:, 1] + 1.0
    gw = pw * dw.exp()
    gh = ph * dh.exp()
    gx = torch.addcmul(px, pw, dx, value=1)   # transforms.py:66 (addcmul(px, 1, pw, dx))
    gy = torch.addcmul(py, ph, dy, value=1)
    x1 = gx - gw * 0.5 + 0.5
    y1 = gy - gh * 0.5 + 0.5
    x2 = gx + gw * 0.5 - 0.5
    y2 = gy + gh * 0.5 - 0.5
    if max_shape is not None:
        x1 = x1.clamp(min=0, max=max_shape[1] - 1)
        y1 = y1.clamp(min=0, max=max_shape[0] - 1)
        x2 = x2.clamp(min=0, max=max_shape[1] - 1)
        y2 = y2.clamp(min=0, max=max_shape[0] - 1)
    return torch.stack([x1, y1, x2, y2], dim=-1)


# --------------------------------------------------------------------------- NMS
def nms(dets, iou_thr, mode="cuda"):
    """Greedy NMS; returns ascending ORIGINAL indices (int64 tensor).

    mode "cuda": suppress at IoU > thr (mmdet/ops/nms/src/nms_kernel.cu:60);
    mode "cpu" : suppress at IoU >= thr (mmdet/ops/nms/src/nms_cpu.cpp:55).
    """
    d = np.ascontiguousarray(dets.detach().cpu().numpy() if isinstance(dets, torch.Tensor) else dets,
                             dtype=np.float32)
    n = d.shape[0]
    keep = np.empty(n, dtype=np.int64)
    k = _lib().oracle_nms(d.ctypes.data, n, float(np.float32(iou_thr)),
                          0 if mode == "cuda" else 1, keep.ctypes.data)
    return torch.from_numpy(keep[:k].copy())


def iou_pair(a, b):
    a = np.ascontiguousarray(a, dtype=np.float32)
    b = np.ascontiguousarray(b, dtype=np.float32)
    return float(_lib().oracle_iou(a.ctypes.data, b.ctypes.data))


def multiclass_nms(boxes, scores, score_thr, iou_thr, max_num, mode="cuda", return_index=False, soft=None):
    """mmdet/core/post_processing/bbox_nms.py:6-67.

    soft = None: nms_cfg type 'nms'.  soft = dict(method=, sigma=, min_score=): type 'soft_nms'
    (nms_wrapper.soft_nms per class, :48-54): rows of a class then come in soft-NMS selection order with
    their decayed scores.

    boxes (n,4); scores (n, 1+C) with background column 0.  Returns dets (k,5),
    labels (k,) int64 and, if asked, the candidate row index of every det.
    Candidate order within a class = ascending row index (nms returns ascending
    indices, nms_wrapper.py:49); classes are concatenated 0..C-1; if more than
    max_num remain they are re-ordered by score descending (:57-62; the
    reference's sort is not stable -- ties are broken here by position).
    """
    n, c1 = scores.shape
    out_d, out_l, out_i = [], [], []
    for c in range(1, c1):
        rows = torch.nonzero(scores[:, c] > score_thr).reshape(-1)   # bbox_nms.py:37 (strict >)
        if rows.numel() == 0:
            continue
        d = torch.cat([boxes[rows], scores[rows, c, None]], dim=1)
        if soft is not None:
            from . import soft_nms as SN
            nd, keep = SN.soft_nms(d.numpy(), iou_thr, **soft)
            out_d.append(torch.from_numpy(nd))
            out_l.append(torch.full((len(keep),), c - 1, dtype=torch.long))
            out_i.append(rows[torch.from_numpy(keep)])
            continue
        keep = nms(d, iou_thr, mode)
        out_d.append(d[keep])
        out_l.append(torch.full((keep.numel(),), c - 1, dtype=torch.long))
        out_i.append(rows[keep])
    if not out_d:
        z = (boxes.new_zeros((0, 5)), torch.zeros((0,), dtype=torch.long))
        return z + (torch.zeros((0,), dtype=torch.long),) if return_index else z
    dets, labels, idx = torch.cat(out_d), torch.cat(out_l), torch.cat(out_i)
    if dets.shape[0] > max_num:
        order = torch.sort(dets[:, 4], descending=True, stable=True)[1][:max_num]
        dets, labels, idx = dets[order], labels[order], idx[order]
    return (dets, labels, idx) if return_index else (dets, labels)


# --------------------------------------------------------------------------- get_bboxes
def reweight_scores(cls_rows, iou_rows, alpha=0.5):
    """iou_aware_retina_head.py:505-531: sigmoid(cls)^alpha * sigmoid(iou)^(1-alpha)."""
    s = cls_rows.sigmoid()
    q = iou_rows.sigmoid()
    return s.pow(alpha) * q.reshape(-1, 1).expand(-1, s.shape[1]).pow(1 - alpha)


def level_candidates(cls_map, reg_map, iou_map, stride, base, img_shape, nms_pre,
                     num_classes=80, means=(0., 0., 0., 0.), stds=(1., 1., 1., 1.), alpha=0.5):
    """One FPN level of get_bboxes_single (iou_aware_retina_head.py:499-551).

    cls_map (A*C,H,W), reg_map (A*4,H,W), iou_map (A,H,W).  Returns candidate
    boxes (k,4), scores (k,C) and the level-local anchor index of every
    candidate, in the reference's order (top-k order, or natural order when the
    level has <= nms_pre anchors).
    """
    h, w = cls_map.shape[-2:]
    cls_rows = cls_map.permute(1, 2, 0).reshape(-1, num_classes)
    reg_rows = reg_map.permute(1, 2, 0).reshape(-1, 4)
    anchors = grid_anchors(base, h, w, stride)
    if iou_map is None:      # plain RetinaHead: scores = sigmoid(cls) (anchor_head.py:404-407)
        scores = cls_rows.sigmoid()
    else:
        scores = reweight_scores(cls_rows, iou_map.permute(1, 2, 0).reshape(-1), alpha)
    idx = torch.arange(scores.shape[0])
    if nms_pre > 0 and scores.shape[0] > nms_pre:
        best = scores.max(dim=1)[0]
        idx = best.topk(nms_pre)[1]                    # :544, sorted descending
    boxes = delta2bbox(anchors[idx], reg_rows[idx], means, stds, img_shape)
    return boxes, scores[idx], idx


def get_bboxes_single(cls_maps, reg_maps, iou_maps, strides, bases, img_shape, scale_factor,
                      cfg, rescale=False, num_classes=80, nms_mode="cuda", return_index=False,
                      **kw):
    """iou_aware_retina_head.py:463-564 for one image; cfg needs nms_pre, score_thr,
    nms['iou_thr'], max_per_img."""
    bs, ss = [], []
    if iou_maps is None:
        iou_maps = [None] * len(cls_maps)
    for c, r, q, s, b in zip(cls_maps, reg_maps, iou_maps, strides, bases):
        bx, sc, _ = level_candidates(c, r, q, s, b, img_shape, cfg.get("nms_pre", -1),
                                     num_classes, **kw)
        bs.append(bx)
        ss.append(sc)
    boxes = torch.cat(bs)
    if rescale:
        boxes = boxes / boxes.new_tensor(scale_factor)   # :553-554 (after the clamp)
    scores = torch.cat(ss)
    scores = torch.cat([scores.new_zeros(scores.shape[0], 1), scores], dim=1)  # :556-558
    nms_cfg = dict(cfg["nms"])
    soft = None
    if nms_cfg.get("type", "nms") == "soft_nms":      # bbox_nms.py:29-31 -> nms_wrapper.soft_nms defaults
        soft = dict(method=nms_cfg.get("method", "linear"), sigma=nms_cfg.get("sigma", 0.5),
                    min_score=nms_cfg.get("min_score", 1e-3))
    return multiclass_nms(boxes, scores, cfg["score_thr"], nms_cfg.get("iou_thr", 0.5),
                          cfg["max_per_img"], nms_mode, return_index, soft=soft)


def distance2bbox(points, distance, max_shape=None):
    """mmdet/core/bbox/transforms.py:169-190."""
    x1 = points[:, 0] - distance[:, 0]
    y1 = points[:, 1] - distance[:, 1]
    x2 = points[:, 0] + distance[:, 2]
    y2 = points[:, 1] + distance[:, 3]
    if max_shape is not None:
        x1 = x1.clamp(min=0, max=max_shape[1] - 1)
        y1 = y1.clamp(min=0, max=max_shape[0] - 1)
        x2 = x2.clamp(min=0, max=max_shape[1] - 1)
        y2 = y2.clamp(min=0, max=max_shape[0] - 1)
    return torch.stack([x1, y1, x2, y2], -1)


def fcos_points(feat_h, feat_w, stride):
    """IoUawareFCOSHead.get_points_single (iou_aware_fcos_head.py:392-401): (x*s + s//2, y*s + s//2), row-major."""
    xs = torch.arange(0, feat_w * stride, stride, dtype=torch.float32)
    ys = torch.arange(0, feat_h * stride, stride, dtype=torch.float32)
    y, x = torch.meshgrid(ys, xs, indexing="ij")
    return torch.stack((x.reshape(-1), y.reshape(-1)), dim=-1) + stride // 2


def fcos_get_bboxes_single(cls_maps, reg_maps, iou_maps, strides, img_shape, scale_factor, cfg, rescale=False,
                           num_classes=80, alpha=0.3, nms_mode="cuda", return_candidates=False):
    """IoUawareFCOSHead.get_bboxes_single (iou_aware_fcos_head.py:270-366) for one image: score =
    sigmoid(cls)^alpha * sigmoid(iou)^(1-alpha) (:312-316), per-level top-k of the best class (:321-331),
    distance2bbox (:332), rescale after the clamp (:338-339), multiclass_nms (:347-352).  The centerness maps
    do not enter the result (their uses are commented out in the reference)."""
    bs, ss, ii = [], [], []
    nms_pre = cfg.get("nms_pre", -1)
    for c, r, q, s in zip(cls_maps, reg_maps, iou_maps, strides):
        scores = c.permute(1, 2, 0).reshape(-1, num_classes).sigmoid()
        iou = q.permute(1, 2, 0).reshape(-1).sigmoid()
        scores = scores.pow(alpha) * iou.view(-1, 1).expand(-1, scores.size(-1)).pow(1 - alpha)
        bbox_pred = r.permute(1, 2, 0).reshape(-1, 4)
        points = fcos_points(c.shape[-2], c.shape[-1], s)
        idx = torch.arange(scores.shape[0])
        if nms_pre > 0 and scores.shape[0] > nms_pre:
            idx = scores.max(dim=1)[0].topk(nms_pre)[1]
        bs.append(distance2bbox(points[idx], bbox_pred[idx], max_shape=img_shape))
        ss.append(scores[idx])
        ii.append(idx)
    boxes = torch.cat(bs)
    if rescale:
        boxes = boxes / boxes.new_tensor(scale_factor)
    scores = torch.cat(ss)
    if return_candidates:
        return boxes, scores, torch.cat(ii)
    padded = torch.cat([scores.new_zeros(scores.shape[0], 1), scores], dim=1)
    return multiclass_nms(boxes, padded, cfg["score_thr"], cfg["nms"]["iou_thr"], cfg["max_per_img"], nms_mode)


def candidates_single(cls_maps, reg_maps, iou_maps, strides, bases, img_shape, scale_factor,
                      nms_pre, rescale=False, num_classes=80, **kw):
    """The (boxes (M,4), scores (M,C), level-local index (M,)) that enter multiclass_nms."""
    bs, ss, ii = [], [], []
    for c, r, q, s, b in zip(cls_maps, reg_maps, iou_maps, strides, bases):
        bx, sc, ix = level_candidates(c, r, q, s, b, img_shape, nms_pre, num_classes, **kw)
        bs.append(bx), ss.append(sc), ii.append(ix)
    boxes = torch.cat(bs)
    if rescale:
        boxes = boxes / boxes.new_tensor(scale_factor)
    return boxes, torch.cat(ss), torch.cat(ii)


def bbox2result(dets, labels, num_classes):
    """mmdet/core/bbox/transforms.py:148-166 (num_classes includes background)."""
    if dets.shape[0] == 0:
        return [np.zeros((0, 5), dtype=np.float32) for _ in range(num_classes - 1)]
    d, l = dets.cpu().numpy(), labels.cpu().numpy()
    return [d[l == i, :] for i in range(num_classes - 1)]


# --------------------------------------------------------------------------- focal loss
def sigmoid_focal_loss_forward(logits, targets, gamma, alpha):
    """mmdet/ops/sigmoid_focal_loss/src/sigmoid_focal_loss_cuda.cu:24-63 (fp32).

    logits (N,C); targets (N,) int64 in 0..C (0 = background, class d <-> d+1).
    """
    x = logits.float()
    n, c = x.shape
    t = targets.reshape(-1, 1)
    d1 = torch.arange(1, c + 1).reshape(1, -1)
    c1 = (t == d1).float()
    c2 = ((t >= 0) & (t != d1)).float()
    p = 1.0 / (1.0 + torch.exp(-x))
    tiny = torch.finfo(torch.float32).tiny
    term1 = torch.pow(1.0 - p, gamma) * torch.log(p.clamp(min=tiny))
    pos = (x >= 0).float()
    term2 = torch.pow(p, gamma) * (-1.0 * x * pos - torch.log(1.0 + torch.exp(x - 2.0 * x * pos)))
    return -c1 * term1 * alpha - c2 * term2 * (1.0 - alpha)


def sigmoid_focal_loss_backward(logits, targets, d_losses, gamma, alpha):
    """sigmoid_focal_loss_cuda.cu:66-105."""
    x = logits.float()
    n, c = x.shape
    t = targets.reshape(-1, 1)
    d1 = torch.arange(1, c + 1).reshape(1, -1)
    c1 = (t == d1).float()
    c2 = ((t >= 0) & (t != d1)).float()
    p = 1.0 / (1.0 + torch.exp(-x))
    tiny = torch.finfo(torch.float32).tiny
    term1 = torch.pow(1.0 - p, gamma) * (1.0 - p - p * gamma * torch.log(p.clamp(min=tiny)))
    pos = (x >= 0).float()
    term2 = torch.pow(p, gamma) * (
        (-1.0 * x * pos - torch.log(1.0 + torch.exp(x - 2.0 * x * pos))) * (1.0 - p) * gamma - p)
    return (-c1 * term1 * alpha - c2 * term2 * (1.0 - alpha)) * d_losses


def _focal_terms_typed(logits, targets, gamma, alpha, dtype, backward):
    """The reference kernel for scalar_t = at::Half / double (sigmoid_focal_loss_cuda.cu:24-105, dispatched at
    :128,167): locals are scalar_t (a Half local is rounded at every assignment and computes in float), every
    transcendental is the FLOAT function (expf / logf / powf) and the literals are double."""
    A = torch.float32 if dtype == torch.float16 else dtype          # arithmetic type of scalar_t
    rnd = lambda v: v.to(dtype).to(A)
    f32 = lambda v: v.to(torch.float32)
    f64 = lambda v: v.to(torch.float64)
    x = logits.to(dtype).to(A)
    n, c = x.shape
    t = targets.reshape(-1, 1)
    d1 = torch.arange(1, c + 1).reshape(1, -1)
    c1 = (t == d1).to(A)
    c2 = ((t >= 0) & (t != d1)).to(A)
    a32 = torch.tensor(alpha, dtype=torch.float32)
    zn, zp = rnd((1.0 - f64(a32)).to(A)), rnd(a32.to(A))
    tiny = torch.finfo(torch.float32).tiny
    p = rnd((1.0 / (1.0 + f64(torch.exp(-f32(x))))).to(A))
    pos = f64(x >= 0)
    soft = f64(torch.log(f32(1.0 + f64(torch.exp(f32(f64(x) - 2.0 * f64(x) * pos))))))
    if not backward:
        term1 = rnd((torch.pow(f32(1.0 - f64(p)), gamma) * torch.log(f32(p).clamp(min=tiny))).to(A))
        term2 = rnd((f64(torch.pow(f32(p), gamma)) * (-1.0 * f64(x) * pos - soft)).to(A))
    else:
        term1 = rnd((f64(torch.pow(f32(1.0 - f64(p)), gamma)) *
                     (1.0 - f64(p) - f64(f32(p) * gamma * torch.log(f32(p).clamp(min=tiny))))).to(A))
        term2 = rnd((f64(torch.pow(f32(p), gamma)) *
                     ((-1.0 * f64(x) * pos - soft) * (1.0 - f64(p)) * float(torch.tensor(gamma, dtype=torch.float32))
                      - f64(p))).to(A))
    out = rnd(rnd(rnd(-c1 * term1) * zp))
    out = rnd(out + rnd(rnd(-c2 * term2) * zn))
    return out, rnd


def sigmoid_focal_loss_forward_typed(logits, targets, gamma, alpha, dtype):
    """fp16 / fp64 flavours of sigmoid_focal_loss_cuda.forward; returns a tensor of `dtype`."""
    out, _ = _focal_terms_typed(logits, targets, gamma, alpha, dtype, False)
    return out.to(dtype)


def sigmoid_focal_loss_backward_typed(logits, targets, d_losses, gamma, alpha, dtype):
    out, rnd = _focal_terms_typed(logits, targets, gamma, alpha, dtype, True)
    return rnd(out * d_losses.to(dtype).to(out.dtype)).to(dtype)

