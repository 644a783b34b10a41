"""multiclass_nms with the reference's signature (mmdet/core/post_processing/bbox_nms.py:6-67),
executed as ONE batched launch pair (class_nms + final_select) instead of an 80-iteration
Python loop with a device->host sync per class."""
import torch

from .. import postproc as PP

_ws_cache = {}
BATCHED_MAX_ROWS = 6144          # IOU_MAX_CANDIDATES (include/iou_b200.h)
BATCHED_MAX_KEPT = 8192          # num_classes * (max_num + 1) of the batched kernels


def _class_by_class(multi_bboxes, multi_scores, score_thr, nms_type, nms_kw, max_num, score_factors):
    """The reference's own schedule (bbox_nms.py:36-62), one `mmdet.ops.nms` / `soft_nms` launch per class, for the
    calls the batched kernels do not cover: max_num <= 0 (the reference default -1), score_factors, per-class boxes
    (n, C*4), more than 6144 rows, more than 8192 kept-row slots.  Same outputs, including the quirks of
    :57-59: the threshold is applied to the UNscaled score, and `inds[:max_num]` with max_num = -1 drops the
    lowest-scoring detection after the sort."""
    from . import ops
    op = {'nms': ops.nms, 'soft_nms': ops.soft_nms}[nms_type]
    own_boxes = multi_bboxes.shape[1] != 4
    dets, labels = [], []
    for c in range(1, multi_scores.shape[1]):
        sel = multi_scores[:, c] > score_thr
        if not bool(sel.any()):
            continue
        b = multi_bboxes[sel, c * 4:(c + 1) * 4] if own_boxes else multi_bboxes[sel]
        sc = multi_scores[sel, c]
        if score_factors is not None:
            sc = sc * score_factors[sel]
        kept, _ = op(torch.cat([b, sc[:, None]], dim=1), **nms_kw)
        dets.append(kept)
        labels.append(torch.full((kept.shape[0],), c - 1, dtype=torch.long, device=multi_bboxes.device))
    if not dets:
        return multi_bboxes.new_zeros((0, 5)), multi_bboxes.new_zeros((0,), dtype=torch.long)
    dets, labels = torch.cat(dets), torch.cat(labels)
    if dets.shape[0] > max_num:
        order = dets[:, -1].sort(descending=True)[1][:max_num]
        dets, labels = dets[order], labels[order]
    return dets, labels


def multiclass_nms(multi_bboxes, multi_scores, score_thr, nms_cfg, max_num=-1, score_factors=None):
    """multi_bboxes (n,4); multi_scores (n, 1+C) with background in column 0.  Returns
    (dets (k,5), labels (k,) int64), class-major / row-minor order, score-sorted if k would
    exceed max_num."""
    if not multi_bboxes.is_cuda:
        raise RuntimeError("multiclass_nms: libiou_b200 handles CUDA tensors only (no CPU fallback)")
    cfg_ = dict(nms_cfg)
    nms_type = cfg_.pop('type', 'nms')
    if nms_type not in ('nms', 'soft_nms'):
        raise NotImplementedError("nms type '%s' is outside the accelerated path" % nms_type)
    iou_thr = cfg_.get('iou_thr', 0.5)
    soft = None
    if nms_type == 'soft_nms':                 # nms_wrapper.soft_nms keyword defaults (nms_wrapper.py:52)
        method = cfg_.get('method', 'linear')
        if method not in PP.SOFT_NMS_METHODS:
            raise ValueError('Invalid method for SoftNMS: {}'.format(method))
        soft = (PP.SOFT_NMS_METHODS[method], float(cfg_.get('sigma', 0.5)), float(cfg_.get('min_score', 1e-3)))
    n, c1 = multi_scores.shape
    C = c1 - 1
    if n == 0 or C == 0:
        return multi_bboxes.new_zeros((0, 5)), multi_bboxes.new_zeros((0,), dtype=torch.long)
    Cp = (C + 3) // 4 * 4                      # kernel needs a multiple of 4 classes; pad with zeros
    if (max_num is None or max_num <= 0 or score_factors is not None or multi_bboxes.shape[1] != 4
            or n > BATCHED_MAX_ROWS or Cp * (max_num + 1) > BATCHED_MAX_KEPT):
        return _class_by_class(multi_bboxes, multi_scores, score_thr, nms_type, cfg_,
                               -1 if max_num is None else max_num, score_factors)
    scores = multi_scores[:, 1:]
    scores_cm = multi_bboxes.new_zeros((1, Cp, n))
    scores_cm[0, :C] = scores.t()
    key = (n, Cp, float(score_thr), float(iou_thr), int(max_num), multi_bboxes.device)
    if key not in _ws_cache:
        # a synthetic single-level description with exactly n candidate rows
        base = [torch.zeros(1, 4)]
        cfg = PP.make_cfg([(1, n)], [1], base, Cp, -1, max_num, score_thr, iou_thr)
        _ws_cache.clear()
        _ws_cache[key] = PP.PostprocWorkspace(cfg, 1, multi_bboxes.device)
    wsp = _ws_cache[key]
    with torch.cuda.device(multi_bboxes.device):
        if soft is None:
            dets, labels, counts = PP.batched_nms(wsp, multi_bboxes.float().reshape(1, n, 4), scores_cm)
        else:
            dets, labels, counts = PP.batched_soft_nms(wsp, multi_bboxes.float().reshape(1, n, 4), scores_cm, *soft)
    k = int(counts.item())
    return dets[0, :k].clone(), labels[0, :k].clone()
