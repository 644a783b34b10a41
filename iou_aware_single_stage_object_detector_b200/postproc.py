"""Host side of the post-processing kernels (thin ctypes glue over include/iou_b200.h).

Mirrors the argument meaning of IoUawareRetinaHead.get_bboxes
(mmdet/models/anchor_heads/iou_aware_retina_head.py:390-564) and multiclass_nms
(mmdet/core/post_processing/bbox_nms.py:6-67) of the reference.
"""
import ctypes

import torch

from . import lib as L


def make_cfg(feat_sizes, strides, base_anchors, num_classes, nms_pre, max_per_img, score_thr, iou_thr,
             target_means=(0., 0., 0., 0.), target_stds=(1., 1., 1., 1.), alpha=0.5,
             wh_ratio_clip=16 / 1000, decode_mode=0):
    cfg = L.PostprocCfg()
    cfg.decode_mode = int(decode_mode)
    cfg.num_levels = len(feat_sizes)
    cfg.num_anchors = int(base_anchors[0].shape[0])
    cfg.num_classes = int(num_classes)
    cfg.nms_pre = int(nms_pre)
    cfg.max_per_img = int(max_per_img)
    if cfg.num_levels > L.MAX_LEVELS or cfg.num_anchors > L.MAX_ANCHORS:
        raise RuntimeError("too many levels / anchors for libiou_b200")
    for l, ((h, w), s, b) in enumerate(zip(feat_sizes, strides, base_anchors)):
        cfg.feat_h[l], cfg.feat_w[l], cfg.stride[l] = int(h), int(w), int(s)
        bl = b.detach().cpu().float().tolist()
        for a in range(cfg.num_anchors):
            for q in range(4):
                cfg.base_anchors[l][a][q] = bl[a][q]
    for q in range(4):
        cfg.target_means[q] = float(target_means[q])
        cfg.target_stds[q] = float(target_stds[q])
    cfg.alpha, cfg.score_thr, cfg.iou_thr = float(alpha), float(score_thr), float(iou_thr)
    cfg.wh_ratio_clip = float(wh_ratio_clip)
    return cfg


def num_candidates(cfg):
    m = L.load().iou_postproc_num_candidates(ctypes.byref(cfg))
    if m < 0:
        L.check(m)
    return m


def make_img_info(img_metas, device):
    """[n][8] fp32 = (img_h, img_w, sf_x1, sf_y1, sf_x2, sf_y2, 0, 0) from mmdet img_meta dicts."""
    rows = []
    for m in img_metas:
        sf = m.get("scale_factor", 1.0)
        sf = [float(sf)] * 4 if not hasattr(sf, "__len__") else [float(v) for v in sf]
        rows.append([float(m["img_shape"][0]), float(m["img_shape"][1])] + sf + [0.0, 0.0])
    return torch.tensor(rows, dtype=torch.float32).to(device, non_blocking=True)


def nhwc_rows(t):
    """(N, CH, H, W) logical tensor -> same tensor with NHWC memory (no copy if it already is)."""
    if t.dtype != torch.float32:
        t = t.float()
    t = t.contiguous(memory_format=torch.channels_last)
    if t.data_ptr() % 16:
        t = t.clone(memory_format=torch.preserve_format)
    return t


def _ptr_array(tensors):
    return (ctypes.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])


class PostprocWorkspace(object):
    """Device scratch + outputs for one (cfg, n_img); reusable across calls (CUDA-graph safe)."""

    def __init__(self, cfg, n_img, device):
        lib = L.load()
        self.cfg, self.n_img, self.device = cfg, n_img, device
        self.M = num_candidates(cfg)
        nbytes = lib.iou_postproc_workspace_bytes(ctypes.byref(cfg), n_img)
        if nbytes == 0:
            L.check(-1)
        self.ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
        self.ws_bytes = nbytes
        # dets | labels | counts live in ONE byte buffer: a step's results leave the GPU as one all-gather / one
        # device->host copy with no pack or cast kernels (dist.PackedGather, detect_stream)
        from .dist import packed_layout, packed_views
        self.packed = torch.zeros(packed_layout(n_img, cfg.max_per_img)[3], dtype=torch.uint8, device=device)
        self.dets, self.labels, self.counts = packed_views(self.packed, n_img, cfg.max_per_img)


def get_bboxes_device(wsp, cls_list, reg_list, iou_list, img_info, rescale, cls_max2=None):
    """Launches the whole get_bboxes pipeline; returns device tensors (dets, labels, counts).
    cls_max2: optional per-level fp32 [n][H][W][A][2] tensors (None entries allowed) whose last-axis max is the anchor's
    max class logit (what the retina_cls conv epilogue writes, Engine.cls_max2): the class maps are then only read for
    the candidates (iou_get_bboxes_premax)."""
    lib = L.load()
    cls_list = [nhwc_rows(t) for t in cls_list]
    reg_list = [nhwc_rows(t) for t in reg_list]
    iou_list = [nhwc_rows(t) for t in iou_list] if iou_list is not None else None
    pm = None
    if cls_max2 is not None:
        assert len(cls_max2) == len(cls_list)
        pm = (ctypes.c_void_p * len(cls_max2))(*[(t.data_ptr() if t is not None else None) for t in cls_max2])
    L.check(lib.iou_get_bboxes_premax(ctypes.byref(wsp.cfg), wsp.n_img, _ptr_array(cls_list), _ptr_array(reg_list),
                                      _ptr_array(iou_list) if iou_list is not None else None, pm, img_info.data_ptr(),
                                      int(bool(rescale)),
                                      wsp.dets.data_ptr(), wsp.labels.data_ptr(), wsp.counts.data_ptr(),
                                      wsp.ws.data_ptr(), wsp.ws_bytes, L.stream_ptr()))
    L.launch_count += 5
    return wsp.dets, wsp.labels, wsp.counts


def decode_candidates(wsp, cls_list, reg_list, iou_list, img_info, rescale):
    """Stage 1 only: (boxes [n,M,4], scores_cm [n,C,M], cand_idx [n,M])."""
    lib = L.load()
    cfg, n = wsp.cfg, wsp.n_img
    cls_list = [nhwc_rows(t) for t in cls_list]
    reg_list = [nhwc_rows(t) for t in reg_list]
    iou_list = [nhwc_rows(t) for t in iou_list] if iou_list is not None else None     # None: alpha == 1 heads
    boxes = torch.empty(n, wsp.M, 4, dtype=torch.float32, device=wsp.device)
    scores = torch.empty(n, cfg.num_classes, wsp.M, dtype=torch.float32, device=wsp.device)
    idx = torch.empty(n, wsp.M, dtype=torch.int32, device=wsp.device)
    L.check(lib.iou_decode_candidates(ctypes.byref(cfg), n, _ptr_array(cls_list), _ptr_array(reg_list),
                                      _ptr_array(iou_list) if iou_list is not None else None,
                                      img_info.data_ptr(), int(bool(rescale)),
                                      boxes.data_ptr(), scores.data_ptr(), idx.data_ptr(),
                                      wsp.ws.data_ptr(), wsp.ws_bytes, L.stream_ptr()))
    L.launch_count += 3
    return boxes, scores, idx


def batched_nms(wsp, boxes, scores_cm):
    """Stage 2 only: boxes [n,M,4], scores_cm [n,C,M] -> (dets, labels, counts)."""
    lib = L.load()
    boxes, scores_cm = boxes.contiguous(), scores_cm.contiguous()
    L.check(lib.iou_batched_nms(ctypes.byref(wsp.cfg), wsp.n_img, boxes.data_ptr(), scores_cm.data_ptr(),
                                wsp.dets.data_ptr(), wsp.labels.data_ptr(), wsp.counts.data_ptr(),
                                wsp.ws.data_ptr(), wsp.ws_bytes, L.stream_ptr()))
    L.launch_count += 2
    return wsp.dets, wsp.labels, wsp.counts


SOFT_NMS_METHODS = {'linear': 1, 'gaussian': 2}


def batched_soft_nms(wsp, boxes, scores_cm, method=1, sigma=0.5, min_score=1e-3):
    """Stage 2 with nms type 'soft_nms': same buffers as batched_nms (wsp.cfg.iou_thr is the soft-NMS iou_thr)."""
    lib = L.load()
    boxes, scores_cm = boxes.contiguous(), scores_cm.contiguous()
    L.check(lib.iou_batched_soft_nms(ctypes.byref(wsp.cfg), wsp.n_img, boxes.data_ptr(), scores_cm.data_ptr(),
                                     int(method), float(sigma), float(min_score),
                                     wsp.dets.data_ptr(), wsp.labels.data_ptr(), wsp.counts.data_ptr(),
                                     wsp.ws.data_ptr(), wsp.ws_bytes, L.stream_ptr()))
    L.launch_count += 2
    return wsp.dets, wsp.labels, wsp.counts


def soft_nms_cuda(dets, iou_thr, method=1, sigma=0.5, min_score=1e-3):
    """Device counterpart of soft_nms_cpu.soft_nms_cpu (ops/nms/src/soft_nms_cpu.pyx:22-127): returns
    (new_dets (k,5) with decayed scores in selection order, inds (k,) int64), both on dets.device."""
    if not dets.is_cuda:
        raise RuntimeError("soft_nms_cuda: dets must be a CUDA tensor")
    d = dets.detach().float().contiguous()
    n = d.shape[0]
    out = torch.empty(n, 5, dtype=torch.float32, device=d.device)
    inds = torch.empty(n, dtype=torch.int64, device=d.device)
    cnt = torch.zeros(1, dtype=torch.int32, device=d.device)
    if n > 0:
        with torch.cuda.device(d.device):
            L.check(L.load().iou_soft_nms(d.data_ptr(), n, float(iou_thr), int(method), float(sigma),
                                          float(min_score), out.data_ptr(), inds.data_ptr(), cnt.data_ptr(),
                                          L.stream_ptr()))
        L.launch_count += 1
    k = int(cnt.item())
    return out[:k], inds[:k]


def split_results(dets, labels, counts):
    """One D2H sync: list[(Tensor(k,5), Tensor(k,))] as get_bboxes returns (:461)."""
    cnt = counts.cpu().tolist()
    return [(dets[i, :k], labels[i, :k]) for i, k in enumerate(cnt)]


def nms_cuda(dets, iou_thr):
    """Drop-in for mmdet.ops.nms.nms_cuda.nms (ops/nms/src/nms_cuda.cpp:8-13)."""
    if not dets.is_cuda:
        raise RuntimeError("nms_cuda: dets must be a CUDA tensor")
    if dets.numel() == 0:
        return torch.empty(0, dtype=torch.long, device="cpu")     # nms_cuda.cpp:10-11
    lib = L.load()
    d = dets.detach().float().contiguous()
    n = d.shape[0]
    keep = torch.empty(n, dtype=torch.int64, device=d.device)
    cnt = torch.zeros(1, dtype=torch.int32, device=d.device)
    ws_bytes = lib.iou_nms_workspace_bytes(n)          # 256 B up to 6144 boxes; ~n*n/8 B of mask words above
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=d.device)
    with torch.cuda.device(d.device):
        L.check(lib.iou_nms(d.data_ptr(), n, float(iou_thr), keep.data_ptr(), cnt.data_ptr(), ws.data_ptr(),
                            ws_bytes, L.stream_ptr()))
    L.launch_count += 1
    return keep[:int(cnt.item())]
