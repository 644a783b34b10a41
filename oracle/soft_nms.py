"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's Soft-NMS.

Follows mmdet/ops/nms/src/soft_nms_cpu.pyx:22-127 (``soft_nms_cpu``) statement by statement, in numpy
float32/float64 exactly as the C generated from the .pyx evaluates it (see ``_rescore``; ``np.exp`` is
evaluated in double on the float argument and rounded back to float, :106), including its in-place bookkeeping:

  * selection (:39-72): the maximum score among positions [i, N) with a strict ``<`` (:52), i.e. the FIRST
    maximum wins; rows i and maxpos are swapped together with their original indices;
  * rescoring (:82-112): only boxes that overlap the selected one (iw > 0 and ih > 0, legacy +1 widths)
    are touched; weight = 1 - ov if ov > iou_thr (linear, :100-104), exp(-ov*ov/sigma) (gaussian, :105-106),
    0/1 (anything else, :107-111);
  * removal (:116-124): a rescored box whose score drops below ``min_score`` is overwritten by the LAST
    live box, N shrinks, and the moved box is examined at the same position.

Pinned bit-for-bit to the reference's own .pyx compiled with Cython (oracle/build_ref.py ->
oracle/_ref/soft_nms_cpu*.so) by tests/golden/soft_nms.npz and tests/test_oracle_golden.py.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline may import this module.
"""
import numpy as np

F = np.float32
METHOD_CODES = {'linear': 1, 'gaussian': 2}


def _rescore(sel, boxes, method, iou_thr, sigma):
    """New scores and the 'touched' mask for `boxes` (m,5) against the selected box `sel` (5,).

    Precision follows the C that Cython generates from the .pyx: coordinate differences are C ``float``
    subtractions, but the literal in ``+ 1`` is emitted as the double ``1.0``, so ``(x2 - x1 + 1) * (y2 - y1 + 1)``
    and the whole ``ua`` expression are evaluated in double and rounded to float once on assignment; ``iw * ih``
    and ``ov = iw * ih / ua`` are float operations; ``1 - ov`` is a double subtraction rounded to float."""
    f32, f64 = np.float32, np.float64
    tx1, ty1, tx2, ty2 = sel[0], sel[1], sel[2], sel[3]
    x1, y1, x2, y2, s = boxes[:, 0], boxes[:, 1], boxes[:, 2], boxes[:, 3], boxes[:, 4]
    area = (((x2 - x1).astype(f64) + 1.0) * ((y2 - y1).astype(f64) + 1.0)).astype(f32)          # :90
    iw = ((np.minimum(tx2, x2) - np.maximum(tx1, x1)).astype(f64) + 1.0).astype(f32)            # :91
    ih = ((np.minimum(ty2, y2) - np.maximum(ty1, y1)).astype(f64) + 1.0).astype(f32)            # :93
    touched = (iw > 0) & (ih > 0)                                                               # :92,94
    with np.errstate(divide='ignore', invalid='ignore'):
        inter = (iw * ih).astype(f32)                                                           # float x float
        sel_area = (f64(tx2 - tx1) + 1.0) * (f64(ty2 - ty1) + 1.0)
        ua = ((sel_area + area.astype(f64)) - inter.astype(f64)).astype(f32)                    # :95
        ov = (inter / ua).astype(f32)                                                           # :96
        if method == 1:
            weight = np.where(ov > f32(iou_thr), (1.0 - ov.astype(f64)).astype(f32), f32(1.0))
        elif method == 2:
            arg = ((-(ov * ov)).astype(f32) / f32(sigma)).astype(f32)
            weight = np.exp(arg.astype(f64)).astype(f32)                   # np.exp on a Python float, :106
        else:
            weight = np.where(ov > f32(iou_thr), f32(0.0), f32(1.0))
        new_s = (weight.astype(f32) * s).astype(f32)                       # :113
    return np.where(touched, new_s, s).astype(f32), touched


def soft_nms_cpu(boxes_in, iou_thr, method=1, sigma=0.5, min_score=0.001):
    """(boxes[:N] with decayed scores in selection order, original indices) -- soft_nms_cpu.pyx:22-127."""
    boxes = np.array(boxes_in, dtype=np.float32, copy=True)
    N = boxes.shape[0]
    inds = np.arange(N)
    min_score = F(min_score)
    i = 0
    while i < N:
        maxpos = i + int(np.argmax(boxes[i:N, 4]))                         # first maximum (strict <, :52)
        boxes[[i, maxpos]] = boxes[[maxpos, i]]
        inds[[i, maxpos]] = inds[[maxpos, i]]
        if i + 1 < N:
            new_s, touched = _rescore(boxes[i], boxes[i + 1:N], method, iou_thr, sigma)
            boxes[i + 1:N, 4] = new_s
            drop = touched & (new_s < min_score)                           # :116 (only rescored boxes)
            if drop.any():
                flags = np.zeros(N, dtype=bool)
                flags[i + 1:N] = drop
                pos = i + 1
                while pos < N:                                             # :116-124, literally
                    if flags[pos]:
                        boxes[pos] = boxes[N - 1]
                        inds[pos] = inds[N - 1]
                        flags[pos] = flags[N - 1]
                        N -= 1
                    else:
                        pos += 1
        i += 1
    return boxes[:N].copy(), inds[:N].copy()


def soft_nms(dets, iou_thr, method='linear', sigma=0.5, min_score=1e-3):
    """nms_wrapper.soft_nms (mmdet/ops/nms/nms_wrapper.py:52-78) on numpy input."""
    if method not in METHOD_CODES:
        raise ValueError('Invalid method for SoftNMS: {}'.format(method))
    new_dets, inds = soft_nms_cpu(np.asarray(dets, dtype=np.float32), iou_thr, METHOD_CODES[method], sigma, min_score)
    return new_dets.astype(np.float32), inds.astype(np.int64)
