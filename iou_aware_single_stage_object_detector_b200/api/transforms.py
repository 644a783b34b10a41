"""Box codec helpers with the reference's signatures (mmdet/core/bbox/transforms.py:44-78,148-166).

``delta2bbox`` is fused into the gather/decode CUDA kernel on the inference path; this
device-agnostic torch statement exists for API parity and for callers outside that path.
"""
import numpy as np
import torch


def delta2bbox(rois, deltas, means=[0, 0, 0, 0], stds=[1, 1, 1, 1], max_shape=None,
               wh_ratio_clip=16 / 1000):
    k = deltas.size(1) // 4
    mean = deltas.new_tensor(means).repeat(1, k)
    std = deltas.new_tensor(stds).repeat(1, k)
    d = deltas * std + mean
    limit = abs(np.log(wh_ratio_clip))
    dx, dy = d[:, 0::4], d[:, 1::4]
    dw = d[:, 2::4].clamp(min=-limit, max=limit)
    dh = d[:, 3::4].clamp(min=-limit, max=limit)
    pw = (rois[:, 2] - rois[:, 0] + 1.0).unsqueeze(1).expand_as(dw)
    ph = (rois[:, 3] - rois[:, 1] + 1.0).unsqueeze(1).expand_as(dh)
    px = ((rois[:, 0] + rois[:, 2]) * 0.5).unsqueeze(1).expand_as(dx)
    py = ((rois[:, 1] + rois[:, 3]) * 0.5).unsqueeze(1).expand_as(dy)
    gw, gh = pw * dw.exp(), ph * dh.exp()
    gx, gy = torch.addcmul(px, pw, dx), torch.addcmul(py, ph, dy)
    x1, y1 = gx - gw * 0.5 + 0.5, gy - gh * 0.5 + 0.5
    x2, y2 = gx + gw * 0.5 - 0.5, gy + gh * 0.5 - 0.5
    if max_shape is not None:
        x1, x2 = x1.clamp(min=0, max=max_shape[1] - 1), x2.clamp(min=0, max=max_shape[1] - 1)
        y1, y2 = y1.clamp(min=0, max=max_shape[0] - 1), y2.clamp(min=0, max=max_shape[0] - 1)
    return torch.stack([x1, y1, x2, y2], dim=-1).view_as(deltas)


def bbox2result(bboxes, labels, num_classes):
    """(k,5),(k,) -> list of num_classes-1 float32 arrays; one GPU->host sync."""
    if bboxes.shape[0] == 0:
        return [np.zeros((0, 5), dtype=np.float32) for _ in range(num_classes - 1)]
    b = bboxes.detach().cpu().numpy()
    lab = labels.detach().cpu().numpy()
    return [b[lab == i, :] for i in range(num_classes - 1)]


def multi_apply(func, *args, **kwargs):
    """mmdet/core/utils/misc.py:21-24."""
    from functools import partial
    pfunc = partial(func, **kwargs) if kwargs else func
    return tuple(map(list, zip(*map(pfunc, *args))))
