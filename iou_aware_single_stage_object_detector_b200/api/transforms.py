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


class ImageTransform(object):
    """Device-side ImageTransform (mmdet/datasets/transforms.py:11-50) for images that are ALREADY at
    their test scale: normalize -> (flip) -> pad to size_divisor -> CHW, for a whole uint8 batch in one
    kernel (iou_preprocess_u8).  Rescaling (mmcv.imrescale) is not part of the accelerated path."""

    def __init__(self, mean=(0, 0, 0), std=(1, 1, 1), to_rgb=True, size_divisor=None):
        import numpy as np
        self.mean = np.array(mean, dtype=np.float32)
        self.std = np.array(std, dtype=np.float32)
        self.to_rgb = to_rgb
        self.size_divisor = size_divisor

    def pad_shape(self, h, w):
        d = self.size_divisor
        return (h, w) if d is None else ((h + d - 1) // d * d, (w + d - 1) // d * d)

    def __call__(self, img_u8, flip=False, out=None):
        """img_u8: CUDA uint8 tensor (n, h, w, 3) BGR -> fp32 (n, 3, hp, wp); returns (tensor, img_shape, pad_shape)."""
        import ctypes
        from .. import lib as L
        if not img_u8.is_cuda or img_u8.dtype != torch.uint8 or img_u8.dim() != 4 or img_u8.shape[-1] != 3:
            raise RuntimeError("ImageTransform: expected a CUDA uint8 tensor of shape (n, h, w, 3)")
        img_u8 = img_u8.contiguous()
        n, h, w, _ = img_u8.shape
        hp, wp = self.pad_shape(h, w)
        if out is None:
            out = torch.empty(n, 3, hp, wp, dtype=torch.float32, device=img_u8.device)
        mean = (ctypes.c_float * 3)(*[float(v) for v in self.mean])
        std = (ctypes.c_float * 3)(*[float(v) for v in self.std])
        with torch.cuda.device(img_u8.device):
            L.check(L.load().iou_preprocess_u8(img_u8.data_ptr(), n, h, w, hp, wp, mean, std, int(self.to_rgb),
                                               int(flip), out.data_ptr(), L.stream_ptr()))
        L.launch_count += 1
        return out, (h, w, 3), (hp, wp, 3)
