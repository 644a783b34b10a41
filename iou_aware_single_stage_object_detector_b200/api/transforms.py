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


def rescale_size(h, w, scale, keep_ratio=True):
    """(new_h, new_w, scale_factor) of mmcv 0.2.8 imrescale / imresize as ImageTransform calls them
    (mmdet/datasets/transforms.py:33-40).  keep_ratio: scale = (long, short) edge limits or a number ->
    factor = min(long / max(h, w), short / min(h, w)), size = int(edge * factor + 0.5), scalar scale_factor;
    otherwise scale = (w, h) exactly and scale_factor = [w_scale, h_scale, w_scale, h_scale]."""
    import numpy as np
    if keep_ratio:
        if isinstance(scale, (int, float)):
            if scale <= 0:
                raise ValueError('Invalid scale {}, must be positive.'.format(scale))
            factor = float(scale)
        elif isinstance(scale, tuple):
            factor = min(max(scale) / max(h, w), min(scale) / min(h, w))
        else:
            raise TypeError('Scale must be a number or tuple of int, but got {}'.format(type(scale)))
        return int(h * float(factor) + 0.5), int(w * float(factor) + 0.5), factor
    new_w, new_h = scale
    w_scale, h_scale = new_w / w, new_h / h
    return int(new_h), int(new_w), np.array([w_scale, h_scale, w_scale, h_scale], dtype=np.float32)


class ImageTransform(object):
    """Device-side ImageTransform (mmdet/datasets/transforms.py:11-50): (rescale ->) normalize -> (flip) -> pad
    to size_divisor -> CHW, for a whole uint8 batch in ONE kernel.  Without `scale` the frames are taken as
    already at their test scale (iou_preprocess_u8); with `scale` they are resized first with OpenCV's 8-bit
    bilinear arithmetic, bit-identical to the cv2.resize call inside mmcv.imrescale (iou_preprocess_resize_u8)."""

    def __init__(self, mean=(0, 0, 0), std=(1, 1, 1), to_rgb=True, size_divisor=None, scale=None, keep_ratio=True):
        import numpy as np
        self.mean = np.array(mean, dtype=np.float32)
        self.std = np.array(std, dtype=np.float32)
        self.to_rgb = to_rgb
        self.size_divisor = size_divisor
        self.scale, self.keep_ratio = scale, keep_ratio

    def out_shape(self, h, w, scale=None, keep_ratio=None):
        """(img_h, img_w, scale_factor) after the optional rescale of an h x w frame."""
        scale = self.scale if scale is None else scale
        keep_ratio = self.keep_ratio if keep_ratio is None else keep_ratio
        if scale is None:
            return h, w, 1.0
        return rescale_size(h, w, scale, keep_ratio)

    def pad_shape(self, h, w, scale=None, keep_ratio=None):
        """Padded (hp, wp) of the transformed h x w frame."""
        h, w, _ = self.out_shape(h, w, scale, keep_ratio)
        d = self.size_divisor
        return (h, w) if d is None else ((h + d - 1) // d * d, (w + d - 1) // d * d)

    def __call__(self, img_u8, scale=None, flip=False, keep_ratio=None, out=None):
        """img_u8: CUDA uint8 tensor (n, h, w, 3) BGR -> fp32 (n, 3, hp, wp).  Returns (tensor, img_shape,
        pad_shape) and, when a scale is in effect, scale_factor as a 4th element (the reference's return order,
        transforms.py:50)."""
        import ctypes
        from .. import lib as L
        if not img_u8.is_cuda or img_u8.dtype != torch.uint8 or img_u8.dim() != 4 or img_u8.shape[-1] != 3:
            raise RuntimeError("ImageTransform: expected a CUDA uint8 tensor of shape (n, h, w, 3)")
        img_u8 = img_u8.contiguous()
        n, h, w, _ = img_u8.shape
        scale = self.scale if scale is None else scale
        nh, nw, factor = self.out_shape(h, w, scale, keep_ratio)
        hp, wp = self.pad_shape(h, w, scale, keep_ratio)
        if out is None:
            out = torch.empty(n, 3, hp, wp, dtype=torch.float32, device=img_u8.device)
        elif tuple(out.shape) != (n, 3, hp, wp):
            raise RuntimeError("ImageTransform: out has shape %s, expected %s" % (tuple(out.shape), (n, 3, hp, wp)))
        mean = (ctypes.c_float * 3)(*[float(v) for v in self.mean])
        std = (ctypes.c_float * 3)(*[float(v) for v in self.std])
        with torch.cuda.device(img_u8.device):
            if scale is None:
                L.check(L.load().iou_preprocess_u8(img_u8.data_ptr(), n, h, w, hp, wp, mean, std, int(self.to_rgb),
                                                   int(flip), out.data_ptr(), L.stream_ptr()))
            else:
                L.check(L.load().iou_preprocess_resize_u8(img_u8.data_ptr(), n, h, w, nh, nw, hp, wp, mean, std,
                                                          int(self.to_rgb), int(flip), out.data_ptr(),
                                                          L.stream_ptr()))
        L.launch_count += 1
        if scale is None:
            return out, (h, w, 3), (hp, wp, 3)
        return out, (nh, nw, 3), (hp, wp, 3), factor
