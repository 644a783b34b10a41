"""Test-time data items with the reference's layout (mmdet/datasets/custom.py:283-359, transforms.py:53-104).

``prepare_test_img`` builds what ``CustomDataset.prepare_test_img`` hands to ``forward_test``:
``dict(img=[Tensor(3, hp, wp)], img_meta=[DataContainer(dict)], gt_bboxes=[DataContainer(Tensor)],
gt_labels=[DataContainer(Tensor)])`` -- one entry per test scale (plus a flipped one when ``flip_ratio > 0``), the
fork's ``gt_bboxes`` / ``gt_labels`` lists included (``custom.py:325-356``).  The image arithmetic (rescale,
normalise, flip, pad, CHW) runs on the GPU in one kernel (api.ImageTransform); boxes and metas are host-side numpy,
as in the reference.
"""
import numpy as np
import torch


class DataContainer(object):
    """The fields of mmcv.parallel.DataContainer that the item layout uses (a tagged box around ``data``)."""

    def __init__(self, data, stack=False, padding_value=0, cpu_only=False, pad_dims=2):
        self._data, self._stack, self._padding_value = data, stack, padding_value
        self._cpu_only, self._pad_dims = cpu_only, pad_dims

    data = property(lambda self: self._data)
    stack = property(lambda self: self._stack)
    padding_value = property(lambda self: self._padding_value)
    cpu_only = property(lambda self: self._cpu_only)
    pad_dims = property(lambda self: self._pad_dims)

    def __repr__(self):
        return '{}({})'.format(self.__class__.__name__, repr(self._data))


def to_tensor(data):
    """mmdet/datasets/utils.py:15-33."""
    if isinstance(data, torch.Tensor):
        return data
    if isinstance(data, np.ndarray):
        return torch.from_numpy(data)
    if isinstance(data, (list, tuple)):
        return torch.tensor(data)
    if isinstance(data, int):
        return torch.LongTensor([data])
    if isinstance(data, float):
        return torch.FloatTensor([data])
    raise TypeError('type {} cannot be converted to tensor.'.format(type(data)))


def bbox_flip(bboxes, img_shape):
    """Horizontal flip of (..., 4k) boxes inside an image of img_shape = (h, w, ...) (transforms.py:53-65)."""
    assert bboxes.shape[-1] % 4 == 0
    w = img_shape[1]
    out = bboxes.copy()
    out[..., 0::4] = w - bboxes[..., 2::4] - 1
    out[..., 2::4] = w - bboxes[..., 0::4] - 1
    return out


class BboxTransform(object):
    """Scale, optionally flip, clip to the image and optionally pad gt boxes (transforms.py:68-104)."""

    def __init__(self, max_num_gts=None):
        self.max_num_gts = max_num_gts

    def __call__(self, bboxes, img_shape, scale_factor, flip=False):
        b = bboxes * scale_factor
        if flip:
            b = bbox_flip(b, img_shape)
        b[:, 0::2] = np.clip(b[:, 0::2], 0, img_shape[1] - 1)
        b[:, 1::2] = np.clip(b[:, 1::2], 0, img_shape[0] - 1)
        if self.max_num_gts is None:
            return b
        padded = np.zeros((self.max_num_gts, 4), dtype=np.float32)
        padded[:b.shape[0], :] = b
        return padded


def prepare_test_img(frame, img_info, ann, img_transform, bbox_transform=None, img_scales=((1333, 800),),
                     flip_ratio=0, resize_keep_ratio=True, device="cuda"):
    """One test item.  frame: uint8 BGR (h, w, 3) ndarray or tensor (what mmcv.imread returns, custom.py:286);
    img_info: dict(height, width); ann: dict(bboxes (k, 4) float32, labels (k,) int64) as get_ann_info returns.
    Entry order follows the reference loop (custom.py:333-351): per scale the plain image, then -- if
    flip_ratio > 0 -- its flipped twin; gt lists get ONE entry per scale.  Like the reference (:341-342), the boxes
    of scale i+1 are transformed from the already transformed boxes of scale i."""
    bbox_transform = bbox_transform or BboxTransform()
    if isinstance(frame, np.ndarray):
        frame = torch.from_numpy(np.ascontiguousarray(frame))
    frame = frame.to(device).unsqueeze(0)
    imgs, metas, gtb_list, gtl_list = [], [], [], []
    gt_bboxes, gt_labels = ann['bboxes'], ann['labels']

    def single(scale, flip):
        out, img_shape, pad_shape, factor = img_transform(frame, scale=scale, flip=flip, keep_ratio=resize_keep_ratio)
        meta = dict(ori_shape=(img_info['height'], img_info['width'], 3), img_shape=img_shape, pad_shape=pad_shape,
                    scale_factor=factor, flip=flip)
        return out[0], meta

    for scale in img_scales:
        img, meta = single(scale, False)
        gt_bboxes = bbox_transform(gt_bboxes, meta['img_shape'], meta['scale_factor'], flip=False)
        gtb_list.append(DataContainer(to_tensor(gt_bboxes)))
        gtl_list.append(DataContainer(to_tensor(gt_labels)))
        imgs.append(img)
        metas.append(DataContainer(meta, cpu_only=True))
        if flip_ratio > 0:
            img, meta = single(scale, True)
            imgs.append(img)
            metas.append(DataContainer(meta, cpu_only=True))
    return dict(img=imgs, img_meta=metas, gt_bboxes=gtb_list, gt_labels=gtl_list)
