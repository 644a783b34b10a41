"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the pre-processing row.

numpy restatement of ``ImageTransform.__call__`` without its resize step
(mmdet/datasets/transforms.py:31-50).  The arithmetic lives in the third-party dependency mmcv
(>=0.2.6 per the reference's setup.py:108-111, 0.2.8 per INSTALL.md:16-20), which is absent from
/root/reference; its published algorithm is restated here:
  imnormalize(img, mean, std, to_rgb): img.astype(float32); BGR->RGB if to_rgb; (img - mean) / std
  imflip(img): horizontal flip;  impad_to_multiple(img, d): zero-pad bottom/right to multiples of d
PARITY UNPINNED for this row: the reference has no tests and mmcv / cv2 are not installed, so there is
no reference output to pin against (the formulas above are float32 numpy semantics).
"""
import numpy as np


def image_transform(img_u8, mean, std, to_rgb=True, size_divisor=None, flip=False):
    """img_u8: (h, w, 3) uint8 BGR -> (3, hp, wp) float32, plus img_shape and pad_shape."""
    mean = np.array(mean, dtype=np.float32)
    std = np.array(std, dtype=np.float32)
    img = img_u8.astype(np.float32)
    if to_rgb:
        img = img[..., ::-1]
    img = (img - mean) / std
    if flip:
        img = img[:, ::-1]
    img_shape = img.shape
    if size_divisor is not None:
        hp = int(np.ceil(img.shape[0] / size_divisor)) * size_divisor
        wp = int(np.ceil(img.shape[1] / size_divisor)) * size_divisor
        out = np.zeros((hp, wp, 3), dtype=np.float32)
        out[:img.shape[0], :img.shape[1]] = img
        img = out
    return np.ascontiguousarray(img.transpose(2, 0, 1)), img_shape, img.shape
