"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the pre-processing row.

numpy restatement of ``ImageTransform.__call__`` (mmdet/datasets/transforms.py:31-50).  The arithmetic lives
in the third-party dependency mmcv (>=0.2.6 per the reference's setup.py:108-111, 0.2.8 per INSTALL.md:16-20),
which is absent from /root/reference; its published algorithm is restated here:
  imrescale(img, scale): factor = min(long / max(h, w), short / min(h, w)); size = int(edge * factor + 0.5);
                         cv2.resize(img, size, interpolation=cv2.INTER_LINEAR)      (imresize: size given)
  imnormalize(img, mean, std, to_rgb): img.astype(float32); BGR->RGB if to_rgb; (img - mean) / std
  imflip(img): horizontal flip;  impad_to_multiple(img, d): zero-pad bottom/right to multiples of d
The resize is OpenCV's: ``resize_linear_u8`` restates the 8-bit fixed-point bilinear path of imgproc/resize.cpp
and IS PINNED -- bit-identical to ``cv2.resize`` (opencv 4.13 in this image) on the goldens of
tests/golden/resize_cv2.npz and on random sizes (tests/test_oracle_golden.py).  The normalise / flip / pad steps and
the whole call sequence are pinned by tests/golden/test_items.npz: the reference's OWN ImageTransform.__call__ and
CustomDataset.prepare_test_img run in the build container on mmcv 0.2.8's image functions restated on cv2 / numpy
(oracle/ref_shim._install_mmcv_image -- mmcv itself is absent), and image_transform_rescaled equals those items bit
for bit.
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


def rescale_size(h, w, scale, keep_ratio=True):
    """mmcv 0.2.8 imrescale / imresize size rule: (new_h, new_w, scale_factor)."""
    if keep_ratio:
        if isinstance(scale, (int, float)):
            factor = float(scale)
        else:
            factor = min(max(scale) / max(h, w), min(scale) / min(h, w))
        return int(h * float(factor) + 0.5), int(w * float(factor) + 0.5), factor
    new_w, new_h = scale
    return int(new_h), int(new_w), np.array([new_w / w, new_h / h, new_w / w, new_h / h], dtype=np.float32)


def _resize_coeffs(dn, sn, scale, zero_frac_at_clamp):
    """resize.cpp: fx = (float)((d + 0.5) * scale - 0.5); s = floor(fx); fx -= s; 11-bit fixed-point weights.
    Columns zero the fraction where s is clamped; rows keep it and clip the row index instead."""
    f0 = ((np.arange(dn, dtype=np.float64) + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f0).astype(np.int64)
    fr = (f0 - s.astype(np.float32)).astype(np.float32)
    if zero_frac_at_clamp:
        lo, hi = s < 0, s >= sn - 1
        fr = np.where(lo | hi, np.float32(0), fr).astype(np.float32)
        s = np.where(lo, 0, np.where(hi, sn - 1, s))
        s0, s1 = s, np.minimum(s + 1, sn - 1)
    else:
        s0, s1 = np.clip(s, 0, sn - 1), np.clip(s + 1, 0, sn - 1)
    w0 = np.rint((np.float32(1.0) - fr) * np.float32(2048)).astype(np.int64)      # saturate_cast<short>(float)
    w1 = np.rint(fr * np.float32(2048)).astype(np.int64)
    return s0, s1, w0, w1


def resize_linear_u8(img, dst_w, dst_h):
    """cv2.resize(img, (dst_w, dst_h), interpolation=cv2.INTER_LINEAR) for uint8 (h, w, c) images."""
    sh, sw = img.shape[:2]
    x0, x1, a0, a1 = _resize_coeffs(dst_w, sw, 1.0 / (dst_w / sw), True)
    y0, y1, b0, b1 = _resize_coeffs(dst_h, sh, 1.0 / (dst_h / sh), False)
    S = img.astype(np.int64).reshape(sh, sw, -1)
    H = S[:, x0, :] * a0[None, :, None] + S[:, x1, :] * a1[None, :, None]
    out = (((b0[:, None, None] * (H[y0] >> 4)) >> 16) + ((b1[:, None, None] * (H[y1] >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8).reshape((dst_h, dst_w) + img.shape[2:])


def image_transform_rescaled(img_u8, scale, mean, std, to_rgb=True, size_divisor=None, flip=False, keep_ratio=True):
    """Full ImageTransform.__call__ (transforms.py:31-50): returns (chw float32, img_shape, pad_shape, scale_factor)."""
    nh, nw, factor = rescale_size(img_u8.shape[0], img_u8.shape[1], scale, keep_ratio)
    chw, img_shape, pad_shape = image_transform(resize_linear_u8(img_u8, nw, nh), mean, std, to_rgb, size_divisor, flip)
    return chw, img_shape, pad_shape, factor
