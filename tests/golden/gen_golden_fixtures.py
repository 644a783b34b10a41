"""Input builders shared by gen_golden.py (needs the reference) and the tests (do not)."""
import numpy as np

import cases


def results_fixture():
    """A small synthetic result list (3 images x 80 classes) + the dataset attributes det2json reads."""
    rs = np.random.RandomState(77)
    results = []
    for img in range(3):
        per_class = []
        for c in range(80):
            k = int(rs.randint(0, 3)) if (c + img) % 7 == 0 else 0
            b = cases.random_dets(rs, k, 300, 100) if k else np.zeros((0, 5), np.float32)
            per_class.append(b.astype(np.float32))
        results.append(per_class)

    class DS(object):
        img_ids = [139, 285, 632]
        cat_ids = [i * 2 + 1 for i in range(80)]

        def __len__(self):
            return 3
    return DS(), results


def resize_cases():
    """name -> (uint8 BGR frame (h, w, 3), (dst_w, dst_h)): up- and down-scaling, odd sizes, the 2x special case."""
    rs = np.random.RandomState(31)
    out = {}
    for name, (sh, sw), (dw, dh) in [("up_60x80", (60, 80), (133, 100)), ("down_97x131", (97, 131), (53, 37)),
                                     ("up2x_32x48", (32, 48), (96, 64)), ("down2x_64x64", (64, 64), (32, 32)),
                                     ("aspect_50x120", (50, 120), (77, 201)), ("row_1x40", (1, 40), (90, 3))]:
        out[name] = (rs.randint(0, 256, (sh, sw, 3)).astype(np.uint8), (dw, dh))
    return out


IMG_NORM = dict(mean=[123.675, 116.28, 103.53], std=[58.395, 57.12, 57.375], to_rgb=True)   # configs/...r50_4gpu.py:75-76


def test_item_cases():
    """name -> dict(frame uint8 BGR (h, w, 3), img_info, ann, img_scales, flip_ratio, resize_keep_ratio) for
    CustomDataset.prepare_test_img (mmdet/datasets/custom.py:283-359)."""
    rs = np.random.RandomState(2025)

    def ann(n, h, w):
        xy = rs.rand(n, 2) * [w * 0.7, h * 0.7]
        wh = rs.rand(n, 2) * [w * 0.4, h * 0.4] + 2
        return dict(bboxes=np.concatenate([xy, xy + wh], axis=1).astype(np.float32),
                    labels=rs.randint(1, 81, n).astype(np.int64))
    out = {}
    for name, (h, w), scales, flip_ratio, keep in [
            ("single_scale", (120, 173), [(333, 200)], 0, True),            # the IoU-aware configs' layout (flip_ratio 0)
            ("with_flip", (97, 131), [(333, 200)], 0.5, True),
            ("multi_scale_flip", (60, 80), [(200, 120), (133, 100)], 0.5, True),
            ("no_keep_ratio", (50, 120), [(201, 77)], 0, False)]:
        out[name] = dict(frame=rs.randint(0, 256, (h, w, 3)).astype(np.uint8), img_info=dict(filename=name + ".jpg", height=h, width=w),
                         ann=ann(5, h, w), img_scales=scales, flip_ratio=flip_ratio, resize_keep_ratio=keep)
    return out
