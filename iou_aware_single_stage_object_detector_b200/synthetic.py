"""Seeded synthetic weights / inputs for benchmarks (no datasets or checkpoints exist offline).

``spread_state_dict_`` re-draws BN statistics and the head so that logits spread like a trained
network's instead of the degenerate reference init where every score is
0.0710; the three output convs are calibrated on a probe image through ``forward_fn``.
"""
import torch


def spread_state_dict_(sd, forward_fn, seed=1, targets=(2.0, 0.5, 1.5), cls_bias=-3.0):
    """In place on a CPU state_dict.  forward_fn(sd, probe) -> (cls list, reg list, iou list)."""
    g = torch.Generator().manual_seed(seed)
    for k in sorted(sd.keys()):
        v = sd[k]
        is_bn = (".bn" in k) or ("downsample.1" in k) or k.startswith("backbone.bn1")
        if k.endswith("running_var"):
            v.copy_(torch.rand(v.shape, generator=g) + 0.5)
        elif k.endswith("running_mean"):
            v.copy_(torch.randn(v.shape, generator=g) * 0.1)
        elif is_bn and k.endswith(".weight"):
            v.copy_(torch.rand(v.shape, generator=g) * 0.3 + 0.1 if ".bn3." in k
                    else torch.rand(v.shape, generator=g) + 0.5)
        elif is_bn and k.endswith(".bias"):
            v.copy_(torch.randn(v.shape, generator=g) * 0.1)
    for i in range(4):
        for t in ("cls_convs", "reg_convs"):
            k = "bbox_head.%s.%d.conv.weight" % (t, i)
            sd[k].copy_(torch.randn(sd[k].shape, generator=g) * (2.0 / (sd[k].shape[1] * 9)) ** 0.5)
    for name in ("retina_cls", "retina_reg", "retina_iou"):
        sd["bbox_head.%s.weight" % name].copy_(torch.randn(sd["bbox_head.%s.weight" % name].shape, generator=g) * 0.01)
        sd["bbox_head.%s.bias" % name].zero_()
    probe = torch.randn(1, 3, 128, 160, generator=g)
    cls, reg, iou = forward_fn(sd, probe)
    for name, maps, tgt in (("retina_cls", cls, targets[0]), ("retina_reg", reg, targets[1]),
                            ("retina_iou", iou, targets[2])):
        std = torch.cat([m.reshape(-1).float().cpu() for m in maps]).std().item()
        sd["bbox_head.%s.weight" % name].mul_(tgt / max(std, 1e-12))
    sd["bbox_head.retina_cls.bias"].fill_(cls_bias)
    return sd


def cuda_forward_fn(det, device):
    """forward_fn that runs the probe through the CUDA path of `det` (weights loaded from sd)."""
    def fn(sd, probe):
        det.load_state_dict(sd)
        det.to(device)
        feats = det.extract_feat(probe.to(device))
        return det.bbox_head(feats)
    return fn


def synthetic_batch(n, h=800, w=1344, seed=0, pin=False):
    """Image batch with ~ImageNet-normalised statistics and the img_meta of a 1333x800 image padded to
    a multiple of 32 (mmdet/datasets/transforms.py:43-46)."""
    g = torch.Generator().manual_seed(seed)
    img = torch.randn(n, 3, h, w, generator=g)
    if pin:
        img = img.pin_memory()
    real_w = w - 11 if w == 1344 else w
    meta = dict(ori_shape=(h, real_w, 3), img_shape=(h, real_w, 3), pad_shape=(h, w, 3), scale_factor=1.0,
                flip=False)
    return img, [dict(meta) for _ in range(n)]
