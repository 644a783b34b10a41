"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the dense half of the path.

Functional, state-dict driven fp32 restatement (torch-CPU ATen ops == the
reference's own arithmetic) of:
  * ResNet / ResNeXt forward, eval-mode BN   (mmdet/models/backbones/resnet.py:224-267,
    507-518; resnext.py:21-56),
  * FPN forward                               (mmdet/models/necks/fpn.py:97-136),
  * IoUawareRetinaHead.forward_single         (mmdet/models/anchor_heads/iou_aware_retina_head.py:171-219).
No nn.Modules are built: the functions walk the reference's state_dict keys
(SURVEY.md Appendix A), so the same weights can be handed to the CUDA path.
"""
import torch
import torch.nn.functional as F

STAGE_BLOCKS = {50: (3, 4, 6, 3), 101: (3, 4, 23, 3), 152: (3, 8, 36, 3)}


def _bn(x, sd, p, eps=1e-5):
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"],
                        sd[p + ".weight"], sd[p + ".bias"], False, 0.0, eps)


def backbone_forward(sd, img, depth=50, groups=1, prefix="backbone.", style="pytorch"):
    """resnet.py:507-518.  Returns (C2, C3, C4, C5), NCHW fp32.  style: where a block's stride sits
    (resnet.py:129-134): 'pytorch' in the 3x3 conv2, 'caffe' in the 1x1 conv1."""
    x = F.conv2d(img, sd[prefix + "conv1.weight"], None, stride=2, padding=3)
    x = F.relu(_bn(x, sd, prefix + "bn1"))
    x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)
    outs = []
    for s, nblocks in enumerate(STAGE_BLOCKS[depth]):
        for b in range(nblocks):
            p = "%slayer%d.%d." % (prefix, s + 1, b)
            stride = 2 if (b == 0 and s > 0) else 1
            s1, s2 = (1, stride) if style == "pytorch" else (stride, 1)
            idt = x
            y = F.relu(_bn(F.conv2d(x, sd[p + "conv1.weight"], None, stride=s1), sd, p + "bn1"))
            y = F.conv2d(y, sd[p + "conv2.weight"], None, stride=s2, padding=1, groups=groups)
            y = F.relu(_bn(y, sd, p + "bn2"))
            y = _bn(F.conv2d(y, sd[p + "conv3.weight"]), sd, p + "bn3")
            if (p + "downsample.0.weight") in sd:
                idt = _bn(F.conv2d(x, sd[p + "downsample.0.weight"], None, stride=stride),
                          sd, p + "downsample.1")
            x = F.relu(y + idt)                               # resnet.py:256,265
        outs.append(x)
    return tuple(outs)


def fpn_forward(sd, feats, prefix="neck.", start_level=1, num_outs=5, extra_convs_on_inputs=True,
                relu_before_extra_convs=False):
    """fpn.py:97-136 with add_extra_convs=True, no norm, no activation.  RetinaNet configs: extra levels from C5,
    no ReLU; FCOS configs: extra_convs_on_inputs=False (P6 from the P5 output) and relu_before_extra_convs=True
    (ReLU in front of every extra conv after the first, :123-128)."""
    used = feats[start_level:]
    lat = [F.conv2d(f, sd["%slateral_convs.%d.conv.weight" % (prefix, i)],
                    sd["%slateral_convs.%d.conv.bias" % (prefix, i)]) for i, f in enumerate(used)]
    for i in range(len(lat) - 1, 0, -1):
        lat[i - 1] = lat[i - 1] + F.interpolate(lat[i], scale_factor=2, mode="nearest")
    outs = [F.conv2d(l, sd["%sfpn_convs.%d.conv.weight" % (prefix, i)],
                     sd["%sfpn_convs.%d.conv.bias" % (prefix, i)], padding=1) for i, l in enumerate(lat)]
    n = len(lat)
    src = feats[-1] if extra_convs_on_inputs else outs[-1]
    for i in range(n, num_outs):
        if relu_before_extra_convs and i > n:
            src = F.relu(src)
        outs.append(F.conv2d(src, sd["%sfpn_convs.%d.conv.weight" % (prefix, i)],
                             sd["%sfpn_convs.%d.conv.bias" % (prefix, i)], stride=2, padding=1))
        src = outs[-1]
    return tuple(outs)


def head_forward_single(sd, x, prefix="bbox_head.", stacked=4):
    """iou_aware_retina_head.py:171-219 with shared_conv=4, no feature alignment."""
    c = r = x
    for i in range(stacked):
        c = F.relu(F.conv2d(c, sd["%scls_convs.%d.conv.weight" % (prefix, i)],
                            sd["%scls_convs.%d.conv.bias" % (prefix, i)], padding=1))
        r = F.relu(F.conv2d(r, sd["%sreg_convs.%d.conv.weight" % (prefix, i)],
                            sd["%sreg_convs.%d.conv.bias" % (prefix, i)], padding=1))
    cls = F.conv2d(c, sd[prefix + "retina_cls.weight"], sd[prefix + "retina_cls.bias"], padding=1)
    reg = F.conv2d(r, sd[prefix + "retina_reg.weight"], sd[prefix + "retina_reg.bias"], padding=1)
    iou = F.conv2d(r, sd[prefix + "retina_iou.weight"], sd[prefix + "retina_iou.bias"], padding=1)
    return cls, reg, iou


def head_forward(sd, feats, prefix="bbox_head."):
    """anchor_head.py:102-103 (multi_apply over levels) -> (list, list, list)."""
    outs = [head_forward_single(sd, f, prefix) for f in feats]
    return tuple(map(list, zip(*outs)))


def fcos_head_forward(sd, feats, prefix="bbox_head.", stacked=4, groups=32, eps=1e-5):
    """IoUawareFCOSHead.forward (iou_aware_fcos_head.py:89-113): per level, towers of conv(no bias) -> GroupNorm ->
    ReLU; cls_score / centerness from cls_feat; bbox_pred = exp(scale_l * fcos_reg(reg_feat)); iou from reg_feat.
    Returns (cls list, bbox_pred list, centerness list, iou list)."""
    outs = []
    for l, x in enumerate(feats):
        c = r = x
        for i in range(stacked):
            for tower in ("cls", "reg"):
                k = "%s%s_convs.%d." % (prefix, tower, i)
                t = F.conv2d(c if tower == "cls" else r, sd[k + "conv.weight"], None, padding=1)
                t = F.relu(F.group_norm(t, groups, sd[k + "gn.weight"], sd[k + "gn.bias"], eps))
                if tower == "cls":
                    c = t
                else:
                    r = t
        cls = F.conv2d(c, sd[prefix + "fcos_cls.weight"], sd[prefix + "fcos_cls.bias"], padding=1)
        cen = F.conv2d(c, sd[prefix + "fcos_centerness.weight"], sd[prefix + "fcos_centerness.bias"], padding=1)
        reg = (F.conv2d(r, sd[prefix + "fcos_reg.weight"], sd[prefix + "fcos_reg.bias"], padding=1) *
               sd["%sscales.%d.scale" % (prefix, l)]).exp()
        iou = F.conv2d(r, sd[prefix + "fcos_iou.weight"], sd[prefix + "fcos_iou.bias"], padding=1)
        outs.append((cls, reg, cen, iou))
    return tuple(map(list, zip(*outs)))


@torch.no_grad()
def fcos_detector_forward(sd, img, depth=50):
    """FCOS (configs/fcos/iou_aware_fcos_r50_caffe_fpn_gn_1x_4gpu.py): caffe-style ResNet, FPN with P6 from P5 and
    ReLU before P7, IoUawareFCOSHead.  Returns (cls, bbox_pred, centerness, iou) lists."""
    feats = fpn_forward(sd, backbone_forward(sd, img, depth, 1, style="caffe"), extra_convs_on_inputs=False,
                        relu_before_extra_convs=True)
    return fcos_head_forward(sd, feats)


@torch.no_grad()
def detector_forward(sd, img, depth=50, groups=1):
    """extract_feat + bbox_head (single_stage.py:39-43, 86-87)."""
    feats = fpn_forward(sd, backbone_forward(sd, img, depth, groups))
    return head_forward(sd, feats)


def spread_weights_(sd, seed=1, depth=50, groups=1, targets=(2.0, 0.5, 1.5), cls_bias=-3.0):
    """re-draw BN statistics / affine (un-zeroing bn3) and the head so that
    logits SPREAD (score gaps >> 1e-4) instead of the degenerate reference init, then calibrate the
    three output convs on a fixed probe image so that cls / reg / iou logits have std `targets`
    (a trained-net-like regime).  Operates in place on a state_dict; deterministic in `seed`."""
    g = torch.Generator().manual_seed(seed)
    for k in sorted(sd.keys()):
        v = sd[k]
        is_bn = (".bn" in k) or ("downsample.1" in k) or k.startswith("backbone.bn1")
        if k.endswith("running_var"):
            v.copy_(torch.rand(v.shape, generator=g) + 0.5)
        elif k.endswith("running_mean"):
            v.copy_(torch.randn(v.shape, generator=g) * 0.1)
        elif is_bn and k.endswith(".weight"):
            if ".bn3." in k:      # keep the residual branch modest so depth does not blow the scale up
                v.copy_(torch.rand(v.shape, generator=g) * 0.3 + 0.1)
            else:
                v.copy_(torch.rand(v.shape, generator=g) + 0.5)
        elif is_bn and k.endswith(".bias"):
            v.copy_(torch.randn(v.shape, generator=g) * 0.1)
    for i in range(4):
        for t in ("cls_convs", "reg_convs"):
            k = "bbox_head.%s.%d.conv.weight" % (t, i)
            fan_in = sd[k].shape[1] * 9
            sd[k].copy_(torch.randn(sd[k].shape, generator=g) * (2.0 / fan_in) ** 0.5)
    for name in ("retina_cls", "retina_reg", "retina_iou"):
        k = "bbox_head.%s.weight" % name
        sd[k].copy_(torch.randn(sd[k].shape, generator=g) * 0.01)
        sd["bbox_head.%s.bias" % name].zero_()
    probe = torch.randn(1, 3, 128, 160, generator=g)
    cls, reg, iou = detector_forward(sd, probe, depth, groups)
    for name, maps, tgt in (("retina_cls", cls, targets[0]), ("retina_reg", reg, targets[1]),
                            ("retina_iou", iou, targets[2])):
        std = torch.cat([m.reshape(-1) for m in maps]).std().item()
        sd["bbox_head.%s.weight" % name].mul_(tgt / max(std, 1e-12))
    sd["bbox_head.retina_cls.bias"].fill_(cls_bias)
    return sd


def spread_fcos_weights_(sd, seed=1, depth=50, targets=(2.0, 0.5, 1.5), cls_bias=-3.0, reg_bias=2.5):
    """The FCOS counterpart of spread_weights_: re-draw BN statistics / affine, GroupNorm affine, per-level Scale
    and the head convs, then calibrate fcos_cls / fcos_reg / fcos_iou on a probe image so that the logits spread
    (cls std 2 around -3, reg std 0.5 around 2.5 -> distances ~ 12 px x scale, iou std 1.5).  In place."""
    g = torch.Generator().manual_seed(seed)
    for k in sorted(sd.keys()):
        v = sd[k]
        is_bn = (".bn" in k) or ("downsample.1" in k) or k.startswith("backbone.bn1")
        if k.endswith("running_var"):
            v.copy_(torch.rand(v.shape, generator=g) + 0.5)
        elif k.endswith("running_mean"):
            v.copy_(torch.randn(v.shape, generator=g) * 0.1)
        elif is_bn and k.endswith(".weight"):
            v.copy_(torch.rand(v.shape, generator=g) * 0.3 + 0.1 if ".bn3." in k else torch.rand(v.shape, generator=g) + 0.5)
        elif is_bn and k.endswith(".bias"):
            v.copy_(torch.randn(v.shape, generator=g) * 0.1)
        elif k.endswith(".gn.weight"):
            v.copy_(torch.rand(v.shape, generator=g) + 0.5)
        elif k.endswith(".gn.bias"):
            v.copy_(torch.randn(v.shape, generator=g) * 0.1)
        elif k.endswith(".scale"):
            v.copy_(torch.rand(v.shape, generator=g) * 0.5 + 0.75)
        elif "bbox_head." in k and k.endswith("_convs.%s.conv.weight" % k.split(".")[-3]):
            v.copy_(torch.randn(v.shape, generator=g) * (2.0 / (v.shape[1] * 9)) ** 0.5)
    for name in ("fcos_cls", "fcos_reg", "fcos_iou", "fcos_centerness"):
        sd["bbox_head.%s.weight" % name].copy_(torch.randn(sd["bbox_head.%s.weight" % name].shape, generator=g) * 0.01)
        sd["bbox_head.%s.bias" % name].zero_()
    probe = torch.randn(1, 3, 128, 160, generator=g)
    feats = fpn_forward(sd, backbone_forward(sd, probe, depth, 1, style="caffe"), extra_convs_on_inputs=False,
                        relu_before_extra_convs=True)
    for l in range(len(feats)):
        sd["bbox_head.scales.%d.scale" % l].fill_(1.0)
    cls, reg, cen, iou = fcos_head_forward(sd, feats)
    for name, maps, tgt, is_exp in (("fcos_cls", cls, targets[0], False), ("fcos_reg", reg, targets[1], True),
                                    ("fcos_iou", iou, targets[2], False)):
        flat = torch.cat([(m.log() if is_exp else m).reshape(-1) for m in maps])
        sd["bbox_head.%s.weight" % name].mul_(tgt / max(flat.std().item(), 1e-12))
    sd["bbox_head.fcos_cls.bias"].fill_(cls_bias)
    sd["bbox_head.fcos_reg.bias"].fill_(reg_bias)
    for l in range(len(feats)):
        sd["bbox_head.scales.%d.scale" % l].fill_(0.8 + 0.1 * l)
    return sd
