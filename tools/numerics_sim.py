"""CPU numerics study (no GPU): how far does a split-operand tensor-core scheme drift from exact arithmetic
over the whole dense path (ResNet-50 + FPN + IoUawareRetinaHead, BN folded as the engine does)?

Every scheme stores an activation as a few narrow numbers and computes a conv as a sum of narrow x narrow
products accumulated wide; here the products are formed in float64 from the EXACTLY representable narrow
values, so what is measured is the representation + dropped-term error of the scheme (the tensor core's
own fp32 accumulation error, profiles/r01_v5_accumulation_probe.txt, comes on top and is the same for all).

  bf16x3   : x = hi + lo (bf16, bf16);  conv = hi*Whi + hi*Wlo + lo*Whi           (3 bf16 passes; shipped)
  f16+2f8  : x = h (fp16) + l8 * 2^-11 (e4m3), x8 = e4m3(x);  w as Wh' = fp16(w S), Wl8 = e4m3(w S - Wh'),
             W8 = e4m3(w s) with a per-output-channel power of two s and S = 2^11 s;
             conv = ( h*Wh' + x8*Wl8 + l8*W8 ) / S                                (1 fp16 pass + 2 fp8 passes
             at twice the rate = 2 bf16-pass equivalents, 4 B per element like bf16x3, ONE accumulator)
  f16x1 / bf16x1 : single pass, for scale.

Run:  python tools/numerics_sim.py [H W]
"""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import model as om  # noqa: E402

D = torch.float64


def rnd(x, dt):
    return x.to(torch.float32).to(dt).to(D)


def e4m3(x):
    return rnd(x.clamp(-448.0, 448.0), torch.float8_e4m3fn)


class Exact:
    name = "exact(f64)"

    def act(self, x):
        return (x.to(D),)

    def val(self, r):
        return r[0]

    def wt(self, w):
        return (w.to(D),)

    def conv(self, r, wr, **kw):
        return F.conv2d(r[0], wr[0], None, **kw)


class F32(Exact):
    name = "fp32 storage"

    def act(self, x):
        return (x.to(torch.float32).to(D),)


class Single(Exact):
    def __init__(self, dt, name):
        self.dt, self.name = dt, name

    def act(self, x):
        return (rnd(x, self.dt),)

    def wt(self, w):
        return (rnd(w, self.dt),)


class Bf16x3:
    name = "bf16x3 (shipped)"
    dt = torch.bfloat16

    def act(self, x):
        x = x.to(torch.float32).to(D)
        hi = rnd(x, self.dt)
        return hi, rnd(x - hi, self.dt)

    def val(self, r):
        return r[0] + r[1]

    wt = act

    def conv(self, r, wr, **kw):
        return (F.conv2d(r[0], wr[0], None, **kw) + F.conv2d(r[0], wr[1], None, **kw) +
                F.conv2d(r[1], wr[0], None, **kw))


class F16x3(Bf16x3):
    name = "fp16x3"
    dt = torch.float16


class F16F8:
    """1 fp16 pass + 2 e4m3 passes (K-concatenated: [x8 | l8] x [Wl8 ; W8])."""
    name = "f16+2f8"
    LS = 2.0 ** 11

    def act(self, x):
        x = x.to(torch.float32).to(D)
        h = rnd(x, torch.float16)
        return h, e4m3((x - h) * self.LS), e4m3(x)

    def val(self, r):
        return r[0] + r[1] / self.LS

    def wt(self, w):
        w = w.to(torch.float32).to(D)
        # per-output-channel power-of-two scale: row max of |w| * s in (8, 16]; S = 2^11 * s is shared by all copies
        m = w.abs().flatten(1).max(1).values.clamp_min(1e-30)
        s = torch.exp2(torch.floor(torch.log2(8.0 / m)) + 1.0).view(-1, 1, 1, 1)
        S = s * self.LS
        h = rnd(w * S, torch.float16)
        return h, e4m3(w * S - h), e4m3(w * s), S.view(1, -1, 1, 1)

    def conv(self, r, wr, **kw):
        acc = (F.conv2d(r[0], wr[0], None, **kw) + F.conv2d(r[2], wr[1], None, **kw) +
               F.conv2d(r[1], wr[2], None, **kw))               # one accumulator at scale S
        return acc / wr[3]


class F16F16F8(F16F8):
    """3-byte element (round-2 candidate): x = h (fp16) + l8 * 2^-11, x8 is NOT stored.  conv = h*Wh' + h*Wl' + l8*W8 with
    Wl' = fp16(w S - Wh') (a second fp16 pass) and the e4m3 pass only over l8 (K = 64 per slab): 2.25 bf16-pass equivalents,
    25 % less activation traffic."""
    name = "f16+f16+f8 (3 B)"

    def act(self, x):
        x = x.to(torch.float32).to(D)
        h = rnd(x, torch.float16)
        return h, e4m3((x - h) * self.LS)

    def wt(self, w):
        w = w.to(torch.float32).to(D)
        m = w.abs().flatten(1).max(1).values.clamp_min(1e-30)
        s = torch.exp2(torch.floor(torch.log2(8.0 / m)) + 1.0).view(-1, 1, 1, 1)
        S = s * self.LS
        h = rnd(w * S, torch.float16)
        return h, rnd(w * S - h, torch.float16), e4m3(w * s), S.view(1, -1, 1, 1)

    def conv(self, r, wr, **kw):
        acc = (F.conv2d(r[0], wr[0], None, **kw) + F.conv2d(r[0], wr[1], None, **kw) +
               F.conv2d(r[1], wr[2], None, **kw))
        return acc / wr[3]


def fold_bn(sd, conv, bn, eps=1e-5):
    w = sd[conv + ".weight"].to(D)
    g, b = sd[bn + ".weight"].to(D), sd[bn + ".bias"].to(D)
    m, v = sd[bn + ".running_mean"].to(D), sd[bn + ".running_var"].to(D)
    sc = g / torch.sqrt(v + eps)
    return w * sc.view(-1, 1, 1, 1), b - m * sc


@torch.no_grad()
def forward(sd, img, S, depth=50):
    def conv(xr, w, shift, relu=False, res=None, **kw):
        y = S.conv(xr, S.wt(w), **kw)
        if shift is not None:
            y = y + shift.view(1, -1, 1, 1)
        if res is not None:
            y = y + S.val(res)
        if relu:
            y = F.relu(y)
        return y

    p = "backbone."
    w, b = fold_bn(sd, p + "conv1", p + "bn1")
    x = conv(S.act(img), w, b, relu=True, stride=2, padding=3)
    x = S.act(F.max_pool2d(S.val(S.act(x)), 3, 2, 1))
    feats, amax = [], 0.0
    for s, nb in enumerate(om.STAGE_BLOCKS[depth]):
        for bi in range(nb):
            q = "%slayer%d.%d." % (p, s + 1, bi)
            stride = 2 if (bi == 0 and s > 0) else 1
            idt = x
            w, b = fold_bn(sd, q + "conv1", q + "bn1")
            y = S.act(conv(x, w, b, relu=True))
            w, b = fold_bn(sd, q + "conv2", q + "bn2")
            y = S.act(conv(y, w, b, relu=True, stride=stride, padding=1))
            if (q + "downsample.0.weight") in sd:
                w, b = fold_bn(sd, q + "downsample.0", q + "downsample.1")
                idt = S.act(conv(x, w, b, stride=stride))
            w, b = fold_bn(sd, q + "conv3", q + "bn3")
            x = S.act(conv(y, w, b, relu=True, res=idt))
            amax = max(amax, float(S.val(x).abs().max()))
        feats.append(x)
    used = feats[1:]
    lat = [conv(f, sd["neck.lateral_convs.%d.conv.weight" % i].to(D), sd["neck.lateral_convs.%d.conv.bias" % i].to(D))
           for i, f in enumerate(used)]
    lat[2] = S.act(lat[2])
    for i in (2, 1):
        lat[i - 1] = S.act(lat[i - 1] + F.interpolate(S.val(lat[i]), scale_factor=2, mode="nearest"))
    outs = [S.act(conv(l, sd["neck.fpn_convs.%d.conv.weight" % i].to(D), sd["neck.fpn_convs.%d.conv.bias" % i].to(D),
                       padding=1)) for i, l in enumerate(lat)]
    src = feats[-1]
    for i in (3, 4):
        outs.append(S.act(conv(src, sd["neck.fpn_convs.%d.conv.weight" % i].to(D),
                               sd["neck.fpn_convs.%d.conv.bias" % i].to(D), stride=2, padding=1)))
        src = outs[-1]
    h = "bbox_head."
    res = []
    for f in outs:
        c = r = f
        for i in range(4):
            c = S.act(conv(c, sd["%scls_convs.%d.conv.weight" % (h, i)].to(D), sd["%scls_convs.%d.conv.bias" % (h, i)].to(D),
                           relu=True, padding=1))
            r = S.act(conv(r, sd["%sreg_convs.%d.conv.weight" % (h, i)].to(D), sd["%sreg_convs.%d.conv.bias" % (h, i)].to(D),
                           relu=True, padding=1))
            amax = max(amax, float(S.val(c).abs().max()), float(S.val(r).abs().max()))
        res.append(tuple(conv(t, sd[h + k + ".weight"].to(D), sd[h + k + ".bias"].to(D), padding=1)
                         for t, k in ((c, "retina_cls"), (r, "retina_reg"), (r, "retina_iou"))))
    return res, amax


def main():
    hh, ww = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (256, 320)
    import parity_util as U
    det, _ = U.small_detector()
    sd = {k: v.clone() for k, v in det.state_dict().items()}
    img = torch.randn(1, 3, hh, ww, generator=torch.Generator().manual_seed(3))
    ref, amax = forward(sd, img, Exact())
    print("image %dx%d; max |activation| %.1f; logit range cls [%.2f, %.2f]" %
          (hh, ww, amax, min(float(l[0].min()) for l in ref), max(float(l[0].max()) for l in ref)))
    for S in (F32(), Bf16x3(), F16F8(), F16F16F8(), F16x3(), Single(torch.float16, "fp16x1"),
              Single(torch.bfloat16, "bf16x1")):
        out, _ = forward(sd, img, S)
        line = []
        for j, nm in enumerate(("cls", "reg", "iou")):
            d = torch.cat([(o[j] - r[j]).flatten() for o, r in zip(out, ref)])
            line.append("%s max %.3g rms %.3g" % (nm, float(d.abs().max()), float(d.pow(2).mean().sqrt())))
        # score-level error: sqrt(sigmoid(cls) * sigmoid(iou)) per anchor/class
        ds = 0.0
        for o, r in zip(out, ref):
            def score(t):
                c = torch.sigmoid(t[0]).view(1, 9, 80, *t[0].shape[2:])
                q = torch.sigmoid(t[2]).view(1, 9, 1, *t[2].shape[2:])
                return (c * q).sqrt()
            ds = max(ds, float((score(o) - score(r)).abs().max()))
        print("%-18s %s | max|dscore| %.3g" % (S.name, "; ".join(line), ds))


if __name__ == "__main__":
    main()
