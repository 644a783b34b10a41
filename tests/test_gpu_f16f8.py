"""-m gpu tests of conv `passes = 2` (IOU_FMT_F16F8: one fp16 tensor-core pass + one e4m3 pass of doubled K,
csrc/split_fmt.cuh): the layout kernels' encode/decode bit for bit against oracle/split_fmt.py, the conv engine
against plain PyTorch fp32 (CPU) convs, the whole detector against the CPU oracle at the north-star tolerance."""
import pytest
import torch
import torch.nn.functional as F

import parity_util as U  # noqa: F401  (sets sys.path)
from oracle import split_fmt as SF
from iou_aware_single_stage_object_detector_b200 import engine as E
from iou_aware_single_stage_object_detector_b200 import lib as L

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
# fp16 main pass + e4m3 corrections: ~2^-16 relative per product (tools/numerics_sim.py), fp32 accumulate
TOL = 2e-4


def rel_err(a, b):
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-6)).item()


def test_pack_encode_is_bit_exact_and_unpack_decodes():
    g = torch.Generator().manual_seed(0)
    for shape in ((2, 64, 7, 11), (1, 256, 13, 21), (3, 8, 5, 4)):
        n, c, h, w = shape
        x = torch.randn(*shape, generator=g) * torch.exp(2 * torch.randn(*shape, generator=g))
        x[0, 0, 0, 0], x[0, 1, 0, 0], x[0, 2, 0, 0] = 1e5, -3e-7, 500.0       # fp16 / e4m3 saturation, underflow
        eng = E.Engine(DEV, passes=2)
        m = eng.pack_input(x.to(DEV))
        y = eng.unpack_output(m)
        eng.run()
        torch.cuda.synchronize()
        t = m.tensor[: n * (h + 2) * (w + 2)].view(torch.uint8).view(n, h + 2, w + 2, 4 * c).cpu()
        want = SF.encode_rows(x.permute(0, 2, 3, 1).reshape(-1, c)).view(n, h, w, 4 * c)
        assert torch.equal(t[:, 1:-1, 1:-1], want)
        assert int(t[:, 0].max()) == 0 and int(t[:, -1].max()) == 0 and int(t[:, :, 0].max()) == 0 and int(t[:, :, -1].max()) == 0
        dec = SF.decode_rows(want.view(-1, 4 * c)).view(n, h, w, c).permute(0, 3, 1, 2)
        assert torch.equal(y.cpu(), dec)
        xs = x.clamp(-65504, 65504)
        # hi (11 bits) + l8 (4 bits): ~2^-16 relative inside the format's window; below |x| ~ 2^-5 the e4m3 residual is
        # subnormal (abs 2^-21), above |x| ~ 448 it saturates and the value falls back to fp16 precision (2^-12)
        big = xs.abs() > 448
        assert torch.allclose(y.cpu()[~big], xs[~big], rtol=2.0 ** -15, atol=5e-7)
        assert torch.allclose(y.cpu()[big], xs[big], rtol=2.0 ** -11)


def run_conv(x, w, bias=None, stride=1, relu=False, residual=None, two_cta=None):
    eng = E.Engine(DEV, passes=2)
    m = eng.pack_input(x.to(DEV).contiguous())
    k, co = w.shape[-1], w.shape[0]
    wp = E.pack_weight(w, E.pick_block_n(co)[1])
    res = eng.pack_input(residual.to(DEV).contiguous()) if residual is not None else None
    kw = dict(shift=bias, relu=relu, two_cta=two_cta)
    if stride == 1:
        out = eng.conv("t", [m], E.TAPS_1X1 if k == 1 else E.TAPS_3X3, wp, x.shape[1], co, residual=res,
                       res_mode=L.RES_SAME if res is not None else L.RES_NONE, **kw)
    elif k == 1:
        out = eng.conv("t", [eng.phase_split("p", m, mask=8)[3]] * 4, E.TAPS_1X1_S2, wp, x.shape[1], co, **kw)
    else:
        out = eng.conv("t", eng.phase_split("p", m), E.TAPS_3X3_S2, wp, x.shape[1], co, **kw)
    y = eng.unpack_output(out)
    eng.run()
    torch.cuda.synchronize()
    return y.cpu()


@pytest.mark.parametrize("cin,cout,k,hw", [(64, 64, 1, (9, 13)), (64, 256, 1, (17, 23)), (256, 64, 1, (8, 8)),
                                           (64, 64, 3, (12, 20)), (128, 128, 3, (25, 42)),
                                           (256, 256, 3, (13, 21)), (2048, 256, 1, (5, 6))])
def test_conv_stride1_vs_torch(cin, cout, k, hw):
    g = torch.Generator().manual_seed(cin + cout + k)
    x = torch.randn(2, cin, *hw, generator=g)
    w = torch.randn(cout, cin, k, k, generator=g) * (2.0 / (cin * k * k)) ** 0.5
    w *= torch.exp(torch.randn(cout, 1, 1, 1, generator=g))          # per-channel scales differ (BN fold)
    b = torch.randn(cout, generator=g)
    ref = F.conv2d(x, w, b, padding=k // 2)
    y = run_conv(x, w, bias=b)
    assert y.shape == ref.shape
    assert rel_err(y, ref) < TOL, rel_err(y, ref)


@pytest.mark.parametrize("cin,cout,k,hw", [(128, 128, 3, (25, 42)), (256, 512, 1, (20, 28)), (2048, 256, 3, (25, 42))])
def test_conv_stride2_vs_torch(cin, cout, k, hw):
    g = torch.Generator().manual_seed(cin * 3 + cout + k)
    x = torch.randn(2, cin, *hw, generator=g)
    w = torch.randn(cout, cin, k, k, generator=g) * (2.0 / (cin * k * k)) ** 0.5
    b = torch.randn(cout, generator=g)
    ref = F.conv2d(x, w, b, stride=2, padding=k // 2)
    y = run_conv(x, w, bias=b, stride=2)
    assert y.shape == ref.shape and rel_err(y, ref) < TOL, rel_err(y, ref)


@pytest.mark.parametrize("cin,cout,k,shape,res", [(64, 256, 1, (2, 50, 70), True),      # pair, N = 256: ONE accumulator stage
                                                  (64, 128, 1, (2, 50, 70), True),      # pair, residual ring, two stages
                                                  (256, 256, 3, (2, 50, 70), False),    # the head-tower shape
                                                  (64, 64, 3, (2, 50, 70), False)])
def test_cta_pair_mode_matches_torch(cin, cout, k, shape, res):
    n, h, w = shape
    g = torch.Generator().manual_seed(cin + cout + h)
    x = torch.randn(n, cin, h, w, generator=g)
    wt = torch.randn(cout, cin, k, k, generator=g) * (2.0 / (cin * k * k)) ** 0.5
    b = torch.randn(cout, generator=g)
    r = torch.randn(n, cout, h, w, generator=g) if res else None
    ref = F.conv2d(x, wt, b, padding=k // 2)
    if res:
        ref = F.relu(ref + r)
    y = run_conv(x, wt, bias=b, relu=res, residual=r, two_cta=True)
    assert rel_err(y, ref) < TOL, rel_err(y, ref)


@pytest.mark.parametrize("pair", [False, True])
def test_multi_segment_head_style_dense_outputs(pair):
    """Shared weights over several levels in ONE launch, dense fp32 NHWC outputs (N = 240 tiles and a 48-wide split)."""
    g = torch.Generator().manual_seed(21)
    sizes = [(2, 30, 44), (2, 15, 22), (2, 8, 11), (2, 2, 3), (2, 1, 2)]
    xs = [torch.randn(n, 256, h, w, generator=g) for (n, h, w) in sizes]
    w_cls, b_cls = torch.randn(720, 256, 3, 3, generator=g) * 0.02, torch.randn(720, generator=g)
    w_ri, b_ri = torch.randn(45, 256, 3, 3, generator=g) * 0.02, torch.randn(45, generator=g)
    eng = E.Engine(DEV, passes=2)
    Fm = eng.new_map(sizes, 256)
    for s, x in enumerate(xs):
        n, c, h, w = x.shape
        xd = x.to(DEV)
        eng.keep.append(xd)
        L.check(eng.lib.iou_pack_nchw_fmt(xd.data_ptr(), n, c, h, w, Fm.ptr, Fm.segs[s][0], 1, L.stream_ptr()))
    cls_out = [torch.zeros(n, h, w, 720, device=DEV) for (n, h, w) in sizes]
    reg_out = [torch.zeros(n, h, w, 36, device=DEV) for (n, h, w) in sizes]
    iou_out = [torch.zeros(n, h, w, 9, device=DEV) for (n, h, w) in sizes]
    # iou_conv_desc.group_max_cols = 80: per pixel and anchor, two partial maxima of the 80 class logits
    gmax = [torch.full((n, h, w, 9, 2), float("nan"), device=DEV) for (n, h, w) in sizes]
    eng.conv("cls", [Fm], E.TAPS_3X3, E.pack_weight(w_cls, 720), 256, 720, shift=b_cls, dense_out=cls_out,
             two_cta=pair, group_max_out=gmax, group_max_cols=80)
    eng.conv("ri", [Fm], E.TAPS_3X3, E.pack_weight(w_ri, 48), 256, 45, shift=b_ri, dense_out=reg_out,
             dense_out2=iou_out, dense_split=36, two_cta=False)
    eng.run()
    torch.cuda.synchronize()
    for s, x in enumerate(xs):
        n, _, h, w = x.shape
        # exactly the maximum of the values the same launch stored (no NaN left: every slot was written)
        assert torch.equal(gmax[s].max(dim=-1).values, cls_out[s].view(n, h, w, 9, 80).max(dim=-1).values)
        ref = F.conv2d(x, w_cls, b_cls, padding=1).permute(0, 2, 3, 1)
        assert rel_err(cls_out[s].cpu(), ref) < TOL
        ref = F.conv2d(x, w_ri, b_ri, padding=1).permute(0, 2, 3, 1)
        assert rel_err(reg_out[s].cpu(), ref[..., :36]) < TOL and rel_err(iou_out[s].cpu(), ref[..., 36:]) < TOL


@pytest.mark.parametrize("pair", [False, True])
@pytest.mark.parametrize("width,groups,stride,hw", [(256, 64, 1, (9, 13)), (256, 32, 2, (20, 28))])
def test_grouped_conv_block_diagonal_vs_torch(width, groups, stride, hw, pair):
    """ResNeXt 3x3 grouped conv (resnext.py:47-56) as a block-diagonal tap-GEMM, passes = 2."""
    g = torch.Generator().manual_seed(width + groups + stride)
    cg = width // groups
    x = torch.randn(2, width, *hw, generator=g)
    w = torch.randn(width, cg, 3, 3, generator=g) * (2.0 / (cg * 9)) ** 0.5
    b = torch.randn(width, generator=g)
    ref = F.conv2d(x, w, b, stride=stride, padding=1, groups=groups)
    eng = E.Engine(DEV, passes=2)
    m = eng.pack_input(x.to(DEV).contiguous())
    wp = E.pack_weight_grouped(w, groups)
    if stride == 1:
        out = eng.conv("g", [m], E.TAPS_3X3, wp, width, width, shift=b, diag_k=True, two_cta=pair)
    else:
        out = eng.conv("g", eng.phase_split("p", m), E.TAPS_3X3_S2, wp, width, width, shift=b, diag_k=True,
                       two_cta=pair)
    y = eng.unpack_output(out)
    eng.run()
    torch.cuda.synchronize()
    assert y.shape == ref.shape and rel_err(y.cpu(), ref) < TOL, rel_err(y.cpu(), ref)


def test_fpn_upsample_residual_and_stem():
    g = torch.Generator().manual_seed(5)
    # lateral 1x1 + nearest-2x top-down add (fpn.py:108-110)
    xf, xc = torch.randn(2, 512, 12, 20, generator=g), torch.randn(2, 256, 6, 10, generator=g)
    w, b = torch.randn(256, 512, 1, 1, generator=g) * 0.05, torch.randn(256, generator=g)
    eng = E.Engine(DEV, passes=2)
    mf, mc = eng.pack_input(xf.to(DEV)), eng.pack_input(xc.to(DEV))
    out = eng.conv("lat", [mf], E.TAPS_1X1, E.pack_weight(w, 256), 512, 256, shift=b, residual=mc,
                   res_mode=L.RES_UPSAMPLE2)
    y = eng.unpack_output(out)
    # stem: 7x7/s2 conv + BN + ReLU + 3x3/s2 max pool (resnet.py:508-511)
    img = torch.randn(2, 3, 70, 90, generator=g)
    sd = {"backbone.conv1.weight": torch.randn(64, 3, 7, 7, generator=g) * 0.1,
          "backbone.bn1.weight": torch.rand(64, generator=g) + 0.5, "backbone.bn1.bias": torch.randn(64, generator=g) * 0.1,
          "backbone.bn1.running_mean": torch.randn(64, generator=g) * 0.1,
          "backbone.bn1.running_var": torch.rand(64, generator=g) + 0.5}
    ys = eng.unpack_output(eng.add_stem(sd, img.to(DEV)))
    eng.run()
    torch.cuda.synchronize()
    ref = F.conv2d(xf, w, b) + F.interpolate(xc, scale_factor=2, mode="nearest")
    assert rel_err(y.cpu(), ref) < TOL, rel_err(y.cpu(), ref)
    t = F.conv2d(img, sd["backbone.conv1.weight"], None, stride=2, padding=3)
    t = F.batch_norm(t, sd["backbone.bn1.running_mean"], sd["backbone.bn1.running_var"], sd["backbone.bn1.weight"],
                     sd["backbone.bn1.bias"], False, 0.0, 1e-5)
    ref = F.max_pool2d(F.relu(t), 3, 2, 1)
    assert ys.shape == ref.shape and rel_err(ys.cpu(), ref) < TOL, rel_err(ys.cpu(), ref)


def test_whole_detector_small_passes2():
    """Backbone + FPN + head + get_bboxes with passes = 2 against the CPU oracle: head maps within 2e-4 of range,
    detections within 1e-4 (scores) / 1e-4 * max(H, W) px (boxes) -- the same bar as the default path."""
    worst = U.check_detector_small(128, 160, 2, verbose=False, passes=2)
    assert worst <= 2e-4


@pytest.mark.parametrize("relu", [True, False])
def test_group_norm_and_relu_phase_split(relu):
    """The FCOS pieces in the fp16 + e4m3 format: in-place GroupNorm(+ReLU) over two segments against torch, and the
    phase split with the fused ReLU against relu() of the plain phase split."""
    g = torch.Generator().manual_seed(9)
    sizes = [(2, 9, 14), (2, 5, 7)]
    xs = [torch.randn(n, 256, h, w, generator=g) * 3 + 0.5 for (n, h, w) in sizes]
    gamma, beta = torch.rand(256, generator=g) + 0.5, torch.randn(256, generator=g) * 0.2
    eng = E.Engine(DEV, passes=2)
    Fm = eng.new_map(sizes, 256)
    for s_, x in enumerate(xs):
        n, c, h, w = x.shape
        xd = x.to(DEV)
        eng.keep.append(xd)
        L.check(eng.lib.iou_pack_nchw_fmt(xd.data_ptr(), n, c, h, w, Fm.ptr, Fm.segs[s_][0], 1, L.stream_ptr()))
    eng.group_norm("gn", Fm, gamma, beta, 32, relu=relu)
    outs = [eng.unpack_output(Fm, s_) for s_ in range(2)]
    m = eng.pack_input((xs[0] - 0.5).to(DEV))                 # mixed signs
    ph_relu = eng.phase_split("pr", m, relu=True)
    ph_plain = eng.phase_split("pp", m)
    pr = [eng.unpack_output(p) for p in ph_relu]
    pp = [eng.unpack_output(p) for p in ph_plain]
    eng.run()
    torch.cuda.synchronize()
    for x, y in zip(xs, outs):
        ref = F.group_norm(x, 32, gamma, beta, 1e-5)
        ref = F.relu(ref) if relu else ref
        assert rel_err(y.cpu(), ref) < TOL, rel_err(y.cpu(), ref)
    for a, b, ra, rb in zip(pr, pp, ph_relu, ph_plain):
        assert torch.equal(a.cpu(), F.relu(b.cpu()))
        # the clamped entries are all-zero bytes (what a conv epilogue would have written for relu(x) = 0)
        neg = (b.cpu() < 0)
        assert bool(neg.any())


# ------------------------------------------------------------------------------------------------
# Dynamic range of the fp16 + e4m3 format outside synthetic weights (trained-checkpoint-like activation ranges)
def _conv_chain(scale, n_layers=2, cin=64, hw=(20, 28)):
    """x -> conv3x3 -> relu -> conv3x3, activations scaled by `scale`; returns (gpu maps, torch-CPU fp32 maps, engine)."""
    g = torch.Generator().manual_seed(5)
    x = (torch.randn(2, cin, *hw, generator=g) * scale).clamp(-65504.0, 65504.0)     # the input pack clamps too
    ws = [torch.randn(cin, cin, 3, 3, generator=g) * (2.0 / (cin * 9)) ** 0.5 for _ in range(n_layers)]
    eng = E.Engine(DEV, passes=2)
    m = eng.pack_input(x.to(DEV))
    outs, refs, r = [], [], x
    for i, w in enumerate(ws):
        m = eng.conv("c%d" % i, [m], E.TAPS_3X3, E.pack_weight(w, cin), cin, cin, relu=(i + 1 < n_layers))
        outs.append(eng.unpack_output(m))
        r = F.conv2d(r, w, padding=1)
        if i + 1 < n_layers:
            r = F.relu(r)
        refs.append(r)
    eng.run()
    torch.cuda.synchronize()
    return [o.cpu() for o in outs], refs, eng


def test_range_above_448_degrades_to_fp16_precision_and_stays_finite():
    """|v| > 448: the e4m3 parts saturate (split_fmt.cuh) and the element keeps fp16 precision -- never worse."""
    outs, refs, eng = _conv_chain(scale=300.0)
    assert refs[0].abs().max() > 448 and refs[0].abs().max() < 65504
    for o, r in zip(outs, refs):
        assert torch.isfinite(o).all()
        assert rel_err(o, r) < 2.0 ** -9, rel_err(o, r)          # fp16-grade (2^-11 per element), summed over K
    rep = {r["label"]: r for r in eng.range_report()}
    assert rep["c0"]["above_448"] > 0 and rep["c0"]["saturated"] == 0
    assert abs(rep["c0"]["max_abs"] - refs[0].clamp_min(0).max().item()) < 1e-2 * refs[0].max().item()


def test_range_beyond_fp16_saturates_finite_and_is_counted():
    """|v| > 65504: the encode clamps at the fp16 limit (no inf, so no NaN downstream) and iou_range_stats counts it."""
    outs, refs, eng = _conv_chain(scale=3.0e4)
    assert refs[0].abs().max() > 65504
    for o in outs:
        assert torch.isfinite(o).all()
    assert outs[0].max().item() <= 65505.0
    rep = {r["label"]: r for r in eng.range_report()}
    lo, hi = int((refs[0] >= 65504 * 1.001).sum()), int((refs[0] >= 65504 * 0.999).sum())
    assert lo > 0 and lo <= rep["c0"]["saturated"] <= hi, (lo, rep["c0"]["saturated"], hi)


def test_range_tiny_activations_keep_absolute_precision():
    """|v| < 2^-5: the e4m3 residual goes subnormal, absolute error 2^-21 per stored element (which a conv then sums
    over its K = 576 inputs: a few 1e-6 on outputs of ~1e-2)."""
    outs, refs, _ = _conv_chain(scale=2.0e-3)
    assert refs[-1].abs().max() < 2.0 ** -5
    for o, r in zip(outs, refs):
        assert (o - r).abs().max().item() < 5e-6, (o - r).abs().max().item()


def test_detector_unnormalised_input_range_guard():
    """Un-normalised input (pixel scale, as a caffe-style checkpoint sees it) pushes activations past 448; a wild input
    scale pushes them past the fp16 limit: the first batch of the plan raises instead of clamping silently, and
    detector.passes = 3 (fp32 exponent range) still matches the oracle."""
    from oracle import model as om
    det, cfg = U.small_detector()
    sd = {k: v.clone() for k, v in det.state_dict().items()}
    det = det.to(DEV)
    g = torch.Generator().manual_seed(3)
    base = torch.randn(1, 3, 128, 160, generator=g)
    metas = [dict(ori_shape=(128, 157, 3), img_shape=(128, 157, 3), pad_shape=(128, 160, 3), scale_factor=1.0, flip=False)]
    # (a) input at 300x the normalised scale: some maps exceed 448, none the fp16 limit -> runs, the report says so,
    #     logits stay close
    img = base * 300.0
    det.use_cuda_graph = False
    det.detect_device(img.to(DEV), metas, rescale=False)
    plan = det.fused_plan(img.shape, DEV, False)
    assert plan.range is not None and sum(r["above_448"] for r in plan.range) > 0
    assert sum(r["saturated"] for r in plan.range) == 0
    ref_cls, ref_reg, ref_iou = om.detector_forward(sd, img)
    for mine, ref in ((plan.outs[0], ref_cls), (plan.outs[1], ref_reg), (plan.outs[2], ref_iou)):
        for a, b in zip(mine, ref):
            assert torch.isfinite(a).all()
            assert rel_err(a.cpu(), b) < 2e-3, rel_err(a.cpu(), b)          # fp16-grade on the saturated elements
    # (b) a wild scale: beyond the fp16 limit -> loud
    wild = base * 3.0e5
    det.detect_device(wild.to(DEV), metas, rescale=False)      # same plan: only a plan's FIRST batch is checked ...
    with pytest.raises(RuntimeError, match="65504"):
        plan.check_range()                                       # ... but the check can be asked for at any time
    det._fused.clear()
    with pytest.raises(RuntimeError, match="65504"):            # a fresh plan meets the wild input first: loud
        det.detect_device(wild.to(DEV), metas, rescale=False)
    # ... unless the caller accepts the clamping: finite, no NaN
    det.range_check = False
    det._fused.clear()
    d, l, c = det.detect_device(wild.to(DEV), metas, rescale=False)
    assert torch.isfinite(d).all()
    # (c) passes = 3 holds the fp32 exponent range
    det.passes = 3
    det.detect_device(wild.to(DEV), metas, rescale=False)
    plan3 = det.fused_plan(wild.shape, DEV, False)
    ref_cls, _, _ = om.detector_forward(sd, wild)
    for a, b in zip(plan3.outs[0], ref_cls):
        assert rel_err(a.cpu(), b) < 5e-4, rel_err(a.cpu(), b)


@pytest.mark.parametrize("passes", [2, 3])
@pytest.mark.parametrize("ks,cin,cout,stride", [(4, 512, 64, 1), (8, 1024, 256, 2), (2, 256, 128, 1)])
def test_split_k_conv_with_channel_group_sum_vs_torch(passes, ks, cin, cout, stride):
    """k_split: the K loop of a conv with few output rows is cut into `ks` channel slices whose partial sums land in
    N-concatenated output channels (iou_conv_desc.k_split) and are added, with the bias, by iou_sum_channel_groups
    (how FPN's P6 conv on C5 runs, fpn.py:126-128)."""
    g = torch.Generator().manual_seed(ks * 7 + cin)
    x = torch.randn(2, cin, 13, 21, generator=g)
    w = torch.randn(cout, cin, 3, 3, generator=g) * (2.0 / (cin * 9)) ** 0.5
    b = torch.randn(cout, generator=g)
    ref = F.conv2d(x, w, b, stride=stride, padding=1)
    eng = E.Engine(DEV, passes=passes)
    m = eng.pack_input(x.to(DEV))
    cs = cin // ks
    wsp = torch.cat([w[:, j * cs:(j + 1) * cs] for j in range(ks)], dim=0)
    n, _, ho, wo = ref.shape
    part = eng.new_map([(n, ho, wo)], ks * cout)
    srcs, taps = ([m], E.TAPS_3X3) if stride == 1 else (eng.phase_split("p", m), E.TAPS_3X3_S2)
    eng.conv("c", srcs, taps, E.pack_weight(wsp, ks * cout), cs, ks * cout, out=part, k_split=ks)
    out = eng.new_map([(n, ho, wo)], cout)
    eng.sum_groups("s", part, out, ks, bias=b)
    y = eng.unpack_output(out)
    eng.run()
    torch.cuda.synchronize()
    assert rel_err(y.cpu(), ref) < TOL, rel_err(y.cpu(), ref)
    # border rows of the summed map stay zero (the next conv's taps read them as padding)
    rows = out.tensor[: n * (ho + 2) * (wo + 2)].view(torch.uint8).view(n, ho + 2, wo + 2, 4 * cout)
    assert int(rows[:, 0].max()) == 0 and int(rows[:, -1].max()) == 0 and int(rows[:, :, 0].max()) == 0


@pytest.mark.parametrize("ca,cb,cc,shape", [(64, 256, 64, (2, 40, 60)),        # layer-1 shape: one N tile, resident weights
                                            (128, 512, 128, (2, 40, 60)),      # two N tiles of conv A, streamed weights
                                            (256, 1024, 256, (1, 50, 84)),     # layer-3 shape: four N tiles, K = 1024 for conv B
                                            (64, 256, 64, (1, 45, 67)),        # odd tile count: the last pair's second CTA idles
                                            (64, 256, 128, (8, 30, 30)),       # several images
                                            (64, 256, 64, (3, 200, 304))])     # ~9 units per CTA pair: the ring wraps, stages change hands
def test_chained_1x1_convs_one_launch_equal_two_launches(ca, cb, cc, shape):
    """conv3(+residual, ReLU) of a bottleneck and conv1 of the next one in ONE launch (iou_conv_chain_plan_create,
    csrc/conv_chain.cu): bit-identical to the two separate launches, and fp32-grade against torch."""
    n, h, w = shape
    g = torch.Generator().manual_seed(ca + cb + cc + h)
    x = torch.randn(n, ca, h, w, generator=g)
    r = torch.randn(n, cb, h, w, generator=g)
    wa = torch.randn(cb, ca, 1, 1, generator=g) * (2.0 / ca) ** 0.5
    wb = torch.randn(cc, cb, 1, 1, generator=g) * (2.0 / cb) ** 0.5
    ba, bb = torch.randn(cb, generator=g), torch.randn(cc, generator=g)
    ref_y = F.relu(F.conv2d(x, wa, ba) + r)
    ref_t = F.relu(F.conv2d(ref_y, wb, bb))
    outs = []
    for chained in (True, False):
        eng = E.Engine(DEV, passes=2)
        m, rm = eng.pack_input(x.to(DEV)), eng.pack_input(r.to(DEV))
        y = eng.conv("a", [m], E.TAPS_1X1, E.pack_weight(wa, cb), ca, cb, shift=ba, relu=True, residual=rm,
                     res_mode=L.RES_SAME, hold=chained, two_cta=True)
        t = eng.conv("b", [y], E.TAPS_1X1, E.pack_weight(wb, cc), cb, cc, shift=bb, relu=True, chain=chained,
                     force_bn=(cc, cc), two_cta=True)
        assert getattr(eng, "chained", 0) == (1 if chained else 0)
        assert len([o for o in eng.ops if not o[0].startswith("pack")]) == (1 if chained else 2)
        yo, to = eng.unpack_output(y), eng.unpack_output(t)
        eng.run()
        torch.cuda.synchronize()
        outs.append((yo.cpu(), to.cpu()))
    (y1, t1), (y2, t2) = outs
    assert torch.equal(y1, y2) and torch.equal(t1, t2)
    assert rel_err(y1, ref_y) < TOL and rel_err(t1, ref_t) < TOL, (rel_err(y1, ref_y), rel_err(t1, ref_t))


@pytest.mark.parametrize("cin,cout,k,shape,res,pair,phase", [
    (64, 256, 1, (3, 200, 304), True, True, False),     # layer-1 conv3: 8 column groups per tile over 3 warps, ~10 tiles per CTA pair
    (256, 1024, 1, (2, 50, 84), True, True, False),     # layer-3 conv3: four N tiles per row tile
    (128, 512, 1, (1, 45, 67), True, None, True),       # odd tile count, phase (1,1) written next to the output
    (256, 64, 1, (2, 100, 150), False, True, False),    # two column groups per tile: every third tile a warp has nothing to do
    (256, 128, 1, (2, 60, 90), False, None, True),      # four groups, all four phase maps
    (64, 64, 3, (2, 60, 90), False, True, False),       # 3x3, resident weights
    (64, 256, 1, (2, 37, 53), True, False, False)])     # single-CTA kernel (no pairs)
def test_wide_epilogue_is_bit_identical_to_the_8_warp_kernel(cin, cout, k, shape, res, pair, phase):
    """iou_conv_desc.wide = 1 (conv_tap_gemm_kernel<., ., 3>: 12 epilogue warps behind setmaxnreg, column groups handed out
    round-robin across tiles, residual slabs overwritten in place by the results) computes exactly what the 8-warp kernel does."""
    n, h, w = shape
    g = torch.Generator().manual_seed(cin + cout + h)
    x = torch.randn(n, cin, h, w, generator=g)
    wt = torch.randn(cout, cin, k, k, generator=g) * (2.0 / (cin * k * k)) ** 0.5
    b = torch.randn(cout, generator=g)
    r = torch.randn(n, cout, h, w, generator=g) if res else None
    ref = F.conv2d(x, wt, b, padding=k // 2)
    ref = F.relu(ref + r) if res else F.relu(ref)
    outs = []
    for wide in (1, -1):
        eng = E.Engine(DEV, passes=2)
        m = eng.pack_input(x.to(DEV))
        rm = eng.pack_input(r.to(DEV)) if res else None
        ph = eng.new_phase_maps(n, h, w, cout, mask=8 if res else 15) if phase else None
        taps = E.TAPS_1X1 if k == 1 else E.TAPS_3X3
        y = eng.conv("c", [m], taps, E.pack_weight(wt, cout), cin, cout, shift=b, relu=True, residual=rm,
                     res_mode=L.RES_SAME if res else L.RES_NONE, two_cta=pair, phase_outs=ph, wide=wide)
        assert eng.epi_warps["c"] == (12 if wide == 1 else 8)
        yo = eng.unpack_output(y)
        eng.run()
        torch.cuda.synchronize()
        outs.append((yo.cpu(), y.tensor.view(torch.int16).cpu(), [p.tensor.view(torch.int16).cpu() for p in (ph or []) if p is not None]))
    assert torch.equal(outs[0][1], outs[1][1])
    for a, c in zip(outs[0][2], outs[1][2]):
        assert torch.equal(a, c)
    assert rel_err(outs[0][0], ref) < TOL, rel_err(outs[0][0], ref)
