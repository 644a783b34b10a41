"""-m gpu numerics tests of the tcgen05 conv engine and the layout kernels against plain
PyTorch fp32 (CPU) references of the same ops."""
import pytest
import torch
import torch.nn.functional as F

import parity_util as U  # noqa: F401  (sets sys.path)
from iou_aware_single_stage_object_detector_b200 import engine as E
from iou_aware_single_stage_object_detector_b200 import lib as L

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
# 3-pass bf16 split accumulates in fp32: expected relative error ~1e-5 of the output scale
TOL = 2e-4


def rel_err(a, b):
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-6)).item()


def run_conv(x, w, bias=None, stride=1, relu=False, scale=None, shift=None, residual=None, passes=3):
    """x (N,C,H,W) cpu, w (Co,Ci,k,k) cpu -> engine output (N,Co,Ho,Wo) cpu."""
    eng = E.Engine(DEV, passes=passes)
    xin = x.to(DEV).contiguous()
    m = eng.pack_input(xin)
    k = w.shape[-1]
    co = w.shape[0]
    wp = E.pack_weight(w, E.pick_block_n(co)[1])
    res = eng.pack_input(residual.to(DEV).contiguous()) if residual is not None else None
    sh = shift if shift is not None else bias
    if stride == 1:
        taps = E.TAPS_1X1 if k == 1 else E.TAPS_3X3
        out = eng.conv("t", [m], taps, wp, x.shape[1], co, scale=scale, shift=sh, relu=relu, residual=res,
                       res_mode=L.RES_SAME if res is not None else L.RES_NONE)
    else:
        if k == 1:
            ph = eng.phase_split("p", m, mask=8)
            out = eng.conv("t", [ph[3]] * 4, E.TAPS_1X1_S2, wp, x.shape[1], co, scale=scale, shift=sh, relu=relu)
        else:
            ph = eng.phase_split("p", m)
            out = eng.conv("t", ph, E.TAPS_3X3_S2, wp, x.shape[1], co, scale=scale, shift=sh, relu=relu)
    y = eng.unpack_output(out)
    eng.run()
    torch.cuda.synchronize()
    return y.cpu()


def test_pack_unpack_roundtrip():
    g = torch.Generator().manual_seed(0)
    for shape in ((2, 64, 7, 11), (1, 256, 13, 21), (3, 8, 5, 4)):
        x = torch.randn(*shape, generator=g)
        eng = E.Engine(DEV)
        xin = x.to(DEV)
        m = eng.pack_input(xin)
        y = eng.unpack_output(m)
        eng.run()
        torch.cuda.synchronize()
        # hi + lo carries 16 mantissa bits
        assert torch.allclose(y.cpu(), x, rtol=2e-5, atol=1e-30), rel_err(y.cpu(), x)
        # border rows of the padded layout are zero
        n, c, h, w = shape
        t = m.tensor[: n * (h + 2) * (w + 2)].view(n, h + 2, w + 2, 2 * c).float()
        assert float(t[:, 0].abs().max()) == 0 and float(t[:, -1].abs().max()) == 0
        assert float(t[:, :, 0].abs().max()) == 0 and float(t[:, :, -1].abs().max()) == 0


@pytest.mark.parametrize("cin,cout,k,hw", [(64, 64, 1, (9, 13)), (64, 256, 1, (17, 23)), (256, 64, 1, (8, 8)),
                                           (64, 64, 3, (12, 20)), (128, 128, 3, (25, 42)),
                                           (256, 256, 3, (13, 21)), (512, 128, 1, (7, 11)),
                                           (256, 512, 1, (10, 10)), (2048, 256, 1, (5, 6))])
def test_conv_stride1_vs_torch(cin, cout, k, hw):
    g = torch.Generator().manual_seed(cin + cout + k)
    x = torch.randn(2, cin, *hw, generator=g)
    w = torch.randn(cout, cin, k, k, generator=g) * (2.0 / (cin * k * k)) ** 0.5
    b = torch.randn(cout, generator=g)
    ref = F.conv2d(x, w, b, padding=k // 2)
    y = run_conv(x, w, bias=b)
    assert y.shape == ref.shape
    assert rel_err(y, ref) < TOL, rel_err(y, ref)


def test_conv_bn_relu_residual_epilogue():
    g = torch.Generator().manual_seed(11)
    x = torch.randn(2, 64, 14, 18, generator=g)
    w = torch.randn(256, 64, 1, 1, generator=g) * 0.2
    scale = torch.rand(256, generator=g) + 0.5
    shift = torch.randn(256, generator=g) * 0.3
    res = torch.randn(2, 256, 14, 18, generator=g)
    ref = F.relu(F.conv2d(x, w) * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1) + res)
    y = run_conv(x, w, scale=scale, shift=shift, relu=True, residual=res)
    assert rel_err(y, ref) < TOL, rel_err(y, ref)


@pytest.mark.parametrize("cin,cout,k,hw", [(64, 64, 3, (12, 20)), (128, 128, 3, (25, 42)), (256, 256, 3, (13, 21)),
                                           (256, 512, 1, (20, 28)), (512, 1024, 1, (25, 42)),
                                           (2048, 256, 3, (25, 42)), (256, 256, 3, (7, 11))])
def test_conv_stride2_vs_torch(cin, cout, k, hw):
    g = torch.Generator().manual_seed(cin * 3 + cout + k)
    x = torch.randn(2, cin, *hw, generator=g)
    w = torch.randn(cout, cin, k, k, generator=g) * (2.0 / (cin * k * k)) ** 0.5
    b = torch.randn(cout, generator=g)
    ref = F.conv2d(x, w, b, stride=2, padding=k // 2)
    y = run_conv(x, w, bias=b, stride=2)
    assert y.shape == ref.shape, (y.shape, ref.shape)
    assert rel_err(y, ref) < TOL, rel_err(y, ref)


def test_conv_single_pass_bf16_is_coarser_but_close():
    g = torch.Generator().manual_seed(5)
    x = torch.randn(1, 128, 10, 10, generator=g)
    w = torch.randn(128, 128, 3, 3, generator=g) * 0.03
    ref = F.conv2d(x, w, padding=1)
    e1 = rel_err(run_conv(x, w, passes=1), ref)
    e3 = rel_err(run_conv(x, w, passes=3), ref)
    assert e3 < TOL and e1 < 3e-2 and e3 < e1 / 10, (e1, e3)


def test_multi_segment_head_style_dense_outputs():
    """Shared weights over several levels in ONE launch, dense fp32 NHWC outputs with a split."""
    g = torch.Generator().manual_seed(21)
    sizes = [(2, 12, 20), (2, 6, 10), (2, 3, 5), (2, 2, 3), (2, 1, 2)]
    xs = [torch.randn(n, 256, h, w, generator=g) for (n, h, w) in sizes]
    w_cls = torch.randn(720, 256, 3, 3, generator=g) * 0.02
    b_cls = torch.randn(720, generator=g)
    w_ri = torch.randn(45, 256, 3, 3, generator=g) * 0.02
    b_ri = torch.randn(45, generator=g)
    eng = E.Engine(DEV)
    Fm = E.FlatMap(sizes, 256, DEV)
    xin = [x.to(DEV) for x in xs]
    for s, x in enumerate(xin):
        n, c, h, w = x.shape
        L.check(eng.lib.iou_pack_nchw(x.data_ptr(), n, c, h, w, Fm.ptr, Fm.segs[s][0], L.stream_ptr()))
    cls_out = [torch.zeros(n, h, w, 720, device=DEV) for (n, h, w) in sizes]
    reg_out = [torch.zeros(n, h, w, 36, device=DEV) for (n, h, w) in sizes]
    iou_out = [torch.zeros(n, h, w, 9, device=DEV) for (n, h, w) in sizes]
    eng.conv("cls", [Fm], E.TAPS_3X3, E.pack_weight(w_cls, 720), 256, 720, shift=b_cls, dense_out=cls_out)
    eng.conv("ri", [Fm], E.TAPS_3X3, E.pack_weight(w_ri, 48), 256, 45, shift=b_ri, dense_out=reg_out,
             dense_out2=iou_out, dense_split=36)
    eng.run()
    torch.cuda.synchronize()
    for s, x in enumerate(xs):
        ref_c = F.conv2d(x, w_cls, b_cls, padding=1).permute(0, 2, 3, 1)
        ref_ri = F.conv2d(x, w_ri, b_ri, padding=1).permute(0, 2, 3, 1)
        assert rel_err(cls_out[s].cpu(), ref_c) < TOL
        assert rel_err(reg_out[s].cpu(), ref_ri[..., :36]) < TOL
        assert rel_err(iou_out[s].cpu(), ref_ri[..., 36:]) < TOL


def test_lateral_upsample_add_epilogue():
    g = torch.Generator().manual_seed(31)
    fine = torch.randn(2, 512, 12, 20, generator=g)
    coarse = torch.randn(2, 256, 6, 10, generator=g)
    w = torch.randn(256, 512, 1, 1, generator=g) * 0.05
    b = torch.randn(256, generator=g)
    ref = F.conv2d(fine, w, b) + F.interpolate(coarse, scale_factor=2, mode="nearest")
    eng = E.Engine(DEV)
    mf = eng.pack_input(fine.to(DEV))
    mc = eng.pack_input(coarse.to(DEV))
    out = eng.conv("lat", [mf], E.TAPS_1X1, E.pack_weight(w, 256), 512, 256, shift=b, residual=mc,
                   res_mode=L.RES_UPSAMPLE2)
    y = eng.unpack_output(out)
    eng.run()
    torch.cuda.synchronize()
    assert rel_err(y.cpu(), ref) < TOL


def test_stem_im2col_conv_maxpool_vs_torch():
    g = torch.Generator().manual_seed(41)
    img = torch.randn(2, 3, 64, 96, generator=g)
    sd = {"backbone.conv1.weight": torch.randn(64, 3, 7, 7, generator=g) * 0.1,
          "backbone.bn1.weight": torch.rand(64, generator=g) + 0.5,
          "backbone.bn1.bias": torch.randn(64, generator=g) * 0.1,
          "backbone.bn1.running_mean": torch.randn(64, generator=g) * 0.1,
          "backbone.bn1.running_var": torch.rand(64, generator=g) + 0.5}
    x = F.conv2d(img, sd["backbone.conv1.weight"], stride=2, padding=3)
    x = F.batch_norm(x, sd["backbone.bn1.running_mean"], sd["backbone.bn1.running_var"],
                     sd["backbone.bn1.weight"], sd["backbone.bn1.bias"], False, 0.0, 1e-5)
    ref = F.max_pool2d(F.relu(x), 3, 2, 1)
    eng = E.Engine(DEV)
    out = eng.add_stem(sd, img.to(DEV))
    y = eng.unpack_output(out)
    eng.run()
    torch.cuda.synchronize()
    assert y.shape == ref.shape
    assert rel_err(y.cpu(), ref) < TOL, rel_err(y.cpu(), ref)


def test_conv_rejects_bad_descriptors_loudly():
    eng = E.Engine(DEV)
    m = E.FlatMap([(1, 4, 4)], 48, DEV)       # cin not a multiple of 64
    with pytest.raises(RuntimeError):
        eng.conv("bad", [m], E.TAPS_1X1, torch.zeros(64, 96, dtype=torch.bfloat16), 48, 64)


@pytest.mark.parametrize("pair", [None, True])
@pytest.mark.parametrize("width,groups,stride,hw", [(128, 32, 1, (14, 18)), (256, 64, 1, (9, 13)),
                                                    (256, 32, 2, (20, 28)), (512, 64, 2, (13, 21)),
                                                    (1024, 32, 1, (7, 11))])
def test_grouped_conv_block_diagonal_vs_torch(width, groups, stride, hw, pair):
    """ResNeXt 3x3 grouped conv (resnext.py:47-56) as a block-diagonal tap-GEMM."""
    g = torch.Generator().manual_seed(width + groups + stride)
    cg = width // groups
    x = torch.randn(2, width, *hw, generator=g)
    w = torch.randn(width, cg, 3, 3, generator=g) * (2.0 / (cg * 9)) ** 0.5
    b = torch.randn(width, generator=g)
    ref = F.conv2d(x, w, b, stride=stride, padding=1, groups=groups)
    eng = E.Engine(DEV)
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
    assert y.shape == ref.shape
    assert rel_err(y.cpu(), ref) < TOL, rel_err(y.cpu(), ref)


@pytest.mark.parametrize("cin,cout,k,shape,res", [(256, 256, 3, (2, 60, 80), False),     # even tile count
                                                  (256, 256, 3, (1, 41, 59), False),     # odd: last pair half idle
                                                  (128, 512, 1, (2, 33, 47), False),     # two N tiles
                                                  (64, 128, 1, (2, 50, 70), True),       # residual ring, N tile 128
                                                  (512, 256, 1, (3, 25, 42), False),
                                                  (64, 64, 3, (2, 50, 70), False),       # narrow N: 32 B rows per CTA
                                                  (128, 128, 3, (1, 37, 53), False),
                                                  (256, 64, 1, (2, 50, 70), False)])
def test_cta_pair_mode_matches_torch(cin, cout, k, shape, res):
    """tcgen05 cta_group::2 path (a cluster of two CTAs shares the B tile) forced on small maps."""
    n, h, w = shape
    g = torch.Generator().manual_seed(cin + cout + h)
    x = torch.randn(n, cin, h, w, generator=g)
    wt = torch.randn(cout, cin, k, k, generator=g) * (2.0 / (cin * k * k)) ** 0.5
    b = torch.randn(cout, generator=g)
    r = torch.randn(n, cout, h, w, generator=g) if res else None
    ref = F.conv2d(x, wt, b, padding=k // 2)
    if res:
        ref = F.relu(ref + r)
    eng = E.Engine(DEV)
    m = eng.pack_input(x.to(DEV).contiguous())
    rm = eng.pack_input(r.to(DEV).contiguous()) if res else None
    out = eng.conv("pair", [m], E.TAPS_1X1 if k == 1 else E.TAPS_3X3, E.pack_weight(wt, cout), cin, cout, shift=b,
                   relu=res, residual=rm, res_mode=L.RES_SAME if res else L.RES_NONE, two_cta=True)
    y = eng.unpack_output(out)
    eng.run()
    torch.cuda.synchronize()
    assert rel_err(y.cpu(), ref) < TOL, rel_err(y.cpu(), ref)


def test_cta_pair_mode_dense_head_outputs():
    g = torch.Generator().manual_seed(77)
    sizes = [(2, 30, 44), (2, 15, 22), (2, 8, 11)]
    xs = [torch.randn(n, 256, h, w, generator=g) for (n, h, w) in sizes]
    w_cls = torch.randn(720, 256, 3, 3, generator=g) * 0.02
    b_cls = torch.randn(720, generator=g)
    eng = E.Engine(DEV)
    Fm = eng.new_map(sizes, 256)
    for s, x in enumerate(xs):
        n, c, h, w = x.shape
        xd = x.to(DEV)
        eng.keep.append(xd)
        L.check(eng.lib.iou_pack_nchw(xd.data_ptr(), n, c, h, w, Fm.ptr, Fm.segs[s][0], L.stream_ptr()))
    cls_out = [torch.zeros(n, h, w, 720, device=DEV) for (n, h, w) in sizes]
    eng.conv("cls", [Fm], E.TAPS_3X3, E.pack_weight(w_cls, 720), 256, 720, shift=b_cls, dense_out=cls_out,
             two_cta=True)
    eng.run()
    torch.cuda.synchronize()
    for s, x in enumerate(xs):
        ref = F.conv2d(x, w_cls, b_cls, padding=1).permute(0, 2, 3, 1)
        assert rel_err(cls_out[s].cpu(), ref) < TOL


@pytest.mark.parametrize("passes", [3, 2])
@pytest.mark.parametrize("cin,cout,hw,res,pair", [(256, 128, (25, 42), False, None), (64, 256, (12, 20), True, None),
                                                  (128, 512, (13, 21), True, True), (256, 64, (50, 70), False, True)])
def test_fused_phase_outputs_equal_the_phase_split_kernel(cin, cout, hw, res, pair, passes):
    """conv(phase_outs=...) writes, from its epilogue, exactly the bytes iou_phase_split produces from the conv's
    ordinary output (odd and even sizes, with / without the ordinary output, residual ring, CTA pairs)."""
    g = torch.Generator().manual_seed(cin + cout + hw[0])
    n = 2
    x = torch.randn(n, cin, *hw, generator=g)
    w = torch.randn(cout, cin, 1, 1, generator=g) * (2.0 / cin) ** 0.5
    b = torch.randn(cout, generator=g)
    r = torch.randn(n, cout, *hw, generator=g) if res else None
    eng = E.Engine(DEV, passes=passes)
    m = eng.pack_input(x.to(DEV).contiguous())
    rm = eng.pack_input(r.to(DEV).contiguous()) if res else None
    kw = dict(shift=b, relu=True, residual=rm, res_mode=L.RES_SAME if res else L.RES_NONE, two_cta=pair)
    plain = eng.conv("plain", [m], E.TAPS_1X1, E.pack_weight(w, cout), cin, cout, **kw)
    want = eng.phase_split("split", plain)
    both = eng.new_phase_maps(n, hw[0], hw[1], cout, mask=8 if res else 15)
    out2 = eng.conv("both", [m], E.TAPS_1X1, E.pack_weight(w, cout), cin, cout, phase_outs=both, **kw)
    only = eng.new_phase_maps(n, hw[0], hw[1], cout)
    none = eng.conv("only", [m], E.TAPS_1X1, E.pack_weight(w, cout), cin, cout, phase_outs=only, phase_only=True, **kw)
    eng.run()
    torch.cuda.synchronize()
    assert none is None
    assert torch.equal(out2.tensor.view(torch.int16), plain.tensor.view(torch.int16))
    for i in range(4):
        if both[i] is not None:
            assert torch.equal(both[i].tensor.view(torch.int16), want[i].tensor.view(torch.int16)), i
        assert torch.equal(only[i].tensor.view(torch.int16), want[i].tensor.view(torch.int16)), i


@pytest.mark.parametrize("passes", [3, 2])
@pytest.mark.parametrize("width,cin,cout,stride,hw", [(64, 64, 256, 1, (12, 20)), (128, 256, 512, 2, (20, 28)),
                                                      (256, 512, 1024, 2, (13, 21))])
def test_k_concatenated_conv3_plus_downsample(width, cin, cout, stride, hw, passes):
    """relu(conv3(t2) + downsample(x)) of a bottleneck's first block as ONE GEMM over two sources with different channel
    counts (iou_conv_desc.src_cin), against torch fp32."""
    g = torch.Generator().manual_seed(width + cin + stride)
    x = torch.randn(2, cin, *hw, generator=g)
    ho, wo = (hw[0] + stride - 1) // stride, (hw[1] + stride - 1) // stride
    t2 = torch.randn(2, width, ho, wo, generator=g)
    w3 = torch.randn(cout, width, 1, 1, generator=g) * (1.0 / width) ** 0.5
    wd = torch.randn(cout, cin, 1, 1, generator=g) * (1.0 / cin) ** 0.5
    b = torch.randn(cout, generator=g)
    ref = F.relu(F.conv2d(t2, w3) + F.conv2d(x, wd, stride=stride) + b.view(1, -1, 1, 1))
    eng = E.Engine(DEV, passes=passes)
    mt, mx = eng.pack_input(t2.to(DEV).contiguous()), eng.pack_input(x.to(DEV).contiguous())
    src_b = eng.phase_split("p", mx, mask=8)[3] if stride == 2 else mx
    kmax = max(width, cin)
    wt = torch.zeros(2, cout, kmax)
    wt[0, :, :width], wt[1, :, :cin] = w3[:, :, 0, 0], wd[:, :, 0, 0]
    out = eng.conv("c3+ds", [mt, src_b], [(0, 0, 0), (1, 0, 0)], E._hi_lo_rows(wt), kmax, cout, shift=b, relu=True)
    y = eng.unpack_output(out)
    eng.run()
    torch.cuda.synchronize()
    assert y.shape == ref.shape and rel_err(y.cpu(), ref) < TOL, rel_err(y.cpu(), ref)
