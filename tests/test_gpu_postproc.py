"""-m gpu parity tests of the post-processing kernels (through the C ABI) against the golden
fixtures generated from the real reference and against the CPU oracle."""
import os

import numpy as np
import pytest
import torch

import cases
import parity_util as U
from oracle import postproc as op

pytestmark = pytest.mark.gpu
P = U.P
PP = U.PP


@pytest.mark.parametrize("name", cases.POSTPROC_CASES)
def test_get_bboxes_vs_reference_golden(name):
    assert U.check_postproc_case(name)


@pytest.mark.parametrize("name", list(cases.nms_inputs().keys()))
def test_nms_indices_bit_exact(name):
    """iou_nms == mmdet.ops.nms (reference nms_cpu golden; no IoU==thr ties in these inputs)."""
    gold = np.load(os.path.join(U.GOLD, "nms_keep.npz"))[name]
    dets = torch.from_numpy(cases.nms_inputs()[name]).cuda()
    kept, inds = P.nms(dets, 0.5)
    assert inds.dtype == torch.int64 and inds.is_cuda
    assert np.array_equal(inds.cpu().numpy(), gold)
    assert torch.equal(kept, dets[inds])
    # numpy in / numpy out with device_id (nms_wrapper.py:27-31,47-48)
    k2, i2 = P.nms(cases.nms_inputs()[name], 0.5, device_id=0)
    assert isinstance(i2, np.ndarray) and np.array_equal(i2, gold)


def test_nms_edge_cases():
    empty = torch.zeros(0, 5, device="cuda")
    kept, inds = P.nms(empty, 0.5)
    assert kept.shape == (0, 5) and inds.numel() == 0 and inds.dtype == torch.int64
    # IoU exactly == thr is NOT suppressed by the CUDA op (nms_kernel.cu:60, strict >)
    d = torch.tensor([[0, 0, 9, 9, 0.9], [0, 0, 9, 19, 0.8]], device="cuda")
    assert P.nms(d, 0.5)[1].tolist() == [0, 1]
    assert P.nms(d, 0.49)[1].tolist() == [0]
    with pytest.raises(TypeError):
        P.nms([1, 2, 3], 0.5)
    with pytest.raises(RuntimeError):
        P.nms(torch.zeros(3, 5), 0.5)                        # CPU tensor: no fallback


@pytest.mark.parametrize("name", list(cases.nms_large_inputs().keys()))
def test_nms_beyond_one_block_bit_exact(name):
    """n > IOU_MAX_NMS_BOXES (6144): the mask-tile path of iou_nms (sort + 64x64 mask tiles + greedy scan, all on the
    device) == the reference's nms_cpu golden; the reference has no size limit (nms_kernel.cu:70-131)."""
    gold = np.load(os.path.join(U.GOLD, "boundary_ops.npz"))["nms_" + name]
    dets = torch.from_numpy(cases.nms_large_inputs()[name]).cuda()
    kept, inds = P.nms(dets, cases.NMS_LARGE_THR)
    assert inds.dtype == torch.int64 and inds.is_cuda
    assert np.array_equal(inds.cpu().numpy(), gold)
    assert torch.equal(kept, dets[inds])


def test_nms_both_paths_agree_across_the_size_limit():
    rs = np.random.RandomState(9)
    big = cases.random_dets(rs, 6500, 700, 140)
    for n in (6143, 6144, 6145, 6208, 6500):
        want = op.nms(big[:n], 0.45, "cuda").numpy()
        got = P.nms(torch.from_numpy(big[:n]).cuda(), 0.45)[1].cpu().numpy()
        assert np.array_equal(got, want), n


def test_nms_random_vs_oracle_many_sizes():
    rs = np.random.RandomState(3)
    for n in (2, 31, 32, 33, 63, 64, 65, 127, 128, 129, 500, 1025, 3000, 6144):
        dets = cases.random_dets(rs, n, 400, 150)
        want = op.nms(dets, 0.45, "cuda").numpy()
        got = P.nms(torch.from_numpy(dets).cuda(), 0.45)[1].cpu().numpy()
        assert np.array_equal(got, want), n


def test_nms_duplicate_scores_tie_rule():
    """Equal scores: visiting order = ascending index (documented tie rule, same as the oracle)."""
    rs = np.random.RandomState(5)
    dets = cases.random_dets(rs, 300, 200, 90)
    dets[:, 4] = np.round(dets[:, 4] * 4) / 4
    want = op.nms(dets, 0.5, "cuda").numpy()
    got = P.nms(torch.from_numpy(dets).cuda(), 0.5)[1].cpu().numpy()
    assert np.array_equal(got, want)


def test_multiclass_nms_api_vs_oracle():
    rs = np.random.RandomState(8)
    n, C = 900, 80
    boxes = cases.random_dets(rs, n, 500, 160)[:, :4]
    scores = (rs.rand(n, C + 1) ** 6).astype(np.float32)
    scores[:, 0] = 0
    for max_num in (100, 3000):
        d_ref, l_ref = op.multiclass_nms(torch.from_numpy(boxes), torch.from_numpy(scores), 0.05, 0.5, max_num)
        # max_num = 3000: 80 x 3001 kept-row slots exceed the batched kernels -> the class-by-class path
        d, l = P.multiclass_nms(torch.from_numpy(boxes).cuda(), torch.from_numpy(scores).cuda(), 0.05,
                                dict(type='nms', iou_thr=0.5), max_num)
        assert np.array_equal(l.cpu().numpy(), l_ref.numpy())
        assert np.array_equal(d.cpu().numpy(), d_ref.numpy())
    # nothing above the threshold -> empty (0,5) / (0,) (bbox_nms.py:63-65)
    d, l = P.multiclass_nms(torch.from_numpy(boxes).cuda(), torch.zeros(n, C + 1).cuda(), 0.05,
                            dict(type='nms', iou_thr=0.5), 100)
    assert d.shape == (0, 5) and l.shape == (0,) and l.dtype == torch.int64


@pytest.mark.parametrize("name", list(cases.multiclass_inputs().keys()))
def test_multiclass_nms_options_vs_reference_golden(name):
    """max_num = -1 (the reference default, bbox_nms.py:6-11; :57-59 then sorts and drops the last row), max_num = 0,
    score_factors (threshold on the unscaled score, :37,46-47), per-class boxes (:43-44) and soft_nms: outputs of
    the reference's own multiclass_nms."""
    gold = np.load(os.path.join(U.GOLD, "boundary_ops.npz"))
    b, sc, thr, nms_cfg, max_num, fac = cases.multiclass_inputs()[name]
    d, l = P.multiclass_nms(torch.from_numpy(b).cuda(), torch.from_numpy(sc).cuda(), thr, dict(nms_cfg), max_num,
                            None if fac is None else torch.from_numpy(fac).cuda())
    gd, gl = gold["mc_%s_dets" % name], gold["mc_%s_labels" % name]
    assert d.shape == gd.shape and l.dtype == torch.int64
    d, l = d.cpu().numpy(), l.cpu().numpy()
    assert np.array_equal(d[:, 4], gd[:, 4])                  # the score order itself is exact
    # rows with EQUAL scores come out of torch.sort (not stable, :58) in an unspecified order: compare those as sets

    def canon(dd, ll):
        o = np.lexsort((dd[:, 0], dd[:, 1], ll, -dd[:, 4].astype(np.float64)))
        return dd[o], ll[o]
    (d, l), (gd, gl) = canon(d, l), canon(gd, gl)
    assert np.array_equal(l, gl)
    assert np.array_equal(d, gd)


def test_get_bboxes_outside_the_kernel_envelope_vs_oracle():
    """test_cfg beyond the batched kernels (nms_pre > 2048, classes x (max_per_img + 1) > 8192): get_bboxes still
    answers, through the reference's own level-by-level / class-by-class schedule on the device."""
    case = cases.postproc_case("small")
    head = U.get_head()
    cfgd = dict(case["cfg"])
    cfgd.update(nms_pre=2500, max_per_img=150)
    cfg = P.ConfigDict(cfgd)
    sizes = [tuple(t.shape[-2:]) for t in case["cls"]]
    assert head.in_kernel_envelope(sizes, P.ConfigDict(case["cfg"])) and not head.in_kernel_envelope(sizes, cfg)
    dev = torch.device("cuda:0")
    cls, reg, iou = [[t.to(dev) for t in ts] for ts in (case["cls"], case["reg"], case["iou"])]
    n_img = cls[0].shape[0]
    res = head.get_bboxes(cls, reg, iou, [None] * n_img, [None] * n_img, case["img_metas"], cfg, rescale=True)
    bases = U.oracle_bases()
    for i, (d, l) in enumerate(res):
        m = case["img_metas"][i]
        d_ref, l_ref = op.get_bboxes_single([c[i] for c in case["cls"]], [r[i] for r in case["reg"]],
                                            [q[i] for q in case["iou"]], cases.STRIDES, bases, m["img_shape"],
                                            m["scale_factor"], cfgd, rescale=True)
        assert d.shape[0] == d_ref.shape[0] == 150 and l.dtype == torch.int64
        frac, ms, mb = U.match_as_sets(d.cpu().numpy(), l.cpu().numpy(), d_ref.numpy(), l_ref.numpy(), min_frac=1.0,
                                       score_tol=1e-5, box_tol=1e-3)
        assert frac == 1.0


def test_topk_with_exact_ties_is_deterministic():
    """Many identical max scores at the top-k boundary: lowest anchor indices win (documented rule)."""
    case = cases.postproc_case("small")
    head = U.get_head()
    dev = torch.device("cuda:0")
    cls = [torch.full_like(t[:1], -2.0).to(dev) for t in case["cls"]]
    iou = [torch.zeros_like(t[:1]).to(dev) for t in case["iou"]]
    reg = [t[:1].to(dev) for t in case["reg"]]
    cfg = P.ConfigDict(case["cfg"])
    wsp = head.postproc_workspace([tuple(t.shape[-2:]) for t in cls], 1, cfg, dev)
    info = PP.make_img_info(case["img_metas"][:1], dev)
    _, _, idx = PP.decode_candidates(wsp, cls, reg, iou, info, False)
    idx = idx[0].cpu().numpy()
    off = 0
    for t in cls:
        n_l = t.shape[-1] * t.shape[-2] * 9
        k = min(n_l, cfg.nms_pre)
        assert np.array_equal(idx[off:off + k], np.arange(k)), "level with %d anchors" % n_l
        off += k


@pytest.mark.parametrize("regime", ["reference_init", "spread", "two_clusters"])
def test_topk_histogram_select_vs_oracle(regime):
    """Per-level top-k = bucket histogram (built by max_score_kernel) + one-pass collect + sort, with the exact radix
    select behind it for score distributions one bucket cannot split.  'reference_init': every fused score is
    0.07098 +- 2e-5 -> the fallback; 'spread': the histogram path; 'two_clusters': both in one call.
    The candidate SET equals torch.topk's on the oracle's scores; the order is (score desc, index asc)."""
    rs = np.random.RandomState({"reference_init": 1, "spread": 2, "two_clusters": 3}[regime])
    sizes = cases.level_sizes(416, 544)            # 31 824 / 7 956 / 1 989 / 567 / 162 anchors
    cls, reg, iou = [], [], []
    for (h, w) in sizes:
        if regime == "reference_init":
            c = rs.randn(1, 720, h, w) * 0.0016 - 4.5952
            q = rs.randn(1, 9, h, w) * 0.0019
        elif regime == "spread":
            c = rs.randn(1, 720, h, w) * 2.0 - 3.0
            q = rs.randn(1, 9, h, w) * 1.5
        else:
            c = np.where(rs.rand(1, 720, h, w) < 0.5, rs.randn(1, 720, h, w) * 0.001 + 1.0, rs.randn(1, 720, h, w) - 6.0)
            q = rs.randn(1, 9, h, w) * 0.001
        cls.append(torch.from_numpy(c.astype(np.float32)))
        iou.append(torch.from_numpy(q.astype(np.float32)))
        reg.append(torch.from_numpy((rs.randn(1, 36, h, w) * 0.5).astype(np.float32)))
    head = U.get_head()
    dev = torch.device("cuda:0")
    cfgd = dict(cases.TEST_CFG)
    cfg = P.ConfigDict(cfgd)
    meta = dict(img_shape=(416, 540, 3), scale_factor=1.0)
    wsp = head.postproc_workspace(sizes, 1, cfg, dev)
    info = PP.make_img_info([meta], dev)
    boxes, scores_cm, idx = PP.decode_candidates(wsp, [t.to(dev) for t in cls], [t.to(dev) for t in reg],
                                                 [t.to(dev) for t in iou], info, False)
    idx, sc = idx[0].cpu().numpy(), scores_cm[0].t().contiguous().cpu().numpy()
    _, o_sc, o_idx = op.candidates_single([t[0] for t in cls], [t[0] for t in reg], [t[0] for t in iou], cases.STRIDES,
                                          U.oracle_bases(), meta["img_shape"], 1.0, cfgd["nms_pre"])
    off = 0
    for (h, w) in sizes:
        n_l = h * w * 9
        k = min(n_l, cfgd["nms_pre"])
        mine, ref = idx[off:off + k], o_idx[off:off + k].numpy()
        my_key = sc[off:off + k].max(1)
        if n_l > cfgd["nms_pre"]:
            # order: score descending, ties by ascending anchor index
            assert np.all(np.diff(my_key) <= 0)
            assert np.all(np.diff(mine)[np.diff(my_key) == 0] > 0)
            # same set as torch.topk, except where keys tie with the k-th key to within fp32 noise of the fused score
            diff = set(mine.tolist()) ^ set(ref.tolist())
            kth = my_key[-1]
            ref_key = o_sc[off:off + k].numpy().max(1)
            assert all(True for _ in diff) and len(diff) <= 2 * int((np.abs(ref_key - kth) <= 2e-7).sum() + 1), (regime, len(diff))
        else:
            assert np.array_equal(mine, np.arange(n_l))
        off += k


def test_full_size_properties():
    """BASELINE-size invariants that need no oracle run: counts, ordering, clamping, idempotence."""
    case = cases.postproc_case("full")
    head = U.get_head()
    dev = torch.device("cuda:0")
    cfg = P.ConfigDict(case["cfg"])
    n = 4
    rs = np.random.RandomState(77)
    cls, reg, iou = cases.random_maps(rs, n, case["sizes"])
    cls, reg, iou = [t.to(dev) for t in cls], [t.to(dev) for t in reg], [t.to(dev) for t in iou]
    metas = case["img_metas"] * n
    r1 = head.get_bboxes(cls, reg, iou, None, None, metas, cfg, rescale=False)
    r2 = head.get_bboxes(cls, reg, iou, None, None, metas, cfg, rescale=False)
    for (d, l), (d2, l2) in zip(r1, r2):
        assert torch.equal(d, d2) and torch.equal(l, l2)                  # deterministic
        assert d.shape == (100, 5) and l.dtype == torch.int64
        s = d[:, 4]
        assert bool((s[:-1] >= s[1:]).all()) and bool((s > 0.05).all())  # sorted, thresholded
        assert bool((d[:, 0] >= 0).all()) and bool((d[:, 2] <= 1332).all()) and bool((d[:, 3] <= 799).all())
        assert bool((l >= 0).all()) and bool((l < 80).all())
    # batch independence: image i alone gives the same detections as inside the batch
    solo = head.get_bboxes([t[1:2] for t in cls], [t[1:2] for t in reg], [t[1:2] for t in iou], None, None,
                           metas[:1], cfg, rescale=False)
    assert torch.equal(solo[0][0], r1[1][0]) and torch.equal(solo[0][1], r1[1][1])


@pytest.mark.parametrize("name", ["small", "full"])
def test_get_bboxes_premax_equals_get_bboxes(name):
    """iou_get_bboxes_premax (per-anchor max class logit handed in as two partial maxima, the way the retina_cls conv
    epilogue produces them) returns bit for bit what iou_get_bboxes reduces from the class maps itself; a level without
    partials (NULL entry) falls back to its class map."""
    case = cases.postproc_case(name)
    head = U.get_head()
    dev = torch.device("cuda:0")
    cfg = P.ConfigDict(case["cfg"])
    cls = [t.to(dev) for t in case["cls"]]
    reg = [t.to(dev) for t in case["reg"]]
    iou = [t.to(dev) for t in case["iou"]]
    n_img = cls[0].shape[0]
    sizes = [tuple(t.shape[-2:]) for t in cls]
    info = PP.make_img_info(case["img_metas"], dev)
    outs = []
    for mode in ("plain", "premax", "mixed"):
        wsp = head.postproc_workspace(sizes, n_img, cfg, dev)
        pm = None
        if mode != "plain":
            pm = []
            for l, t in enumerate(cls):                      # (n, A*C, H, W) -> (n, H, W, A, C): even / odd 16-class chunks
                n, ac, h, w = t.shape
                v = t.permute(0, 2, 3, 1).reshape(n, h, w, 9, 5, 16)
                pm.append(torch.stack([v[..., 0::2, :].amax(dim=(-1, -2)), v[..., 1::2, :].amax(dim=(-1, -2))], dim=-1).contiguous())
            if mode == "mixed":
                pm[0] = None
        d, l_, c = PP.get_bboxes_device(wsp, cls, reg, iou, info, case["rescale"], cls_max2=pm)
        torch.cuda.synchronize()
        outs.append((d.clone(), l_.clone(), c.clone()))
    for o in outs[1:]:
        assert torch.equal(o[0], outs[0][0]) and torch.equal(o[1], outs[0][1]) and torch.equal(o[2], outs[0][2])
    assert int(outs[0][2].sum()) > 0


def test_focal_loss_forward_backward_vs_oracle():
    torch.manual_seed(0)
    x = (torch.randn(513, 80) * 3)
    t = torch.randint(0, 81, (513,))
    for gamma, alpha in ((2.0, 0.25), (1.5, 0.4)):
        ref_f = op.sigmoid_focal_loss_forward(x, t, gamma, alpha)
        ref_b = op.sigmoid_focal_loss_backward(x, t, torch.ones_like(x), gamma, alpha)
        f = P.sigmoid_focal_loss_cuda.forward(x.cuda(), t.cuda(), 80, gamma, alpha)
        b = P.sigmoid_focal_loss_cuda.backward(x.cuda(), t.cuda(), torch.ones_like(x).cuda(), 80, gamma, alpha)
        assert torch.allclose(f.cpu(), ref_f, rtol=1e-4, atol=1e-6)
        assert torch.allclose(b.cpu(), ref_b, rtol=1e-4, atol=1e-6)
    xg = x.cuda().requires_grad_(True)
    loss = P.sigmoid_focal_loss(xg, t.cuda(), 2.0, 0.25, 'mean')
    loss.backward()
    assert torch.allclose(loss.detach().cpu(), op.sigmoid_focal_loss_forward(x, t, 2.0, 0.25).mean(), rtol=1e-4)
    assert torch.allclose(xg.grad.cpu(), op.sigmoid_focal_loss_backward(
        x, t, torch.full_like(x, 1.0 / x.numel()), 2.0, 0.25), rtol=1e-4, atol=1e-9)
    # the module returns what modules/sigmoid_focal_loss.py:13-16 returns: sigmoid_focal_loss(...) with its default
    # reduction ('mean') and .sum() of that 0-dim tensor = the MEAN over the N*C elements
    assert P.SigmoidFocalLoss(2.0, 0.25)(x.cuda(), t.cuda()).item() == pytest.approx(
        op.sigmoid_focal_loss_forward(x, t, 2.0, 0.25).mean().item(), rel=1e-4)
    for red, f in (('none', lambda v: v), ('sum', lambda v: v.sum())):
        got = P.sigmoid_focal_loss(x.cuda(), t.cuda(), 2.0, 0.25, red)
        assert torch.allclose(got.cpu(), f(op.sigmoid_focal_loss_forward(x, t, 2.0, 0.25)), rtol=1e-4, atol=1e-6)
    with pytest.raises(ValueError):
        P.sigmoid_focal_loss(x.cuda(), t.cuda(), 2.0, 0.25, 'bogus')


@pytest.mark.parametrize("dtype,rtol,atol", [(torch.float16, 2e-3, 2e-4), (torch.float64, 1e-5, 1e-7)])
def test_focal_loss_fp16_fp64_dispatch_vs_oracle(dtype, rtol, atol):
    """AT_DISPATCH_FLOATING_TYPES_AND_HALF (sigmoid_focal_loss_cuda.cu:128,167): fp16 and fp64 logits keep their
    dtype; the oracle restates the templated kernel per scalar_t (float transcendentals, scalar_t locals)."""
    torch.manual_seed(1)
    x = (torch.randn(257, 80) * 3).to(dtype)
    t = torch.randint(0, 81, (257,))
    g = torch.rand(257, 80).to(dtype)
    for gamma, alpha in ((2.0, 0.25), (1.5, 0.4)):
        f = P.sigmoid_focal_loss_cuda.forward(x.cuda(), t.cuda(), 80, gamma, alpha)
        b = P.sigmoid_focal_loss_cuda.backward(x.cuda(), t.cuda(), g.cuda(), 80, gamma, alpha)
        assert f.dtype == dtype and b.dtype == dtype
        ref_f = op.sigmoid_focal_loss_forward_typed(x, t, gamma, alpha, dtype)
        ref_b = op.sigmoid_focal_loss_backward_typed(x, t, g, gamma, alpha, dtype)
        assert torch.allclose(f.cpu().double(), ref_f.double(), rtol=rtol, atol=atol), (f.cpu().double() - ref_f.double()).abs().max()
        assert torch.allclose(b.cpu().double(), ref_b.double(), rtol=rtol, atol=atol), (b.cpu().double() - ref_b.double()).abs().max()
        # and both agree with the fp32 formula to the dtype's own precision
        f32 = op.sigmoid_focal_loss_forward(x.float(), t, gamma, alpha)
        assert torch.allclose(f.cpu().float(), f32, rtol=5e-3 if dtype == torch.float16 else 1e-5, atol=1e-3 if dtype == torch.float16 else 1e-6)
    with pytest.raises(RuntimeError):
        P.sigmoid_focal_loss_cuda.forward(x.cuda().to(torch.bfloat16), t.cuda(), 80, 2.0, 0.25)


def test_plain_retina_head_get_bboxes_vs_reference_golden():
    """RetinaHead (alpha = 1, no IoU maps) through the same kernels vs the reference's own RetinaHead."""
    gold = np.load(os.path.join(U.GOLD, "postproc_plain_retina_small.npz"))
    case = cases.postproc_case("small")
    cfgd = dict(U.head_cfg(), type='RetinaHead')
    torch.manual_seed(0)
    head = P.build_head(cfgd)
    assert sorted(k.split('.')[0] for k in head.state_dict()).count('retina_iou') == 0
    dev = torch.device("cuda:0")
    cls = [t.to(dev) for t in case["cls"]]
    reg = [t.to(dev) for t in case["reg"]]
    n = cls[0].shape[0]
    res = head.get_bboxes(cls, reg, [None] * n, [None] * n, case["img_metas"], P.ConfigDict(case["cfg"]),
                          rescale=case["rescale"])
    for i, (d, l) in enumerate(res):
        gd, gl = gold["dets_%d" % i], gold["labels_%d" % i]
        assert d.shape[0] == gd.shape[0]
        U.match_as_sets(d.cpu().numpy(), l.cpu().numpy(), gd, gl, min_frac=0.97)


# ------------------------------------------------------------------ Soft-NMS
@pytest.mark.parametrize("name", list(cases.soft_nms_inputs().keys()))
def test_soft_nms_bit_exact_vs_reference_golden(name):
    """iou_soft_nms == the reference's soft_nms_cpu.pyx: survivors, order and decayed scores, bit for bit."""
    g = np.load(os.path.join(U.GOLD, "soft_nms.npz"))
    d, thr, method, sigma, min_score = cases.soft_nms_inputs()[name]
    nd, inds = P.soft_nms(torch.from_numpy(d).cuda(), thr, method=method, sigma=sigma, min_score=min_score)
    assert nd.is_cuda and inds.is_cuda and inds.dtype == torch.int64
    assert np.array_equal(inds.cpu().numpy(), g[name + "_inds"])
    assert np.array_equal(nd.cpu().numpy().view(np.uint32), g[name + "_dets"].view(np.uint32))
    # numpy in -> numpy out (nms_wrapper.py:53-58,75-78)
    nd2, i2 = P.soft_nms(d, thr, method=method, sigma=sigma, min_score=min_score)
    assert isinstance(nd2, np.ndarray) and nd2.dtype == np.float32 and i2.dtype == np.int64
    assert np.array_equal(i2, g[name + "_inds"])


def test_soft_nms_random_vs_oracle_and_errors():
    from oracle import soft_nms as SN
    rs = np.random.RandomState(21)
    for trial in range(40):
        n = int(rs.choice([1, 2, 31, 33, 255, 256, 257, 600, 1025, 2500]))
        d = cases.random_dets(rs, n, float(rs.choice([60., 300.])), float(rs.choice([40., 120.])))
        if trial % 3 == 0:
            d[:, 4] = np.round(d[:, 4] * 8) / 8                         # tie order must follow the swap bookkeeping
        method = ['linear', 'gaussian'][trial % 2]
        thr, sig, ms = float(rs.choice([0.3, 0.5])), float(rs.choice([0.3, 0.5])), float(rs.choice([1e-3, 0.05, 0.3]))
        want_d, want_i = SN.soft_nms(d, thr, method=method, sigma=sig, min_score=ms)
        got_d, got_i = P.soft_nms(torch.from_numpy(d).cuda(), thr, method=method, sigma=sig, min_score=ms)
        assert np.array_equal(got_i.cpu().numpy(), want_i), (trial, n, method)
        assert np.array_equal(got_d.cpu().numpy().view(np.uint32), want_d.view(np.uint32)), (trial, n, method)
    e_d, e_i = P.soft_nms(torch.zeros(0, 5, device="cuda"), 0.5)
    assert e_d.shape == (0, 5) and e_i.shape == (0,)
    with pytest.raises(ValueError):
        P.soft_nms(torch.zeros(3, 5, device="cuda"), 0.5, method='bogus')
    with pytest.raises(TypeError):
        P.soft_nms([[0, 0, 1, 1, 0.5]], 0.5)
    with pytest.raises(RuntimeError):
        P.soft_nms(torch.zeros(3, 5), 0.5)                               # CPU tensor: no fallback
    with pytest.raises(RuntimeError):
        P.soft_nms(torch.zeros(7000, 5, device="cuda"), 0.5)             # above IOU_MAX_NMS_BOXES


def test_multiclass_soft_nms_vs_reference_golden():
    """multiclass_nms(nms_cfg type 'soft_nms') == the reference's on the 'small' case candidates."""
    g = np.load(os.path.join(U.GOLD, "soft_nms.npz"))
    c = np.load(os.path.join(U.GOLD, "postproc_small.npz"))
    for i in range(2):
        boxes = torch.from_numpy(c["cand_boxes_%d" % i]).cuda()
        scores = torch.from_numpy(c["cand_scores_%d" % i]).cuda()
        padded = torch.cat([scores.new_zeros(scores.shape[0], 1), scores], dim=1)
        d, l = P.multiclass_nms(boxes, padded, 0.05, dict(cases.SOFT_MULTICLASS), 100)
        want_d, want_l = g["mc_dets_%d" % i], g["mc_labels_%d" % i]
        # the final top-100 cut sorts by score (torch.sort, unstable in the reference): compare as sorted sets
        assert d.shape == want_d.shape
        key = lambda D, Lb: sorted(map(tuple, np.concatenate([D, Lb[:, None].astype(np.float32)], 1).tolist()))
        assert key(d.cpu().numpy(), l.cpu().numpy()) == key(want_d, want_l)
    # few candidates (no top-100 cut): class-major, selection order inside a class, exactly as the oracle
    rs = np.random.RandomState(3)
    n, C = 300, 8
    boxes = cases.random_dets(rs, n, 300, 120)[:, :4]
    scores = (rs.rand(n, C + 1) ** 8).astype(np.float32)
    scores[:, 0] = 0
    want_d, want_l = op.multiclass_nms(torch.from_numpy(boxes), torch.from_numpy(scores), 0.3, 0.5, 1000,
                                       soft=dict(method='gaussian', sigma=0.5, min_score=0.05))
    d, l = P.multiclass_nms(torch.from_numpy(boxes).cuda(), torch.from_numpy(scores).cuda(), 0.3,
                            dict(type='soft_nms', iou_thr=0.5, method='gaussian', sigma=0.5, min_score=0.05), 1000)
    assert np.array_equal(l.cpu().numpy(), want_l.numpy())
    assert np.array_equal(d.cpu().numpy().view(np.uint32), want_d.numpy().view(np.uint32))


def test_head_get_bboxes_with_soft_nms_test_cfg():
    """IoUawareRetinaHead.get_bboxes with test_cfg.nms = dict(type='soft_nms', ...) vs the reference's own
    output on the 'small' case maps (scores/boxes within the fp32 tolerance of the decode stage)."""
    g = np.load(os.path.join(U.GOLD, "soft_nms.npz"))
    case = cases.postproc_case("small")
    cfgd = dict(case["cfg"])
    cfgd["nms"] = dict(cases.SOFT_MULTICLASS)
    head = U.get_head()
    dev = torch.device("cuda:0")
    n_img = case["cls"][0].shape[0]
    res = head.get_bboxes([t.to(dev) for t in case["cls"]], [t.to(dev) for t in case["reg"]],
                          [t.to(dev) for t in case["iou"]], [None] * n_img, [None] * n_img, case["img_metas"],
                          P.ConfigDict(cfgd), rescale=case["rescale"])
    for i, (d, l) in enumerate(res):
        gd, gl = g["gb_dets_%d" % i], g["gb_labels_%d" % i]
        assert d.shape == gd.shape and l.dtype == torch.int64
        d, l = d.cpu().numpy(), l.cpu().numpy()
        o1 = np.lexsort((d[:, 0], l, -d[:, 4]))
        o2 = np.lexsort((gd[:, 0], gl, -gd[:, 4]))
        assert np.array_equal(l[o1], gl[o2])
        assert np.allclose(d[o1], gd[o2], rtol=U.RTOL, atol=U.ATOL), np.abs(d[o1] - gd[o2]).max()


def test_fcos_head_get_bboxes_vs_reference_golden():
    """IoUawareFCOSHead.get_bboxes (alpha = 0.3, one location per cell, distance2bbox) through the same kernels.
    Candidates: same set/order as the reference (near-ties aside), boxes/scores within the fp32 tolerance (pow with
    a non-trivial exponent differs by ~1 ulp between torch-CPU and CUDA); stage 2 on the GOLDEN candidates is
    bit-exact; the head's public get_bboxes matches the reference detections within tolerance."""
    g = np.load(os.path.join(U.GOLD, "postproc_fcos.npz"))
    case = cases.fcos_case()
    dev = torch.device("cuda:0")
    head = P.IoUawareFCOSHead(81, 256, strides=cases.FCOS_STRIDES)
    cfg = P.ConfigDict(case["cfg"])
    cls = [t.to(dev) for t in case["cls"]]
    reg = [t.to(dev) for t in case["reg"]]
    cen = [t.to(dev) for t in case["cen"]]
    iou = [t.to(dev) for t in case["iou"]]
    n_img = 2
    wsp = head.postproc_workspace([tuple(t.shape[-2:]) for t in cls], n_img, cfg, dev)
    info = PP.make_img_info(case["img_metas"], dev)
    boxes, scores_cm, idx = PP.decode_candidates(wsp, cls, reg, iou, info, True)
    for i in range(n_img):
        my_idx, g_idx = idx[i].cpu().numpy(), g["cand_idx_%d" % i]
        bad = int((my_idx != g_idx).sum())
        assert bad <= 8 and sorted(my_idx.tolist()) == sorted(g_idx.tolist()), bad
        if bad == 0:
            assert np.allclose(boxes[i].cpu().numpy(), g["cand_boxes_%d" % i], rtol=U.RTOL, atol=U.ATOL)
            assert np.allclose(scores_cm[i].t().cpu().numpy(), g["cand_scores_%d" % i], rtol=U.RTOL, atol=U.ATOL)
    gb = torch.stack([torch.from_numpy(g["cand_boxes_%d" % i]) for i in range(n_img)]).to(dev)
    gs = torch.stack([torch.from_numpy(g["cand_scores_%d" % i]).t().contiguous() for i in range(n_img)]).to(dev)
    dets, labels, counts = PP.batched_nms(wsp, gb, gs)
    for i in range(n_img):
        want_d, want_l = op.multiclass_nms(torch.from_numpy(g["cand_boxes_%d" % i]),
                                           torch.cat([torch.zeros(gs.shape[2], 1), torch.from_numpy(g["cand_scores_%d" % i])], 1),
                                           0.05, 0.5, 100)
        k = int(counts[i])
        assert k == want_d.shape[0]
        assert np.array_equal(labels[i, :k].cpu().numpy(), want_l.numpy())
        assert np.array_equal(dets[i, :k].cpu().numpy(), want_d.numpy())
    res = head.get_bboxes(cls, reg, cen, iou, [None] * n_img, [None] * n_img, case["img_metas"], cfg, rescale=True)
    for i, (d, l) in enumerate(res):
        gd, gl = g["dets_%d" % i], g["labels_%d" % i]
        d, l = d.cpu().numpy(), l.cpu().numpy()
        assert d.shape == gd.shape
        o1, o2 = np.lexsort((d[:, 0], l, -d[:, 4])), np.lexsort((gd[:, 0], gl, -gd[:, 4]))
        assert np.array_equal(l[o1], gl[o2])
        assert np.allclose(d[o1], gd[o2], rtol=U.RTOL, atol=U.ATOL), np.abs(d[o1] - gd[o2]).max()
