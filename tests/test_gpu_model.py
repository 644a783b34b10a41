"""-m gpu: module-level and whole-detector parity against the oracle's torch-CPU fp32 forward."""
import os

import numpy as np
import pytest
import torch

import cases
import parity_util as U
from oracle import model as om
from oracle import postproc as op

pytestmark = pytest.mark.gpu
P = U.P
DEV = "cuda:0"


def _rel(a, b):
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-6)).item()


def test_backbone_fpn_head_modules_standalone():
    """Each module called on its own with plain NCHW tensors, like the reference's modules."""
    det, cfg = U.small_detector(seed=2)
    sd = {k: v.clone() for k, v in det.state_dict().items()}
    det = det.to(DEV)
    img = torch.randn(1, 3, 96, 128, generator=torch.Generator().manual_seed(9))
    ref_feats = om.backbone_forward(sd, img)
    feats = det.backbone(img.to(DEV))
    assert len(feats) == 4
    for a, b in zip(feats, ref_feats):
        assert a.shape == b.shape and _rel(a.cpu(), b) < 5e-4, _rel(a.cpu(), b)
    ref_p = om.fpn_forward(sd, ref_feats)
    outs = det.neck([f.to(DEV) for f in ref_feats])
    assert len(outs) == 5
    for a, b in zip(outs, ref_p):
        assert a.shape == b.shape and _rel(a.cpu(), b) < 5e-4, _rel(a.cpu(), b)
    ref_h = om.head_forward(sd, ref_p)
    cls, reg, iou = det.bbox_head(tuple(p.to(DEV) for p in ref_p))
    for mine, ref in ((cls, ref_h[0]), (reg, ref_h[1]), (iou, ref_h[2])):
        assert len(mine) == 5
        for a, b in zip(mine, ref):
            assert tuple(a.shape) == tuple(b.shape)          # logical (N, A*C, H, W) like the reference
            assert _rel(a.cpu(), b) < 5e-4, _rel(a.cpu(), b)
    c0, r0, q0 = det.bbox_head.forward_single(ref_p[2].to(DEV))
    assert _rel(c0.cpu(), ref_h[0][2]) < 5e-4
    # extract_feat == backbone + neck
    x = det.extract_feat(img.to(DEV))
    for a, b in zip(x, ref_p):
        assert _rel(a.cpu(), b) < 5e-4


def test_public_forwards_return_fresh_tensors():
    """The reference's modules return independent tensors: a second call with the same input shape must not
    overwrite what the first call returned (the plans' own buffers stay internal)."""
    det, cfg = U.small_detector(seed=2)
    det = det.to(DEV)
    g = torch.Generator().manual_seed(4)
    x1, x2 = torch.randn(1, 3, 96, 128, generator=g).to(DEV), torch.randn(1, 3, 96, 128, generator=g).to(DEV)
    f1 = det.backbone(x1)
    keep = [t.clone() for t in f1]
    f2 = det.backbone(x2)
    assert all(torch.equal(a, b) for a, b in zip(f1, keep)) and not torch.equal(f1[0], f2[0])
    p1 = det.neck(f1)
    keep = [t.clone() for t in p1]
    p2 = det.neck(f2)
    assert all(torch.equal(a, b) for a, b in zip(p1, keep)) and not torch.equal(p1[0], p2[0])
    h1 = det.bbox_head(p1)
    keep = [[t.clone() for t in ts] for ts in h1]
    h2 = det.bbox_head(p2)
    assert all(torch.equal(a, b) for ts, ks in zip(h1, keep) for a, b in zip(ts, ks))
    assert not torch.equal(h1[0][0], h2[0][0])
    metas = [dict(ori_shape=(96, 128, 3), img_shape=(96, 128, 3), pad_shape=(96, 128, 3), scale_factor=1.0, flip=False)]
    d1 = det.detect_device(x1, metas)
    keep = [t.clone() for t in d1]
    det.detect_device(x2, metas)
    assert all(torch.equal(a, b) for a, b in zip(d1, keep))


def test_detector_end_to_end_small():
    worst = U.check_detector_small(use_graph=False)
    print("worst relative head error", worst)
    assert worst < 1e-3


def test_detector_cuda_graph_replay_matches_eager():
    U.check_detector_small(use_graph=True)


def test_forward_test_reference_signature():
    """detector(return_loss=False, rescale=..., img=[T], img_meta=[[meta]], gt_*) (base.py:105-123)."""
    det, cfg = U.small_detector(seed=4)
    det = det.to(DEV)
    h, w = 96, 128
    img = torch.randn(1, 3, h, w, generator=torch.Generator().manual_seed(1)).to(DEV)
    meta = dict(ori_shape=(h, w, 3), img_shape=(h, w, 3), pad_shape=(h, w, 3), scale_factor=1.0, flip=False)
    out = det(return_loss=False, rescale=True, img=[img], img_meta=[[meta]],
              gt_bboxes=[[torch.zeros(0, 4)]], gt_labels=[[torch.zeros(0, dtype=torch.long)]])
    assert isinstance(out, list) and len(out) == 80
    assert all(isinstance(a, np.ndarray) and a.dtype == np.float32 and a.shape[1] == 5 for a in out)
    with pytest.raises(AssertionError):      # the reference asserts one image per GPU here
        det(return_loss=False, img=[img.repeat(2, 1, 1, 1)], img_meta=[[meta, meta]],
            gt_bboxes=[[None, None]], gt_labels=[[None, None]])
    with pytest.raises(TypeError):
        det(return_loss=False, img=img, img_meta=[[meta]], gt_bboxes=[[None]], gt_labels=[[None]])


def test_reference_init_degenerate_scores():
    """Reference init (no spread): every score ~ sqrt(0.01*0.5)=0.0707 -> worst-case NMS
    load; checks the path stays finite and returns max_per_img detections."""
    det, cfg = U.small_detector(seed=0, spread=False)
    det = det.to(DEV)
    h, w = 96, 128
    img = torch.randn(2, 3, h, w, generator=torch.Generator().manual_seed(1)).to(DEV)
    meta = dict(ori_shape=(h, w, 3), img_shape=(h, w, 3), pad_shape=(h, w, 3), scale_factor=1.0, flip=False)
    dets, labels, counts = det.detect_device(img, [meta, meta])
    torch.cuda.synchronize()
    assert counts.tolist() == [100, 100]
    s = dets[..., 4]
    assert bool(torch.isfinite(dets).all()) and float((s - 0.07098).abs().max()) < 2e-3


@pytest.mark.parametrize("cfg_name,depth,groups", [("iou_aware_retinanet_r101_fpn_1x_4gpu.py", 101, 1),
                                                    ("iou_aware_retinanet_x101_32x4d_fpn_1x_4gpu.py", 101, 32),
                                                    ("iou_aware_retinanet_x101_64x4d_fpn_1x.py", 101, 64)])
def test_other_backbones_head_maps_vs_oracle(cfg_name, depth, groups):
    """BASELINE configs 2-4: R101 / X101-32x4d / X101-64x4d through the same engine."""
    cfg = P.Config.fromfile(os.path.join(U.CFG_DIR, cfg_name))
    cfg.model.pretrained = None
    torch.manual_seed(5)
    det = P.build_detector(cfg.model, train_cfg=None, test_cfg=cfg.test_cfg).eval()
    sd = det.state_dict()
    om.spread_weights_(sd, seed=6, depth=depth, groups=groups)
    det.load_state_dict(sd)
    sd = {k: v.clone() for k, v in det.state_dict().items()}
    det = det.to(DEV)
    det.use_cuda_graph = False
    h, w = 96, 128
    img = torch.randn(1, 3, h, w, generator=torch.Generator().manual_seed(2))
    meta = dict(ori_shape=(h, w, 3), img_shape=(h, w, 3), pad_shape=(h, w, 3), scale_factor=1.0, flip=False)
    det.simple_test_batch(img.to(DEV), [meta])
    plan = det.fused_plan(img.shape, torch.device(DEV), False)
    torch.cuda.synchronize()
    ref = om.detector_forward(sd, img, depth, groups)
    for mine, r in zip(plan.outs, ref):
        for a, b in zip(mine, r):
            err = (a.cpu().contiguous() - b).abs().max().item()
            assert err <= 3e-4 * max(b.abs().max().item(), 1.0), (cfg_name, err)


def test_detect_stream_matches_detect_device():
    """The pipelined host-batch API returns exactly what the one-shot API returns, batch by batch."""
    det, cfg = U.small_detector(seed=7)
    det = det.to(DEV)
    h, w = 96, 128
    meta = dict(ori_shape=(h, w, 3), img_shape=(h, w, 3), pad_shape=(h, w, 3), scale_factor=1.0, flip=False)
    imgs = [torch.randn(2, 3, h, w, generator=torch.Generator().manual_seed(s)).pin_memory() for s in range(4)]
    want = []
    for im in imgs:
        d, l, c = det.detect_device(im, [meta, meta], rescale=True)
        want.append((d.cpu().clone(), l.cpu().clone(), c.cpu().clone()))
    imgs5 = imgs + [imgs[1]]                       # odd count: the last batch sits alone in its pipeline slot
    want5 = want + [want[1]]
    for depth in (2, 1):                           # two plans on two streams / one plan on the caller's stream
        got = list(det.detect_stream(((im, [meta, meta]) for im in imgs5), rescale=True, depth=depth))
        assert len(got) == 5
        for (d, l, c), (d2, l2, c2) in zip(want5, got):
            assert torch.equal(c, c2) and torch.equal(d, d2) and torch.equal(l, l2)
    assert list(det.detect_stream(iter(()), rescale=True)) == []
    # the side-stream packed gather (what N > 1 ranks use; world 1 = a device copy) and the legacy gather callable
    from iou_aware_single_stage_object_detector_b200 import dist as D
    for gather in (D.PackedGather(1, DEV), lambda d, l, c: (d, l, c)):
        got = list(det.detect_stream(((im, [meta, meta]) for im in imgs5), rescale=True, gather=gather))
        assert len(got) == 5
        for (d, l, c), (d2, l2, c2) in zip(want5, got):
            assert torch.equal(c, c2) and torch.equal(d, d2) and torch.equal(l, l2)


def test_unpadded_input_raises_like_the_reference():
    """An input whose FPN levels are not exactly 2x apart makes the reference fail in the top-down add
    (fpn.py:108-110); the engine raises as well instead of silently producing something."""
    det, cfg = U.small_detector(seed=0, spread=False)
    det = det.to(DEV)
    img = torch.randn(1, 3, 100, 136).to(DEV)            # 100/8 = 13 (ceil) vs 2*7 = 14
    meta = dict(ori_shape=(100, 136, 3), img_shape=(100, 136, 3), pad_shape=(100, 136, 3), scale_factor=1.0, flip=False)
    with pytest.raises(RuntimeError):
        det.simple_test_batch(img, [meta])


def test_preprocess_kernel_bit_exact_and_uint8_stream():
    """iou_preprocess_u8 == numpy ImageTransform restatement (oracle/preprocess.py), bit for bit; and
    the uint8 streaming path gives the same detections as feeding the normalised fp32 tensor."""
    from oracle import preprocess as opp
    rs = np.random.RandomState(4)
    mean, std = [123.675, 116.28, 103.53], [58.395, 57.12, 57.375]
    tf = P.ImageTransform(mean, std, to_rgb=True, size_divisor=32)
    for (h, w, flip) in ((90, 125, False), (96, 128, True), (33, 47, False)):
        imgs = rs.randint(0, 256, size=(2, h, w, 3)).astype(np.uint8)
        out, img_shape, pad_shape = tf(torch.from_numpy(imgs).to(DEV), flip=flip)
        for i in range(2):
            ref, rs_shape, rp_shape = opp.image_transform(imgs[i], mean, std, True, 32, flip)
            assert tuple(out[i].shape) == ref.shape and pad_shape == rp_shape and img_shape == rs_shape
            assert np.array_equal(out[i].cpu().numpy(), ref)
    det, cfg = U.small_detector(seed=9)
    det = det.to(DEV)
    h, w = 90, 125
    frames = [torch.from_numpy(rs.randint(0, 256, size=(2, h, w, 3)).astype(np.uint8)).pin_memory() for _ in range(3)]
    meta = dict(ori_shape=(h, w, 3), img_shape=(h, w, 3), pad_shape=(96, 128, 3), scale_factor=1.0, flip=False)
    got = list(det.detect_stream(((f, [meta, meta]) for f in frames), rescale=False, img_transform=tf))
    for f, (d, l, c) in zip(frames, got):
        x = torch.stack([torch.from_numpy(opp.image_transform(f[i].numpy(), mean, std, True, 32)[0]) for i in range(2)])
        d2, l2, c2 = det.detect_device(x.to(DEV), [meta, meta], rescale=False)
        assert torch.equal(c, c2.cpu()) and torch.equal(d, d2.cpu()) and torch.equal(l, l2.cpu())


def test_full_size_dense_path_properties():
    """BASELINE size (800x1344, R50-FPN): size-independent properties of the whole dense path + get_bboxes, and
    one image's head maps against a torch fp32 reference of the same network evaluated on the GPU (cuDNN, TF32
    off) -- the CPU oracle needs ~10 s per full-size image, which is what tests/test_gpu_model.py's small cases
    and bench.py's cpu_baseline spend it on."""
    from oracle import model as om
    det, cfg = U.small_detector()
    sd = {k: v.clone() for k, v in det.state_dict().items()}
    dev = torch.device("cuda:0")
    det = det.to(dev)
    n, h, w = 4, 800, 1344
    g = torch.Generator().manual_seed(11)
    img = torch.randn(n, 3, h, w, generator=g).to(dev)
    metas = [dict(ori_shape=(h, 1333, 3), img_shape=(h, 1333, 3), pad_shape=(h, w, 3), scale_factor=1.0,
                  flip=False) for _ in range(n)]

    def run(x):
        d, l, c = det.detect_device(x, metas, rescale=False)
        torch.cuda.synchronize()
        return d.clone(), l.clone(), c.clone()
    d0, l0, c0 = run(img)
    assert int(c0.min()) > 0
    # (1) idempotence: the same launch sequence gives the same bits
    d1, l1, c1 = run(img)
    assert torch.equal(d0, d1) and torch.equal(l0, l1) and torch.equal(c0, c1)
    # (2) images are independent (iou_aware_retina_head.py:434-460): permuting the batch permutes the detections
    perm = torch.tensor([2, 0, 3, 1], device=dev)
    dp, lp, cp = run(img[perm].contiguous())
    assert torch.equal(dp, d0[perm]) and torch.equal(lp, l0[perm]) and torch.equal(cp, c0[perm])
    # (3) linearity of the conv engine at full size is not observable through ReLU/NMS, so pin the dense maps:
    # image 0 through the same weights in torch fp32 on the GPU
    run(img)                                         # the plan's output maps now hold the un-permuted batch again
    plan = det.fused_plan(img.shape, dev, False)
    mine = [[t[:1].clone() for t in plan.outs[k]] for k in range(3)]
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        with torch.no_grad():
            ref = om.detector_forward({k: v.to(dev) for k, v in sd.items()}, img[:1])
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    for name, a_l, b_l in zip(("cls", "reg", "iou"), mine, ref):
        for lvl, (a, b) in enumerate(zip(a_l, b_l)):
            err = (a - b).abs().max().item()
            scale = max(b.abs().max().item(), 1.0)
            assert err <= 2e-4 * scale, "full-size head %s level %d: max|d| %g of range %g" % (name, lvl, err, scale)


def test_preprocess_resize_kernel_bit_exact_vs_cv2_restatement():
    """iou_preprocess_resize_u8 (resize + normalise + flip + pad + CHW in one kernel) == the oracle's restatement of
    ImageTransform.__call__ with mmcv.imrescale -> cv2.resize(INTER_LINEAR), bit for bit, incl. the cv2 goldens."""
    from oracle import preprocess as OP
    from gen_golden_fixtures import resize_cases
    mean, std = [123.675, 116.28, 103.53], [58.395, 57.12, 57.375]
    g = np.load(os.path.join(U.GOLD, "resize_cv2.npz"))
    for name, (img, (dw, dh)) in resize_cases().items():        # keep_ratio=False: exact (w, h) as in the goldens
        tf = P.ImageTransform(mean, std, to_rgb=True, size_divisor=32, scale=(dw, dh), keep_ratio=False)
        out, img_shape, pad_shape, factor = tf(torch.from_numpy(img[None]).to(DEV))
        want, w_shape, w_pad = OP.image_transform(g[name], mean, std, True, 32, False)
        assert img_shape == (dh, dw, 3) and tuple(pad_shape) == tuple(w_pad)
        assert np.array_equal(out[0].cpu().numpy().view(np.uint32), want.view(np.uint32)), name
    rs = np.random.RandomState(9)
    frames = rs.randint(0, 256, (3, 120, 160, 3)).astype(np.uint8)            # a batch, keep_ratio, flip
    tf = P.ImageTransform(mean, std, to_rgb=True, size_divisor=32, scale=(333, 200))
    for flip in (False, True):
        out, img_shape, pad_shape, factor = tf(torch.from_numpy(frames).to(DEV), flip=flip)
        for i in range(3):
            want, w_shape, w_pad, w_f = OP.image_transform_rescaled(frames[i], (333, 200), mean, std, True, 32, flip)
            assert tuple(img_shape) == tuple(w_shape) and tuple(pad_shape) == tuple(w_pad) and factor == w_f
            assert np.array_equal(out[i].cpu().numpy().view(np.uint32), want.view(np.uint32)), (i, flip)
    assert tf.pad_shape(120, 160) == (224, 288) and tf.out_shape(120, 160)[:2] == (200, 267)


def _spread_fcos_head(seed=5):
    head = P.IoUawareFCOSHead(81, 256, strides=cases.FCOS_STRIDES)
    head.init_weights()
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for k, v in head.state_dict().items():
            if ".gn.weight" in k:
                v.copy_(torch.rand(v.shape, generator=g) + 0.5)
            elif ".gn.bias" in k:
                v.copy_(torch.randn(v.shape, generator=g) * 0.1)
            elif k.endswith(".scale"):
                v.copy_(torch.rand(v.shape, generator=g) + 0.5)
            elif k.endswith("conv.weight"):
                v.copy_(torch.randn(v.shape, generator=g) * 0.03)
    return head.eval(), g


def test_group_norm_kernel_vs_torch():
    """iou_group_norm_relu on a two-segment padded-rows map vs torch.nn.functional.group_norm (+ReLU)."""
    import torch.nn.functional as F
    from iou_aware_single_stage_object_detector_b200 import engine as E
    g = torch.Generator().manual_seed(1)
    xs = [torch.randn(2, 256, 13, 17, generator=g) * 3 + 1, torch.randn(2, 256, 5, 7, generator=g)]
    gamma, beta = torch.rand(256, generator=g) + 0.5, torch.randn(256, generator=g) * 0.2
    for groups, relu in ((32, True), (8, False)):
        eng = E.Engine(DEV)
        Fm = eng.new_map([(2, 13, 17), (2, 5, 7)], 256)
        from iou_aware_single_stage_object_detector_b200 import lib as L
        for s_, x in enumerate(xs):
            xd = x.to(DEV).contiguous()
            eng.keep.append(xd)
            L.check(eng.lib.iou_pack_nchw(xd.data_ptr(), 2, 256, x.shape[2], x.shape[3], Fm.ptr, Fm.segs[s_][0],
                                          L.stream_ptr()))
        eng.group_norm("gn", Fm, gamma, beta, groups, 1e-5, relu=relu)
        outs = [eng.unpack_output(Fm, s_) for s_ in range(2)]
        eng.run()
        torch.cuda.synchronize()
        for x, y in zip(xs, outs):
            ref = F.group_norm(x, groups, gamma, beta, 1e-5)
            ref = F.relu(ref) if relu else ref
            # inputs/outputs live as bf16 hi+lo pairs (16 mantissa bits): 2e-5 of the range
            assert (y.cpu() - ref).abs().max().item() <= 2e-5 * max(ref.abs().max().item(), 1.0)


def test_fcos_head_forward_vs_oracle_and_get_bboxes():
    """IoUawareFCOSHead.forward on the conv engine (GN towers, split output GEMMs, exp(scale * reg)) vs the oracle
    restatement that equals the reference module bit for bit; then forward -> get_bboxes end to end vs the oracle."""
    head, g = _spread_fcos_head()
    sd = {"bbox_head." + k: v.clone() for k, v in head.state_dict().items()}
    sizes = [(16, 20), (8, 10), (4, 5), (2, 3), (1, 2)]
    feats = [torch.randn(2, 256, h, w, generator=g) for (h, w) in sizes]
    ref = om.fcos_head_forward(sd, feats)
    head = head.to(DEV)
    outs = head([f.to(DEV) for f in feats])
    torch.cuda.synchronize()
    for name, a_l, b_l in zip(("cls", "bbox_pred", "centerness", "iou"), outs, ref):
        for lvl, (a, b) in enumerate(zip(a_l, b_l)):
            assert tuple(a.shape) == tuple(b.shape), (name, lvl, a.shape, b.shape)
            err = (a.cpu() - b).abs().max().item()
            assert err <= 2e-4 * max(b.abs().max().item(), 1.0), (name, lvl, err)
    # end to end through the reference signature
    cfg = P.ConfigDict(dict(cases.TEST_CFG, nms_pre=300))
    metas = [dict(ori_shape=(128, 157, 3), img_shape=(128, 157, 3), pad_shape=(128, 160, 3), scale_factor=1.0,
                  flip=False)] * 2
    res = head.get_bboxes(outs[0], outs[1], outs[2], outs[3], [None] * 2, [None] * 2, metas, cfg, rescale=False)
    for i, (d, l) in enumerate(res):
        want_d, want_l = op.fcos_get_bboxes_single([t[i] for t in ref[0]], [t[i] for t in ref[1]], [t[i] for t in ref[3]],
                                                   cases.FCOS_STRIDES, metas[i]["img_shape"], 1.0, dict(cfg))
        assert abs(d.shape[0] - want_d.shape[0]) <= 2
        if want_d.shape[0]:
            frac, ms, mb = U.match_as_sets(d.cpu().numpy(), l.cpu().numpy(), want_d.numpy(), want_l.numpy(),
                                           min_frac=0.97, score_tol=1e-4, box_tol=1e-4 * 160)


def test_fcos_detector_end_to_end_small():
    """configs/fcos/iou_aware_fcos_r50_caffe_fpn_gn_1x_4gpu.py through build_detector: caffe-style ResNet (stride in
    the 1x1), FPN with P6 from P5 and ReLU before P7, GroupNorm head, distance decode, alpha 0.3 -- head maps and
    detections vs the oracle (which equals the reference FCOS detector bit for bit)."""
    cfg = P.Config.fromfile(os.path.join(U.ROOT, "configs", "fcos", "iou_aware_fcos_r50_caffe_fpn_gn_1x_4gpu.py"))
    cfg.model.pretrained = None
    torch.manual_seed(0)
    det = P.build_detector(cfg.model, train_cfg=None, test_cfg=cfg.test_cfg).eval()
    assert type(det).__name__ == "FCOS"
    sd = det.state_dict()
    om.spread_fcos_weights_(sd, seed=1)
    det.load_state_dict(sd)
    sd = {k: v.clone() for k, v in det.state_dict().items()}
    det = det.to(DEV)
    n, h, w = 2, 128, 160
    img = torch.randn(n, 3, h, w, generator=torch.Generator().manual_seed(3))
    metas = [dict(ori_shape=(h, w - 3, 3), img_shape=(h, w - 3, 3), pad_shape=(h, w, 3), scale_factor=1.0,
                  flip=False) for _ in range(n)]
    results = det.simple_test_batch(img.to(DEV), metas, rescale=False)
    results = det.simple_test_batch(img.to(DEV), metas, rescale=False)        # second call replays the CUDA graph
    plan = det.fused_plan(img.shape, torch.device(DEV), False)
    torch.cuda.synchronize()
    ref = om.fcos_detector_forward(sd, img)
    for name, mine, r in zip(("cls", "bbox_pred", "centerness", "iou"), plan.outs, ref):
        for lvl, (a, b) in enumerate(zip(mine, r)):
            err = (a.cpu() - b).abs().max().item()
            assert err <= 3e-4 * max(b.abs().max().item(), 1.0), (name, lvl, err, b.abs().max().item())
    for i in range(n):
        d_ref, l_ref = op.fcos_get_bboxes_single([c[i] for c in ref[0]], [r_[i] for r_ in ref[1]],
                                                 [q[i] for q in ref[3]], cases.FCOS_STRIDES, metas[i]["img_shape"],
                                                 1.0, dict(cfg.test_cfg), rescale=False)
        d_my, l_my = U.results_to_arrays(results[i])
        assert abs(len(d_my) - d_ref.shape[0]) <= 3 and d_ref.shape[0] > 0
        U.match_as_sets(d_my, l_my, d_ref.numpy(), l_ref.numpy(), min_frac=0.97, score_tol=1e-4,
                        box_tol=1e-4 * max(h, w))


def test_prepare_test_img_item_layout_vs_reference_golden():
    """prepare_test_img == the reference's CustomDataset.prepare_test_img (custom.py:283-359) on the same frames:
    entry order (scale-major, flipped twin second), the fork's gt_bboxes / gt_labels lists, img_meta fields, and every
    image bit for bit (rescale + normalise + flip + pad + CHW in one device kernel)."""
    from gen_golden_fixtures import test_item_cases, IMG_NORM
    gold = np.load(os.path.join(U.GOLD, "test_items.npz"))
    tf = P.ImageTransform(size_divisor=32, **IMG_NORM)
    for name, c in test_item_cases().items():
        item = P.prepare_test_img(c["frame"], c["img_info"], c["ann"], tf, P.BboxTransform(), c["img_scales"],
                                  c["flip_ratio"], c["resize_keep_ratio"], device=DEV)
        assert sorted(item.keys()) == ["gt_bboxes", "gt_labels", "img", "img_meta"]
        assert len(item["img"]) == len(item["img_meta"]) == int(gold[name + "_n_img"])
        assert len(item["gt_bboxes"]) == len(item["gt_labels"]) == int(gold[name + "_n_gt"])
        for i, (im, meta) in enumerate(zip(item["img"], item["img_meta"])):
            assert meta.cpu_only and not meta.stack
            m = meta.data
            assert np.array_equal(im.cpu().numpy(), gold["%s_img_%d" % (name, i)]), (name, i)
            want = gold["%s_meta_%d" % (name, i)]
            assert list(m["ori_shape"]) + list(m["img_shape"]) + list(m["pad_shape"]) + [int(m["flip"])] == want.tolist()
            assert np.array_equal(np.asarray(m["scale_factor"], dtype=np.float64).reshape(-1), gold["%s_sf_%d" % (name, i)])
        for i, (gb, gl) in enumerate(zip(item["gt_bboxes"], item["gt_labels"])):
            assert np.array_equal(gb.data.numpy(), gold["%s_gtb_%d" % (name, i)])
            assert np.array_equal(gl.data.numpy(), gold["%s_gtl_%d" % (name, i)])
    # the item feeds forward_test unchanged (base.py:62-103): one image per call, gt lists forwarded
    det, cfg = U.small_detector()
    det = det.to(DEV)
    c = test_item_cases()["single_scale"]
    item = P.prepare_test_img(c["frame"], c["img_info"], c["ann"], tf, None, c["img_scales"], 0, True, device=DEV)
    res = det(return_loss=False, rescale=True, img=[t.unsqueeze(0) for t in item["img"]],
              img_meta=[[m.data] for m in item["img_meta"]], gt_bboxes=[[g.data] for g in item["gt_bboxes"]],
              gt_labels=[[g.data] for g in item["gt_labels"]])
    assert len(res) == 80 and all(r.shape[1] == 5 for r in res)
