"""Pins the CPU oracle (oracle/) to outputs of the REAL reference (tests/golden/*.npz,
made by tests/golden/gen_golden.py from /root/reference).  CPU only."""
import os

import numpy as np
import pytest
import torch

import cases
from oracle import postproc as op

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _bases():
    sc = op.retina_anchor_scales(4, 3)
    return [op.base_anchors(s, sc, [0.5, 1.0, 2.0]) for s in cases.STRIDES]


def test_anchor_known_answers():
    kat = np.load(os.path.join(G, "kat_anchor_codec.npz"))
    bases = _bases()
    for s, b in zip(cases.STRIDES, bases):
        assert np.array_equal(b.numpy(), kat["base_anchors_%d" % s])
    # SURVEY.md 8(a7) probe values
    assert bases[0].tolist()[0] == [-19, -7, 26, 14] and bases[0].tolist()[8] == [-14, -32, 21, 39]
    assert bases[4].tolist()[0] == [-298, -117, 425, 244] and bases[4].tolist()[8] == [-223, -511, 350, 638]
    assert np.array_equal(op.grid_anchors(bases[0], 3, 5, 8).numpy(), kat["grid_s8_3x5"])


def test_delta2bbox_known_answers():
    kat = np.load(os.path.join(G, "kat_anchor_codec.npz"))
    rois, deltas = cases.codec_inputs()
    r, d = torch.from_numpy(rois), torch.from_numpy(deltas)
    out = op.delta2bbox(r, d, max_shape=(800, 1333, 3)).numpy()
    assert np.array_equal(out, kat["delta2bbox"])
    assert np.allclose(out[0], [0, 0, 39.7977, 799], atol=1e-3)
    out2 = op.delta2bbox(r, d, [0.1, 0., -0.1, 0.], [0.1, 0.1, 0.2, 0.2], (300, 400, 3)).numpy()
    assert np.array_equal(out2, kat["delta2bbox_std"])


@pytest.mark.parametrize("name", list(cases.nms_inputs().keys()))
def test_nms_matches_reference_nms_cpu(name):
    gold = np.load(os.path.join(G, "nms_keep.npz"))[name]
    dets = cases.nms_inputs()[name]
    # no IoU == thr pair exists in these inputs (asserted at generation), so the
    # ">" (nms_cuda) and ">=" (nms_cpu) variants must both equal the reference result
    for mode in ("cuda", "cpu"):
        keep = op.nms(dets, 0.5, mode).numpy()
        assert np.array_equal(keep, gold), (name, mode)


def test_nms_threshold_semantics():
    # two boxes with IoU exactly 0.5: 10x10 vs 10x20 sharing the 10x10 -> inter 100, union 200
    d = np.array([[0, 0, 9, 9, 0.9], [0, 0, 9, 19, 0.8]], dtype=np.float32)
    assert op.iou_pair(d[0, :4], d[1, :4]) == 0.5
    assert op.nms(d, 0.5, "cuda").tolist() == [0, 1]      # nms_kernel.cu:60  (>)
    assert op.nms(d, 0.5, "cpu").tolist() == [0]          # nms_cpu.cpp:55    (>=)
    assert op.nms(np.zeros((0, 5), np.float32), 0.5).numel() == 0


@pytest.mark.parametrize("name", cases.POSTPROC_CASES)
def test_get_bboxes_matches_reference(name):
    gold = np.load(os.path.join(G, "postproc_%s.npz" % name))
    case = cases.postproc_case(name)
    bases = _bases()
    n_img = case["cls"][0].shape[0]
    for i in range(n_img):
        m = case["img_metas"][i]
        args = ([c[i] for c in case["cls"]], [r[i] for r in case["reg"]], [q[i] for q in case["iou"]],
                cases.STRIDES, bases, m["img_shape"], m["scale_factor"])
        boxes, scores, idx = op.candidates_single(*args, nms_pre=case["cfg"]["nms_pre"],
                                                  rescale=case["rescale"])
        assert np.array_equal(idx.numpy(), gold["cand_idx_%d" % i])
        assert np.array_equal(boxes.numpy(), gold["cand_boxes_%d" % i])
        assert np.array_equal(scores.numpy(), gold["cand_scores_%d" % i])
        dets, labels = op.get_bboxes_single(*args, cfg=case["cfg"], rescale=case["rescale"],
                                            nms_mode="cpu")
        assert dets.shape == gold["dets_%d" % i].shape
        assert np.array_equal(labels.numpy(), gold["labels_%d" % i])
        assert np.array_equal(dets.numpy(), gold["dets_%d" % i])
        # the ">" variant (what the CUDA library implements) agrees on these inputs too
        d2, l2 = op.get_bboxes_single(*args, cfg=case["cfg"], rescale=case["rescale"], nms_mode="cuda")
        assert np.array_equal(d2.numpy(), dets.numpy()) and np.array_equal(l2.numpy(), labels.numpy())


def test_focal_loss_matches_python_formula():
    """Cross-check against the reference's pure-torch statement of the same loss
    (mmdet/core/loss/losses.py:226-247 py_sigmoid_focal_loss, one-hot targets)."""
    torch.manual_seed(0)
    x = torch.randn(37, 80) * 3
    t = torch.randint(0, 81, (37,))
    g, a = 2.0, 0.25
    loss = op.sigmoid_focal_loss_forward(x, t, g, a)
    onehot = torch.zeros(37, 81).scatter_(1, t[:, None], 1.0)[:, 1:]
    p = x.sigmoid()
    pt = (1 - p) * onehot + p * (1 - onehot)
    w = (a * onehot + (1 - a) * (1 - onehot)) * pt.pow(g)
    ref = torch.nn.functional.binary_cross_entropy_with_logits(x, onehot, reduction="none") * w
    assert torch.allclose(loss, ref, rtol=1e-4, atol=1e-6)
    xg = x.clone().requires_grad_(True)
    p2 = xg.sigmoid()
    pt2 = (1 - p2) * onehot + p2 * (1 - onehot)
    w2 = (a * onehot + (1 - a) * (1 - onehot)) * pt2.pow(g)
    (torch.nn.functional.binary_cross_entropy_with_logits(xg, onehot, reduction="none") * w2).sum().backward()
    grad = op.sigmoid_focal_loss_backward(x, t, torch.ones_like(x), g, a)
    assert torch.allclose(grad, xg.grad, rtol=1e-3, atol=1e-6)


def test_plain_retina_head_matches_reference():
    """Sibling RetinaHead (score = sigmoid(cls), anchor_head.py:364-450) on the 'small' maps."""
    gold = np.load(os.path.join(G, "postproc_plain_retina_small.npz"))
    case = cases.postproc_case("small")
    bases = _bases()
    for i in range(case["cls"][0].shape[0]):
        m = case["img_metas"][i]
        d, l = op.get_bboxes_single([c[i] for c in case["cls"]], [r[i] for r in case["reg"]], None,
                                    cases.STRIDES, bases, m["img_shape"], m["scale_factor"], cfg=case["cfg"],
                                    rescale=case["rescale"], nms_mode="cpu")
        assert np.array_equal(l.numpy(), gold["labels_%d" % i]) and np.array_equal(d.numpy(), gold["dets_%d" % i])


# ------------------------------------------------------------------ Soft-NMS
@pytest.mark.parametrize("name", list(cases.soft_nms_inputs().keys()))
def test_soft_nms_oracle_matches_reference_pyx(name):
    """oracle/soft_nms.py == the reference's own soft_nms_cpu.pyx (compiled with Cython) bit for bit:
    decayed scores, survivors and their order."""
    from oracle import soft_nms as SN
    g = np.load(os.path.join(G, "soft_nms.npz"))
    d, thr, method, sigma, min_score = cases.soft_nms_inputs()[name]
    nd, inds = SN.soft_nms(d, thr, method=method, sigma=sigma, min_score=min_score)
    assert nd.dtype == np.float32 and inds.dtype == np.int64
    assert np.array_equal(inds, g[name + "_inds"])
    assert np.array_equal(nd.view(np.uint32), g[name + "_dets"].view(np.uint32))


def test_soft_nms_oracle_multiclass_matches_reference():
    """multiclass_nms with nms_cfg type 'soft_nms' (bbox_nms.py:29-67) on the 'small' case candidates."""
    g = np.load(os.path.join(G, "soft_nms.npz"))
    c = np.load(os.path.join(G, "postproc_small.npz"))
    cfg = dict(cases.SOFT_MULTICLASS)
    for i in range(2):
        boxes = torch.from_numpy(c["cand_boxes_%d" % i])
        scores = torch.from_numpy(c["cand_scores_%d" % i])
        padded = torch.cat([scores.new_zeros(scores.shape[0], 1), scores], dim=1)
        d, l = op.multiclass_nms(boxes, padded, 0.05, cfg['iou_thr'], 100,
                                 soft=dict(method=cfg['method'], sigma=cfg['sigma'], min_score=cfg['min_score']))
        assert np.array_equal(l.numpy(), g["mc_labels_%d" % i])
        assert np.array_equal(d.numpy().view(np.uint32), g["mc_dets_%d" % i].view(np.uint32))


def test_soft_nms_oracle_rejects_unknown_method():
    from oracle import soft_nms as SN
    with pytest.raises(ValueError):
        SN.soft_nms(np.zeros((1, 5), np.float32), 0.5, method='bogus')


@pytest.mark.reference
def test_soft_nms_oracle_vs_compiled_reference_random():
    """Build-container only: 200 random scenes through oracle/_ref/soft_nms_cpu*.so and the oracle."""
    from oracle import build_ref, soft_nms as SN
    ref = build_ref.load_ref_soft_nms_cpu()
    rs = np.random.RandomState(11)
    for trial in range(200):
        n = int(rs.randint(1, 250))
        d = cases.random_dets(rs, n, float(rs.choice([50., 200., 600.])), float(rs.choice([30., 120.])))
        if trial % 5 == 0:
            d[:, 4] = np.round(d[:, 4] * 8) / 8
        method = int(rs.randint(1, 4))
        thr, sig, ms = float(rs.choice([0.3, 0.5, 0.7])), float(rs.choice([0.3, 0.5, 1.0])), float(rs.choice([1e-3, 0.05, 0.2]))
        rb, ri = ref.soft_nms_cpu(d, thr, method=method, sigma=sig, min_score=ms)
        ob, oi = SN.soft_nms_cpu(d, thr, method, sig, ms)
        assert np.array_equal(ri, oi) and np.array_equal(rb.view(np.uint32), ob.view(np.uint32)), (trial, method, n)


def test_soft_nms_oracle_get_bboxes_matches_reference():
    """get_bboxes with test_cfg.nms type 'soft_nms' on the 'small' case vs the reference's own output."""
    g = np.load(os.path.join(G, "soft_nms.npz"))
    case = cases.postproc_case("small")
    cfgd = dict(case["cfg"])
    cfgd["nms"] = dict(cases.SOFT_MULTICLASS)
    strides, bases = [8, 16, 32, 64, 128], _bases()
    for i, meta in enumerate(case["img_metas"]):
        d, l = op.get_bboxes_single([t[i] for t in case["cls"]], [t[i] for t in case["reg"]],
                                    [t[i] for t in case["iou"]], strides, bases, meta["img_shape"],
                                    meta["scale_factor"], cfgd, rescale=case["rescale"], nms_mode="cpu")
        assert np.array_equal(l.numpy(), g["gb_labels_%d" % i])
        assert np.array_equal(d.numpy().view(np.uint32), g["gb_dets_%d" % i].view(np.uint32))


# ------------------------------------------------------------------ IoUawareFCOSHead.get_bboxes
def test_fcos_get_bboxes_oracle_matches_reference():
    g = np.load(os.path.join(G, "postproc_fcos.npz"))
    case = cases.fcos_case()
    for i, meta in enumerate(case["img_metas"]):
        args = ([t[i] for t in case["cls"]], [t[i] for t in case["reg"]], [t[i] for t in case["iou"]],
                cases.FCOS_STRIDES, meta["img_shape"], meta["scale_factor"], case["cfg"])
        b, s, idx = op.fcos_get_bboxes_single(*args, rescale=True, return_candidates=True)
        assert np.array_equal(idx.numpy(), g["cand_idx_%d" % i])
        assert np.array_equal(b.numpy().view(np.uint32), g["cand_boxes_%d" % i].view(np.uint32))
        assert np.array_equal(s.numpy().view(np.uint32), g["cand_scores_%d" % i].view(np.uint32))
        d, l = op.fcos_get_bboxes_single(*args, rescale=True, nms_mode="cpu")
        assert np.array_equal(l.numpy(), g["labels_%d" % i])
        assert np.array_equal(d.numpy().view(np.uint32), g["dets_%d" % i].view(np.uint32))
    # known answer for the points (iou_aware_fcos_head.py:392-401)
    assert op.fcos_points(2, 3, 8).tolist() == [[4, 4], [12, 4], [20, 4], [4, 12], [12, 12], [20, 12]]


# ------------------------------------------------------------------ ImageTransform resize
def test_resize_oracle_matches_cv2_golden():
    """oracle.preprocess.resize_linear_u8 == cv2.resize(INTER_LINEAR) on uint8 frames, bit for bit (goldens written by
    cv2 itself in the build container; when cv2 is importable, 30 random sizes are checked live as well)."""
    from oracle import preprocess as OP
    from gen_golden_fixtures import resize_cases
    g = np.load(os.path.join(G, "resize_cv2.npz"))
    for name, (img, (dw, dh)) in resize_cases().items():
        assert np.array_equal(OP.resize_linear_u8(img, dw, dh), g[name]), name
    try:
        import cv2
    except ImportError:
        cv2 = None
    if cv2 is not None:
        rs = np.random.RandomState(5)
        for _ in range(30):
            sh, sw, dh, dw = (int(v) for v in rs.randint(8, 400, size=4))
            img = rs.randint(0, 256, (sh, sw, 3)).astype(np.uint8)
            assert np.array_equal(OP.resize_linear_u8(img, dw, dh), cv2.resize(img, (dw, dh), interpolation=cv2.INTER_LINEAR))
    # mmcv 0.2.8 imrescale size rule: COCO 480x640 -> 800x1067, 427x640 -> 800x1199
    assert OP.rescale_size(480, 640, (1333, 800))[:2] == (800, 1067)
    assert OP.rescale_size(427, 640, (1333, 800))[:2] == (800, 1199)
    assert OP.rescale_size(640, 427, (1333, 800))[:2] == (1199, 800)
    assert OP.rescale_size(100, 200, 0.5)[:2] == (50, 100)
    nh, nw, f = OP.rescale_size(100, 200, (300, 150), keep_ratio=False)
    assert (nh, nw) == (150, 300) and np.allclose(f, [1.5, 1.5, 1.5, 1.5])


# ------------------------------------------------------------------ test items
def test_image_transform_oracle_and_bbox_transform_vs_reference_items():
    """tests/golden/test_items.npz was written by the reference's own ImageTransform / BboxTransform /
    CustomDataset.prepare_test_img (mmcv 0.2.8's image functions restated in the shim on cv2 / numpy): the oracle's
    image_transform_rescaled reproduces every image bit for bit (normalise / flip / pad are thereby pinned), and the
    host-side BboxTransform of this repo reproduces the gt boxes."""
    from oracle import preprocess as OP
    from gen_golden_fixtures import test_item_cases, IMG_NORM
    import iou_aware_single_stage_object_detector_b200 as P
    gold = np.load(os.path.join(G, "test_items.npz"))
    for name, c in test_item_cases().items():
        i = 0
        boxes = c["ann"]["bboxes"]
        for s_i, scale in enumerate(c["img_scales"]):
            for flip in ([False, True] if c["flip_ratio"] > 0 else [False]):
                chw, img_shape, pad_shape, factor = OP.image_transform_rescaled(
                    c["frame"], scale, IMG_NORM["mean"], IMG_NORM["std"], IMG_NORM["to_rgb"], 32, flip,
                    c["resize_keep_ratio"])
                assert np.array_equal(chw, gold["%s_img_%d" % (name, i)]), (name, i)
                meta = gold["%s_meta_%d" % (name, i)]
                assert tuple(meta[3:6]) == tuple(img_shape) and tuple(meta[6:9]) == tuple(pad_shape) and bool(meta[9]) == flip
                assert np.allclose(np.asarray(factor, dtype=np.float64).reshape(-1), gold["%s_sf_%d" % (name, i)], rtol=0, atol=0)
                if not flip:
                    boxes = P.BboxTransform()(boxes, img_shape, factor, flip=False)
                    assert np.array_equal(boxes, gold["%s_gtb_%d" % (name, s_i)]), (name, s_i)
                i += 1
        assert i == int(gold[name + "_n_img"]) and len(c["img_scales"]) == int(gold[name + "_n_gt"])
