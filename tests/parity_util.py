"""Shared parity checkers (used by the -m gpu tests and by __graft_entry__.smoke()).

Everything here compares the CUDA path (called through the package == through the C ABI) with the
CPU oracle and with the golden fixtures generated from the real reference.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests", "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)

import cases  # noqa: E402
from oracle import model as om  # noqa: E402
from oracle import postproc as op  # noqa: E402

import iou_aware_single_stage_object_detector_b200 as P  # noqa: E402
from iou_aware_single_stage_object_detector_b200 import postproc as PP  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
CFG_DIR = os.path.join(ROOT, "configs", "iou_aware_single_stage_detector")
# parity bar of BASELINE.json north_star: scores / boxes within 1e-4 (fp32), indices exact
RTOL = ATOL = 1e-4


def head_cfg():
    return dict(type='IoUawareRetinaHead', num_classes=81, in_channels=256, stacked_convs=4,
                feat_channels=256, octave_base_scale=4, scales_per_octave=3,
                anchor_ratios=[0.5, 1.0, 2.0], anchor_strides=[8, 16, 32, 64, 128],
                target_means=[.0, .0, .0, .0], target_stds=[1.0, 1.0, 1.0, 1.0],
                loss_cls=dict(type='FocalLoss', use_sigmoid=True, gamma=2.0, alpha=0.25, loss_weight=1.0),
                loss_bbox=dict(type='SmoothL1Loss', beta=0.11, loss_weight=1.0))


_head = None


def get_head():
    global _head
    if _head is None:
        torch.manual_seed(0)
        _head = P.build_head(head_cfg())
    return _head


def oracle_bases():
    sc = op.retina_anchor_scales(4, 3)
    return [op.base_anchors(s, sc, [0.5, 1.0, 2.0]) for s in cases.STRIDES]


def near_tie_explained(scores_sorted_desc, tol=2e-6):
    """True if some adjacent pair of the sorted key list is closer than tol (order then undefined)."""
    d = np.diff(scores_sorted_desc.astype(np.float64))
    return bool(np.any(np.abs(d) < tol))


def check_postproc_case(name, verbose=True):
    """get_bboxes on the GPU vs golden fixtures made from the real reference."""
    gold = np.load(os.path.join(GOLD, "postproc_%s.npz" % name))
    case = cases.postproc_case(name)
    head = get_head()
    dev = torch.device("cuda:0")
    cfg = P.ConfigDict(case["cfg"])
    cls = [t.to(dev) for t in case["cls"]]
    reg = [t.to(dev) for t in case["reg"]]
    iou = [t.to(dev) for t in case["iou"]]
    n_img = cls[0].shape[0]
    metas = case["img_metas"]
    sizes = [tuple(t.shape[-2:]) for t in cls]
    wsp = head.postproc_workspace(sizes, n_img, cfg, dev)
    info = PP.make_img_info(metas, dev)
    # ---- stage 1: candidates
    boxes, scores_cm, idx = PP.decode_candidates(wsp, cls, reg, iou, info, case["rescale"])
    torch.cuda.synchronize()
    idx_exact = True
    for i in range(n_img):
        g_idx = gold["cand_idx_%d" % i]
        g_box = gold["cand_boxes_%d" % i]
        g_sc = gold["cand_scores_%d" % i]
        my_idx = idx[i].cpu().numpy()
        my_box = boxes[i].cpu().numpy()
        my_sc = scores_cm[i].t().contiguous().cpu().numpy()
        same = np.array_equal(my_idx, g_idx)
        if not same:
            idx_exact = False
            # order may only differ where two keys are closer than fp32 noise
            assert sorted(my_idx.tolist()) == sorted(g_idx.tolist()) or \
                near_tie_explained(np.sort(g_sc.max(1))[::-1]), "candidate set differs (%s, img %d)" % (name, i)
            bad = int((my_idx != g_idx).sum())
            assert bad <= 8, "%d candidate order mismatches (%s, img %d)" % (bad, name, i)
            if verbose:
                print("[%s] img %d: %d candidate order differences (near-ties)" % (name, i, bad))
        else:
            assert np.allclose(my_box, g_box, rtol=RTOL, atol=ATOL), \
                "boxes differ: max %g" % np.abs(my_box - g_box).max()
            assert np.allclose(my_sc, g_sc, rtol=RTOL, atol=ATOL), \
                "scores differ: max %g" % np.abs(my_sc - g_sc).max()
            if verbose:
                print("[%s] img %d: candidates exact order; max |dbox| %.3g, max |dscore| %.3g" % (
                    name, i, np.abs(my_box - g_box).max(), np.abs(my_sc - g_sc).max()))
    # ---- stage 2 on the GOLDEN candidates: bit-exact indices/labels/values
    M = wsp.M
    gb = torch.stack([torch.from_numpy(gold["cand_boxes_%d" % i]) for i in range(n_img)]).to(dev)
    gs = torch.stack([torch.from_numpy(gold["cand_scores_%d" % i]).t().contiguous() for i in range(n_img)]).to(dev)
    assert gb.shape[1] == M
    dets, labels, counts = PP.batched_nms(wsp, gb, gs)
    for i, (d, l) in enumerate(PP.split_results(dets, labels, counts)):
        gd, gl = gold["dets_%d" % i], gold["labels_%d" % i]
        assert d.shape[0] == gd.shape[0], "det count %d != %d (%s img %d)" % (d.shape[0], gd.shape[0], name, i)
        assert np.array_equal(l.cpu().numpy(), gl), "labels differ (%s img %d)" % (name, i)
        assert np.array_equal(d.cpu().numpy(), gd), "dets differ (%s img %d)" % (name, i)
    # ---- end to end through the reference-signature API
    res = head.get_bboxes(cls, reg, iou, [None] * n_img, [None] * n_img, metas, cfg, rescale=case["rescale"])
    for i, (d, l) in enumerate(res):
        gd, gl = gold["dets_%d" % i], gold["labels_%d" % i]
        assert d.shape[0] == gd.shape[0]
        if idx_exact:
            assert np.array_equal(l.cpu().numpy(), gl), "e2e labels differ (%s img %d)" % (name, i)
            assert np.allclose(d.cpu().numpy(), gd, rtol=RTOL, atol=ATOL), \
                "e2e dets differ: %g" % np.abs(d.cpu().numpy() - gd).max()
        else:
            match_as_sets(d.cpu().numpy(), l.cpu().numpy(), gd, gl)
    return True


def match_as_sets(d, l, gd, gl, min_frac=0.97, score_tol=ATOL, box_tol=None):
    """Order-free comparison keyed by label + nearest box.  A detection matches
    when |dscore| <= score_tol and every coordinate is within box_tol pixels (default: allclose with
    rtol = atol = 1e-4).  Returns (matched fraction, max |dscore|, max |dbox| px) over the matches."""
    used = np.zeros(len(gd), bool)
    hit, ms, mb = 0, 0.0, 0.0
    for k in range(len(d)):
        cand = np.where((gl == l[k]) & ~used)[0]
        if cand.size == 0:
            continue
        err = np.abs(gd[cand, :4] - d[k, :4]).max(1)
        j = cand[err.argmin()]
        ok_box = err.min() <= box_tol if box_tol is not None else \
            np.allclose(d[k, :4], gd[j, :4], rtol=RTOL, atol=ATOL)
        if ok_box and abs(d[k, 4] - gd[j, 4]) <= score_tol:
            used[j] = True
            hit += 1
            ms, mb = max(ms, float(abs(d[k, 4] - gd[j, 4]))), max(mb, float(err.min()))
    frac = hit / max(len(gd), 1)
    assert frac >= min_frac, "only %.3f of detections matched" % frac
    return frac, ms, mb


def small_detector(seed=0, spread=True, cfg_name="iou_aware_retinanet_r50_fpn_1x_4gpu.py"):
    cfg = P.Config.fromfile(os.path.join(CFG_DIR, cfg_name))
    cfg.model.pretrained = None
    torch.manual_seed(seed)
    det = P.build_detector(cfg.model, train_cfg=None, test_cfg=cfg.test_cfg)
    det.eval()
    if spread:
        sd = det.state_dict()
        om.spread_weights_(sd, seed=seed + 1)
        det.load_state_dict(sd)
    return det, cfg


def results_to_arrays(per_class):
    n = sum(len(m) for m in per_class)
    if n == 0:
        return np.zeros((0, 5), np.float32), np.zeros((0,), np.int64)
    d = np.concatenate(list(per_class), 0)
    l = np.concatenate([np.full(len(m), c) for c, m in enumerate(per_class)]).astype(np.int64)
    return d, l


def check_detector_small(h=128, w=160, n=2, verbose=True, use_graph=False, passes=None):
    """Whole path on a small image: CUDA head maps vs oracle (torch-CPU fp32) maps, then detections.
    passes=None runs the detector's DEFAULT scheme (fp16 + e4m3, the one bench.py times)."""
    det, cfg = small_detector()
    sd = {k: v.clone() for k, v in det.state_dict().items()}
    dev = torch.device("cuda:0")
    det = det.to(dev)
    g = torch.Generator().manual_seed(3)
    img = torch.randn(n, 3, h, w, generator=g)
    metas = [dict(ori_shape=(h, w - 3, 3), img_shape=(h, w - 3, 3), pad_shape=(h, w, 3), scale_factor=1.0,
                  flip=False) for _ in range(n)]
    det.use_cuda_graph = use_graph
    det.passes = passes
    results = det.simple_test_batch(img.to(dev), metas, rescale=False)
    if use_graph:   # second call replays the captured graph
        results = det.simple_test_batch(img.to(dev), metas, rescale=False)
    plan = det.fused_plan(img.shape, dev, False)
    torch.cuda.synchronize()
    ref_cls, ref_reg, ref_iou = om.detector_forward(sd, img)
    worst = 0.0
    for name, mine, ref in (("cls", plan.outs[0], ref_cls), ("reg", plan.outs[1], ref_reg),
                            ("iou", plan.outs[2], ref_iou)):
        for l, (a, b) in enumerate(zip(mine, ref)):
            a = a.cpu().contiguous()
            err = (a - b).abs().max().item()
            scale = b.abs().max().item()
            worst = max(worst, err / max(scale, 1e-6))
            if verbose:
                print("head %s level %d: max|d| %.3g (ref max %.3g)" % (name, l, err, scale))
            # budget of the 3-pass split-bf16 path over ~60 stacked convs: <= 2e-4 of the map's range
            assert err <= 2e-4 * max(scale, 1.0), "head %s level %d differs: %g" % (name, l, err)
    bases = oracle_bases()
    for i in range(n):
        d_ref, l_ref = op.get_bboxes_single([c[i] for c in ref_cls], [r[i] for r in ref_reg],
                                            [q[i] for q in ref_iou], cases.STRIDES, bases,
                                            metas[i]["img_shape"], 1.0, dict(cfg.test_cfg), rescale=False)
        d_my, l_my = results_to_arrays(results[i])
        if verbose:
            print("image %d: %d dets (oracle %d)" % (i, len(d_my), d_ref.shape[0]))
        assert len(d_my) == d_ref.shape[0], "detection count %d != %d" % (len(d_my), d_ref.shape[0])
        if d_ref.shape[0]:
            # bar: every detection matched 1:1, scores within 1e-4, boxes within 1e-4 of the coordinate range
            frac, ms, mb = match_as_sets(d_my, l_my, d_ref.numpy(), l_ref.numpy(), min_frac=1.0,
                                         score_tol=1e-4, box_tol=1e-4 * max(h, w))
            if verbose:
                print("image %d: matched %.3f, max|dscore| %.3g, max|dbox| %.3g px" % (i, frac, ms, mb))
    return worst


# ------------------------------------------------------------------------------------------------
# Whole detector vs goldens written by the LIVE shimmed reference (tests/golden/gen_golden_detector.py)
import detector_cases as DC  # noqa: E402

# Tolerances of the dense half, asserted by check_detector_golden (north star: scores / boxes within 1e-4, fp32):
SCORE_TOL = 1e-4            # final scores and candidate scores: allclose(rtol=1e-4, atol=1e-4)
# A decoded coordinate is px + pw*dx (transforms.py:66): an error e in a regression delta moves it by pw*e pixels and
# anchors are up to ~1150 px wide, so a per-coordinate allclose(rtol=1e-4, atol=1e-4) cannot hold between ANY two
# implementations that sum in a different order.  What the network computes are the deltas: the
# padded-rows activation format keeps 15-16 significant bits per element per layer (csrc/split_fmt.cuh), which puts
# the regression deltas within 1.5e-4 (rms 3e-5) of the reference at 800x1344 and the decoded boxes within
# 1e-4 of their own scale.  Asserted:
#   |dcoord| <= BOX_DELTA_TOL * (1 + |coord| + side),  side = max(box side, largest anchor side of the box's level)
# and the count of coordinates outside the strict allclose(1e-4, 1e-4) is reported next to it.
BOX_DELTA_TOL = 1e-4      # measured at 800x1344: 5.4e-5 (worst coordinate 0.04 px on a 700 px anchor)
LOGIT_ABS_TOL = 1.5e-3      # head logits (range ~ +-15): max |d| over the stored samples
LOGIT_RMS_TOL = 2.5e-4      # ... and their rms
ANCHOR_SIDE = [s * 4 * 2 ** (2.0 / 3) * 2 ** 0.5 for s in DC.STRIDES]       # ratio 0.5 / 2 anchors, largest scale


def _box_ok(b, g, level_side=0.0, tol=BOX_DELTA_TOL):
    size = max(g[2] - g[0], g[3] - g[1], level_side, 1.0)
    return bool(np.all(np.abs(b - g) <= tol * (1.0 + np.abs(g) + size)))


def check_detector_golden(name, passes=None, verbose=True, use_graph=False):
    """The DEFAULT detector scheme (passes=None -> fp16 + e4m3) on a golden case: head-map samples, candidates and
    final detections vs the live reference's outputs.  Detections are matched 1:1 by label + box; a golden detection
    may stay unmatched only when its score is within SCORE_TOL of the top-`max_per_img` cut or of score_thr, or when
    an NMS decision on it sits within 1e-3 of iou_thr (the order / decision is then inside the score tolerance)."""
    gold = np.load(os.path.join(GOLD, "detector_%s.npz" % name))
    det, sd, cfg = DC.case_state_dict(name, P, om, CFG_DIR)
    img, metas = DC.case_inputs(name)
    dev = torch.device("cuda:0")
    det = det.to(dev)
    det.use_cuda_graph = use_graph
    det.passes = passes
    n = img.shape[0]
    dets, labels, counts = det.detect_device(img.to(dev), metas, rescale=True)
    if use_graph:
        dets, labels, counts = det.detect_device(img.to(dev), metas, rescale=True)
    torch.cuda.synchronize()
    plan = det.fused_plan(img.shape, dev, True)
    report = {"case": name, "passes": det.resolved_passes()}
    # ---- head maps (fixed random sample of every map)
    sq, cnt, worst = 0.0, 0, 0.0
    for kind, maps in (("cls", plan.outs[0]), ("reg", plan.outs[1]), ("iou", plan.outs[2])):
        k_sq, k_cnt, k_worst = 0.0, 0, 0.0
        for lv, m in enumerate(maps):
            for i in range(n):
                flat = m[i].contiguous().reshape(-1)
                idx = torch.from_numpy(DC.sample_index(name, "%s%d" % (kind, i), lv, flat.numel())).to(dev)
                mine = flat[idx].cpu().numpy().astype(np.float64)
                ref = gold["%s_l%d_%d" % (kind, lv, i)].astype(np.float64)
                assert np.isfinite(mine).all(), "non-finite %s logits (level %d)" % (kind, lv)
                e = np.abs(mine - ref)
                k_sq += float((e ** 2).sum()); k_cnt += e.size; k_worst = max(k_worst, float(e.max()))
        report["logit_%s" % kind] = (k_worst, (k_sq / k_cnt) ** 0.5)
        sq += k_sq; cnt += k_cnt; worst = max(worst, k_worst)
        if verbose:
            print("[%s] head %s: max|d| %.3g rms %.3g over %d samples" % (name, kind, k_worst, (k_sq / k_cnt) ** 0.5, k_cnt))
    assert worst <= LOGIT_ABS_TOL, "head logits differ: max %g" % worst
    assert (sq / cnt) ** 0.5 <= LOGIT_RMS_TOL, "head logits differ: rms %g" % (sq / cnt) ** 0.5
    # ---- candidates entering multiclass_nms, keyed by (level, anchor index)
    boxes, scores_cm, cidx = PP.decode_candidates(plan.wsp, plan.post_in[0], plan.post_in[1], plan.post_in[2],
                                                  plan.img_info, True)
    torch.cuda.synchronize()
    tc = cfg.test_cfg
    sizes = [tuple(t.shape[-2:]) for t in plan.outs[0]]
    per_level = [min(h * w * 9, tc["nms_pre"]) for (h, w) in sizes]
    lvl_of = np.concatenate([np.full(k, l) for l, k in enumerate(per_level)])
    c_strict = c_tot = 0
    for i in range(n):
        g_idx, g_box, g_max = gold["cand_idx_%d" % i], gold["cand_boxes_%d" % i], gold["cand_max_%d" % i]
        my_idx, my_box = cidx[i].cpu().numpy(), boxes[i].cpu().numpy()
        my_sc = scores_cm[i].t().contiguous().cpu().numpy()
        gmap = {(int(l), int(a)): r for r, (l, a) in enumerate(zip(lvl_of, g_idx))}
        missing = 0
        matched_rows = {}
        for r, (l, a) in enumerate(zip(lvl_of, my_idx)):
            gr = gmap.get((int(l), int(a)))
            if gr is None:
                missing += 1
                # a candidate the reference did not select must sit at the level's top-k boundary
                cut = g_max[lvl_of == l].min()
                assert abs(my_sc[r].max() - cut) <= 2 * SCORE_TOL, \
                    "candidate set differs away from the top-k boundary (img %d level %d)" % (i, l)
                continue
            matched_rows[gr] = r
            assert abs(my_sc[r].max() - g_max[gr]) <= SCORE_TOL * (1 + abs(g_max[gr])), "candidate score differs"
            assert _box_ok(my_box[r], g_box[gr], ANCHOR_SIDE[int(l)]), "candidate box differs (level %d): %s vs %s" % (l, my_box[r], g_box[gr])
            c_strict += int(np.allclose(my_box[r], g_box[gr], rtol=RTOL, atol=ATOL)); c_tot += 1
        rows = DC.sample_index(name, "candrows%d" % i, 0, len(g_idx))[:DC.CAND_ROWS]
        g_rows = gold["cand_rows_%d" % i]
        for k_, gr in enumerate(rows):
            if int(gr) in matched_rows:
                assert np.allclose(my_sc[matched_rows[int(gr)]], g_rows[k_], rtol=SCORE_TOL, atol=SCORE_TOL), \
                    "candidate class scores differ"
        order_same = int((my_idx == g_idx).sum())
        report["cand_%d" % i] = dict(missing=missing, same_position=order_same, total=len(g_idx))
        if verbose:
            print("[%s] img %d candidates: %d of %d at the reference's position, %d at the top-k boundary swapped"
                  % (name, i, order_same, len(g_idx), missing))
        assert missing <= 0.01 * len(g_idx), "%d candidates differ" % missing
    report["cand_boxes_strict_allclose"] = (c_strict, c_tot)
    # ---- final detections, 1:1
    res = PP.split_results(dets, labels, counts)
    ms = mb = mrel = 0.0
    strict_fail = strict_tot = unmatched_tot = 0
    for i, (d, l) in enumerate(res):
        d, l = d.cpu().numpy(), l.cpu().numpy()
        gd, gl = gold["dets_%d" % i], gold["labels_%d" % i]
        assert d.shape[0] == gd.shape[0], "detection count %d != %d (img %d)" % (d.shape[0], gd.shape[0], i)
        used = np.zeros(len(d), bool)
        cut = min(gd[:, 4].min(), d[:, 4].min()) if len(gd) else 0.0
        lvl_by_box = {tuple(np.round(b_, 3)): int(l_) for b_, l_ in zip(gold["cand_boxes_%d" % i], lvl_of)}
        side = [ANCHOR_SIDE[lvl_by_box.get(tuple(np.round(gd[j, :4], 3)), 0)] for j in range(len(gd))]
        unmatched = []
        for j in range(len(gd)):
            cand = np.where((l == gl[j]) & ~used)[0]
            k = -1
            if cand.size:
                err = np.abs(d[cand, :4] - gd[j, :4]).max(1)
                k = cand[err.argmin()]
                if not (_box_ok(d[k, :4], gd[j, :4], side[j]) and abs(d[k, 4] - gd[j, 4]) <= SCORE_TOL * (1 + abs(gd[j, 4]))):
                    k = -1
            if k < 0:
                unmatched.append(j)
                continue
            used[k] = True
            ms = max(ms, float(abs(d[k, 4] - gd[j, 4])))
            e = np.abs(d[k, :4] - gd[j, :4])
            mb = max(mb, float(e.max()))
            mrel = max(mrel, float((e / (1.0 + np.abs(gd[j, :4]) + max(gd[j, 2] - gd[j, 0], gd[j, 3] - gd[j, 1], side[j]))).max()))
            bad = ~np.isclose(d[k, :4], gd[j, :4], rtol=RTOL, atol=ATOL)
            strict_fail += int(bad.sum()); strict_tot += 4
            if verbose and bad.any() and strict_fail <= 12:
                print("    img %d det %d label %d: coords %s vs ref %s (|d| %s)" % (i, j, gl[j], d[k, :4], gd[j, :4], e))
        for j in unmatched:
            # only the score-order cut (top-100 / score_thr) may drop a reference detection
            assert gd[j, 4] - cut <= 2 * SCORE_TOL or abs(gd[j, 4] - tc["score_thr"]) <= 2 * SCORE_TOL, \
                "reference detection %d of image %d (label %d score %.6f box %s) has no counterpart" % (
                    j, i, gl[j], gd[j, 4], gd[j, :4])
        unmatched_tot += len(unmatched)
        if verbose:
            print("[%s] img %d: %d detections, %d matched 1:1, %d swapped at the score cut" % (
                name, i, len(gd), len(gd) - len(unmatched), len(unmatched)))
    report.update(max_dscore=ms, max_dbox_px=mb, max_dbox_rel=mrel, strict_coord_fail=(strict_fail, strict_tot),
                  unmatched=unmatched_tot)
    if verbose:
        print("[%s] detections: max|dscore| %.3g, max|dbox| %.3g px (%.3g of 1+|coord|+size), %d of %d coordinates "
              "outside the strict per-coordinate allclose(1e-4, 1e-4)" % (name, ms, mb, mrel, strict_fail, strict_tot))
    return report
