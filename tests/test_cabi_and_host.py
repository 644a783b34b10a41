"""CPU-side tests: the C-ABI library loads and exports every symbol declared in include/iou_b200.h,
the host-side mirror of the reference interface behaves, and the N>1 plumbing works over gloo."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
import iou_aware_single_stage_object_detector_b200 as P
from iou_aware_single_stage_object_detector_b200 import dist as D
from iou_aware_single_stage_object_detector_b200 import engine as E
from iou_aware_single_stage_object_detector_b200 import lib as L

CFG_DIR = os.path.join(ROOT, "configs", "iou_aware_single_stage_detector")


def header_functions():
    txt = open(os.path.join(ROOT, "include", "iou_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(iou_[a-z0-9_]+)\s*\(", txt)))


def test_library_builds_and_exports_header_symbols():
    path = L.build()
    assert os.path.isfile(path)
    lib = ctypes.CDLL(path)
    names = header_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "symbol %s declared in the header is not exported" % n
    assert set(names) == set(L.EXPORTED_SYMBOLS)
    lib.iou_abi_version.restype = ctypes.c_int
    assert lib.iou_abi_version() == 8


def test_sass_is_blackwell_native():
    """tcgen05 / TMA show up in SASS as UTCHMMA / UTMALDG / LDTM (B200_PROFILING.md)."""
    out = subprocess.run(["cuobjdump", "-sass", L.build()], capture_output=True, text=True).stdout
    for mnem in ("UTCHMMA", "UTMALDG", "LDTM"):
        assert mnem in out, mnem
    assert "HMMA.16816" not in out       # no legacy mma.sync path


def test_struct_layouts_match_header():
    # sizes the C side computes for the same structs (plain-C layout rules)
    assert ctypes.sizeof(L.ConvSegment) == 16
    assert ctypes.sizeof(L.PostprocCfg) == 5 * 4 + 3 * 8 * 4 + 8 * 16 * 4 * 4 + 8 * 4 + 4 * 4 + 4   # + decode_mode (ABI 2)
    assert L.ConvDesc.src.offset % 8 == 0 and L.ConvDesc.weight.offset % 8 == 0
    lib = L.load()                                   # the compiled C structs agree with the ctypes mirror
    assert lib.iou_sizeof(0) == ctypes.sizeof(L.PostprocCfg)
    assert lib.iou_sizeof(1) == ctypes.sizeof(L.ConvDesc)
    assert lib.iou_sizeof(2) == ctypes.sizeof(L.ConvSegment)


def test_errors_are_loud_without_gpu():
    lib = L.load()
    cfg = L.PostprocCfg()
    assert lib.iou_postproc_num_candidates(ctypes.byref(cfg)) < 0
    assert b"num_levels" in lib.iou_last_error()
    with pytest.raises(RuntimeError):
        P.nms(torch.zeros(4, 5), 0.5)                 # CPU tensor -> no fallback
    with pytest.raises(RuntimeError):
        P.multiclass_nms(torch.zeros(4, 4), torch.zeros(4, 81), 0.05, dict(type='nms', iou_thr=0.5), 100)
    cfg = P.Config.fromfile(os.path.join(CFG_DIR, "iou_aware_retinanet_r50_fpn_1x_4gpu.py"))
    with pytest.raises(RuntimeError):          # model-zoo URLs need a network: loud, not silent
        P.build_detector(cfg.model)
    cfg.model.pretrained = None
    det = P.build_detector(cfg.model)
    with pytest.raises(RuntimeError):
        det.backbone(torch.zeros(1, 3, 64, 64))


@pytest.mark.parametrize("name,params_m", [("iou_aware_retinanet_r50_fpn_1x_4gpu.py", 37.99),
                                           ("iou_aware_retinanet_r50_fpn_1x_2gpu.py", 37.99),
                                           ("iou_aware_retinanet_r101_fpn_1x_4gpu.py", 56.98),
                                           ("iou_aware_retinanet_x101_32x4d_fpn_1x_4gpu.py", None),
                                           ("iou_aware_retinanet_x101_64x4d_fpn_1x.py", None)])
def test_configs_build(name, params_m):
    cfg = P.Config.fromfile(os.path.join(CFG_DIR, name))
    cfg.model.pretrained = None
    det = P.build_detector(cfg.model, train_cfg=cfg.train_cfg, test_cfg=cfg.test_cfg)
    assert type(det).__name__ == "RetinaNet" and type(det.bbox_head).__name__ == "IoUawareRetinaHead"
    if params_m:
        assert abs(sum(p.numel() for p in det.parameters()) / 1e6 - params_m) < 0.01
    assert det.test_cfg.nms.iou_thr == 0.5 and det.test_cfg.get('nms_pre', -1) == 1000
    h = det.bbox_head
    assert h.num_anchors == 9 and h.cls_out_channels == 80
    assert h.retina_cls.weight.shape == (720, 256, 3, 3) and h.retina_iou.weight.shape == (9, 256, 3, 3)
    # reference init of the head (iou_aware_retina_head.py:153-165)
    assert float(h.retina_cls.bias[0]) == pytest.approx(-4.59512, abs=1e-4)
    assert float(h.retina_cls.weight.std()) == pytest.approx(0.01, rel=0.05)


@pytest.mark.reference
def test_reference_config_files_load_unchanged_and_keys_match():
    ref_dir = "/root/reference/configs/iou_aware_single_stage_detector"
    for f in sorted(os.listdir(ref_dir)):
        cfg = P.Config.fromfile(os.path.join(ref_dir, f))
        cfg.model.pretrained = None
        det = P.build_detector(cfg.model, train_cfg=cfg.train_cfg, test_cfg=cfg.test_cfg)
        assert cfg.test_cfg.max_per_img == 100 and cfg.dist_params.backend == 'nccl'
        mine = os.path.join(CFG_DIR, f)
        if os.path.isfile(mine):
            c2 = P.Config.fromfile(mine)
            assert dict(c2.test_cfg) == dict(cfg.test_cfg)
            c2.model.pretrained = None
            assert c2.model == cfg.model
    # state_dict schema == the reference's, checked in a subprocess so that the
    # reference's `mmdet` package never shares a process with this repo's modules
    code = ("import sys; sys.path.insert(0, %r)\n"
            "from oracle import ref_shim\n"
            "m, cfg = ref_shim.build_reference_detector()\n"
            "print('\\n'.join('%%s %%s' %% (k, tuple(v.shape)) for k, v in m.state_dict().items()))\n") % ROOT
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    ref_keys = [l for l in out.stdout.strip().splitlines() if l.startswith(("backbone", "neck", "bbox_head"))]
    cfg = P.Config.fromfile(os.path.join(ref_dir, "iou_aware_retinanet_r50_fpn_1x_4gpu.py"))
    cfg.model.pretrained = None
    det = P.build_detector(cfg.model, test_cfg=cfg.test_cfg)
    my_keys = ['%s %s' % (k, tuple(v.shape)) for k, v in det.state_dict().items()]
    assert sorted(my_keys) == sorted(ref_keys)
    assert len(my_keys) == 356


def test_registry_and_builder_semantics():
    assert 'IoUawareRetinaHead' in P.HEADS.module_dict and 'RetinaNet' in P.DETECTORS.module_dict
    assert 'ResNet' in P.BACKBONES.module_dict and 'ResNeXt' in P.BACKBONES.module_dict
    assert 'FPN' in P.NECKS.module_dict and 'FocalLoss' in P.LOSSES.module_dict
    with pytest.raises(KeyError):
        P.build_head(dict(type='NoSuchHead'))
    with pytest.raises(KeyError):
        P.HEADS.register_module(P.IoUawareRetinaHead)
    with pytest.raises(TypeError):
        P.HEADS.register_module(int)
    seq = P.build([dict(type='FocalLoss', use_sigmoid=True), dict(type='SmoothL1Loss')], P.LOSSES)
    assert isinstance(seq, torch.nn.Sequential) and len(seq) == 2


def test_anchor_generator_matches_oracle():
    from oracle import postproc as op
    for s in (8, 16, 32, 64, 128):
        g = P.AnchorGenerator(s, op.retina_anchor_scales(4, 3), [0.5, 1.0, 2.0])
        assert torch.equal(g.base_anchors, op.base_anchors(s, op.retina_anchor_scales(4, 3), [0.5, 1.0, 2.0]))
        assert torch.equal(g.grid_anchors((3, 5), s), op.grid_anchors(g.base_anchors, 3, 5, s))
    rois = torch.tensor([[0., 0., 31., 31.]])
    out = P.delta2bbox(rois, torch.tensor([[0.1, -0.2, 0.3, 5.0]]), max_shape=(800, 1333, 3))
    assert np.allclose(out.numpy(), [[0, 0, 39.7977, 799]], atol=1e-3)


def test_bbox2result_formats():
    d = torch.tensor([[1., 2., 3., 4., .9], [5., 6., 7., 8., .8]])
    l = torch.tensor([3, 0])
    r = P.bbox2result(d, l, 81)
    assert len(r) == 80 and r[3].shape == (1, 5) and r[0][0, 4] == np.float32(.8) and r[1].shape == (0, 5)
    e = P.bbox2result(torch.zeros(0, 5), torch.zeros(0, dtype=torch.long), 81)
    assert len(e) == 80 and all(a.shape == (0, 5) and a.dtype == np.float32 for a in e)


def test_weight_packing_and_layout_helpers():
    w = torch.randn(7, 64, 3, 3)
    p = E.pack_weight(w, 16)
    assert p.shape == (9 * 16, 128) and p.dtype == torch.bfloat16
    hi, lo = p[:, :64].float(), p[:, 64:].float()
    tap = (hi + lo).reshape(9, 16, 64)
    assert torch.allclose(tap[4, :7], w[:, :, 1, 1], rtol=2e-5, atol=1e-30)
    assert float(tap[:, 7:].abs().max()) == 0
    assert E.pick_block_n(720) == (240, 720) and E.pick_block_n(45) == (48, 48)
    assert E.pick_block_n(256) == (256, 256) and E.pick_block_n(2048) == (256, 2048)
    assert E.pick_block_n(64) == (64, 64)
    m = E.FlatMap([(8, 100, 168), (8, 50, 84), (8, 25, 42), (8, 13, 21), (8, 7, 11)], 256, "meta",
                  tensor=torch.empty(0), ptr=0, rows=0)
    starts = [s[0] for s in m.segs]
    assert all(s % 128 == 0 for s in starts) and starts[1] == 138752
    # stride-2 taps: tap (r,s) reads phase (r&1, s&1) at offset (r>>1, s>>1)
    assert E.TAPS_3X3_S2[0] == (0, 0, 0) and E.TAPS_3X3_S2[4] == (3, 0, 0) and E.TAPS_3X3_S2[8] == (0, 1, 1)
    # tile-count-aware N tile: small maps get narrow tiles, big maps keep N=256
    assert E.pick_block_n(256, 22)[0] == 64 and E.pick_block_n(256, 1469)[0] == 256
    assert E.pick_block_n(256, 4266, max_bn=128)[0] == 128
    # ResNeXt grouped weights -> block-diagonal 64-channel blocks
    wg = torch.randn(256, 4, 3, 3)
    pg = E.pack_weight_grouped(wg, 64)
    tg = (pg[:, :64].float() + pg[:, 64:].float()).reshape(9, 256, 64)
    assert torch.allclose(tg[4, 70, 4:8], wg[70, :, 1, 1], rtol=2e-5, atol=1e-30)
    assert float(tg[4, 70, :4].abs().max()) == 0 and float(tg[4, 70, 8:].abs().max()) == 0


def test_f16f8_weight_packing_pairs_with_the_activation_bytes():
    """engine.pack_f16f8 (conv passes = 2) against the byte-level emulation of the two tensor-core passes in
    oracle/split_fmt.py: the K order of [Wl8 | W8] must meet the activations' [x8 | l8], the per-channel power-of-two
    scale S_n shared by all three weight copies must be undone by the returned 1 / S_n, and the result must be fp32-grade (~2^-16 per product)."""
    from oracle import split_fmt as SF
    g = torch.Generator().manual_seed(0)
    x = torch.randn(300, 128, generator=g) * torch.exp(torch.randn(300, 128, generator=g))
    w = torch.randn(40, 128, generator=g) * 0.05 * torch.exp(2 * torch.randn(40, 1, generator=g))
    w[7] = 0                                                    # an all-zero (padding) output channel
    rows, cs = E.pack_f16f8(w.view(1, 40, 128))
    assert rows.shape == (40, 256) and rows.dtype == torch.bfloat16 and cs.shape == (40,)
    assert bool((torch.log2(cs) == torch.log2(cs).round()).all())          # powers of two
    wh, wl8, w8 = E.unpack_f16f8_rows(rows, 128)
    assert float(w8.abs().max()) <= 16 and float(w8.abs().amax(1)[w.abs().amax(1) > 0].min()) >= 7
    assert float(wh.abs().max()) <= 32768
    y = SF.emulate_gemm(SF.encode_rows(x), rows, cs)
    ref = x.double() @ w.double().t()
    bound = x.double().abs() @ w.double().abs().t()
    assert float(((y - ref).abs() / bound.clamp_min(1e-30)).max()) < 4e-5
    assert float(y[:, 7].abs().max()) == 0
    # the packed default weights carry their fp32 source for the repack in Engine.conv
    p = E.pack_weight(torch.randn(16, 64, 3, 3, generator=g), 16)
    assert p.src32.shape == (9, 16, 64)
    assert float((SF.decode_rows(SF.encode_rows(x)) - x).abs().max() / x.abs().max()) < 2.0 ** -15


def test_k_concatenated_weights_share_one_scale():
    """conv3(t2) + downsample(x) as one GEMM (engine.add_backbone): both taps' weights are packed under ONE per-channel
    scale, the narrower tap is zero-padded to the wider K, and the byte-level emulation of the two passes over each
    source's own channels adds up to the fp64 sum of both products."""
    from oracle import split_fmt as SF
    g = torch.Generator().manual_seed(4)
    width, cin, cout = 64, 128, 32
    t2 = torch.randn(200, width, generator=g).relu() * 2
    x = torch.randn(200, cin, generator=g).relu() * 3
    w3 = torch.randn(cout, width, generator=g) * 0.05
    wd = torch.randn(cout, cin, generator=g) * 0.3 * torch.exp(torch.randn(cout, 1, generator=g))
    wt = torch.zeros(2, cout, cin)
    wt[0, :, :width], wt[1] = w3, wd
    rows, inv_s = E.pack_f16f8(wt)
    assert rows.shape == (2 * cout, 2 * cin)
    wh, wl8, w8 = E.unpack_f16f8_rows(rows, cin)
    assert float(wh[:cout, width:].abs().max()) == 0 and float(w8[:cout, width:].abs().max()) == 0     # padding
    # tap 0 contracts only t2's `width` channels: the first `width` K columns of both planes of its weight rows
    def tap(a, rows_t, k):
        b = rows_t.contiguous().view(torch.uint8).view(cout, 4 * cin)
        hi, lo = b[:, :2 * k], b[:, 2 * cin:2 * cin + 2 * k]
        sub = torch.cat([hi, lo], dim=1).contiguous().view(torch.bfloat16)
        return SF.emulate_gemm(SF.encode_rows(a), sub, torch.ones(cout))
    acc = tap(t2, rows[:cout], width) + tap(x, rows[cout:], cin)          # ONE accumulator at scale S_n
    y = acc * inv_s.double().view(1, -1)
    ref = t2.double() @ w3.double().t() + x.double() @ wd.double().t()
    bound = t2.double().abs() @ w3.double().abs().t() + x.double().abs() @ wd.double().abs().t()
    assert float(((y - ref).abs() / bound).max()) < 6e-5


def test_tile_counts_and_default_scheme():
    # tiles stop at the last interior pixel: the 25x42 maps of 8 images are 74 tiles (75 with the trailing border rows)
    assert E.seg_tiles(8, 25, 42) == 74 and E.seg_tiles(8, 100, 168) == 1083 and E.seg_tiles(1, 1, 2) == 1
    assert sum(E.seg_tiles(8, h, w) for h, w in ((100, 168), (50, 84), (25, 42), (13, 21), (7, 11))) == 1466
    # detectors default to the fp16 + e4m3 scheme (every op exists for both element formats)
    cfg = P.Config.fromfile(os.path.join(CFG_DIR, "iou_aware_retinanet_r50_fpn_1x_4gpu.py"))
    cfg.model.pretrained = None
    det = P.build_detector(cfg.model, train_cfg=None, test_cfg=cfg.test_cfg)
    assert det.resolved_passes() == 2
    det.passes = 3
    assert det.resolved_passes() == 3
    fcfg = P.Config.fromfile(os.path.join(os.path.dirname(CFG_DIR), "fcos", "iou_aware_fcos_r50_caffe_fpn_gn_1x_4gpu.py"))
    fcfg.model.pretrained = None
    assert P.build_detector(fcfg.model, train_cfg=None, test_cfg=fcfg.test_cfg).resolved_passes() == 2


def test_shard_ranges_cover_batch():
    for world in (1, 2, 3, 4, 8):
        for B in (8, 13, 64):
            r = [D.shard_range(k, world, B) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == B and all(a[1] == b[0] for a, b in zip(r, r[1:]))


def _gloo_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    r, w, _ = D.init_dist("gloo")
    b, k = 3, 100
    g = torch.Generator().manual_seed(100 + rank)
    dets = torch.rand(b, k, 5, generator=g)
    labels = torch.randint(0, 80, (b, k), generator=g)
    counts = torch.tensor([rank, 50 + rank, 100], dtype=torch.int32)
    gd, gl, gc = D.gather_detections(dets, labels, counts)
    q.put((rank, gd, gl, gc, dets, labels, counts))
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


def test_all_gather_of_detections_world2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    import socket
    with socket.socket() as sk:            # a free port: a fixed one collides with a lingering earlier run
        sk.bind(('127.0.0.1', 0))
        port = sk.getsockname()[1]
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted([q.get(timeout=120) for _ in range(2)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, gd, gl, gc, *_ in got:
        assert gd.shape == (6, 100, 5) and gl.shape == (6, 100) and gl.dtype == torch.int64
        for src in range(2):
            assert torch.equal(gd[src * 3:(src + 1) * 3], got[src][4])
            assert torch.equal(gl[src * 3:(src + 1) * 3], got[src][5])
            assert torch.equal(gc[src * 3:(src + 1) * 3], got[src][6])


def _gloo_packed_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    D.init_dist("gloo")
    b, k = 3, 100
    g = torch.Generator().manual_seed(200 + rank)
    packed = torch.zeros(D.packed_layout(b, k)[3], dtype=torch.uint8)
    dets, labels, counts = D.packed_views(packed, b, k)
    dets.copy_(torch.rand(b, k, 5, generator=g))
    labels.copy_(torch.randint(0, 80, (b, k), generator=g))
    counts.copy_(torch.tensor([rank, 50 + rank, 100], dtype=torch.int32))
    pg = D.PackedGather(world, "cpu")
    outs = []
    for _ in range(3):                                  # buffers rotate between two slots
        buf, done = pg(packed)
        assert done is None
        outs.append(tuple(t.clone() for t in D.unpack_gathered(buf, world, b, k)))
    q.put((rank, outs, dets.clone(), labels.clone(), counts.clone()))
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


def test_packed_gather_world2_gloo():
    """The N > 1 result path of bench.py / detect_stream: ONE all-gather of the packed dets | labels | counts buffer."""
    import socket
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    with socket.socket() as sk:
        sk.bind(('127.0.0.1', 0))
        port = sk.getsockname()[1]
    procs = [ctx.Process(target=_gloo_packed_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted([q.get(timeout=120) for _ in range(2)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, outs, *_ in got:
        for gd, gl, gc in outs:
            assert gd.shape == (6, 100, 5) and gl.shape == (6, 100) and gl.dtype == torch.int64 and gc.dtype == torch.int32
            for src in range(2):
                assert torch.equal(gd[src * 3:(src + 1) * 3], got[src][2])
                assert torch.equal(gl[src * 3:(src + 1) * 3], got[src][3])
                assert torch.equal(gc[src * 3:(src + 1) * 3], got[src][4])


def test_packed_layout_is_aligned_and_round_trips():
    for b, k in ((1, 1), (8, 100), (3, 7), (64, 100)):
        o_d, o_l, o_c, total = D.packed_layout(b, k)
        assert o_d == 0 and o_l % 16 == 0 and o_c % 16 == 0 and total % 16 == 0
        assert o_l >= b * k * 20 and o_c >= o_l + b * k * 8 and total >= o_c + b * 4
        buf = torch.zeros(total, dtype=torch.uint8)
        d, l, c = D.packed_views(buf, b, k)
        assert d.shape == (b, k, 5) and l.shape == (b, k) and c.shape == (b,)
        d.fill_(1.5); l.fill_(7); c.fill_(3)
        d2, l2, c2 = D.unpack_gathered(buf, 1, b, k)
        assert float(d2.sum()) == 1.5 * b * k * 5 and int(l2.sum()) == 7 * b * k and int(c2.sum()) == 3 * b


def test_bench_reference_arm_prints_the_contract_line():
    """bench.py --impl reference (the CPU arm of the measurement contract) runs here: one JSON line with the keys the
    driver reads; the bounded sample stops at the time cap."""
    import json
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--reference-budget-s", "1"], capture_output=True, text=True, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert line["impl"] == "reference" and line["unit"] == "images/sec" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["higher_is_better"] is True


def test_result_formats_match_reference(tmp_path):
    """det2json / xyxy2xywh / results2json vs the golden written by the reference's own functions
    (tests/golden/gen_golden.py::gen_results_json); batch_bbox2result == per-image bbox2result."""
    import json
    import pickle
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import gen_golden_fixtures as GF
    ds, results = GF.results_fixture()
    want = json.load(open(os.path.join(ROOT, "tests", "golden", "results_det2json.json")))
    assert P.det2json(ds, results) == want
    out = str(tmp_path / "r.pkl.json")
    assert P.results2json(ds, results, out) == want and json.load(open(out)) == want
    P.dump_results(results, str(tmp_path / "r.pkl"))
    back = pickle.load(open(str(tmp_path / "r.pkl"), "rb"))
    assert all(np.array_equal(a, b) for ra, rb in zip(results, back) for a, b in zip(ra, rb))
    with pytest.raises(TypeError):
        P.results2json(ds, [3], out)
    with pytest.raises(TypeError):
        P.dump_results(results, str(tmp_path / "r.txt"))
    # padded batch tensors -> per-image per-class arrays
    rs = np.random.RandomState(2)
    n, K = 3, 100
    dets = rs.rand(n, K, 5).astype(np.float32)
    labels = rs.randint(0, 80, size=(n, K)).astype(np.int64)
    counts = np.array([100, 0, 37], dtype=np.int32)
    got = P.batch_bbox2result(torch.from_numpy(dets), torch.from_numpy(labels), torch.from_numpy(counts), 81)
    for i in range(n):
        k = counts[i]
        want_i = P.bbox2result(torch.from_numpy(dets[i, :k]), torch.from_numpy(labels[i, :k]), 81)
        assert len(got[i]) == 80
        for a, b in zip(got[i], want_i):
            assert a.dtype == np.float32 and a.shape == b.shape and np.array_equal(a, b)
