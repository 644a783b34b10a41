#!/usr/bin/env python
"""bench.py -- images/sec of IoU-aware RetinaNet R50-FPN inference (backbone + FPN + head +
get_bboxes), 800x1344 padded input (img_shape 800x1333), bs=8 per GPU, synthetic data.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port)

One "step" = one pass of the whole hot path over one batch of 8 images per GPU.  Rank 0 prints ONE
JSON line.  `value` = whole-job images/sec with the batch already resident in HBM (device-timed, max
over ranks; consecutive steps alternate between two launch plans on two streams, `config.pipeline`);
`e2e` = the same through the public API (detect_stream) with pinned HOST images copied in and detections read
back every step; `roofline` = per-launch CUDA-event timing of the tcgen05 conv kernel (eager pass, rescaled to
the forward's CUDA-graph replay time, `graph_over_eager`) against the measured bf16 peak (algorithmic 2*MAC
flops: the default split scheme -- one fp16 pass + one e4m3 pass of doubled K, `--passes 2` -- costs two bf16-pass
equivalents of tensor-pipe time, so its ceiling is `frac` = 0.5; `--passes 3`, bf16 hi|lo x3: 0.33);
`cpu_baseline` = the oracle (a port of the
reference's CPU algorithm) on this box's host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "images/sec (1333x800, bs=8/GPU) IoU-aware RetinaNet-R50 inference"
UNIT = "images/sec"
CFG_DIR = os.path.join(ROOT, "configs", "iou_aware_single_stage_detector")
MODELS = {   # name -> (config file, depth, groups, default per-GPU batch)   (BASELINE.json configs 1-3)
    "r50": ("iou_aware_retinanet_r50_fpn_1x_4gpu.py", 50, 1, 8),
    "r101": ("iou_aware_retinanet_r101_fpn_1x_4gpu.py", 101, 1, 8),
    "x101_32x4d": ("iou_aware_retinanet_x101_32x4d_fpn_1x_4gpu.py", 101, 32, 8),
    "x101_64x4d": ("iou_aware_retinanet_x101_64x4d_fpn_1x.py", 101, 64, 4),
}
CFG = os.path.join(CFG_DIR, MODELS["r50"][0])
H, W, BATCH = 800, 1344, 8
DEPTH, GROUPS, MODEL = 50, 1, "r50"
GFLOP_PER_IMG = {"head": 290.36, "fpn": 36.39, "backbone": 175.16}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return dict(hbm=d.get("hbm_gbs", 6650.0), tf_burst=d.get("bf16_tflops", 1590.0),
                    tf_sustained=d.get("bf16_tflops_sustained", 1400.0), source="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback")


def measured_traffic(model, batch):
    """(bytes per conv launch, source) from profiles/conv_dram_traffic.json (ncu dram__bytes_read+write summed over
    the conv launches of one step / launches; written by tools/ncu_traffic.py), or (None, why)."""
    p = os.path.join(ROOT, "profiles", "conv_dram_traffic.json")
    if not os.path.isfile(p):
        return None, "no ncu capture committed for this build"
    d = json.load(open(p))
    e = d.get("%s_bs%d" % (model, batch))
    if e is None:
        return None, "no ncu capture of %s bs=%d in profiles/conv_dram_traffic.json" % (model, batch)
    return e["bytes_per_launch"], "ncu, %s (%d conv launches, %.2f GB per step)" % (e["source"], e["launches"], e["bytes_per_step"] / 1e9)


class ClockSampler(object):
    """nvidia-smi clock / throttle sampling DURING the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        self.index, self.proc, self.rows = index, None, []

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "25"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return None
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            try:
                pw.append(float(f[2]))
            except ValueError:
                pass
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return None
        busy = sorted(sm)[len(sm) // 2:]          # upper half ~ samples under load
        out = {"sm_mhz": sorted(busy)[len(busy) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons),
               "samples": len(sm), "sm_mhz_min": min(sm)}
        if pw:
            out["power_w_median"], out["power_w_max"] = sorted(pw)[len(pw) // 2], max(pw)
        return out


# ------------------------------------------------------------------------------------------------
def build_detector(device, weights, seed=0):
    import torch
    import iou_aware_single_stage_object_detector_b200 as P
    from iou_aware_single_stage_object_detector_b200 import synthetic
    cfg = P.Config.fromfile(CFG)
    cfg.model.pretrained = None                       # tools/test.py:138
    torch.manual_seed(seed)
    det = P.build_detector(cfg.model, train_cfg=None, test_cfg=cfg.test_cfg)
    det.eval()
    if weights == "spread":
        sd = {k: v.clone() for k, v in det.state_dict().items()}
        synthetic.spread_state_dict_(sd, synthetic.cuda_forward_fn(det, device), seed=seed + 1)
        det.load_state_dict(sd)
    return det.to(device), cfg


def run_ours(args):
    import torch
    import torch.distributed as dist
    from iou_aware_single_stage_object_detector_b200 import dist as D
    from iou_aware_single_stage_object_detector_b200 import lib as L
    from iou_aware_single_stage_object_detector_b200 import synthetic
    rank, world, local = D.init_dist("nccl")
    if world != args.gpus and rank == 0:
        print("note: WORLD_SIZE=%d, --gpus=%d" % (world, args.gpus), file=sys.stderr)
    dev = torch.device("cuda", torch.cuda.current_device())
    det, cfg = build_detector(dev, args.weights)
    det.passes = args.passes
    det.use_cuda_graph = not args.no_graph
    img_host, metas = synthetic.synthetic_batch(BATCH, H, W, seed=rank, pin=True)
    img_dev = img_host.to(dev)
    plan = det.fused_plan(img_dev.shape, dev, rescale=True)
    plan.img.copy_(img_dev)
    from iou_aware_single_stage_object_detector_b200 import postproc as PP
    plan.img_info.copy_(PP.make_img_info(metas, "cpu"))

    # Two independent launch plans (own buffers, own CUDA graph) replayed alternately on two streams: the tail of
    # every persistent conv kernel and the latency-bound post-processing of step i are filled by kernels of step
    # i+1 (the same schedule detect_stream uses).  Profiling / eager runs keep ONE plan on one stream.
    pipelined = not (args.no_graph or args.ncu_range or args.no_pipeline)
    plans, main_stream = [plan], torch.cuda.current_stream()
    streams = [main_stream]
    if pipelined:
        for slot in range(1, args.plans):
            plan_b = det.fused_plan(img_dev.shape, dev, rescale=True, slot=slot)
            plan_b.img.copy_(img_dev)
            plan_b.img_info.copy_(PP.make_img_info(metas, "cpu"))
            plans.append(plan_b)
            streams.append(torch.cuda.Stream(dev))
    step_no = [0]
    # N > 1: ONE all-gather of the packed detections per step, on a side stream (dist.PackedGather): the compute
    # streams never wait for the collective, only a plan's NEXT run waits for the gather of its previous results
    pg = D.PackedGather(world, dev) if world > 1 and not os.environ.get("IOU_BENCH_NO_GATHER") else None   # (diagnostic knob)
    busy = [None] * len(plans)

    def step_device():
        k = step_no[0] % len(plans)
        step_no[0] += 1
        with torch.cuda.stream(streams[k]):
            if busy[k] is not None:
                streams[k].wait_event(busy[k])
            out = plans[k].run()
            if pg is not None:
                out, busy[k] = pg(plans[k].wsp.packed)
            return out

    def fork():                                   # the side streams start after everything queued on the main one
        if pipelined:
            ev = torch.cuda.Event()
            ev.record(main_stream)
            for st_ in streams[1:]:
                st_.wait_event(ev)

    def join():                                   # ... and the main stream ends after the side streams
        if pipelined:
            for st_ in streams[1:]:
                main_stream.wait_stream(st_)
        if pg is not None:
            main_stream.wait_stream(pg.stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    fork()
    for _ in range(2 * max(args.warmup, 3)):
        step_device()
    join()
    barrier()
    # ---- device-timed region (inputs resident in HBM) ------------------------------------------
    L.launch_count = 0
    sampler = ClockSampler(torch.cuda.current_device())
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if args.ncu_range:
        torch.cuda.profiler.start()
    e0.record()
    fork()
    for _ in range(args.steps):
        out = step_device()
    join()
    e1.record()
    if args.ncu_range:
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    launches = L.launch_count
    t = torch.tensor([ms], device=dev)
    ms_ranks = None
    if world > 1:
        every = torch.zeros(world, device=dev)
        dist.all_gather_into_tensor(every, t)            # which rank set the max: a slow GPU or the collective
        ms_ranks = [round(float(v) / args.steps, 3) for v in every.tolist()]
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = world * BATCH * args.steps / (ms / 1e3)
    # ---- end to end through the public API: pinned host images in, detections out, every step ----
    # Public API call: detect_stream() takes HOST batches, copies each to the GPU (overlapping the copy of
    # batch i+1 with the compute of batch i) and returns the detections on the host.  Every step's input
    # crosses PCIe inside the timed region and every step's result is read back.
    host_bufs = [img_host, img_host.clone().pin_memory()]

    def host_batches(k):
        for i in range(k):
            yield host_bufs[i & 1], metas

    def run_e2e(k):
        out = None
        for dets, labels, counts in det.detect_stream(host_batches(k), rescale=True, device=dev, gather=pg):
            out = (dets, labels, counts)
        return out
    res = run_e2e(2)
    barrier()
    t0 = time.perf_counter()
    res = run_e2e(args.steps)
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * BATCH * args.steps / float(t.item())
    h2d = img_host.numel() * 4 + BATCH * 8 * 4
    # extra: the same end-to-end loop fed with raw uint8 frames (800x1333x3 BGR); normalise / pad / CHW run
    # on the device (api.ImageTransform == mmdet/datasets/transforms.py:31-50 without the resize)
    import iou_aware_single_stage_object_detector_b200 as P
    tf = P.ImageTransform(cfg.img_norm_cfg["mean"], cfg.img_norm_cfg["std"], cfg.img_norm_cfg["to_rgb"], 32)
    gen = torch.Generator().manual_seed(100 + rank)
    frames = [torch.randint(0, 256, (BATCH, H, 1333, 3), generator=gen, dtype=torch.uint8).pin_memory() for _ in range(2)]

    def run_e2e_u8(k):
        out = None
        for out in det.detect_stream(((frames[i & 1], metas) for i in range(k)), rescale=True, device=dev,
                                     gather=pg, img_transform=tf):
            pass
        return out
    run_e2e_u8(2)
    barrier()
    t0 = time.perf_counter()
    run_e2e_u8(args.steps)
    barrier()
    t = torch.tensor([time.perf_counter() - t0], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_u8_value = world * BATCH * args.steps / float(t.item())
    d2h = world * plan.wsp.packed.numel()          # ONE packed buffer (dets | labels | counts of every rank) per step
    # ---- live roofline of the dominant kernel (conv_tap_gemm_kernel), rank 0 ---------------------
    roof, extra = None, {}
    if rank == 0:
        peaks = measured_peaks()
        prof = plan.eng.profile(iters=5)
        conv_ms_eager = sum(ms_ for name, ms_ in prof if name in plan.eng.op_flops)
        other_ms_eager = sum(ms_ for name, ms_ in prof if name not in plan.eng.op_flops)
        # Eager launches carry a launch gap inside every event pair (81 x a few us).  The per-launch durations are
        # therefore rescaled so that the forward's launches sum to its measured CUDA-graph replay time (same plan,
        # one stream, no step overlap): duration_i = graph_ms x eager_i / sum(eager).
        graph_scale = 1.0
        if plan.graph is not None:
            ge0, ge1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            pe_ = torch.cuda.Event(enable_timing=True)
            for _ in range(3):
                plan.run()
            ge0.record()
            for _ in range(10):
                plan.run()
            ge1.record()
            pe_.record()
            torch.cuda.synchronize()
            graph_ms = ge0.elapsed_time(ge1) / 10
            pp0, pp1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            pp0.record()
            for _ in range(5):
                PP.get_bboxes_device(plan.wsp, plan.post_in[0], plan.post_in[1], plan.post_in[2], plan.img_info, True,
                                     cls_max2=plan.cls_max2)        # the same post-processing launches the graph holds
            pp1.record()
            torch.cuda.synchronize()
            post_eager = pp0.elapsed_time(pp1) / 5
            graph_scale = graph_ms / (conv_ms_eager + other_ms_eager + post_eager)
            prof = [(name, ms_ * graph_scale) for name, ms_ in prof]
        conv_ms = conv_ms_eager * graph_scale
        other_ms = other_ms_eager * graph_scale
        conv_flops = sum(plan.eng.op_flops.values())
        head_ms = sum(ms_ for name, ms_ in prof if name.startswith("bbox_head."))
        head_flops = sum(f for name, f in plan.eng.op_flops.items() if name.startswith("bbox_head."))
        n_conv = sum(1 for name, _ in prof if name in plan.eng.op_flops)
        achieved = conv_flops / (conv_ms / 1e3) / 1e12
        # DRAM traffic of the conv kernel per launch, measured by ncu on THIS model and build (never a literal):
        # profiles/conv_dram_traffic.json is written by tools/ncu_traffic.py from a committed ncu launch list
        traffic, traffic_src = measured_traffic(MODEL, BATCH)
        # a 20-step timed region lasts ~0.2 s: the clocks are still near 1965 MHz (burst regime); the driver's
        # sustained peak was measured after seconds under the power cap (~1320 MHz).  Both fractions are given.
        roof = {"bound": "tensor", "kernel": "conv_tap_gemm_kernel", "achieved": round(achieved, 2),
                "peak": peaks["tf_sustained"], "peak_source": peaks["source"] + " bf16 sustained (cuBLAS)",
                "unit": "TFLOP/s", "frac": round(achieved / peaks["tf_sustained"], 4),
                "peak_burst": peaks["tf_burst"], "frac_burst": round(achieved / peaks["tf_burst"], 4),
                "regime": "timed region %.2f s at %s MHz: nearer the burst than the sustained peak; the scheme's "
                          "ceiling is 1/passes of either" % (ms / 1e3, clocks["sm_mhz"] if clocks else "?"),
                "traffic": traffic, "traffic_source": traffic_src,
                "launches_per_step": n_conv,
                "avg_launch_ms": round(conv_ms / max(n_conv, 1), 4),
                "algorithmic_gflop_per_step": round(conv_flops / 1e9, 1),
                "mma_passes": args.passes,
                "tensor_pipe_frac_est": round(args.passes * achieved / peaks["tf_sustained"], 4),
                "head_tower_tflops": round(head_flops / (head_ms / 1e3) / 1e12, 2),
                "head_tower_frac": round(head_flops / (head_ms / 1e3) / 1e12 / peaks["tf_sustained"], 4)}
        extra = {"conv_ms_per_step": round(conv_ms, 3), "layout_kernels_ms_per_step": round(other_ms, 3),
                 "conv_ms_per_step_eager": round(conv_ms_eager, 3), "graph_over_eager": round(graph_scale, 4)}
        if args.dump_ops:
            rows = [{"op": name, "ms": round(ms_, 4), "gflop": round(plan.eng.op_flops.get(name, 0.0) / 1e9, 2),
                     "tflops": round(plan.eng.op_flops.get(name, 0.0) / (ms_ / 1e3) / 1e12, 1) if ms_ > 0 else 0}
                    for name, ms_ in prof]
            json.dump(rows, open(args.dump_ops, "w"), indent=0)
        # post-processing pass alone (HBM-bound decode + NMS)
        pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        pe0.record()
        for _ in range(5):
            PP.get_bboxes_device(plan.wsp, plan.post_in[0], plan.post_in[1], plan.post_in[2], plan.img_info, True,
                                 cls_max2=plan.cls_max2)
        pe1.record()
        torch.cuda.synchronize()
        post_ms = pe0.elapsed_time(pe1) / 5
        # bytes the decode stage must read: per anchor the max class logit (two partial maxima from the retina_cls
        # epilogue, or the 80 class logits when it reduces them itself) + the iou logit; per candidate its 85 logits
        per_anchor = (2 + 1) if plan.cls_max2 is not None else (80 + 1)
        logits_bytes = BATCH * (201600 * per_anchor + plan.wsp.M * 85) * 4
        extra["postproc_premax"] = plan.cls_max2 is not None
        extra["postproc_ms_per_step"] = round(post_ms, 3)
        extra["decode_read_gbs_lower_bound"] = round(logits_bytes / (post_ms / 1e3) / 1e9, 1)
    line = None
    if rank == 0:
        line = {"metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": round(ms / args.steps, 3),
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                # what tests/test_gpu_detector_golden.py asserts against the live reference at 800x1344 (parity_util.py)
                "dtype": {3: "bf16 hi|lo split operands x3 on tcgen05, fp32 accumulate; activations stored with 16 "
                             "significant bits.  vs the fp32 reference: scores allclose(1e-4), box deltas within 1e-4 "
                             "(|dbox| <= 1e-4*(1+|coord|+box size)), detections matched 1:1",
                          2: "fp16 + e4m3 split operands (1 fp16 + 1 e4m3 MMA pass) on tcgen05, fp32 accumulate; "
                             "activations stored with 15-16 significant bits.  vs the fp32 reference: scores "
                             "allclose(1e-4) (measured 8e-6), box deltas within 1e-4 (|dbox| <= 1e-4*(1+|coord|+box "
                             "size); measured 0.04 px max, 19 of 800 coordinates outside a strict per-coordinate "
                             "allclose(1e-4,1e-4)), detections matched 1:1, NMS indices bit-exact on equal candidates"
                          }.get(args.passes, "bf16 (single pass; fails the parity bar, for scale only)"),
                "data": "synthetic",
                "config": {"workload": "IoU-aware RetinaNet %s-FPN inference bs=%d/GPU, synthetic 800x1344 "
                                       "(img_shape 800x1333), backbone+FPN+head+get_bboxes" % (MODEL.upper(), BATCH),
                           "global_batch": world * BATCH, "weights": args.weights + " (seeded random)",
                           "parallelism": "dp%d (image batch sharded, one all-gather of detections)" % world,
                           "l2": "per-step working set (~10 GB of activations) exceeds the 126 MB L2; no flush",
                           "cuda_graph": not args.no_graph,
                           "pipeline": "%d launch plans alternating on %d streams" % (len(plans), len(plans)) if pipelined else "none"},
                "e2e": {"value": round(e2e_value, 2), "unit": UNIT, "h2d_bytes_per_step": h2d,
                        "d2h_bytes_per_step": d2h},
                "e2e_uint8_frames": {"value": round(e2e_u8_value, 2), "unit": UNIT,
                                     "h2d_bytes_per_step": frames[0].numel() + BATCH * 8 * 4,
                                     "note": "extra: raw 800x1333x3 uint8 frames in, device-side ImageTransform"},
                "gpu_launches": launches, "clocks": clocks, "roofline": roof}
        line.update(extra)
        if ms_ranks is not None:
            line["ms_per_step_ranks"] = ms_ranks          # `ms_per_step` is their max
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args.weights, images=args.cpu_images)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
def oracle_setup(weights, seed=0):
    """Reference CPU path = the oracle port (torch-CPU ATen forward + oracle post-processing)."""
    import torch
    import iou_aware_single_stage_object_detector_b200 as P
    from oracle import model as om
    from oracle import postproc as op
    cfg = P.Config.fromfile(CFG)
    cfg.model.pretrained = None
    torch.manual_seed(seed)
    det = P.build_detector(cfg.model, train_cfg=None, test_cfg=cfg.test_cfg)   # parameter container only
    sd = {k: v.clone() for k, v in det.state_dict().items()}
    if weights == "spread":
        om.spread_weights_(sd, seed=seed + 1, depth=DEPTH, groups=GROUPS)
    sc = op.retina_anchor_scales(4, 3)
    bases = [op.base_anchors(s, sc, [0.5, 1.0, 2.0]) for s in (8, 16, 32, 64, 128)]
    return sd, dict(cfg.test_cfg), bases, om, op


def oracle_image(sd, test_cfg, bases, om, op, img, meta):
    """One image exactly like tools/test.py:single_gpu_test drives the reference (1 image / iteration)."""
    cls, reg, iou = om.detector_forward(sd, img, DEPTH, GROUPS)
    d, l = op.get_bboxes_single([c[0] for c in cls], [r[0] for r in reg], [q[0] for q in iou],
                                [8, 16, 32, 64, 128], bases, meta["img_shape"], meta["scale_factor"],
                                test_cfg, rescale=True, nms_mode="cpu")
    return op.bbox2result(d, l, 81)


def cpu_baseline(weights, images=3):
    import torch
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    from iou_aware_single_stage_object_detector_b200 import synthetic
    sd, test_cfg, bases, om, op = oracle_setup(weights)
    img, metas = synthetic.synthetic_batch(1, H, W, seed=0)
    oracle_image(sd, test_cfg, bases, om, op, img, metas[0])          # warm-up
    t0 = time.perf_counter()
    for _ in range(images):
        oracle_image(sd, test_cfg, bases, om, op, img, metas[0])
    dt = time.perf_counter() - t0
    return {"value": round(images / dt, 4), "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": "%d full-size images (1x3x800x1344), one per iteration as tools/test.py does, "
                      "oracle port of the reference CPU path (torch-CPU fp32 convs + C greedy NMS)" % images}


def run_reference(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return                                     # rank 0 alone runs and prints the reference arm
    # torchrun exports OMP_NUM_THREADS=1: the CPU arm must still use every host core it can
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    from iou_aware_single_stage_object_detector_b200 import synthetic
    sd, test_cfg, bases, om, op = oracle_setup(args.weights)
    img, metas = synthetic.synthetic_batch(1, H, W, seed=0)
    for _ in range(min(args.warmup, 1)):
        oracle_image(sd, test_cfg, bases, om, op, img, metas[0])
    # bounded sample: each step = ONE image of the batch-of-8 workload (the reference itself runs one
    # image per iteration, base.py:97-98); steps are capped so the arm ends within a few minutes
    t0 = time.perf_counter()
    done = 0
    for _ in range(args.steps):
        oracle_image(sd, test_cfg, bases, om, op, img, metas[0])
        done += 1
        if time.perf_counter() - t0 > args.reference_budget_s:
            break
    dt = time.perf_counter() - t0
    v = done / dt
    cores = torch.get_num_threads()
    sample = ("%d of %d steps executed (time cap %ds); each step = 1 image of the 8-image batch, run the "
              "way the reference runs (1 image/iteration)" % (done, args.steps, args.reference_budget_s))
    print(json.dumps({"impl": "reference", "metric": METRIC, "value": round(v, 4), "unit": UNIT,
                      "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                      "ms_per_step": round(dt / done * 1e3, 1), "higher_is_better": True, "scaling": "weak",
                      "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
                      "config": {"workload": "IoU-aware RetinaNet %s-FPN inference, synthetic 800x1344, CPU "
                                             "reference path (oracle port), 1 image per step" % MODEL.upper(),
                                 "weights": args.weights + " (seeded random)"},
                      "cpu_baseline": {"value": round(v, 4), "unit": UNIT, "cores": cores, "kind": "port",
                                       "sample": sample},
                      "e2e": {"value": round(v, 4), "unit": UNIT, "h2d_bytes_per_step": 0,
                              "d2h_bytes_per_step": 0}}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--weights", default="spread", choices=["spread", "reference-init"])
    ap.add_argument("--passes", type=int, default=2, choices=[1, 2, 3, 4],
                    help="2: fp16 pass + e4m3 correction pass = 2 bf16-pass equivalents (default, fp32-grade); "
                         "3: bf16 hi|lo x3 (fp32-grade); 1: plain bf16 (fails the parity bar, for scale only)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="eager launches (for ncu launch lists)")
    ap.add_argument("--no-pipeline", action="store_true", help="one launch plan on one stream (no step overlap)")
    ap.add_argument("--plans", type=int, default=2, help="launch plans alternating on as many streams (device-timed loop)")
    ap.add_argument("--ncu-range", action="store_true",
                    help="bracket the timed steps with cudaProfilerStart/Stop (ncu --profile-from-start off)")
    ap.add_argument("--dump-ops", default=None, help="write the per-launch CUDA-event table to this JSON file")
    ap.add_argument("--cpu-images", type=int, default=3)
    ap.add_argument("--reference-budget-s", type=int, default=150)
    ap.add_argument("--model", default="r50", choices=sorted(MODELS),
                    help="r50 = the headline workload; the others are BASELINE configs 2-3 (extra lines)")
    ap.add_argument("--batch", type=int, default=0, help="images per GPU (default: 8, 4 for x101_64x4d)")
    args = ap.parse_args()
    global CFG, BATCH, DEPTH, GROUPS, MODEL, METRIC
    cfg_file, DEPTH, GROUPS, default_bs = MODELS[args.model]
    CFG, MODEL = os.path.join(CFG_DIR, cfg_file), args.model
    BATCH = args.batch or default_bs
    if args.model != "r50" or BATCH != 8:
        METRIC = "images/sec (1333x800, bs=%d/GPU) IoU-aware RetinaNet-%s inference" % (BATCH, args.model.upper())
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
