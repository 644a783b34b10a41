"""Multi-GPU plumbing: one process per GPU, the image batch is the only thing partitioned
(images are independent: per-image loop at iou_aware_retina_head.py:434-460, eval-mode BN), and the
only collective is ONE all-gather of the fixed-size padded detections per step -- the replacement for
the reference's pickle-files-in-a-tmpdir gather (tools/test.py:63-102).  Process-group set-up follows
mmdet/apis/env.py:26-31 (env:// rendezvous, backend from dist_params).
"""
import os

import torch
import torch.distributed as dist


def init_dist(backend="nccl"):
    """RANK / LOCAL_RANK / WORLD_SIZE / MASTER_* from the environment (torchrun)."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    if backend == "nccl":
        torch.cuda.set_device(local % max(torch.cuda.device_count(), 1))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def shard_range(rank, world, global_batch):
    """Contiguous shard [lo, hi) of the global image batch owned by `rank` (DistributedSampler-like,
    datasets/loader/build_loader.py:22-31, but contiguous so the gathered order is the global order)."""
    base, rem = divmod(global_batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def pack_detections(dets, labels, counts):
    """(b,K,5) f32, (b,K) i64, (b,) i32 -> one flat fp32 buffer [b*K*6 + b] (labels/counts are small
    integers, exactly representable)."""
    b, k, _ = dets.shape
    body = torch.cat([dets, labels.to(dets.dtype).unsqueeze(-1)], dim=-1).reshape(-1)
    return torch.cat([body, counts.to(dets.dtype)])


def unpack_detections(flat, b, k):
    body = flat[: b * k * 6].reshape(b, k, 6)
    return body[..., :5].contiguous(), body[..., 5].to(torch.int64), flat[b * k * 6:].to(torch.int32)


def gather_detections(dets, labels, counts, world=None):
    """ONE all-gather (NCCL over NVLink / NVSwitch; gloo in the CPU tests) of every rank's padded
    detections.  All ranks must hold the same per-rank batch b.  Returns global (B,K,5), (B,K), (B,)."""
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        return dets, labels, counts
    b, k, _ = dets.shape
    mine = pack_detections(dets, labels, counts)
    out = torch.empty(world * mine.numel(), dtype=mine.dtype, device=mine.device)
    dist.all_gather_into_tensor(out, mine)
    parts = [unpack_detections(out[r * mine.numel():(r + 1) * mine.numel()], b, k) for r in range(world)]
    return (torch.cat([p[0] for p in parts]), torch.cat([p[1] for p in parts]),
            torch.cat([p[2] for p in parts]))


# ------------------------------------------------------------------------------------------------
# Packed results: the post-processing kernels write dets | labels | counts of a batch into ONE byte buffer
# (postproc.PostprocWorkspace.packed), so a step's results leave the GPU as one all-gather + one device->host copy
# with no pack / cast kernels in between.
def _a16(x):
    return (x + 15) // 16 * 16


def packed_layout(b, k):
    """(dets offset, labels offset, counts offset, total bytes) of a packed result buffer for b images x k rows:
    [b*k*5 fp32][b*k int64][b int32], each section 16-byte aligned."""
    o_lab = _a16(b * k * 5 * 4)
    o_cnt = o_lab + _a16(b * k * 8)
    return 0, o_lab, o_cnt, o_cnt + _a16(b * 4)


def packed_views(buf, b, k):
    """(dets [b,k,5] f32, labels [b,k] i64, counts [b] i32) as VIEWS of a packed uint8 buffer (device or host)."""
    o_d, o_l, o_c, total = packed_layout(b, k)
    assert buf.dtype == torch.uint8 and buf.numel() >= total
    dets = buf[o_d:o_d + b * k * 20].view(torch.float32).view(b, k, 5)
    labels = buf[o_l:o_l + b * k * 8].view(torch.int64).view(b, k)
    counts = buf[o_c:o_c + b * 4].view(torch.int32)
    return dets, labels, counts


def unpack_gathered(buf, world, b, k):
    """Packed buffers of `world` ranks back to back (what the all-gather returns) -> global (B,k,5), (B,k), (B,)."""
    total = packed_layout(b, k)[3]
    parts = [packed_views(buf[r * total:(r + 1) * total], b, k) for r in range(world)]
    if world == 1:
        return parts[0]
    return (torch.cat([p[0] for p in parts]), torch.cat([p[1] for p in parts]), torch.cat([p[2] for p in parts]))


class PackedGather(object):
    """ONE ncclAllGather of the packed results per step, issued on a SIDE stream: the compute stream records an
    event behind final_select and carries on with the next batch, the collective (19 KB per rank, latency-bound)
    and the device->host read-back overlap it.  Replaces tools/test.py:63-102 (pickle files + barrier).
    Usage per step:   buf, done = gather(plan.wsp.packed)   ->  `done` (event on the side stream) guards both the
    gathered buffer and the re-use of the plan's packed buffer."""

    def __init__(self, world, device, slots=2):
        self.world, self.device, self.slots = world, torch.device(device), slots
        self.stream = torch.cuda.Stream(self.device) if self.device.type == "cuda" else None
        self.bufs, self.turn = {}, 0

    def __call__(self, packed):
        n = packed.numel()
        key = (n, self.turn % self.slots)
        self.turn += 1
        if key not in self.bufs:
            self.bufs[key] = torch.empty(self.world * n, dtype=torch.uint8, device=packed.device)
        out = self.bufs[key]
        if self.stream is None:                      # CPU tensors (gloo tests): no streams
            if self.world > 1:
                dist.all_gather_into_tensor(out, packed)
            else:
                out.copy_(packed)
            return out, None
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(self.device))
        self.stream.wait_event(ready)
        with torch.cuda.stream(self.stream):
            if self.world > 1:
                dist.all_gather_into_tensor(out, packed)
            else:
                out.copy_(packed, non_blocking=True)
            done = torch.cuda.Event()
            done.record(self.stream)
        return out, done

