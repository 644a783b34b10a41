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
