"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the fp16 + e4m3 split-operand format (IOU_FMT_F16F8) that
`passes = 2` of the conv engine computes in (csrc/split_fmt.cuh, engine.pack_f16f8), and a byte-level emulation
of what the two tensor-core passes accumulate.  There is no reference counterpart (the reference computes in
fp32, mmdet/models/anchor_heads/iou_aware_retina_head.py:171-219); the checks built on this file are about the
format being self-consistent and fp32-grade, parity itself stays with oracle/model.py."""
import torch

LO_SCALE = 2048.0   # 2^11


def _e4m3(x):
    return x.clamp(-448.0, 448.0).to(torch.float8_e4m3fn)


def encode_rows(x):
    """fp32 [rows, C] -> uint8 [rows, 4C]: [C x fp16][per 8 channels: x8 x 8 | l8 x 8] (C % 8 == 0)."""
    rows, c = x.shape
    assert c % 8 == 0
    x = x.float().clamp(-65504.0, 65504.0)
    h = x.to(torch.float16)
    x8 = _e4m3(x).view(torch.uint8).view(rows, c // 8, 8)
    l8 = _e4m3((x - h.float()) * LO_SCALE).view(torch.uint8).view(rows, c // 8, 8)
    lo = torch.stack([x8, l8], dim=2).reshape(rows, 2 * c)
    return torch.cat([h.contiguous().view(torch.uint8).view(rows, 2 * c), lo], dim=1).contiguous()


def decode_rows(b):
    """uint8 [rows, 4C] -> fp32 [rows, C]: hi + l8 * 2^-11 (what residual adds / max pool / unpack read)."""
    rows, c = b.shape[0], b.shape[1] // 4
    h = b[:, :2 * c].contiguous().view(torch.float16).float()
    l8 = b[:, 2 * c:].reshape(rows, c // 8, 2, 8)[:, :, 1].contiguous().view(torch.float8_e4m3fn).float()
    return h + l8.reshape(rows, c) / LO_SCALE


def emulate_gemm(a_bytes, w_rows, inv_scale):
    """What the tensor core sums for one tap: a_bytes uint8 [rows, 4K] (encode_rows), w_rows the bf16-typed
    [cout, 2K] matrix of engine.pack_f16f8, inv_scale [cout] = 1 / S_n.  fp16 pass: hi x Wh' over K; e4m3 pass: the
    2K lo BYTES in storage order (this is what pins the [x8|l8] vs [Wl8|W8] pairing).  Both land in one accumulator
    (= S_n x result); products and sums in float64."""
    k = a_bytes.shape[1] // 4
    wb = w_rows.contiguous().view(torch.uint8).view(w_rows.shape[0], 4 * k)
    ah = a_bytes[:, :2 * k].contiguous().view(torch.float16).double()
    wh = wb[:, :2 * k].contiguous().view(torch.float16).double()
    al = a_bytes[:, 2 * k:].contiguous().view(torch.float8_e4m3fn).double()
    wl = wb[:, 2 * k:].contiguous().view(torch.float8_e4m3fn).double()
    return (ah @ wh.t() + al @ wl.t()) * inv_scale.double().view(1, -1)
