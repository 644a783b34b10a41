"""Execution plans over the padded-rows activation layout (see include/iou_b200.h).

An ``Engine`` turns a state_dict of the reference's modules into a flat list of
C-ABI launches: every convolution becomes one ``iou_conv_run`` of the tcgen05
tap-GEMM kernel with BatchNorm / bias / ReLU / residual folded into its
epilogue; layout kernels (stem im2col, max pool, stride-2 phase split) sit
between them.  Buffers and TMA plans are created once per input shape, so a whole
forward is a fixed launch sequence that can be captured in a CUDA graph.

What each builder mirrors in the reference:
  add_backbone : ResNet.forward            mmdet/models/backbones/resnet.py:224-267,507-518
  add_fpn      : FPN.forward               mmdet/models/necks/fpn.py:97-136
  add_head     : IoUawareRetinaHead.forward_single (all levels at once)
                                           mmdet/models/anchor_heads/iou_aware_retina_head.py:171-219
"""
import ctypes
import os

import torch

from . import lib as L

TILE_M = 128
STAGE_BLOCKS = {50: (3, 4, 6, 3), 101: (3, 4, 23, 3), 152: (3, 8, 36, 3)}
TAPS_1X1 = [(0, 0, 0)]
TAPS_3X3 = [(0, r - 1, s - 1) for r in range(3) for s in range(3)]                    # stride 1, pad 1
TAPS_3X3_S2 = [((r & 1) * 2 + (s & 1), r >> 1, s >> 1) for r in range(3) for s in range(3)]  # on phase maps
TAPS_1X1_S2 = [(3, 0, 0)]
TAPS_STEM = [(0, r - 2, 0) for r in range(4)]
TAPS_STEM_V = [(0, 0, s - 2) for s in range(4)]                                        # stem, vertical pack: dx = -2..1                                          # stem: dy = -2..1                                                             # phase (1,1)


def _round_up(x, m):
    return (x + m - 1) // m * m


def seg_tiles(n, h, w):
    """128-row tiles iou_conv_run computes for one segment: up to the last interior pixel (the trailing border rows
    stay zero from the allocation)."""
    return max(1, -(-(n * (h + 2) * (w + 2) - (w + 2) - 1) // TILE_M))


class FlatMap(object):
    """A bf16 [rows][2*c] padded-rows buffer holding one or more (n, h, w) segments."""

    def __init__(self, segs, c, device, tensor=None, ptr=None, rows=None):
        self.c = c
        self.segs = []          # (row_start, n, h, w)
        off = 0
        for (n, h, w) in segs:
            self.segs.append((off, n, h, w))
            off += _round_up(n * (h + 2) * (w + 2), TILE_M)
        if tensor is None and ptr is None:
            # zeros, not empty: the conv engine never writes the border rows behind a segment's last interior pixel
            tensor = torch.zeros(off, 2 * c, dtype=torch.bfloat16, device=device)
        self.tensor = tensor
        self.ptr = tensor.data_ptr() if ptr is None else ptr
        self.rows = off if rows is None else rows

    def view(self, s):
        """Single-segment view starting at segment s (rows run to the end of the parent)."""
        rs, n, h, w = self.segs[s]
        m = FlatMap([(n, h, w)], self.c, None, tensor=self.tensor, ptr=self.ptr + rs * 2 * self.c * 2,
                    rows=self.rows - rs)
        return m


def split_hi_lo(x):
    """fp32 -> (bf16 hi, bf16 lo) with hi + lo == x to ~2^-17 relative."""
    hi = x.to(torch.bfloat16)
    lo = (x - hi.float()).to(torch.bfloat16)
    return hi, lo


def _hi_lo_rows(t):
    """fp32 [taps, cout_pad, K] -> the C ABI's bf16 weight matrix [taps*cout_pad][2*K] (hi | lo).  The fp32 source
    rides along as ``.src32`` so that Engine.conv can repack it for the fp16 + e4m3 scheme (passes == 2)."""
    hi, lo = split_hi_lo(t)
    out = torch.cat([hi, lo], dim=2).reshape(t.shape[0] * t.shape[1], 2 * t.shape[2]).contiguous()
    out.src32 = t
    return out


F8_LO_SCALE = 2048.0      # 2^11, csrc/split_fmt.cuh


def pack_f16f8(t):
    """fp32 [taps, cout_pad, K] -> (weight matrix for passes == 2, per-channel 1 / S_n).

    Row layout mirrors the activations (csrc/split_fmt.cuh): [Wh': K x fp16][per 8 channels: Wl8 x 8 | W8 x 8].
    With s_n the power of two per OUTPUT channel that puts the row maximum of |w| * s_n in (8, 16] and
    S_n = 2^11 * s_n:  Wh' = fp16(w * S_n) (<= 32768),  Wl8 = e4m3(w * S_n - Wh'),  W8 = e4m3(w * s_n).
    The activation's fp16 hi meets Wh' in the fp16 MMA and its bytes [x8 x 8 | l8 x 8] (l8 = (x - hi) * 2^11) meet
    [Wl8 x 8 | W8 x 8] in the e4m3 MMA: all three products carry the factor S_n, so they add up in ONE fp32
    accumulator = S_n * (hi*w + x*wl + xl*w), which the epilogue multiplies by the returned 1 / S_n.
    Returned as a bf16-typed [taps*cout_pad][2*K] tensor (same bytes per row as hi | lo)."""
    taps, co, k = t.shape
    assert k % 8 == 0, k
    t = t.float()
    m = t.abs().amax(dim=(0, 2))
    s = torch.where(m > 0, torch.exp2(torch.floor(torch.log2(8.0 / m.clamp_min(1e-38))) + 1.0), torch.ones_like(m))
    s = s.clamp(2.0 ** -60, 2.0 ** 60)
    sv = s.view(1, co, 1)
    ts = t * (F8_LO_SCALE * sv)                                    # exact: powers of two
    wh = ts.clamp(-65504.0, 65504.0).to(torch.float16)
    wl8 = (ts - wh.float()).clamp(-448.0, 448.0).to(torch.float8_e4m3fn).view(torch.uint8)
    w8 = (t * sv).clamp(-448.0, 448.0).to(torch.float8_e4m3fn).view(torch.uint8)
    lo = torch.stack([wl8.view(taps, co, k // 8, 8), w8.view(taps, co, k // 8, 8)], dim=3).reshape(taps, co, 2 * k)
    hi = wh.contiguous().view(torch.uint8).view(taps, co, 2 * k)
    rows = torch.cat([hi, lo], dim=2).contiguous().view(torch.bfloat16).reshape(taps * co, 2 * k)
    return rows, (1.0 / (F8_LO_SCALE * s)).float().contiguous()


def unpack_f16f8_rows(rows, k):
    """Inverse view of pack_f16f8 for tests: -> (Wh' fp32, Wl8 fp32, W8 fp32), each [rows, K]."""
    b = rows.contiguous().view(torch.uint8).view(rows.shape[0], 4 * k)
    wh = b[:, :2 * k].contiguous().view(torch.float16).float()
    lo = b[:, 2 * k:].reshape(rows.shape[0], k // 8, 2, 8)
    wl8 = lo[:, :, 0].contiguous().view(torch.float8_e4m3fn).float().reshape(rows.shape[0], k)
    w8 = lo[:, :, 1].contiguous().view(torch.float8_e4m3fn).float().reshape(rows.shape[0], k)
    return wh, wl8, w8


def pack_weight(w, cout_pad, kpad=None):
    """(Cout, Cin, kh, kw) fp32 -> bf16 [kh*kw*cout_pad][2*Cin'] (hi | lo), tap-major, K contiguous."""
    cout, cin, kh, kw = w.shape
    t = w.permute(2, 3, 0, 1).reshape(kh * kw, cout, cin).float()
    if kpad is not None and kpad != cin:
        t = torch.nn.functional.pad(t, (0, kpad - cin))
        cin = kpad
    if cout_pad != cout:
        t = torch.nn.functional.pad(t, (0, 0, 0, cout_pad - cout))
    return _hi_lo_rows(t)


def pack_weight_stem(w):
    """(64, 3, 7, 7) stem weight -> [4 taps * 64][2*64] for the packed space-to-depth layout of
    iou_stem_pack: tap r (dy = r-2), K index kk = j*16 + (py*2+px)*3 + ch  <->  w[o, ch, 2r+py-1, 2j+px-1]."""
    cout = w.shape[0]
    t = torch.zeros(4, cout, 64, dtype=torch.float32, device=w.device)
    for r in range(4):
        for j in range(4):
            for py in range(2):
                for px in range(2):
                    ky, kx = 2 * r + py - 1, 2 * j + px - 1
                    if 0 <= ky <= 6 and 0 <= kx <= 6:
                        k0 = j * 16 + (py * 2 + px) * 3
                        t[r, :, k0:k0 + 3] = w[:, :, ky, kx].float()
    return _hi_lo_rows(t)


def pack_weight_stem_v(w):
    """(64, 3, 7, 7) stem weight -> [4 taps * 64][2*64] for iou_stem_pack_v: tap s (dx = s-2), K index
    kk = j*16 + (py*2+px)*3 + ch  <->  w[o, ch, 2j+py-1, 2s+px-1] (j = vertical neighbour y'-2+j)."""
    cout = w.shape[0]
    t = torch.zeros(4, cout, 64, dtype=torch.float32, device=w.device)
    for s_ in range(4):
        for j in range(4):
            for py in range(2):
                for px in range(2):
                    ky, kx = 2 * j + py - 1, 2 * s_ + px - 1
                    if 0 <= ky <= 6 and 0 <= kx <= 6:
                        k0 = j * 16 + (py * 2 + px) * 3
                        t[s_, :, k0:k0 + 3] = w[:, :, ky, kx].float()
    return _hi_lo_rows(t)


def pack_weight_grouped(w, groups):
    """Grouped (Cout, Cout/groups, kh, kw) fp32 -> block-diagonal bf16 [kh*kw*Cout][2*64]: row o holds the
    64 input channels of its own 64-channel block (zeros outside o's group), hi | lo."""
    cout, cg, kh, kw = w.shape
    assert cout % 64 == 0 and 64 % cg == 0 and cout // groups == cg, (w.shape, groups)
    t = torch.zeros(kh * kw, cout, 64, dtype=torch.float32, device=w.device)
    wt = w.permute(2, 3, 0, 1).reshape(kh * kw, cout, cg).float()
    o = torch.arange(cout, device=w.device)
    base = (o // cg) * cg - (o // 64) * 64                     # first local input channel of o's group
    idx = (base[:, None] + torch.arange(cg, device=w.device)[None, :])          # (cout, cg)
    t.scatter_(2, idx[None].expand(kh * kw, cout, cg), wt)
    return _hi_lo_rows(t)


def bn_fold(sd, prefix, eps=1e-5):
    """Eval-mode BN as per-channel (scale, shift): y = conv*scale + shift."""
    scale = sd[prefix + ".weight"].float() / torch.sqrt(sd[prefix + ".running_var"].float() + eps)
    shift = sd[prefix + ".bias"].float() - sd[prefix + ".running_mean"].float() * scale
    return scale.contiguous(), shift.contiguous()


NUM_SMS = 148


def fold_scale(w, scale):
    """Fold the per-output-channel BN scale into the conv weight (y = conv(x; w*s) + shift), so that the
    GEMM epilogue only adds the shift."""
    return w.float() * scale.to(w.device).float().view(-1, 1, 1, 1)


def pick_block_n(cout, m_tiles=None, max_bn=256, pair_min_tiles=None):
    """(block_n, cout_pad) for the tap-GEMM.  With the number of 128-row tiles known, the N tile is
    chosen to minimise waves x per-tile cost on 148 SMs (small maps get narrower tiles so that more
    CTAs share the K loop; big maps keep N=256 so the A tile is loaded once).  Maps with at least
    `pair_min_tiles` tiles run as CTA pairs: the unit of work is then two 128-row tiles on one of 74 pairs."""
    if cout % 64 != 0:                       # head outputs: 720 -> 3 x 240, 45 -> 48
        if cout % 240 == 0:
            return 240, cout
        pad = _round_up(cout, 16)
        return (pad, pad) if pad <= 256 else (256, _round_up(cout, 256))
    import os
    max_bn = min(max_bn, int(os.environ.get("IOU_MAX_BN", "256")))     # tuning knob for experiments
    cands = [bn for bn in (256, 128, 64) if cout % bn == 0 and bn <= max_bn]
    if m_tiles is None:
        return cands[0], cout
    pair = pair_min_tiles is not None and m_tiles >= pair_min_tiles

    def cost(bn):
        n_tiles = cout // bn
        if pair:
            waves = -(-(-(-m_tiles // 2) * n_tiles) // (NUM_SMS // 2))
        else:
            waves = -(-(m_tiles * n_tiles) // NUM_SMS)
        return waves * (bn + 64)
    return min(cands, key=cost), cout


class Engine(object):
    def __init__(self, device, passes=3):
        self.device = torch.device(device)
        self.passes = passes        # 3: bf16 hi|lo x3 (default), 2: fp16 + e4m3 corrections (csrc/split_fmt.cuh), 1: bf16
        self.fmt = 1 if passes == 2 else 0
        self.ops = []            # (name, callable(stream_ptr))
        self.plans = []
        self.keep = []           # tensors that must outlive the plan
        self.maps = []           # every padded-rows buffer of this engine (range_report)
        self.side = []           # (fork_at, first, last): ops[first..last] run on a side stream forked before op fork_at
        self._side_stream = None
        self.flops = 0.0
        self.op_flops = {}
        self.epi_warps = {}                     # conv name -> 8 / 12 (which kernel variant its plan launches)
        import os
        self.two_cta = os.environ.get("IOU_TWO_CTA", "1") != "0"
        self.pair_min_bn = int(os.environ.get("IOU_PAIR_MIN_BN", "48"))   # 48: the reg+iou output conv as a CTA pair (-28 %)
        self.pair_min_tiles = int(os.environ.get("IOU_PAIR_MIN_TILES", "32"))
        self.stem_vertical = os.environ.get("IOU_STEM_VERTICAL", "1") != "0"
        self.res_bn256 = os.environ.get("IOU_RES_BN256", "1") != "0"        # N = 256 tiles for residual convs in pair mode
        self.pair_diag = os.environ.get("IOU_PAIR_DIAG", "1") != "0"       # grouped (block-diagonal) convs as CTA pairs
        self.lib = L.load()

    # ------------------------------------------------------------------ primitive ops
    def new_map(self, segs, c):
        """Allocates a padded-rows buffer owned by this engine (plans only hold raw pointers)."""
        m = FlatMap(segs, c, self.device)
        self.keep.append(m.tensor)
        self.maps.append(m)
        return m

    def _dev(self, t):
        t = t.detach().to(self.device, torch.float32).contiguous()
        self.keep.append(t)
        return t

    def conv(self, name, srcs, taps, weight, cin, cout, out=None, scale=None, shift=None, relu=False,
             residual=None, res_mode=L.RES_NONE, dense_out=None, dense_out2=None, dense_split=0,
             segs_from=None, diag_k=False, true_flops_scale=1.0, two_cta=None, force_bn=None,
             phase_outs=None, phase_only=False, k_split=1, hold=False, chain=False, wide=0,
             group_max_out=None, group_max_cols=0):
        """srcs: list[FlatMap] (same geometry); out: FlatMap or None (dense).  Returns out.
        phase_outs: 4 FlatMaps / None from new_phase_maps(): the epilogue also writes the stride-2 phase split of the
        output (what iou_phase_split would produce from it); phase_only: nothing else is written (returns None).
        hold=True keeps the launch back so that the NEXT conv(chain=True) -- a plain 1x1 conv reading this conv's output --
        can be chained into the same launch (iou_conv_chain_plan_create); if the pair does not qualify both run as usual.
        group_max_out / group_max_cols: dense outputs only -- per pixel and group of `group_max_cols` output channels, two
        partial maxima (fp32 [n*h*w][cout/cols][2] per segment; iou_conv_desc.group_max_cols).
        wide: iou_conv_desc.wide (0 = library default, 1 = 12 epilogue warps, -1 = 8).
        k_split = S > 1: the sources hold S * cin channels and `cout` = S * (real cout) output channels are the S partial
        sums over the channel slices (weight rows packed to match, see split_k_weight); sum_groups() adds them."""
        geo = segs_from or srcs[0]
        m_tiles = sum(seg_tiles(n, h, w) for (_, n, h, w) in geo.segs)
        # a same-geometry residual needs its TMA staging ring in shared memory: N tile <= 128, or 256 when the conv
        # runs as a CTA pair (half a B tile per CTA); plan_create decides, so 256 is tried first and 128 is the retry
        pair_aware = self.pair_min_tiles if (self.two_cta and two_cta is None and
                                             os.environ.get('IOU_PAIR_AWARE', '1') == '1') else None
        res_wide = (res_mode == L.RES_SAME and self.res_bn256 and self.two_cta and two_cta is None and
                    m_tiles >= self.pair_min_tiles and cout % 256 == 0)
        block_n, cout_pad = pick_block_n(cout, m_tiles, max_bn=128 if (res_mode == L.RES_SAME and not res_wide) else 256,
                                         pair_min_tiles=pair_aware)
        # passes == 2: main + correction accumulators of an N > 128 tile fill all 512 TMEM columns, so the epilogue of
        # tile i no longer overlaps the MMAs of tile i+1.  Convs with a short K loop (1x1 convs, where the epilogue IS
        # the work) keep two accumulator stages with N = 128 instead; deep K loops (3x3 towers) amortise it
        if (self.passes == 2 and block_n > 128 and cout % 128 == 0 and not diag_k and
                len(taps) * (cin // 64) <= int(os.environ.get("IOU_F8_SHALLOW", "0"))):
            block_n, cout_pad = 128, cout
        retry_bn = 128 if (res_wide and block_n == 256) else None
        if force_bn is not None:                 # (block_n, cout_pad) chosen by the caller (weight packed to match)
            block_n, cout_pad = force_bn
            retry_bn = None
        if diag_k:
            block_n, cout_pad = 64, cout
        d = L.ConvDesc()
        d.cin, d.cout, d.cout_pad, d.block_n = cin, cout, cout_pad, block_n
        d.num_taps = len(taps)
        for i, (s, dy, dx) in enumerate(taps):
            d.tap_src[i], d.tap_dy[i], d.tap_dx[i] = s, dy, dx
        d.num_src = len(srcs)
        for i, m in enumerate(srcs):
            # sources may have fewer channels than cin (= the K of the weight rows): their taps contract m.c channels
            assert (m.c == cin * k_split) if k_split > 1 else (m.c <= cin and (m.c == cin or not diag_k)), (name, m.c, cin)
            d.src[i] = m.ptr
            d.src_cin[i] = m.c
        d.src_rows = min(m.rows for m in srcs)
        kdim = 64 if diag_k else cin
        assert weight.shape == (len(taps) * cout_pad, 2 * kdim), (name, weight.shape, len(taps), cout_pad, cin)
        corr_scale = None
        if self.passes == 2:                     # repack for the fp16 + e4m3 scheme, from the fp32 source if it rode along
            src32 = getattr(weight, "src32", None)
            if src32 is None:
                src32 = (weight[:, :kdim].float() + weight[:, kdim:].float())
            assert scale is None, "passes == 2 uses `scale` for 1 / S_n (fold a BN scale into the weights)"
            weight, corr_scale = pack_f16f8(src32.reshape(len(taps), cout_pad, kdim))
        elif hasattr(weight, "src32"):
            del weight.src32
        wp = weight.to(self.device)
        self.keep.append(wp)
        d.diag_k = int(diag_k)
        # big maps with a wide N tile run as CTA pairs (cta_group::2): half the B traffic, deeper pipeline
        d.two_cta = int(self.two_cta and (not diag_k or self.pair_diag) and block_n % 16 == 0 and block_n >= self.pair_min_bn
                        and m_tiles >= self.pair_min_tiles) if two_cta is None else int(two_cta)
        d.weight = wp.data_ptr()

        def padc(v):
            v = self._dev(v)
            if v.numel() != cout_pad:
                v = torch.nn.functional.pad(v, (0, cout_pad - v.numel())).contiguous()
                self.keep.append(v)
            return v
        if corr_scale is not None:
            d.scale = padc(corr_scale).data_ptr()
        elif scale is not None:
            d.scale = padc(scale).data_ptr()
        if shift is not None:
            d.shift = padc(shift).data_ptr()
        d.relu = int(relu)
        d.res_mode = res_mode
        if residual is not None:
            d.residual = residual.ptr
            d.res_rows = residual.rows
            for i, (rs, n, h, w) in enumerate(residual.segs):
                d.res_seg[i] = L.ConvSegment(rs, n, h, w)
        d.num_seg = len(geo.segs)
        for i, (rs, n, h, w) in enumerate(geo.segs):
            d.seg[i] = L.ConvSegment(rs, n, h, w)
        if phase_outs is not None:
            for i, m in enumerate(phase_outs):
                if m is not None:
                    assert m.c == cout
                    d.phase_out[i] = m.ptr
            d.phase_only = int(phase_only)
        if dense_out is None and phase_only:
            d.out_mode = L.OUT_PADDED
        elif dense_out is None:
            if out is None:
                out = self.new_map([(n, h, w) for (_, n, h, w) in geo.segs], cout)
            assert out.c == cout
            d.out_mode, d.out = L.OUT_PADDED, out.ptr
            d.out_rows = out.rows
        else:
            d.out_mode = L.OUT_DENSE
            d.dense_split = dense_split
            for i, t in enumerate(dense_out):
                d.out_dense[i] = t.data_ptr()
            if dense_out2 is not None:
                for i, t in enumerate(dense_out2):
                    d.out_dense2[i] = t.data_ptr()
            if group_max_out is not None:
                d.group_max_cols = int(group_max_cols)
                for i, t in enumerate(group_max_out):
                    d.group_max_out[i] = t.data_ptr()
        d.passes = self.passes
        d.k_split = int(k_split)
        d.wide = -1 if (hold or chain) else int(wide)      # the chained launch has its own (8-warp) epilogue
        plan = ctypes.c_void_p()
        rc = self.lib.iou_conv_plan_create(ctypes.byref(d), ctypes.byref(plan))
        if rc != 0 and retry_bn is not None:         # the wide residual tile did not fit next to the staging rings
            d.block_n = retry_bn
            rc = self.lib.iou_conv_plan_create(ctypes.byref(d), ctypes.byref(plan))
        L.check(rc)
        self.plans.append(plan)
        for m in ([out] if out is not None else []) + [m for m in (phase_outs or []) if m is not None]:
            if getattr(m, "label", None) is None:
                m.label = name
        self.epi_warps[name] = self.lib.iou_conv_plan_epilogue_warps(plan)
        f = self.lib.iou_conv_plan_flops(plan) * true_flops_scale
        self.flops += f
        lib = self.lib
        held, self._held = getattr(self, "_held", None), None
        if held is not None:
            h_name, h_plan, h_f = held
            merged = ctypes.c_void_p()
            rc = lib.iou_conv_chain_plan_create(h_plan, plan, ctypes.byref(merged)) if chain else -1
            if rc == 0:                          # one launch computes both convs (the first plan now owns the second)
                self.plans.remove(plan)
                name2 = h_name + "+" + name.split(".")[-1]
                self.op_flops[name2] = self.op_flops.get(name2, 0.0) + h_f + f
                self.ops.append((name2, lambda st, p=h_plan: L.check(lib.iou_conv_run(p, st))))
                self.chained = getattr(self, "chained", 0) + 1
                return out
            self.op_flops[h_name] = self.op_flops.get(h_name, 0.0) + h_f
            self.ops.append((h_name, lambda st, p=h_plan: L.check(lib.iou_conv_run(p, st))))
        if hold:
            self._held = (name, plan, f)
            return out
        self.op_flops[name] = self.op_flops.get(name, 0.0) + f
        self.ops.append((name, lambda st, p=plan: L.check(lib.iou_conv_run(p, st))))
        return out

    def sum_groups(self, name, part, out, groups, bias=None):
        """out (FlatMap, one segment, c channels) <- bias + sum of the `groups` channel groups of part (groups * c
        channels): the reduction behind conv(..., k_split=groups)."""
        (_, n, h, w), c = out.segs[0], out.c
        assert part.c == groups * c and part.segs[0][1:] == (n, h, w)
        b = self._dev(bias) if bias is not None else None
        lib, pp, op, bp, fmt = self.lib, part.ptr, out.ptr, (b.data_ptr() if b is not None else None), self.fmt
        self.ops.append((name, lambda st: L.check(lib.iou_sum_channel_groups(pp, n, h, w, c, groups, bp, op, fmt, st))))
        return out

    def new_phase_maps(self, n, h, w, c, mask=15):
        """Zero-initialised stride-2 phase maps of an (n, h, w) map with c channels, for conv(phase_outs=...)."""
        ho, wo = (h + 1) // 2, (w + 1) // 2
        return [self.new_map([(n, ho, wo)], c) if (mask >> i) & 1 else None for i in range(4)]

    def phase_split(self, name, src, mask=15, relu=False):
        """-> list of 4 FlatMaps (None where masked out) in the stride-2 output geometry; relu=True clamps the
        copied values at zero (FPN relu_before_extra_convs)."""
        assert len(src.segs) == 1
        _, n, h, w = src.segs[0]
        ho, wo = (h + 1) // 2, (w + 1) // 2
        outs = [self.new_map([(n, ho, wo)], src.c) if (mask >> i) & 1 else None for i in range(4)]
        arr = (ctypes.c_void_p * 4)(*[(o.ptr if o is not None else None) for o in outs])
        self.keep.append(arr)
        lib, c, sp, fmt = self.lib, src.c, src.ptr, self.fmt
        kmask = mask | (16 if relu else 0)
        self.ops.append((name, lambda st: L.check(lib.iou_phase_split_fmt(sp, n, c, h, w, arr, kmask, fmt, st))))
        return outs

    # ------------------------------------------------------------------ network builders
    def add_stem(self, sd, img, prefix="backbone."):
        """conv1 7x7/s2 + bn1 + relu + maxpool (resnet.py:508-511).  img: (N,3,H,W) fp32 cuda tensor.
        The 7x7/s2 conv runs as 4 taps of K=64 over the packed space-to-depth map (iou_stem_pack)."""
        n, _, h, w = img.shape
        ho, wo = (h - 1) // 2 + 1, (w - 1) // 2 + 1
        packed = self.new_map([(n, ho, wo)], 64)
        lib, ip, pp = self.lib, img.data_ptr(), packed.ptr
        self.keep.append(img)
        scale, shift = bn_fold(sd, prefix + "bn1")
        wf = fold_scale(sd[prefix + "conv1.weight"], scale)
        fmt = self.fmt
        if fmt != 0 and not self.stem_vertical:
            raise RuntimeError("passes == 2 needs the vertical stem pack (IOU_STEM_VERTICAL=1)")
        if self.stem_vertical:      # four dx taps sharing one A window (half the shared-memory fill of the stem conv)
            self.ops.append(("stem.pack", lambda st: L.check(lib.iou_stem_pack_v_fmt(ip, n, h, w, pp, fmt, st))))
            s1 = self.conv("stem.conv1", [packed], TAPS_STEM_V, pack_weight_stem_v(wf), 64, 64, shift=shift, relu=True,
                           true_flops_scale=147.0 / 256.0)
        else:
            self.ops.append(("stem.pack", lambda st: L.check(lib.iou_stem_pack(ip, n, h, w, pp, st))))
            s1 = self.conv("stem.conv1", [packed], TAPS_STEM, pack_weight_stem(wf), 64, 64, shift=shift, relu=True,
                           true_flops_scale=147.0 / 256.0)
        hp, wq = (ho - 1) // 2 + 1, (wo - 1) // 2 + 1
        x = self.new_map([(n, hp, wq)], 64)
        sp, xp = s1.ptr, x.ptr
        self.ops.append(("stem.maxpool", lambda st: L.check(lib.iou_maxpool3x3s2_fmt(sp, n, 64, ho, wo, xp, fmt, st))))
        return x

    def add_backbone(self, sd, img, depth=50, groups=1, prefix="backbone.", style="pytorch"):
        """ResNet (groups == 1) or ResNeXt (grouped 3x3 run as a block-diagonal tap-GEMM).  style 'pytorch' puts a
        block's stride in the 3x3 conv2, 'caffe' in the 1x1 conv1 (resnet.py:129-134)."""
        x = self.add_stem(sd, img, prefix)
        outs = []
        fuse = os.environ.get("IOU_FUSE_PHASE", "1") != "0"     # producers write the stride-2 phase maps themselves
        fuse_ds = os.environ.get("IOU_FUSE_DS", "1") != "0"     # conv3 + downsample of a stage's first block in one GEMM
        fuse_ph3 = os.environ.get("IOU_FUSE_PH3", "1") != "0"   # a stage's last conv3 also writes phase (1,1) of its output
        # conv3 of block b and conv1 of block b+1 (same stage) in ONE launch: conv1 reads x back from L2 (conv_chain.cu)
        chain = os.environ.get("IOU_CHAIN", "0") != "0" and self.passes == 2 and groups == 1 and style == "pytorch"
        chain_stages = [int(v) for v in os.environ.get("IOU_CHAIN_STAGES", "0,1,2").split(",") if v != ""]
        chained_in = False                                      # this block's conv1 completes a held conv3
        x_ph3 = None                                            # phase (1,1) of x, if its producer wrote it
        for s, nblocks in enumerate(STAGE_BLOCKS[depth]):
            planes = 64 * 2 ** s
            for b in range(nblocks):
                p = "%slayer%d.%d." % (prefix, s + 1, b)
                stride = 2 if (b == 0 and s > 0) else 1
                cin = x.c
                width = sd[p + "conv1.weight"].shape[0]          # == planes for ResNet (resnext.py:21-24)
                sc1, sh1 = bn_fold(sd, p + "bn1")
                w1 = pack_weight(fold_scale(sd[p + "conv1.weight"], sc1), width)
                xs = [None, None, None, x_ph3] if (stride == 2 and x_ph3 is not None) else None
                x_ph3 = None
                t1_ph = None
                if stride == 2 and style == "caffe":        # 1x1 stride 2 reads phase (1,1) of x; conv2 has stride 1
                    if xs is None:
                        xs = self.phase_split(p + "conv1.phase", x, mask=8)
                    t1 = self.conv(p + "conv1", [xs[3]] * 4, TAPS_1X1_S2, w1, cin, width, shift=sh1, relu=True)
                elif stride == 2 and fuse and width % 64 == 0:
                    # t1 is only read by the stride-2 conv2: conv1's epilogue writes its four phase maps directly
                    _, n_, h_, w_ = x.segs[0]
                    t1_ph = self.new_phase_maps(n_, h_, w_, width)
                    t1 = self.conv(p + "conv1", [x], TAPS_1X1, w1, cin, width, shift=sh1, relu=True,
                                   phase_outs=t1_ph, phase_only=True)
                else:
                    t1 = self.conv(p + "conv1", [x], TAPS_1X1, w1, cin, width, shift=sh1, relu=True, chain=chained_in,
                                   force_bn=((width, width) if (chained_in and width <= 256) else None))
                chained_in = False
                sc2, sh2 = bn_fold(sd, p + "bn2")
                if groups == 1:
                    w2, kw = pack_weight(fold_scale(sd[p + "conv2.weight"], sc2), width), {}
                else:
                    cg = sd[p + "conv2.weight"].shape[1]
                    w2 = pack_weight_grouped(fold_scale(sd[p + "conv2.weight"], sc2), groups)
                    kw = dict(diag_k=True, true_flops_scale=cg / 64.0)
                if stride == 2 and style != "caffe":
                    ph = t1_ph if t1_ph is not None else self.phase_split(p + "conv2.phase", t1)
                    t2 = self.conv(p + "conv2", ph, TAPS_3X3_S2, w2, width, width, shift=sh2,
                                   relu=True, **kw)
                else:
                    t2 = self.conv(p + "conv2", [t1], TAPS_3X3, w2, width, width, shift=sh2,
                                   relu=True, **kw)
                idt = x
                sc3, sh3 = bn_fold(sd, p + "bn3")
                ph3 = None
                if fuse and fuse_ph3 and b == nblocks - 1 and s + 1 < len(STAGE_BLOCKS[depth]):
                    # the next stage's first block reads phase (1,1) of this block's output (stride-2 1x1 convs)
                    _, n_, h_, w_ = t2.segs[0]
                    ph3 = self.new_phase_maps(n_, h_, w_, planes * 4, mask=8)
                if (p + "downsample.0.weight") in sd and fuse_ds:
                    # relu(bn3(conv3(t2)) + bn_d(downsample(x))) as ONE K-concatenated GEMM: tap 0 contracts t2's `width`
                    # channels, tap 1 the `cin` channels of x (stride 2: its phase (1,1) map) into the same accumulator --
                    # the identity branch is never written to or read back from HBM
                    scd, shd = bn_fold(sd, p + "downsample.1")
                    if stride == 2 and xs is None:
                        xs = self.phase_split(p + "downsample.phase", x, mask=8)
                    w3 = fold_scale(sd[p + "conv3.weight"], sc3)
                    wdn = fold_scale(sd[p + "downsample.0.weight"], scd)
                    kmax = max(width, cin)
                    wt = torch.zeros(2, planes * 4, kmax, dtype=torch.float32, device=w3.device)
                    wt[0, :, :width] = w3[:, :, 0, 0]
                    wt[1, :, :cin] = wdn[:, :, 0, 0]
                    x = self.conv(p + "conv3+downsample", [t2, xs[3] if stride == 2 else x], [(0, 0, 0), (1, 0, 0)],
                                  _hi_lo_rows(wt), kmax, planes * 4, shift=sh3 + shd, relu=True, phase_outs=ph3)
                    x_ph3 = ph3[3] if ph3 is not None else None
                    continue
                if (p + "downsample.0.weight") in sd:
                    scd, shd = bn_fold(sd, p + "downsample.1")
                    wd = pack_weight(fold_scale(sd[p + "downsample.0.weight"], scd), planes * 4)
                    if stride == 2:
                        if xs is None:
                            xs = self.phase_split(p + "downsample.phase", x, mask=8)
                        srcs = [xs[3], xs[3], xs[3], xs[3]]
                        idt = self.conv(p + "downsample", srcs, TAPS_1X1_S2, wd, cin, planes * 4,
                                        shift=shd)
                    else:
                        idt = self.conv(p + "downsample", [x], TAPS_1X1, wd, cin, planes * 4,
                                        shift=shd)
                hold = chain and s in chain_stages and ph3 is None and b + 1 < nblocks and idt is x
                x = self.conv(p + "conv3", [t2], TAPS_1X1, pack_weight(fold_scale(sd[p + "conv3.weight"], sc3), planes * 4),
                              width, planes * 4, shift=sh3, relu=True, residual=idt,
                              res_mode=L.RES_SAME, phase_outs=ph3, hold=hold)
                chained_in = hold
                x_ph3 = ph3[3] if ph3 is not None else None
            outs.append(x)
        return outs

    def add_fpn(self, sd, feats, prefix="neck.", start_level=1, num_outs=5, out_channels=256,
                extra_convs_on_inputs=True, relu_before_extra_convs=False):
        """feats: list of FlatMaps C2..C5.  Returns one multi-segment FlatMap (P3..P7).  The extra stride-2 levels
        read C5 (RetinaNet configs) or the last FPN output (FCOS, extra_convs_on_inputs=False), optionally through
        a ReLU (relu_before_extra_convs; fpn.py:117-128)."""
        used = feats[start_level:]
        nl = len(used)
        fpn_first = len(self.ops)
        lat = [None] * nl
        for i in range(nl - 1, -1, -1):
            kp = "%slateral_convs.%d.conv." % (prefix, i)
            res = lat[i + 1] if i + 1 < nl else None
            if res is not None:
                # F.interpolate(scale_factor=2) + in-place add (fpn.py:108-110) needs exact 2x sizes; the
                # reference raises on a mismatch (e.g. an input not padded to a multiple of 32), so do we
                (_, _, hf, wf), (_, _, hc, wc) = used[i].segs[0], res.segs[0]
                if (hf, wf) != (2 * hc, 2 * wc):
                    raise RuntimeError("FPN top-down add: level sizes %dx%d and %dx%d are not 2x apart "
                                       "(pad the input to a multiple of 32)" % (hf, wf, hc, wc))
            lat[i] = self.conv(kp[:-1], [used[i]], TAPS_1X1, pack_weight(sd[kp + "weight"], out_channels),
                               used[i].c, out_channels, shift=sd[kp + "bias"], residual=res,
                               res_mode=L.RES_UPSAMPLE2 if res is not None else L.RES_NONE)
        geos = [m.segs[0][1:] for m in lat]
        n, h, w = geos[-1]
        for _ in range(nl, num_outs):
            h, w = (h + 1) // 2, (w + 1) // 2
            geos.append((n, h, w))
        F = self.new_map(geos, out_channels)
        for i in range(nl):
            kp = "%sfpn_convs.%d.conv." % (prefix, i)
            self.conv(kp[:-1], [lat[i]], TAPS_3X3, pack_weight(sd[kp + "weight"], out_channels),
                      out_channels, out_channels, out=F.view(i), shift=sd[kp + "bias"])
        src = feats[-1] if extra_convs_on_inputs else F.view(nl - 1)     # fpn.py:118-122
        first_extra = len(self.ops)
        for i in range(nl, num_outs):
            kp = "%sfpn_convs.%d.conv." % (prefix, i)
            # fpn.py:123-128: ReLU only in front of the extra convs AFTER the first one
            ph = self.phase_split(kp + "phase", src, relu=(relu_before_extra_convs and i > nl))
            ks = int(os.environ.get("IOU_P6_KSPLIT", "4"))
            _, n_, h_, w_ = F.segs[i]
            if ks > 1 and src.c % (64 * ks) == 0 and src.c >= 1024 and seg_tiles(n_, h_, w_) < NUM_SMS // 2:
                # few output rows, long K (P6 on C5: 22 row tiles x K = 18 432): S times more work items of 1/S the
                # length through N-concatenated partial sums, added (+ bias) by one small kernel
                wk = sd[kp + "weight"]
                cs = src.c // ks
                wsp = torch.cat([wk[:, j * cs:(j + 1) * cs] for j in range(ks)], dim=0)     # (S*cout, cin/S, 3, 3)
                part = self.new_map([(n_, h_, w_)], ks * out_channels)
                self.conv(kp[:-1], ph, TAPS_3X3_S2, pack_weight(wsp, ks * out_channels), cs, ks * out_channels,
                          out=part, k_split=ks, two_cta=(True if os.environ.get("IOU_P6_PAIR", "1") != "0" else None))
                self.sum_groups(kp + "sum", part, F.view(i), ks, bias=sd[kp + "bias"])
            else:
                self.conv(kp[:-1], ph, TAPS_3X3_S2, pack_weight(sd[kp + "weight"], out_channels), src.c,
                          out_channels, out=F.view(i), shift=sd[kp + "bias"])
            src = F.view(i)
        if extra_convs_on_inputs and num_outs > nl and os.environ.get("IOU_FPN_SIDE", "0") != "0":
            # P6 / P7 only read C5 (fpn.py:118-128): their launches (few tiles, long K) run on a side stream next to the
            # laterals and the P3..P5 output convs, joined before the head
            self.side.append((fpn_first, first_extra, len(self.ops) - 1))
        return F

    def add_head(self, sd, F, prefix="bbox_head.", stacked=4, num_anchors=9, num_classes=80, with_iou=True):
        """All levels in one launch per conv.  Returns (cls, reg, iou) lists of NHWC fp32 tensors
        exposed with the reference's logical shape (N, A*C, H, W)."""
        c = r = F
        fc = sd[prefix + "cls_convs.0.conv.weight"].shape[0]
        for i in range(stacked):
            kc, kr = "%scls_convs.%d.conv." % (prefix, i), "%sreg_convs.%d.conv." % (prefix, i)
            c = self.conv(kc[:-1], [c], TAPS_3X3, pack_weight(sd[kc + "weight"], fc), c.c, fc,
                          shift=sd[kc + "bias"], relu=True)
            r = self.conv(kr[:-1], [r], TAPS_3X3, pack_weight(sd[kr + "weight"], fc), r.c, fc,
                          shift=sd[kr + "bias"], relu=True)
        ncls, nreg, niou = num_anchors * num_classes, num_anchors * 4, num_anchors
        cls_out = [torch.empty(n, h, w, ncls, dtype=torch.float32, device=self.device) for (_, n, h, w) in F.segs]
        reg_out = [torch.empty(n, h, w, nreg, dtype=torch.float32, device=self.device) for (_, n, h, w) in F.segs]
        iou_out = [torch.empty(n, h, w, niou, dtype=torch.float32, device=self.device) for (_, n, h, w) in F.segs]
        bn_c, pad_c = pick_block_n(ncls)
        force = None
        cls_bn = int(os.environ.get("IOU_F8_CLS_BN", "0"))
        if self.passes == 2 and cls_bn:          # two accumulator stages (N <= 128) at the price of padded columns
            pad_c = _round_up(ncls, cls_bn)
            force = (cls_bn, pad_c)
        # the per-anchor max class logit leaves the retina_cls epilogue (two partial maxima per anchor), so that the top-k
        # pre-selection of get_bboxes does not re-read the class maps (iou_get_bboxes_premax)
        bn_used = force[0] if force is not None else bn_c
        self.cls_max2 = None
        if (os.environ.get("IOU_FUSE_MAX", "1") != "0" and num_classes >= 32 and num_classes % 16 == 0 and
                bn_used % num_classes == 0):
            self.cls_max2 = [torch.empty(n, h, w, num_anchors, 2, dtype=torch.float32, device=self.device)
                             for (_, n, h, w) in F.segs]
            self.keep += self.cls_max2
        self.conv(prefix + "retina_cls", [c], TAPS_3X3, pack_weight(sd[prefix + "retina_cls.weight"], pad_c),
                  fc, ncls, shift=sd[prefix + "retina_cls.bias"], dense_out=cls_out, force_bn=force,
                  group_max_out=self.cls_max2, group_max_cols=num_classes if self.cls_max2 is not None else 0)
        if with_iou:
            # retina_reg and retina_iou read the same feature (shared_conv=4, :198-204): one GEMM, split store
            w_ri = torch.cat([sd[prefix + "retina_reg.weight"], sd[prefix + "retina_iou.weight"]], dim=0)
            b_ri = torch.cat([sd[prefix + "retina_reg.bias"], sd[prefix + "retina_iou.bias"]], dim=0)
            _, pad_ri = pick_block_n(nreg + niou)
            self.conv(prefix + "retina_reg+iou", [r], TAPS_3X3, pack_weight(w_ri, pad_ri), fc, nreg + niou,
                      shift=b_ri, dense_out=reg_out, dense_out2=iou_out, dense_split=nreg)
        else:                                 # plain RetinaHead (retina_head.py:78-96)
            _, pad_r = pick_block_n(nreg)
            self.conv(prefix + "retina_reg", [r], TAPS_3X3, pack_weight(sd[prefix + "retina_reg.weight"], pad_r),
                      fc, nreg, shift=sd[prefix + "retina_reg.bias"], dense_out=reg_out)
        self.keep += cls_out + reg_out + iou_out
        as_nchw = lambda ts: [t.permute(0, 3, 1, 2) for t in ts]
        return as_nchw(cls_out), as_nchw(reg_out), (as_nchw(iou_out) if with_iou else None)

    def group_norm(self, name, m, gamma, beta, groups, eps=1e-5, relu=True):
        """In-place GroupNorm(+ReLU) of every segment of FlatMap m (ConvModule with norm_cfg type 'GN')."""
        g, b = self._dev(gamma), self._dev(beta)
        segs = (L.ConvSegment * len(m.segs))()
        for i, (rs, n, h, w) in enumerate(m.segs):
            segs[i].row_start, segs[i].n_img, segs[i].h, segs[i].w = rs, n, h, w
        nbytes = self.lib.iou_group_norm_workspace_bytes(sum(n for (_, n, _, _) in m.segs), groups)
        ws = torch.empty(max(nbytes, 8), dtype=torch.uint8, device=self.device)
        self.keep += [ws, segs]
        lib, mp, c, ns, gp, bp, wp = self.lib, m.ptr, m.c, len(m.segs), g.data_ptr(), b.data_ptr(), ws.data_ptr()
        fmt = self.fmt
        self.ops.append((name, lambda st: L.check(lib.iou_group_norm_relu_fmt(mp, c, ns, segs, groups, gp, bp, float(eps),
                                                                              int(relu), wp, nbytes, fmt, st))))
        self.extra_launches = getattr(self, "extra_launches", 0) + 2      # memset + 2 kernels behind one op
        return m

    def scale_exp(self, name, t, scale):
        """t <- exp(t * scale) on a dense fp32 tensor owned by this engine."""
        lib, tp, n, sc = self.lib, t.data_ptr(), t.numel(), float(scale)
        self.ops.append((name, lambda st: L.check(lib.iou_scale_exp(tp, n, sc, st))))

    def add_fcos_head(self, sd, F, prefix="bbox_head.", stacked=4, num_classes=80, groups=32, eps=1e-5):
        """IoUawareFCOSHead.forward_single for all levels at once (iou_aware_fcos_head.py:92-113): towers of
        conv(no bias) -> GN -> ReLU; fcos_cls + fcos_centerness share cls_feat (one GEMM, split store), fcos_reg +
        fcos_iou share reg_feat; bbox_pred = exp(scale_l * fcos_reg).  Returns (cls, bbox_pred, centerness, iou)."""
        c = r = F
        fc = sd[prefix + "cls_convs.0.conv.weight"].shape[0]
        for i in range(stacked):
            for tower in ("cls", "reg"):
                k = "%s%s_convs.%d." % (prefix, tower, i)
                src = c if tower == "cls" else r
                out = self.conv(k + "conv", [src], TAPS_3X3, pack_weight(sd[k + "conv.weight"], fc), src.c, fc)
                self.group_norm(k + "gn", out, sd[k + "gn.weight"], sd[k + "gn.bias"], groups, eps, relu=True)
                if tower == "cls":
                    c = out
                else:
                    r = out
        mk = lambda ch: [torch.empty(n, h, w, ch, dtype=torch.float32, device=self.device) for (_, n, h, w) in F.segs]
        cls_out, cen_out, reg_out, iou_out = mk(num_classes), mk(1), mk(4), mk(1)
        w_cc = torch.cat([sd[prefix + "fcos_cls.weight"], sd[prefix + "fcos_centerness.weight"]], dim=0)
        b_cc = torch.cat([sd[prefix + "fcos_cls.bias"], sd[prefix + "fcos_centerness.bias"]], dim=0)
        _, pad_cc = pick_block_n(num_classes + 1)
        self.conv(prefix + "fcos_cls+centerness", [c], TAPS_3X3, pack_weight(w_cc, pad_cc), fc, num_classes + 1,
                  shift=b_cc, dense_out=cls_out, dense_out2=cen_out, dense_split=num_classes)
        w_ri = torch.cat([sd[prefix + "fcos_reg.weight"], sd[prefix + "fcos_iou.weight"]], dim=0)
        b_ri = torch.cat([sd[prefix + "fcos_reg.bias"], sd[prefix + "fcos_iou.bias"]], dim=0)
        _, pad_ri = pick_block_n(5)
        self.conv(prefix + "fcos_reg+iou", [r], TAPS_3X3, pack_weight(w_ri, pad_ri), fc, 5, shift=b_ri,
                  dense_out=reg_out, dense_out2=iou_out, dense_split=4)
        for l, t in enumerate(reg_out):
            self.scale_exp("%sscales.%d.exp" % (prefix, l), t, float(sd["%sscales.%d.scale" % (prefix, l)]))
        self.keep += cls_out + cen_out + reg_out + iou_out
        as_nchw = lambda ts: [t.permute(0, 3, 1, 2) for t in ts]
        return as_nchw(cls_out), as_nchw(reg_out), as_nchw(cen_out), as_nchw(iou_out)

    # ------------------------------------------------------------------ layout I/O
    def pack_input(self, x):
        """(N,C,H,W) fp32 cuda tensor -> FlatMap (op appended)."""
        n, c, h, w = x.shape
        m = self.new_map([(n, h, w)], c)
        lib, xp, mp, fmt = self.lib, x.data_ptr(), m.ptr, self.fmt
        self.keep.append(x)
        self.ops.append(("pack", lambda st: L.check(lib.iou_pack_nchw_fmt(xp, n, c, h, w, mp, 0, fmt, st))))
        return m

    def unpack_output(self, m, s=0):
        rs, n, h, w = m.segs[s]
        out = torch.empty(n, m.c, h, w, dtype=torch.float32, device=self.device)
        self.keep.append(out)
        lib, mp, op, c, fmt = self.lib, m.ptr, out.data_ptr(), m.c, self.fmt
        self.ops.append(("unpack", lambda st: L.check(lib.iou_unpack_nchw_fmt(mp, rs, n, c, h, w, op, fmt, st))))
        return out

    # ------------------------------------------------------------------ execution
    def run(self):
        st = L.stream_ptr()
        if not self.side:
            for _, fn in self.ops:
                fn(st)
        else:
            cur = torch.cuda.current_stream(self.device)
            if self._side_stream is None:
                self._side_stream = torch.cuda.Stream(self.device)
            side = self._side_stream
            sst = ctypes.c_void_p(side.cuda_stream)
            fork = {f: (a, b) for (f, a, b) in self.side}
            on_side = set(i for (_, a, b) in self.side for i in range(a, b + 1))
            join = set(b for (_, a, b) in self.side)
            for i, (_, fn) in enumerate(self.ops):
                if i in fork:
                    side.wait_stream(cur)                 # fork: the side branch sees everything queued so far
                if i in on_side:
                    fn(sst)
                else:
                    fn(st)
                if i in join:
                    cur.wait_stream(side)                 # join (also what ends a CUDA-graph capture cleanly)
        L.launch_count += len(self.ops)

    def profile(self, iters=3):
        """Eager pass with a CUDA event pair around every launch (on the launching stream).
        Returns [(name, median_ms)] in launch order; used by bench.py for the live roofline."""
        st_t = torch.cuda.current_stream()
        st = L.stream_ptr()
        acc = [[] for _ in self.ops]
        for _ in range(iters):
            evs = []
            for _, fn in self.ops:
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(st_t)
                fn(st)
                b.record(st_t)
                evs.append((a, b))
            torch.cuda.synchronize()
            for i, (a, b) in enumerate(evs):
                acc[i].append(a.elapsed_time(b))
        L.launch_count += iters * len(self.ops)
        return [(name, sorted(t)[len(t) // 2]) for (name, _), t in zip(self.ops, acc)]      # median over the passes

    def num_launches(self):
        return len(self.ops)

    def range_report(self):
        """Range statistics of every activation map as the LAST run left them (iou_range_stats): list of dicts
        (label, max_abs, saturated, above_448, nonzero, elements).  `saturated` counts values at or beyond the fp16
        limit: with passes == 2 the encode clamps there, i.e. the value the reference's fp32 holds was lost."""
        out = torch.zeros(len(self.maps), 4, dtype=torch.int64, device=self.device)
        st = L.stream_ptr()
        for i, m in enumerate(self.maps):
            L.check(self.lib.iou_range_stats(m.ptr, m.rows, m.c, self.fmt, out[i].data_ptr(), st))
        rows = out.cpu()
        rep = []
        for i, m in enumerate(self.maps):
            mx = torch.tensor([int(rows[i, 0]) & 0xffffffff], dtype=torch.int64).to(torch.int32).view(torch.float32).item()
            rep.append(dict(label=getattr(m, "label", None) or "map%d" % i, max_abs=mx, saturated=int(rows[i, 1]),
                            above_448=int(rows[i, 2]), nonzero=int(rows[i, 3]), elements=m.rows * m.c))
        return rep

    def __del__(self):
        try:
            for p in self.plans:
                self.lib.iou_conv_plan_destroy(p)
        except Exception:
            pass
