// tcgen05 "tap GEMM" convolution for sm_100a.
//
// One persistent, warp-specialised kernel computes, for every 128-row tile of a padded-rows
// output map and every BLOCK_N-wide slice of output channels,
//     acc[m, n] = sum_taps sum_k  A_tap[m + off_tap, k] * W_tap[n, k]
// where A_tap is a 2-D view [rows][2*Cin] (bf16 hi | lo) of an activation map in the
// padded-rows layout (include/iou_b200.h) and off_tap = dy*(W+2)+dx is a constant row offset:
// a 3x3 convolution is nine shifted TMA box loads of the SAME matrix, no im2col buffer exists.
// Stride-2 convolutions read the four phase maps written by iou_phase_split, which turns
// them into the same constant-offset form.  Out-of-range rows are zero-filled by TMA.
//
//   warp 0      : TMA producer (cp.async.bulk.tensor.2d, 128B swizzle, mbarrier complete_tx)
//   warp 1      : TMEM allocator + tcgen05.mma issuer (kind::f16, bf16 x bf16 -> fp32 in TMEM)
//   warps 2..5  : epilogue (tcgen05.ld -> scale/shift (BN fold or bias) -> +residual -> ReLU
//                 -> bf16 hi|lo padded rows, or dense fp32 NHWC for the head outputs)
// fp32-grade accuracy comes from three bf16 MMAs per K step (hi*hi + hi*lo + lo*hi) into the
// same fp32 accumulator; `passes = 1` runs plain bf16.  Two TMEM accumulator stages let the
// epilogue of tile i overlap the MMAs of tile i+1.
// `passes = 2` (template kF8) is the fp16 + e4m3 scheme of split_fmt.cuh: per K step one kind::f16 MMA
// (fp16 hi x fp16 Wh') and one kind::f8f6f4 MMA ([x8|l8] x [Wl8;W8], K = 32).  All three weight copies carry the
// same per-output-channel power-of-two scale S_n, so both MMAs add into ONE fp32 accumulator (= S_n x result) and
// the epilogue multiplies by 1/S_n (`scale`): two bf16-pass equivalents of tensor-pipe time instead of three, the
// same TMEM footprint.  Narrow tiles (N <= 64) keep two column blocks so that consecutive MMAs do not serialise
// on one accumulator.
//
// Replaces the F.conv2d / cuDNN calls of mmdet/models/backbones/resnet.py:224-267,
// mmdet/models/necks/fpn.py:97-136, mmdet/models/anchor_heads/iou_aware_retina_head.py:171-219.
#include "conv_common.cuh"

namespace iou {

// ------------------------------------------------------------------------------------ kernel
// kEpiQ = epilogue warps per TMEM lane quadrant: 2 (10 warps) or 3 (the "wide" variant, 16 warps with setmaxnreg)
template <bool kTwoCta, bool kF8, int kEpiQ = 2>
__global__ void __launch_bounds__(kEpiQ == 3 ? kNumThreadsWide : kNumThreads, 1) conv_tap_gemm_kernel(const __grid_constant__ ConvParams P) {
  constexpr int kFmt = kF8 ? kFmtF16F8 : kFmtBf16x2;
  constexpr int kEpiWarps = 4 * kEpiQ;
  constexpr int kFirstEpiWarp = kEpiQ == 3 ? 4 : 2;
  constexpr uint32_t kBarRes = kEpiQ == 3 ? kBarResWide : 192u;
  extern __shared__ unsigned char smem_dyn[];
  // 1024-byte alignment is required by the 128B swizzle atoms
  const uint32_t raw = smem_u32(smem_dyn);
  const uint32_t base = (raw + 1023u) & ~1023u;
  unsigned char* ctrl = smem_dyn + (base - raw);
  const uint32_t ctrl_addr = base;
  const uint32_t tiles_addr = base + kCtrlBytes;
  // control block: tfull[2] @128, tempty[2] @144, tmem ptr @160, residual ring @192..320, A full[8] @320,
  // A empty[8] @384, B full[16] @448, B empty[16] @576
  const uint32_t bar_full = ctrl_addr + 448, bar_empty = ctrl_addr + 576, bar_tfull = ctrl_addr + 128,
                 bar_tempty = ctrl_addr + 144, bar_afull = ctrl_addr + 320, bar_aempty = ctrl_addr + 384;
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(ctrl + 160);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // work distribution: one item = (128-row tile, N tile) per CTA, or (two 128-row tiles, N tile) per CTA pair
  const int rank = kTwoCta ? (int)cluster_ctarank() : 0;
  const int w_first = kTwoCta ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int w_stride = kTwoCta ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int w_total = kTwoCta ? P.total_pair_tiles : P.total_tiles;
  auto decode_tile = [&](int wi, int& m_tile, int& n_tile, int& sidx) {
    const int q = wi / P.num_n_tiles;
    n_tile = wi - q * P.num_n_tiles;
    m_tile = kTwoCta ? 2 * q + rank : q;
    sidx = 0;
    while (sidx + 1 < P.num_seg && m_tile >= P.seg_tile_off[sidx + 1]) ++sidx;
  };

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < P.num_b_stages; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
    for (int s = 0; s < P.num_a_stages; ++s) { mbar_init(bar_afull + 8 * s, 1); mbar_init(bar_aempty + 8 * s, 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(bar_tfull + 8 * a, 1); mbar_init(bar_tempty + 8 * a, (kTwoCta ? 2 : 1) * kEpiWarps * 32); }
    for (int r = 0; r < 2 * kEpiWarps; ++r) mbar_init(ctrl_addr + kBarRes + 8 * r, 1);   // residual ring: two per epilogue warp
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    for (int i = 0; i < IOU_CONV_MAX_SRC; ++i)
      asm volatile("prefetch.tensormap [%0];" ::"l"(&P.tmap_src[i]) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&P.tmap_w) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&P.tmap_out) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&P.tmap_res) : "memory");
  }
  if (warp == 1) {
    if constexpr (kTwoCta) {       // the same warp of both CTAs allocates collectively
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;"
                   ::"r"(ctrl_addr + 160), "r"(kTmemCols) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                   ::"r"(ctrl_addr + 160), "r"(kTmemCols) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  if constexpr (kTwoCta) cluster_sync_all();   // the peer's barriers must be initialised before any remote signal
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  if (P.pdl) griddep_launch_dependents();      // every CTA of this (persistent) grid is resident: dependents fill freed SMs

  const uint32_t a_lo_off = (uint32_t)P.a_rows * 128u;       // A entry: hi window | lo window
  const uint32_t b_lo_off = (uint32_t)P.b_tile_bytes;        // B entry: hi tile | lo tile
  const uint32_t b_ring_addr = tiles_addr + (uint32_t)(P.num_a_stages * P.a_entry_bytes);

  // wide: register budget per scheduler partition: 4 warps x 128 at launch -> 56 (warpgroup 0) + 3 x 152 (epilogue).  The
  // setmaxnreg instructions sit INSIDE the role branches: ptxas sizes the register allocation of the code they dominate
  if (warp < kFirstEpiWarp) {
  if constexpr (kEpiQ == 3) asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
  if (warp == 0) {
    // =============================== TMA producer ===============================
    if (elect_one()) {
      if (P.pdl) griddep_wait();                 // the activations are the previous kernel's output
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      const uint32_t mult = kTwoCta ? 2u : 1u;   // pair: both CTAs' loads complete on the LEADER's barrier
      for (int tile = w_first; tile < w_total; tile += w_stride) {
        int m_tile, n_tile, s;
        decode_tile(tile, m_tile, n_tile, s);
        const int row0 = P.seg[s].row_start + (m_tile - P.seg_tile_off[s]) * kBlockM;
        const int wp = P.seg[s].w + 2;
        const int b_rows = kTwoCta ? (P.block_n >> 1) : P.block_n;     // a pair splits the B tile by rows
        for (int g = 0; g < P.num_groups; ++g) {
          const int arow = row0 + P.grp_dy[g] * wp + P.grp_dx0[g];
          const CUtensorMap* tm = &P.tmap_src[P.grp_src[g]];
          const int nt = P.grp_nt[g];
          const int g_cin = P.src_cin[P.grp_src[g]];
          for (int ks = 0; ks < P.grp_ks[g]; ++ks) {
            // ---- one A window (hi, lo) serves all dx taps of the group
            mbar_wait(bar_aempty + 8 * as, aph ^ 1u);
            const uint32_t fa = bar_afull + 8 * as;
            const uint32_t sa = tiles_addr + as * P.a_entry_bytes;
            const int a_col = (P.diag_k ? n_tile : (P.ksplit_ntiles ? (n_tile / P.ksplit_ntiles) * P.grp_ks[g] + ks : ks)) * kBlockK;
            if (!kTwoCta || rank == 0) mbar_expect_tx(fa, mult * (uint32_t)P.a_entry_bytes);
            if constexpr (kTwoCta) {
              tma_load_2d_pair(tm, fa, sa, a_col, arow);
              if (P.passes == 3) tma_load_2d_pair(tm, fa, sa + a_lo_off, g_cin + a_col, arow);
            } else {
              tma_load_2d(tm, fa, sa, a_col, arow);
              if (P.passes == 3) tma_load_2d(tm, fa, sa + a_lo_off, g_cin + a_col, arow);
            }
            if (++as == P.num_a_stages) { as = 0; aph ^= 1u; }
            // ---- one B tile (hi, lo) per tap
            for (int j = 0; j < nt; ++j) {
              // resident weights: the ring has one entry per (tap, slab) of a tile and is filled once, with the first tile
              if (!P.b_resident || tile == w_first) {
                const int wrow = P.grp_tap[g][j] * P.cout_pad + n_tile * P.block_n + rank * b_rows;
                mbar_wait(bar_empty + 8 * bs, bph ^ 1u);
                const uint32_t fb = bar_full + 8 * bs;
                const uint32_t sb = b_ring_addr + bs * P.b_entry_bytes;
                if (!kTwoCta || rank == 0) mbar_expect_tx(fb, mult * (uint32_t)P.b_entry_bytes);
                if constexpr (kTwoCta) {
                  tma_load_2d_pair(&P.tmap_w, fb, sb, ks * kBlockK, wrow);
                  if (P.passes == 3) tma_load_2d_pair(&P.tmap_w, fb, sb + b_lo_off, P.b_cin + ks * kBlockK, wrow);
                } else {
                  tma_load_2d(&P.tmap_w, fb, sb, ks * kBlockK, wrow);
                  if (P.passes == 3) tma_load_2d(&P.tmap_w, fb, sb + b_lo_off, P.b_cin + ks * kBlockK, wrow);
                }
              }
              if (++bs == P.num_b_stages) { bs = 0; bph ^= 1u; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer (pair mode: leader CTA only) ===============================
    int as = 0, bs = 0;
    uint32_t aph = 0, bph = 0;
    int it = 0;
    const int mode = kF8 ? 6 : (P.combine ? 5 : (P.passes == 3 ? (P.lolo ? 4 : 3) : 1));
    const uint32_t corr_off = (uint32_t)P.corr_off;
    const bool two_acc = P.num_acc == 2;
    const uint32_t a_lo_d = a_lo_off >> 4, b_lo_d = b_lo_off >> 4;
    const uint32_t idesc = P.idesc, idesc2 = P.idesc2, col2 = 2u * (uint32_t)P.block_n;
    auto mma_i = [&](uint32_t d_tmem, uint64_t ad, uint64_t bd, uint32_t id, uint32_t accum) {
      if constexpr (kTwoCta) tc_mma_bf16_pair(d_tmem, ad, bd, id, accum);
      else tc_mma_bf16(d_tmem, ad, bd, id, accum);
    };
    auto mma = [&](uint32_t d_tmem, uint64_t ad, uint64_t bd, uint32_t accum) { mma_i(d_tmem, ad, bd, idesc, accum); };
    for (int tile = w_first; tile < w_total && rank == 0; tile += w_stride, ++it) {
      const int acc = two_acc ? (it & 1) : 0;
      const uint32_t acc_phase = (uint32_t)(two_acc ? (it >> 1) : it) & 1u;
      mbar_wait(bar_tempty + 8 * acc, acc_phase ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * kAccStride);
      int done = 0;                                    // taps x slabs issued for this tile
      const int todo = P.tap_slabs_per_tile;
      for (int g = 0; g < P.num_groups; ++g) {
        const int nt = P.grp_nt[g];
        for (int ks = 0; ks < P.grp_ks[g]; ++ks) {
          mbar_wait(bar_afull + 8 * as, aph);
          const uint32_t sa = tiles_addr + as * P.a_entry_bytes;
          for (int j = 0; j < nt; ++j, ++done) {
            if (!P.b_resident || it == 0) mbar_wait(bar_full + 8 * bs, bph);
            tc_fence_after();
            if (elect_one()) {
              // descriptor low words: (addr >> 4) | LBO; every operand lives below 256 KiB, so advancing an
              // address by x bytes is adding x >> 4 (the single issuing thread is the bottleneck of narrow-N
              // tiles: 12 MMAs of 128 x 64 x 16 retire in ~400 cycles, so the loop body is kept minimal)
              const uint32_t sb = b_ring_addr + bs * P.b_entry_bytes;
              const uint32_t da = umma_desc_lo(sa + (uint32_t)P.grp_shift[g][j] * 128u), db = umma_desc_lo(sb);
              const uint32_t dal = da + a_lo_d, dbl = db + b_lo_d;
              const uint32_t first = done > 0 ? 1u : 0u;
              if constexpr (kF8) {
                // fp16 hi x Wh' and [x8|l8] x [Wl8;W8] (e4m3, K = 32 per 32 bytes), both at scale S_n, into the same
                // accumulator (narrow tiles: two column blocks).  Both instruction descriptors have the same bits
                // (formats 0 = F16 / E4M3).
                const uint32_t cfirst = corr_off ? first : 1u;     // same column block: the fp16 MMA initialised it
#pragma unroll
                for (uint32_t kk = 0; kk < kBlockK / 16; ++kk) {
                  mma_i(d_tmem, umma_desc(da + 2 * kk), umma_desc(db + 2 * kk), idesc, kk ? 1u : first);
                  if constexpr (kTwoCta) tc_mma_f8_pair(d_tmem + corr_off, umma_desc(dal + 2 * kk), umma_desc(dbl + 2 * kk), idesc, kk ? 1u : cfirst);
                  else tc_mma_f8(d_tmem + corr_off, umma_desc(dal + 2 * kk), umma_desc(dbl + 2 * kk), idesc, kk ? 1u : cfirst);
                }
              } else if (mode == 5) {
                // back-to-back MMAs into the SAME accumulator serialise on its read-modify-write latency
                // (~120 cycles, more than a 128 x 64 x 16 MMA computes for): B_hi and B_lo are adjacent in
                // the B entry, so hi*hi and hi*lo are one N = 2*block_n MMA into columns [0, 2bn) and lo*hi
                // goes to columns [2bn, 3bn); the epilogue adds the three column blocks
#pragma unroll
                for (uint32_t kk = 0; kk < kBlockK / 16; ++kk) {
                  mma_i(d_tmem, umma_desc(da + 2 * kk), umma_desc(db + 2 * kk), idesc2, kk ? 1u : first);
                  mma_i(d_tmem + col2, umma_desc(dal + 2 * kk), umma_desc(db + 2 * kk), idesc, kk ? 1u : first);
                }
              } else if (mode == 3) {
#pragma unroll
                for (uint32_t kk = 0; kk < kBlockK / 16; ++kk) {
                  mma(d_tmem, umma_desc(da + 2 * kk), umma_desc(db + 2 * kk), kk ? 1u : first);
                  mma(d_tmem, umma_desc(da + 2 * kk), umma_desc(dbl + 2 * kk), 1u);
                  mma(d_tmem, umma_desc(dal + 2 * kk), umma_desc(db + 2 * kk), 1u);
                }
              } else if (mode == 1) {
#pragma unroll
                for (uint32_t kk = 0; kk < kBlockK / 16; ++kk)
                  mma(d_tmem, umma_desc(da + 2 * kk), umma_desc(db + 2 * kk), kk ? 1u : first);
              } else {
#pragma unroll
                for (uint32_t kk = 0; kk < kBlockK / 16; ++kk) {
                  mma(d_tmem, umma_desc(da + 2 * kk), umma_desc(db + 2 * kk), kk ? 1u : first);
                  mma(d_tmem, umma_desc(da + 2 * kk), umma_desc(dbl + 2 * kk), 1u);
                  mma(d_tmem, umma_desc(dal + 2 * kk), umma_desc(db + 2 * kk), 1u);
                  mma(d_tmem, umma_desc(dal + 2 * kk), umma_desc(dbl + 2 * kk), 1u);
                }
              }
              // commits fire when the MMAs issued so far retire (pair mode: multicast to BOTH CTAs)
              if constexpr (kTwoCta) {
                if (!P.b_resident) tc_commit_pair(bar_empty + 8 * bs);
                if (j == nt - 1) tc_commit_pair(bar_aempty + 8 * as);
                if (done == todo - 1) tc_commit_pair(bar_tfull + 8 * acc);
              } else {
                if (!P.b_resident) tc_commit(bar_empty + 8 * bs);  // frees the B tile
                if (j == nt - 1) tc_commit(bar_aempty + 8 * as);   // frees the A window
                if (done == todo - 1) tc_commit(bar_tfull + 8 * acc);   // accumulator complete
              }
            }
            __syncwarp();
            if (++bs == P.num_b_stages) { bs = 0; bph ^= 1u; }
          }
          if (++as == P.num_a_stages) { as = 0; aph ^= 1u; }
        }
      }
    }
  }
  } else {
    // =============================== epilogue ===============================
    if constexpr (kEpiQ == 3) asm volatile("setmaxnreg.inc.sync.aligned.u32 152;");
    if (P.pdl) griddep_wait();                   // residual reads (and, conservatively, every store) follow the previous grid
    const int lane_group = warp & 3;                      // TMEM lanes 32*lane_group .. +31
    const int m_local = lane_group * 32 + lane;
    const int ew = warp - kFirstEpiWarp;                   // 0..7 (wide: 0..11)
    const int half = ew >> 2;                              // the kEpiQ warps of a quadrant take alternate column groups
    const uint32_t st_out = tiles_addr + P.ring_bytes + ew * P.staging_per_warp;
    // kEpiQ == 2: staging tile + 2 residual buffers.  kEpiQ == 3: two buffers, used IN PLACE -- the residual slab
    // arrives in buffer q & 1, every lane reads its own row and writes its results over it, the TMA store leaves from there
    const uint32_t st_res = kEpiQ == 3 ? st_out : st_out + 4096;
    const uint32_t bar_res = ctrl_addr + kBarRes + ew * 16;
    auto issue_res_at = [&](int row, int col, int q_) {    // one elected lane only
      const uint32_t bar = bar_res + 8 * (q_ & 1), dst = st_res + (q_ & 1) * 4096;
      mbar_expect_tx(bar, 4096u);
      tma_load_2d(&P.tmap_res, bar, dst, col, row);
      tma_load_2d(&P.tmap_res, bar, dst + 2048, P.cout + col, row);
    };
    auto issue_res = [&](int tile_, int g_, int q_) {      // one elected lane only; decodes the tile (divisions)
      int mt, nt, s_;
      decode_tile(tile_, mt, nt, s_);
      issue_res_at(P.seg[s_].row_start + (mt - P.seg_tile_off[s_]) * kBlockM + lane_group * 32,
                   nt * P.block_n + g_ * 32, q_);
    };
    // the 2-deep shared-memory ring covers one slab of latency; the slabs of the tiles further ahead are pulled
    // into L2 (no shared memory needed) so that the ring's loads hit L2 instead of HBM
    auto prefetch_res = [&](int tile_) {                   // one elected lane only
      if (tile_ >= w_total) return;
      int mt, nt, s_;
      decode_tile(tile_, mt, nt, s_);
      const int row = P.seg[s_].row_start + (mt - P.seg_tile_off[s_]) * kBlockM + lane_group * 32;
      for (int g_ = half; g_ < (P.block_n >> 5); g_ += 2) {
        const int col = nt * P.block_n + g_ * 32;
        tma_prefetch_l2_2d(&P.tmap_res, col, row);
        tma_prefetch_l2_2d(&P.tmap_res, P.cout + col, row);
      }
    };
    int rq = 0;                                            // running slab counter of the residual ring
    // wide: the three warps of a quadrant take the column groups round-robin ACROSS tiles (group id = it * n_groups + g,
    // warp = id % 3), so that 8, 4 or 2 groups per tile still keep all three busy
    const int n_groups_w = P.block_n >> 5;
    auto first_group = [&](int it_) { return (half + 3 - (it_ * n_groups_w) % 3) % 3; };
    // this warp's first slab at or after (tile_, it_, g_): false when there is none
    auto find_slab = [&](int& tile_, int& it_, int& g_) {
      for (;;) {
        if (g_ < n_groups_w) return true;
        tile_ += w_stride; ++it_;
        if (tile_ >= w_total) return false;
        g_ = first_group(it_);
      }
    };
    if constexpr (kEpiQ == 3) {
      if (P.res_staged && w_first < w_total && elect_one()) {
        int t0 = w_first, i0 = 0, g0 = first_group(0);
        if (find_slab(t0, i0, g0)) issue_res(t0, g0, 0);
      }
    } else {
      if (P.res_staged && w_first < w_total && elect_one()) {
        issue_res(w_first, half, 0);
        for (int a = 1; a <= P.res_prefetch; ++a) prefetch_res(w_first + a * w_stride);
      }
    }
    int it = 0;
    const bool two_acc = P.num_acc == 2;
    for (int tile = w_first; tile < w_total; tile += w_stride, ++it) {
      const int acc = two_acc ? (it & 1) : 0;
      const uint32_t acc_phase = (uint32_t)(two_acc ? (it >> 1) : it) & 1u;
      int m_tile, n_tile, s;
      decode_tile(tile, m_tile, n_tile, s);
      if (kEpiQ == 2 && P.res_staged && P.res_prefetch > 0 && elect_one()) prefetch_res(tile + (P.res_prefetch + 1) * w_stride);
      __syncwarp();
      const bool tile_valid = m_tile < P.num_m_tiles;        // pair mode: the odd CTA of the last pair may idle
      const SegDev sg = P.seg[s];
      const int grow = sg.row_start + (m_tile - P.seg_tile_off[s]) * kBlockM + m_local;
      const int wp = sg.w + 2, plane = (sg.h + 2) * wp;
      const int rel = grow - sg.row_start;
      const int img = rel / plane, rem = rel - img * plane;
      const int yp = rem / wp, xp = rem - yp * wp;
      const bool interior = tile_valid && (img < sg.n_img) && (yp >= 1) && (yp <= sg.h) && (xp >= 1) && (xp <= sg.w);
      const __nv_bfloat16* res_row = nullptr;
      if (P.res_mode == IOU_RES_SAME && !P.res_staged) {
        res_row = P.residual + (size_t)grow * (2 * P.cout);
      } else if (P.res_mode == IOU_RES_UPSAMPLE2 && interior) {
        const SegDev rs = P.res_seg[s];
        const int ry = ((yp - 1) >> 1) + 1, rx = ((xp - 1) >> 1) + 1;
        const size_t rrow = (size_t)rs.row_start + ((size_t)img * (rs.h + 2) + ry) * (rs.w + 2) + rx;
        res_row = P.residual + rrow * (2 * P.cout);
      }
      __nv_bfloat16* ph_row = nullptr;                       // this row's place in a stride-2 phase map, if any
      if (P.phase_any && tile_valid && img < sg.n_img) {
        __nv_bfloat16* pb = P.phase_out[((yp & 1) << 1) | (xp & 1)];
        if (pb != nullptr)
          ph_row = pb + ((size_t)(img * (P.ph_h + 2) + (yp >> 1) + 1) * (P.ph_w + 2) + (xp >> 1) + 1) * (size_t)(2 * P.cout);
      }
      // phase rows leave through the staging tile: lane l stores chunk (l & 3) of rows 8*i + (l >> 2), i = 0..3, so one
      // store instruction writes 8 rows x 64 contiguous bytes (not 32 rows x 16 bytes)
      unsigned long long ph_dst[4] = {0ull, 0ull, 0ull, 0ull};
      if (P.phase_any) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          ph_dst[i] = __shfl_sync(0xffffffffu, (unsigned long long)(uintptr_t)ph_row, 8 * i + (lane >> 2));
      }
      mbar_wait(bar_tfull + 8 * acc, acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(lane_group * 32) << 16) + (uint32_t)(acc * kAccStride);
      const int n_chunks = P.block_n >> 4;
      if (!P.staged) {
      // column-group maxima (retina_cls: the per-anchor max class logit): a running max over this warp's chunks of the
      // current group, flushed when the group changes
      float gm = -INFINITY;
      int gcur = -1;
      const int cpg = P.gmax_cols >> 4;                      // 16-column chunks per group
      const size_t gpix = ((size_t)img * sg.h + (yp - 1)) * sg.w + (xp - 1);
      auto gmax_flush = [&]() {
        const int gidx = n_tile * (P.block_n / (P.gmax_cols > 0 ? P.gmax_cols : 1)) + gcur;
        if (gcur >= 0 && interior && gidx < P.gmax_groups)
          P.gmax_out[s][(gpix * P.gmax_groups + gidx) * 2 + half] = gm;
      };
      for (int ch = half; ch < n_chunks; ch += kEpiQ) {
        uint32_t v[16];
        tc_ld16(t_row + ch * 16, v);
        tc_wait_ld();
        const int c0 = n_tile * P.block_n + ch * 16;      // first output channel of this chunk
        if constexpr (kF8) {                               // (+ second column block) x per-channel 1 / S_n
          if (P.corr_off) {
            uint32_t w1[16];
            tc_ld16(t_row + P.corr_off + ch * 16, w1);
            tc_wait_ld();
#pragma unroll
            for (int q = 0; q < 16; ++q) v[q] = __float_as_uint(__uint_as_float(v[q]) + __uint_as_float(w1[q]));
          }
#pragma unroll
          for (int q = 0; q < 16; ++q) v[q] = __float_as_uint(__uint_as_float(v[q]) * __ldg(P.scale + c0 + q));
        } else if (P.combine) {                            // sum the hi*hi, hi*lo and lo*hi column blocks
          uint32_t w1[16], w2[16];
          tc_ld16(t_row + P.block_n + ch * 16, w1);
          tc_ld16(t_row + 2 * P.block_n + ch * 16, w2);
          tc_wait_ld();
#pragma unroll
          for (int q = 0; q < 16; ++q)
            v[q] = __float_as_uint(__uint_as_float(v[q]) + (__uint_as_float(w1[q]) + __uint_as_float(w2[q])));
        }
        float f[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) {
          float x = __uint_as_float(v[q]);
          if (!kF8 && P.scale) x *= __ldg(P.scale + c0 + q);
          if (P.shift) x += __ldg(P.shift + c0 + q);
          f[q] = x;
        }
        if (res_row != nullptr && interior) {
          const uint4* rh = reinterpret_cast<const uint4*>(res_row + c0);
          const uint4* rl = reinterpret_cast<const uint4*>(res_row + P.cout + c0);
#pragma unroll
          for (int h2 = 0; h2 < 2; ++h2) {
            float t8[8];
            decode8<kFmt>(__ldg(rh + h2), __ldg(rl + h2), t8);
#pragma unroll
            for (int q = 0; q < 8; ++q) f[h2 * 8 + q] += t8[q];
          }
        }
        if (P.relu) {
#pragma unroll
          for (int q = 0; q < 16; ++q) f[q] = fmaxf(f[q], 0.f);
        }
        if (P.gmax_cols) {
          const int gi = ch / cpg;
          if (gi != gcur) { gmax_flush(); gcur = gi; gm = -INFINITY; }
#pragma unroll
          for (int q = 0; q < 16; ++q) gm = fmaxf(gm, f[q]);
        }
        if (interior) {
          const size_t pix = ((size_t)img * sg.h + (yp - 1)) * sg.w + (xp - 1);
          const int split = P.dense_split > 0 ? P.dense_split : P.cout;
          if (c0 + 16 <= split && (split & 3) == 0) {
            float4* o = reinterpret_cast<float4*>(P.out_dense[s] + pix * split + c0);
#pragma unroll
            for (int q = 0; q < 4; ++q) o[q] = make_float4(f[4 * q], f[4 * q + 1], f[4 * q + 2], f[4 * q + 3]);
          } else {
            const int w2 = P.cout - split;
#pragma unroll
            for (int q = 0; q < 16; ++q) {
              const int c = c0 + q;
              if (c < split) P.out_dense[s][pix * split + c] = f[q];
              else if (c < P.cout) P.out_dense2[s][pix * w2 + (c - split)] = f[q];
            }
          }
        }
      }
      if (P.gmax_cols) gmax_flush();
      } else {
        // ---- staged path (padded-rows output): 32 output channels per step and warp.  Results go to a
        // 64B-swizzled shared-memory tile (32 rows x 32 ch, hi and lo) and leave through TMA stores
        // (coalesced, clipped at the end of the buffer); a same-geometry residual arrives through a
        // 2-deep TMA-load ring per warp.
        const int row_tile0 = grow - lane;                 // first row of this warp's 32-row slab
        const int n_groups = P.block_n >> 5;
        const uint32_t swz = (uint32_t)((lane >> 1) & 3);  // 64B swizzle: 16B chunk j of row r sits at j ^ ((r>>1)&3)
        for (int g = kEpiQ == 3 ? first_group(it) : half; g < n_groups; g += kEpiQ, ++rq) {
          const int c0 = n_tile * P.block_n + g * 32;
          // where this slab's results are staged (wide + residual: over the residual slab itself)
          const uint32_t buf_out = (kEpiQ == 3 && P.res_staged) ? st_out + (uint32_t)(rq & 1) * 4096u : st_out;
          if (P.res_staged) {
            if (kEpiQ == 2 && elect_one()) {               // prefetch this warp's next slab
              if (g + 2 < n_groups) issue_res_at(row_tile0, n_tile * P.block_n + (g + 2) * 32, rq + 1);
              else if (tile + w_stride < w_total) issue_res(tile + w_stride, half, rq + 1);
            }
            mbar_wait(bar_res + 8 * (rq & 1), (uint32_t)(rq >> 1) & 1u);
          }
          uint32_t v[32];
          tc_ld32(t_row + g * 32, v);
          uint32_t wc[kF8 ? 32 : 1];
          if constexpr (kF8) {
            if (P.corr_off) tc_ld32(t_row + P.corr_off + g * 32, wc);
          }
          // per-channel shift (and optional scale) of the slab's 32 channels: 8 broadcast 16-byte loads (every lane
          // reads the same address) instead of one load + 32 shuffles
          float shv[32];
          if (P.shift) {
            const float4* sp = reinterpret_cast<const float4*>(P.shift + c0);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float4 t4 = __ldg(sp + q);
              shv[4 * q] = t4.x; shv[4 * q + 1] = t4.y; shv[4 * q + 2] = t4.z; shv[4 * q + 3] = t4.w;
            }
          } else {
#pragma unroll
            for (int q = 0; q < 32; ++q) shv[q] = 0.f;
          }
          const float sc_l = (!kF8 && P.scale) ? __ldg(P.scale + c0 + lane) : 1.f;
          tc_wait_ld();
          if constexpr (kF8) {                             // (+ second column block); x 1 / S_n happens with the shift
            if (P.corr_off) {
#pragma unroll
              for (int q = 0; q < 32; ++q) v[q] = __float_as_uint(__uint_as_float(v[q]) + __uint_as_float(wc[kF8 ? q : 0]));
            }
          } else if (P.combine) {                          // sum the hi*hi, hi*lo and lo*hi column blocks
            uint32_t w1[32];
            tc_ld32(t_row + P.block_n + g * 32, w1);
            tc_wait_ld();
#pragma unroll
            for (int q = 0; q < 32; ++q) v[q] = __float_as_uint(__uint_as_float(v[q]) + __uint_as_float(w1[q]));
            tc_ld32(t_row + 2 * P.block_n + g * 32, w1);
            tc_wait_ld();
#pragma unroll
            for (int q = 0; q < 32; ++q) v[q] = __float_as_uint(__uint_as_float(v[q]) + __uint_as_float(w1[q]));
          }
          float f[32];
          if constexpr (kF8) {                             // acc / S_n + shift: 8 broadcast 16-byte loads of 1 / S_n
            const float4* cp = reinterpret_cast<const float4*>(P.scale + c0);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float4 t4 = __ldg(cp + q);
              f[4 * q] = fmaf(__uint_as_float(v[4 * q]), t4.x, shv[4 * q]);
              f[4 * q + 1] = fmaf(__uint_as_float(v[4 * q + 1]), t4.y, shv[4 * q + 1]);
              f[4 * q + 2] = fmaf(__uint_as_float(v[4 * q + 2]), t4.z, shv[4 * q + 2]);
              f[4 * q + 3] = fmaf(__uint_as_float(v[4 * q + 3]), t4.w, shv[4 * q + 3]);
            }
          } else if (P.scale) {
#pragma unroll
            for (int q = 0; q < 32; ++q)
              f[q] = fmaf(__uint_as_float(v[q]), __shfl_sync(0xffffffffu, sc_l, q), shv[q]);
          } else {
#pragma unroll
            for (int q = 0; q < 32; ++q) f[q] = __uint_as_float(v[q]) + shv[q];
          }
          if (P.res_staged) {
            const uint32_t rb = st_res + (rq & 1) * 4096 + lane * 64;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint32_t sw = (uint32_t)((j ^ swz) << 4);
              decode8_add<kFmt>(lds128(rb + sw), lds128(rb + 2048 + sw), f + j * 8);
            }
          } else if (res_row != nullptr && interior) {     // nearest-2x upsampled residual (FPN laterals)
            const uint4* rh = reinterpret_cast<const uint4*>(res_row + c0);
            const uint4* rl = reinterpret_cast<const uint4*>(res_row + P.cout + c0);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              decode8_add<kFmt>(__ldg(rh + j), __ldg(rl + j), f + j * 8);
            }
          }
          if (kEpiQ == 3 && P.res_staged) {
            // the other buffer: once the previous slab's store has read it, the NEXT slab's residual may land there.  (This
            // slab's buffer needs no wait: its last store was waited for one slab ago, before its residual load was issued.)
            if (elect_one()) {                             // the lane that issued the stores: bulk groups are per thread
              tma_store_wait_read();
              int t2 = tile, i2 = it, g2 = g + 3;
              if (find_slab(t2, i2, g2)) {
                if (t2 == tile) issue_res_at(row_tile0, n_tile * P.block_n + g2 * 32, rq + 1);
                else issue_res(t2, g2, rq + 1);
              }
            }
          } else if (!P.phase_only) {
            if (elect_one()) tma_store_wait_read();        // (same elected lane as the stores) previous slab has left the staging tile
            __syncwarp();
          }
          const uint32_t ob = buf_out + lane * 64;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float x8[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) x8[q] = P.relu ? fmaxf(f[j * 8 + q], 0.f) : f[j * 8 + q];
            // packed conversions (F2FP): bf16 hi | lo, or fp16 hi | e4m3 pairs (split_fmt.cuh); border rows are zero
            uint4 hi, lo;
            encode8<kFmt, false>(x8, hi, lo);
            if (!interior) { hi = make_uint4(0, 0, 0, 0); lo = hi; }
            const uint32_t sw = (uint32_t)((j ^ swz) << 4);
            sts128(ob + sw, hi);
            sts128(ob + 2048 + sw, lo);
          }
          if (!P.phase_only) fence_async_smem();
          __syncwarp();
          if (P.phase_any) {                               // staged rows -> phase maps, 64 contiguous bytes per 4 lanes
            const uint32_t cch = (uint32_t)(lane & 3);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const uint32_t r = (uint32_t)(8 * i + (lane >> 2));
              const uint32_t sa = buf_out + r * 64 + ((cch ^ ((r >> 1) & 3u)) << 4);
              if (ph_dst[i] != 0ull) {
                __nv_bfloat16* dp = reinterpret_cast<__nv_bfloat16*>((uintptr_t)ph_dst[i]) + c0 + 8 * cch;
                *reinterpret_cast<uint4*>(dp) = lds128(sa);
                *reinterpret_cast<uint4*>(dp + P.cout) = lds128(sa + 2048);
              }
            }
            if (P.phase_only) { __syncwarp(); continue; }  // the next slab overwrites the staging tile
          }
          if (tile_valid && elect_one()) {
            tma_store_2d(&P.tmap_out, buf_out, c0, row_tile0);
            tma_store_2d(&P.tmap_out, buf_out + 2048, P.cout + c0, row_tile0);
            tma_store_commit();
          }
        }
      }
      tc_fence_before();
      if constexpr (kTwoCta) mbar_arrive_leader(bar_tempty + 8 * acc);   // only the leader's MMA warp waits on it
      else mbar_arrive(bar_tempty + 8 * acc);
    }
    if (P.staged && elect_one()) tma_store_wait_read();    // staging must outlive the last TMA store
  }

  tc_fence_before();
  if constexpr (kTwoCta) cluster_sync_all();   // neither CTA may release TMEM / exit while the pair is still in flight
  else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if constexpr (kTwoCta)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

// ------------------------------------------------------------------------------------ host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2-D bf16 matrix [rows][cols] (row-major), box = 64 cols x box_rows rows, 128B swizzle, zero OOB fill.
static int encode_2d(CUtensorMap* tm, const void* ptr, uint64_t rows, uint64_t cols, uint32_t box_rows,
                     uint32_t box_cols = kBlockK, CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return fail(IOU_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstr[1] = {cols * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(IOU_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return IOU_OK;
}

}  // namespace iou

using namespace iou;

extern "C" int iou_conv_plan_create(const iou_conv_desc* d, iou_conv_plan** plan_out) {
  IOU_REQUIRE(d && plan_out, "NULL argument");
  IOU_REQUIRE(d->cin > 0 && d->cin % kBlockK == 0, "cin must be a positive multiple of %d (got %d)", kBlockK, d->cin);
  IOU_REQUIRE(d->block_n >= 16 && d->block_n <= 256 && d->block_n % 16 == 0, "block_n must be in 16..256, multiple of 16");
  IOU_REQUIRE(d->cout >= 1 && d->cout_pad >= d->cout && d->cout_pad % d->block_n == 0, "cout_pad must be a multiple of block_n >= cout");
  IOU_REQUIRE(d->num_taps >= 1 && d->num_taps <= IOU_CONV_MAX_TAPS, "num_taps out of range");
  IOU_REQUIRE(d->num_src >= 1 && d->num_src <= IOU_CONV_MAX_SRC, "num_src out of range");
  IOU_REQUIRE(d->num_seg >= 1 && d->num_seg <= IOU_CONV_MAX_SEG, "num_seg out of range");
  IOU_REQUIRE(d->passes >= 1 && d->passes <= 4, "passes must be 1, 2 (fp16 + e4m3), 3 or 4");
  IOU_REQUIRE(d->passes != 2 || d->scale != nullptr, "passes == 2 needs `scale` = 1 / S_n, the inverse of the weights' per-channel scale");
  IOU_REQUIRE(d->weight != nullptr, "weight is NULL");
  IOU_REQUIRE(!d->diag_k || (d->block_n == 64 && d->cin == d->cout), "diag_k (grouped conv) needs block_n == 64 and cin == cout");
  IOU_REQUIRE(d->src_rows > 0 && d->src_rows < (1ll << 31), "src_rows out of range");
  const bool phase_any = d->phase_out[0] || d->phase_out[1] || d->phase_out[2] || d->phase_out[3];
  IOU_REQUIRE(!d->phase_only || phase_any, "phase_only without any phase_out");
  if (phase_any) {
    IOU_REQUIRE(d->out_mode == IOU_OUT_PADDED_BF16X2 && d->num_seg == 1 && d->block_n % 64 == 0,
                "phase_out needs a single-segment padded-rows output and block_n %% 64 == 0");
    for (int i = 0; i < 4; ++i) IOU_REQUIRE(((uintptr_t)d->phase_out[i] & 15) == 0, "phase_out[%d] must be 16-byte aligned", i);
  }
  if (d->out_mode == IOU_OUT_PADDED_BF16X2) {
    IOU_REQUIRE(d->out != nullptr || d->phase_only, "out is NULL");
    IOU_REQUIRE(d->cout == d->cout_pad, "padded output needs cout == cout_pad");
    IOU_REQUIRE(((uintptr_t)d->out & 15) == 0, "out must be 16-byte aligned");
  } else {
    IOU_REQUIRE(d->out_mode == IOU_OUT_DENSE_F32, "bad out_mode");
    IOU_REQUIRE(d->dense_split >= 0 && d->dense_split < d->cout, "dense_split out of range");
  }
  if (d->group_max_cols != 0) {
    IOU_REQUIRE(d->out_mode == IOU_OUT_DENSE_F32 && d->dense_split == 0, "group_max_cols needs a dense, unsplit output");
    IOU_REQUIRE(d->group_max_cols >= 32 && d->group_max_cols % 16 == 0 && d->block_n % d->group_max_cols == 0 &&
                d->cout % d->group_max_cols == 0, "group_max_cols must be a multiple of 16, >= 32, and divide block_n and cout");
    for (int s = 0; s < d->num_seg; ++s) IOU_REQUIRE(d->group_max_out[s] != nullptr, "group_max_out[%d] is NULL", s);
  }
  if (d->res_mode != IOU_RES_NONE) {
    IOU_REQUIRE(d->residual != nullptr, "residual is NULL");
    IOU_REQUIRE(d->cout == d->cout_pad && d->cout % 16 == 0, "residual needs cout == cout_pad, multiple of 16");
  }
  iou_conv_plan* plan = new (std::nothrow) iou_conv_plan();
  if (!plan) return fail(IOU_ERR_INVALID, "out of host memory");
  ConvParams& P = plan->params;
  memset(&P, 0, sizeof(P));
  P.cin = d->cin; P.cout = d->cout; P.cout_pad = d->cout_pad; P.block_n = d->block_n;
  P.num_taps = d->num_taps;
  P.diag_k = d->diag_k ? 1 : 0;
  P.b_cin = P.diag_k ? kBlockK : d->cin;
  P.k_slabs = P.diag_k ? 1 : d->cin / kBlockK;
  const int ksplit = d->k_split > 1 ? d->k_split : 1;
  if (ksplit > 1 && (P.diag_k || (d->cout_pad / d->block_n) % ksplit != 0 || d->out_mode != IOU_OUT_PADDED_BF16X2)) {
    delete plan;
    return fail(IOU_ERR_INVALID, "k_split needs a padded-rows output, no diag_k and cout_pad/block_n a multiple of k_split");
  }
  for (int i = 0; i < IOU_CONV_MAX_SRC; ++i) {
    P.src_cin[i] = (i < d->num_src && d->src_cin[i] > 0) ? d->src_cin[i] : d->cin * ksplit;
    if (ksplit > 1 && P.src_cin[i] != d->cin * ksplit) { delete plan; return fail(IOU_ERR_INVALID, "k_split: every source must hold k_split * cin channels"); }
    if (P.src_cin[i] % kBlockK != 0 || (ksplit == 1 && P.src_cin[i] > d->cin)) { delete plan; return fail(IOU_ERR_INVALID, "src_cin[%d] must be a multiple of %d and <= cin", i, kBlockK); }
    if (P.diag_k && P.src_cin[i] != d->cin) { delete plan; return fail(IOU_ERR_INVALID, "diag_k needs equal source channel counts"); }
  }
  P.passes = d->passes == 1 ? 1 : 3;   // 3 = hi/lo operands staged; lolo adds the fourth product
  P.lolo = d->passes == 4;
  P.f8 = d->passes == 2;
  // f8: one accumulator per stage (all weight copies share the scale S_n); narrow tiles split the two MMA kinds
  // over two column blocks, which the epilogue adds
  P.num_acc = 2;
  const int corr_max_bn = getenv("IOU_F8_CORR_MAX_BN") ? atoi(getenv("IOU_F8_CORR_MAX_BN")) : 64;
  P.corr_off = (P.f8 && d->block_n <= corr_max_bn && d->block_n <= 128 && !getenv("IOU_F8_ONE_BLOCK")) ? 128 : 0;
  for (int t = 0; t < d->num_taps; ++t) {
    if (d->tap_src[t] < 0 || d->tap_src[t] >= d->num_src) { delete plan; return fail(IOU_ERR_INVALID, "tap_src out of range"); }
    P.tap_src[t] = d->tap_src[t]; P.tap_dy[t] = d->tap_dy[t]; P.tap_dx[t] = d->tap_dx[t];
  }
  P.num_seg = d->num_seg;
  int toff = 0;
  double real_rows = 0;
  for (int s = 0; s < d->num_seg; ++s) {
    const iou_conv_segment& g = d->seg[s];
    if (g.row_start % kBlockM != 0 || g.n_img < 1 || g.h < 1 || g.w < 1) {
      delete plan;
      return fail(IOU_ERR_INVALID, "segment %d: row_start must be a multiple of %d and dims positive", s, kBlockM);
    }
    P.seg[s] = SegDev{g.row_start, g.n_img, g.h, g.w};
    P.res_seg[s] = SegDev{d->res_seg[s].row_start, d->res_seg[s].n_img, d->res_seg[s].h, d->res_seg[s].w};
    P.seg_tile_off[s] = toff;
    // tiles cover the rows up to the last interior pixel; the (w+2)+1 border rows behind it are never written and
    // stay zero from the allocation (padded-rows buffers must be zero-initialised, include/iou_b200.h).  At
    // 800x1344 this turns the 25x42 maps from 75 into 74 tiles = exactly 2 waves of 148 for N = 128 x 4
    long long rows = (long long)g.n_img * (g.h + 2) * (g.w + 2);
    rows -= (g.w + 2) + 1;
    toff += (int)((rows + kBlockM - 1) / kBlockM);
    real_rows += (double)g.n_img * g.h * g.w;
    P.out_dense[s] = (float*)d->out_dense[s];
    P.out_dense2[s] = (float*)d->out_dense2[s];
    P.gmax_out[s] = (float*)d->group_max_out[s];
    if (d->out_mode == IOU_OUT_DENSE_F32 && !d->out_dense[s]) { delete plan; return fail(IOU_ERR_INVALID, "out_dense[%d] is NULL", s); }
    if (d->out_mode == IOU_OUT_DENSE_F32 && d->dense_split > 0 && !d->out_dense2[s]) { delete plan; return fail(IOU_ERR_INVALID, "out_dense2[%d] is NULL", s); }
  }
  for (int s = d->num_seg; s <= IOU_CONV_MAX_SEG; ++s) P.seg_tile_off[s] = toff;
  P.gmax_cols = d->group_max_cols;
  P.gmax_groups = d->group_max_cols > 0 ? d->cout / d->group_max_cols : 0;
  P.num_m_tiles = toff;
  P.num_n_tiles = d->cout_pad / d->block_n;
  P.total_tiles = P.num_m_tiles * P.num_n_tiles;
  P.ksplit_ntiles = ksplit > 1 ? P.num_n_tiles / ksplit : 0;
  // CTA-pair mode (tcgen05 cta_group::2): two CTAs of a cluster own two consecutive 128-row tiles and share
  // one BLOCK_N-wide B tile, each staging half of its rows -> half the B traffic per CTA and a deeper pipeline
  P.two_cta = (d->two_cta && d->block_n % 16 == 0) ? 1 : 0;   // each CTA stages block_n/2 rows of B (whole 8-row swizzle atoms)
  P.total_pair_tiles = ((P.num_m_tiles + 1) / 2) * P.num_n_tiles;
  P.b_tile_bytes = (P.two_cta ? d->block_n / 2 : d->block_n) * kBlockK * 2;
  // padded-rows outputs leave through a per-warp 64B-swizzled staging tile (32 rows x 32 ch, hi + lo)
  // and TMA stores; a same-geometry residual arrives through a 2-deep TMA-load ring per warp
  P.staged = (d->out_mode == IOU_OUT_PADDED_BF16X2) && (d->block_n % 64 == 0);
  P.res_staged = P.staged && d->res_mode == IOU_RES_SAME;
  // wide variant (12 epilogue warps), decided below once the rings are known: IOU_WIDE = 0 never, 1 (default) the convs
  // whose epilogue, not their operand stream, is the work -- resident weights, or a residual with a short K loop --,
  // 2 every padded-rows conv of the fp16 + e4m3 scheme.  iou_conv_desc.wide = 1 / -1 overrides.
  const int wide_mode = getenv("IOU_WIDE") ? atoi(getenv("IOU_WIDE")) : 1;
  const bool wide_ok = P.f8 && P.staged && d->wide >= 0 && (d->wide > 0 || wide_mode >= 1);
  const bool wide_forced = d->wide > 0 || wide_mode >= 2;
  for (int i = 0; i < 4; ++i) P.phase_out[i] = (__nv_bfloat16*)d->phase_out[i];
  P.phase_any = phase_any ? 1 : 0;
  P.phase_only = d->phase_only ? 1 : 0;
  P.ph_h = (d->seg[0].h + 1) / 2;
  P.ph_w = (d->seg[0].w + 1) / 2;
  P.pdl = (getenv("IOU_PDL") && atoi(getenv("IOU_PDL")) != 0) ? 1 : 0;
  P.res_prefetch = getenv("IOU_RES_PREFETCH") ? atoi(getenv("IOU_RES_PREFETCH")) : 0;   // measured: no gain
  if (P.res_prefetch < 0 || P.res_prefetch > 8) P.res_prefetch = 0;
  const int nsplit = d->passes >= 2 ? 2 : 1;
  P.b_entry_bytes = nsplit * P.b_tile_bytes;
  if (P.b_tile_bytes % 1024 != 0) { delete plan; return fail(IOU_ERR_INVALID, "block_n %d: B tile is not a whole number of swizzle atoms", d->block_n); }
  P.taps_per_tile = d->num_taps;
  // tap groups: taps reading the same source at the same dy with dx within a span of 4 share one A window
  // (tried first; if the rings do not fit shared memory that way, every tap loads its own 128-row window)
  bool fits = false;
  int epi_warps = kNumEpiWarps;
  for (int try_wide = wide_ok ? 1 : 0; try_wide >= 0 && !fits; --try_wide) {
  P.wide = try_wide;
  epi_warps = P.wide ? kNumEpiWarpsWide : kNumEpiWarps;
  // per epilogue warp: a 4 KB staging tile (+ two 4 KB residual buffers); wide: the two residual buffers double as staging
  P.staging_per_warp = P.staged ? (P.wide ? (P.res_staged ? 8192 : 4096) : (4096 + (P.res_staged ? 8192 : 0))) : 0;
  for (int share = getenv("IOU_NO_A_SHARE") ? 0 : 1; share >= 0 && !fits; --share) {
    P.num_groups = 0;
    bool used[IOU_CONV_MAX_TAPS] = {false};
    for (int t = 0; t < d->num_taps; ++t) {
      if (used[t]) continue;
      const int g = P.num_groups++;
      int members[4] = {t, -1, -1, -1}, nm = 1, dxmin = d->tap_dx[t], dxmax = d->tap_dx[t];
      used[t] = true;
      for (int u = t + 1; u < d->num_taps && nm < 4 && share; ++u) {
        if (used[u] || d->tap_src[u] != d->tap_src[t] || d->tap_dy[u] != d->tap_dy[t]) continue;
        const int lo = d->tap_dx[u] < dxmin ? d->tap_dx[u] : dxmin, hi = d->tap_dx[u] > dxmax ? d->tap_dx[u] : dxmax;
        if (hi - lo > 3) continue;                      // the window has 8 spare rows; the stem uses dx = -2..1
        members[nm++] = u; used[u] = true; dxmin = lo; dxmax = hi;
      }
      P.grp_src[g] = d->tap_src[t]; P.grp_dy[g] = d->tap_dy[t]; P.grp_dx0[g] = dxmin; P.grp_nt[g] = nm;
      for (int j = 0; j < nm; ++j) { P.grp_tap[g][j] = members[j]; P.grp_shift[g][j] = d->tap_dx[members[j]] - dxmin; }
    }
    P.a_rows = (P.num_groups == d->num_taps) ? kBlockM : kBlockM + 8;    // shifted windows need up to 2 extra rows
    P.a_entry_bytes = nsplit * P.a_rows * kBlockK * 2;
    // ring depths: the A ring is worth (taps per group) B tiles per entry; maximise the shallower of the two
    const int budget = kSmemBudget - kCtrlBytes - 1024 - epi_warps * P.staging_per_warp;
    int best_na = 0, best_nb = 0, best_score = 0;
    // resident weights: with ONE N tile every tile of a CTA multiplies by the same few weight tiles -- load them once
    // (entry e = the e-th (tap, slab) of a tile) and spend the rest of shared memory on the A ring
    int b_entries = 0;
    for (int g = 0; g < P.num_groups; ++g) {
      P.grp_ks[g] = P.diag_k ? 1 : (ksplit > 1 ? d->cin / kBlockK : P.src_cin[P.grp_src[g]] / kBlockK);
      b_entries += P.grp_nt[g] * P.grp_ks[g];
    }
    P.tap_slabs_per_tile = b_entries;
    P.b_resident = 0;
    if (P.num_n_tiles == 1 && b_entries <= kMaxBStages && !getenv("IOU_NO_B_RESIDENT")) {
      int na = (budget - b_entries * P.b_entry_bytes) / P.a_entry_bytes;
      if (na > kMaxStages) na = kMaxStages;
      if (na >= 2) { P.b_resident = 1; best_na = na; best_nb = b_entries; best_score = 1; }
    }
    const int force_na = (getenv("IOU_FORCE_NA") && d->num_taps == 9 && d->cin <= 128) ? atoi(getenv("IOU_FORCE_NA")) : 0;   // (experiments)
    for (int na = 2; na <= kMaxStages && !P.b_resident; ++na) {
      if (force_na > 0 && na != force_na) continue;
      int nb = (budget - na * P.a_entry_bytes) / P.b_entry_bytes;
      if (nb > kMaxStages) nb = kMaxStages;
      if (nb < 2) break;
      const int cover = na * d->num_taps / P.num_groups;
      const int score = cover < nb ? cover : nb;
      if (score > best_score) { best_score = score; best_na = na; best_nb = nb; }
    }
    if (best_score > 0) {
      fits = true;
      P.num_a_stages = best_na; P.num_b_stages = best_nb;
      P.ring_bytes = best_na * P.a_entry_bytes + best_nb * P.b_entry_bytes;
    }
    if (share == 0) break;
  }
  if (fits && P.wide && !wide_forced &&
      !(P.b_resident || (P.tap_slabs_per_tile <= 8 && (d->res_mode != IOU_RES_NONE || d->block_n <= 128))))
    fits = false;                                        // operand-bound: 8 epilogue warps and deeper rings (measured)
  }
  if (P.wide) P.res_prefetch = 0;
  if (!fits) { delete plan; return fail(IOU_ERR_INVALID, "tile does not fit shared memory (block_n %d, residual %d): use a smaller block_n", d->block_n, d->res_mode); }
  if (getenv("IOU_CONV_DEBUG"))
    fprintf(stderr, "[iou_conv] cin %d cout %d taps %d block_n %d pair %d res %d staged %d wide %d | groups %d a_rows %d nA %d nB %d%s ring %d KB\n",
            d->cin, d->cout, d->num_taps, d->block_n, P.two_cta, d->res_mode, P.staged, P.wide, P.num_groups, P.a_rows,
            P.num_a_stages, P.num_b_stages, P.b_resident ? " (resident)" : "", P.ring_bytes / 1024);
  P.scale = d->scale; P.shift = d->shift; P.relu = d->relu; P.res_mode = d->res_mode;
  P.residual = (const __nv_bfloat16*)d->residual;
  P.out_mode = d->out_mode; P.out = (__nv_bfloat16*)d->out; P.dense_split = d->dense_split;
  // cute::UMMA::InstrDescriptor: c=F32 (1<<4), a=b=BF16 (1<<7, 1<<10), K-major both, N>>3 @17, M>>4 @24
  P.combine = (d->passes == 3 && !P.two_cta && d->block_n <= 80 && !getenv("IOU_NO_COMBINE")) ? 1 : 0;
  // f8: a/b format fields 0 = F16 (kind::f16) and 0 = E4M3 (kind::f8f6f4): one descriptor serves both MMAs
  const uint32_t fmt_bits = P.f8 ? 0u : ((1u << 7) | (1u << 10));
  P.idesc2 = (1u << 4) | fmt_bits | ((uint32_t)((2 * d->block_n) >> 3) << 17) | ((uint32_t)(kBlockM >> 4) << 24);
  P.idesc = (1u << 4) | fmt_bits | ((uint32_t)(d->block_n >> 3) << 17) |
            ((uint32_t)((P.two_cta ? 2 * kBlockM : kBlockM) >> 4) << 24);
  for (int i = 0; i < d->num_src; ++i) {
    if (!d->src[i] || ((uintptr_t)d->src[i] & 15)) { delete plan; return fail(IOU_ERR_INVALID, "src[%d] NULL or misaligned", i); }
    if (int e = encode_2d(&P.tmap_src[i], d->src[i], (uint64_t)d->src_rows, (uint64_t)2 * P.src_cin[i], (uint32_t)P.a_rows)) { delete plan; return e; }
  }
  for (int i = d->num_src; i < IOU_CONV_MAX_SRC; ++i) P.tmap_src[i] = P.tmap_src[0];
  P.tmap_out = P.tmap_w; P.tmap_res = P.tmap_w;
  if (P.staged && !P.phase_only) {
    const uint64_t orow = d->out_rows > 0 ? (uint64_t)d->out_rows : (uint64_t)d->src_rows;
    if (int e = encode_2d(&P.tmap_out, d->out, orow, (uint64_t)2 * d->cout, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B)) { delete plan; return e; }
  }
  if (P.res_staged) {
    const uint64_t rrow = d->res_rows > 0 ? (uint64_t)d->res_rows : (uint64_t)d->src_rows;
    if (((uintptr_t)d->residual & 15) != 0) { delete plan; return fail(IOU_ERR_INVALID, "residual must be 16-byte aligned"); }
    if (int e = encode_2d(&P.tmap_res, d->residual, rrow, (uint64_t)2 * d->cout, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B)) { delete plan; return e; }
  }
  if (int e = encode_2d(&P.tmap_w, d->weight, (uint64_t)d->num_taps * d->cout_pad, (uint64_t)2 * P.b_cin, (uint32_t)(P.two_cta ? d->block_n / 2 : d->block_n))) { delete plan; return e; }
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (getenv("IOU_MAX_SMS") && atoi(getenv("IOU_MAX_SMS")) > 1 && atoi(getenv("IOU_MAX_SMS")) < sms) sms = atoi(getenv("IOU_MAX_SMS"));   // (experiments)
  if (P.two_cta) {
    const int pairs = sms / 2;
    plan->grid = 2 * (P.total_pair_tiles < pairs ? P.total_pair_tiles : pairs);
  } else {
    plan->grid = P.total_tiles < sms ? P.total_tiles : sms;
  }
  plan->smem_bytes = (size_t)kCtrlBytes + 1024 + (size_t)P.ring_bytes + (size_t)epi_warps * P.staging_per_warp;
  double k_total = 0;
  for (int t = 0; t < d->num_taps; ++t) k_total += P.diag_k ? kBlockK : (ksplit > 1 ? d->cin : P.src_cin[d->tap_src[t]]);
  plan->flops = 2.0 * real_rows * d->cout * k_total;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_tap_gemm_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(conv_tap_gemm_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(conv_tap_gemm_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(conv_tap_gemm_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(conv_tap_gemm_kernel<false, true, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(conv_tap_gemm_kernel<true, true, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget);
    if (e != cudaSuccess) { delete plan; return fail(IOU_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e)); }
    attr_set = true;
  }
  *plan_out = plan;
  return IOU_OK;
}

extern "C" int iou_conv_run(const iou_conv_plan* plan, void* stream) {
  IOU_REQUIRE(plan != nullptr, "plan is NULL");
  if (plan->chained) return iou::launch_chain(plan, stream);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(plan->grid);
  cfg.blockDim = dim3(plan->params.wide ? kNumThreadsWide : kNumThreads);
  cfg.dynamicSmemBytes = plan->smem_bytes;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (plan->params.two_cta) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 2; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (plan->params.pdl) {        // may start while the previous kernel of the stream drains (it waits in griddepcontrol.wait)
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr; cfg.numAttrs = na;
  cudaError_t e;
  if (plan->params.wide)
    e = plan->params.two_cta ? cudaLaunchKernelEx(&cfg, conv_tap_gemm_kernel<true, true, 3>, plan->params)
                             : cudaLaunchKernelEx(&cfg, conv_tap_gemm_kernel<false, true, 3>, plan->params);
  else if (plan->params.two_cta)
    e = plan->params.f8 ? cudaLaunchKernelEx(&cfg, conv_tap_gemm_kernel<true, true>, plan->params)
                        : cudaLaunchKernelEx(&cfg, conv_tap_gemm_kernel<true, false>, plan->params);
  else
    e = plan->params.f8 ? cudaLaunchKernelEx(&cfg, conv_tap_gemm_kernel<false, true>, plan->params)
                        : cudaLaunchKernelEx(&cfg, conv_tap_gemm_kernel<false, false>, plan->params);
  if (e != cudaSuccess) return fail(IOU_ERR_CUDA, "conv_tap_gemm_kernel launch failed: %s", cudaGetErrorString(e));
  return IOU_OK;
}

extern "C" void iou_conv_plan_destroy(iou_conv_plan* plan) { delete plan; }

extern "C" double iou_conv_plan_flops(const iou_conv_plan* plan) { return plan ? plan->flops : 0.0; }

extern "C" int iou_conv_plan_epilogue_warps(const iou_conv_plan* plan) {
  return plan ? (plan->params.wide ? kNumEpiWarpsWide : kNumEpiWarps) : 0;
}
