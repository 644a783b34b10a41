// Two chained 1x1 convolutions in ONE persistent launch (sm_100a, CTA pairs, fp16 + e4m3 scheme):
//     y = relu(bn3(conv3(t2)) + x)            (the last conv of bottleneck i,      resnet.py:256-265)
//     t1' = relu(bn1(conv1'(y)))              (the first conv of bottleneck i + 1, resnet.py:224-226)
// Layer by layer, conv1' re-reads the 4*planes-channel map y from HBM right after conv3 wrote it -- a quarter of a
// bottleneck's DRAM traffic.  Here a CTA pair owns a pair of 128-row tiles ("unit"): it computes every N tile of conv3 for
// them, stores y through TMA as before, and then runs conv1' on the SAME rows, whose A operand (y) it loads back through
// TMA while the lines are still in L2 (it has just written them itself).  y still goes to HBM once (the next bottleneck's
// residual needs it); its re-read does not.
//
// Same roles as conv_tap_gemm_kernel (conv_tc.cu): warp 0 TMA producer, warp 1 MMA issuer (leader CTA), warps 2..9
// epilogue.  The item sequence of a pair is software-pipelined by one unit so that the producer rarely waits for y:
//     c3(u0, n = 0..N0-1), c3(u1, *), c1(u0), c3(u2, *), c1(u1), ..., c1(u_last)
// and an mbarrier per unit parity ("y of unit j is in global memory") orders conv1'(u_j)'s loads behind the TMA stores of
// conv3(u_j): the epilogue warps wait for their bulk stores to COMPLETE (cp.async.bulk.wait_group 0, not .read), fence
// the async proxy and arrive; the producer waits, fences and loads.  Every CTA only reads back rows it wrote itself.
//
// DIRECT mode (conv3 with ONE N tile, i.e. layer 1; ChainParams.direct): y is not reloaded at all.  The conv3 epilogue also
// writes each 64-channel slab of y into an A-ring stage in the UMMA K-major 128B-swizzled layout (16-byte chunk c of row r
// at c ^ (r & 7) -- what TMA would have produced), fences the async proxy and arrives on a per-stage "direct full" barrier
// of the leader (16 arrivals: the epilogue warps of both CTAs); the MMA warp waits on that barrier for conv1' items and
// releases the stage through the usual commit.  Item order c3(u0), c1(u0), c3(u1), ...; A-stage parities are tracked per
// producer kind (TMA / epilogue), and the TMA thread still waits on the stages it does not fill, so that every parity
// wait stays within one phase of its barrier.
#include "conv_common.cuh"

namespace iou {

struct ChainParams {
  ConvParams p[2];
  int units;                 // 128-row tile pairs
  int n0;                    // N tiles of conv 0 (conv 1 has exactly one)
  int num_a_stages, num_b_stages, a_entry_bytes, b_entry_bytes, ring_bytes, staging_per_warp;
  int b_resident;            // both convs keep all their weight tiles in shared memory (loaded with their first item)
  int b_res_off1;            // byte offset of conv 1's resident entries inside the B area
  int direct;                // n0 == 1: the epilogue of conv 0 writes y's 64-channel slabs straight into the A ring (UMMA layout) --
                             // conv 1 never reloads y; item order c3(u0), c1(u0), c3(u1), c1(u1), ...
};

constexpr uint32_t kBarOutDone = 704;     // control block offsets 704, 712: "y of unit parity 0 / 1 is stored"
constexpr uint32_t kBarDirect = 720;      // 8 barriers: "A stage s holds a slab written by the epilogue warps of both CTAs"

__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

__global__ void __launch_bounds__(kNumThreads, 1) conv_chain_kernel(const __grid_constant__ ChainParams C) {
  constexpr int kFmt = kFmtF16F8;
  extern __shared__ unsigned char smem_dyn[];
  const uint32_t raw = smem_u32(smem_dyn);
  const uint32_t base = (raw + 1023u) & ~1023u;
  unsigned char* ctrl = smem_dyn + (base - raw);
  const uint32_t ctrl_addr = base;
  const uint32_t tiles_addr = base + kCtrlBytes;
  const uint32_t bar_full = ctrl_addr + 448, bar_empty = ctrl_addr + 576, bar_tfull = ctrl_addr + 128,
                 bar_tempty = ctrl_addr + 144, bar_afull = ctrl_addr + 320, bar_aempty = ctrl_addr + 384,
                 bar_outdone = ctrl_addr + kBarOutDone, bar_dfull = ctrl_addr + kBarDirect;
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(ctrl + 160);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = (int)cluster_ctarank();
  const int u_first = (int)(blockIdx.x >> 1), u_stride = (int)(gridDim.x >> 1);
  const int n0 = C.n0;
  const int U = u_first < C.units ? (C.units - u_first + u_stride - 1) / u_stride : 0;     // units of this pair
  const int total_items = U * (n0 + 1);
  // item s of this pair -> (conv k, unit q, N tile n, unit sequence number j)
  auto item_at = [&](int s, int& k, int& q, int& n, int& j) {
    if (C.direct) { k = s & 1; j = s >> 1; n = 0; }
    else if (s < n0) { k = 0; j = 0; n = s; }
    else {
      const int s2 = s - n0, blk = s2 / (n0 + 1), r = s2 - blk * (n0 + 1);
      if (blk + 1 < U && r < n0) { k = 0; j = blk + 1; n = r; }
      else { k = 1; j = (blk + 1 < U) ? blk : U - 1; n = 0; }
    }
    q = u_first + j * u_stride;
  };
  auto tile_rows = [&](const ConvParams& P, int q, int& m_tile, int& sidx) {
    m_tile = 2 * q + rank;
    sidx = 0;
    while (sidx + 1 < P.num_seg && m_tile >= P.seg_tile_off[sidx + 1]) ++sidx;
  };

  if (warp == 0 && lane == 0) {
    const int nb_bars = C.b_resident ? kMaxBStages : C.num_b_stages;
    for (int s = 0; s < nb_bars; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
    for (int s = 0; s < C.num_a_stages; ++s) { mbar_init(bar_afull + 8 * s, 1); mbar_init(bar_aempty + 8 * s, 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(bar_tfull + 8 * a, 1); mbar_init(bar_tempty + 8 * a, 2 * kNumEpiWarps * 32); }
    for (int r = 0; r < 2 * kNumEpiWarps; ++r) mbar_init(ctrl_addr + 192 + 8 * r, 1);
    for (int a = 0; a < 2; ++a) mbar_init(bar_outdone + 8 * a, kNumEpiWarps);
    for (int a = 0; a < kMaxStages; ++a) mbar_init(bar_dfull + 8 * a, 2 * kNumEpiWarps);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    for (int k = 0; k < 2; ++k) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&C.p[k].tmap_src[0]) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&C.p[k].tmap_w) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&C.p[k].tmap_out) : "memory");
    }
    asm volatile("prefetch.tensormap [%0];" ::"l"(&C.p[0].tmap_res) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(ctrl_addr + 160), "r"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const uint32_t a_lo_off = (uint32_t)kBlockM * 128u;                 // A entry: hi tile | lo tile (128 rows each)
  const uint32_t b_area = tiles_addr + (uint32_t)(C.num_a_stages * C.a_entry_bytes);

  if (warp == 0) {
    // =============================== TMA producer ===============================
    if (elect_one()) {
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      bool b_loaded[2] = {false, false};
      for (int s = 0; s < total_items; ++s) {
        int k, q, n, j;
        item_at(s, k, q, n, j);
        const ConvParams& P = C.p[k];
        int m_tile, sidx;
        tile_rows(P, q, m_tile, sidx);
        const int row0 = P.seg[sidx].row_start + (m_tile - P.seg_tile_off[sidx]) * kBlockM;
        const int b_rows = P.block_n >> 1;
        const int ksl = P.cin / kBlockK;
        const bool direct_item = C.direct && k == 1;    // its A slabs are written by the epilogue warps, not loaded
        if (k == 1 && !C.direct) {                      // y of this unit must be in global memory
          mbar_wait(bar_outdone + 8 * (j & 1), (uint32_t)(j >> 1) & 1u);
          fence_proxy_async_all();
        }
        for (int ks = 0; ks < ksl; ++ks) {
          if (!direct_item) {
            mbar_wait(bar_aempty + 8 * as, aph ^ 1u);
            const uint32_t fa = bar_afull + 8 * as;
            const uint32_t sa = tiles_addr + as * C.a_entry_bytes;
            if (rank == 0) mbar_expect_tx(fa, 2u * (uint32_t)C.a_entry_bytes);
            tma_load_2d_pair(&P.tmap_src[0], fa, sa, ks * kBlockK, row0);
            tma_load_2d_pair(&P.tmap_src[0], fa, sa + a_lo_off, P.cin + ks * kBlockK, row0);
          } else {
            // not loaded here, but observed: every use of a stage is waited for in order, so that a parity wait is never
            // more than one phase behind its barrier
            mbar_wait(bar_aempty + 8 * as, aph ^ 1u);
          }
          if (++as == C.num_a_stages) { as = 0; aph ^= 1u; }
          const int wrow = n * P.block_n + rank * b_rows;
          if (C.b_resident) {
            if (!b_loaded[k]) {
              const int e = (k ? C.p[0].cin / kBlockK : 0) + ks;          // entry index == barrier index
              const uint32_t fb = bar_full + 8 * e;
              const uint32_t sb = b_area + (k ? (uint32_t)C.b_res_off1 : 0u) + (uint32_t)(ks * P.b_entry_bytes);
              if (rank == 0) mbar_expect_tx(fb, 2u * (uint32_t)P.b_entry_bytes);
              tma_load_2d_pair(&P.tmap_w, fb, sb, ks * kBlockK, wrow);
              tma_load_2d_pair(&P.tmap_w, fb, sb + (uint32_t)P.b_tile_bytes, P.b_cin + ks * kBlockK, wrow);
            }
          } else {
            mbar_wait(bar_empty + 8 * bs, bph ^ 1u);
            const uint32_t fb = bar_full + 8 * bs;
            const uint32_t sb = b_area + bs * C.b_entry_bytes;
            if (rank == 0) mbar_expect_tx(fb, 2u * (uint32_t)P.b_entry_bytes);
            tma_load_2d_pair(&P.tmap_w, fb, sb, ks * kBlockK, wrow);
            tma_load_2d_pair(&P.tmap_w, fb, sb + (uint32_t)P.b_tile_bytes, P.b_cin + ks * kBlockK, wrow);
            if (++bs == C.num_b_stages) { bs = 0; bph ^= 1u; }
          }
        }
        b_loaded[k] = true;
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer (leader CTA only) ===============================
    int as = 0, bs = 0;
    uint32_t bph = 0;
    uint32_t tma_par = 0, dir_par = 0;               // per A stage: parity of its next TMA-filled / epilogue-filled use
    bool b_seen[2] = {false, false};
    const uint32_t a_lo_d = a_lo_off >> 4;
    for (int s = 0; s < total_items && rank == 0; ++s) {
      int k, q, n, j;
      item_at(s, k, q, n, j);
      const ConvParams& P = C.p[k];
      const int acc = s & 1;
      const uint32_t acc_phase = (uint32_t)(s >> 1) & 1u;
      mbar_wait(bar_tempty + 8 * acc, acc_phase ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * kAccStride);
      const int ksl = P.cin / kBlockK;
      const uint32_t corr_off = (uint32_t)P.corr_off, idesc = P.idesc, b_lo_d = (uint32_t)P.b_tile_bytes >> 4;
      const bool direct_item = C.direct && k == 1;
      for (int ks = 0; ks < ksl; ++ks) {
        if (direct_item) { mbar_wait(bar_dfull + 8 * as, (dir_par >> as) & 1u); dir_par ^= 1u << as; }
        else { mbar_wait(bar_afull + 8 * as, (tma_par >> as) & 1u); tma_par ^= 1u << as; }
        const uint32_t sa = tiles_addr + as * C.a_entry_bytes;
        uint32_t sb;
        if (C.b_resident) {
          const int e = (k ? C.p[0].cin / kBlockK : 0) + ks;
          if (!b_seen[k]) mbar_wait(bar_full + 8 * e, 0u);
          sb = b_area + (k ? (uint32_t)C.b_res_off1 : 0u) + (uint32_t)(ks * P.b_entry_bytes);
        } else {
          mbar_wait(bar_full + 8 * bs, bph);
          sb = b_area + bs * C.b_entry_bytes;
        }
        tc_fence_after();
        if (elect_one()) {
          const uint32_t da = umma_desc_lo(sa), db = umma_desc_lo(sb);
          const uint32_t dal = da + a_lo_d, dbl = db + b_lo_d;
          const uint32_t first = ks > 0 ? 1u : 0u;
          const uint32_t cfirst = corr_off ? first : 1u;
#pragma unroll
          for (uint32_t kk = 0; kk < kBlockK / 16; ++kk) {
            tc_mma_bf16_pair(d_tmem, umma_desc(da + 2 * kk), umma_desc(db + 2 * kk), idesc, kk ? 1u : first);
            tc_mma_f8_pair(d_tmem + corr_off, umma_desc(dal + 2 * kk), umma_desc(dbl + 2 * kk), idesc, kk ? 1u : cfirst);
          }
          if (!C.b_resident) tc_commit_pair(bar_empty + 8 * bs);
          tc_commit_pair(bar_aempty + 8 * as);
          if (ks == ksl - 1) tc_commit_pair(bar_tfull + 8 * acc);
        }
        __syncwarp();
        if (!C.b_resident) { if (++bs == C.num_b_stages) { bs = 0; bph ^= 1u; } }
        if (++as == C.num_a_stages) as = 0;
      }
      b_seen[k] = true;
    }
  } else {
    // =============================== epilogue ===============================
    const int lane_group = warp & 3;
    const int m_local = lane_group * 32 + lane;
    const int ew = warp - 2;
    const int half = ew >> 2;
    const uint32_t st_out = tiles_addr + C.ring_bytes + ew * C.staging_per_warp;
    const uint32_t st_res = st_out + 4096;
    const uint32_t bar_res = ctrl_addr + 192 + ew * 16;
    const ConvParams& P0 = C.p[0];
    const bool res = P0.res_staged != 0;
    auto res_row_of = [&](int q_) {
      int mt, sx;
      tile_rows(P0, q_, mt, sx);
      return P0.seg[sx].row_start + (mt - P0.seg_tile_off[sx]) * kBlockM + lane_group * 32;
    };
    auto issue_res_at = [&](int row, int col, int q_) {      // one elected lane only
      const uint32_t bar = bar_res + 8 * (q_ & 1), dst = st_res + (q_ & 1) * 4096;
      mbar_expect_tx(bar, 4096u);
      tma_load_2d(&P0.tmap_res, bar, dst, col, row);
      tma_load_2d(&P0.tmap_res, bar, dst + 2048, P0.cout + col, row);
    };
    // the next conv-0 item after item s (same unit: next N tile; else the next unit's first N tile), if any
    auto next_c0 = [&](int s, int& q2, int& n2) {
      for (int t = s + 1; t < total_items; ++t) {
        int k2, j2;
        item_at(t, k2, q2, n2, j2);
        if (k2 == 0) return true;
      }
      return false;
    };
    int rq = 0;                                            // running slab counter of the residual ring
    int ring_pos = 0;                                      // A-ring position (in slabs) of the item being processed
    const int ksl0 = C.p[0].cin / kBlockK, ksl1 = C.p[1].cin / kBlockK;
    if (res && total_items > 0 && elect_one()) issue_res_at(res_row_of(u_first), half * 32, 0);
    for (int s = 0; s < total_items; ++s) {
      int k, q, n, j;
      item_at(s, k, q, n, j);
      const ConvParams& P = C.p[k];
      const int acc = s & 1;
      const uint32_t acc_phase = (uint32_t)(s >> 1) & 1u;
      int m_tile, sidx;
      tile_rows(P, q, m_tile, sidx);
      __syncwarp();
      const bool tile_valid = m_tile < P.num_m_tiles;
      const SegDev sg = P.seg[sidx];
      const int grow = sg.row_start + (m_tile - P.seg_tile_off[sidx]) * kBlockM + m_local;
      const int wp = sg.w + 2, plane = (sg.h + 2) * wp;
      const int rel = grow - sg.row_start;
      const int img = rel / plane, rem = rel - img * plane;
      const int yp = rem / wp, xp = rem - yp * wp;
      const bool interior = tile_valid && (img < sg.n_img) && (yp >= 1) && (yp <= sg.h) && (xp >= 1) && (xp <= sg.w);
      mbar_wait(bar_tfull + 8 * acc, acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(lane_group * 32) << 16) + (uint32_t)(acc * kAccStride);
      const int row_tile0 = grow - lane;
      const int n_groups = P.block_n >> 5;
      const uint32_t swz = (uint32_t)((lane >> 1) & 3);
      const bool use_res = res && k == 0;
      const bool feed = C.direct && k == 0;                  // this item's output slabs are conv 1's A operand
      const int feed_pos = ring_pos + ksl0;                  // ring position of conv 1's first slab (behind this item's own)
      ring_pos += (k == 0) ? ksl0 : ksl1;
      for (int g = half; g < n_groups; g += 2) {
        const int c0 = n * P.block_n + g * 32;
        if (use_res) {
          if (elect_one()) {                               // prefetch this warp's next residual slab
            int q2, n2;
            if (g + 2 < n_groups) issue_res_at(row_tile0, n * P.block_n + (g + 2) * 32, rq + 1);
            else if (next_c0(s, q2, n2)) issue_res_at(res_row_of(q2), n2 * P0.block_n + half * 32, rq + 1);
          }
          mbar_wait(bar_res + 8 * (rq & 1), (uint32_t)(rq >> 1) & 1u);
        }
        uint32_t v[32], wc[32];
        tc_ld32(t_row + g * 32, v);
        if (P.corr_off) tc_ld32(t_row + P.corr_off + g * 32, wc);
        float shv[32];
        if (P.shift) {
          const float4* sp = reinterpret_cast<const float4*>(P.shift + c0);
#pragma unroll
          for (int t = 0; t < 8; ++t) {
            const float4 t4 = __ldg(sp + t);
            shv[4 * t] = t4.x; shv[4 * t + 1] = t4.y; shv[4 * t + 2] = t4.z; shv[4 * t + 3] = t4.w;
          }
        } else {
#pragma unroll
          for (int t = 0; t < 32; ++t) shv[t] = 0.f;
        }
        tc_wait_ld();
        if (P.corr_off) {
#pragma unroll
          for (int t = 0; t < 32; ++t) v[t] = __float_as_uint(__uint_as_float(v[t]) + __uint_as_float(wc[t]));
        }
        float f[32];
        {
          const float4* cp = reinterpret_cast<const float4*>(P.scale + c0);
#pragma unroll
          for (int t = 0; t < 8; ++t) {
            const float4 t4 = __ldg(cp + t);
            f[4 * t] = fmaf(__uint_as_float(v[4 * t]), t4.x, shv[4 * t]);
            f[4 * t + 1] = fmaf(__uint_as_float(v[4 * t + 1]), t4.y, shv[4 * t + 1]);
            f[4 * t + 2] = fmaf(__uint_as_float(v[4 * t + 2]), t4.z, shv[4 * t + 2]);
            f[4 * t + 3] = fmaf(__uint_as_float(v[4 * t + 3]), t4.w, shv[4 * t + 3]);
          }
        }
        if (use_res) {
          const uint32_t rb = st_res + (rq & 1) * 4096 + lane * 64;
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const uint32_t sw = (uint32_t)((t ^ swz) << 4);
            decode8_add<kFmt>(lds128(rb + sw), lds128(rb + 2048 + sw), f + t * 8);
          }
          ++rq;
        }
        if (lane == 0) tma_store_wait_read();              // the previous slab has left the staging tile
        __syncwarp();
        const uint32_t ob = st_out + lane * 64;
        uint32_t feed_row = 0, feed_bar = 0;
        if (feed) {
          // slab sl = g / 2 of conv 1's K: stage and parity from its ring position; the stage is free once the MMAs of
          // the slab that used it before have retired
          const int pos = feed_pos + (g >> 1);
          const int stage = pos % C.num_a_stages;
          mbar_wait(bar_aempty + 8 * stage, (((uint32_t)(pos / C.num_a_stages)) & 1u) ^ 1u);
          feed_row = tiles_addr + stage * C.a_entry_bytes + (uint32_t)m_local * 128u;
          feed_bar = bar_dfull + 8 * stage;
        }
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          float x8[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) x8[e] = P.relu ? fmaxf(f[t * 8 + e], 0.f) : f[t * 8 + e];
          uint4 hi, lo;
          encode8<kFmt, false>(x8, hi, lo);
          if (!interior) { hi = make_uint4(0, 0, 0, 0); lo = hi; }
          const uint32_t sw = (uint32_t)((t ^ swz) << 4);
          sts128(ob + sw, hi);
          sts128(ob + 2048 + sw, lo);
          if (feed) {                                      // K-major 128B-swizzled A tile: 16-byte chunk c of row r sits at c ^ (r & 7)
            const uint32_t chunk = (uint32_t)((g & 1) * 4 + t);
            const uint32_t fo = feed_row + ((chunk ^ ((uint32_t)m_local & 7u)) << 4);
            sts128(fo, hi);
            sts128(fo + a_lo_off, lo);
          }
        }
        fence_async_smem();
        __syncwarp();
        // (a .release.cluster arrive costs MEMBAR.GPU + ERRBAR per slab, a quarter of the kernel's stall samples; the
        // writes are already ordered for the async proxy of THIS SM -- which is the one that reads them, each CTA's
        // tensor core fetches its own half of A -- by the proxy fence above)
        if (feed && lane == 0) mbar_arrive_leader(feed_bar);
        if (tile_valid && lane == 0) {
          tma_store_2d(&P.tmap_out, st_out, c0, row_tile0);
          tma_store_2d(&P.tmap_out, st_out + 2048, P.cout + c0, row_tile0);
          tma_store_commit();
        }
      }
      tc_fence_before();
      mbar_arrive_leader(bar_tempty + 8 * acc);
      if (k == 0 && n == n0 - 1 && !C.direct) {
        // y of unit j is complete for this warp's rows once its bulk stores have been PERFORMED (not only read)
        if (lane == 0) {
          tma_store_wait_all();
          fence_proxy_async_all();
          __threadfence();
          mbar_arrive(bar_outdone + 8 * (j & 1));
        }
        __syncwarp();
      }
    }
    if (lane == 0) tma_store_wait_read();
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

}  // namespace iou

using namespace iou;

// Builds the chained plan from two ordinary plans (already validated, tensor maps encoded).  Takes ownership of both
// on success; on failure both are left untouched.
extern "C" int iou_conv_chain_plan_create(iou_conv_plan* first, iou_conv_plan* second, iou_conv_plan** plan_out) {
  IOU_REQUIRE(first && second && plan_out, "NULL argument");
  const ConvParams &A = first->params, &B = second->params;
  IOU_REQUIRE(!first->chained && !second->chained, "plans are already chained");
  IOU_REQUIRE(!first->params.wide && !second->params.wide, "chain: plans of the wide variant (create them with IOU_WIDE=0)");
  IOU_REQUIRE(A.two_cta && B.two_cta && A.f8 && B.f8, "chain: both convs must run as CTA pairs with passes == 2");
  IOU_REQUIRE(A.num_taps == 1 && B.num_taps == 1 && A.tap_dy[0] == 0 && A.tap_dx[0] == 0 && B.tap_dy[0] == 0 && B.tap_dx[0] == 0 &&
              A.tap_src[0] == 0 && B.tap_src[0] == 0, "chain: both convs must be plain 1x1 convs (one tap, one source)");
  IOU_REQUIRE(!A.diag_k && !B.diag_k && !A.ksplit_ntiles && !B.ksplit_ntiles && !A.phase_any && !B.phase_any,
              "chain: no grouped / split-K / phase-writing convs");
  IOU_REQUIRE(A.staged && B.staged && B.res_mode == IOU_RES_NONE && (A.res_mode == IOU_RES_NONE || A.res_staged),
              "chain: padded-rows outputs; only the first conv may add a (same-geometry) residual");
  IOU_REQUIRE(A.src_cin[0] == A.cin && B.src_cin[0] == B.cin && B.cin == A.cout, "chain: conv 2 contracts conv 1's output channels");
  IOU_REQUIRE(B.num_n_tiles == 1, "chain: conv 2 must have a single N tile");
  IOU_REQUIRE(A.a_rows == kBlockM && B.a_rows == kBlockM, "chain: 128-row A tiles");
  IOU_REQUIRE(A.num_seg == B.num_seg && A.num_m_tiles == B.num_m_tiles, "chain: both convs must cover the same rows");
  for (int s = 0; s < A.num_seg; ++s)
    IOU_REQUIRE(A.seg[s].row_start == B.seg[s].row_start && A.seg[s].n_img == B.seg[s].n_img && A.seg[s].h == B.seg[s].h &&
                A.seg[s].w == B.seg[s].w, "chain: segment %d differs", s);
  static_assert(sizeof(ChainParams) <= 16 * 1024, "kernel parameter block too large");
  const int staging = 4096 + (A.res_staged ? 8192 : 0);
  const int budget = kSmemBudget - kCtrlBytes - 1024 - kNumEpiWarps * staging;
  const int a_entry = 2 * kBlockM * kBlockK * 2;                                    // hi + lo, 32 KB
  const int e0 = A.cin / kBlockK, e1 = B.cin / kBlockK;
  int b_res = 0, na = 0, nb = 0, b_entry = A.b_entry_bytes > B.b_entry_bytes ? A.b_entry_bytes : B.b_entry_bytes;
  const int res_bytes = e0 * A.b_entry_bytes + e1 * B.b_entry_bytes;
  if (A.num_n_tiles == 1 && e0 + e1 <= kMaxBStages && !getenv("IOU_CHAIN_NO_RESIDENT")) {
    const int n = (budget - res_bytes) / a_entry;
    if (n >= 2) { b_res = 1; na = n > kMaxStages ? kMaxStages : n; }
  }
  if (!b_res) {
    for (int n = 2; n <= kMaxStages; ++n) {            // balance the two rings
      const int m = (budget - n * a_entry) / b_entry;
      if (m < 2) break;
      if ((m < n ? m : n) > (nb < na ? nb : na)) { na = n; nb = m > kMaxStages ? kMaxStages : m; }
    }
    if (na < 2 || nb < 2) return fail(IOU_ERR_INVALID, "chain: the rings do not fit shared memory");
  }
  first->chained = 1;
  first->params2 = B;
  first->chain_units = (A.num_m_tiles + 1) / 2;
  first->chain_n0 = A.num_n_tiles;
  first->chain_a_stages = na;
  first->chain_b_stages = b_res ? -(e0) : nb;               // negative: resident, |value| = conv 1's entry count
  first->chain_b_entry_bytes = b_entry;
  const int ring = na * a_entry + (b_res ? res_bytes : nb * b_entry);
  first->smem_bytes = (size_t)kCtrlBytes + 1024 + (size_t)ring + (size_t)kNumEpiWarps * staging;
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int pairs = sms / 2;
  first->grid = 2 * (first->chain_units < pairs ? first->chain_units : pairs);
  first->flops += second->flops;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget);
    if (e != cudaSuccess) { first->chained = 0; return fail(IOU_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e)); }
    attr_set = true;
  }
  if (getenv("IOU_CONV_DEBUG"))
    fprintf(stderr, "[iou_conv_chain] %d -> %d -> %d  units %d n0 %d | nA %d nB %d%s ring %d KB smem %zu\n", A.cin, A.cout, B.cout,
            first->chain_units, first->chain_n0, na, b_res ? e0 + e1 : nb, b_res ? " (resident)" : "", ring / 1024, first->smem_bytes);
  delete second;
  *plan_out = first;
  return IOU_OK;
}

namespace iou {
int launch_chain(const iou_conv_plan* plan, void* stream) {
  ChainParams C;
  C.p[0] = plan->params;
  C.p[1] = plan->params2;
  C.units = plan->chain_units;
  C.n0 = plan->chain_n0;
  C.num_a_stages = plan->chain_a_stages;
  C.a_entry_bytes = 2 * kBlockM * kBlockK * 2;
  C.b_entry_bytes = plan->chain_b_entry_bytes;
  C.b_resident = plan->chain_b_stages < 0 ? 1 : 0;
  C.num_b_stages = C.b_resident ? 0 : plan->chain_b_stages;
  const int e0 = plan->params.cin / kBlockK, e1 = plan->params2.cin / kBlockK;
  C.b_res_off1 = e0 * plan->params.b_entry_bytes;
  C.ring_bytes = C.num_a_stages * C.a_entry_bytes +
                 (C.b_resident ? e0 * plan->params.b_entry_bytes + e1 * plan->params2.b_entry_bytes : C.num_b_stages * C.b_entry_bytes);
  C.staging_per_warp = 4096 + (plan->params.res_staged ? 8192 : 0);
  static const bool no_direct = getenv("IOU_CHAIN_NO_DIRECT") != nullptr;
  // direct hand-over needs conv 0's whole N in one item (its 64-channel slabs are conv 1's K slabs) and, for the parity
  // waits of the epilogue warps, no more than two ring wraps inside one conv-1 item
  C.direct = (!no_direct && C.n0 == 1 && C.b_resident && e1 <= 2 * C.num_a_stages) ? 1 : 0;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(plan->grid);
  cfg.blockDim = dim3(kNumThreads);
  cfg.dynamicSmemBytes = plan->smem_bytes;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, conv_chain_kernel, C);
  if (e != cudaSuccess) return fail(IOU_ERR_CUDA, "conv_chain_kernel launch failed: %s", cudaGetErrorString(e));
  return IOU_OK;
}
}  // namespace iou
