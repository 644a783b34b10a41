// Shared pieces of the tcgen05 conv kernels (conv_tc.cu: the general tap GEMM; conv_chain.cu: two chained 1x1 convs in one
// launch): tile constants, the per-launch parameter block, the PTX wrappers (mbarrier, TMA, tcgen05, PDL) and the plan object.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cstdio>
#include <cstdlib>
#include <new>
#include "common.cuh"
#include "split_fmt.cuh"

namespace iou {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;                  // bf16 elements = one 128-byte swizzle row
constexpr int kATileBytes = kBlockM * kBlockK * 2;   // 16 KiB
constexpr int kNumThreads = 320;             // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue (two per TMEM lane quadrant)
constexpr int kNumEpiWarps = 8;
// "wide" variant: 16 warps = warpgroup 0 (warp 0 TMA, warp 1 MMA, warps 2-3 idle) + 12 epilogue warps (three per TMEM lane
// quadrant); setmaxnreg moves registers from warpgroup 0 to the epilogue warpgroups (56 / 152 per thread)
constexpr int kNumThreadsWide = 512;
constexpr int kNumEpiWarpsWide = 12;
constexpr uint32_t kBarResWide = 704;        // control-block offset of the wide variant's 24 residual barriers
constexpr int kMaxStages = 8;
constexpr int kMaxBStages = 16;              // B ring entries (resident weights: one entry per (tap, K slab) of a tile)
constexpr int kAccStride = 256;              // TMEM columns between the two accumulator stages
constexpr int kTmemCols = 512;
constexpr int kSmemBudget = 227 * 1024;
constexpr int kCtrlBytes = 1024;

struct SegDev { int row_start, n_img, h, w; };

struct ConvParams {
  CUtensorMap tmap_src[IOU_CONV_MAX_SRC];
  CUtensorMap tmap_w;
  CUtensorMap tmap_out;     // padded-rows output, box 64 cols x 32 rows (TMA store)
  CUtensorMap tmap_res;     // residual (same geometry), box 64 cols x 32 rows (TMA load)
  int cin, cout, cout_pad, block_n, num_taps, k_slabs, passes, lolo;
  int diag_k, b_cin;        // grouped conv: N tile j contracts only input channels [64j, 64j+64)
  int tap_src[IOU_CONV_MAX_TAPS], tap_dy[IOU_CONV_MAX_TAPS], tap_dx[IOU_CONV_MAX_TAPS];
  int num_seg;
  SegDev seg[IOU_CONV_MAX_SEG];
  int seg_tile_off[IOU_CONV_MAX_SEG + 1];
  int num_m_tiles, num_n_tiles, total_tiles;
  int two_cta, total_pair_tiles;   // cta_group::2: a CTA pair owns two consecutive 128-row tiles x BLOCK_N
  // A windows and B tiles travel through separate rings: taps that read the same source at the same dy
  // (dx = dx0..dx0+2) share ONE (128+8)-row A window and address it through row-shifted descriptors
  int num_groups;
  int grp_src[IOU_CONV_MAX_TAPS], grp_dy[IOU_CONV_MAX_TAPS], grp_dx0[IOU_CONV_MAX_TAPS], grp_nt[IOU_CONV_MAX_TAPS];
  int grp_tap[IOU_CONV_MAX_TAPS][4], grp_shift[IOU_CONV_MAX_TAPS][4];   // up to 4 taps per window (shift 0..3 rows)
  int a_rows, a_entry_bytes, b_entry_bytes, num_a_stages, num_b_stages, ring_bytes, taps_per_tile;
  // sources may differ in channel count (K-concatenated GEMMs, e.g. conv3(t2) + downsample(x) of a bottleneck's first
  // block in one accumulator): K slabs per tap group and the lo-plane offset follow the group's source
  int src_cin[IOU_CONV_MAX_SRC], grp_ks[IOU_CONV_MAX_TAPS], tap_slabs_per_tile;
  int b_tile_bytes;
  int b_resident;            // one N tile and few (tap, slab) weight tiles: loaded once per CTA, kept for every tile
  int staged, res_staged, staging_per_warp;
  int res_prefetch;          // residual slabs are prefetched into L2 this many tiles ahead (0 = off)
  // fused stride-2 phase split (iou_phase_split's layout, written by the epilogue): row (img, yp, xp) also goes to
  // phase (yp&1, xp&1) at (u, v) = ((yp>>1)+1, (xp>>1)+1) of a [n][ph_h+2][ph_w+2] map; phase_only skips the normal output
  __nv_bfloat16* phase_out[4];
  int phase_any, phase_only, ph_h, ph_w;
  const float* scale;
  const float* shift;
  int relu, res_mode;
  const __nv_bfloat16* residual;
  SegDev res_seg[IOU_CONV_MAX_SEG];
  int out_mode;
  __nv_bfloat16* out;
  float* out_dense[IOU_CONV_MAX_SEG];
  float* out_dense2[IOU_CONV_MAX_SEG];
  int dense_split;
  unsigned int idesc, idesc2;
  int combine;               // narrow N: A_hi x [B_hi|B_lo] as ONE MMA of N = 2*block_n, A_lo x B_hi into a third column block
  int f8;                    // passes == 2: fp16 main pass + e4m3 correction pass (split_fmt.cuh); `scale` = 1 / S_n
  int num_acc;               // TMEM accumulator stages (2)
  int corr_off;              // f8, narrow tiles: the e4m3 MMAs accumulate into a second column block at this offset (0: same block)
  int ksplit_ntiles;         // split-K: N tiles per K split (0 = off); split j = n_tile / ksplit_ntiles reads K slabs [j*grp_ks, (j+1)*grp_ks)
  float* gmax_out[IOU_CONV_MAX_SEG];   // dense outputs: per pixel and group of gmax_cols columns, two partial maxima
  int gmax_cols, gmax_groups;          // columns per group (0 = off), groups in cout
  int wide;                  // 12 epilogue warps (kEpiQ == 3): the HBM-bound convs whose epilogue is issue-latency bound
  int pdl;                   // programmatic dependent launch: prologue overlaps the previous kernel's tail (griddepcontrol)
};

// ------------------------------------------------------------------------------------ PTX
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
// Bounded wait: a pipeline bug must surface as a trapped launch (an error the host sees), never
// as a hung GPU.  The bound (~4 s of wall clock) is far above any legitimate wait in this kernel.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  unsigned long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0x3ffu) == 0) {
      unsigned long long t1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t1 - t0 > 4000000000ull) {
        printf("conv_tap_gemm_kernel: mbarrier wait timed out (block %d thread %d bar %u parity %u)\n",
               (int)blockIdx.x, (int)threadIdx.x, bar, parity);
        __trap();
      }
    }
  }
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* tmap, uint32_t bar, uint32_t dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"(tmap), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
// ---- CTA-pair (cta_group::2) variants.  kPeerMask clears the pair bit of a shared::cluster address, so a
// barrier operand built from a local address names the EVEN (leader) CTA's barrier (CUTLASS Sm100MmaPeerBitMask).
constexpr uint32_t kPeerMask = 0xFEFFFFFFu;
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(const CUtensorMap* tmap, uint32_t bar, uint32_t dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"(tmap), "r"(c0), "r"(c1), "r"(bar & kPeerMask) : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar & kPeerMask) : "memory");
}
__device__ __forceinline__ void tc_commit_pair(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16_pair(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                 uint32_t accumulate) {
  const uint32_t z = 0;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t"
      "}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(z) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_mma_f8_pair(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                               uint32_t accumulate) {
  const uint32_t z = 0;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f8f6f4 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t"
      "}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(z) : "memory");
}
__device__ __forceinline__ void tc_mma_f8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tmap, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
               ::"l"(tmap), "r"(c0), "r"(c1), "r"(src) : "memory");
}
__device__ __forceinline__ void tma_prefetch_l2_2d(const CUtensorMap* tmap, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(tmap), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 r;
  asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr));
  return r;
}
__device__ __forceinline__ void sts128(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// Programmatic dependent launch (PDL): `launch_dependents` lets the NEXT kernel of the stream start its prologue on SMs
// this grid has already left; `wait` blocks until the PREVIOUS grid has completed and its writes are visible.
__device__ __forceinline__ void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, 128-byte swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
// start>>4 | LBO(16B)>>4 <<16 | SBO(1024B: 8 rows x 128B)>>4 <<32 | version=1 <<46 | SWIZZLE_128B(2) <<61
// The start address may sit `shift` rows (shift*128 B) into a 1024-byte swizzle atom (the row-shifted A
// windows of the dx taps): the tensor core applies the 128B XOR swizzle to ABSOLUTE shared-memory address
// bits, exactly as TMA wrote them, so the base_offset field (bits 49..51) stays 0 -- measured on B200:
// base_offset = shift gives wrong sums, 0 is exact (tests/test_gpu_conv.py).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}

// one lane of a converged warp (elect.sync): unlike `lane == 0`, the compiler knows the region is
// single-threaded and emits tcgen05 / TMA instructions without a per-instruction ELECT loop
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ uint32_t umma_desc_lo(uint32_t smem_addr) { return ((smem_addr & 0x3FFFFu) >> 4) | (1u << 16); }
__device__ __forceinline__ uint64_t umma_desc(uint32_t lo) { return ((uint64_t)0x40004040u << 32) | lo; }

__device__ __forceinline__ uint32_t pack_bf16x2(__nv_bfloat16 a, __nv_bfloat16 b) {
  return (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(b) << 16);
}

}  // namespace iou

struct iou_conv_plan {
  iou::ConvParams params;
  int grid;
  size_t smem_bytes;
  double flops;
  // chained launch (conv_chain.cu): params = the first conv, params2 = the 1x1 conv that consumes its output
  int chained;
  iou::ConvParams params2;
  int chain_units, chain_n0, chain_a_stages, chain_b_stages, chain_b_entry_bytes;
};

namespace iou {
int launch_chain(const iou_conv_plan* plan, void* stream);      // conv_chain.cu
}
