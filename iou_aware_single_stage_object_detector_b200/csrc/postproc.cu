// Post-processing half of IoUawareRetinaHead.get_bboxes on the device:
//   K1 max_score   : streaming pass over the class logits, one fused score per anchor
//   K2 topk        : per (image, level) radix-select + bitonic sort of nms_pre anchors
//   K3 gather      : decode (delta2bbox) + full class scores of the selected candidates
//   K4 class_nms   : one CTA per (class, image): threshold, sort, blocked greedy NMS
//   K5 final_select: per image top max_per_img over all classes
// Reference semantics: mmdet/models/anchor_heads/iou_aware_retina_head.py:463-564,
// mmdet/core/bbox/transforms.py:44-78, mmdet/core/post_processing/bbox_nms.py:6-67,
// mmdet/ops/nms/src/nms_kernel.cu:13-131.  Arithmetic that feeds a comparison
// (IoU, decode) uses explicit round-to-nearest intrinsics so no FMA contraction
// can move a threshold decision away from the reference's unfused fp32 result.
#include <cooperative_groups.h>
#include <math.h>
#include "common.cuh"

namespace cg = cooperative_groups;

namespace iou {

// ----------------------------------------------------------------------------------------
struct PostParams {
  int num_levels, A, C, nms_pre, n_img, M, A_total, max_per_img, kcap;
  int feat_h[IOU_MAX_LEVELS], feat_w[IOU_MAX_LEVELS], stride[IOU_MAX_LEVELS];
  int n_anchor[IOU_MAX_LEVELS];      // H*W*A
  int anchor_off[IOU_MAX_LEVELS];    // prefix of n_anchor
  int keep[IOU_MAX_LEVELS];          // min(n_anchor, nms_pre)
  int cand_off[IOU_MAX_LEVELS + 1];  // prefix of keep
  int is_topk[IOU_MAX_LEVELS];
  long long group_off[IOU_MAX_LEVELS + 1];  // K1 work prefix (only top-k levels have width)
  int num_topk_levels;
  int topk_level[IOU_MAX_LEVELS];
  int topk_slot[IOU_MAX_LEVELS];     // inverse of topk_level (-1: the level keeps all its anchors)
  const float* cls[IOU_MAX_LEVELS];
  const float* reg[IOU_MAX_LEVELS];
  const float* iou[IOU_MAX_LEVELS];
  const float* cmax2[IOU_MAX_LEVELS];   // optional: per anchor, two partial maxima of its class logits (iou_get_bboxes_premax)
  float base[IOU_MAX_LEVELS][IOU_MAX_ANCHORS][4];
  float mean[4], stdv[4];
  float alpha, score_thr, iou_thr, max_ratio;
  int rescale, decode_mode;
};

__device__ __forceinline__ float sigmoidf_(float x) { return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x))); }

// score = sigmoid(cls)^alpha * sigmoid(iou)^(1-alpha)   (iou_aware_retina_head.py:510,531)
__device__ __forceinline__ float fuse_score(float cls_logit, float iou_logit, float alpha) {
  const float s = sigmoidf_(cls_logit);
  if (alpha == 1.0f) return s;          // plain RetinaHead: score = sigmoid(cls) (anchor_head.py:404-407)
  const float q = sigmoidf_(iou_logit);
  if (alpha == 0.5f) return __fmul_rn(__fsqrt_rn(s), __fsqrt_rn(q));
  return __fmul_rn(powf(s, alpha), powf(q, 1.0f - alpha));
}

__device__ __forceinline__ float4 ldg_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}

// Monotone bucket of a fused score (scores live in [0, 1]): b(s1) > b(s2) implies s1 > s2, so the per-level top-k is
// "every anchor in a bucket above the boundary bucket + the best of the boundary bucket" (topk_hist_kernel).
#define TOPK_BINS 4096
#define TOPK_HSTRIDE (TOPK_BINS + 16)   // per (image, level): the histogram, then [0] = collected-list length, [1] = chunk ticket
#define TOPK_CHUNKS 8                  // CTAs that share one level's collect pass
__device__ __forceinline__ int score_bucket(float s) {
  return (int)fminf(fmaxf(s * (float)TOPK_BINS, 0.0f), (float)(TOPK_BINS - 1));
}

// ---------------------------------------------------------------------------------------- K1
// One warp per group of 32 consecutive anchors (32*C contiguous floats): coalesced 16-byte
// streaming loads, per-slot max staged in shared memory, then lane a reduces anchor a.  The same pass builds the
// per-(image, level) histogram of score buckets that the top-k selection starts from (warp-aggregated atomics).
__global__ void __launch_bounds__(256) max_score_kernel(const __grid_constant__ PostParams P,
                                                        float* __restrict__ maxscore,
                                                        unsigned int* __restrict__ hist) {
  extern __shared__ float smem_f[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Q = P.C >> 2;
  float* sm = smem_f + warp * 32 * Q;
  const long long total = P.group_off[P.num_levels];
  for (long long g = (long long)blockIdx.x * 8 + warp; g < total; g += (long long)gridDim.x * 8) {
    int l = 0;
    while (g >= P.group_off[l + 1]) ++l;
    const int n_l = P.n_anchor[l];
    const int gpi = (n_l + 31) >> 5;
    const int gl = (int)(g - P.group_off[l]);
    const int img = gl / gpi, gi = gl - img * gpi;
    const int a0 = gi * 32;
    const int v = min(32, n_l - a0);
    const bool premax = P.cmax2[l] != nullptr;            // the class-map producer already reduced the classes
    if (!premax) {
      const float4* src = reinterpret_cast<const float4*>(P.cls[l] + ((size_t)img * n_l + a0) * P.C);
      const int nslots = v * Q;
#pragma unroll 4
      for (int f = lane; f < nslots; f += 32) {
        float4 x = ldg_stream(src + f);
        sm[f] = fmaxf(fmaxf(x.x, x.y), fmaxf(x.z, x.w));
      }
      __syncwarp();
    }
    int bucket = TOPK_BINS + lane;                       // idle lanes: a bucket of their own
    if (lane < v) {
      float m = -INFINITY;
      if (premax) {
        const float2 p2 = __ldg(reinterpret_cast<const float2*>(P.cmax2[l]) + (size_t)img * n_l + a0 + lane);
        m = fmaxf(p2.x, p2.y);
      } else {
        for (int t = 0; t < Q; ++t) m = fmaxf(m, sm[lane * Q + t]);
      }
      const float q = P.iou[l] ? __ldg(P.iou[l] + (size_t)img * n_l + a0 + lane) : 0.f;
      const float sc = fuse_score(m, q, P.alpha);
      maxscore[(size_t)img * P.A_total + P.anchor_off[l] + a0 + lane] = sc;
      bucket = score_bucket(sc);
    }
    const unsigned int peers = __match_any_sync(0xffffffffu, bucket);
    if (lane < v && (__ffs(peers) - 1) == lane)
      atomicAdd(hist + ((size_t)img * P.num_topk_levels + P.topk_slot[l]) * TOPK_HSTRIDE + bucket, (unsigned int)__popc(peers));
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------- sort
// In-place descending bitonic sort of P (power of two) u64 keys in shared memory.
__device__ void bitonic_sort_desc(unsigned long long* a, int P) {
  __syncthreads();
  for (int k = 2; k <= P; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < P; i += blockDim.x) {
        int l = i ^ j;
        if (l > i) {
          unsigned long long x = a[i], y = a[l];
          bool desc = ((i & k) == 0);
          if (desc ? (x < y) : (x > y)) { a[i] = y; a[l] = x; }
        }
      }
      __syncthreads();
    }
  }
}

__device__ __forceinline__ int next_pow2(int x) {
  int p = 1;
  while (p < x) p <<= 1;
  return p;
}

// torch.topk(max_scores, nms_pre) (iou_aware_retina_head.py:544): the selected set is ordered by (score desc, index asc).
#define TOPK_MAX 2048
// Exact radix select over all n keys of one (image, level), one CTA of 1024 threads: the path for score distributions
// the bucket histogram cannot split (e.g. the reference init, where every score is 0.07098 +- 2e-5).
__device__ void topk_radix_select(const float* __restrict__ keys, const int n, const int k,
                                  unsigned long long* sortbuf /* [TOPK_MAX] shared */, int32_t* __restrict__ out) {
  __shared__ unsigned int hist[256];
  __shared__ unsigned int s_prefix, s_mask, s_kleft, s_cnt, s_eq_total, s_eq_run;
  __shared__ unsigned int warp_eq[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  __syncthreads();
  if (tid == 0) { s_prefix = 0; s_mask = 0; s_kleft = k; s_cnt = 0; s_eq_run = 0; }
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = 24 - 8 * pass;
    if (tid < 256) hist[tid] = 0;
    __syncthreads();
    const unsigned int prefix = s_prefix, mask = s_mask;
    for (int base = 0; base < n; base += 4096) {
      unsigned int u4[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {                  // four independent loads in flight per thread
        const int i = base + q * 1024 + tid;
        u4[q] = (i < n) ? float_to_ordered(__ldg(keys + i)) : 0u;
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int i = base + q * 1024 + tid;
        const unsigned int u = u4[q];
        const bool valid = (i < n) && ((u & mask) == prefix);
        const unsigned int d = valid ? ((u >> shift) & 255u) : (256u + lane);
        const unsigned int peers = __match_any_sync(0xffffffffu, d);
        if (valid && (__ffs(peers) - 1) == lane) atomicAdd(&hist[d], __popc(peers));
      }
    }
    __syncthreads();
    if (tid == 0) {
      unsigned int c = 0, kl = s_kleft;
      int d = 255;
      for (; d > 0; --d) {
        if (c + hist[d] >= kl) break;
        c += hist[d];
      }
      s_kleft = kl - c;
      s_prefix = prefix | ((unsigned int)d << shift);
      s_mask = mask | (255u << shift);
      s_eq_total = hist[d];
    }
    __syncthreads();
  }
  const unsigned int T = s_prefix, need_eq = s_kleft;
  const bool ties = (s_eq_total != need_eq);   // more keys equal to T than we may take
  for (int base = 0; base < n; base += 1024) {
    int i = base + tid;
    unsigned int u = (i < n) ? float_to_ordered(keys[i]) : 0u;
    bool gt = (i < n) && (u > T);
    bool eq = (i < n) && (u == T);
    bool take = gt;
    if (!ties) {
      take = gt || eq;
    } else {  // lowest indices first among the ties (block-wide ordered rank)
      unsigned int be = __ballot_sync(0xffffffffu, eq);
      if (lane == 0) warp_eq[warp] = __popc(be);
      __syncthreads();
      unsigned int before = s_eq_run;
      for (int w = 0; w < warp; ++w) before += warp_eq[w];
      unsigned int rank = before + __popc(be & ((1u << lane) - 1u));
      if (eq && rank < need_eq) take = true;
      __syncthreads();
      if (tid == 0) {
        unsigned int t = 0;
        for (int w = 0; w < 32; ++w) t += warp_eq[w];
        s_eq_run += t;
      }
      __syncthreads();
    }
    unsigned int bt = __ballot_sync(0xffffffffu, take);
    unsigned int slot0 = 0;
    if (lane == 0 && bt) slot0 = atomicAdd(&s_cnt, __popc(bt));
    slot0 = __shfl_sync(0xffffffffu, slot0, 0);
    if (take) {
      unsigned int slot = slot0 + __popc(bt & ((1u << lane) - 1u));
      if (slot < TOPK_MAX)
        sortbuf[slot] = ((unsigned long long)u << 32) | (unsigned long long)(0xffffffffu - (unsigned int)i);
    }
  }
  __syncthreads();
  const int Psort = next_pow2(k);
  for (int i = k + tid; i < Psort; i += 1024) sortbuf[i] = 0ull;
  bitonic_sort_desc(sortbuf, Psort);
  for (int r = tid; r < k; r += 1024)
    out[r] = (int32_t)(0xffffffffu - (unsigned int)(sortbuf[r] & 0xffffffffull));
}

// ---------------------------------------------------------------------------------------- K2
// Per (image, level), one CTA: the bucket histogram written by max_score_kernel gives the boundary bucket tb (the
// bucket holding the k-th best score) and the number of anchors above it; ONE pass over the level's keys collects
// every anchor with bucket >= tb (k + a bucket's worth of keys, not n), a bitonic sort orders them by
// (score desc, index asc) and the first k are the level's candidates -- the same set and order as an exact select.
#define TOPK_LIST 4096
#define TOPK_SUB 2048
// position of a score inside its bucket, in TOPK_SUB steps (monotone; s * 4096 and the subtraction are exact)
__device__ __forceinline__ int score_sub_bucket(float s, int bucket) {
  const float t = s * (float)TOPK_BINS - (float)bucket;
  return (int)fminf(fmaxf(t * (float)TOPK_SUB, 0.0f), (float)(TOPK_SUB - 1));
}
// Block-wide search over a histogram split as NB consecutive bins per thread (thread t owns [NB*t, NB*t + NB), the
// top bins belong to the top threads): finds the bin holding the need-th key counted from the top.  Returns through
// shared memory: *s_bin (-1 if the histogram holds fewer than `need` keys) and *s_ge = keys in bins >= that bin.
template <int NB>
__device__ __forceinline__ void find_boundary_bin(const unsigned int (&hv)[NB], const unsigned int need,
                                                  unsigned int* warp_sum, int* s_bin, unsigned int* s_ge) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  unsigned int mine = 0;
#pragma unroll
  for (int q = 0; q < NB; ++q) mine += hv[q];
  unsigned int incl = mine;                                  // inclusive suffix sum inside the warp (higher lanes first)
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned int t = __shfl_down_sync(0xffffffffu, incl, o);
    if (lane + o < 32) incl += t;
  }
  __syncthreads();                                            // warp_sum / s_bin may still be read from an earlier call
  if (lane == 0) warp_sum[warp] = incl;
  if (tid == 0) *s_bin = -1;
  __syncthreads();
  unsigned int above = incl - mine;                          // keys in bins owned by higher threads
  for (int w = warp + 1; w < 32; ++w) above += warp_sum[w];
  if (above < need && above + mine >= need) {                 // the need-th key sits in one of my bins
    unsigned int c = above;
    int b = NB - 1;
    for (; b > 0; --b) {
      if (c + hv[b] >= need) break;
      c += hv[b];
    }
    *s_bin = NB * tid + b;
    *s_ge = c + hv[b];
  }
  __syncthreads();
}

__global__ void __launch_bounds__(1024) topk_hist_kernel(const __grid_constant__ PostParams P,
                                                         const float* __restrict__ maxscore,
                                                         unsigned int* __restrict__ hist,
                                                         unsigned long long* __restrict__ glist,
                                                         int32_t* __restrict__ cand_idx) {
  // grid (level slot, image, chunk): the TOPK_CHUNKS CTAs of one (image, level) each scan a slice of its keys into a
  // shared global list; the CTA that finishes last sorts the list and writes the level's candidates.  Score
  // distributions the histogram cannot split take the single-CTA paths in chunk 0 (the other chunks leave at once).
  __shared__ unsigned long long list[TOPK_LIST];
  __shared__ unsigned int sub[TOPK_SUB];
  __shared__ unsigned int warp_sum[32];
  __shared__ int s_bin;
  __shared__ unsigned int s_ge, s_cnt, s_intb, s_last;
  const int slot = blockIdx.x, l = P.topk_level[slot], img = blockIdx.y, chunk = blockIdx.z;
  const int n = P.n_anchor[l], k = P.keep[l];
  const float* keys = maxscore + (size_t)img * P.A_total + P.anchor_off[l];
  int32_t* out = cand_idx + (size_t)img * P.M + P.cand_off[l];
  unsigned int* h = hist + ((size_t)img * P.num_topk_levels + slot) * TOPK_HSTRIDE;
  unsigned int* ctr = h + TOPK_BINS;
  unsigned long long* gl = glist + ((size_t)img * P.num_topk_levels + slot) * TOPK_LIST;
  const int tid = threadIdx.x, lane = tid & 31;
  const uint4 h4 = *reinterpret_cast<const uint4*>(h + 4 * tid);
  const unsigned int hv[4] = {h4.x, h4.y, h4.z, h4.w};
  find_boundary_bin<4>(hv, (unsigned int)k, warp_sum, &s_bin, &s_ge);
  const int tb = s_bin;
  const bool multi = (tb >= 0) && (s_ge <= TOPK_LIST);          // the plain case: every chunk collects its slice
  if (!multi && chunk != 0) return;
  if (tb < 0) { topk_radix_select(keys, n, k, list, out); return; }
  int tb2 = 0;                                                // keys of bucket tb are taken from sub-bucket tb2 upwards
  if (!multi) {
    // the boundary bucket alone holds too many keys (clustered scores, e.g. the reference init: every score is
    // 0.07098 +- 2e-5): split it once more, TOPK_SUB sub-buckets of width 2^-23
    for (int i = tid; i < TOPK_SUB; i += 1024) sub[i] = 0;
    if (tid == 0) s_intb = 0;
    __syncthreads();
    for (int base = 0; base < n; base += 1024) {
      const int i = base + tid;
      const float sc = (i < n) ? __ldg(keys + i) : 0.f;
      const bool hit = (i < n) && score_bucket(sc) == tb;
      const int d = hit ? score_sub_bucket(sc, tb) : (TOPK_SUB + lane);
      const unsigned int peers = __match_any_sync(0xffffffffu, d);
      if (hit && (__ffs(peers) - 1) == lane) atomicAdd(&sub[d], (unsigned int)__popc(peers));
    }
    __syncthreads();
    const unsigned int sv[2] = {sub[2 * tid], sub[2 * tid + 1]};
    unsigned int tot = sv[0] + sv[1];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
    if (lane == 0) atomicAdd(&s_intb, tot);
    __syncthreads();
    const unsigned int gt_tb = s_ge - s_intb;                 // anchors in buckets above tb (< k)
    find_boundary_bin<2>(sv, (unsigned int)k - gt_tb, warp_sum, &s_bin, &s_ge);
    if (s_bin < 0 || gt_tb + s_ge > TOPK_LIST) {              // still one lump (exact ties): exact radix select
      topk_radix_select(keys, n, k, list, out);
      return;
    }
    tb2 = s_bin;
  }
  const int slice = multi ? (n + TOPK_CHUNKS - 1) / TOPK_CHUNKS : n;
  const int lo = multi ? min(n, chunk * slice) : 0, hi = multi ? min(n, lo + slice) : n;
  if (tid == 0) s_cnt = 0;
  __syncthreads();
  for (int base = lo; base < hi; base += 1024) {
    const int i = base + tid;
    const float sc = (i < hi) ? __ldg(keys + i) : 0.f;
    const int b = score_bucket(sc);
    const bool take = (i < hi) && (b > tb || (b == tb && (tb2 == 0 || score_sub_bucket(sc, tb) >= tb2)));
    const unsigned int bt = __ballot_sync(0xffffffffu, take);
    unsigned int slot0 = 0;
    if (lane == 0 && bt) slot0 = multi ? atomicAdd(&ctr[0], (unsigned int)__popc(bt)) : atomicAdd(&s_cnt, (unsigned int)__popc(bt));
    slot0 = __shfl_sync(0xffffffffu, slot0, 0);
    if (take) {
      const unsigned int pos = slot0 + __popc(bt & ((1u << lane) - 1u));
      const unsigned long long key = ((unsigned long long)float_to_ordered(sc) << 32) | (unsigned long long)(0xffffffffu - (unsigned int)i);
      if (pos < TOPK_LIST) { if (multi) gl[pos] = key; else list[pos] = key; }
    }
  }
  if (multi) {
    __threadfence();                                          // my list entries are visible before my ticket
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(&ctr[1], 1u) == (unsigned int)(TOPK_CHUNKS - 1)) ? 1u : 0u;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const int total_g = min((int)__ldcg(&ctr[0]), TOPK_LIST);
    for (int i = tid; i < total_g; i += 1024) list[i] = __ldcg(gl + i);
    if (tid == 0) s_cnt = (unsigned int)total_g;
  }
  __syncthreads();
  const int total = min((int)s_cnt, TOPK_LIST);
  const int Psort = next_pow2(max(total, 2));
  for (int i = total + tid; i < Psort; i += 1024) list[i] = 0ull;
  bitonic_sort_desc(list, Psort);
  for (int r = tid; r < k; r += 1024)
    out[r] = (int32_t)(0xffffffffu - (unsigned int)(list[r] & 0xffffffffull));
}

// ---------------------------------------------------------------------------------------- K3
// 32 candidates per CTA: every warp computes the C class scores of 4 candidates (coalesced
// row reads), lane 0 decodes the box; the score tile is transposed through shared memory so
// the class-major output rows are written 128 B at a time.
__device__ __forceinline__ float clampf_(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }

__global__ void __launch_bounds__(256) gather_decode_kernel(const __grid_constant__ PostParams P,
                                                            const float* __restrict__ img_info,
                                                            int32_t* __restrict__ cand_idx,
                                                            float* __restrict__ boxes,
                                                            float* __restrict__ scores_cm) {
  extern __shared__ float tile[];   // [C][33]
  const int img = blockIdx.y, j0 = blockIdx.x * 32;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* info = img_info + img * 8;
  for (int t = warp * 4; t < warp * 4 + 4; ++t) {
    const int j = j0 + t;
    if (j >= P.M) {
      for (int c = lane; c < P.C; c += 32) tile[c * 33 + t] = 0.f;
      continue;
    }
    int l = 0;
    while (j >= P.cand_off[l + 1]) ++l;
    int i;
    if (P.is_topk[l]) {
      i = cand_idx[(size_t)img * P.M + j];
    } else {
      i = j - P.cand_off[l];     // natural order when the level has <= nms_pre anchors (:536)
      if (lane == 0) cand_idx[(size_t)img * P.M + j] = i;
    }
    const int n_l = P.n_anchor[l];
    const size_t row = (size_t)img * n_l + i;
    const float ql = P.iou[l] ? __ldg(P.iou[l] + row) : 0.f;
    const float* crow = P.cls[l] + row * P.C;
    for (int c = lane; c < P.C; c += 32) tile[c * 33 + t] = fuse_score(__ldg(crow + c), ql, P.alpha);
    if (lane == 0) {
      // anchor (anchor_generator.py:57-67): base[a] + (x*s, y*s, x*s, y*s)
      const int a = i % P.A, pos = i / P.A;
      const int x = pos % P.feat_w[l], y = pos / P.feat_w[l];
      const float sx = (float)(x * P.stride[l]), sy = (float)(y * P.stride[l]);
      const float ax1 = __fadd_rn(P.base[l][a][0], sx), ay1 = __fadd_rn(P.base[l][a][1], sy);
      const float ax2 = __fadd_rn(P.base[l][a][2], sx), ay2 = __fadd_rn(P.base[l][a][3], sy);
      const float4 d = __ldg(reinterpret_cast<const float4*>(P.reg[l]) + row);
      float x1, y1, x2, y2;
      if (P.decode_mode == IOU_DECODE_DISTANCE) {
        // FCOS point (iou_aware_fcos_head.py:392-401) + distance2bbox (transforms.py:181-184)
        const float ptx = __fadd_rn(sx, (float)(P.stride[l] / 2)), pty = __fadd_rn(sy, (float)(P.stride[l] / 2));
        x1 = __fsub_rn(ptx, d.x); y1 = __fsub_rn(pty, d.y);
        x2 = __fadd_rn(ptx, d.z); y2 = __fadd_rn(pty, d.w);
      } else {
      // delta2bbox (transforms.py:44-78)
      const float dx = __fadd_rn(__fmul_rn(d.x, P.stdv[0]), P.mean[0]);
      const float dy = __fadd_rn(__fmul_rn(d.y, P.stdv[1]), P.mean[1]);
      float dw = __fadd_rn(__fmul_rn(d.z, P.stdv[2]), P.mean[2]);
      float dh = __fadd_rn(__fmul_rn(d.w, P.stdv[3]), P.mean[3]);
      dw = clampf_(dw, -P.max_ratio, P.max_ratio);
      dh = clampf_(dh, -P.max_ratio, P.max_ratio);
      const float px = __fmul_rn(__fadd_rn(ax1, ax2), 0.5f), py = __fmul_rn(__fadd_rn(ay1, ay2), 0.5f);
      const float pw = __fadd_rn(__fsub_rn(ax2, ax1), 1.0f), ph = __fadd_rn(__fsub_rn(ay2, ay1), 1.0f);
      const float gw = __fmul_rn(pw, expf(dw)), gh = __fmul_rn(ph, expf(dh));
      const float gx = __fadd_rn(px, __fmul_rn(pw, dx)), gy = __fadd_rn(py, __fmul_rn(ph, dy));
      const float hw = __fmul_rn(gw, 0.5f), hh = __fmul_rn(gh, 0.5f);
      x1 = __fadd_rn(__fsub_rn(gx, hw), 0.5f); y1 = __fadd_rn(__fsub_rn(gy, hh), 0.5f);
      x2 = __fsub_rn(__fadd_rn(gx, hw), 0.5f); y2 = __fsub_rn(__fadd_rn(gy, hh), 0.5f);
      }
      const float xmax = __fsub_rn(info[1], 1.0f), ymax = __fsub_rn(info[0], 1.0f);
      x1 = clampf_(x1, 0.f, xmax); y1 = clampf_(y1, 0.f, ymax);
      x2 = clampf_(x2, 0.f, xmax); y2 = clampf_(y2, 0.f, ymax);
      if (P.rescale) {   // mlvl_bboxes /= scale_factor, after the clamp (:553-554)
        x1 = __fdiv_rn(x1, info[2]); y1 = __fdiv_rn(y1, info[3]);
        x2 = __fdiv_rn(x2, info[4]); y2 = __fdiv_rn(y2, info[5]);
      }
      reinterpret_cast<float4*>(boxes)[(size_t)img * P.M + j] = make_float4(x1, y1, x2, y2);
    }
  }
  __syncthreads();
  const int jj = j0 + lane;
  if (jj < P.M)
    for (int c = warp; c < P.C; c += 8)
      scores_cm[((size_t)img * P.C + c) * P.M + jj] = tile[c * 33 + lane];
}

// ---------------------------------------------------------------------------------------- NMS core
// IoU > thr with the reference's operation order (nms_kernel.cu:13-21,60).
__device__ __forceinline__ bool iou_gt(const float4 a, const float sa, const float4 b, const float sb,
                                       const float thr) {
  const float left = fmaxf(a.x, b.x), right = fminf(a.z, b.z);
  const float top = fmaxf(a.y, b.y), bottom = fminf(a.w, b.w);
  const float w = fmaxf(__fadd_rn(__fsub_rn(right, left), 1.0f), 0.0f);
  const float h = fmaxf(__fadd_rn(__fsub_rn(bottom, top), 1.0f), 0.0f);
  const float inter = __fmul_rn(w, h);
  if (inter == 0.0f) return 0.0f > thr;
  const float uni = __fsub_rn(__fadd_rn(sa, sb), inter);
  return __fdiv_rn(inter, uni) > thr;
}
__device__ __forceinline__ float box_area(const float4 a) {
  return __fmul_rn(__fadd_rn(__fsub_rn(a.z, a.x), 1.0f), __fadd_rn(__fsub_rn(a.w, a.y), 1.0f));
}

#define NMS_ROUND 512
#define NMS_FIRST 256
struct NmsSmem {
  unsigned long long* sortbuf;   // [P]
  float4* sbox;                  // [n]
  float* sarea;                  // [n]
  int* kept;                     // [n] sorted positions of the boxes kept so far
};

// Greedy NMS over n boxes already ordered by (score desc, index asc) in shared memory.
// Scans sorted positions [begin, n).  Calls emit(r, pos) from thread 0 for every kept position r, in
// order; stops once `stop_after` boxes are kept in total (0 = never).  *s_total_p (shared memory) holds
// the kept count and persists across calls, so the scan can proceed in rounds.
// The scan is LAZY: a 64-box chunk is only tested when the scan reaches it --
//   (a) every (chunk box, previously kept box) pair is tested in parallel,
//   (b) the chunk's own 64x64 suppression mask is built with warp ballots,
//   (c) one thread resolves the chunk serially and appends to the kept list --
// so with the max_per_img+1 early stop only the first few chunks are ever touched.
template <typename Emit>
__device__ int greedy_nms_range(const NmsSmem& S, const int begin, const int n, const float thr,
                                const int stop_after, int* s_total_p, Emit emit) {
  __shared__ unsigned long long cmask[64];
  __shared__ unsigned int dead[2];
  int& s_total = *s_total_p;                   // kept count so far (shared memory, persists across ranges)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  __syncthreads();
  for (int cs = begin; cs < n; cs += 64) {
    const int cnt = min(64, n - cs);
    const int total0 = s_total;
    if (tid < 2) dead[tid] = 0u;
    __syncthreads();
    // (a) chunk boxes vs everything kept so far: thread -> (box b, kept slots k0, k0+stride, ...)
    {
      const int b = tid & 63, k0 = tid >> 6, kstride = blockDim.x >> 6;
      if (b < cnt && total0 > 0) {
        const float4 bb = S.sbox[cs + b];
        const float sb = S.sarea[cs + b];
        for (int k = k0; k < total0; k += kstride) {
          if ((((volatile unsigned int*)dead)[b >> 5] >> (b & 31)) & 1u) break;
          const int kp = S.kept[k];
          if (iou_gt(S.sbox[kp], S.sarea[kp], bb, sb, thr)) {
            atomicOr(&dead[b >> 5], 1u << (b & 31));
            break;
          }
        }
      }
    }
    // (b) intra-chunk mask rows (bits > row only)
    for (int r = warp; r < cnt; r += nwarps) {
      const float4 a = S.sbox[cs + r];
      const float sa = S.sarea[cs + r];
      bool h0 = false, h1 = false;
      const int c0 = lane, c1 = lane + 32;
      if (c0 > r && c0 < cnt) h0 = iou_gt(a, sa, S.sbox[cs + c0], S.sarea[cs + c0], thr);
      if (c1 > r && c1 < cnt) h1 = iou_gt(a, sa, S.sbox[cs + c1], S.sarea[cs + c1], thr);
      const unsigned int b0 = __ballot_sync(0xffffffffu, h0), b1 = __ballot_sync(0xffffffffu, h1);
      if (lane == 0) cmask[r] = ((unsigned long long)b1 << 32) | b0;
    }
    __syncthreads();
    // (c) serial resolve of the chunk
    if (tid == 0) {
      unsigned long long R = (unsigned long long)dead[0] | ((unsigned long long)dead[1] << 32);
      int total = total0;
      for (int i = 0; i < cnt; ++i) {
        if (!((R >> i) & 1ull)) {
          R |= cmask[i];
          S.kept[total] = cs + i;
          emit(cs + i, total);
          ++total;
          if (stop_after && total >= stop_after) break;
        }
      }
      s_total = total;
    }
    __syncthreads();
    if (stop_after && s_total >= stop_after) break;
  }
  return s_total;
}

// Bounded variant for the per-class kernel, where the scan always stops after `stop_after` kept boxes:
// only the kept boxes (<= stop_after) and the current 64-box chunk live in shared memory, the rest is
// fetched from global memory by candidate index, so several CTAs fit on one SM.
// keys: the range's composite keys sorted descending (shared memory); bx: the image's boxes (global).
template <typename Emit>
__device__ int greedy_nms_bounded(const unsigned long long* keys, const int n, const float4* __restrict__ bx,
                                  const float thr, const int stop_after, int* s_total_p, float4* kbox,
                                  float* karea, Emit emit) {
  __shared__ unsigned long long cmask[64];
  __shared__ float4 cbox[64];
  __shared__ float carea[64];
  __shared__ unsigned int dead[2];
  __shared__ unsigned long long s_keep;
  int& s_total = *s_total_p;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  __syncthreads();
  for (int cs = 0; cs < n; cs += 64) {
    const int cnt = min(64, n - cs);
    const int total0 = s_total;
    if (tid < cnt) {
      const unsigned int j = 0xffffffffu - (unsigned int)(keys[cs + tid] & 0xffffffffull);
      const float4 b = __ldg(bx + j);
      cbox[tid] = b;
      carea[tid] = box_area(b);
    }
    if (tid < 2) dead[tid] = 0u;
    __syncthreads();
    {
      const int b = tid & 63, k0 = tid >> 6, kstride = blockDim.x >> 6;
      if (b < cnt && total0 > 0) {
        const float4 bb = cbox[b];
        const float sb = carea[b];
        for (int k = k0; k < total0; k += kstride) {
          if ((((volatile unsigned int*)dead)[b >> 5] >> (b & 31)) & 1u) break;
          if (iou_gt(kbox[k], karea[k], bb, sb, thr)) {
            atomicOr(&dead[b >> 5], 1u << (b & 31));
            break;
          }
        }
      }
    }
    for (int r = warp; r < cnt; r += nwarps) {
      const float4 a = cbox[r];
      const float sa = carea[r];
      bool h0 = false, h1 = false;
      const int c0 = lane, c1 = lane + 32;
      if (c0 > r && c0 < cnt) h0 = iou_gt(a, sa, cbox[c0], carea[c0], thr);
      if (c1 > r && c1 < cnt) h1 = iou_gt(a, sa, cbox[c1], carea[c1], thr);
      const unsigned int b0 = __ballot_sync(0xffffffffu, h0), b1 = __ballot_sync(0xffffffffu, h1);
      if (lane == 0) cmask[r] = ((unsigned long long)b1 << 32) | b0;
    }
    __syncthreads();
    if (tid == 0) {
      // the sequential part is only the keep / suppress decision (a 64-bit mask walk, branch-free so that the mask loads
      // pipeline); copying the kept boxes and emitting their keys is done by all threads below
      unsigned long long R = (unsigned long long)dead[0] | ((unsigned long long)dead[1] << 32), keep = 0ull;
      int total = total0;
#pragma unroll 8
      for (int i = 0; i < cnt; ++i) {
        const unsigned long long m = cmask[i];
        const bool take = !((R >> i) & 1ull) && total < stop_after;
        R |= take ? m : 0ull;
        keep |= take ? (1ull << i) : 0ull;
        total += take ? 1 : 0;
      }
      s_total = total;
      s_keep = keep;
    }
    __syncthreads();
    if (tid < cnt) {
      const unsigned long long keep = s_keep;
      if ((keep >> tid) & 1ull) {
        const int pos = total0 + __popcll(keep & ((1ull << tid) - 1ull));
        kbox[pos] = cbox[tid];
        karea[pos] = carea[tid];
        emit(cs + tid, pos);
      }
    }
    __syncthreads();                                   // kbox complete, cbox free for the next chunk
    if (s_total >= stop_after) break;
  }
  return s_total;
}

// ---------------------------------------------------------------------------------------- K4
__global__ void __launch_bounds__(384, 5) class_nms_kernel(const __grid_constant__ PostParams P,
                                                        const float* __restrict__ boxes,
                                                        const float* __restrict__ scores_cm,
                                                        unsigned long long* __restrict__ kept_keys,
                                                        int32_t* __restrict__ kept_cnt, const int Pmax) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // candidates above the threshold, compacted: score bits (u32) and row (u16) = 6 bytes each (composite 64-bit
  // keys are only materialised for the <= NMS_ROUND members of a round, in rb)
  unsigned int* c_bits = reinterpret_cast<unsigned int*>(smem_raw);                              // [Pmax]
  unsigned short* c_row = reinterpret_cast<unsigned short*>(smem_raw + (size_t)Pmax * 4);          // [Pmax]
  float4* kbox = reinterpret_cast<float4*>(smem_raw + (size_t)Pmax * 6);       // [kcap] kept boxes
  float* karea = reinterpret_cast<float*>(smem_raw + (size_t)Pmax * 6 + (size_t)P.kcap * 16);
  auto key_at = [&](int i) {
    return ((unsigned long long)c_bits[i] << 32) | (unsigned long long)(0xffffffffu - (unsigned int)c_row[i]);
  };
  __shared__ unsigned int s_n, warp_cnt[16];
  const int c = blockIdx.x, img = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int T = blockDim.x, nw = T >> 5;            // 384 threads: five CTAs per SM
  const float* sc = scores_cm + ((size_t)img * P.C + c) * P.M;
  // compaction of rows with score > score_thr (bbox_nms.py:37), ascending row index: warp w owns the contiguous rows
  // [w * per, (w + 1) * per) -- count, one barrier, prefix over the warps, write (the second read hits L1)
  {
    const int per = ((P.M + nw - 1) / nw + 31) & ~31;
    const int lo = warp * per, hi = min(P.M, lo + per);
    unsigned int mine = 0;
    for (int j = lo + lane; j < lo + per; j += 32) {
      const bool pass = (j < hi) && (__ldg(sc + j) > P.score_thr);
      mine += __popc(__ballot_sync(0xffffffffu, pass));
    }
    if (lane == 0) warp_cnt[warp] = mine;
    __syncthreads();
    unsigned int off = 0, tot = 0;
    for (int w = 0; w < nw; ++w) { const unsigned int c_ = warp_cnt[w]; tot += c_; if (w < warp) off += c_; }
    for (int j = lo + lane; j < lo + per; j += 32) {
      const float sv = (j < hi) ? __ldg(sc + j) : 0.f;
      const bool pass = (j < hi) && (sv > P.score_thr);
      const unsigned int b = __ballot_sync(0xffffffffu, pass);
      if (pass) {
        const unsigned int slot = off + __popc(b & ((1u << lane) - 1u));
        c_bits[slot] = float_to_ordered(sv);
        c_row[slot] = (unsigned short)j;
      }
      off += __popc(b);
    }
    if (tid == 0) s_n = tot;
    __syncthreads();
  }
  const int n = (int)s_n;
  int32_t* cnt_out = kept_cnt + (size_t)img * P.C + c;
  if (n == 0) {
    if (tid == 0) *cnt_out = 0;
    return;
  }
  const float4* bx = reinterpret_cast<const float4*>(boxes) + (size_t)img * P.M;
  unsigned long long* keys_out = kept_keys + ((size_t)img * P.C + c) * P.kcap;
  __shared__ int s_total;
  if (tid == 0) s_total = 0;
  // Bucket histogram of the class's scores: bucket(s) is monotone in s, so "all candidates of buckets
  // [blo, bhi)" is a contiguous slice of the score order.  The scan stops after max_per_img+1 kept boxes, so
  // the order is consumed from the top in rounds: a first round of <= NMS_FIRST candidates (one cheap 256-wide
  // sort usually already yields the 101 kept boxes), then rounds of <= NMS_ROUND.  One histogram pass + a scan
  // of 1024 bins replaces the 8-pass exact radix select, which remains the fallback when a single bucket holds
  // more than a round (many near-equal scores).
  __shared__ unsigned int bh[1024];
  __shared__ unsigned int s_maxb;
  __shared__ int s_blo, s_teff;
  __shared__ unsigned long long rb[NMS_ROUND];
  const float b_scale = 1024.0f / fmaxf(1.0f - P.score_thr, 1e-6f);
  auto bucket_of = [&](unsigned long long key) {
    const float sv = ordered_to_float((unsigned int)(key >> 32));
    const int bq = (int)((sv - P.score_thr) * b_scale);
    return min(1023, max(0, bq));
  };
  bool use_buckets = false;
  if (n > NMS_FIRST) {
    for (int i = tid; i < 1024; i += T) bh[i] = 0;
    if (tid == 0) s_maxb = 0;
    __syncthreads();
    for (int i = tid; i < n; i += T) atomicAdd(&bh[bucket_of(key_at(i))], 1u);
    __syncthreads();
    for (int i = tid; i < 1024; i += T) if (bh[i] > NMS_FIRST) atomicMax(&s_maxb, bh[i]);
    __syncthreads();
    use_buckets = (s_maxb == 0);               // every bucket fits the first (smallest) round
  }
  if (n <= NMS_FIRST) {
    // small class: sort everything once
    const int Ps = next_pow2(n);
    for (int i = tid; i < Ps; i += T) rb[i] = (i < n) ? key_at(i) : 0ull;
    bitonic_sort_desc(rb, Ps);
    const unsigned long long* sb = rb;
    greedy_nms_bounded(sb, n, bx, P.iou_thr, P.kcap, &s_total, kbox, karea,
                       [&](int r, int pos) { keys_out[pos] = sb[r]; });
  } else if (use_buckets) {
    __shared__ unsigned int s_slot2;
    int bhi = 1024, cap = NMS_FIRST;
    while (bhi > 0) {
      __syncthreads();
      if (warp == 0) {
        // lowest blo with sum(bh[blo .. bhi)) <= cap, walking down from bhi and stopping at the first bin that would
        // overflow: lane L owns bins [32L, 32L + 32); suffix sums over the lanes find the lane where the walk stops,
        // that lane finishes inside its own bins
        unsigned int mine = 0;
        for (int k = 0; k < 32; ++k) { const int b_ = 32 * lane + k; if (b_ < bhi) mine += bh[b_]; }
        unsigned int suf = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const unsigned int v_ = __shfl_down_sync(0xffffffffu, suf, o);
          if (lane + o < 32) suf += v_;
        }
        const unsigned int over = __ballot_sync(0xffffffffu, suf > (unsigned int)cap);
        const unsigned int above_mine = __shfl_down_sync(0xffffffffu, suf, 1);
        int bq = 0, cum = (int)__shfl_sync(0xffffffffu, suf, 0);
        if (over != 0u) {
          const int lc = 31 - __clz(over);
          if (lane == lc) {
            bq = min(bhi, 32 * lc + 32);
            cum = (lc < 31) ? (int)above_mine : 0;
            while (bq > 32 * lc && cum + (int)bh[bq - 1] <= cap) cum += (int)bh[--bq];
          }
          bq = __shfl_sync(0xffffffffu, bq, lc);
          cum = __shfl_sync(0xffffffffu, cum, lc);
        }
        if (lane == 0) { s_blo = bq; s_teff = cum; s_slot2 = 0; }
      }
      __syncthreads();
      const int blo = s_blo, teff = s_teff;
      if (teff > 0) {
        for (int base = 0; base < n; base += T) {
          const int i = base + tid;
          const unsigned long long u = (i < n) ? key_at(i) : 0ull;
          const int bq = (i < n) ? bucket_of(u) : -1;
          const bool take = bq >= blo && bq < bhi;
          const unsigned int bt = __ballot_sync(0xffffffffu, take);
          unsigned int slot0 = 0;
          if (lane == 0 && bt) slot0 = atomicAdd(&s_slot2, __popc(bt));
          slot0 = __shfl_sync(0xffffffffu, slot0, 0);
          if (take) rb[slot0 + __popc(bt & ((1u << lane) - 1u))] = u;
        }
        __syncthreads();
        const int Ps = next_pow2(teff);
        for (int i = teff + tid; i < Ps; i += T) rb[i] = 0ull;
        bitonic_sort_desc(rb, Ps);
        const int total = greedy_nms_bounded(rb, teff, bx, P.iou_thr, P.kcap, &s_total, kbox, karea,
                                             [&](int r, int pos) { keys_out[pos] = rb[r]; });
        if (total >= P.kcap) break;
      }
      bhi = blo;
      cap = NMS_ROUND;
    }
  } else {
    // big class: the scan stops after max_per_img+1 kept boxes, so only the head of the score order is
    // ever needed.  Rounds of NMS_ROUND boxes: exact radix select of the round's lowest composite key
    // (keys are unique: score bits | ~index), compaction, a 1024-wide sort, then the lazy greedy scan.
    unsigned int* hist = bh;                    // [256]: the bucket histogram is not used on this path
    __shared__ unsigned long long s_prefix;
    __shared__ unsigned int s_kleft, s_slot;
    unsigned long long bound = ~0ull;           // keys of this round are < bound
    int processed = 0;
    while (processed < n) {
      const int teff = min(NMS_ROUND, n - processed);
      if (tid == 0) { s_prefix = 0ull; s_kleft = teff; s_slot = 0; }
      unsigned long long mask = 0ull;
      for (int pass = 0; pass < 8; ++pass) {
        const int shift = 56 - 8 * pass;
        if (tid < 256) hist[tid] = 0;
        __syncthreads();
        const unsigned long long prefix = s_prefix;
        for (int base = 0; base < n; base += T) {
          const int i = base + tid;
          const unsigned long long u = (i < n) ? key_at(i) : 0ull;
          const bool valid = (i < n) && (u < bound) && ((u & mask) == prefix);
          const unsigned int d = valid ? (unsigned int)((u >> shift) & 255ull) : (256u + lane);
          const unsigned int peers = __match_any_sync(0xffffffffu, d);
          if (valid && (__ffs(peers) - 1) == lane) atomicAdd(&hist[d], __popc(peers));
        }
        __syncthreads();
        if (tid == 0) {
          unsigned int cacc = 0, kl = s_kleft;
          int d = 255;
          for (; d > 0; --d) {
            if (cacc + hist[d] >= kl) break;
            cacc += hist[d];
          }
          s_kleft = kl - cacc;
          s_prefix = prefix | ((unsigned long long)d << shift);
        }
        mask |= 255ull << shift;
        __syncthreads();
      }
      const unsigned long long kt = s_prefix;   // the teff-th largest key below `bound`
      for (int base = 0; base < n; base += T) {
        const int i = base + tid;
        const unsigned long long u = (i < n) ? key_at(i) : 0ull;
        const bool take = (i < n) && (u < bound) && (u >= kt);
        const unsigned int bt = __ballot_sync(0xffffffffu, take);
        unsigned int slot0 = 0;
        if (lane == 0 && bt) slot0 = atomicAdd(&s_slot, __popc(bt));
        slot0 = __shfl_sync(0xffffffffu, slot0, 0);
        if (take) rb[slot0 + __popc(bt & ((1u << lane) - 1u))] = u;
      }
      __syncthreads();
      const int Ps = next_pow2(teff);
      for (int i = teff + tid; i < Ps; i += T) rb[i] = 0ull;
      bitonic_sort_desc(rb, Ps);
      const int total = greedy_nms_bounded(rb, teff, bx, P.iou_thr, P.kcap, &s_total, kbox, karea,
                                           [&](int r, int pos) { keys_out[pos] = rb[r]; });
      processed += teff;
      bound = kt;
      if (total >= P.kcap) break;
      __syncthreads();
    }
  }
  __syncthreads();
  if (tid == 0) *cnt_out = s_total;
}

// ---------------------------------------------------------------------------------------- K5
// bbox_nms.py:55-62: concat classes; if more than max_per_img remain, order by score desc.
__global__ void __launch_bounds__(1024) final_select_kernel(const __grid_constant__ PostParams P,
                                                            const float* __restrict__ boxes,
                                                            const unsigned long long* __restrict__ kept_keys,
                                                            const int32_t* __restrict__ kept_cnt,
                                                            float* __restrict__ dets,
                                                            long long* __restrict__ labels,
                                                            int32_t* __restrict__ counts,
                                                            const int rank_order) {
  // rank_order (soft NMS): rows of a class keep their SELECTION order (the order soft_nms returns them in),
  // so the unsorted key is (class, rank) and the candidate index is looked up again afterwards
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned long long* sortbuf = reinterpret_cast<unsigned long long*>(smem_raw);   // [Pcap]
  __shared__ int s_off[257];
  const int img = blockIdx.x, tid = threadIdx.x;
  if (tid < 256) s_off[tid + 1] = (tid < P.C) ? kept_cnt[(size_t)img * P.C + tid] : 0;   // parallel loads, serial scan
  __syncthreads();
  if (tid == 0) {
    int t = 0;
    for (int c = 0; c < P.C; ++c) { const int n_c = s_off[c + 1]; s_off[c] = t; t += n_c; }
    s_off[P.C] = t;
  }
  __syncthreads();
  const int total = s_off[P.C];
  const bool by_score = total > P.max_per_img;
  const int Ps = next_pow2(max(total, 1));
  for (int i = total + tid; i < Ps; i += 1024) sortbuf[i] = 0ull;
  // one kept row per thread and trip (all loads of a trip in flight together; a class-by-class loop would wait for
  // global memory C times in a row): the row's class comes from a binary search in the class offsets
  for (int e = tid; e < total; e += 1024) {
    int lo_c = 0, hi_c = P.C;                                  // s_off[lo_c] <= e < s_off[hi_c]
    while (hi_c - lo_c > 1) {
      const int mid = (lo_c + hi_c) >> 1;
      if (s_off[mid] <= e) lo_c = mid; else hi_c = mid;
    }
    const int c = lo_c, r = e - s_off[c];
    const unsigned long long key = kept_keys[((size_t)img * P.C + c) * P.kcap + r];
    const unsigned int sbits = (unsigned int)(key >> 32);
    const unsigned int j = 0xffffffffu - (unsigned int)(key & 0xffffffffull);
    const unsigned int pos = rank_order ? (unsigned int)(c * P.kcap + r)
                                        : (unsigned int)c * (unsigned int)P.M + j;     // class-major, row-minor
    // payload (pos) must survive the sort: by_score -> key = (score, ~pos); else key = (~pos, score)
    sortbuf[e] = by_score ? (((unsigned long long)sbits << 32) | (0xffffffffu - pos))
                          : (((unsigned long long)(0xffffffffu - pos) << 32) | sbits);
  }
  const int k = min(total, P.max_per_img);
  if (by_score && total > 4 * P.max_per_img) {
    // Only the max_per_img largest keys are needed: exact radix select of the k-th largest key (keys are unique:
    // score bits | ~position), compaction of the keys >= it to the front, and a sort of just those --
    // instead of a bitonic sort of up to 8192 keys (91 barrier-separated stages on one SM per image).
    __shared__ unsigned int hist[256];
    __shared__ unsigned long long s_prefix;
    __shared__ unsigned int s_kleft, s_slot;
    __syncthreads();
    if (tid == 0) { s_prefix = 0ull; s_kleft = (unsigned int)k; s_slot = 0; }
    unsigned long long mask = 0ull;
    for (int pass = 0; pass < 8; ++pass) {
      const int shift = 56 - 8 * pass;
      if (tid < 256) hist[tid] = 0;
      __syncthreads();
      const unsigned long long prefix = s_prefix;
      for (int i = tid; i < total; i += 1024) {
        const unsigned long long u = sortbuf[i];
        if ((u & mask) == prefix) atomicAdd(&hist[(unsigned int)((u >> shift) & 255ull)], 1u);
      }
      __syncthreads();
      if (tid == 0) {
        unsigned int cacc = 0, kl = s_kleft;
        int d = 255;
        for (; d > 0; --d) {
          if (cacc + hist[d] >= kl) break;
          cacc += hist[d];
        }
        s_kleft = kl - cacc;
        s_prefix = prefix | ((unsigned long long)d << shift);
      }
      mask |= 255ull << shift;
      __syncthreads();
    }
    const unsigned long long kth = s_prefix;                     // exactly k keys are >= kth
    unsigned long long mine[8];                                  // total <= 8192 = 8 x 1024 (checked by the host)
    int nm = 0;
    for (int i = tid; i < total; i += 1024) {
      const unsigned long long u = sortbuf[i];
      if (u >= kth) mine[nm++] = u;
    }
    __syncthreads();
    const int Pk = next_pow2(k);
    for (int q = 0; q < nm; ++q) sortbuf[atomicAdd(&s_slot, 1u)] = mine[q];
    __syncthreads();
    for (int i = k + tid; i < Pk; i += 1024) sortbuf[i] = 0ull;
    bitonic_sort_desc(sortbuf, Pk);
  } else {
    bitonic_sort_desc(sortbuf, Ps);
  }
  float* d_out = dets + (size_t)img * P.max_per_img * 5;
  long long* l_out = labels + (size_t)img * P.max_per_img;
  const float4* bx = reinterpret_cast<const float4*>(boxes) + (size_t)img * P.M;
  for (int r = tid; r < P.max_per_img; r += 1024) {
    if (r < k) {
      const unsigned long long key = sortbuf[r];
      const unsigned int hi = (unsigned int)(key >> 32), lo = (unsigned int)(key & 0xffffffffull);
      const unsigned int sbits = by_score ? hi : lo;
      const unsigned int pos = 0xffffffffu - (by_score ? lo : hi);
      unsigned int c, j;
      if (rank_order) {
        c = pos / (unsigned int)P.kcap;
        const unsigned long long k2 = kept_keys[((size_t)img * P.C + c) * P.kcap + (pos - c * (unsigned int)P.kcap)];
        j = 0xffffffffu - (unsigned int)(k2 & 0xffffffffull);
      } else {
        c = pos / (unsigned int)P.M; j = pos - c * (unsigned int)P.M;
      }
      const float4 b = __ldg(bx + j);
      d_out[r * 5 + 0] = b.x; d_out[r * 5 + 1] = b.y; d_out[r * 5 + 2] = b.z; d_out[r * 5 + 3] = b.w;
      d_out[r * 5 + 4] = ordered_to_float(sbits);
      l_out[r] = (long long)c;
    } else {
      for (int q = 0; q < 5; ++q) d_out[r * 5 + q] = 0.f;
      l_out[r] = 0;
    }
  }
  if (tid == 0) counts[img] = k;
}

// ---------------------------------------------------------------------------------------- single NMS
// Drop-in for nms_cuda.nms: all n boxes take part; output = ascending original indices.
__global__ void __launch_bounds__(512) single_nms_kernel(const float* __restrict__ dets, const int n,
                                                         const float thr, long long* __restrict__ keep_idx,
                                                         int32_t* __restrict__ keep_count, const int Pmax) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  NmsSmem S;
  S.sortbuf = reinterpret_cast<unsigned long long*>(smem_raw);
  S.sbox = reinterpret_cast<float4*>(smem_raw + (size_t)Pmax * 8);
  S.sarea = reinterpret_cast<float*>(smem_raw + (size_t)Pmax * 8 + (size_t)n * 16);
  S.kept = reinterpret_cast<int*>(smem_raw + (size_t)Pmax * 8 + (size_t)n * 20);
  unsigned int* keepbits = reinterpret_cast<unsigned int*>(S.kept + n);
  __shared__ unsigned int s_run, warp_cnt[16];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int Ps = next_pow2(n);
  for (int i = tid; i < Ps; i += 512)
    S.sortbuf[i] = (i < n) ? (((unsigned long long)float_to_ordered(dets[(size_t)i * 5 + 4]) << 32) |
                              (unsigned long long)(0xffffffffu - (unsigned int)i))
                           : 0ull;
  for (int i = tid; i < (n + 31) / 32; i += 512) keepbits[i] = 0u;
  bitonic_sort_desc(S.sortbuf, Ps);
  for (int r = tid; r < n; r += 512) {
    const unsigned int j = 0xffffffffu - (unsigned int)(S.sortbuf[r] & 0xffffffffull);
    const float4 b = make_float4(dets[(size_t)j * 5], dets[(size_t)j * 5 + 1], dets[(size_t)j * 5 + 2],
                                 dets[(size_t)j * 5 + 3]);
    S.sbox[r] = b;
    S.sarea[r] = box_area(b);
  }
  __syncthreads();
  const unsigned long long* sb = S.sortbuf;
  __shared__ int s_total;
  if (tid == 0) s_total = 0;
  greedy_nms_range(S, 0, n, thr, 0, &s_total, [&](int r, int) {
    const unsigned int j = 0xffffffffu - (unsigned int)(sb[r] & 0xffffffffull);
    keepbits[j >> 5] |= 1u << (j & 31);     // only thread 0 emits
  });
  __syncthreads();
  if (tid == 0) s_run = 0;
  __syncthreads();
  for (int base = 0; base < n; base += 512) {
    const int j = base + tid;
    const bool kp = (j < n) && ((keepbits[j >> 5] >> (j & 31)) & 1u);
    const unsigned int b = __ballot_sync(0xffffffffu, kp);
    if (lane == 0) warp_cnt[warp] = __popc(b);
    __syncthreads();
    unsigned int off = s_run;
    for (int w = 0; w < warp; ++w) off += warp_cnt[w];
    if (kp) keep_idx[off + __popc(b & ((1u << lane) - 1u))] = (long long)j;
    __syncthreads();
    if (tid == 0) {
      unsigned int t = 0;
      for (int w = 0; w < 16; ++w) t += warp_cnt[w];
      s_run += t;
    }
    __syncthreads();
  }
  if (tid == 0) *keep_count = (int32_t)s_run;
}


// ---------------------------------------------------------------------------------------- single NMS, any n
// n > IOU_MAX_NMS_BOXES does not fit one block's shared memory: the reference's own structure instead
// (nms_kernel.cu:23-67: 64 x 64 suppression-mask tiles over the score-sorted boxes), with the sort and the greedy
// scan (which the reference runs on the HOST after a D2H copy of the mask, :99-123) kept on the device.
__global__ void nms_large_keys_kernel(const float* __restrict__ dets, const int n, const int P,
                                      unsigned long long* __restrict__ keys) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < P)
    keys[i] = (i < n) ? (((unsigned long long)float_to_ordered(dets[(size_t)i * 5 + 4]) << 32) |
                         (unsigned long long)(0xffffffffu - (unsigned int)i))
                      : 0ull;                       // padding sorts behind every real key
}
// one compare-exchange pass of a bitonic network over global memory (descending)
__global__ void nms_large_bitonic_kernel(unsigned long long* __restrict__ a, const int P, const int j, const int k) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  const int l = i ^ j;
  if (l > i) {
    const unsigned long long x = a[i], y = a[l];
    const bool desc = (i & k) == 0;
    if (desc ? (x < y) : (x > y)) { a[i] = y; a[l] = x; }
  }
}
__global__ void nms_large_gather_kernel(const float* __restrict__ dets, const int n,
                                        const unsigned long long* __restrict__ keys, float4* __restrict__ sbox,
                                        float* __restrict__ sarea) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const unsigned int j = 0xffffffffu - (unsigned int)(keys[r] & 0xffffffffull);
  const float4 b = make_float4(dets[(size_t)j * 5], dets[(size_t)j * 5 + 1], dets[(size_t)j * 5 + 2],
                               dets[(size_t)j * 5 + 3]);
  sbox[r] = b;
  sarea[r] = box_area(b);
}
// mask[r][cb] bit c: sorted box cb*64+c is suppressed by sorted box r (only pairs with a higher position, :51-57)
__global__ void __launch_bounds__(64) nms_large_mask_kernel(const float4* __restrict__ sbox, const float* __restrict__ sarea,
                                                            const int n, const int cb, const float thr,
                                                            unsigned long long* __restrict__ mask) {
  const int rb = blockIdx.y, cbk = blockIdx.x;
  if (cbk < rb) return;
  __shared__ float4 cbox[64];
  __shared__ float carea[64];
  const int cn = min(n - cbk * 64, 64), rn = min(n - rb * 64, 64);
  if ((int)threadIdx.x < cn) { cbox[threadIdx.x] = sbox[cbk * 64 + threadIdx.x]; carea[threadIdx.x] = sarea[cbk * 64 + threadIdx.x]; }
  __syncthreads();
  if ((int)threadIdx.x < rn) {
    const int r = rb * 64 + threadIdx.x;
    const float4 a = sbox[r];
    const float sa = sarea[r];
    unsigned long long t = 0ull;
    for (int c = (rb == cbk) ? (int)threadIdx.x + 1 : 0; c < cn; ++c)
      if (iou_gt(a, sa, cbox[c], carea[c], thr)) t |= 1ull << c;
    mask[(size_t)r * cb + cbk] = t;
  }
}
// greedy scan over the mask (the reference's host loop, :105-123) + ascending original indices (:127-130)
__global__ void __launch_bounds__(1024) nms_large_scan_kernel(const unsigned long long* __restrict__ mask, const int n,
                                                              const int cb, const unsigned long long* __restrict__ keys,
                                                              unsigned int* __restrict__ keepbits,
                                                              long long* __restrict__ keep_idx,
                                                              int32_t* __restrict__ keep_count) {
  extern __shared__ unsigned long long remv[];
  __shared__ unsigned long long s_kept;
  __shared__ unsigned int s_run, warp_cnt[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int w = tid; w < cb; w += 1024) remv[w] = 0ull;
  for (int w = tid; w < (n + 31) / 32; w += 1024) keepbits[w] = 0u;
  __syncthreads();
  for (int blk = 0; blk < cb; ++blk) {
    if (tid == 0) {
      unsigned long long r = remv[blk], kept = 0ull;
      const int rn = min(n - blk * 64, 64);
      for (int b = 0; b < rn; ++b) {
        if (!((r >> b) & 1ull)) {
          kept |= 1ull << b;
          r |= mask[(size_t)(blk * 64 + b) * cb + blk];
          const unsigned int j = 0xffffffffu - (unsigned int)(keys[blk * 64 + b] & 0xffffffffull);
          keepbits[j >> 5] |= 1u << (j & 31);
        }
      }
      s_kept = kept;
    }
    __syncthreads();
    const unsigned long long kept = s_kept;
    for (int w = blk + 1 + tid; w < cb; w += 1024) {
      unsigned long long acc = remv[w], kk = kept;
      while (kk) {
        const int b = __ffsll((long long)kk) - 1;
        kk &= kk - 1;
        acc |= mask[(size_t)(blk * 64 + b) * cb + w];
      }
      remv[w] = acc;
    }
    __syncthreads();
  }
  if (tid == 0) s_run = 0;
  __syncthreads();
  for (int base = 0; base < n; base += 1024) {
    const int j = base + tid;
    const bool kp = (j < n) && ((keepbits[j >> 5] >> (j & 31)) & 1u);
    const unsigned int b = __ballot_sync(0xffffffffu, kp);
    if (lane == 0) warp_cnt[warp] = __popc(b);
    __syncthreads();
    unsigned int off = s_run;
    for (int w = 0; w < warp; ++w) off += warp_cnt[w];
    if (kp) keep_idx[off + __popc(b & ((1u << lane) - 1u))] = (long long)j;
    __syncthreads();
    if (tid == 0) {
      unsigned int t = 0;
      for (int w = 0; w < 32; ++w) t += warp_cnt[w];
      s_run += t;
    }
    __syncthreads();
  }
  if (tid == 0) *keep_count = (int32_t)s_run;
}

struct NmsLargeWs { size_t keys, sbox, sarea, mask, keepbits, total; int P, cb; };
static NmsLargeWs nms_large_carve(int n) {
  NmsLargeWs W;
  W.P = next_pow2_host(n) < 2 ? 2 : next_pow2_host(n);
  W.cb = (n + 63) / 64;
  size_t off = 0;
  W.keys = off; off = align_up(off + (size_t)W.P * 8, 256);
  W.sbox = off; off = align_up(off + (size_t)n * 16, 256);
  W.sarea = off; off = align_up(off + (size_t)n * 4, 256);
  W.mask = off; off = align_up(off + (size_t)n * W.cb * 8, 256);
  W.keepbits = off; off = align_up(off + (size_t)((n + 31) / 32) * 4, 256);
  W.total = off;
  return W;
}


// ---------------------------------------------------------------------------------------- soft NMS
// Drop-in for soft_nms_cpu (mmdet/ops/nms/src/soft_nms_cpu.pyx:22-127), SURVEY 8(f) rank 3.  The reference is
// a sequential in-place loop; what is kept here is its exact OUTPUT, including the order that its swap /
// swap-with-last bookkeeping produces (that order decides ties between equal scores):
//   select  : first maximum of score over positions [i, N)            (strict '<', :52)  -> block arg-max
//   swap    : rows i <-> maxpos                                        (:57-72)
//   rescore : every later box that overlaps the selected one           (:82-113)          -> one box per thread
//   remove  : boxes whose new score < min_score; the sequential "overwrite with the last live box and look
//             again" loop (:116-124) is a two-pointer partition: survivors in front stay where they are and the
//             holes, in ascending order, receive the surviving boxes of the tail in DESCENDING position order
// Precision follows the C that Cython generates: '+ 1' is emitted as the double 1.0, so areas and the union are
// double expressions rounded to float once, iw*ih and the division are float, 1 - ov is a double subtraction;
// np.exp runs in double on the float argument.  Explicit _rn intrinsics keep nvcc from contracting into FMAs.
struct SoftSmem {
  float4* box;           // [cap]
  float* score;          // [cap]
  int* idx;              // [cap] original row
  int* list;             // [2*cap] holes | fillers
  unsigned char* drop;   // [cap]
};
__host__ __device__ inline size_t soft_smem_bytes(int cap) { return (size_t)cap * (16 + 4 + 4 + 8 + 1) + 16; }
__device__ __forceinline__ SoftSmem soft_carve(unsigned char* base, int cap) {
  SoftSmem S;
  S.box = reinterpret_cast<float4*>(base);
  S.score = reinterpret_cast<float*>(base + (size_t)cap * 16);
  S.idx = reinterpret_cast<int*>(base + (size_t)cap * 20);
  S.list = reinterpret_cast<int*>(base + (size_t)cap * 24);
  S.drop = base + (size_t)cap * 32;
  return S;
}

__device__ __forceinline__ float soft_rescore(const float4 t, const double t_area, const float4 b, const float s,
                                              const int method, const float thr, const float sigma, bool& touched) {
  touched = false;
  const float iw = (float)__dadd_rn((double)__fsub_rn(fminf(t.z, b.z), fmaxf(t.x, b.x)), 1.0);        // :91
  if (!(iw > 0.f)) return s;
  const float ih = (float)__dadd_rn((double)__fsub_rn(fminf(t.w, b.w), fmaxf(t.y, b.y)), 1.0);        // :93
  if (!(ih > 0.f)) return s;
  touched = true;
  const float area = (float)__dmul_rn(__dadd_rn((double)__fsub_rn(b.z, b.x), 1.0),
                                      __dadd_rn((double)__fsub_rn(b.w, b.y), 1.0));                    // :90
  const float inter = __fmul_rn(iw, ih);
  const float ua = (float)__dsub_rn(__dadd_rn(t_area, (double)area), (double)inter);                  // :95
  const float ov = __fdiv_rn(inter, ua);                                                             // :96
  float w;
  if (method == 1) w = (ov > thr) ? (float)__dsub_rn(1.0, (double)ov) : 1.f;                         // :98-102
  else if (method == 2) w = (float)exp((double)__fdiv_rn(-__fmul_rn(ov, ov), sigma));                // :103-104
  else w = (ov > thr) ? 0.f : 1.f;                                                                   // :105-109
  return __fmul_rn(w, s);                                                                            // :111
}

// Exclusive rank of `flag` among the block's threads (thread order) + block total; two barriers.
__device__ __forceinline__ int block_rank(const bool flag, int* warp_cnt, int& total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const unsigned int b = __ballot_sync(0xffffffffu, flag);
  if (lane == 0) warp_cnt[warp] = __popc(b);
  __syncthreads();
  int off = 0, tot = 0;
  for (int w = 0; w < nw; ++w) { const int c = warp_cnt[w]; tot += c; if (w < warp) off += c; }
  __syncthreads();
  total = tot;
  return off + __popc(b & ((1u << lane) - 1u));
}

// Runs the loop on n boxes staged in S; returns the number of live boxes N (positions [0, N) hold the result in
// selection order).  stop_after > 0 ends the loop after that many selections (later rows are then unfinished).
__device__ int soft_nms_core(const SoftSmem S, const int n, const int stop_after, const int method,
                             const float thr, const float sigma, const float min_score) {
  __shared__ float w_s[32];
  __shared__ int w_p[32], warp_cnt[32];
  __shared__ int s_drops;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, T = blockDim.x, nw = T >> 5;
  int N = n;
  __syncthreads();
  for (int i = 0; i < N; ++i) {
    if (stop_after > 0 && i >= stop_after) break;
    // ---- select: first maximum over [i, N)
    float bs = -INFINITY;
    int bp = 0x7fffffff;
    for (int p = i + tid; p < N; p += T) {
      const float s = S.score[p];
      if (s > bs || bp == 0x7fffffff) { bs = s; bp = p; }       // ascending p per thread: strict > keeps the first
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float os = __shfl_xor_sync(0xffffffffu, bs, o);
      const int op = __shfl_xor_sync(0xffffffffu, bp, o);
      if (op != 0x7fffffff && (bp == 0x7fffffff || os > bs || (os == bs && op < bp))) { bs = os; bp = op; }
    }
    if (lane == 0) { w_s[warp] = bs; w_p[warp] = bp; }
    __syncthreads();
    if (warp == 0) {
      bs = lane < nw ? w_s[lane] : -INFINITY;
      bp = lane < nw ? w_p[lane] : 0x7fffffff;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float os = __shfl_xor_sync(0xffffffffu, bs, o);
        const int op = __shfl_xor_sync(0xffffffffu, bp, o);
        if (op != 0x7fffffff && (bp == 0x7fffffff || os > bs || (os == bs && op < bp))) { bs = os; bp = op; }
      }
      if (lane == 0) {                                           // swap rows i <-> maxpos (:57-72)
        const float4 tb = S.box[i]; const float ts = S.score[i]; const int ti = S.idx[i];
        S.box[i] = S.box[bp]; S.score[i] = S.score[bp]; S.idx[i] = S.idx[bp];
        S.box[bp] = tb; S.score[bp] = ts; S.idx[bp] = ti;
        s_drops = 0;
      }
    }
    __syncthreads();
    // ---- rescore positions (i, N)
    const float4 t = S.box[i];
    const double t_area = __dmul_rn(__dadd_rn((double)__fsub_rn(t.z, t.x), 1.0), __dadd_rn((double)__fsub_rn(t.w, t.y), 1.0));
    int my_drops = 0;
    for (int p = i + 1 + tid; p < N; p += T) {
      bool touched;
      const float ns = soft_rescore(t, t_area, S.box[p], S.score[p], method, thr, sigma, touched);
      S.score[p] = ns;
      const bool d = touched && (ns < min_score);                // :114 (only rescored boxes are examined)
      S.drop[p] = d ? 1 : 0;
      my_drops += d ? 1 : 0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) my_drops += __shfl_xor_sync(0xffffffffu, my_drops, o);
    if (lane == 0 && my_drops) atomicAdd(&s_drops, my_drops);
    __syncthreads();
    const int D = s_drops;
    if (D > 0) {
      // ---- remove: two-pointer partition of (i, N)
      const int Nn = N - D;
      int H = 0, run = 0;
      for (int base = i + 1; base < Nn; base += T) {             // holes, ascending
        const int p = base + tid;
        const bool f = p < Nn && S.drop[p];
        int tot;
        const int r = block_rank(f, warp_cnt, tot);
        if (f) S.list[run + r] = p;
        run += tot;
      }
      H = run; run = 0;
      for (int base = 0; base < D; base += T) {                  // surviving tail boxes, descending
        const int q = N - 1 - (base + tid);
        const bool f = q >= Nn && !S.drop[q];
        int tot;
        const int r = block_rank(f, warp_cnt, tot);
        if (f) S.list[n + run + r] = q;
        run += tot;
      }
      __syncthreads();
      for (int k = tid; k < H; k += T) {
        const int dst = S.list[k], src = S.list[n + k];
        S.box[dst] = S.box[src]; S.score[dst] = S.score[src]; S.idx[dst] = S.idx[src];
      }
      N = Nn;
    }
    __syncthreads();
  }
  return N;
}

__global__ void __launch_bounds__(1024) soft_nms_kernel(const float* __restrict__ dets, const int n, const float thr,
                                                        const int method, const float sigma, const float min_score,
                                                        float* __restrict__ out_dets, long long* __restrict__ out_inds,
                                                        int32_t* __restrict__ out_count) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const SoftSmem S = soft_carve(smem_raw, n);
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    S.box[i] = make_float4(dets[(size_t)i * 5], dets[(size_t)i * 5 + 1], dets[(size_t)i * 5 + 2], dets[(size_t)i * 5 + 3]);
    S.score[i] = dets[(size_t)i * 5 + 4];
    S.idx[i] = i;
  }
  const int N = soft_nms_core(S, n, 0, method, thr, sigma, min_score);
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    const float4 b = S.box[i];
    out_dets[(size_t)i * 5] = b.x; out_dets[(size_t)i * 5 + 1] = b.y; out_dets[(size_t)i * 5 + 2] = b.z;
    out_dets[(size_t)i * 5 + 3] = b.w; out_dets[(size_t)i * 5 + 4] = S.score[i];
    out_inds[i] = (long long)S.idx[i];
  }
  if (threadIdx.x == 0) *out_count = N;
}

// multiclass_nms with nms_cfg type 'soft_nms' (bbox_nms.py:29-54): one CTA per (class, image).  Selection scores
// never increase, so only the first max_per_img+1 selections of a class can reach the final top max_per_img.
__global__ void __launch_bounds__(512) class_soft_nms_kernel(const __grid_constant__ PostParams P,
                                                             const float* __restrict__ boxes,
                                                             const float* __restrict__ scores_cm,
                                                             unsigned long long* __restrict__ kept_keys,
                                                             int32_t* __restrict__ kept_cnt, const int method,
                                                             const float sigma, const float min_score) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const SoftSmem S = soft_carve(smem_raw, P.M);
  __shared__ int warp_cnt[32];
  const int c = blockIdx.x, img = blockIdx.y, tid = threadIdx.x;
  const float* sc = scores_cm + ((size_t)img * P.C + c) * P.M;
  const float4* bx = reinterpret_cast<const float4*>(boxes) + (size_t)img * P.M;
  int n = 0;
  for (int base = 0; base < P.M; base += blockDim.x) {            // rows with score > score_thr, ascending (bbox_nms.py:37)
    const int j = base + tid;
    const float s = (j < P.M) ? __ldg(sc + j) : 0.f;
    const bool pass = (j < P.M) && (s > P.score_thr);
    int tot;
    const int r = block_rank(pass, warp_cnt, tot);
    if (pass) { S.box[n + r] = __ldg(bx + j); S.score[n + r] = s; S.idx[n + r] = j; }
    n += tot;
  }
  int32_t* cnt_out = kept_cnt + (size_t)img * P.C + c;
  if (n == 0) {
    if (tid == 0) *cnt_out = 0;
    return;
  }
  // S.list is indexed with the capacity the core is given; the staged count n <= P.M
  const int N = soft_nms_core(S, n, P.kcap, method, P.iou_thr, sigma, min_score);
  const int k = min(N, P.kcap);
  unsigned long long* keys_out = kept_keys + ((size_t)img * P.C + c) * P.kcap;
  for (int r = tid; r < k; r += blockDim.x)
    keys_out[r] = ((unsigned long long)float_to_ordered(S.score[r]) << 32) |
                  (unsigned long long)(0xffffffffu - (unsigned int)S.idx[r]);
  if (tid == 0) *cnt_out = k;
}

// ---------------------------------------------------------------------------------------- host
static int fill_params(const iou_postproc_cfg* cfg, int n_img, PostParams& P) {
  IOU_REQUIRE(cfg != nullptr, "cfg is NULL");
  IOU_REQUIRE(cfg->num_levels >= 1 && cfg->num_levels <= IOU_MAX_LEVELS, "num_levels out of range");
  IOU_REQUIRE(cfg->num_anchors >= 1 && cfg->num_anchors <= IOU_MAX_ANCHORS, "num_anchors out of range");
  IOU_REQUIRE(cfg->num_classes >= 1 && cfg->num_classes <= 256, "num_classes out of range");
  IOU_REQUIRE(n_img >= 1, "n_img must be >= 1");
  IOU_REQUIRE(cfg->max_per_img >= 1, "max_per_img must be >= 1");
  if (cfg->num_classes % 4 != 0)
    return fail(IOU_ERR_UNSUPPORTED, "num_classes %% 4 != 0 is not supported (got %d)", cfg->num_classes);
  if (cfg->nms_pre > TOPK_MAX)
    return fail(IOU_ERR_UNSUPPORTED, "nms_pre > %d is not supported (got %d)", TOPK_MAX, cfg->nms_pre);
  memset(&P, 0, sizeof(P));
  P.num_levels = cfg->num_levels; P.A = cfg->num_anchors; P.C = cfg->num_classes;
  P.nms_pre = cfg->nms_pre; P.n_img = n_img; P.max_per_img = cfg->max_per_img;
  P.kcap = cfg->max_per_img + 1;
  if (next_pow2_host(P.C * P.kcap) > 8192)
    return fail(IOU_ERR_UNSUPPORTED, "num_classes*(max_per_img+1) > 8192 is not supported");
  int aoff = 0, coff = 0;
  long long goff = 0;
  for (int l = 0; l < P.num_levels; ++l) {
    IOU_REQUIRE(cfg->feat_h[l] > 0 && cfg->feat_w[l] > 0, "empty feature map at level %d", l);
    P.feat_h[l] = cfg->feat_h[l]; P.feat_w[l] = cfg->feat_w[l]; P.stride[l] = cfg->stride[l];
    P.n_anchor[l] = cfg->feat_h[l] * cfg->feat_w[l] * P.A;
    P.anchor_off[l] = aoff; aoff += P.n_anchor[l];
    P.is_topk[l] = (P.nms_pre > 0 && P.n_anchor[l] > P.nms_pre);
    P.keep[l] = P.is_topk[l] ? P.nms_pre : P.n_anchor[l];
    P.cand_off[l] = coff; coff += P.keep[l];
    P.group_off[l] = goff;
    P.topk_slot[l] = -1;
    if (P.is_topk[l]) {
      goff += (long long)n_img * ((P.n_anchor[l] + 31) / 32);
      P.topk_slot[l] = P.num_topk_levels;
      P.topk_level[P.num_topk_levels++] = l;
    }
    for (int a = 0; a < P.A; ++a)
      for (int q = 0; q < 4; ++q) P.base[l][a][q] = cfg->base_anchors[l][a][q];
  }
  P.cand_off[P.num_levels] = coff;
  P.group_off[P.num_levels] = goff;
  for (int l = P.num_levels + 1; l <= IOU_MAX_LEVELS; ++l) { P.cand_off[l] = coff; P.group_off[l] = goff; }
  P.A_total = aoff; P.M = coff;
  if (P.M > IOU_MAX_CANDIDATES)
    return fail(IOU_ERR_UNSUPPORTED, "%d candidates per image exceed IOU_MAX_CANDIDATES=%d", P.M,
                IOU_MAX_CANDIDATES);
  for (int q = 0; q < 4; ++q) { P.mean[q] = cfg->target_means[q]; P.stdv[q] = cfg->target_stds[q]; }
  P.alpha = cfg->alpha; P.score_thr = cfg->score_thr; P.iou_thr = cfg->iou_thr;
  P.max_ratio = (float)fabs(log((double)cfg->wh_ratio_clip));
  IOU_REQUIRE(cfg->decode_mode == IOU_DECODE_DELTA || cfg->decode_mode == IOU_DECODE_DISTANCE, "bad decode_mode %d", cfg->decode_mode);
  IOU_REQUIRE(cfg->decode_mode == IOU_DECODE_DELTA || cfg->num_anchors == 1, "distance decoding needs num_anchors == 1");
  P.decode_mode = cfg->decode_mode;
  return IOU_OK;
}

struct PostWorkspace {
  float* maxscore; unsigned long long* kept_keys; int32_t* kept_cnt;
  float* boxes; float* scores_cm; int32_t* cand_idx;
  unsigned int* topk_hist;           // [n_img][num_topk_levels][TOPK_HSTRIDE] score-bucket histograms + counters (K1 -> K2)
  unsigned long long* topk_list;     // [n_img][num_topk_levels][TOPK_LIST] keys collected by the chunks of a level
  size_t total;
};
static PostWorkspace carve(const PostParams& P, void* base) {
  PostWorkspace W;
  size_t off = 0;
  unsigned char* b = static_cast<unsigned char*>(base);
  auto take = [&](size_t bytes) { void* p = b ? b + off : nullptr; off += align_up(bytes, 256); return p; };
  W.maxscore = (float*)take((size_t)P.n_img * P.A_total * 4);
  W.kept_keys = (unsigned long long*)take((size_t)P.n_img * P.C * P.kcap * 8);
  W.kept_cnt = (int32_t*)take((size_t)P.n_img * P.C * 4);
  W.boxes = (float*)take((size_t)P.n_img * P.M * 16);
  W.scores_cm = (float*)take((size_t)P.n_img * P.C * P.M * 4);
  W.cand_idx = (int32_t*)take((size_t)P.n_img * P.M * 4);
  W.topk_hist = (unsigned int*)take((size_t)P.n_img * (P.num_topk_levels > 0 ? P.num_topk_levels : 1) * TOPK_HSTRIDE * 4);
  W.topk_list = (unsigned long long*)take((size_t)P.n_img * (P.num_topk_levels > 0 ? P.num_topk_levels : 1) * TOPK_LIST * 8);
  W.total = off;
  return W;
}

static int run_decode(PostParams& P, const float* const* cls, const float* const* reg,
                      const float* const* iou, const float* const* cls_max2, const float* img_info, int rescale, float* boxes,
                      float* scores_cm, int32_t* cand_idx, float* maxscore, unsigned int* topk_hist, unsigned long long* topk_list,
                      cudaStream_t st) {
  for (int l = 0; l < P.num_levels; ++l) {
    IOU_REQUIRE(cls[l] && reg[l], "NULL level pointer at level %d", l);
    IOU_REQUIRE((iou && iou[l]) || P.alpha == 1.0f, "iou maps may only be omitted when alpha == 1 (level %d)", l);
    IOU_REQUIRE(((uintptr_t)cls[l] & 15) == 0 && ((uintptr_t)reg[l] & 15) == 0,
                "cls/reg pointers must be 16-byte aligned (level %d)", l);
    P.cls[l] = cls[l]; P.reg[l] = reg[l]; P.iou[l] = iou ? iou[l] : nullptr;
    P.cmax2[l] = cls_max2 ? cls_max2[l] : nullptr;
    IOU_REQUIRE(((uintptr_t)P.cmax2[l] & 7) == 0, "cls_max2 pointers must be 8-byte aligned (level %d)", l);
  }
  P.rescale = rescale;
  const long long groups = P.group_off[P.num_levels];
  if (groups > 0) {
    const int blocks = (int)((groups + 7) / 8 < 148 * 8 ? (groups + 7) / 8 : 148 * 8);
    const size_t sm = (size_t)8 * 32 * (P.C / 4) * sizeof(float);
    if (sm > 48 * 1024)     // more than 192 classes: above the default dynamic shared-memory limit
      IOU_CHECK_CUDA(cudaFuncSetAttribute(max_score_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    IOU_CHECK_CUDA(cudaMemsetAsync(topk_hist, 0, (size_t)P.n_img * P.num_topk_levels * TOPK_HSTRIDE * 4, st));
    max_score_kernel<<<blocks, 256, sm, st>>>(P, maxscore, topk_hist);
    if (int e = launch_status("max_score_kernel")) return e;
    topk_hist_kernel<<<dim3(P.num_topk_levels, P.n_img, TOPK_CHUNKS), 1024, 0, st>>>(P, maxscore, topk_hist, topk_list, cand_idx);
    if (int e = launch_status("topk_hist_kernel")) return e;
  }
  const size_t sm3 = (size_t)P.C * 33 * sizeof(float);
  gather_decode_kernel<<<dim3((P.M + 31) / 32, P.n_img), 256, sm3, st>>>(P, img_info, cand_idx, boxes,
                                                                        scores_cm);
  return launch_status("gather_decode_kernel");
}

static int run_nms(const PostParams& P, const float* boxes, const float* scores_cm, float* dets,
                   int64_t* labels, int32_t* counts, unsigned long long* kept_keys, int32_t* kept_cnt,
                   cudaStream_t st) {
  const int Pmax = (P.M + 7) & ~7;          // multiple of 8 keeps kbox 16-byte aligned behind the 6-byte entries
  const size_t sm4 = (size_t)Pmax * 6 + (size_t)P.kcap * 20 + 16;
  IOU_CHECK_CUDA(cudaFuncSetAttribute(class_nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm4));
  // 640 (class, image) CTAs at bs = 8 must be ONE wave (the kernel is a chain of block-wide latencies, a second wave
  // doubles its time): 44 KB of shared memory and 384 threads per CTA -> 5 CTAs per SM = 740 slots
  IOU_CHECK_CUDA(cudaFuncSetAttribute(class_nms_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  static const int nms_threads = getenv("IOU_NMS_THREADS") ? atoi(getenv("IOU_NMS_THREADS")) : 384;
  const int nt = (nms_threads == 256 || nms_threads == 128) ? nms_threads : 384;
  class_nms_kernel<<<dim3(P.C, P.n_img), nt, sm4, st>>>(P, boxes, scores_cm, kept_keys, kept_cnt, Pmax);
  if (int e = launch_status("class_nms_kernel")) return e;
  const size_t sm5 = (size_t)next_pow2_host(P.C * P.kcap) * 8 + 64;
  IOU_CHECK_CUDA(cudaFuncSetAttribute(final_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm5));
  final_select_kernel<<<P.n_img, 1024, sm5, st>>>(P, boxes, kept_keys, kept_cnt, dets,
                                                   reinterpret_cast<long long*>(labels), counts, 0);
  return launch_status("final_select_kernel");
}

}  // namespace iou

using namespace iou;

extern "C" int iou_postproc_num_candidates(const iou_postproc_cfg* cfg) {
  PostParams P;
  int e = fill_params(cfg, 1, P);
  return e ? e : P.M;
}

extern "C" size_t iou_postproc_workspace_bytes(const iou_postproc_cfg* cfg, int n_img) {
  PostParams P;
  if (fill_params(cfg, n_img, P)) return 0;
  return carve(P, nullptr).total;
}

extern "C" int iou_decode_candidates(const iou_postproc_cfg* cfg, int n_img, const float* const* cls,
                                     const float* const* reg, const float* const* iou,
                                     const float* img_info, int rescale, float* boxes, float* scores_cm,
                                     int32_t* cand_idx, void* workspace, size_t workspace_bytes,
                                     void* stream) {
  PostParams P;
  if (int e = fill_params(cfg, n_img, P)) return e;
  IOU_REQUIRE(cls && reg && img_info && boxes && scores_cm && cand_idx, "NULL argument");
  PostWorkspace W = carve(P, workspace);
  if (!workspace || workspace_bytes < W.total)
    return fail(IOU_ERR_WORKSPACE, "workspace too small: need %zu bytes", W.total);
  return run_decode(P, cls, reg, iou, nullptr, img_info, rescale, boxes, scores_cm, cand_idx, W.maxscore,
                    W.topk_hist, W.topk_list, (cudaStream_t)stream);
}

extern "C" int iou_batched_nms(const iou_postproc_cfg* cfg, int n_img, const float* boxes,
                               const float* scores_cm, float* dets, int64_t* labels, int32_t* counts,
                               void* workspace, size_t workspace_bytes, void* stream) {
  PostParams P;
  if (int e = fill_params(cfg, n_img, P)) return e;
  IOU_REQUIRE(boxes && scores_cm && dets && labels && counts, "NULL argument");
  PostWorkspace W = carve(P, workspace);
  if (!workspace || workspace_bytes < W.total)
    return fail(IOU_ERR_WORKSPACE, "workspace too small: need %zu bytes", W.total);
  return run_nms(P, boxes, scores_cm, dets, labels, counts, W.kept_keys, W.kept_cnt, (cudaStream_t)stream);
}

extern "C" int iou_get_bboxes(const iou_postproc_cfg* cfg, int n_img, const float* const* cls,
                              const float* const* reg, const float* const* iou, const float* img_info,
                              int rescale, float* dets, int64_t* labels, int32_t* counts,
                              void* workspace, size_t workspace_bytes, void* stream) {
  return iou_get_bboxes_premax(cfg, n_img, cls, reg, iou, nullptr, img_info, rescale, dets, labels, counts, workspace,
                               workspace_bytes, stream);
}

extern "C" int iou_get_bboxes_premax(const iou_postproc_cfg* cfg, int n_img, const float* const* cls,
                                     const float* const* reg, const float* const* iou, const float* const* cls_max2,
                                     const float* img_info, int rescale, float* dets, int64_t* labels, int32_t* counts,
                                     void* workspace, size_t workspace_bytes, void* stream) {
  PostParams P;
  if (int e = fill_params(cfg, n_img, P)) return e;
  IOU_REQUIRE(cls && reg && img_info && dets && labels && counts, "NULL argument");
  PostWorkspace W = carve(P, workspace);
  if (!workspace || workspace_bytes < W.total)
    return fail(IOU_ERR_WORKSPACE, "workspace too small: need %zu bytes", W.total);
  if (int e = run_decode(P, cls, reg, iou, cls_max2, img_info, rescale, W.boxes, W.scores_cm, W.cand_idx,
                         W.maxscore, W.topk_hist, W.topk_list, (cudaStream_t)stream))
    return e;
  return run_nms(P, W.boxes, W.scores_cm, dets, labels, counts, W.kept_keys, W.kept_cnt,
                 (cudaStream_t)stream);
}

extern "C" size_t iou_nms_workspace_bytes(int n) {
  if (n <= IOU_MAX_NMS_BOXES) return 256;          // one block, everything in shared memory
  return nms_large_carve(n).total;                 // keys + sorted boxes + n x ceil(n/64) mask words
}

extern "C" int iou_nms(const float* dets, int n, float iou_thr, int64_t* keep_idx, int32_t* keep_count,
                       void* workspace, size_t workspace_bytes, void* stream) {
  IOU_REQUIRE(n >= 0, "n must be >= 0");
  IOU_REQUIRE(keep_count != nullptr, "keep_count is NULL");
  cudaStream_t st = (cudaStream_t)stream;
  if (n == 0) {
    IOU_CHECK_CUDA(cudaMemsetAsync(keep_count, 0, sizeof(int32_t), st));
    return IOU_OK;
  }
  IOU_REQUIRE(dets && keep_idx, "NULL argument");
  if (n > IOU_MAX_NMS_BOXES) {
    // the reference has no size limit (nms_kernel.cu:70-131): mask tiles in global memory, sort and scan on the device
    const NmsLargeWs W = nms_large_carve(n);
    IOU_REQUIRE(W.cb * 8 <= 200 * 1024, "iou_nms: more than 1 638 400 boxes");
    if (!workspace || workspace_bytes < W.total)
      return fail(IOU_ERR_WORKSPACE, "iou_nms with %d boxes needs a workspace of %zu bytes (iou_nms_workspace_bytes)", n, W.total);
    unsigned char* base = reinterpret_cast<unsigned char*>(workspace);
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(base + W.keys);
    float4* sbox = reinterpret_cast<float4*>(base + W.sbox);
    float* sarea = reinterpret_cast<float*>(base + W.sarea);
    unsigned long long* mask = reinterpret_cast<unsigned long long*>(base + W.mask);
    unsigned int* keepbits = reinterpret_cast<unsigned int*>(base + W.keepbits);
    const int pb = (W.P + 255) / 256;
    nms_large_keys_kernel<<<pb, 256, 0, st>>>(dets, n, W.P, keys);
    for (int k = 2; k <= W.P; k <<= 1)
      for (int j = k >> 1; j > 0; j >>= 1) nms_large_bitonic_kernel<<<pb, 256, 0, st>>>(keys, W.P, j, k);
    nms_large_gather_kernel<<<(n + 255) / 256, 256, 0, st>>>(dets, n, keys, sbox, sarea);
    nms_large_mask_kernel<<<dim3(W.cb, W.cb), 64, 0, st>>>(sbox, sarea, n, W.cb, iou_thr, mask);
    const size_t sm = (size_t)W.cb * 8;
    if (sm > 48 * 1024)
      IOU_CHECK_CUDA(cudaFuncSetAttribute(nms_large_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    nms_large_scan_kernel<<<1, 1024, sm, st>>>(mask, n, W.cb, keys, keepbits, reinterpret_cast<long long*>(keep_idx), keep_count);
    return launch_status("nms_large_scan_kernel");
  }
  const int Pmax = next_pow2_host(n) < 2 ? 2 : next_pow2_host(n);
  const size_t sm = (size_t)Pmax * 8 + (size_t)n * 24 + (size_t)((n + 31) / 32) * 4 + 16;
  IOU_CHECK_CUDA(cudaFuncSetAttribute(single_nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  single_nms_kernel<<<1, 512, sm, st>>>(dets, n, iou_thr, reinterpret_cast<long long*>(keep_idx),
                                        keep_count, Pmax);
  return launch_status("single_nms_kernel");
}

extern "C" int iou_soft_nms(const float* dets, int n, float iou_thr, int method, float sigma, float min_score,
                            float* out_dets, int64_t* out_inds, int32_t* out_count, void* stream) {
  IOU_REQUIRE(n >= 0, "n must be >= 0");
  IOU_REQUIRE(out_count != nullptr, "out_count is NULL");
  IOU_REQUIRE(method >= 1 && method <= 3, "method must be 1 (linear), 2 (gaussian) or 3 (hard)");
  IOU_REQUIRE(method != 2 || sigma != 0.f, "sigma must be non-zero for the gaussian method");
  cudaStream_t st = (cudaStream_t)stream;
  if (n == 0) {
    IOU_CHECK_CUDA(cudaMemsetAsync(out_count, 0, sizeof(int32_t), st));
    return IOU_OK;
  }
  IOU_REQUIRE(dets && out_dets && out_inds, "NULL argument");
  if (n > IOU_MAX_NMS_BOXES)
    return fail(IOU_ERR_UNSUPPORTED, "iou_soft_nms supports at most %d boxes per call (got %d)", IOU_MAX_NMS_BOXES, n);
  const size_t sm = soft_smem_bytes(n);
  IOU_CHECK_CUDA(cudaFuncSetAttribute(soft_nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)soft_smem_bytes(IOU_MAX_NMS_BOXES)));
  soft_nms_kernel<<<1, n > 512 ? 1024 : 256, sm, st>>>(dets, n, iou_thr, method, sigma, min_score, out_dets,
                                                       reinterpret_cast<long long*>(out_inds), out_count);
  return launch_status("soft_nms_kernel");
}

extern "C" int iou_batched_soft_nms(const iou_postproc_cfg* cfg, int n_img, const float* boxes,
                                    const float* scores_cm, int method, float sigma, float min_score,
                                    float* dets, int64_t* labels, int32_t* counts, void* workspace,
                                    size_t workspace_bytes, void* stream) {
  PostParams P;
  if (int e = fill_params(cfg, n_img, P)) return e;
  IOU_REQUIRE(boxes && scores_cm && dets && labels && counts, "NULL argument");
  IOU_REQUIRE(method >= 1 && method <= 3, "method must be 1 (linear), 2 (gaussian) or 3 (hard)");
  IOU_REQUIRE(method != 2 || sigma != 0.f, "sigma must be non-zero for the gaussian method");
  PostWorkspace W = carve(P, workspace);
  if (!workspace || workspace_bytes < W.total)
    return fail(IOU_ERR_WORKSPACE, "workspace too small: need %zu bytes", W.total);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t sm = soft_smem_bytes(P.M);
  IOU_CHECK_CUDA(cudaFuncSetAttribute(class_soft_nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  class_soft_nms_kernel<<<dim3(P.C, P.n_img), 512, sm, st>>>(P, boxes, scores_cm, W.kept_keys, W.kept_cnt, method,
                                                             sigma, min_score);
  if (int e = launch_status("class_soft_nms_kernel")) return e;
  const size_t sm5 = (size_t)next_pow2_host(P.C * P.kcap) * 8 + 64;
  IOU_CHECK_CUDA(cudaFuncSetAttribute(final_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm5));
  final_select_kernel<<<P.n_img, 1024, sm5, st>>>(P, boxes, W.kept_keys, W.kept_cnt, dets,
                                                   reinterpret_cast<long long*>(labels), counts, 1);
  return launch_status("final_select_kernel");
}
