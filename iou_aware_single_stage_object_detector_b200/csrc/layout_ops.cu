// Elementwise layout kernels around the conv engine (all HBM-bound, fully coalesced along the
// channel dimension): NCHW fp32 <-> padded-rows bf16 hi|lo, the stem's im2col, the stem's 3x3/s2
// max pool and the stride-2 phase split.  Layout definition: include/iou_b200.h.
#include <cuda_bf16.h>
#include "common.cuh"

namespace iou {

__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}
__device__ __forceinline__ float join_bf16(__nv_bfloat16 hi, __nv_bfloat16 lo) {
  return __bfloat162float(hi) + __bfloat162float(lo);
}

// ---- NCHW fp32 -> padded rows.  One CTA per (padded row y, image); tile transpose via smem.
__global__ void __launch_bounds__(256) pack_nchw_kernel(const float* __restrict__ src, int n, int c, int h, int w,
                                                        __nv_bfloat16* __restrict__ dst) {
  extern __shared__ float tile[];                 // [32 channels][w + 1]
  const int yp = blockIdx.x, img = blockIdx.y;
  const int wp = w + 2;
  __nv_bfloat16* drow = dst + ((size_t)img * (h + 2) + yp) * wp * (2 * c);
  if (yp == 0 || yp == h + 1) {
    for (size_t i = threadIdx.x; i < (size_t)wp * 2 * c; i += blockDim.x) drow[i] = __float2bfloat16_rn(0.f);
    return;
  }
  const int y = yp - 1;
  for (int i = threadIdx.x; i < 2 * c; i += blockDim.x) {   // left / right border pixels
    drow[i] = __float2bfloat16_rn(0.f);
    drow[(size_t)(wp - 1) * 2 * c + i] = __float2bfloat16_rn(0.f);
  }
  for (int c0 = 0; c0 < c; c0 += 32) {
    const int cc = min(32, c - c0);
    for (int i = threadIdx.x; i < cc * w; i += blockDim.x) {
      const int ch = i / w, x = i - ch * w;
      tile[ch * (w + 1) + x] = src[(((size_t)img * c + c0 + ch) * h + y) * w + x];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < cc * w; i += blockDim.x) {
      const int x = i / cc, ch = i - x * cc;
      __nv_bfloat16 hi, lo;
      split_bf16(tile[ch * (w + 1) + x], hi, lo);
      __nv_bfloat16* px = drow + (size_t)(x + 1) * 2 * c;
      px[c0 + ch] = hi;
      px[c + c0 + ch] = lo;
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256) unpack_nchw_kernel(const __nv_bfloat16* __restrict__ src, int n, int c,
                                                          int h, int w, float* __restrict__ dst) {
  extern __shared__ float tile[];                 // [32 channels][w + 1]
  const int y = blockIdx.x, img = blockIdx.y;
  const int wp = w + 2;
  const __nv_bfloat16* srow = src + ((size_t)img * (h + 2) + y + 1) * wp * (2 * c);
  for (int c0 = 0; c0 < c; c0 += 32) {
    const int cc = min(32, c - c0);
    for (int i = threadIdx.x; i < cc * w; i += blockDim.x) {
      const int x = i / cc, ch = i - x * cc;
      const __nv_bfloat16* px = srow + (size_t)(x + 1) * 2 * c;
      tile[ch * (w + 1) + x] = join_bf16(px[c0 + ch], px[c + c0 + ch]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < cc * w; i += blockDim.x) {
      const int ch = i / w, x = i - ch * w;
      dst[(((size_t)img * c + c0 + ch) * h + y) * w + x] = tile[ch * (w + 1) + x];
    }
    __syncthreads();
  }
}

// ---- stem im2col: conv 7x7 stride 2 pad 3 on 3 channels (resnet.py:454-462).
// One CTA per (output row yo in padded coords, image).  k = (r*7+s)*3 + ch, zero-padded to kpad.
__global__ void __launch_bounds__(256) im2col_stem_kernel(const float* __restrict__ img, int n, int h, int w,
                                                          int ho, int wo, int kpad, __nv_bfloat16* __restrict__ dst) {
  extern __shared__ float rows[];                 // [3 ch][7 rows][w + 6] zero-padded input rows
  const int yp = blockIdx.x, im = blockIdx.y;
  const int wop = wo + 2, wpad = w + 6;
  __nv_bfloat16* drow = dst + ((size_t)im * (ho + 2) + yp) * wop * (2 * kpad);
  const __nv_bfloat16 z = __float2bfloat16_rn(0.f);
  if (yp == 0 || yp == ho + 1) {
    for (size_t i = threadIdx.x; i < (size_t)wop * 2 * kpad; i += blockDim.x) drow[i] = z;
    return;
  }
  const int yo = yp - 1;
  for (int i = threadIdx.x; i < 3 * 7 * wpad; i += blockDim.x) {
    const int ch = i / (7 * wpad), rem = i - ch * 7 * wpad;
    const int r = rem / wpad, xx = rem - r * wpad;
    const int y = yo * 2 - 3 + r, x = xx - 3;
    rows[i] = (y >= 0 && y < h && x >= 0 && x < w) ? img[(((size_t)im * 3 + ch) * h + y) * w + x] : 0.f;
  }
  __syncthreads();
  for (size_t i = threadIdx.x; i < (size_t)wop * kpad; i += blockDim.x) {
    const int xp = (int)(i / kpad), k = (int)(i - (size_t)xp * kpad);
    float v = 0.f;
    if (xp >= 1 && xp <= wo && k < 147) {
      const int tap = k / 3, ch = k - tap * 3;
      const int r = tap / 7, s = tap - r * 7;
      v = rows[(ch * 7 + r) * wpad + (xp - 1) * 2 + s];
    }
    __nv_bfloat16 hi, lo;
    split_bf16(v, hi, lo);
    drow[(size_t)xp * 2 * kpad + k] = hi;
    drow[(size_t)xp * 2 * kpad + kpad + k] = lo;
  }
}

// ---- 3x3 stride-2 pad-1 max pool on padded rows (inputs are post-ReLU, so the zero border
// is equivalent to -inf padding).  One thread per (output pixel, channel pair).
__global__ void __launch_bounds__(256) maxpool_kernel(const __nv_bfloat16* __restrict__ src, int n, int c, int h,
                                                      int w, int ho, int wo, __nv_bfloat16* __restrict__ dst) {
  const int c2 = c >> 1;
  const size_t total = (size_t)n * (ho + 2) * (wo + 2) * c2;
  const int wp = w + 2, wop = wo + 2;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int cp = (int)(i % c2);
    size_t t = i / c2;
    const int xp = (int)(t % wop); t /= wop;
    const int yp = (int)(t % (ho + 2));
    const int img = (int)(t / (ho + 2));
    float m0 = 0.f, m1 = 0.f;
    if (xp >= 1 && xp <= wo && yp >= 1 && yp <= ho) {
      // window rows 2*yo-1 .. 2*yo+1 (unpadded) == padded rows 2*yo .. 2*yo+2
      const int py0 = 2 * (yp - 1), px0 = 2 * (xp - 1);
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int s = 0; s < 3; ++s) {
          const int py = py0 + r, px = px0 + s;
          if (py <= h + 1 && px <= w + 1) {
            const __nv_bfloat16* p = src + (((size_t)img * (h + 2) + py) * wp + px) * (2 * c);
            const __nv_bfloat162 hv = *reinterpret_cast<const __nv_bfloat162*>(p + 2 * cp);
            const __nv_bfloat162 lv = *reinterpret_cast<const __nv_bfloat162*>(p + c + 2 * cp);
            m0 = fmaxf(m0, __bfloat162float(hv.x) + __bfloat162float(lv.x));
            m1 = fmaxf(m1, __bfloat162float(hv.y) + __bfloat162float(lv.y));
          }
        }
    }
    __nv_bfloat162 ho2, lo2;
    split_bf16(m0, ho2.x, lo2.x);
    split_bf16(m1, ho2.y, lo2.y);
    __nv_bfloat16* o = dst + (((size_t)img * (ho + 2) + yp) * wop + xp) * (2 * c);
    *reinterpret_cast<__nv_bfloat162*>(o + 2 * cp) = ho2;
    *reinterpret_cast<__nv_bfloat162*>(o + c + 2 * cp) = lo2;
  }
}

// ---- stride-2 phase split: dst[py][px][img][u][v][:] = in_padded[img][2(u-1)+py][2(v-1)+px][:]
// for u,v >= 1 and in-range sources, else 0.  16-byte vectors along the 2*c channel axis.
struct PhasePtrs { uint4* p[4]; };
__global__ void __launch_bounds__(256) phase_split_kernel(const uint4* __restrict__ src, int n, int c, int h, int w,
                                                          int ho, int wo, PhasePtrs dst, int phase_mask) {
  const int vec = (2 * c) >> 3;                    // uint4 per pixel
  const int hop = ho + 2, wop = wo + 2, hp = h + 2, wp = w + 2;
  const size_t total = (size_t)n * hop * wop * vec;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int q = (int)(i % vec);
    size_t t = i / vec;
    const int v = (int)(t % wop); t /= wop;
    const int u = (int)(t % hop);
    const int img = (int)(t / hop);
#pragma unroll
    for (int ph = 0; ph < 4; ++ph) {
      if (!((phase_mask >> ph) & 1)) continue;
      const int py = ph >> 1, px = ph & 1;
      const int sy = 2 * (u - 1) + py, sx = 2 * (v - 1) + px;
      uint4 val = make_uint4(0, 0, 0, 0);
      if (u >= 1 && v >= 1 && sy < hp && sx < wp) val = __ldg(src + (((size_t)img * hp + sy) * wp + sx) * vec + q);
      dst.p[ph][i] = val;
    }
  }
}

}  // namespace iou

using namespace iou;

extern "C" int iou_pack_nchw(const float* src, int n, int c, int h, int w, void* dst, int64_t dst_row_start,
                             void* stream) {
  IOU_REQUIRE(src && dst && n > 0 && c > 0 && h > 0 && w > 0, "bad argument");
  __nv_bfloat16* d = (__nv_bfloat16*)dst + (size_t)dst_row_start * 2 * c;
  pack_nchw_kernel<<<dim3(h + 2, n), 256, (size_t)32 * (w + 1) * 4, (cudaStream_t)stream>>>(src, n, c, h, w, d);
  return launch_status("pack_nchw_kernel");
}

extern "C" int iou_unpack_nchw(const void* src, int64_t src_row_start, int n, int c, int h, int w, float* dst,
                               void* stream) {
  IOU_REQUIRE(src && dst && n > 0 && c > 0 && h > 0 && w > 0, "bad argument");
  const __nv_bfloat16* s = (const __nv_bfloat16*)src + (size_t)src_row_start * 2 * c;
  unpack_nchw_kernel<<<dim3(h, n), 256, (size_t)32 * (w + 1) * 4, (cudaStream_t)stream>>>(s, n, c, h, w, dst);
  return launch_status("unpack_nchw_kernel");
}

extern "C" int iou_im2col_stem(const float* img, int n, int h, int w, int kpad, void* dst, void* stream) {
  IOU_REQUIRE(img && dst && n > 0 && h > 0 && w > 0, "bad argument");
  IOU_REQUIRE(kpad >= 147 && kpad % 64 == 0, "kpad must be a multiple of 64 >= 147");
  const int ho = (h + 6 - 7) / 2 + 1, wo = (w + 6 - 7) / 2 + 1;
  const size_t sm = (size_t)3 * 7 * (w + 6) * 4;
  IOU_REQUIRE(sm <= 200 * 1024, "image too wide for the stem im2col kernel");
  static bool attr = false;
  if (!attr) {
    IOU_CHECK_CUDA(cudaFuncSetAttribute(im2col_stem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr = true;
  }
  im2col_stem_kernel<<<dim3(ho + 2, n), 256, sm, (cudaStream_t)stream>>>(img, n, h, w, ho, wo, kpad, (__nv_bfloat16*)dst);
  return launch_status("im2col_stem_kernel");
}

extern "C" int iou_maxpool3x3s2(const void* src, int n, int c, int h, int w, void* dst, void* stream) {
  IOU_REQUIRE(src && dst && n > 0 && c > 0 && (c & 1) == 0 && h > 0 && w > 0, "bad argument");
  const int ho = (h + 2 - 3) / 2 + 1, wo = (w + 2 - 3) / 2 + 1;
  const size_t total = (size_t)n * (ho + 2) * (wo + 2) * (c / 2);
  const int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  maxpool_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)src, n, c, h, w, ho, wo,
                                                           (__nv_bfloat16*)dst);
  return launch_status("maxpool_kernel");
}

extern "C" int iou_phase_split(const void* src, int n, int c, int h, int w, void* const* dst4, int phase_mask,
                               void* stream) {
  IOU_REQUIRE(src && dst4 && n > 0 && c > 0 && (c % 4) == 0 && h > 0 && w > 0, "bad argument");
  IOU_REQUIRE(phase_mask > 0 && phase_mask < 16, "phase_mask out of range");
  PhasePtrs P;
  for (int i = 0; i < 4; ++i) {
    P.p[i] = (uint4*)dst4[i];
    if (((phase_mask >> i) & 1) && !dst4[i]) return fail(IOU_ERR_INVALID, "dst4[%d] is NULL", i);
  }
  const int ho = (h + 1) / 2, wo = (w + 1) / 2;
  const size_t total = (size_t)n * (ho + 2) * (wo + 2) * ((2 * c) / 8);
  const int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  phase_split_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((const uint4*)src, n, c, h, w, ho, wo, P, phase_mask);
  return launch_status("phase_split_kernel");
}
