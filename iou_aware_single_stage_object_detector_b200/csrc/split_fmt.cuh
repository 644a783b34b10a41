// The two split-operand element formats of the padded-rows activation layout (4 bytes per element each).
// A row of C channels is [hi plane: C x 16 bit][lo plane: C x 16 bit]; a group of 8 channels owns one 16-byte
// vector in each plane (so every kernel addresses both formats identically):
//
//   IOU_FMT_BF16X2 (0): hi = bf16(v), lo = bf16(v - hi).                     conv = hi*Whi + hi*Wlo + lo*Whi
//       (three bf16 tensor-core passes)
//   IOU_FMT_F16F8  (1): hi = fp16(v); lo vector = [x8 x 8 | l8 x 8] with x8 = e4m3(v), l8 = e4m3((v - hi) * 2^11).
//       conv = hi*Wh (one fp16 pass) + 2^-11/s_n * ([x8 | l8] * [Wl8 ; W8]) (one e4m3 pass with K doubled, which
//       runs at twice the rate): two bf16-pass equivalents instead of three.  The e4m3 terms only carry the
//       2^-12-relative corrections, so their 4 significant bits put the total error at ~2^-16 relative
//       (tools/numerics_sim.py).  v is recovered as hi + l8 * 2^-11.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include <stdint.h>

namespace iou {

constexpr int kFmtBf16x2 = 0;
constexpr int kFmtF16F8 = 1;
constexpr float kF8LoScale = 2048.f;            // 2^11
constexpr float kF8LoInv = 1.f / 2048.f;

__device__ __forceinline__ uint32_t e4m3x2(float a, float b) {         // cvt.rn.satfinite.e4m3x2.f32 (low byte = a)
  return (uint32_t)__nv_cvt_float2_to_fp8x2(make_float2(a, b), __NV_SATFINITE, __NV_E4M3);
}
__device__ __forceinline__ float2 e4m3x2_to_float2(uint32_t v) {
  const __half2_raw h = __nv_cvt_fp8x2_to_halfraw2((__nv_fp8x2_storage_t)(v & 0xffffu), __NV_E4M3);
  return __half22float2(*reinterpret_cast<const __half2*>(&h));
}

// mixed-precision ALU ops of sm_100 (SASS FHADD / FHFMA): an fp16 operand is widened inside the instruction
__device__ __forceinline__ float fhadd(uint16_t h, float f) {           // f + float(h)
  float d;
  asm("add.rn.f32.f16 %0, %1, %2;" : "=f"(d) : "h"(h), "f"(f));
  return d;
}
__device__ __forceinline__ float fhfma(uint16_t a, uint16_t b, float c) {   // float(a) * float(b) + c
  float d;
  asm("fma.rn.f32.f16 %0, %1, %2, %3;" : "=f"(d) : "h"(a), "h"(b), "f"(c));
  return d;
}
constexpr uint16_t kF16LoInv = 0x1000;       // 2^-11 as fp16
constexpr uint16_t kF16NegLoScale = 0xE800;  // -2048 as fp16

// fp32 pair -> packed fp16 pair, saturating at +-65504 inside the conversion (SASS F2FP.SATFINITE.F16.F32.PACK_AB: one
// instruction, like the non-saturating form)
__device__ __forceinline__ uint32_t f16x2_sat(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

// 8 consecutive channels -> the group's hi and lo vectors.  Values beyond the fp16 range SATURATE at +-65504 (never inf,
// so nothing downstream turns into NaN); iou_range_stats counts them, and the detector checks that count on the first
// batch of every plan.  kClamp is kept for source compatibility: the conversion itself saturates in both cases.
template <int kFmt, bool kClamp = true>
__device__ __forceinline__ void encode8(const float (&v)[8], uint4& hi, uint4& lo) {
  uint32_t h[4], l[4];
  if constexpr (kFmt == kFmtBf16x2) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const __nv_bfloat162 h2 = __floats2bfloat162_rn(v[2 * q], v[2 * q + 1]);
      const uint32_t hb = *reinterpret_cast<const uint32_t*>(&h2);
      const __nv_bfloat162 l2 = __floats2bfloat162_rn(v[2 * q] - __uint_as_float(hb << 16),
                                                      v[2 * q + 1] - __uint_as_float(hb & 0xffff0000u));
      h[q] = hb;
      l[q] = *reinterpret_cast<const uint32_t*>(&l2);
    }
  } else {
    uint32_t x8[4], l8[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float a = v[2 * q], b = v[2 * q + 1];
      h[q] = f16x2_sat(a, b);
      x8[q] = e4m3x2(a, b);
      // (v - hi) * 2^11 = fma(hi, -2^11, v * 2^11): exact (the difference has <= 13 significant bits)
      l8[q] = e4m3x2(fhfma((uint16_t)(h[q] & 0xffffu), kF16NegLoScale, a * kF8LoScale),
                     fhfma((uint16_t)(h[q] >> 16), kF16NegLoScale, b * kF8LoScale));
    }
    l[0] = x8[0] | (x8[1] << 16); l[1] = x8[2] | (x8[3] << 16);
    l[2] = l8[0] | (l8[1] << 16); l[3] = l8[2] | (l8[3] << 16);
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}

template <int kFmt>
__device__ __forceinline__ void decode8(const uint4 hi, const uint4 lo, float (&v)[8]) {
  const uint32_t h[4] = {hi.x, hi.y, hi.z, hi.w};
  if constexpr (kFmt == kFmtBf16x2) {
    const uint32_t l[4] = {lo.x, lo.y, lo.z, lo.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      v[2 * q] = __uint_as_float(h[q] << 16) + __uint_as_float(l[q] << 16);
      v[2 * q + 1] = __uint_as_float(h[q] & 0xffff0000u) + __uint_as_float(l[q] & 0xffff0000u);
    }
  } else {
    const uint32_t l8[4] = {lo.z & 0xffffu, lo.z >> 16, lo.w & 0xffffu, lo.w >> 16};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&h[q]));
      const float2 lf = e4m3x2_to_float2(l8[q]);
      v[2 * q] = fmaf(lf.x, kF8LoInv, hf.x);
      v[2 * q + 1] = fmaf(lf.y, kF8LoInv, hf.y);
    }
  }
}

// f[i] += value of channel i (residual adds).  fp16 + e4m3 format: two mixed-precision ops per element
// (f + hi, then + l8 * 2^-11) instead of convert / convert / fma / add.
template <int kFmt>
__device__ __forceinline__ void decode8_add(const uint4 hi, const uint4 lo, float* f) {
  if constexpr (kFmt == kFmtBf16x2) {
    float t8[8];
    decode8<kFmt>(hi, lo, t8);
#pragma unroll
    for (int q = 0; q < 8; ++q) f[q] += t8[q];
  } else {
    const uint32_t h[4] = {hi.x, hi.y, hi.z, hi.w};
    const uint32_t l8[4] = {lo.z & 0xffffu, lo.z >> 16, lo.w & 0xffffu, lo.w >> 16};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const __half2_raw lr = __nv_cvt_fp8x2_to_halfraw2((__nv_fp8x2_storage_t)l8[q], __NV_E4M3);
      f[2 * q] = fhfma(lr.x, kF16LoInv, fhadd((uint16_t)(h[q] & 0xffffu), f[2 * q]));
      f[2 * q + 1] = fhfma(lr.y, kF16LoInv, fhadd((uint16_t)(h[q] >> 16), f[2 * q + 1]));
    }
  }
}

// sign test used by the fused ReLU of the phase split: v < 0 exactly when its hi half is negative (both formats)

}  // namespace iou
