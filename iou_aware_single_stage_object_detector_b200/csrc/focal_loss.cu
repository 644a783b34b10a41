// Sigmoid focal loss forward / backward, drop-in for the reference's CUDA-only op
// (mmdet/ops/sigmoid_focal_loss/src/sigmoid_focal_loss_cuda.cu:24-105).  Training-side op:
// it is on the boundary named by the north star but is not executed by RetinaNet inference.
//
// The reference dispatches fp16 / fp32 / fp64 (AT_DISPATCH_FLOATING_TYPES_AND_HALF, .cu:128,167) over one templated
// kernel whose locals are scalar_t while every transcendental is the float one (expf / logf / powf, .cu:42-58) and
// the literals are double.  The kernels below keep that structure: T holds the rounded intermediates (p, term1,
// term2 and the running loss), the functions run in float, the literal arithmetic in the widest type involved.
#include <cuda_fp16.h>
#include <float.h>
#include "common.cuh"

namespace iou {

template <typename T> struct Num;       // scalar_t <-> arithmetic type (at::Half computes in float)
template <> struct Num<float> {
  typedef float A;
  static __device__ __forceinline__ float ld(const float* p, int i) { return p[i]; }
  static __device__ __forceinline__ void st(float* p, int i, float v) { p[i] = v; }
  static __device__ __forceinline__ float rnd(float v) { return v; }
};
template <> struct Num<double> {
  typedef double A;
  static __device__ __forceinline__ double ld(const double* p, int i) { return p[i]; }
  static __device__ __forceinline__ void st(double* p, int i, double v) { p[i] = v; }
  static __device__ __forceinline__ double rnd(double v) { return v; }
};
template <> struct Num<__half> {
  typedef float A;
  static __device__ __forceinline__ float ld(const __half* p, int i) { return __half2float(p[i]); }
  static __device__ __forceinline__ void st(__half* p, int i, float v) { p[i] = __float2half_rn(v); }
  static __device__ __forceinline__ float rnd(float v) { return __half2float(__float2half_rn(v)); }   // a scalar_t local
};

template <typename T>
__global__ void focal_fwd_kernel(const T* __restrict__ logits, const long long* __restrict__ targets,
                                 const int total, const int C, const float gamma, const float alpha,
                                 T* __restrict__ losses) {
  typedef Num<T> N;
  typedef typename N::A A;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += blockDim.x * gridDim.x) {
    const int n = i / C, d = i - n * C;
    const int t = (int)targets[n];                 // 0 = background, class d <-> d+1 (.cu:33-37)
    const A c1 = (t == d + 1) ? A(1) : A(0);
    const A c2 = (t >= 0 && t != d + 1) ? A(1) : A(0);
    const A zn = N::rnd((A)(1.0 - (double)alpha)), zp = N::rnd((A)alpha);
    const A x = N::ld(logits, i);
    const A p = N::rnd((A)(1.0 / (1.0 + (double)expf(-(float)x))));
    const A term1 = N::rnd((A)(powf((float)(1.0 - (double)p), gamma) * logf(fmaxf((float)p, FLT_MIN))));
    const double pos = (x >= A(0)) ? 1.0 : 0.0;
    const A term2 = N::rnd((A)((double)powf((float)p, gamma) *
                               (-1.0 * (double)x * pos -
                                (double)logf((float)(1.0 + (double)expf((float)((double)x - 2.0 * (double)x * pos)))))));
    A l = A(0);
    l = N::rnd(l + N::rnd(N::rnd(-c1 * term1) * zp));
    l = N::rnd(l + N::rnd(N::rnd(-c2 * term2) * zn));
    N::st(losses, i, l);
  }
}

template <typename T>
__global__ void focal_bwd_kernel(const T* __restrict__ logits, const long long* __restrict__ targets,
                                 const T* __restrict__ d_losses, const int total, const int C,
                                 const float gamma, const float alpha, T* __restrict__ d_logits) {
  typedef Num<T> N;
  typedef typename N::A A;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += blockDim.x * gridDim.x) {
    const int n = i / C, d = i - n * C;
    const int t = (int)targets[n];
    const A c1 = (t == d + 1) ? A(1) : A(0);
    const A c2 = (t >= 0 && t != d + 1) ? A(1) : A(0);
    const A zn = N::rnd((A)(1.0 - (double)alpha)), zp = N::rnd((A)alpha);
    const A x = N::ld(logits, i);
    const A p = N::rnd((A)(1.0 / (1.0 + (double)expf(-(float)x))));
    // (1-p)**g * (1 - p - g*p*log(p))                                                       (.cu:86-87)
    const A term1 = N::rnd((A)((double)powf((float)(1.0 - (double)p), gamma) *
                               (1.0 - (double)p - (double)((float)p * gamma * logf(fmaxf((float)p, FLT_MIN))))));
    const double pos = (x >= A(0)) ? 1.0 : 0.0;
    // (p**g) * (g*(1-p)*log(1-p) - p)                                                       (.cu:94-100)
    const A term2 = N::rnd((A)((double)powf((float)p, gamma) *
                               ((-1.0 * (double)x * pos -
                                 (double)logf((float)(1.0 + (double)expf((float)((double)x - 2.0 * (double)x * pos))))) *
                                    (1.0 - (double)p) * (double)gamma -
                                (double)p)));
    A g = A(0);
    g = N::rnd(g + N::rnd(N::rnd(-c1 * term1) * zp));
    g = N::rnd(g + N::rnd(N::rnd(-c2 * term2) * zn));
    N::st(d_logits, i, N::rnd(g * N::ld(d_losses, i)));
  }
}

static int focal_blocks(long long total) {       // .cu:120-121: min(ceil(total / 512), 4096) blocks of 512
  const long long b = (total + 511) / 512;
  return (int)(b < 4096 ? b : 4096);
}

}  // namespace iou

extern "C" int iou_sigmoid_focal_loss_forward_dtype(const void* logits, int dtype, const int64_t* targets, int n, int c,
                                                    float gamma, float alpha, void* losses, void* stream) {
  IOU_REQUIRE(n >= 0 && c >= 1, "bad shape");
  IOU_REQUIRE(dtype == IOU_DTYPE_F32 || dtype == IOU_DTYPE_F16 || dtype == IOU_DTYPE_F64, "dtype must be IOU_DTYPE_F32 / F16 / F64");
  if (n == 0) return IOU_OK;
  IOU_REQUIRE(logits && targets && losses, "NULL argument");
  const long long total = (long long)n * c;
  IOU_REQUIRE(total < (1ll << 31), "n*c too large");
  const int blocks = iou::focal_blocks(total);
  const long long* t = reinterpret_cast<const long long*>(targets);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == IOU_DTYPE_F32)
    iou::focal_fwd_kernel<float><<<blocks, 512, 0, st>>>((const float*)logits, t, (int)total, c, gamma, alpha, (float*)losses);
  else if (dtype == IOU_DTYPE_F16)
    iou::focal_fwd_kernel<__half><<<blocks, 512, 0, st>>>((const __half*)logits, t, (int)total, c, gamma, alpha, (__half*)losses);
  else
    iou::focal_fwd_kernel<double><<<blocks, 512, 0, st>>>((const double*)logits, t, (int)total, c, gamma, alpha, (double*)losses);
  return iou::launch_status("focal_fwd_kernel");
}

extern "C" int iou_sigmoid_focal_loss_backward_dtype(const void* logits, int dtype, const int64_t* targets,
                                                     const void* d_losses, int n, int c, float gamma, float alpha,
                                                     void* d_logits, void* stream) {
  IOU_REQUIRE(n >= 0 && c >= 1, "bad shape");
  IOU_REQUIRE(dtype == IOU_DTYPE_F32 || dtype == IOU_DTYPE_F16 || dtype == IOU_DTYPE_F64, "dtype must be IOU_DTYPE_F32 / F16 / F64");
  if (n == 0) return IOU_OK;
  IOU_REQUIRE(logits && targets && d_losses && d_logits, "NULL argument");
  const long long total = (long long)n * c;
  IOU_REQUIRE(total < (1ll << 31), "n*c too large");
  const int blocks = iou::focal_blocks(total);
  const long long* t = reinterpret_cast<const long long*>(targets);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == IOU_DTYPE_F32)
    iou::focal_bwd_kernel<float><<<blocks, 512, 0, st>>>((const float*)logits, t, (const float*)d_losses, (int)total, c, gamma, alpha, (float*)d_logits);
  else if (dtype == IOU_DTYPE_F16)
    iou::focal_bwd_kernel<__half><<<blocks, 512, 0, st>>>((const __half*)logits, t, (const __half*)d_losses, (int)total, c, gamma, alpha, (__half*)d_logits);
  else
    iou::focal_bwd_kernel<double><<<blocks, 512, 0, st>>>((const double*)logits, t, (const double*)d_losses, (int)total, c, gamma, alpha, (double*)d_logits);
  return iou::launch_status("focal_bwd_kernel");
}

extern "C" int iou_sigmoid_focal_loss_forward(const float* logits, const int64_t* targets, int n, int c,
                                              float gamma, float alpha, float* losses, void* stream) {
  return iou_sigmoid_focal_loss_forward_dtype(logits, IOU_DTYPE_F32, targets, n, c, gamma, alpha, losses, stream);
}

extern "C" int iou_sigmoid_focal_loss_backward(const float* logits, const int64_t* targets,
                                               const float* d_losses, int n, int c, float gamma,
                                               float alpha, float* d_logits, void* stream) {
  return iou_sigmoid_focal_loss_backward_dtype(logits, IOU_DTYPE_F32, targets, d_losses, n, c, gamma, alpha, d_logits, stream);
}
