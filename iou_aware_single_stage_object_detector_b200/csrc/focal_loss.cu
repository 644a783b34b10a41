// Sigmoid focal loss forward / backward, drop-in for the reference's CUDA-only op
// (mmdet/ops/sigmoid_focal_loss/src/sigmoid_focal_loss_cuda.cu:24-105).  Training-side op:
// it is on the boundary named by the north star but is not executed by RetinaNet inference.
#include <float.h>
#include "common.cuh"

namespace iou {

__global__ void focal_fwd_kernel(const float* __restrict__ logits, const long long* __restrict__ targets,
                                 const int total, const int C, const float gamma, const float alpha,
                                 float* __restrict__ losses) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += blockDim.x * gridDim.x) {
    const int n = i / C, d = i - n * C;
    const int t = (int)targets[n];                 // 0 = background, class d <-> d+1 (.cu:33-37)
    const float c1 = (t == d + 1) ? 1.f : 0.f;
    const float c2 = (t >= 0 && t != d + 1) ? 1.f : 0.f;
    const float x = logits[i];
    const float p = 1.f / (1.f + expf(-x));
    const float term1 = powf(1.f - p, gamma) * logf(fmaxf(p, FLT_MIN));
    const float pos = (x >= 0.f) ? 1.f : 0.f;
    const float term2 = powf(p, gamma) * (-1.f * x * pos - logf(1.f + expf(x - 2.f * x * pos)));
    float l = 0.f;
    l += -c1 * term1 * alpha;
    l += -c2 * term2 * (1.f - alpha);
    losses[i] = l;
  }
}

__global__ void focal_bwd_kernel(const float* __restrict__ logits, const long long* __restrict__ targets,
                                 const float* __restrict__ d_losses, const int total, const int C,
                                 const float gamma, const float alpha, float* __restrict__ d_logits) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += blockDim.x * gridDim.x) {
    const int n = i / C, d = i - n * C;
    const int t = (int)targets[n];
    const float c1 = (t == d + 1) ? 1.f : 0.f;
    const float c2 = (t >= 0 && t != d + 1) ? 1.f : 0.f;
    const float x = logits[i];
    const float p = 1.f / (1.f + expf(-x));
    const float term1 = powf(1.f - p, gamma) * (1.f - p - (p * gamma * logf(fmaxf(p, FLT_MIN))));
    const float pos = (x >= 0.f) ? 1.f : 0.f;
    const float term2 = powf(p, gamma) *
                        ((-1.f * x * pos - logf(1.f + expf(x - 2.f * x * pos))) * (1.f - p) * gamma - p);
    float g = 0.f;
    g += -c1 * term1 * alpha;
    g += -c2 * term2 * (1.f - alpha);
    d_logits[i] = g * d_losses[i];
  }
}

}  // namespace iou

extern "C" int iou_sigmoid_focal_loss_forward(const float* logits, const int64_t* targets, int n, int c,
                                              float gamma, float alpha, float* losses, void* stream) {
  IOU_REQUIRE(n >= 0 && c >= 1, "bad shape");
  if (n == 0) return IOU_OK;
  IOU_REQUIRE(logits && targets && losses, "NULL argument");
  const long long total = (long long)n * c;
  IOU_REQUIRE(total < (1ll << 31), "n*c too large");
  const int blocks = (int)((total + 511) / 512 < 4096 ? (total + 511) / 512 : 4096);  // .cu:120-121
  iou::focal_fwd_kernel<<<blocks, 512, 0, (cudaStream_t)stream>>>(
      logits, reinterpret_cast<const long long*>(targets), (int)total, c, gamma, alpha, losses);
  return iou::launch_status("focal_fwd_kernel");
}

extern "C" int iou_sigmoid_focal_loss_backward(const float* logits, const int64_t* targets,
                                               const float* d_losses, int n, int c, float gamma,
                                               float alpha, float* d_logits, void* stream) {
  IOU_REQUIRE(n >= 0 && c >= 1, "bad shape");
  if (n == 0) return IOU_OK;
  IOU_REQUIRE(logits && targets && d_losses && d_logits, "NULL argument");
  const long long total = (long long)n * c;
  IOU_REQUIRE(total < (1ll << 31), "n*c too large");
  const int blocks = (int)((total + 511) / 512 < 4096 ? (total + 511) / 512 : 4096);
  iou::focal_bwd_kernel<<<blocks, 512, 0, (cudaStream_t)stream>>>(
      logits, reinterpret_cast<const long long*>(targets), d_losses, (int)total, c, gamma, alpha, d_logits);
  return iou::launch_status("focal_bwd_kernel");
}
