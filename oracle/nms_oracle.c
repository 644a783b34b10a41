/* TEST INFRASTRUCTURE ONLY -- CPU oracle, never linked into the product library.
 *
 * Plain-C restatement of the reference's greedy NMS:
 *   - IoU arithmetic: mmdet/ops/nms/src/nms_kernel.cu:13-21 (devIoU) ==
 *     mmdet/ops/nms/src/nms_cpu.cpp:18,45-54 (legacy "+1" widths, fp32,
 *     separate mul/add/div -- compiled here with -ffp-contract=off);
 *   - visiting order: boxes sorted by score descending (nms_kernel.cu:73-75,
 *     nms_cpu.cpp:20); ties are unspecified in the reference, this oracle
 *     breaks them by ascending input index (the rule the CUDA path follows);
 *   - suppression test: mode 0 = "IoU >  thr" (nms_kernel.cu:60, the op the
 *     CUDA library is a drop-in for), mode 1 = "IoU >= thr" (nms_cpu.cpp:55);
 *   - result: ORIGINAL indices of the kept boxes in ascending order
 *     (nms_kernel.cu:127-130, nms_cpu.cpp:58).
 */
#include <stdint.h>
#include <stdlib.h>

typedef struct { float score; int64_t idx; } key_t_;

static int cmp_desc(const void* a, const void* b) {
  const key_t_* x = (const key_t_*)a; const key_t_* y = (const key_t_*)b;
  if (x->score > y->score) return -1;
  if (x->score < y->score) return 1;
  return (x->idx > y->idx) - (x->idx < y->idx);
}

static inline float fmaxf_(float a, float b) { return a > b ? a : b; }
static inline float fminf_(float a, float b) { return a < b ? a : b; }

/* dets: n x 5 (x1,y1,x2,y2,score) fp32.  keep_out: capacity n.  Returns #kept. */
int64_t oracle_nms(const float* dets, int64_t n, float thr, int mode, int64_t* keep_out) {
  if (n <= 0) return 0;
  key_t_* order = (key_t_*)malloc(sizeof(key_t_) * (size_t)n);
  float* area = (float*)malloc(sizeof(float) * (size_t)n);
  uint8_t* dead = (uint8_t*)calloc((size_t)n, 1);
  for (int64_t i = 0; i < n; ++i) {
    const float* d = dets + 5 * i;
    order[i].score = d[4]; order[i].idx = i;
    area[i] = (d[2] - d[0] + 1.0f) * (d[3] - d[1] + 1.0f);
  }
  qsort(order, (size_t)n, sizeof(key_t_), cmp_desc);
  for (int64_t a = 0; a < n; ++a) {
    int64_t i = order[a].idx;
    if (dead[i]) continue;
    const float* bi = dets + 5 * i;
    for (int64_t b = a + 1; b < n; ++b) {
      int64_t j = order[b].idx;
      if (dead[j]) continue;
      const float* bj = dets + 5 * j;
      float left = fmaxf_(bi[0], bj[0]), right = fminf_(bi[2], bj[2]);
      float top = fmaxf_(bi[1], bj[1]), bottom = fminf_(bi[3], bj[3]);
      float w = fmaxf_(right - left + 1.0f, 0.0f), h = fmaxf_(bottom - top + 1.0f, 0.0f);
      float inter = w * h;
      float iou = inter / (area[i] + area[j] - inter);
      if (mode == 0 ? (iou > thr) : (iou >= thr)) dead[j] = 1;
    }
  }
  int64_t k = 0;
  for (int64_t i = 0; i < n; ++i) if (!dead[i]) keep_out[k++] = i;
  free(order); free(area); free(dead);
  return k;
}

/* Pairwise IoU of two boxes with the same arithmetic (for threshold-tie audits). */
float oracle_iou(const float* a, const float* b) {
  float left = fmaxf_(a[0], b[0]), right = fminf_(a[2], b[2]);
  float top = fmaxf_(a[1], b[1]), bottom = fminf_(a[3], b[3]);
  float w = fmaxf_(right - left + 1.0f, 0.0f), h = fmaxf_(bottom - top + 1.0f, 0.0f);
  float inter = w * h;
  float sa = (a[2] - a[0] + 1.0f) * (a[3] - a[1] + 1.0f);
  float sb = (b[2] - b[0] + 1.0f) * (b[3] - b[1] + 1.0f);
  return inter / (sa + sb - inter);
}
