// TEST INFRASTRUCTURE ONLY.
// Wrapper translation unit that compiles the UNMODIFIED reference source
// /root/reference/mmdet/ops/nms/src/nms_cpu.cpp from where it lies (no copy).
// torch >= 2 dropped the at::DeprecatedTypeProperties overload that the
// reference passes to AT_DISPATCH_FLOATING_TYPES (nms_cpu.cpp:63), so the macro
// is re-pointed at the ScalarType of that object before the include.
#include <torch/extension.h>

static inline at::ScalarType iou_oracle_scalar_type(const at::DeprecatedTypeProperties& t) {
  return t.scalarType();
}
static inline at::ScalarType iou_oracle_scalar_type(at::ScalarType t) { return t; }

#undef AT_DISPATCH_FLOATING_TYPES
#define AT_DISPATCH_FLOATING_TYPES(TYPE, NAME, ...) \
  AT_DISPATCH_SWITCH(iou_oracle_scalar_type(TYPE), NAME, AT_DISPATCH_CASE_FLOATING_TYPES(__VA_ARGS__))

#include IOU_REFERENCE_NMS_CPU_SOURCE
