#include <stdarg.h>
#include "common.cuh"

namespace iou {
static thread_local std::string g_last_error;
void set_error(const std::string& msg) { g_last_error = msg; }
int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}
}  // namespace iou

extern "C" const char* iou_last_error(void) { return iou::g_last_error.c_str(); }
extern "C" int iou_abi_version(void) { return 8; }
extern "C" size_t iou_sizeof(int what) {
  switch (what) {
    case 0: return sizeof(iou_postproc_cfg);
    case 1: return sizeof(iou_conv_desc);
    case 2: return sizeof(iou_conv_segment);
    default: return 0;
  }
}
