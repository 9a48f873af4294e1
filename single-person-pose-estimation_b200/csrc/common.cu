// Error plumbing and debug knobs shared by every translation unit of libhgb200.
#include "common.cuh"

namespace hgb {

static thread_local char t_err[1024] = "";
int g_debug[48] = {0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(t_err, sizeof(t_err), fmt, ap);
  va_end(ap);
}

}  // namespace hgb

extern "C" const char* hgb_last_error(void) { return hgb::t_err; }
extern "C" int hgb_version(void) { return 100; }
extern "C" int hgb_debug_set(int key, int value) {
  if (key < 0 || key >= 48) {
    hgb::set_error("hgb_debug_set: unknown key %d", key);
    return HGB_ERR_INVALID;
  }
  hgb::g_debug[key] = value;
  return HGB_OK;
}
