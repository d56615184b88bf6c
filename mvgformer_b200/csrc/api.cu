// Error plumbing + version for libmvg_b200.
#include <cstdlib>

#include "common.cuh"

namespace mvg {

static thread_local char t_err[512] = "";
std::atomic<int64_t> g_launches{0};

bool pdl_enabled() {
  static const bool on = []() {
    const char* e = getenv("MVG_PDL");
    return e == nullptr || e[0] != '0';
  }();
  return on;
}

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(t_err, sizeof(t_err), fmt, ap);
  va_end(ap);
}

}  // namespace mvg

extern "C" const char* mvg_last_error(void) { return mvg::t_err; }
extern "C" int mvg_abi_version(void) { return 10; }
extern "C" int64_t mvg_launch_count(void) { return mvg::g_launches.load(); }
