// Error reporting and version of the C ABI (include/gbnerf.h).
#include <atomic>
#include <stdarg.h>

#include "common.cuh"

namespace gbn {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
  return GBN_ECUDA;
}

}  // namespace gbn

extern "C" int gbn_version(void) { return 100; }  // 0.1.0

extern "C" const char* gbn_last_error_string(void) { return gbn::g_err; }

namespace gbn {
static std::atomic<unsigned long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
}  // namespace gbn

extern "C" unsigned long long gbn_kernel_launches(void) { return gbn::g_launches.load(std::memory_order_relaxed); }
