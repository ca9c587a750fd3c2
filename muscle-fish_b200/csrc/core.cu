// Error string, launch accounting, device properties.
#include <atomic>
#include <stdarg.h>

#include "common.cuh"

namespace mfish {

static thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

void count_launch(int k) { g_launches.fetch_add(k, std::memory_order_relaxed); }

}  // namespace mfish

extern "C" {

int mfish_abi_version(void) { return MFISH_ABI_VERSION; }

const char* mfish_last_error(void) { return mfish::g_err; }

int64_t mfish_launch_count(void) { return mfish::g_launches.load(); }

}  // extern "C"
