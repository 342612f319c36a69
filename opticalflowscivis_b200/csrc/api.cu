// Library-wide state of libofsv.so: thread-local error string, launch counter, version.
#include "ofsv_common.cuh"

namespace ofsv {

static thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int device_num_sms() {
  static std::atomic<int> cache[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  const int slot = dev & 63;
  int n = dev < 64 ? cache[slot].load(std::memory_order_relaxed) : 0;
  if (n == 0) {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    if (dev < 64) cache[slot].store(n, std::memory_order_relaxed);
  }
  return n;
}

}  // namespace ofsv

extern "C" const char* ofsv_version(void) { return "ofsv 0.2 (sm_100a)"; }
extern "C" const char* ofsv_last_error(void) { return ofsv::g_err; }
extern "C" int64_t ofsv_launch_count(void) { return ofsv::g_launches.load(); }
