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

}  // namespace ofsv

extern "C" const char* ofsv_version(void) { return "ofsv 0.1 (sm_100a)"; }
extern "C" const char* ofsv_last_error(void) { return ofsv::g_err; }
extern "C" int64_t ofsv_launch_count(void) { return ofsv::g_launches.load(); }
