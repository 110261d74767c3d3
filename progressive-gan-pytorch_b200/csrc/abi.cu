// Error channel + device query of the C-ABI (include/progan_b200.h).
#include "common.cuh"
#include <stdlib.h>

namespace pg {
static thread_local char g_err[512] = "";
void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace pg

extern "C" const char *pg_last_error(void) { return pg::g_err; }
extern "C" int pg_abi_version(void) { return 1; }

namespace pg {
// PG_PDL: 0 = no programmatic-serialization attribute, 1 = small launches (default), 2 = all
// (read once per process)
int pdl_mode() {
  static const int mode = [] {
    const char *e = getenv("PG_PDL");
    return e ? atoi(e) : 1;
  }();
  return mode;
}
}  // namespace pg

extern "C" int pg_device_info(int *sm_count, int *cc_major, int *cc_minor) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    pg::set_error("pg_device_info: %s", cudaGetErrorString(e));
    return PG_ERR_CUDA;
  }
  cudaDeviceProp p;
  e = cudaGetDeviceProperties(&p, dev);
  if (e != cudaSuccess) {
    pg::set_error("pg_device_info: %s", cudaGetErrorString(e));
    return PG_ERR_CUDA;
  }
  if (sm_count) *sm_count = p.multiProcessorCount;
  if (cc_major) *cc_major = p.major;
  if (cc_minor) *cc_minor = p.minor;
  return PG_OK;
}
