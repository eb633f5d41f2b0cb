// Error reporting + device queries for libtsw_sm100.so
#include <cstdlib>
#include <stdarg.h>

#include "common.cuh"

namespace tsw {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
int g_sm_reserve = 0;   // SMs the persistent kernels leave free (for NCCL's kernels in data-parallel runs)
int g_fmha_dynamic = getenv("TSW_FMHA_DYNAMIC") != nullptr && atoi(getenv("TSW_FMHA_DYNAMIC")) != 0 ? 1 : 0;
int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev] - g_sm_reserve > 8 ? cached[dev] - g_sm_reserve : cached[dev];
}
}  // namespace tsw

extern "C" int tsw_set_sm_reserve(int n_sm) {
  TSW_CHECK_ARG(n_sm >= 0 && n_sm % 2 == 0 && n_sm <= 64, "set_sm_reserve: expected an even count in [0, 64]");
  tsw::g_sm_reserve = n_sm;
  return TSW_OK;
}
extern "C" int tsw_set_fmha_work_list(int dynamic) {
  TSW_CHECK_ARG(dynamic == 0 || dynamic == 1, "set_fmha_work_list: expected 0 or 1");
  tsw::g_fmha_dynamic = dynamic;
  return TSW_OK;
}
extern "C" int tsw_abi_version(void) { return TSW_ABI_VERSION; }
extern "C" const char* tsw_last_error(void) { return tsw::g_err; }
extern "C" int tsw_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0, maj = 0, min = 0;
  TSW_CUDA(cudaGetDevice(&dev));
  TSW_CUDA(cudaDeviceGetAttribute(&maj, cudaDevAttrComputeCapabilityMajor, dev));
  TSW_CUDA(cudaDeviceGetAttribute(&min, cudaDevAttrComputeCapabilityMinor, dev));
  if (sm_count) *sm_count = tsw::sm_count();
  if (cc_major) *cc_major = maj;
  if (cc_minor) *cc_minor = min;
  if (maj != 10) { tsw::set_error("device cc %d.%d is not sm_100 (B200)", maj, min); return TSW_E_UNSUPPORTED; }
  return TSW_OK;
}
