// ABI plumbing: version, thread-local error string, launch status.
#include "rdm_common.cuh"

namespace rdm {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int launch_status(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) return 0;
  set_error("%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
  return (int)e;
}

}  // namespace rdm

extern "C" int rdm_abi_version(void) { return RDM_ABI_VERSION; }
extern "C" const char* rdm_last_error(void) { return rdm::g_err; }
