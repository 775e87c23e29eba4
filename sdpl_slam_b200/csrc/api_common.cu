// api_common.cu -- error reporting shared by the C-ABI entry points.
#include "common.cuh"

namespace sdpl {
static thread_local std::string g_last_error;
thread_local int g_launches = 0;
void set_last_error(const std::string& s) { g_last_error = s; }
}  // namespace sdpl

extern "C" {
const char* sdpl_strerror(int code) {
  switch (code) {
    case SDPL_OK: return "ok";
    case SDPL_ERR_ARG: return "invalid argument";
    case SDPL_ERR_CUDA: return "CUDA error (no CPU fallback exists)";
    case SDPL_ERR_CAPACITY: return "output capacity too small";
    case SDPL_ERR_OVERFLOW: return "internal device buffer overflow";
    case SDPL_ERR_UNSUPPORTED: return "unsupported configuration";
    default: return "unknown error";
  }
}
const char* sdpl_last_error(void) { return sdpl::g_last_error.c_str(); }
int sdpl_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}
}
