// api_common.cu -- error reporting shared by the C-ABI entry points.
#include "common.cuh"

namespace sdpl {
static thread_local std::string g_last_error;
thread_local int g_launches = 0;
void set_last_error(const std::string& s) { g_last_error = s; }
}  // namespace sdpl

namespace sdpl {
// 64-bit digest of the valid rows of one per-frame array: every row is hashed with its index (FNV-1a over its 32-bit words,
// then a finaliser) and the row hashes are added up, so the digest does not depend on which thread gets there first.
__global__ void __launch_bounds__(128) k_rows_digest(const uint32_t* __restrict__ rows, int row_words, size_t frame_words, const int* __restrict__ n,
                                                     int max_rows, unsigned long long salt, unsigned long long* __restrict__ digest) {
  const int f = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
  const int cnt = min(n[f], max_rows);
  unsigned long long h = 0;
  if (i < cnt) {
    const uint32_t* r = rows + (size_t)f * frame_words + (size_t)i * row_words;
    h = 0xcbf29ce484222325ull ^ salt ^ ((unsigned long long)(i + 1) * 0x9e3779b97f4a7c15ull);
    for (int k = 0; k < row_words; k++) { h ^= r[k]; h *= 0x100000001b3ull; }
    h ^= h >> 33; h *= 0xff51afd7ed558ccdull; h ^= h >> 33;
  }
#pragma unroll
  for (int s = 16; s; s >>= 1) h += __shfl_xor_sync(0xffffffffu, h, s);
  if ((threadIdx.x & 31) == 0 && h) atomicAdd(digest + f, h);
}
}  // namespace sdpl

extern "C" {
// Adds the digest of rows [0, min(d_n[f], max_rows)) of frame f's block (row_bytes per row, a multiple of 4; blocks frame_stride
// bytes apart) to d_digest[f], on `stream`.  bench.py / the sharding tests compare per-frame results across ranks with it.
int sdpl_rows_digest_dev(const void* d_rows, int row_bytes, size_t frame_stride, const int* d_n, int nframes, int max_rows,
                         unsigned long long salt, unsigned long long* d_digest, void* stream) {
  if (!d_rows || row_bytes < 4 || (row_bytes & 3) || (frame_stride & 3) || !d_n || nframes < 1 || max_rows < 1 || !d_digest) {
    sdpl::set_last_error("sdpl_rows_digest_dev: bad argument");
    return SDPL_ERR_ARG;
  }
  sdpl::k_rows_digest<<<dim3((max_rows + 127) / 128, nframes), 128, 0, (cudaStream_t)stream>>>((const uint32_t*)d_rows, row_bytes / 4, frame_stride / 4, d_n,
                                                                                                max_rows, salt, d_digest);
  SDPL_LAUNCH_CHECK();
  return SDPL_OK;
}

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
