// common.cuh -- shared host/device helpers for the sm_100a front-end kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>
#include "../../include/sdpl_frontend.h"
#include "../../include/sdpl_trig.h"   // strict-IEEE double sin / cos, the same source the CPU oracle compiles

namespace sdpl {

void set_last_error(const std::string& s);
extern thread_local int g_launches;  // kernels launched by the current API call

#define SDPL_CUDA(call)                                                                              \
  do {                                                                                               \
    cudaError_t e__ = (call);                                                                        \
    if (e__ != cudaSuccess) {                                                                        \
      sdpl::set_last_error(std::string(#call) + ": " + cudaGetErrorString(e__) + " @" + __FILE__ +   \
                           ":" + std::to_string(__LINE__));                                          \
      return SDPL_ERR_CUDA;                                                                          \
    }                                                                                                \
  } while (0)

#define SDPL_LAUNCH_CHECK()                                                                          \
  do {                                                                                               \
    sdpl::g_launches++;                                                                              \
    cudaError_t e__ = cudaGetLastError();                                                            \
    if (e__ != cudaSuccess) {                                                                        \
      sdpl::set_last_error(std::string("kernel launch: ") + cudaGetErrorString(e__) + " @" +         \
                           __FILE__ + ":" + std::to_string(__LINE__));                               \
      return SDPL_ERR_CUDA;                                                                          \
    }                                                                                                \
  } while (0)

// Simple owning device buffer (grow-only).
struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  int reserve(size_t n) {
    if (n <= bytes) return SDPL_OK;
    if (p) cudaFree(p);
    p = nullptr; bytes = 0;
    cudaError_t e = cudaMalloc(&p, n);
    if (e != cudaSuccess) { set_last_error(std::string("cudaMalloc: ") + cudaGetErrorString(e)); return SDPL_ERR_CUDA; }
    bytes = n;
    return SDPL_OK;
  }
  void release() { if (p) cudaFree(p); p = nullptr; bytes = 0; }
  template <class T> T* as() const { return (T*)p; }
};

// Per-stage device timing with CUDA events recorded on the handle's stream (bench.py's live roofline numbers).
// mark(name) closes the stage `name` that started at the previous mark; the first mark of a call is "begin".
struct StageTimer {
  std::vector<cudaEvent_t> ev;
  std::vector<const char*> names;
  std::vector<int> launches;      // kernels launched in the stage
  int n = 0;
  bool enabled = false;
  void begin(cudaStream_t st) { n = 0; mark(st, "begin"); }
  void mark(cudaStream_t st, const char* name) {
    if (!enabled) return;
    if (n == (int)ev.size()) {
      cudaEvent_t e;
      if (cudaEventCreate(&e) != cudaSuccess) { enabled = false; return; }
      ev.push_back(e); names.push_back(name); launches.push_back(0);
    }
    names[n] = name; launches[n] = g_launches;
    cudaEventRecord(ev[n], st);
    n++;
  }
  // stage i (0-based) = interval between mark i and mark i+1; returns the number of stages of the last call
  int read(float* ms, const char** nm, int* nlaunch, int cap) {
    if (!enabled || n < 2) return 0;
    cudaEventSynchronize(ev[n - 1]);
    int k = 0;
    for (int i = 0; i + 1 < n && k < cap; i++, k++) {
      float t = 0.f;
      cudaEventElapsedTime(&t, ev[i], ev[i + 1]);
      if (ms) ms[k] = t;
      if (nm) nm[k] = names[i + 1];
      if (nlaunch) nlaunch[k] = launches[i + 1] - launches[i];
    }
    return k;
  }
  void release() { for (auto e : ev) cudaEventDestroy(e); ev.clear(); names.clear(); launches.clear(); n = 0; }
};

// ---- host helpers (OpenCV rounding conventions) ----
static inline int h_cv_round(float v) { return (int)lrintf(v); }
static inline int h_cv_round(double v) { return (int)lrint(v); }
static inline int h_cv_floor(float v) { int i = (int)v; return i - (i > v); }
static inline int h_cv_floor(double v) { int i = (int)v; return i - (i > v); }

// ---- device helpers ----
__device__ __forceinline__ int reflect101(int p, int len) {
  if (len == 1) return 0;
  while (p < 0 || p >= len) p = p < 0 ? -p : 2 * (len - 1) - p;
  return p;
}
__device__ __forceinline__ int cv_round_f(float v) { return __float2int_rn(v); }    // round-half-even
__device__ __forceinline__ int cv_round_d(double v) { return __double2int_rn(v); }

// cv::fastAtan2 in strict, non-contracted fp32 (degrees, [0,360)).  Constants are float(c_k)*float(180/pi)
// products evaluated in fp32 exactly as OpenCV's static initialisers do.
__device__ __forceinline__ float fast_atan2_deg(float y, float x) {
  const float scale = 57.295779513082320876798f;  // (float)(180/CV_PI)
  const float p1 = __fmul_rn(0.9997878412794807f, scale), p3 = __fmul_rn(-0.3258083974640975f, scale),
              p5 = __fmul_rn(0.1555786518463281f, scale), p7 = __fmul_rn(-0.04432655554792128f, scale);
  const float eps = 2.2204460492503131e-16f;  // (float)DBL_EPSILON
  float ax = fabsf(x), ay = fabsf(y), a, c, c2;
  if (ax >= ay) {
    c = __fdiv_rn(ay, __fadd_rn(ax, eps));
    c2 = __fmul_rn(c, c);
    a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
  } else {
    c = __fdiv_rn(ax, __fadd_rn(ay, eps));
    c2 = __fmul_rn(c, c);
    a = __fsub_rn(90.f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c));
  }
  if (x < 0) a = __fsub_rn(180.f, a);
  if (y < 0) a = __fsub_rn(360.f, a);
  return a;
}

static inline int div_up(int a, int b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t a, size_t b) { return (a + b - 1) / b * b; }

}  // namespace sdpl
