// match.cu -- brute-force 256-bit Hamming matching on sm_100a.
// Metric: cv::line_descriptor::match(P,Q,32), 3rdparty/line_descriptor/src/bitops_custom.hpp:86-99.
// Surface: BinaryDescriptorMatcher::match / knnMatch / radiusMatch (binary_descriptor_matcher.cpp:197-504).
//
// K10: one warp owns QW query descriptors (8 x u32 each, in registers of every lane); the 32 lanes stride
// over the train rows with two 128-bit loads per row, popc the XOR, and keep a private top-2 per query;
// the lanes' top-2 are merged with warp shuffles under the total order (distance, train index), which is
// exactly "strict '<' while scanning ascending train index" (ties -> lowest index).  The train set is split
// across blockIdx.y to fill the 148 SMs; partial top-2 are merged by a second tiny kernel.
// This kernel is popc-issue bound, not HBM bound (Q*T*8 popc vs (Q+T)*32 bytes).
#include "common.cuh"
#include <algorithm>

namespace sdpl {

constexpr int kQW = 4;           // queries per warp
constexpr int kWarps = 8;        // warps per block
constexpr uint32_t kNoDist = 257;

struct Top2 { uint32_t d1, i1, d2, i2; };

__device__ __forceinline__ void top2_insert(Top2& t, uint32_t d, uint32_t i) {
  // (d,i) lexicographic; callers feed ascending i per lane so '<' on d alone would do inside a lane,
  // but the merge needs the full order.
  if (d < t.d1 || (d == t.d1 && i < t.i1)) { t.d2 = t.d1; t.i2 = t.i1; t.d1 = d; t.i1 = i; }
  else if (d < t.d2 || (d == t.d2 && i < t.i2)) { t.d2 = d; t.i2 = i; }
}

__device__ __forceinline__ Top2 top2_warp_merge(Top2 t) {
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    Top2 u;
    u.d1 = __shfl_xor_sync(0xffffffffu, t.d1, o); u.i1 = __shfl_xor_sync(0xffffffffu, t.i1, o);
    u.d2 = __shfl_xor_sync(0xffffffffu, t.d2, o); u.i2 = __shfl_xor_sync(0xffffffffu, t.i2, o);
    top2_insert(t, u.d1, u.i1);
    top2_insert(t, u.d2, u.i2);
  }
  return t;
}

// 256-bit Hamming distance.  popc is the scarce unit (a quarter of the logic rate), so seven of the eight XOR words go
// through a carry-save adder tree first (Harley-Seal: sum = a^b^c, carry = maj(a,b,c), one LOP3 each) and only the ones /
// twos / fours words and the eighth word are counted: 4 popc instead of 8 per pair.
__device__ __forceinline__ void csa(uint32_t a, uint32_t b, uint32_t c, uint32_t& sum, uint32_t& carry) {
  sum = a ^ b ^ c;
  carry = (a & b) | (c & (a | b));
}
__device__ __forceinline__ uint32_t hamming256(const uint32_t (&q)[8], const uint4& a, const uint4& b) {
  const uint32_t x0 = q[0] ^ a.x, x1 = q[1] ^ a.y, x2 = q[2] ^ a.z, x3 = q[3] ^ a.w;
  const uint32_t x4 = q[4] ^ b.x, x5 = q[5] ^ b.y, x6 = q[6] ^ b.z, x7 = q[7] ^ b.w;
  uint32_t s1, c1, s2, c2, ones, c3, twos, fours;
  csa(x0, x1, x2, s1, c1);
  csa(x3, x4, x5, s2, c2);
  csa(s1, s2, x6, ones, c3);
  csa(c1, c2, c3, twos, fours);
  return __popc(ones) + __popc(x7) + 2u * __popc(twos) + 4u * __popc(fours);
}

// partial[(p*max_q + q)*nsplit + s]
template <int MINB>
__global__ void __launch_bounds__(kWarps * 32, MINB) k_match_partial(const uint8_t* __restrict__ q, const int* __restrict__ nq_arr,
                                                               size_t q_stride, const uint8_t* __restrict__ t,
                                                               const int* __restrict__ nt_arr, size_t t_stride, int max_q,
                                                               int max_t, int nsplit, Top2* __restrict__ partial) {
  const int p = blockIdx.z, s = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // device-side counts are clamped to the row capacity of their blocks: an extractor that found more features than its output
  // capacity reports the raw count (the caller sees SDPL_ERR_CAPACITY) but only `capacity` rows exist
  const int nq = min(nq_arr[p], max_q), nt = min(nt_arr[p], max_t);
  const int q0 = (blockIdx.x * kWarps + warp) * kQW;
  if (q0 >= nq) return;
  const uint8_t* qp = q + (size_t)p * q_stride;
  const uint8_t* tp = t + (size_t)p * t_stride;
  uint32_t qw[kQW][8];
#pragma unroll
  for (int k = 0; k < kQW; k++) {
    int qi = min(q0 + k, nq - 1);
    const uint4* src = (const uint4*)(qp + (size_t)qi * 32);
    uint4 a = __ldg(src), b = __ldg(src + 1);
    qw[k][0] = a.x; qw[k][1] = a.y; qw[k][2] = a.z; qw[k][3] = a.w;
    qw[k][4] = b.x; qw[k][5] = b.y; qw[k][6] = b.z; qw[k][7] = b.w;
  }
  // top-2 per query as packed keys (distance << 23 | train row): one unsigned compare orders by distance, then by row, so an update
  // is three min / max instructions instead of two compare-and-move chains, and "strict '<' over ascending rows" is the key order
  // (23 bits of row: match_run_dev rejects problems with more than 8 M train rows)
  uint32_t k1[kQW], k2[kQW];
#pragma unroll
  for (int k = 0; k < kQW; k++) { k1[k] = 0xFFFFFFFFu; k2[k] = 0xFFFFFFFFu; }
  // this split's train range
  const int per = (nt + nsplit - 1) / nsplit;
  const int tb = s * per, te = min(nt, tb + per);
  for (int j = tb + lane; j < te; j += 32) {
    const uint4* src = (const uint4*)(tp + (size_t)j * 32);
    uint4 a = __ldg(src), b = __ldg(src + 1);
#pragma unroll
    for (int k = 0; k < kQW; k++) {
      const uint32_t key = (hamming256(qw[k], a, b) << 23) | (uint32_t)j;
      const uint32_t lo = min(key, k1[k]), hi = max(key, k1[k]);
      k1[k] = lo; k2[k] = min(k2[k], hi);
    }
  }
#pragma unroll
  for (int k = 0; k < kQW; k++) {
    uint32_t a1 = k1[k], a2 = k2[k];
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      const uint32_t b1 = __shfl_xor_sync(0xffffffffu, a1, o), b2 = __shfl_xor_sync(0xffffffffu, a2, o);
      // two smallest of {a1 <= a2, b1 <= b2}: the lanes hold disjoint rows, so equal keys only occur as the 0xFFFFFFFF filler
      const uint32_t lo = min(a1, b1), hi = max(a1, b1);
      a2 = min(hi, min(a2, b2)); a1 = lo;
    }
    if (lane == 0 && q0 + k < nq) {
      Top2 m;
      m.d1 = a1 == 0xFFFFFFFFu ? kNoDist : (a1 >> 23); m.i1 = a1 == 0xFFFFFFFFu ? 0xFFFFFFFFu : (a1 & 0x7FFFFFu);
      m.d2 = a2 == 0xFFFFFFFFu ? kNoDist : (a2 >> 23); m.i2 = a2 == 0xFFFFFFFFu ? 0xFFFFFFFFu : (a2 & 0x7FFFFFu);
      partial[((size_t)p * max_q + q0 + k) * nsplit + s] = m;
    }
  }
}

__global__ void __launch_bounds__(256) k_match_merge(const Top2* __restrict__ partial, const int* __restrict__ nq_arr, int max_q,
                                                     int nsplit, sdpl_dmatch* __restrict__ best, sdpl_dmatch* __restrict__ second) {
  const int p = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= min(nq_arr[p], max_q)) return;
  Top2 t{kNoDist, 0xFFFFFFFFu, kNoDist, 0xFFFFFFFFu};
  const Top2* src = partial + ((size_t)p * max_q + i) * nsplit;
  for (int s = 0; s < nsplit; s++) { Top2 u = src[s]; top2_insert(t, u.d1, u.i1); top2_insert(t, u.d2, u.i2); }
  sdpl_dmatch b, c;
  b.query = i; b.train = t.d1 == kNoDist ? -1 : (int)t.i1; b.img = 0; b.distance = (float)t.d1;
  c.query = i; c.train = t.d2 == kNoDist ? -1 : (int)t.i2; c.img = 0; c.distance = (float)t.d2;
  best[(size_t)p * max_q + i] = b;
  second[(size_t)p * max_q + i] = c;
}

// ratio + max-distance filter; counts accepted matches per problem
__global__ void __launch_bounds__(256) k_match_ratio(const sdpl_dmatch* __restrict__ best, const sdpl_dmatch* __restrict__ second, int nq,
                                                     float ratio, float max_dist, sdpl_dmatch* __restrict__ out, int* __restrict__ n_acc) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  bool ok = false;
  if (i < nq) {
    sdpl_dmatch b = best[i], s = second[i];
    ok = b.train >= 0 && b.distance <= max_dist && b.distance < __fmul_rn(ratio, s.distance);
    if (!ok) b.train = -1;
    out[i] = b;
  }
  unsigned m = __ballot_sync(0xffffffffu, ok);
  if ((threadIdx.x & 31) == 0 && m) atomicAdd(n_acc, __popc(m));
}

// batched ratio + max-distance filter: problem p = blockIdx.y; out may be NULL (count only)
__global__ void __launch_bounds__(256) k_match_ratio_batch(const sdpl_dmatch* __restrict__ best, const sdpl_dmatch* __restrict__ second,
                                                           const int* __restrict__ nq_arr, int max_q, float ratio, float max_dist,
                                                           sdpl_dmatch* __restrict__ out, int* __restrict__ n_acc) {
  const int p = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  bool ok = false;
  if (i < min(nq_arr[p], max_q)) {
    sdpl_dmatch b = best[(size_t)p * max_q + i], s = second[(size_t)p * max_q + i];
    ok = b.train >= 0 && b.distance <= max_dist && b.distance < __fmul_rn(ratio, s.distance);
    if (!ok) b.train = -1;
    if (out) out[(size_t)p * max_q + i] = b;
  }
  unsigned m = __ballot_sync(0xffffffffu, ok);
  if ((threadIdx.x & 31) == 0 && m) atomicAdd(n_acc + p, __popc(m));
}

// radius search: one warp per query; counts all train rows within radius and keeps the k nearest
// (ascending (distance, index)) by k rounds of warp-wide arg-min selection over a per-lane scan.
__global__ void __launch_bounds__(256) k_match_radius(const uint8_t* __restrict__ q, int nq, const uint8_t* __restrict__ t, int nt,
                                                      int radius, int k, int* __restrict__ counts, sdpl_dmatch* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int qi = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (qi >= nq) return;
  uint32_t qw[8];
  {
    const uint4* src = (const uint4*)(q + (size_t)qi * 32);
    uint4 a = __ldg(src), b = __ldg(src + 1);
    qw[0] = a.x; qw[1] = a.y; qw[2] = a.z; qw[3] = a.w; qw[4] = b.x; qw[5] = b.y; qw[6] = b.z; qw[7] = b.w;
  }
  int cnt = 0;
  uint32_t last_d = 0, last_i = 0;   // last emitted (distance, index); next must be strictly greater
  bool first = true;
  for (int r = 0; r <= k; r++) {
    // pass r: find the smallest (d,i) > (last_d,last_i) with d <= radius; pass 0 also counts
    uint32_t bd = kNoDist, bi = 0xFFFFFFFFu;
    for (int j = lane; j < nt; j += 32) {
      const uint4* src = (const uint4*)(t + (size_t)j * 32);
      uint4 a = __ldg(src), b = __ldg(src + 1);
      uint32_t d = __popc(qw[0] ^ a.x) + __popc(qw[1] ^ a.y) + __popc(qw[2] ^ a.z) + __popc(qw[3] ^ a.w) +
                   __popc(qw[4] ^ b.x) + __popc(qw[5] ^ b.y) + __popc(qw[6] ^ b.z) + __popc(qw[7] ^ b.w);
      if ((int)d > radius) continue;
      if (r == 0) cnt++;
      bool after = first || d > last_d || (d == last_d && (uint32_t)j > last_i);
      if (after && (d < bd || (d == bd && (uint32_t)j < bi))) { bd = d; bi = j; }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      uint32_t od = __shfl_xor_sync(0xffffffffu, bd, o), oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (od < bd || (od == bd && oi < bi)) { bd = od; bi = oi; }
    }
    if (r == 0) {
#pragma unroll
      for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
      if (lane == 0) counts[qi] = cnt;
    }
    if (r == k) break;
    if (lane == 0) {
      sdpl_dmatch m;
      m.query = qi; m.img = 0;
      if (bd == kNoDist) { m.train = -1; m.distance = 257.f; } else { m.train = (int)bi; m.distance = (float)bd; }
      out[(size_t)qi * k + r] = m;
    }
    if (bd == kNoDist) {
      // nothing left: fill the remaining slots
      if (lane == 0) for (int rr = r + 1; rr < k; rr++) { sdpl_dmatch m{qi, -1, 0, 257.f}; out[(size_t)qi * k + rr] = m; }
      break;
    }
    last_d = bd; last_i = bi; first = false;
  }
}

}  // namespace sdpl

using namespace sdpl;

struct sdpl_matcher {
  int device;
  cudaStream_t own_stream = nullptr, stream = nullptr;
  DevBuf q, t, best, second, out, partial, scal, counts;
  // train set kept on the device (BinaryDescriptorMatcher::add / train / clear): the concatenation of every added image's
  // descriptors; img_start[i] = first row of image i (the reference's indexesMap, binary_descriptor_matcher.cpp:127-160)
  DevBuf train;
  int train_n = 0;
  std::vector<int> img_start;
  int launches = 0;
  int sm_count = 148;
  StageTimer timer;
};

static int match_run_dev(sdpl_matcher* m, const uint8_t* d_q, const int* d_nq, size_t q_stride, const uint8_t* d_t, const int* d_nt,
                         size_t t_stride, int npairs, int max_q, int max_t, sdpl_dmatch* d_best, sdpl_dmatch* d_second) {
  if (max_t > (1 << 23)) { set_last_error("matcher: more than 2^23 train rows per problem (packed top-2 keys hold 23 bits of row)"); return SDPL_ERR_UNSUPPORTED; }
  // choose the train split so that the grid covers the SMs a few times
  int qblocks = div_up(max_q, kWarps * kQW);
  int nsplit = std::max(1, std::min(div_up(4 * m->sm_count, std::max(1, qblocks * npairs)), div_up(max_t, 64)));
  nsplit = std::min(nsplit, 64);
  int rc = m->partial.reserve(sizeof(Top2) * (size_t)npairs * max_q * nsplit);
  if (rc) return rc;
  m->timer.begin(m->stream);
  {
    const int mbk = getenv("SDPL_MATCH_MINB") ? atoi(getenv("SDPL_MATCH_MINB")) : 4;   // 78 registers at 3 CTAs/SM: 2.80 ms per 512 point problems; 64 at 4: 2.67; 48 at 5: 3.25
    if (mbk >= 5) k_match_partial<5><<<dim3(qblocks, nsplit, npairs), kWarps * 32, 0, m->stream>>>(d_q, d_nq, q_stride, d_t, d_nt, t_stride, max_q, max_t,
                                                                                               nsplit, m->partial.as<Top2>());
    else if (mbk == 4) k_match_partial<4><<<dim3(qblocks, nsplit, npairs), kWarps * 32, 0, m->stream>>>(d_q, d_nq, q_stride, d_t, d_nt, t_stride, max_q,
                                                                                                    max_t, nsplit, m->partial.as<Top2>());
    else k_match_partial<3><<<dim3(qblocks, nsplit, npairs), kWarps * 32, 0, m->stream>>>(d_q, d_nq, q_stride, d_t, d_nt, t_stride, max_q, max_t, nsplit,
                                                                                          m->partial.as<Top2>());
  }
  SDPL_LAUNCH_CHECK();
  m->timer.mark(m->stream, "match_partial");
  k_match_merge<<<dim3(div_up(max_q, 256), npairs), 256, 0, m->stream>>>(m->partial.as<Top2>(), d_nq, max_q, nsplit, d_best, d_second);
  SDPL_LAUNCH_CHECK();
  m->timer.mark(m->stream, "match_merge");
  return SDPL_OK;
}

extern "C" {

int sdpl_matcher_create(sdpl_matcher** out, int device) {
  if (!out) return SDPL_ERR_ARG;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
    set_last_error("sdpl_matcher_create: no such CUDA device (this library has no CPU fallback)");
    return SDPL_ERR_CUDA;
  }
  SDPL_CUDA(cudaSetDevice(device));
  sdpl_matcher* m = new sdpl_matcher;
  m->device = device;
  cudaDeviceGetAttribute(&m->sm_count, cudaDevAttrMultiProcessorCount, device);
  SDPL_CUDA(cudaStreamCreateWithFlags(&m->own_stream, cudaStreamNonBlocking));
  m->stream = m->own_stream;
  *out = m;
  return SDPL_OK;
}

void sdpl_matcher_destroy(sdpl_matcher* m) {
  if (!m) return;
  cudaSetDevice(m->device);
  cudaStreamSynchronize(m->stream);
  for (DevBuf* b : {&m->q, &m->t, &m->best, &m->second, &m->out, &m->partial, &m->scal, &m->counts, &m->train}) b->release();
  m->timer.release();
  if (m->own_stream) cudaStreamDestroy(m->own_stream);
  delete m;
}
int sdpl_matcher_set_stream(sdpl_matcher* m, void* s) { if (!m) return SDPL_ERR_ARG; m->stream = s ? (cudaStream_t)s : m->own_stream; return SDPL_OK; }
int sdpl_matcher_last_launches(const sdpl_matcher* m) { return m ? m->launches : 0; }
int sdpl_matcher_set_profiling(sdpl_matcher* m, int on) { if (!m) return SDPL_ERR_ARG; m->timer.enabled = on != 0; return SDPL_OK; }
int sdpl_matcher_stage_times(sdpl_matcher* m, float* ms, const char** names, int* launches, int cap) {
  if (!m) return 0;
  cudaSetDevice(m->device);
  return m->timer.read(ms, names, launches, cap);
}

int sdpl_match_knn2_batch_dev(sdpl_matcher* m, const uint8_t* d_q, const int* d_nq, size_t q_stride, const uint8_t* d_t, const int* d_nt,
                              size_t t_stride, int npairs, int max_q, int max_t, sdpl_dmatch* d_best, sdpl_dmatch* d_second, int sync) {
  if (!m || !d_q || !d_nq || !d_t || !d_nt || npairs < 1 || max_q < 1 || max_t < 1 || !d_best || !d_second) {
    set_last_error("sdpl_match_knn2_batch_dev: bad argument");
    return SDPL_ERR_ARG;
  }
  SDPL_CUDA(cudaSetDevice(m->device));
  g_launches = 0;
  int rc = match_run_dev(m, d_q, d_nq, q_stride, d_t, d_nt, t_stride, npairs, max_q, max_t, d_best, d_second);
  m->launches = g_launches;
  if (rc) return rc;
  if (sync) SDPL_CUDA(cudaStreamSynchronize(m->stream));
  return SDPL_OK;
}

int sdpl_match_ratio_batch_dev(sdpl_matcher* m, const sdpl_dmatch* d_best, const sdpl_dmatch* d_second, const int* d_nq, int npairs,
                               int max_q, float ratio, int max_dist, sdpl_dmatch* d_out, int* d_n_acc, int sync) {
  if (!m || !d_best || !d_second || !d_nq || npairs < 1 || max_q < 1 || !d_n_acc) {
    set_last_error("sdpl_match_ratio_batch_dev: bad argument");
    return SDPL_ERR_ARG;
  }
  SDPL_CUDA(cudaSetDevice(m->device));
  g_launches = 0;
  SDPL_CUDA(cudaMemsetAsync(d_n_acc, 0, sizeof(int) * npairs, m->stream));
  k_match_ratio_batch<<<dim3(div_up(max_q, 256), npairs), 256, 0, m->stream>>>(d_best, d_second, d_nq, max_q, ratio, (float)max_dist, d_out,
                                                                              d_n_acc);
  SDPL_LAUNCH_CHECK();
  m->launches = g_launches;
  if (sync) SDPL_CUDA(cudaStreamSynchronize(m->stream));
  return SDPL_OK;
}

static int match_host_common(sdpl_matcher* m, const uint8_t* q, int nq, const uint8_t* t, int nt) {
  int rc;
  if ((rc = m->q.reserve((size_t)nq * 32))) return rc;
  if ((rc = m->t.reserve((size_t)std::max(nt, 1) * 32))) return rc;
  if ((rc = m->best.reserve(sizeof(sdpl_dmatch) * nq))) return rc;
  if ((rc = m->second.reserve(sizeof(sdpl_dmatch) * nq))) return rc;
  if ((rc = m->out.reserve(sizeof(sdpl_dmatch) * nq))) return rc;
  if ((rc = m->scal.reserve(sizeof(int) * 4))) return rc;
  SDPL_CUDA(cudaMemcpyAsync(m->q.p, q, (size_t)nq * 32, cudaMemcpyHostToDevice, m->stream));
  if (nt > 0) SDPL_CUDA(cudaMemcpyAsync(m->t.p, t, (size_t)nt * 32, cudaMemcpyHostToDevice, m->stream));
  int scal[4] = {nq, nt, 0, 0};
  SDPL_CUDA(cudaMemcpyAsync(m->scal.p, scal, sizeof(scal), cudaMemcpyHostToDevice, m->stream));
  return match_run_dev(m, m->q.as<uint8_t>(), m->scal.as<int>(), 0, m->t.as<uint8_t>(), m->scal.as<int>() + 1, 0, 1, nq, std::max(nt, 1),
                       m->best.as<sdpl_dmatch>(), m->second.as<sdpl_dmatch>());
}

int sdpl_match_knn2(sdpl_matcher* m, const uint8_t* q, int nq, const uint8_t* t, int nt, sdpl_dmatch* best, sdpl_dmatch* second) {
  if (!m || nq < 0 || nt < 0) { set_last_error("sdpl_match_knn2: bad argument"); return SDPL_ERR_ARG; }
  if (nq == 0) return SDPL_OK;   // the reference matcher prints and returns on empty inputs (binary_descriptor_matcher.cpp:201-205)
  if (!q || (nt > 0 && !t) || !best || !second) { set_last_error("sdpl_match_knn2: null pointer"); return SDPL_ERR_ARG; }
  SDPL_CUDA(cudaSetDevice(m->device));
  g_launches = 0;
  int rc = match_host_common(m, q, nq, t, nt);
  m->launches = g_launches;
  if (rc) return rc;
  SDPL_CUDA(cudaMemcpyAsync(best, m->best.p, sizeof(sdpl_dmatch) * nq, cudaMemcpyDeviceToHost, m->stream));
  SDPL_CUDA(cudaMemcpyAsync(second, m->second.p, sizeof(sdpl_dmatch) * nq, cudaMemcpyDeviceToHost, m->stream));
  SDPL_CUDA(cudaStreamSynchronize(m->stream));
  return SDPL_OK;
}

int sdpl_match_ratio(sdpl_matcher* m, const uint8_t* q, int nq, const uint8_t* t, int nt, float ratio, int max_dist, sdpl_dmatch* out,
                     int* n_acc) {
  if (!m || nq < 0 || nt < 0 || !n_acc) { set_last_error("sdpl_match_ratio: bad argument"); return SDPL_ERR_ARG; }
  *n_acc = 0;
  if (nq == 0) return SDPL_OK;
  if (!q || (nt > 0 && !t) || !out) { set_last_error("sdpl_match_ratio: null pointer"); return SDPL_ERR_ARG; }
  SDPL_CUDA(cudaSetDevice(m->device));
  g_launches = 0;
  int rc = match_host_common(m, q, nq, t, nt);
  if (rc) { m->launches = g_launches; return rc; }
  k_match_ratio<<<div_up(nq, 256), 256, 0, m->stream>>>(m->best.as<sdpl_dmatch>(), m->second.as<sdpl_dmatch>(), nq, ratio, (float)max_dist,
                                                        m->out.as<sdpl_dmatch>(), m->scal.as<int>() + 2);
  SDPL_LAUNCH_CHECK();
  m->launches = g_launches;
  SDPL_CUDA(cudaMemcpyAsync(out, m->out.p, sizeof(sdpl_dmatch) * nq, cudaMemcpyDeviceToHost, m->stream));
  SDPL_CUDA(cudaMemcpyAsync(n_acc, m->scal.as<int>() + 2, sizeof(int), cudaMemcpyDeviceToHost, m->stream));
  SDPL_CUDA(cudaStreamSynchronize(m->stream));
  return SDPL_OK;
}

int sdpl_match_radius(sdpl_matcher* m, const uint8_t* q, int nq, const uint8_t* t, int nt, int radius, int k, int* counts, sdpl_dmatch* out) {
  if (!m || nq < 0 || nt < 0 || k < 0 || radius < 0) { set_last_error("sdpl_match_radius: bad argument"); return SDPL_ERR_ARG; }
  if (nq == 0) return SDPL_OK;
  if (!q || (nt > 0 && !t) || !counts || (k > 0 && !out)) { set_last_error("sdpl_match_radius: null pointer"); return SDPL_ERR_ARG; }
  SDPL_CUDA(cudaSetDevice(m->device));
  int rc;
  if ((rc = m->q.reserve((size_t)nq * 32))) return rc;
  if ((rc = m->t.reserve((size_t)std::max(nt, 1) * 32))) return rc;
  if ((rc = m->counts.reserve(sizeof(int) * nq))) return rc;
  if ((rc = m->out.reserve(sizeof(sdpl_dmatch) * (size_t)nq * std::max(k, 1)))) return rc;
  SDPL_CUDA(cudaMemcpyAsync(m->q.p, q, (size_t)nq * 32, cudaMemcpyHostToDevice, m->stream));
  if (nt > 0) SDPL_CUDA(cudaMemcpyAsync(m->t.p, t, (size_t)nt * 32, cudaMemcpyHostToDevice, m->stream));
  g_launches = 0;
  k_match_radius<<<div_up(nq, 8), 256, 0, m->stream>>>(m->q.as<uint8_t>(), nq, m->t.as<uint8_t>(), nt, radius, k, m->counts.as<int>(),
                                                       m->out.as<sdpl_dmatch>());
  SDPL_LAUNCH_CHECK();
  m->launches = g_launches;
  SDPL_CUDA(cudaMemcpyAsync(counts, m->counts.p, sizeof(int) * nq, cudaMemcpyDeviceToHost, m->stream));
  if (k > 0) SDPL_CUDA(cudaMemcpyAsync(out, m->out.p, sizeof(sdpl_dmatch) * (size_t)nq * k, cudaMemcpyDeviceToHost, m->stream));
  SDPL_CUDA(cudaStreamSynchronize(m->stream));
  return SDPL_OK;
}

// ---- general k, explicit train set: k <= 2 on the top-2 kernels, larger k by k rounds of arg-min selection (k_match_radius with
//      the radius wide open).  out[i*k + j] = j-th nearest of query i, ascending (distance, train index); missing: train = -1 ----
static int match_knn_dev_train(sdpl_matcher* m, const uint8_t* q, int nq, const uint8_t* d_t, int nt, int k, sdpl_dmatch* out) {
  int rc;
  if ((rc = m->q.reserve((size_t)nq * 32))) return rc;
  if ((rc = m->out.reserve(sizeof(sdpl_dmatch) * (size_t)nq * std::max(k, 2)))) return rc;
  SDPL_CUDA(cudaMemcpyAsync(m->q.p, q, (size_t)nq * 32, cudaMemcpyHostToDevice, m->stream));
  if (k <= 2) {
    if ((rc = m->best.reserve(sizeof(sdpl_dmatch) * nq))) return rc;
    if ((rc = m->second.reserve(sizeof(sdpl_dmatch) * nq))) return rc;
    if ((rc = m->scal.reserve(sizeof(int) * 4))) return rc;
    int scal[4] = {nq, nt, 0, 0};
    SDPL_CUDA(cudaMemcpyAsync(m->scal.p, scal, sizeof(scal), cudaMemcpyHostToDevice, m->stream));
    // dummy non-null train pointer for an empty set (never dereferenced: nt == 0)
    if ((rc = match_run_dev(m, m->q.as<uint8_t>(), m->scal.as<int>(), 0, d_t ? d_t : m->q.as<uint8_t>(), m->scal.as<int>() + 1, 0, 1, nq,
                            std::max(nt, 1), m->best.as<sdpl_dmatch>(), m->second.as<sdpl_dmatch>()))) return rc;
    std::vector<sdpl_dmatch> b(nq), c(k == 2 ? nq : 0);
    SDPL_CUDA(cudaMemcpyAsync(b.data(), m->best.p, sizeof(sdpl_dmatch) * nq, cudaMemcpyDeviceToHost, m->stream));
    if (k == 2) SDPL_CUDA(cudaMemcpyAsync(c.data(), m->second.p, sizeof(sdpl_dmatch) * nq, cudaMemcpyDeviceToHost, m->stream));
    SDPL_CUDA(cudaStreamSynchronize(m->stream));
    for (int i = 0; i < nq; i++) { out[(size_t)i * k] = b[i]; if (k == 2) out[(size_t)i * k + 1] = c[i]; }
    return SDPL_OK;
  }
  if ((rc = m->counts.reserve(sizeof(int) * nq))) return rc;
  k_match_radius<<<div_up(nq, 8), 256, 0, m->stream>>>(m->q.as<uint8_t>(), nq, d_t, nt, 256, k, m->counts.as<int>(), m->out.as<sdpl_dmatch>());
  SDPL_LAUNCH_CHECK();
  SDPL_CUDA(cudaMemcpyAsync(out, m->out.p, sizeof(sdpl_dmatch) * (size_t)nq * k, cudaMemcpyDeviceToHost, m->stream));
  SDPL_CUDA(cudaStreamSynchronize(m->stream));
  return SDPL_OK;
}

int sdpl_match_knn(sdpl_matcher* m, const uint8_t* q, int nq, const uint8_t* t, int nt, int k, sdpl_dmatch* out) {
  if (!m || nq < 0 || nt < 0 || k < 1) { set_last_error("sdpl_match_knn: bad argument"); return SDPL_ERR_ARG; }
  if (nq == 0) return SDPL_OK;
  if (!q || (nt > 0 && !t) || !out) { set_last_error("sdpl_match_knn: null pointer"); return SDPL_ERR_ARG; }
  SDPL_CUDA(cudaSetDevice(m->device));
  int rc;
  if ((rc = m->t.reserve((size_t)std::max(nt, 1) * 32))) return rc;
  if (nt > 0) SDPL_CUDA(cudaMemcpyAsync(m->t.p, t, (size_t)nt * 32, cudaMemcpyHostToDevice, m->stream));
  g_launches = 0;
  rc = match_knn_dev_train(m, q, nq, m->t.as<uint8_t>(), nt, k, out);
  m->launches = g_launches;
  return rc;
}

// ---- the stored train set ----
static int matcher_append(sdpl_matcher* m, const uint8_t* desc, int n, cudaMemcpyKind kind) {
  SDPL_CUDA(cudaSetDevice(m->device));
  const size_t need = (size_t)(m->train_n + n) * 32;
  if (need > m->train.bytes) {
    // grow geometrically; the rows already stored move device to device
    DevBuf bigger;
    int rc = bigger.reserve(std::max(need, std::max<size_t>(2 * m->train.bytes, (size_t)4096 * 32)));
    if (rc) return rc;
    if (m->train_n) SDPL_CUDA(cudaMemcpyAsync(bigger.p, m->train.p, (size_t)m->train_n * 32, cudaMemcpyDeviceToDevice, m->stream));
    SDPL_CUDA(cudaStreamSynchronize(m->stream));
    m->train.release();
    m->train = bigger;
  }
  if (n > 0) SDPL_CUDA(cudaMemcpyAsync((uint8_t*)m->train.p + (size_t)m->train_n * 32, desc, (size_t)n * 32, kind, m->stream));
  SDPL_CUDA(cudaStreamSynchronize(m->stream));       // the caller may free `desc` on return
  m->img_start.push_back(m->train_n);
  m->train_n += n;
  return SDPL_OK;
}
int sdpl_matcher_add(sdpl_matcher* m, const uint8_t* desc, int n) {
  if (!m || n < 0 || (n > 0 && !desc)) { set_last_error("sdpl_matcher_add: bad argument"); return SDPL_ERR_ARG; }
  return matcher_append(m, desc, n, cudaMemcpyHostToDevice);
}
int sdpl_matcher_add_dev(sdpl_matcher* m, const uint8_t* d_desc, int n) {
  if (!m || n < 0 || (n > 0 && !d_desc)) { set_last_error("sdpl_matcher_add_dev: bad argument"); return SDPL_ERR_ARG; }
  return matcher_append(m, d_desc, n, cudaMemcpyDeviceToDevice);
}
int sdpl_matcher_train(sdpl_matcher* m) { return m ? SDPL_OK : SDPL_ERR_ARG; }   // brute force: nothing to index
int sdpl_matcher_clear(sdpl_matcher* m) {
  if (!m) return SDPL_ERR_ARG;
  m->train_n = 0; m->img_start.clear();
  return SDPL_OK;
}
int sdpl_matcher_train_size(const sdpl_matcher* m, int* n_desc, int* n_imgs) {
  if (!m) return SDPL_ERR_ARG;
  if (n_desc) *n_desc = m->train_n;
  if (n_imgs) *n_imgs = (int)m->img_start.size();
  return SDPL_OK;
}
// imgIdx of a row of the concatenated set: the image whose first row is the last one <= train (the reference's
// indexesMap.upper_bound(idx) - 1, binary_descriptor_matcher.cpp:161-163); trainIdx stays the row of the concatenation, as there
static void matcher_set_img(const sdpl_matcher* m, sdpl_dmatch* out, size_t n) {
  for (size_t i = 0; i < n; i++) {
    if (out[i].train < 0) continue;
    const auto it = std::upper_bound(m->img_start.begin(), m->img_start.end(), out[i].train);
    out[i].img = (int)(it - m->img_start.begin()) - 1;
  }
}
int sdpl_matcher_knn(sdpl_matcher* m, const uint8_t* q, int nq, int k, sdpl_dmatch* out) {
  if (!m || nq < 0 || k < 1) { set_last_error("sdpl_matcher_knn: bad argument"); return SDPL_ERR_ARG; }
  if (nq == 0) return SDPL_OK;
  if (!q || !out) { set_last_error("sdpl_matcher_knn: null pointer"); return SDPL_ERR_ARG; }
  SDPL_CUDA(cudaSetDevice(m->device));
  g_launches = 0;
  int rc = match_knn_dev_train(m, q, nq, m->train.as<uint8_t>(), m->train_n, k, out);
  m->launches = g_launches;
  if (rc) return rc;
  matcher_set_img(m, out, (size_t)nq * k);
  return SDPL_OK;
}
int sdpl_matcher_radius(sdpl_matcher* m, const uint8_t* q, int nq, int radius, int k, int* counts, sdpl_dmatch* out) {
  if (!m || nq < 0 || k < 0 || radius < 0) { set_last_error("sdpl_matcher_radius: bad argument"); return SDPL_ERR_ARG; }
  if (nq == 0) return SDPL_OK;
  if (!q || !counts || (k > 0 && !out)) { set_last_error("sdpl_matcher_radius: null pointer"); return SDPL_ERR_ARG; }
  SDPL_CUDA(cudaSetDevice(m->device));
  int rc;
  if ((rc = m->q.reserve((size_t)nq * 32))) return rc;
  if ((rc = m->counts.reserve(sizeof(int) * nq))) return rc;
  if ((rc = m->out.reserve(sizeof(sdpl_dmatch) * (size_t)nq * std::max(k, 1)))) return rc;
  SDPL_CUDA(cudaMemcpyAsync(m->q.p, q, (size_t)nq * 32, cudaMemcpyHostToDevice, m->stream));
  g_launches = 0;
  k_match_radius<<<div_up(nq, 8), 256, 0, m->stream>>>(m->q.as<uint8_t>(), nq, m->train.as<uint8_t>(), m->train_n, radius, k, m->counts.as<int>(),
                                                       m->out.as<sdpl_dmatch>());
  SDPL_LAUNCH_CHECK();
  m->launches = g_launches;
  SDPL_CUDA(cudaMemcpyAsync(counts, m->counts.p, sizeof(int) * nq, cudaMemcpyDeviceToHost, m->stream));
  if (k > 0) SDPL_CUDA(cudaMemcpyAsync(out, m->out.p, sizeof(sdpl_dmatch) * (size_t)nq * k, cudaMemcpyDeviceToHost, m->stream));
  SDPL_CUDA(cudaStreamSynchronize(m->stream));
  if (k > 0) matcher_set_img(m, out, (size_t)nq * k);
  return SDPL_OK;
}
// the stored set for the batched device entry points (sdpl_match_knn2_batch_dev with t_stride = 0): DEVICE pointer + row count
int sdpl_matcher_train_dev(const sdpl_matcher* m, const uint8_t** d_train, int* n_desc) {
  if (!m || !d_train) return SDPL_ERR_ARG;
  *d_train = m->train.as<uint8_t>();
  if (n_desc) *n_desc = m->train_n;
  return SDPL_OK;
}

}  // extern "C"
