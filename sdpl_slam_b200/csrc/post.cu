// post.cu -- the per-frame post-processing Frame::Frame runs on the extractor outputs, on the device:
//   semi-dense object sampling            src/Frame.cc:769-809   (P1: a stream compaction over the mask / depth / flow planes)
//   line filters (depth step, mask)       src/Frame.cc:349-389   (P2)
//   static point correspondences + depth  src/Frame.cc:482-512, 728-745   (P3)
//   line correspondences + depth          src/Frame.cc:513-604, 746-763   (P4)
//   AssignFeaturesToGrid / PosInGrid      src/Frame.cc:910-925, 1023-1035 (P5)
// Inputs are the reference's per-frame planes as they sit in cv::Mat: maskSEM int32 [h][w], imDepth float [h][w], imFlow float
// [h][w][2]; batched over frames (blockIdx carries the frame).  Every output list keeps the order of the sequential loops
// (push_back order), so results are bit-identical to the oracle (oracle/post_oracle.cpp).  With the key points / key lines left on
// the device by the extractors (sdpl_*_extract_batch_dev) nothing of a frame has to visit the host before tracking needs it.
// All stages are HBM-bound integer / float gathers; no tensor cores.
#include "common.cuh"
#include <algorithm>

namespace sdpl {

struct PostPlanes {
  const int32_t* mask; const float* depth; const float* flow;   // frame 0; frames plane elements apart
  int w, h; size_t plane;
  __device__ __forceinline__ int m(int f, int y, int x) const { return __ldg(mask + (size_t)f * plane + (size_t)y * w + x); }
  __device__ __forceinline__ float d(int f, int y, int x) const { return __ldg(depth + (size_t)f * plane + (size_t)y * w + x); }
  __device__ __forceinline__ float2 fl(int f, int y, int x) const {
    return __ldg(reinterpret_cast<const float2*>(flow) + (size_t)f * plane + (size_t)y * w + x);
  }
};

__device__ __forceinline__ sdpl_keypoint make_kp(float x, float y, int octave) {
  sdpl_keypoint k; k.x = x; k.y = y; k.size = 0.f; k.angle = 0.f; k.response = 0.f; k.octave = octave; k.class_id = -1;
  return k;
}

// ------------------------------------------------------------------------------------------------
// P1: semi-dense features on objects.  Pass A (all samples in parallel): a warp takes 32 consecutive samples of a sampled row,
//     evaluates the reference's conditions and stores the ballot as one bitmap word.  Pass B (one CTA per frame): exclusive scan
//     of the words' popcounts.  Pass C (all words in parallel): every set bit writes its records at scan + rank -- row-major
//     order, as the nested loops of the reference push them.  The planes are read once in pass A (every step-th row; the
//     kept samples once more in pass C).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool obj_sample_keep(const PostPlanes& P, int f, int i, int j, float th_obj, int& label, float& depth, float2& fl) {
  label = P.m(f, i, j);
  if (label == 0) return false;
  depth = P.d(f, i, j);
  if (!(depth < th_obj && depth > 0.f)) return false;
  fl = P.fl(f, i, j);
  const float xj = __fadd_rn((float)j, fl.x), yi = __fadd_rn((float)i, fl.y);
  return xj < (float)P.w && xj > 0.f && yi < (float)P.h && yi > 0.f;
}

__global__ void __launch_bounds__(256) k_obj_flags(PostPlanes P, int step, int cols_s, int words_per_row, int rows_s, float th_obj,
                                                   uint32_t* __restrict__ bitmap) {
  const int f = blockIdx.y, lane = threadIdx.x & 31;
  const int wq = blockIdx.x * 8 + (threadIdx.x >> 5);              // word index inside the frame
  if (wq >= words_per_row * rows_s) return;
  const int rs = wq / words_per_row, k = wq - rs * words_per_row;
  const int cs = k * 32 + lane;
  bool keep = false;
  if (cs < cols_s) {
    int label; float depth; float2 fl;
    keep = obj_sample_keep(P, f, rs * step, cs * step, th_obj, label, depth, fl);
  }
  const uint32_t m = __ballot_sync(0xffffffffu, keep);
  if (lane == 0) bitmap[(size_t)f * words_per_row * rows_s + wq] = m;
}

// exclusive scan of popc(bitmap word) per frame: one CTA of 1024 threads per frame
__global__ void __launch_bounds__(1024) k_obj_scan(const uint32_t* __restrict__ bitmap, int nwords, uint32_t* __restrict__ offs, int* __restrict__ n_out) {
  __shared__ uint32_t wsum[32];
  __shared__ uint32_t carry;
  const int f = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t* bm = bitmap + (size_t)f * nwords;
  uint32_t* of = offs + (size_t)f * nwords;
  if (tid == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < nwords; base += 1024) {
    const int i = base + tid;
    const uint32_t c = i < nwords ? (uint32_t)__popc(bm[i]) : 0u;
    uint32_t inc = c;
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, s); if (lane >= s) inc += t; }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      uint32_t v = wsum[lane], iv = v;
#pragma unroll
      for (int s = 1; s < 32; s <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, iv, s); if (lane >= s) iv += t; }
      wsum[lane] = iv - v;
    }
    __syncthreads();
    const uint32_t excl = carry + wsum[warp] + inc - c;
    if (i < nwords) of[i] = excl;
    __syncthreads();
    if (tid == 1023) carry = excl + c;
    __syncthreads();
  }
  if (tid == 0) n_out[f] = (int)carry;
}

__global__ void __launch_bounds__(256) k_obj_emit(PostPlanes P, int step, int cols_s, int words_per_row, int rows_s, const uint32_t* __restrict__ bitmap,
                                                  const uint32_t* __restrict__ offs, sdpl_keypoint* __restrict__ keys,
                                                  sdpl_keypoint* __restrict__ corres, float2* __restrict__ flow_next, float* __restrict__ depth_out,
                                                  int32_t* __restrict__ label_out, int capacity) {
  const int f = blockIdx.y, lane = threadIdx.x & 31;
  const int nwords = words_per_row * rows_s;
  const int wq = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (wq >= nwords) return;
  const uint32_t m = bitmap[(size_t)f * nwords + wq];
  if (!((m >> lane) & 1u)) return;
  const int idx = (int)offs[(size_t)f * nwords + wq] + __popc(m & ((1u << lane) - 1u));
  if (idx >= capacity) return;
  const int rs = wq / words_per_row, k = wq - rs * words_per_row;
  const int i = rs * step, j = (k * 32 + lane) * step;
  const int label = P.m(f, i, j);
  const float depth = P.d(f, i, j);
  const float2 fl = P.fl(f, i, j);
  const size_t o = (size_t)f * capacity + idx;
  flow_next[o] = fl;
  corres[o] = make_kp(__fadd_rn((float)j, fl.x), __fadd_rn((float)i, fl.y), -1);
  keys[o] = make_kp((float)j, (float)i, -1);
  depth_out[o] = depth;
  label_out[o] = label;
}

// ------------------------------------------------------------------------------------------------
// Ordered compaction inside one CTA of 256 threads: returns the output slot of a kept item (base + rank among the kept items of
// lower thread index) and advances base by the number of items kept in this round.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int block_slot(bool keep, int& base, int* wsum /* [9] shared */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t m = __ballot_sync(0xffffffffu, keep);
  __syncthreads();                                    // wsum of the previous round has been read
  if (lane == 0) wsum[warp] = __popc(m);
  __syncthreads();
  int before = 0, total = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) { const int c = wsum[k]; if (k < warp) before += c; total += c; }
  const int slot = base + before + __popc(m & ((1u << lane) - 1u));
  base += total;
  return slot;
}

// P2: the two erase loops on mvKeys_Line (Frame.cc:349-389)
__global__ void __launch_bounds__(256) k_filter_lines(PostPlanes P, const sdpl_keyline* __restrict__ kls, const int* __restrict__ n_in, int capacity,
                                                      sdpl_keyline* __restrict__ out, int32_t* __restrict__ keep_idx, int* __restrict__ n_out) {
  __shared__ int wsum[9];
  const int f = blockIdx.x;
  const int n = min(n_in[f], capacity);
  int base = 0;
  for (int i0 = 0; i0 < n; i0 += 256) {
    const int i = i0 + threadIdx.x;
    bool keep = false;
    sdpl_keyline l;
    if (i < n) {
      l = kls[(size_t)f * capacity + i];
      const int x1 = (int)l.sx, y1 = (int)l.sy, x2 = (int)l.ex, y2 = (int)l.ey;
      const int xm = (x1 + x2) / 2, ym = (y1 + y2) / 2;
      const float ds = P.d(f, y1, x1), dm = P.d(f, ym, xm), de = P.d(f, y2, x2);
      const float expected = __fdiv_rn(__fadd_rn(ds, de), 2.f);
      const double dx = (double)(x2 - x1), dy = (double)(y2 - y1);
      const float len = (float)sqrt(dx * dx + dy * dy);            // pow(int, 2) is an exact product in double
      const float thr = __fmul_rn(10.0f, __fdiv_rn(len, 1000.f));
      keep = !(fabsf(__fsub_rn(dm, expected)) > thr) && P.m(f, y1, x1) == P.m(f, y2, x2);
    }
    const int slot = block_slot(keep, base, wsum);
    if (keep) { out[(size_t)f * capacity + slot] = l; keep_idx[(size_t)f * capacity + slot] = i; }
  }
  if (threadIdx.x == 0) n_out[f] = base;
}

// P3: static point correspondences (Frame.cc:482-512) and their depths (:728-745)
__global__ void __launch_bounds__(256) k_point_corres(PostPlanes P, const sdpl_keypoint* __restrict__ kps, const int* __restrict__ n_in, int capacity,
                                                      float th_depth, sdpl_keypoint* __restrict__ stat, sdpl_keypoint* __restrict__ corres,
                                                      float2* __restrict__ flow_next, float* __restrict__ stat_depth, int32_t* __restrict__ src_idx,
                                                      int* __restrict__ n_out) {
  __shared__ int wsum[9];
  const int f = blockIdx.x;
  const int n = min(n_in[f], capacity);
  int base = 0;
  for (int i0 = 0; i0 < n; i0 += 256) {
    const int i = i0 + threadIdx.x;
    bool keep = false;
    sdpl_keypoint kp; float2 fl = make_float2(0.f, 0.f); float d = 0.f;
    if (i < n) {
      kp = kps[(size_t)f * capacity + i];
      const int x = (int)kp.x, y = (int)kp.y;
      if (P.m(f, y, x) == 0) {
        d = P.d(f, y, x);
        if (!(d > th_depth || d <= 0.f)) {
          fl = P.fl(f, y, x);
          keep = fl.x != 0.f && fl.y != 0.f && __fadd_rn(kp.x, fl.x) < (float)P.w && __fadd_rn(kp.y, fl.y) < (float)P.h && kp.x < (float)P.w &&
                 kp.y < (float)P.h;
        }
      }
    }
    const int slot = block_slot(keep, base, wsum);
    if (keep) {
      const size_t o = (size_t)f * capacity + slot;
      stat[o] = kp;
      corres[o] = make_kp(__fadd_rn(kp.x, fl.x), __fadd_rn(kp.y, fl.y), kp.octave);
      flow_next[o] = fl;
      stat_depth[o] = d > 0.f ? d : -1.f;
      src_idx[o] = i;
    }
  }
  if (threadIdx.x == 0) n_out[f] = base;
}

// P4: line correspondences (Frame.cc:513-604) and their depths (:746-763); object lines go to a list of their own
__global__ void __launch_bounds__(256) k_line_corres(PostPlanes P, const sdpl_keyline* __restrict__ kls, const int* __restrict__ n_in, int capacity,
                                                     float th_depth, sdpl_keyline* __restrict__ obj, int* __restrict__ n_obj,
                                                     sdpl_keyline* __restrict__ stat, sdpl_keyline* __restrict__ corres, float4* __restrict__ flow_next,
                                                     double* __restrict__ inf_line, float2* __restrict__ stat_depth, int32_t* __restrict__ src_idx,
                                                     int* __restrict__ n_out) {
  __shared__ int wsum[9];
  const int f = blockIdx.x;
  const int n = min(n_in[f], capacity);
  int base = 0, obase = 0;
  for (int i0 = 0; i0 < n; i0 += 256) {
    const int i = i0 + threadIdx.x;
    bool keep = false, is_obj = false;
    sdpl_keyline l; float4 fw = make_float4(0.f, 0.f, 0.f, 0.f);
    int sx = 0, sy = 0, ex = 0, ey = 0;
    if (i < n) {
      l = kls[(size_t)f * capacity + i];
      sx = (int)l.sx; sy = (int)l.sy; ex = (int)l.ex; ey = (int)l.ey;
      const int ms = P.m(f, sy, sx), me = P.m(f, ey, ex);
      if (ms != 0 && me != 0) {
        is_obj = ms == me;
      } else if (ms == 0 && me == 0 && !(sx == ex && sy == ey)) {
        const float ds = P.d(f, sy, sx), de = P.d(f, ey, ex);
        if (!(ds > th_depth || ds <= 0.f || de > th_depth || de <= 0.f)) {
          const float2 a = P.fl(f, sy, sx), b = P.fl(f, ey, ex);
          fw = make_float4(a.x, a.y, b.x, b.y);
          const float csx = __fadd_rn((float)sx, a.x), csy = __fadd_rn((float)sy, a.y), cex = __fadd_rn((float)ex, b.x), cey = __fadd_rn((float)ey, b.y);
          keep = a.x != 0.f && a.y != 0.f && b.x != 0.f && b.y != 0.f && csx < (float)P.w && csy < (float)P.h && cex < (float)P.w && cey < (float)P.h &&
                 csx > 0.f && csy > 0.f && cex > 0.f && cey > 0.f;
        }
      }
    }
    const int oslot = block_slot(is_obj, obase, wsum);
    if (is_obj) obj[(size_t)f * capacity + oslot] = l;
    const int slot = block_slot(keep, base, wsum);
    if (keep) {
      const size_t o = (size_t)f * capacity + slot;
      stat[o] = l;
      sdpl_keyline c;
      c.sx = __fadd_rn((float)sx, fw.x); c.sy = __fadd_rn((float)sy, fw.y); c.ex = __fadd_rn((float)ex, fw.z); c.ey = __fadd_rn((float)ey, fw.w);
      c.octave = l.octave;
      const float ddx = __fsub_rn(c.ex, c.sx), ddy = __fsub_rn(c.ey, c.sy);
      c.angle = (float)atan2((double)ddy, (double)ddx);
      c.pt_x = __fdiv_rn(__fadd_rn(c.sx, c.ex), 2.f); c.pt_y = __fdiv_rn(__fadd_rn(c.sy, c.ey), 2.f);
      c.size = __fmul_rn(ddx, ddy);
      c.length = (float)sqrt((double)ddx * (double)ddx + (double)ddy * (double)ddy);
      c.response = 0.f; c.class_id = -1; c.num_pixels = 0;
      c.sx_oct = 0.f; c.sy_oct = 0.f; c.ex_oct = 0.f; c.ey_oct = 0.f;
      corres[o] = c;
      flow_next[o] = fw;
      const double a0 = c.sx, a1 = c.sy, b0 = c.ex, b1 = c.ey;
      const double c0 = a1 - b1, c1 = b0 - a0, c2 = a0 * b1 - a1 * b0;      // (a0, a1, 1) x (b0, b1, 1)
      const double nn = sqrt(c0 * c0 + c1 * c1 + c2 * c2);
      inf_line[3 * o] = nn > 0 ? c0 / nn : c0; inf_line[3 * o + 1] = nn > 0 ? c1 / nn : c1; inf_line[3 * o + 2] = nn > 0 ? c2 / nn : c2;
      const float d_start = P.d(f, sy, sx), d_end = P.d(f, ey, ex);
      stat_depth[o] = make_float2((d_start > 0.f && d_end > 0.f) ? d_start : -1.f, d_end);     // the reference's missing braces, :759-761
      src_idx[o] = i;
    }
  }
  if (threadIdx.x == 0) { n_out[f] = base; n_obj[f] = obase; }
}

// P5: AssignFeaturesToGrid (Frame.cc:910-925): a stable counting sort of the key point indices by grid cell
__global__ void __launch_bounds__(256) k_grid(const sdpl_keypoint* __restrict__ kps, const int* __restrict__ n_in, int capacity, int w, int h, int gcols,
                                              int grows, int32_t* __restrict__ cell_start, int32_t* __restrict__ items, int32_t* __restrict__ cell_of) {
  extern __shared__ int s_cnt[];                      // gcols * grows + 1
  const int f = blockIdx.x, ncell = gcols * grows;
  const int n = min(n_in[f], capacity);
  const float wInv = __fdiv_rn((float)gcols, (float)w), hInv = __fdiv_rn((float)grows, (float)h);
  for (int c = threadIdx.x; c <= ncell; c += 256) s_cnt[c] = 0;
  __syncthreads();
  int32_t* cof = cell_of + (size_t)f * capacity;
  for (int i = threadIdx.x; i < n; i += 256) {
    const sdpl_keypoint kp = kps[(size_t)f * capacity + i];
    const int px = (int)roundf(__fmul_rn(kp.x, wInv)), py = (int)roundf(__fmul_rn(kp.y, hInv));
    const int c = (px < 0 || px >= gcols || py < 0 || py >= grows) ? -1 : px * grows + py;
    cof[i] = c;
    if (c >= 0) atomicAdd(&s_cnt[c], 1);
  }
  __syncthreads();
  if (threadIdx.x == 0) {                              // 3072 cells: a serial scan by one thread is a few microseconds
    int run = 0;
    for (int c = 0; c < ncell; c++) { const int v = s_cnt[c]; s_cnt[c] = run; run += v; }
    s_cnt[ncell] = run;
  }
  __syncthreads();
  int32_t* cs = cell_start + (size_t)f * (ncell + 1);
  for (int c = threadIdx.x; c <= ncell; c += 256) cs[c] = s_cnt[c];
  // stable placement: key point i goes behind the key points of its cell with a lower index
  for (int i = threadIdx.x; i < n; i += 256) {
    const int c = cof[i];
    if (c < 0) continue;
    int rank = 0;
    for (int j = 0; j < i; j++) rank += cof[j] == c;
    items[(size_t)f * capacity + s_cnt[c] + rank] = i;
  }
}

// P6: Frame::GetFeaturesInArea (Frame.cc:970-1023) on the grid of P5.  One warp per query: the cells of the window are visited in
// the reference's order (ix, then iy, then the cell's list), 32 list entries at a time, and the hits are appended in that order with
// a ballot -- what the reference's push_back gives.  query = {x, y, r, minLevel, maxLevel} (five floats)
__global__ void __launch_bounds__(128) k_features_in_area(const sdpl_keypoint* __restrict__ kps, int capacity, const int32_t* __restrict__ cell_start,
                                                          const int32_t* __restrict__ items, int w, int h, int gcols, int grows,
                                                          const float* __restrict__ queries, int nq, int32_t* __restrict__ out, int max_out,
                                                          int* __restrict__ counts) {
  const int f = blockIdx.y, lane = threadIdx.x & 31;
  const int q = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (q >= nq) return;
  const float* Q = queries + ((size_t)f * nq + q) * 5;
  const float x = Q[0], y = Q[1], r = Q[2];
  const int minLevel = (int)Q[3], maxLevel = (int)Q[4];
  const float wInv = __fdiv_rn((float)gcols, (float)w), hInv = __fdiv_rn((float)grows, (float)h);
  const int x0 = max(0, (int)floorf(__fmul_rn(__fsub_rn(x, r), wInv))), x1 = min(gcols - 1, (int)ceilf(__fmul_rn(__fadd_rn(x, r), wInv)));
  const int y0 = max(0, (int)floorf(__fmul_rn(__fsub_rn(y, r), hInv))), y1 = min(grows - 1, (int)ceilf(__fmul_rn(__fadd_rn(y, r), hInv)));
  int n = 0;
  if (!(x0 >= gcols || x1 < 0 || y0 >= grows || y1 < 0)) {
    const bool check = minLevel > 0 || maxLevel >= 0;
    const int32_t* cs = cell_start + (size_t)f * (gcols * grows + 1);
    const int32_t* it = items + (size_t)f * capacity;
    const sdpl_keypoint* K = kps + (size_t)f * capacity;
    int32_t* o = out + ((size_t)f * nq + q) * max_out;
    for (int ix = x0; ix <= x1; ix++) {
      // the cells (ix, y0 .. y1) are adjacent in the cell array: one contiguous run of the item list
      const int j0 = cs[ix * grows + y0], j1 = cs[ix * grows + y1 + 1];
      for (int jb = j0; jb < j1; jb += 32) {
        const int j = jb + lane;
        bool hit = false; int idx = 0;
        if (j < j1) {
          idx = it[j];
          const sdpl_keypoint kp = K[idx];
          const bool lvl = !check || (kp.octave >= minLevel && (maxLevel < 0 || kp.octave <= maxLevel));
          hit = lvl && fabsf(__fsub_rn(kp.x, x)) < r && fabsf(__fsub_rn(kp.y, y)) < r;
        }
        const uint32_t m = __ballot_sync(0xffffffffu, hit);
        if (hit) { const int s = n + __popc(m & ((1u << lane) - 1u)); if (s < max_out) o[s] = idx; }
        n += __popc(m);
      }
    }
  }
  if (lane == 0) counts[(size_t)f * nq + q] = n;
}

__device__ __forceinline__ int hamming32(const uint4& a0, const uint4& a1, const uint4& b0, const uint4& b1) {
  return __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) + __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) +
         __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
}

// P7: the descriptor search of a projection match on the window of P6: one warp per query walks the window in the reference's
// visiting order, 32 candidates at a time; every lane keeps the best / second best of its candidates as (distance << 22 | visiting
// rank) keys (strict '<' over the visiting order = smallest key), the lanes' pairs are merged with shuffles
__global__ void __launch_bounds__(128) k_search_area(const sdpl_keypoint* __restrict__ kps, const uint8_t* __restrict__ desc, int capacity,
                                                     const int32_t* __restrict__ cell_start, const int32_t* __restrict__ items, int w, int h, int gcols,
                                                     int grows, const float* __restrict__ queries, const uint8_t* __restrict__ qdesc, int nq,
                                                     int32_t* __restrict__ out5) {
  const int f = blockIdx.y, lane = threadIdx.x & 31;
  const int q = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (q >= nq) return;
  const float* Q = queries + ((size_t)f * nq + q) * 5;
  const float x = Q[0], y = Q[1], r = Q[2];
  const int minLevel = (int)Q[3], maxLevel = (int)Q[4];
  const uint4* qd = (const uint4*)(qdesc + ((size_t)f * nq + q) * 32);
  const uint4 q0 = __ldg(qd), q1 = __ldg(qd + 1);
  const float wInv = __fdiv_rn((float)gcols, (float)w), hInv = __fdiv_rn((float)grows, (float)h);
  const int x0 = max(0, (int)floorf(__fmul_rn(__fsub_rn(x, r), wInv))), x1 = min(gcols - 1, (int)ceilf(__fmul_rn(__fadd_rn(x, r), wInv)));
  const int y0 = max(0, (int)floorf(__fmul_rn(__fsub_rn(y, r), hInv))), y1 = min(grows - 1, (int)ceilf(__fmul_rn(__fadd_rn(y, r), hInv)));
  // key = distance (9 bits) << 22 | visiting rank (22 bits); payload = key point index
  uint32_t k1 = 0xFFFFFFFFu, k2 = 0xFFFFFFFFu; int i1 = -1, i2 = -1;
  if (!(x0 >= gcols || x1 < 0 || y0 >= grows || y1 < 0)) {
    const bool check = minLevel > 0 || maxLevel >= 0;
    const int32_t* cs = cell_start + (size_t)f * (gcols * grows + 1);
    const int32_t* it = items + (size_t)f * capacity;
    const sdpl_keypoint* K = kps + (size_t)f * capacity;
    const uint8_t* Dd = desc + (size_t)f * capacity * 32;
    int visited = 0;
    for (int ix = x0; ix <= x1; ix++) {
      const int j0 = cs[ix * grows + y0], j1 = cs[ix * grows + y1 + 1];
      for (int jb = j0; jb < j1; jb += 32) {
        const int j = jb + lane;
        if (j < j1) {
          const int idx = it[j];
          const sdpl_keypoint kp = K[idx];
          const bool lvl = !check || (kp.octave >= minLevel && (maxLevel < 0 || kp.octave <= maxLevel));
          if (lvl && fabsf(__fsub_rn(kp.x, x)) < r && fabsf(__fsub_rn(kp.y, y)) < r) {
            const uint4* dp = (const uint4*)(Dd + (size_t)idx * 32);
            const uint32_t key = ((uint32_t)hamming32(q0, q1, __ldg(dp), __ldg(dp + 1)) << 22) | (uint32_t)(visited + j - jb);
            if (key < k1) { k2 = k1; i2 = i1; k1 = key; i1 = idx; } else if (key < k2) { k2 = key; i2 = idx; }
          }
        }
        visited += min(32, j1 - jb);       // ranks follow the position in the visiting order (hits and misses alike)
      }
    }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    const uint32_t b1 = __shfl_xor_sync(0xffffffffu, k1, o), b2 = __shfl_xor_sync(0xffffffffu, k2, o);
    const int j1 = __shfl_xor_sync(0xffffffffu, i1, o), j2 = __shfl_xor_sync(0xffffffffu, i2, o);
    // two smallest of {k1 <= k2, b1 <= b2} (keys are unique: the rank is)
    uint32_t n1, n2; int m1, m2;
    if (b1 < k1) { n1 = b1; m1 = j1; if (k1 < b2) { n2 = k1; m2 = i1; } else { n2 = b2; m2 = j2; } }
    else { n1 = k1; m1 = i1; if (b1 < k2) { n2 = b1; m2 = j1; } else { n2 = k2; m2 = i2; } }
    k1 = n1; i1 = m1; k2 = n2; i2 = m2;
  }
  if (lane == 0) {
    int32_t* o = out5 + ((size_t)f * nq + q) * 5;
    const sdpl_keypoint* K = kps + (size_t)f * capacity;
    o[0] = i1; o[1] = i1 >= 0 ? (int)(k1 >> 22) : 256; o[2] = i1 >= 0 ? K[i1].octave : -1;
    o[3] = i2 >= 0 ? (int)(k2 >> 22) : 256; o[4] = i2 >= 0 ? K[i2].octave : -1;
  }
}

// P8: MapPoint::ComputeDistinctiveDescriptors (MapPoint.cc:242-307): one warp per map point with N <= 64 observations.  Row i of the
// distance matrix sits in the lanes (columns lane, lane + 32); its median is the element of rank (int)(0.5 (N - 1)) under the order
// (distance, column), found by counting; the first row with the least median wins.
__global__ void __launch_bounds__(128) k_distinctive(const uint8_t* __restrict__ desc, const int32_t* __restrict__ start, int n_points,
                                                     int32_t* __restrict__ best_idx, uint8_t* __restrict__ out_desc, int* __restrict__ err) {
  const int lane = threadIdx.x & 31;
  const int p = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (p >= n_points) return;
  const int s0 = start[p], N = start[p + 1] - s0;
  if (N <= 0) { if (lane == 0) best_idx[p] = -1; return; }
  if (N > 64) { if (lane == 0) { best_idx[p] = -1; atomicExch(err, SDPL_ERR_OVERFLOW); } return; }
  const uint4* D4 = (const uint4*)(desc + (size_t)s0 * 32);
  const int c0 = lane, c1 = lane + 32;
  uint4 a0 = make_uint4(0, 0, 0, 0), a1 = a0, b0 = a0, b1 = a0;
  if (c0 < N) { a0 = __ldg(D4 + 2 * c0); a1 = __ldg(D4 + 2 * c0 + 1); }
  if (c1 < N) { b0 = __ldg(D4 + 2 * c1); b1 = __ldg(D4 + 2 * c1 + 1); }
  const int kth = (int)(0.5 * (double)(N - 1));
  int bestMedian = 0x7fffffff, bestI = 0;
  for (int i = 0; i < N; i++) {
    // descriptor i to every lane
    const int src = i & 31;
    uint4 r0, r1;
    const uint4 s_lo0 = i < 32 ? a0 : b0, s_lo1 = i < 32 ? a1 : b1;
    r0.x = __shfl_sync(0xffffffffu, s_lo0.x, src); r0.y = __shfl_sync(0xffffffffu, s_lo0.y, src);
    r0.z = __shfl_sync(0xffffffffu, s_lo0.z, src); r0.w = __shfl_sync(0xffffffffu, s_lo0.w, src);
    r1.x = __shfl_sync(0xffffffffu, s_lo1.x, src); r1.y = __shfl_sync(0xffffffffu, s_lo1.y, src);
    r1.z = __shfl_sync(0xffffffffu, s_lo1.z, src); r1.w = __shfl_sync(0xffffffffu, s_lo1.w, src);
    const int d0 = c0 < N ? hamming32(r0, r1, a0, a1) : 1 << 20, d1 = c1 < N ? hamming32(r0, r1, b0, b1) : 1 << 20;
    // rank of my two elements among the N of the row, order (distance, column)
    int rk0 = 0, rk1 = 0;
    for (int j = 0; j < N; j++) {
      const int dj = j < 32 ? __shfl_sync(0xffffffffu, d0, j) : __shfl_sync(0xffffffffu, d1, j - 32);
      rk0 += (dj < d0) || (dj == d0 && j < c0);
      rk1 += (dj < d1) || (dj == d1 && j < c1);
    }
    const uint32_t hit0 = __ballot_sync(0xffffffffu, c0 < N && rk0 == kth), hit1 = __ballot_sync(0xffffffffu, c1 < N && rk1 == kth);
    int median;
    if (hit0) median = __shfl_sync(0xffffffffu, d0, __ffs(hit0) - 1); else median = __shfl_sync(0xffffffffu, d1, __ffs(hit1) - 1);
    if (median < bestMedian) { bestMedian = median; bestI = i; }
  }
  if (lane == 0) best_idx[p] = bestI;
  if (lane < 8) ((uint32_t*)(out_desc + (size_t)p * 32))[lane] = ((const uint32_t*)(desc + (size_t)(s0 + bestI) * 32))[lane];
}

// P9: MapPoint::PredictScale (MapPoint.cc:385-417)
__global__ void k_predict_scale(const float* __restrict__ max_distance, const float* __restrict__ current_dist, int n, float log_sf, int n_levels,
                                int32_t* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float ratio = __fdiv_rn(max_distance[i], current_dist[i]);
  // std::log(float) of the reference: the double logarithm rounded to float is the correctly rounded value
  int s = (int)ceilf(__fdiv_rn((float)log((double)ratio), log_sf));
  s = s < 0 ? 0 : (s >= n_levels ? n_levels - 1 : s);
  out[i] = s;
}

}  // namespace sdpl

using namespace sdpl;

struct sdpl_post {
  int device = 0;
  cudaStream_t own_stream = nullptr, stream = nullptr;
  DevBuf bitmap, offs, cellof;
  DevBuf stage[12];                       // host-entry staging
  int launches = 0;
  StageTimer timer;
};

static PostPlanes make_planes(const int32_t* mask, const float* depth, const float* flow, int w, int h) {
  PostPlanes P; P.mask = mask; P.depth = depth; P.flow = flow; P.w = w; P.h = h; P.plane = (size_t)w * h;
  return P;
}

extern "C" {

int sdpl_post_create(sdpl_post** out, int device) {
  if (!out) return SDPL_ERR_ARG;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
    set_last_error("sdpl_post_create: no such CUDA device (this library has no CPU fallback)");
    return SDPL_ERR_CUDA;
  }
  SDPL_CUDA(cudaSetDevice(device));
  sdpl_post* p = new sdpl_post;
  p->device = device;
  SDPL_CUDA(cudaStreamCreateWithFlags(&p->own_stream, cudaStreamNonBlocking));
  p->stream = p->own_stream;
  *out = p;
  return SDPL_OK;
}
void sdpl_post_destroy(sdpl_post* p) {
  if (!p) return;
  cudaSetDevice(p->device);
  cudaStreamSynchronize(p->stream);
  p->bitmap.release(); p->offs.release(); p->cellof.release();
  for (DevBuf& b : p->stage) b.release();
  p->timer.release();
  if (p->own_stream) cudaStreamDestroy(p->own_stream);
  delete p;
}
int sdpl_post_set_stream(sdpl_post* p, void* s) { if (!p) return SDPL_ERR_ARG; p->stream = s ? (cudaStream_t)s : p->own_stream; return SDPL_OK; }
int sdpl_post_last_launches(const sdpl_post* p) { return p ? p->launches : 0; }
int sdpl_post_set_profiling(sdpl_post* p, int on) { if (!p) return SDPL_ERR_ARG; p->timer.enabled = on != 0; return SDPL_OK; }
int sdpl_post_stage_times(sdpl_post* p, float* ms, const char** names, int* launches, int cap) {
  if (!p) return 0;
  cudaSetDevice(p->device);
  return p->timer.read(ms, names, launches, cap);
}

int sdpl_post_sample_objects_dev(sdpl_post* p, const int32_t* d_mask, const float* d_depth, const float* d_flow, int nframes, int w, int h, int step,
                                 float th_depth_obj, sdpl_keypoint* d_keys, sdpl_keypoint* d_corres, float* d_flow_next, float* d_depth_out,
                                 int32_t* d_label, int capacity, int* d_n, int sync) {
  if (!p || !d_mask || !d_depth || !d_flow || nframes < 1 || w < 1 || h < 1 || step < 1 || !d_keys || !d_corres || !d_flow_next || !d_depth_out ||
      !d_label || capacity < 1 || !d_n) { set_last_error("sdpl_post_sample_objects_dev: bad argument"); return SDPL_ERR_ARG; }
  SDPL_CUDA(cudaSetDevice(p->device));
  const int cols_s = div_up(w, step), rows_s = div_up(h, step), wpr = div_up(cols_s, 32), nwords = wpr * rows_s;
  int rc;
  if ((rc = p->bitmap.reserve(sizeof(uint32_t) * (size_t)nwords * nframes))) return rc;
  if ((rc = p->offs.reserve(sizeof(uint32_t) * (size_t)nwords * nframes))) return rc;
  const PostPlanes P = make_planes(d_mask, d_depth, d_flow, w, h);
  g_launches = 0;
  p->timer.begin(p->stream);
  k_obj_flags<<<dim3(div_up(nwords, 8), nframes), 256, 0, p->stream>>>(P, step, cols_s, wpr, rows_s, th_depth_obj, p->bitmap.as<uint32_t>());
  SDPL_LAUNCH_CHECK();
  k_obj_scan<<<nframes, 1024, 0, p->stream>>>(p->bitmap.as<uint32_t>(), nwords, p->offs.as<uint32_t>(), d_n);
  SDPL_LAUNCH_CHECK();
  k_obj_emit<<<dim3(div_up(nwords, 8), nframes), 256, 0, p->stream>>>(P, step, cols_s, wpr, rows_s, p->bitmap.as<uint32_t>(), p->offs.as<uint32_t>(), d_keys,
                                                                      d_corres, (float2*)d_flow_next, d_depth_out, d_label, capacity);
  SDPL_LAUNCH_CHECK();
  p->timer.mark(p->stream, "object_sampling");
  p->launches = g_launches;
  if (sync) SDPL_CUDA(cudaStreamSynchronize(p->stream));
  return SDPL_OK;
}

int sdpl_post_filter_lines_dev(sdpl_post* p, const int32_t* d_mask, const float* d_depth, int nframes, int w, int h, const sdpl_keyline* d_kls,
                               const int* d_n_in, int capacity, sdpl_keyline* d_out, int32_t* d_keep_idx, int* d_n_out, int sync) {
  if (!p || !d_mask || !d_depth || nframes < 1 || w < 1 || h < 1 || !d_kls || !d_n_in || capacity < 1 || !d_out || !d_keep_idx || !d_n_out) {
    set_last_error("sdpl_post_filter_lines_dev: bad argument"); return SDPL_ERR_ARG;
  }
  SDPL_CUDA(cudaSetDevice(p->device));
  g_launches = 0;
  p->timer.begin(p->stream);
  k_filter_lines<<<nframes, 256, 0, p->stream>>>(make_planes(d_mask, d_depth, nullptr, w, h), d_kls, d_n_in, capacity, d_out, d_keep_idx, d_n_out);
  SDPL_LAUNCH_CHECK();
  p->timer.mark(p->stream, "line_filters");
  p->launches = g_launches;
  if (sync) SDPL_CUDA(cudaStreamSynchronize(p->stream));
  return SDPL_OK;
}

int sdpl_post_point_corres_dev(sdpl_post* p, const int32_t* d_mask, const float* d_depth, const float* d_flow, int nframes, int w, int h,
                               const sdpl_keypoint* d_kps, const int* d_n_in, int capacity, float th_depth, sdpl_keypoint* d_stat,
                               sdpl_keypoint* d_corres, float* d_flow_next, float* d_stat_depth, int32_t* d_src_idx, int* d_n_out, int sync) {
  if (!p || !d_mask || !d_depth || !d_flow || nframes < 1 || w < 1 || h < 1 || !d_kps || !d_n_in || capacity < 1 || !d_stat || !d_corres ||
      !d_flow_next || !d_stat_depth || !d_src_idx || !d_n_out) { set_last_error("sdpl_post_point_corres_dev: bad argument"); return SDPL_ERR_ARG; }
  SDPL_CUDA(cudaSetDevice(p->device));
  g_launches = 0;
  p->timer.begin(p->stream);
  k_point_corres<<<nframes, 256, 0, p->stream>>>(make_planes(d_mask, d_depth, d_flow, w, h), d_kps, d_n_in, capacity, th_depth, d_stat, d_corres,
                                                 (float2*)d_flow_next, d_stat_depth, d_src_idx, d_n_out);
  SDPL_LAUNCH_CHECK();
  p->timer.mark(p->stream, "point_corres");
  p->launches = g_launches;
  if (sync) SDPL_CUDA(cudaStreamSynchronize(p->stream));
  return SDPL_OK;
}

int sdpl_post_line_corres_dev(sdpl_post* p, const int32_t* d_mask, const float* d_depth, const float* d_flow, int nframes, int w, int h,
                              const sdpl_keyline* d_kls, const int* d_n_in, int capacity, float th_depth, sdpl_keyline* d_obj, int* d_n_obj,
                              sdpl_keyline* d_stat, sdpl_keyline* d_corres, float* d_flow_next, double* d_inf_line, float* d_stat_depth,
                              int32_t* d_src_idx, int* d_n_out, int sync) {
  if (!p || !d_mask || !d_depth || !d_flow || nframes < 1 || w < 1 || h < 1 || !d_kls || !d_n_in || capacity < 1 || !d_obj || !d_n_obj || !d_stat ||
      !d_corres || !d_flow_next || !d_inf_line || !d_stat_depth || !d_src_idx || !d_n_out) {
    set_last_error("sdpl_post_line_corres_dev: bad argument"); return SDPL_ERR_ARG;
  }
  SDPL_CUDA(cudaSetDevice(p->device));
  g_launches = 0;
  p->timer.begin(p->stream);
  k_line_corres<<<nframes, 256, 0, p->stream>>>(make_planes(d_mask, d_depth, d_flow, w, h), d_kls, d_n_in, capacity, th_depth, d_obj, d_n_obj, d_stat,
                                                d_corres, (float4*)d_flow_next, d_inf_line, (float2*)d_stat_depth, d_src_idx, d_n_out);
  SDPL_LAUNCH_CHECK();
  p->timer.mark(p->stream, "line_corres");
  p->launches = g_launches;
  if (sync) SDPL_CUDA(cudaStreamSynchronize(p->stream));
  return SDPL_OK;
}

int sdpl_post_grid_dev(sdpl_post* p, int nframes, int w, int h, const sdpl_keypoint* d_kps, const int* d_n_in, int capacity, int grid_cols,
                       int grid_rows, int32_t* d_cell_start, int32_t* d_items, int sync) {
  if (!p || nframes < 1 || w < 1 || h < 1 || !d_kps || !d_n_in || capacity < 1 || grid_cols < 1 || grid_rows < 1 || grid_cols * grid_rows > 11000 ||
      !d_cell_start || !d_items) { set_last_error("sdpl_post_grid_dev: bad argument"); return SDPL_ERR_ARG; }
  SDPL_CUDA(cudaSetDevice(p->device));
  int rc;
  if ((rc = p->cellof.reserve(sizeof(int32_t) * (size_t)capacity * nframes))) return rc;
  g_launches = 0;
  p->timer.begin(p->stream);
  k_grid<<<nframes, 256, sizeof(int) * (grid_cols * grid_rows + 1), p->stream>>>(d_kps, d_n_in, capacity, w, h, grid_cols, grid_rows, d_cell_start, d_items,
                                                                                  p->cellof.as<int32_t>());
  SDPL_LAUNCH_CHECK();
  p->timer.mark(p->stream, "grid");
  p->launches = g_launches;
  if (sync) SDPL_CUDA(cudaStreamSynchronize(p->stream));
  return SDPL_OK;
}

int sdpl_post_features_in_area_dev(sdpl_post* p, int nframes, int w, int h, const sdpl_keypoint* d_kps, int capacity, const int32_t* d_cell_start,
                                   const int32_t* d_items, int grid_cols, int grid_rows, const float* d_queries, int nq, int32_t* d_out, int max_out,
                                   int* d_counts, int sync) {
  if (!p || nframes < 1 || w < 1 || h < 1 || !d_kps || capacity < 1 || !d_cell_start || !d_items || grid_cols < 1 || grid_rows < 1 || !d_queries ||
      nq < 1 || !d_out || max_out < 1 || !d_counts) { set_last_error("sdpl_post_features_in_area_dev: bad argument"); return SDPL_ERR_ARG; }
  SDPL_CUDA(cudaSetDevice(p->device));
  g_launches = 0;
  p->timer.begin(p->stream);
  k_features_in_area<<<dim3(div_up(nq, 4), nframes), 128, 0, p->stream>>>(d_kps, capacity, d_cell_start, d_items, w, h, grid_cols, grid_rows, d_queries, nq,
                                                                          d_out, max_out, d_counts);
  SDPL_LAUNCH_CHECK();
  p->timer.mark(p->stream, "features_in_area");
  p->launches = g_launches;
  if (sync) SDPL_CUDA(cudaStreamSynchronize(p->stream));
  return SDPL_OK;
}

int sdpl_post_search_area_dev(sdpl_post* p, int nframes, int w, int h, const sdpl_keypoint* d_kps, const uint8_t* d_desc, int capacity,
                              const int32_t* d_cell_start, const int32_t* d_items, int grid_cols, int grid_rows, const float* d_queries,
                              const uint8_t* d_qdesc, int nq, int32_t* d_out5, int sync) {
  if (!p || nframes < 1 || w < 1 || h < 1 || !d_kps || !d_desc || capacity < 1 || !d_cell_start || !d_items || grid_cols < 1 || grid_rows < 1 ||
      !d_queries || !d_qdesc || nq < 1 || !d_out5) { set_last_error("sdpl_post_search_area_dev: bad argument"); return SDPL_ERR_ARG; }
  SDPL_CUDA(cudaSetDevice(p->device));
  g_launches = 0;
  p->timer.begin(p->stream);
  k_search_area<<<dim3(div_up(nq, 4), nframes), 128, 0, p->stream>>>(d_kps, d_desc, capacity, d_cell_start, d_items, w, h, grid_cols, grid_rows, d_queries,
                                                                     d_qdesc, nq, d_out5);
  SDPL_LAUNCH_CHECK();
  p->timer.mark(p->stream, "search_area");
  p->launches = g_launches;
  if (sync) SDPL_CUDA(cudaStreamSynchronize(p->stream));
  return SDPL_OK;
}

int sdpl_post_distinctive_descriptors_dev(sdpl_post* p, const uint8_t* d_desc, const int32_t* d_start, int n_points, int32_t* d_best_idx,
                                          uint8_t* d_out_desc, int sync) {
  if (!p || !d_desc || !d_start || n_points < 1 || !d_best_idx || !d_out_desc) { set_last_error("sdpl_post_distinctive_descriptors_dev: bad argument"); return SDPL_ERR_ARG; }
  SDPL_CUDA(cudaSetDevice(p->device));
  int rc;
  if ((rc = p->stage[11].reserve(sizeof(int)))) return rc;
  SDPL_CUDA(cudaMemsetAsync(p->stage[11].p, 0, sizeof(int), p->stream));
  g_launches = 0;
  p->timer.begin(p->stream);
  k_distinctive<<<div_up(n_points, 4), 128, 0, p->stream>>>(d_desc, d_start, n_points, d_best_idx, d_out_desc, p->stage[11].as<int>());
  SDPL_LAUNCH_CHECK();
  p->timer.mark(p->stream, "distinctive_descriptors");
  p->launches = g_launches;
  if (sync) {
    int e = 0;
    SDPL_CUDA(cudaMemcpyAsync(&e, p->stage[11].p, sizeof(int), cudaMemcpyDeviceToHost, p->stream));
    SDPL_CUDA(cudaStreamSynchronize(p->stream));
    if (e) { set_last_error("sdpl_post_distinctive_descriptors_dev: a map point has more than 64 observations"); return SDPL_ERR_OVERFLOW; }
  }
  return SDPL_OK;
}

int sdpl_post_predict_scale_dev(sdpl_post* p, const float* d_max_distance, const float* d_current_dist, int n, float log_scale_factor, int n_levels,
                                int32_t* d_out, int sync) {
  if (!p || !d_max_distance || !d_current_dist || n < 1 || !(log_scale_factor > 0.f) || n_levels < 1 || !d_out) {
    set_last_error("sdpl_post_predict_scale_dev: bad argument"); return SDPL_ERR_ARG;
  }
  SDPL_CUDA(cudaSetDevice(p->device));
  g_launches = 0;
  k_predict_scale<<<div_up(n, 256), 256, 0, p->stream>>>(d_max_distance, d_current_dist, n, log_scale_factor, n_levels, d_out);
  SDPL_LAUNCH_CHECK();
  p->launches = g_launches;
  if (sync) SDPL_CUDA(cudaStreamSynchronize(p->stream));
  return SDPL_OK;
}

// ---- host-buffer entry point for one frame (what Frame::Frame would call): planes in, object samples out ----
int sdpl_post_sample_objects(sdpl_post* p, const int32_t* mask, const float* depth, const float* flow, int w, int h, int step, float th_depth_obj,
                             sdpl_keypoint* keys, sdpl_keypoint* corres, float* flow_next, float* depth_out, int32_t* label, int capacity, int* n_out) {
  if (!p || !mask || !depth || !flow || w < 1 || h < 1 || step < 1 || !keys || !corres || !flow_next || !depth_out || !label || capacity < 1 || !n_out) {
    set_last_error("sdpl_post_sample_objects: bad argument"); return SDPL_ERR_ARG;
  }
  SDPL_CUDA(cudaSetDevice(p->device));
  const size_t px = (size_t)w * h;
  int rc;
  const size_t sizes[9] = {px * 4, px * 4, px * 8, sizeof(sdpl_keypoint) * capacity, sizeof(sdpl_keypoint) * capacity, (size_t)8 * capacity,
                           (size_t)4 * capacity, (size_t)4 * capacity, sizeof(int)};
  for (int i = 0; i < 9; i++) if ((rc = p->stage[i].reserve(sizes[i]))) return rc;
  cudaStream_t st = p->stream;
  SDPL_CUDA(cudaMemcpyAsync(p->stage[0].p, mask, px * 4, cudaMemcpyHostToDevice, st));
  SDPL_CUDA(cudaMemcpyAsync(p->stage[1].p, depth, px * 4, cudaMemcpyHostToDevice, st));
  SDPL_CUDA(cudaMemcpyAsync(p->stage[2].p, flow, px * 8, cudaMemcpyHostToDevice, st));
  rc = sdpl_post_sample_objects_dev(p, p->stage[0].as<int32_t>(), p->stage[1].as<float>(), p->stage[2].as<float>(), 1, w, h, step, th_depth_obj,
                                    p->stage[3].as<sdpl_keypoint>(), p->stage[4].as<sdpl_keypoint>(), p->stage[5].as<float>(), p->stage[6].as<float>(),
                                    p->stage[7].as<int32_t>(), capacity, p->stage[8].as<int>(), 0);
  if (rc) return rc;
  SDPL_CUDA(cudaMemcpyAsync(n_out, p->stage[8].p, sizeof(int), cudaMemcpyDeviceToHost, st));
  SDPL_CUDA(cudaStreamSynchronize(st));
  const int n = std::min(*n_out, capacity);
  if (n > 0) {
    SDPL_CUDA(cudaMemcpyAsync(keys, p->stage[3].p, sizeof(sdpl_keypoint) * n, cudaMemcpyDeviceToHost, st));
    SDPL_CUDA(cudaMemcpyAsync(corres, p->stage[4].p, sizeof(sdpl_keypoint) * n, cudaMemcpyDeviceToHost, st));
    SDPL_CUDA(cudaMemcpyAsync(flow_next, p->stage[5].p, (size_t)8 * n, cudaMemcpyDeviceToHost, st));
    SDPL_CUDA(cudaMemcpyAsync(depth_out, p->stage[6].p, (size_t)4 * n, cudaMemcpyDeviceToHost, st));
    SDPL_CUDA(cudaMemcpyAsync(label, p->stage[7].p, (size_t)4 * n, cudaMemcpyDeviceToHost, st));
    SDPL_CUDA(cudaStreamSynchronize(st));
  }
  if (*n_out > capacity) { set_last_error("sdpl_post_sample_objects: more samples than the capacity"); return SDPL_ERR_CAPACITY; }
  return SDPL_OK;
}

}  // extern "C"
