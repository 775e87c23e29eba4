// sdpl_slam_b200/csrc/edlines.cuh -- the EDLines back-end of Lineextractor (extractor == 1; reference: LSDDetectorC::detectImpl_ED,
// 3rdparty/line_descriptor/src/LSDDetector_custom.cpp:386-461 -> EDLines::EDLines(Mat), ED_Lib/EDLines.cpp:8-70, ED_Lib/ED.cpp:8-62).
// Included by line.cu after LineDev / OctDev.  Per (frame, octave) task:
//   k_ed_smooth   5x5 sigma 1 Gaussian of the pyramid level (OpenCV fixed point: Q8.8 rows, Q16.16 columns, round half up)   [pixel-parallel]
//   k_ed_grad     Sobel-weighted |gx| + |gy| with threshold 36, vertical / horizontal direction map (ED.cpp:275-362)           [pixel-parallel]
//   k_ed_anchor   anchors: gradient maxima by >= 8 across the edge direction (ED.cpp:364-398)                                  [pixel-parallel]
//   k_ed_sort     anchors by descending gradient, row-major inside a value (ED.cpp:1000-1047 + the descending loop of :414)   [one warp per task]
//   k_ed_link     anchor linking (include/sdpl_edlines_core.h)                                                                [one thread per task]
//   k_ed_fit      line fitting, joining, validation (same header; lane per segment / per line)                                [one warp per task]
// The linking is sequential per image by definition (a walk stops at the pixels earlier walks have drawn), like LSD's region growing;
// this first version runs the whole sequential tail in one thread per task, so a batch is as fast as its slowest image, and all tasks
// of the batch run side by side.  Results go into the Pending slots of the LSD path (accepted = 1), so key-line construction, the top-N
// filter and LBD are the kernels of the LSD back-end.
#pragma once
#include "../../include/sdpl_edlines_core.h"

namespace sdpl {

constexpr int kEdNfaN = 8192;            // validation look-up table: pixel counts below this (a 2-px-wide rectangle along a joined line: up to 3 x the image diagonal)

struct EdOct {
  int w, h, npx;
  unsigned long long img_off;            // bytes inside one frame's image block: smooth[npx] dir[npx] edge[npx] pad grad[npx] (s16)
  unsigned long long work_off;           // bytes inside one frame's scratch block
  int anchors_cap, pixels_cap, stack_cap, chains_cap, chain_nos_cap, seg_px_cap, seg_cap, lines_cap;
  int min_line_len;
};
struct EdDev {
  EdOct O[kMaxOct];
  unsigned long long img_frame, work_frame;
  uint8_t* img; uint8_t* work;
  int* n_anchors; int* nseg;             // per task
  const double* atan_lut;                // [1025]
  const int* nfa_min_k;                  // [nl][kEdNfaN]
};

__device__ __forceinline__ size_t ed_align(size_t v) { return (v + 15) & ~(size_t)15; }
struct EdPtrs {
  uint8_t *smooth, *dir, *edge; int16_t* grad;
  int* anchors; int* pixels; sdpl_ed::Node* stack; sdpl_ed::Chain* chains; int* chain_nos; int* seg_px; int* seg_off; sdpl_ed::Line* lines;
};
__host__ __device__ inline size_t ed_img_bytes(int npx) { return (((size_t)3 * npx + 15) & ~(size_t)15) + (size_t)2 * npx; }
__host__ __device__ inline size_t ed_work_bytes(const EdOct& O) {
  size_t b = 0;
  b += (((size_t)O.anchors_cap * 4 + 15) & ~(size_t)15);
  b += (((size_t)O.pixels_cap * 4 + 15) & ~(size_t)15);
  b += (size_t)O.stack_cap * sizeof(sdpl_ed::Node);
  b += (((size_t)O.chains_cap * sizeof(sdpl_ed::Chain) + 15) & ~(size_t)15);
  b += (((size_t)O.chain_nos_cap * 4 + 15) & ~(size_t)15);
  b += (((size_t)O.seg_px_cap * 4 + 15) & ~(size_t)15);
  b += (((size_t)O.seg_cap * 4 + 15) & ~(size_t)15);
  b += (size_t)O.lines_cap * sizeof(sdpl_ed::Line);
  return (b + 15) & ~(size_t)15;
}
__device__ __forceinline__ EdPtrs ed_ptrs(const EdDev& E, int f, int o) {
  const EdOct& O = E.O[o];
  EdPtrs P;
  uint8_t* im = E.img + (size_t)f * E.img_frame + O.img_off;
  P.smooth = im; P.dir = im + O.npx; P.edge = im + 2 * (size_t)O.npx;
  P.grad = (int16_t*)(im + ed_align((size_t)3 * O.npx));
  uint8_t* wk = E.work + (size_t)f * E.work_frame + O.work_off;
  P.anchors = (int*)wk; wk += ed_align((size_t)O.anchors_cap * 4);
  P.pixels = (int*)wk; wk += ed_align((size_t)O.pixels_cap * 4);
  P.stack = (sdpl_ed::Node*)wk; wk += (size_t)O.stack_cap * sizeof(sdpl_ed::Node);
  P.chains = (sdpl_ed::Chain*)wk; wk += ed_align((size_t)O.chains_cap * sizeof(sdpl_ed::Chain));
  P.chain_nos = (int*)wk; wk += ed_align((size_t)O.chain_nos_cap * 4);
  P.seg_px = (int*)wk; wk += ed_align((size_t)O.seg_px_cap * 4);
  P.seg_off = (int*)wk; wk += ed_align((size_t)O.seg_cap * 4);
  P.lines = (sdpl_ed::Line*)wk;
  return P;
}
__device__ __forceinline__ void ed_level(const LineDev& D, int f, int o, const uint8_t*& src, int& stride) {
  if (o == 0) { src = D.in + (size_t)f * D.in_frame; stride = D.in_stride; }
  else { src = D.lvl + (size_t)f * D.lvl_frame + D.O[o].lvl_off; stride = D.O[o].w; }
}
__device__ __forceinline__ int ed_r101(int v, int n) { if (v < 0) v = -v; if (v >= n) v = 2 * n - 2 - v; return v; }

__global__ void __launch_bounds__(256) k_ed_smooth(LineDev D, EdDev E, int o) {
  const EdOct& O = E.O[o];
  const int f = blockIdx.z;
  const int x = blockIdx.x * 64 + (threadIdx.x & 63), y = blockIdx.y * 4 + (threadIdx.x >> 6);
  if (x >= O.w || y >= O.h) return;
  const uint8_t* src; int stride;
  ed_level(D, f, o, src, stride);
  const int k[5] = {14, 62, 104, 62, 14};
  int xs[5];
#pragma unroll
  for (int j = 0; j < 5; j++) xs[j] = ed_r101(x + j - 2, O.w);
  uint32_t acc = 0;
#pragma unroll
  for (int i = 0; i < 5; i++) {
    const uint8_t* row = src + (size_t)ed_r101(y + i - 2, O.h) * stride;
    uint32_t hsum = 0;
#pragma unroll
    for (int j = 0; j < 5; j++) hsum += (uint32_t)k[j] * row[xs[j]];
    acc += (uint32_t)k[i] * hsum;
  }
  ed_ptrs(E, f, o).smooth[(size_t)y * O.w + x] = (uint8_t)min(255u, (acc + (1u << 15)) >> 16);
}

__global__ void __launch_bounds__(256) k_ed_grad(EdDev E, int o) {
  const EdOct& O = E.O[o];
  const int f = blockIdx.z, w = O.w, h = O.h;
  const int j = blockIdx.x * 64 + (threadIdx.x & 63), i = blockIdx.y * 4 + (threadIdx.x >> 6);
  if (j >= w || i >= h) return;
  const EdPtrs P = ed_ptrs(E, f, o);
  const uint8_t* s = P.smooth;
  int sum = sdpl_ed::kGradThresh - 1, d = 0;
  if (i >= 1 && i < h - 1 && j >= 1 && j < w - 1) {
    const int com1 = (int)s[(i + 1) * w + j + 1] - (int)s[(i - 1) * w + j - 1];
    const int com2 = (int)s[(i - 1) * w + j + 1] - (int)s[(i + 1) * w + j - 1];
    const int gx = abs(com1 + com2 + 2 * ((int)s[i * w + j + 1] - (int)s[i * w + j - 1]));
    const int gy = abs(com1 - com2 + 2 * ((int)s[(i + 1) * w + j] - (int)s[(i - 1) * w + j]));
    sum = gx + gy;
    if (sum >= sdpl_ed::kGradThresh) d = gx >= gy ? sdpl_ed::kVertical : sdpl_ed::kHorizontal;
  }
  P.grad[i * w + j] = (int16_t)sum;
  P.dir[i * w + j] = (uint8_t)d;
}

__global__ void __launch_bounds__(256) k_ed_anchor(EdDev E, int o) {
  const EdOct& O = E.O[o];
  const int f = blockIdx.z, w = O.w, h = O.h;
  const int j = blockIdx.x * 64 + (threadIdx.x & 63), i = blockIdx.y * 4 + (threadIdx.x >> 6);
  if (j >= w || i >= h) return;
  const EdPtrs P = ed_ptrs(E, f, o);
  int e = 0;
  if (i >= 2 && i < h - 2 && j >= 2 && j < w - 2) {
    const int g = P.grad[i * w + j];
    if (g >= sdpl_ed::kGradThresh) {
      int d1, d2;
      if (P.dir[i * w + j] == sdpl_ed::kVertical) { d1 = g - P.grad[i * w + j - 1]; d2 = g - P.grad[i * w + j + 1]; }
      else { d1 = g - P.grad[(i - 1) * w + j]; d2 = g - P.grad[(i + 1) * w + j]; }
      if (d1 >= sdpl_ed::kAnchorThresh && d2 >= sdpl_ed::kAnchorThresh) e = sdpl_ed::kAnchor;
    }
  }
  P.edge[i * w + j] = (uint8_t)e;
}

// one warp per task: counting sort of the anchors by gradient value (descending), row-major inside a value.  Pass 1 compacts the anchors
// in pixel order (four pixels per lane and trip, warp scan of the counts) into the walk's pixel buffer, which is idle until the
// linking, and counts the values; pass 2 places 32 listed anchors per trip: lanes with the same value are ranked with match_any, so
// the position inside a value's run stays the row-major rank.
__global__ void __launch_bounds__(32) k_ed_sort(LineDev D, EdDev E) {
  __shared__ int hist[2048];
  const uint32_t FULL = 0xffffffffu;
  const int task = blockIdx.x, f = task / D.nl, o = task % D.nl, lane = threadIdx.x;
  const uint32_t lt = (1u << lane) - 1u;
  const EdOct& O = E.O[o];
  const EdPtrs P = ed_ptrs(E, f, o);
  int* list = P.pixels;
  for (int i = lane; i < 2048; i += 32) hist[i] = 0;
  __syncwarp();
  int total = 0;
  bool over = false;
  for (int q0 = 0; q0 < O.npx; q0 += 128) {
    uint32_t mine = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int q = q0 + 4 * lane + j;
      if (q < O.npx && P.edge[q] == sdpl_ed::kAnchor) mine |= 1u << j;
    }
    const int c = __popc(mine);
    int incl = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int v = __shfl_up_sync(FULL, incl, d); if (lane >= d) incl += v; }
    const int chunk = __shfl_sync(FULL, incl, 31);
    if (chunk == 0) continue;
    if (total + chunk > O.anchors_cap || total + chunk > O.pixels_cap) { over = true; break; }
    int pos = total + incl - c;
    for (uint32_t b = mine; b; b &= b - 1) {
      const int q = q0 + 4 * lane + (__ffs(b) - 1);
      list[pos++] = q;
      atomicAdd(&hist[min((int)P.grad[q], 2047)], 1);
    }
    total += chunk;
  }
  if (over) { if (lane == 0) { atomicOr(D.err, SDPL_ERR_OVERFLOW); E.n_anchors[task] = 0; } return; }
  __syncwarp();
  // start of every value's run in descending order of the value: lane-strided exclusive scan (64 values per lane), highest value first
  int local = 0;
  for (int b = 0; b < 64; b++) local += hist[2047 - (lane * 64 + b)];
  int incl = local;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) { const int v = __shfl_up_sync(FULL, incl, d); if (lane >= d) incl += v; }
  int run = incl - local;
  for (int b = 0; b < 64; b++) { const int idx = 2047 - (lane * 64 + b); const int c = hist[idx]; hist[idx] = run; run += c; }
  __syncwarp();
  for (int i0 = 0; i0 < total; i0 += 32) {
    const int i = i0 + lane;
    const bool have = i < total;
    const int q = have ? list[i] : 0;
    const int g = have ? min((int)P.grad[q], 2047) : -1 - lane;
    const uint32_t grp = __match_any_sync(FULL, g);
    int at = 0;
    if (have) at = hist[g] + __popc(grp & lt);
    __syncwarp();
    if (have && (grp & lt) == 0) hist[g] += __popc(grp);      // the first lane of a value's group moves the run on
    if (have) P.anchors[at] = q;
    __syncwarp();
  }
  if (lane == 0) E.n_anchors[task] = total;
}

// one task per CTA, worked by its first lane: tasks must not share a warp (32 divergent tasks in a warp run one after the other)
__device__ __forceinline__ void ed_work(const LineDev& D, const EdDev& E, int task, sdpl_ed::Work& W) {
  const int f = task / D.nl, o = task % D.nl;
  const EdOct& O = E.O[o];
  const EdPtrs P = ed_ptrs(E, f, o);
  W.w = O.w; W.h = O.h; W.grad = P.grad; W.dir = P.dir; W.edge = P.edge;
  W.anchors = P.anchors; W.n_anchors = E.n_anchors[task];
  W.pixels = P.pixels; W.pixels_cap = O.pixels_cap;
  W.stack = P.stack; W.stack_cap = O.stack_cap;
  W.chains = P.chains; W.chains_cap = O.chains_cap;
  W.chain_nos = P.chain_nos; W.chain_nos_cap = O.chain_nos_cap;
  W.seg_px = P.seg_px; W.seg_px_cap = O.seg_px_cap;
  W.seg_off = P.seg_off; W.seg_cap = O.seg_cap; W.nseg = 0;
  W.lines = P.lines; W.lines_cap = O.lines_cap; W.nlines = 0;
  ed_level(D, f, o, W.src, W.src_stride);
  W.src_magic = sdpl_ed::src_magic_of(O.w);
  W.atan_lut = E.atan_lut;
  W.nfa_min_k = E.nfa_min_k + (size_t)o * kEdNfaN; W.nfa_n = kEdNfaN;
  W.min_line_len = O.min_line_len;
  W.err = 0;
}
// anchor linking: the walks of one image are sequential by definition
__global__ void __launch_bounds__(32) k_ed_link(LineDev D, EdDev E) {
  const int task = blockIdx.x;
  if (threadIdx.x != 0) return;
  sdpl_ed::Work W;
  ed_work(D, E, task, W);
  sdpl_ed::link_anchors(W);
  E.nseg[task] = W.err ? -1 : W.nseg;
  if (W.err) atomicOr(D.err, SDPL_ERR_OVERFLOW);
}
// line fitting, joining and validation of one task, by one CTA of kEdFitThreads threads: segments are independent of each other
// (fitting + joining: thread per segment, the lines of segment s written at the prefix of the bounds len(s) / min_line_len -- every line
// uses up at least min_line_len pixels -- so they keep the reference's (segment, position) order), and so are the lines in the
// validation (thread per line, verdicts into a flag array, ordered compaction into the Pending slots by the first warp).  Threads of a
// warp diverge in these loops (ncu: 4.8 of 32 threads per instruction), so the parallelism that counts is the number of warps.
constexpr int kEdFitThreads = 256;
__global__ void __launch_bounds__(kEdFitThreads) k_ed_fit(LineDev D, EdDev E) {
  const uint32_t FULL = 0xffffffffu;
  const int task = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
  sdpl_ed::Work W;
  ed_work(D, E, task, W);
  const int nseg = E.nseg[task];
  lsd::Pending* pend = D.pend + (size_t)task * D.pend_cap;
  if (nseg < 0) { if (tid == 0) D.npend[task] = 0; return; }
  const int o = task % D.nl;
  const int seg_cap = E.O[o].seg_cap;
  int* cnt = W.chain_nos;                  // [nseg]     lines per segment after joining
  int* off = W.chain_nos + seg_cap;        // [nseg + 1] first line slot of a segment in W.lines
  int* pre = W.chain_nos + 2 * seg_cap;    // [nseg + 1] prefix of cnt
  uint8_t* vflag = (uint8_t*)W.pixels;     // [J] verdicts (the walk's pixel buffer is idle here)
  __shared__ int s_J, s_bad;
  if (tid == 0) {
    int run = 0;
    for (int s = 0; s < nseg; s++) { off[s] = run; run += (W.seg_off[s + 1] - W.seg_off[s]) / W.min_line_len; }
    off[nseg] = run;
    s_bad = run > W.lines_cap;
  }
  __syncthreads();
  if (s_bad) { if (tid == 0) { atomicOr(D.err, SDPL_ERR_OVERFLOW); D.npend[task] = 0; } return; }
  for (int s = tid; s < nseg; s += kEdFitThreads) {
    const int n = sdpl_ed::split_segment(W, s, W.lines + off[s], off[s + 1] - off[s]);
    cnt[s] = sdpl_ed::join_segment(W.lines + off[s], n);
  }
  __syncthreads();
  if (tid == 0) { int run = 0; for (int s = 0; s < nseg; s++) { pre[s] = run; run += cnt[s]; } pre[nseg] = run; s_J = run; }
  __syncthreads();
  const int J = s_J;
  for (int t = tid; t < J; t += kEdFitThreads) {
    int lo = 0, hi = nseg;                 // the segment of line t: pre[lo] <= t < pre[lo + 1]
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (pre[mid] <= t) lo = mid; else hi = mid; }
    vflag[t] = sdpl_ed::validate_one(W, W.lines[off[lo] + (t - pre[lo])]) ? 1 : 0;
  }
  if (W.err) atomicOr(&s_bad, 1);
  __syncthreads();
  if (tid >= 32) return;
  int base = 0;
  for (int t0 = 0; t0 < J; t0 += 32) {
    const int t = t0 + lane;
    const bool valid = t < J && vflag[t];
    const uint32_t m = __ballot_sync(FULL, valid);
    if (valid) {
      int lo = 0, hi = nseg;
      while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (pre[mid] <= t) lo = mid; else hi = mid; }
      const sdpl_ed::Line& l = W.lines[off[lo] + (t - pre[lo])];
      const int idx = base + __popc(m & ((1u << lane) - 1u));
      if (idx < D.pend_cap) {
        pend[idx].accepted = 1;
        pend[idx].seg[0] = (float)l.sx; pend[idx].seg[1] = (float)l.sy; pend[idx].seg[2] = (float)l.ex; pend[idx].seg[3] = (float)l.ey;
        pend[idx].tag = 0; pend[idx].seed = 0; pend[idx].npix = l.len;
      }
    }
    base += __popc(m);
  }
  const bool bad = s_bad != 0 || base > D.pend_cap;
  if (lane == 0) {
    if (bad) atomicOr(D.err, SDPL_ERR_OVERFLOW);
    D.npend[task] = bad ? 0 : base;
  }
}

}  // namespace sdpl
