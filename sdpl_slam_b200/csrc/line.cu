// line.cu -- sm_100a CUDA implementation of the SDPL-SLAM line front-end:
//   Lineextractor::operator()                 src/Lineextractor.cc:42-99
//   LSDDetectorC::ComputePyramid / detectImpl  3rdparty/line_descriptor/src/LSDDetector_custom.cpp:76-138, 254-369
//   cv::LineSegmentDetector::detect            (OpenCV imgproc/lsd.cpp, un-vendored; restated in oracle/lsd_oracle.cpp)
//   BinaryDescriptor::compute (LBD)            3rdparty/line_descriptor/src/binary_descriptor_custom.cpp:350-412,524-687,1026-1372
//
// Batched over frames (blockIdx carries the frame / task index).  Device layout for a chunk of B frames, nl octaves:
//   lvl    [B][sum_{l>=1} w_l*h_l]   u8   octave images (octave 0 is the caller's frame, never copied)          (L1)
//   scaled [B][sum_l sw_l*sh_l]      u8   sigma-0.75 blurred, 0.8x INTER_LINEAR_EXACT resized octaves            (L2)
//   px     [B][sum_l npx_l]          16 B level-line angle f64 + cos/sin f32 of every pixel                     (L3)
//   g2     [B][sum_l npx_l]          i32  gx^2+gy^2                                                             (L3)
//   state  [B][sum_l npx_l]          u32  used bit + speculative stamp                                          (L5)
//   order  [B][sum_l npx_l]          u32  defined pixels sorted by descending 10-bit gradient bin (stable)      (L4)
//   hist   [tasks][chunks][1024]     u32  counting-sort histograms / offsets                                    (L4)
//   reg    [tasks][2*npx_l]          i32  region lists (32 speculative lane segments + 1 sequential buffer)     (L5)
//   pend   [tasks][pend_cap]         rectangles in seed order -> accepted segments                              (L5/L6)
//   g      [B][sum_o w_o*h_o] u8, dx/dy [B][sum_o w_o*h_o] i16   LBD Gaussian octaves and Sobel derivatives     (L8)
// No tensor cores: integer / byte stencils (HBM-bound when batched), a latency-bound greedy stage, fp32/fp64 scalar math.
#include "lsd_grow.cuh"
#include "lsd_grow2.cuh"
#include <math.h>
#include <string.h>
#include <algorithm>

namespace sdpl {

constexpr int kMaxOct = 4;
constexpr int kBins = 1024;
constexpr int kSortChunk = 4096;        // pixels per warp in the counting sort
constexpr int kSortWarps = 4;           // warps per block in the counting sort

struct OctDev {
  int w, h;                 // octave image size
  int sw, sh, npx;          // LSD working size (0.8x) and pixel count
  int lw, lh;               // LBD octave size (w>>o, h>>o)
  int sc_stride; unsigned long long sc_off;   // LSD working image (0.8x): 16-byte row pitch, offset inside one frame's block
  int gstride;              // row pitch of the LBD Gaussian octave (multiple of 16 bytes: aligned word loads / stores)
  unsigned long long g_off; // its offset inside one frame's block of D.g
  unsigned long long lvl_off, px_off, lbd_off;   // element offsets inside one frame's block
  int nchunks; unsigned long long hist_off;       // per task (frame-independent) offsets inside one frame's hist block
  unsigned long long reg_off;                     // inside one frame's reg block (ints)
  int lane_cap;
  int area_fast;
  const unsigned short* xofs; const short2* xa; const unsigned short* yofs; const short2* ya;   // INTER_LINEAR tables (o>=1)
  const int* ex_ofs; const unsigned short* ex_c1; const int* ey_ofs; const unsigned short* ey_c1;  // INTER_LINEAR_EXACT tables
  double log_nt; int min_reg;
  float oct_scale;          // pow(scale, o)
};

struct LineDev {
  int nl, B;
  int in_w, in_h, in_stride; unsigned long long in_frame;
  const uint8_t* in;
  unsigned long long lvl_frame, px_frame, lbd_frame, g_frame, sc_frame, hist_frame, reg_frame;
  uint8_t *lvl, *scaled;
  lsd::PxRec* px; double* ang; int* g2; uint32_t* order;
  uint32_t* hist; int* maxg2; int* ndef; int* task_order;
  int* reg; lsd::Pending* pend; int pend_cap; int* npend;
  lsd::SlotCtx* ctx; int grow_ta;      // per-task seed-slot contexts of the two-phase schedule; phase-A expansion cap
  uint8_t* g; short2* sd;               // LBD: Gaussian octaves; Sobel (dx, dy) pairs of every pixel (one 4-byte gather per sample in k_lbd)
  int* err;
  long long* prof; int prof_detail;
  const double* lgam; int lgam_n;
  const double* nfa_tab;               // [nl][kNfaTabLevels][kNfaTabTri] (lsd::nfa_lookup), filled by k_nfa_table at set-up
  int* nbig; int* bigidx;
  double rho, prec, p, density_th, log_eps, scale;
  int g2_min;                          // smallest 4 * |gradient|^2 whose norm sqrt(g2 / 4.0) exceeds rho (host-computed, exact)
  int refine, serial_mode;
  double min_length;
  int nfeatures;
  OctDev O[kMaxOct];
  float gaussL[21], gaussG[63];
};

}  // namespace sdpl
#include "edlines.cuh"
namespace sdpl {
// ------------------------------------------------------------------------------------------------
// Shared-memory staging of an u8 image window as aligned 32-bit words, for the word / dp4a stencils below (L2, L8).
// Staged word m of a row holds the image columns xf + 4m - 4 .. xf + 4m - 1, whatever the alignment of the image rows in
// global memory (a frame of width 1242 has rows that are not word aligned): every staged word is cut from two aligned global
// words with a funnel shift.  Rows beyond the top / bottom edge are reflected (reflect-101) by index; columns beyond the left /
// right edge -- up to column xlast -- are patched in shared memory from their mirror columns.  [lo, hi) bounds the aligned words
// that overlap the caller's buffer: nothing outside is read.  One warp per row; warps = blockDim.x / 32.
// ------------------------------------------------------------------------------------------------
template <int NWORDS>
__device__ __forceinline__ void stage_tile(uint32_t (*s_in)[NWORDS], int nrows, int nwords, const uint8_t* img, int stride, int W, int H, int xf,
                                           int yf, int xlast, uintptr_t lo, uintptr_t hi) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  for (int r = warp; r < nrows; r += nwarp) {
    const int ry = reflect101(min(max(yf + r, -H + 1), 2 * H - 2), H);
    const uintptr_t a = (uintptr_t)(img + (size_t)ry * stride) + xf - 4;
    const int sh = (int)(a & 3) * 8;
    const uintptr_t a0 = a - (a & 3);
    for (int m = lane; m < nwords; m += 32) {
      const uintptr_t wa = a0 + 4 * (uintptr_t)m;
      const bool need = xf - 4 + 4 * m <= xlast;
      const uint32_t w0 = (need && wa >= lo && wa < hi) ? __ldg((const uint32_t*)wa) : 0u;
      const uint32_t w1 = (need && sh && wa + 4 >= lo && wa + 4 < hi) ? __ldg((const uint32_t*)(wa + 4)) : 0u;
      s_in[r][m] = __funnelshift_r(w0, w1, sh);
    }
    __syncwarp();
    if (lane == 0) {
      uint8_t* row = (uint8_t*)s_in[r] + 4;                   // row[c] = image column xf + c
      for (int c = -4; c < 0; c++) if (xf + c < 0) row[c] = row[-(xf + c) - xf];
      for (int xi = W; xi <= xlast && xi - xf < 4 * nwords - 4; xi++) row[xi - xf] = row[2 * (W - 1) - xi - xf];
    }
  }
}
// horizontal 5-tap pass (k0, k1, k2, k1, k0) over staged words: word m -> the four 16-bit row sums of its columns
template <int NWORDS>
__device__ __forceinline__ void hz5_tile(const uint32_t (*s_in)[NWORDS], uint2 (*s_hz)[NWORDS], int nrows, int nwords, uint32_t k0, uint32_t k1,
                                         uint32_t k2) {
  const uint32_t K_lo = k0 | (k1 << 8) | (k2 << 16) | (k1 << 24), K_hi = k0;
  for (int i = threadIdx.x; i < nrows * nwords; i += blockDim.x) {
    const int r = i / nwords, m = i - r * nwords;
    const uint32_t P = m > 0 ? s_in[r][m - 1] : 0u, C = s_in[r][m], N = m < nwords - 1 ? s_in[r][m + 1] : 0u;
    const uint32_t h0 = __dp4a(__funnelshift_r(P, C, 16), K_lo, __dp4a(__funnelshift_r(C, N, 16), K_hi, 0u));
    const uint32_t h1 = __dp4a(__funnelshift_r(P, C, 24), K_lo, __dp4a(__funnelshift_r(C, N, 24), K_hi, 0u));
    const uint32_t h2 = __dp4a(C, K_lo, __dp4a(N, K_hi, 0u));
    const uint32_t h3 = __dp4a(__funnelshift_r(C, N, 8), K_lo, __dp4a(N >> 8, K_hi, 0u));
    s_hz[r][m] = make_uint2(h0 | (h1 << 16), h2 | (h3 << 16));
  }
}
// vertical 5-tap pass on the row sums of rows gr .. gr + 4 of word m: four blurred bytes ((v + 2^15) >> 16; the weights add up to
// 2^16, so a result cannot exceed 255)
template <int NWORDS>
__device__ __forceinline__ uint32_t vt5_word(const uint2 (*s_hz)[NWORDS], int gr, int m, uint32_t k0, uint32_t k1, uint32_t k2) {
  const uint2 a = s_hz[gr][m], b = s_hz[gr + 1][m], c = s_hz[gr + 2][m], d = s_hz[gr + 3][m], e = s_hz[gr + 4][m];
  uint32_t v[4];
  v[0] = k0 * ((a.x & 0xffffu) + (e.x & 0xffffu)) + k1 * ((b.x & 0xffffu) + (d.x & 0xffffu)) + k2 * (c.x & 0xffffu) + (1u << 15);
  v[1] = k0 * ((a.x >> 16) + (e.x >> 16)) + k1 * ((b.x >> 16) + (d.x >> 16)) + k2 * (c.x >> 16) + (1u << 15);
  v[2] = k0 * ((a.y & 0xffffu) + (e.y & 0xffffu)) + k1 * ((b.y & 0xffffu) + (d.y & 0xffffu)) + k2 * (c.y & 0xffffu) + (1u << 15);
  v[3] = k0 * ((a.y >> 16) + (e.y >> 16)) + k1 * ((b.y >> 16) + (d.y >> 16)) + k2 * (c.y >> 16) + (1u << 15);
  const uint32_t lo2 = __byte_perm(v[0], v[1], 0x0062), hi2 = __byte_perm(v[2], v[3], 0x0062);
  return __byte_perm(lo2, hi2, 0x5410);
}

// ------------------------------------------------------------------------------------------------
// L1: octave o from octave o-1, cv::resize INTER_LINEAR (LSDDetectorC::ComputePyramid, LSDDetector_custom.cpp:76-109).
//     The 19-px reflect-101 border of the reference buffers is never read by LSD beyond what reflect-101 of the
//     interior gives (the LSD blur needs 3 px), so the octaves are stored un-padded.
// ------------------------------------------------------------------------------------------------
constexpr int kResizeRows = 8;
__global__ void __launch_bounds__(128) k_line_resize(LineDev D, int o) {
  const OctDev& Od = D.O[o];
  const OctDev& Os = D.O[o - 1];
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= Od.w) return;
  const uint8_t* src; int sstride;
  if (o == 1) { src = D.in + (size_t)blockIdx.z * D.in_frame; sstride = D.in_stride; }
  else { src = D.lvl + (size_t)blockIdx.z * D.lvl_frame + Os.lvl_off; sstride = Os.w; }
  uint8_t* dst = D.lvl + (size_t)blockIdx.z * D.lvl_frame + Od.lvl_off;
  // kResizeRows rows per thread (a one-row block is too short-lived to pay for its launch)
#pragma unroll 1
  for (int y = blockIdx.y * kResizeRows; y < min(Od.h, (int)(blockIdx.y + 1) * kResizeRows); y++) {
  int v;
  if (Od.area_fast) {
    const uint8_t* s0 = src + (size_t)(2 * y) * sstride;
    const uint8_t* s1 = s0 + sstride;
    v = (s0[2 * x] + s0[2 * x + 1] + s1[2 * x] + s1[2 * x + 1] + 2) >> 2;
  } else {
    const int sy = (short)Od.yofs[y];
    const short2 bb = Od.ya[y];
    const int sy0 = min(max(sy, 0), Os.h - 1), sy1 = min(max(sy + 1, 0), Os.h - 1);
    const uint8_t* s0 = src + (size_t)sy0 * sstride;
    const uint8_t* s1 = src + (size_t)sy1 * sstride;
    const int sx = Od.xofs[x];
    const short2 aa = Od.xa[x];
    const int sx1 = min(sx + 1, Os.w - 1);
    const int r0 = s0[sx] * aa.x + s0[sx1] * aa.y;
    const int r1 = s1[sx] * aa.x + s1[sx1] * aa.y;
    v = (((bb.x * (r0 >> 4)) >> 16) + ((bb.y * (r1 >> 4)) >> 16) + 2) >> 2;
    v = min(max(v, 0), 255);
  }
  dst[(size_t)y * Od.w + x] = (uint8_t)v;
  }
}

// ------------------------------------------------------------------------------------------------
// L2: LSD pre-scaling: GaussianBlur 7x7 sigma=0.6/0.8 (Q8.8 kernel [0,4,56,136,56,4,0], reflect-101) fused with
//     resize(fx=fy=0.8, INTER_LINEAR_EXACT) (Q8.8 coefficients).  One CTA = 64x16 output pixels; the source tile is
//     staged in shared memory, blurred there (horizontal Q8.8 -> u16, vertical Q16.16 -> u8), then down-sampled.
// ------------------------------------------------------------------------------------------------
constexpr int kST_W = 128, kST_H = 32;
constexpr int kSS_W = 168, kSS_H = 48;    // max source tile (incl. +-3 halo): 128/0.8+2+6, 32/0.8+2+6 (bounds checked at set-up)
constexpr int kSS_WW = 44;                // staged words per row: columns -4 .. 171 relative to the first blurred column
//     The source window is staged as aligned words (stage_tile), blurred with the dp4a row pass / 16-bit column pass of the
//     LBD kernel (the outer taps of the 7-tap kernel are zero: five taps 4, 56, 136, 56, 4), and a thread then interpolates four
//     adjacent output pixels and stores them as one word (the scaled plane has a 16-byte row pitch of its own).
__global__ void __launch_bounds__(256) k_lsd_scale(LineDev D, int o) {
  __shared__ uint32_t s_in[kSS_H][kSS_WW];
  __shared__ uint2 s_hz[kSS_H][kSS_WW];
  uint32_t (*s_bl)[kSS_WW] = s_in;          // the blurred window takes the place of the staged one (dead after the row pass)
  const OctDev& O = D.O[o];
  const int ox0 = blockIdx.x * kST_W, oy0 = blockIdx.y * kST_H;
  const int ox1 = min(ox0 + kST_W, O.sw) - 1, oy1 = min(oy0 + kST_H, O.sh) - 1;
  const uint8_t* src; int sstride; uintptr_t lo, hi;
  if (o == 0) {
    src = D.in + (size_t)blockIdx.z * D.in_frame; sstride = D.in_stride;
    lo = (uintptr_t)D.in & ~(uintptr_t)3;
    hi = ((uintptr_t)D.in + (size_t)(D.B - 1) * D.in_frame + (size_t)(O.h - 1) * D.in_stride + O.w + 3) & ~(uintptr_t)3;
  } else {
    src = D.lvl + (size_t)blockIdx.z * D.lvl_frame + O.lvl_off; sstride = O.w;
    lo = (uintptr_t)D.lvl; hi = (uintptr_t)D.lvl + (size_t)D.B * D.lvl_frame;        // own arena, 256-byte granules
  }
  // blurred source range needed by this tile
  const int bx0 = O.ex_ofs[ox0], bx1 = min(O.ex_ofs[ox1] + 1, O.w - 1);
  const int by0 = O.ey_ofs[oy0], by1 = min(O.ey_ofs[oy1] + 1, O.h - 1);
  const int bw = bx1 - bx0 + 1, bh = by1 - by0 + 1;      // <= 162 x 42
  constexpr int nwords = kSS_WW;                         // columns -4 .. 171 >= bw + 1 (words beyond the need are staged as zeros)
  const int nrows = bh + 4;                              // rows by0 - 2 .. by1 + 2
  stage_tile<kSS_WW>(s_in, nrows, nwords, src, sstride, O.w, O.h, bx0, by0 - 2, bx1 + 2, lo, hi);
  __syncthreads();
  hz5_tile<kSS_WW>(s_in, s_hz, nrows, nwords, 4u, 56u, 136u);
  __syncthreads();
  for (int i = threadIdx.x; i < bh * nwords; i += 256) {
    const int gr = i / nwords, m = i - gr * nwords;
    s_bl[gr][m] = vt5_word<kSS_WW>(s_hz, gr, m, 4u, 56u, 136u);
  }
  __syncthreads();
  // INTER_LINEAR_EXACT down-sampling of the blurred window; bl(y, x) = byte 4 + x of row y.  A thread owns four adjacent
  // output columns (their source offsets and weights are read once) and every eighth row of the tile
  const int xq = (threadIdx.x & 31) * 4;
  if (ox0 + xq >= O.sw) return;
  int sx[4], sx1[4]; uint32_t cx1[4];
#pragma unroll
  for (int e = 0; e < 4; e++) {
    const int ox = min(ox0 + xq + e, O.sw - 1);
    sx[e] = O.ex_ofs[ox] - bx0; cx1[e] = O.ex_c1[ox]; sx1[e] = min(sx[e] + 1, bw - 1);
  }
  uint8_t* dst = D.scaled + (size_t)blockIdx.z * D.sc_frame + O.sc_off;
#pragma unroll
  for (int k = 0; k < kST_H / 8; k++) {
    const int oy = oy0 + (threadIdx.x >> 5) + 8 * k;
    if (oy >= O.sh) break;
    const int sy = O.ey_ofs[oy] - by0;
    const uint32_t cy1 = O.ey_c1[oy];
    const int sy1 = min(sy + 1, bh - 1);
    const uint8_t* b0 = (const uint8_t*)s_bl[sy] + 4;
    const uint8_t* b1 = (const uint8_t*)s_bl[sy1] + 4;
    uint32_t out = 0;
#pragma unroll
    for (int e = 0; e < 4; e++) {
      const uint32_t r0 = (256u - cx1[e]) * b0[sx[e]] + cx1[e] * b0[sx1[e]];
      const uint32_t r1 = (256u - cx1[e]) * b1[sx[e]] + cx1[e] * b1[sx1[e]];
      const uint32_t v = (256u - cy1) * (r0 & 0xFFFFu) + cy1 * (r1 & 0xFFFFu);
      out |= min(255u, (v + (1u << 15)) >> 16) << (8 * e);
    }
    *(uint32_t*)(dst + (size_t)oy * O.sc_stride + ox0 + xq) = out;
  }
}

// ------------------------------------------------------------------------------------------------
// L3: ll_angle: 2x2 gradient, level-line angle, per-pixel cos/sin, gradient-magnitude maximum.
// ------------------------------------------------------------------------------------------------
constexpr int kGradTiles = 4;
__global__ void __launch_bounds__(256) k_lsd_grad(LineDev D, int o) {
  const OctDev& O = D.O[o];
  const int f = blockIdx.z;
  const int x = blockIdx.x * 64 + (threadIdx.x & 63);
  int m = 0;
  // kGradTiles 64 x 4 tiles per block, one below the other: a block of one tile lives a few thousand cycles, and its launch, the
  // block maximum, the barrier and the atomic were a visible share of that
#pragma unroll 1
  for (int t = 0; t < kGradTiles; t++) {
  const int y = (blockIdx.y * kGradTiles + t) * 4 + (threadIdx.x >> 6);
  int g2 = 0;
  if (x < O.sw && y < O.sh) {
    const uint8_t* img = D.scaled + (size_t)f * D.sc_frame + O.sc_off;
    const size_t q = (size_t)f * D.px_frame + O.px_off + (size_t)y * O.sw + x;
    lsd::PxRec a; a.state = 0u; a.deg = lsd::kNotDefDeg; a.c = 0.f; a.s = 0.f;
    double ang = lsd::kNotDef;
    if (x < O.sw - 1 && y < O.sh - 1) {
      const uint8_t* r0 = img + (size_t)y * O.sc_stride + x;
      const uint8_t* r1 = r0 + O.sc_stride;
      const int DA = (int)r1[1] - (int)r0[0], BC = (int)r0[1] - (int)r1[0];
      const int gx = DA + BC, gy = DA - BC;
      g2 = gx * gx + gy * gy;
      if (g2 >= D.g2_min) {           // norm = sqrt(g2 / 4.0) > rho: sqrt is monotonic, the integer bound is found on the host
        a.deg = fast_atan2_deg((float)gx, (float)-gy);
        ang = (double)a.deg * lsd::kDegToRad;
        const double af = (double)(float)ang;
        double sn_, cs_; sdpl_sincos(af, &sn_, &cs_);
        a.c = (float)cs_; a.s = (float)sn_;
      } else {
        g2 = -g2 - 1;           // undefined pixels carry a negative code: excluded from max / bins, modgrad unused
      }
    } else {
      g2 = -1;
    }
    D.px[q] = a;
    D.ang[q] = ang;
    D.g2[q] = g2;
  }
  m = max(m, g2);
  }
  // block maximum of the defined gradient magnitudes
#pragma unroll
  for (int s = 16; s; s >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, s));
  __shared__ int wm[8];
  if ((threadIdx.x & 31) == 0) wm[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = wm[0];
    for (int i = 1; i < 8; i++) t = max(t, wm[i]);
    if (t > 0) atomicMax(D.maxg2 + f * D.nl + o, t);
  }
}

// ------------------------------------------------------------------------------------------------
// L4: pseudo-ordering = stable counting sort of the defined pixels by descending bin, bin = int(norm*(1023/max_norm)).
//     Each warp owns a contiguous run of kSortChunk pixels (row-major), so "stable" == row-major inside a bin.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int lsd_bin(int g2, double bin_coef) { return (int)(sqrt((double)g2 / 4.0) * bin_coef); }

template <bool SCATTER>
__global__ void __launch_bounds__(kSortWarps * 32) k_lsd_sort(LineDev D, int o) {
  __shared__ uint32_t cnt[kSortWarps][kBins];
  const OctDev& O = D.O[o];
  const int f = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int chunk = blockIdx.x * kSortWarps + warp;
  const int task = f * D.nl + o;
  uint32_t* h = D.hist + (size_t)f * D.hist_frame + O.hist_off + (size_t)chunk * kBins;
  uint32_t* c = cnt[warp];
  if (chunk < O.nchunks) {
    for (int i = lane; i < kBins; i += 32) c[i] = SCATTER ? h[i] : 0u;
  }
  __syncwarp();
  if (chunk >= O.nchunks) return;
  const int mg = D.maxg2[task];
  const double max_grad = sqrt((double)mg / 4.0);
  const double bin_coef = mg > 0 ? (double)(kBins - 1) / max_grad : 0.0;
  const int* g2 = D.g2 + (size_t)f * D.px_frame + O.px_off;
  uint32_t* order = D.order + (size_t)f * D.px_frame + O.px_off;
  const int p0 = chunk * kSortChunk, p1 = min(p0 + kSortChunk, O.npx);
  // four groups of 32 pixels per trip: the four loads (and the square roots behind them) are in flight together; the ranking
  // steps then run group after group, which is what keeps the order stable (ncu on the one-group loop: 2400 cycles per trip,
  // a DRAM round trip in front of every dependent ranking step)
  for (int pb = p0; pb < p1; pb += 128) {
    int g[4];
#pragma unroll
    for (int u = 0; u < 4; u++) { const int p = pb + 32 * u + lane; g[u] = p < p1 ? g2[p] : -1; }
    int bq[4];
#pragma unroll
    for (int u = 0; u < 4; u++) bq[u] = g[u] >= 0 ? (kBins - 1 - lsd_bin(g[u], bin_coef)) : -1;      // descending bins
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int p = pb + 32 * u + lane;
      const bool def = g[u] >= 0;
      const int b = bq[u];
      if (!SCATTER) {                      // counting only: the order inside the group does not matter
        if (def) atomicAdd(&c[b], 1u);
        continue;
      }
      const uint32_t act = __ballot_sync(0xffffffffu, def);
      if (def) {
        const uint32_t peers = __match_any_sync(act, b);
        const int rank = __popc(peers & ((1u << lane) - 1u));
        const int leader = __ffs(peers) - 1;
        uint32_t base = 0;
        if (lane == leader) { base = c[b]; c[b] = base + __popc(peers); }
        base = __shfl_sync(peers, base, leader);
        if (SCATTER) order[base + rank] = (uint32_t)p;
      }
      __syncwarp();
    }
  }
  if (!SCATTER) {
    __syncwarp();
    for (int i = lane; i < kBins; i += 32) h[i] = c[i];
  }
}

// per task: turn the [chunk][bin] histogram into scatter offsets: offset(bin, chunk) = #(lower bin index) + #(same bin,
// earlier chunk).  One CTA of 1024 threads per task, thread == bin.
__global__ void __launch_bounds__(kBins) k_lsd_sort_scan(LineDev D) {
  __shared__ uint32_t tot[kBins];
  __shared__ uint32_t wsum[32];
  const int o = blockIdx.x, f = blockIdx.y, b = threadIdx.x;
  const OctDev& O = D.O[o];
  uint32_t* h = D.hist + (size_t)f * D.hist_frame + O.hist_off;
  uint32_t run = 0;
  for (int c = 0; c < O.nchunks; c++) { const uint32_t v = h[(size_t)c * kBins + b]; h[(size_t)c * kBins + b] = run; run += v; }
  // exclusive scan of the per-bin totals over bins
  uint32_t inc = run;
  const int lane = b & 31, w = b >> 5;
#pragma unroll
  for (int s = 1; s < 32; s <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, s); if (lane >= s) inc += t; }
  if (lane == 31) wsum[w] = inc;
  __syncthreads();
  if (w == 0) {
    uint32_t v = wsum[lane], iv = v;
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, iv, s); if (lane >= s) iv += t; }
    wsum[lane] = iv - v;
    if (lane == 31) D.ndef[f * D.nl + o] = (int)iv;
  }
  __syncthreads();
  tot[b] = wsum[w] + inc - run;
  __syncthreads();
  const uint32_t base = tot[b];
  for (int c = 0; c < O.nchunks; c++) h[(size_t)c * kBins + b] += base;
}

// ------------------------------------------------------------------------------------------------
// L5: region growing + rectangle fitting + refine (lsd_grow.cuh).  One warp (= one CTA) per (frame, octave).
// L6: NFA validation, one thread per pending rectangle.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void make_task(const LineDev& D, int f, int o, lsd::Task& T) {
  const OctDev& O = D.O[o];
  const size_t pb = (size_t)f * D.px_frame + O.px_off;
  const int task = f * D.nl + o;
  T.w = O.sw; T.h = O.sh; T.npx = O.npx;
  T.px.p = D.px + pb; T.ang = D.ang + pb; T.g2 = D.g2 + pb; T.state.p = D.px + pb; T.order = D.order + pb;
  T.ndef = D.ndef[task];
  T.reg_spec = D.reg + (size_t)f * D.reg_frame + O.reg_off;
  T.lane_cap = O.lane_cap;
  T.reg_serial = T.reg_spec + (size_t)32 * O.lane_cap;
  T.pend = D.pend + (size_t)task * D.pend_cap; T.pend_cap = D.pend_cap; T.npend = D.npend + task;
  T.prec = D.prec; T.p = D.p; T.log_nt = O.log_nt; T.density_th = D.density_th; T.log_eps = D.log_eps; T.scale = D.scale;
  T.min_reg = O.min_reg; T.refine = D.refine; T.err = D.err;
  T.prof = D.prof ? D.prof + (size_t)task * 16 : nullptr;
  T.prof_detail = D.prof_detail;
  T.lgam = D.lgam; T.lgam_n = D.lgam_n;
  T.nfa_tab = D.nfa_tab ? D.nfa_tab + (size_t)o * lsd::kNfaTabLevels * lsd::kNfaTabTri : nullptr;
}

// L5 scheduling.  A task's run time follows its number of candidate seeds, and tasks differ by 2x within a batch; CTAs are
// handed to the SMs in blockIdx order, so the longest tasks get the lowest indices (longest-processing-time-first) and the
// short ones fill the tail of the launch.
__global__ void k_lsd_task_rank(LineDev D, int ntask) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ntask) return;
  const int mine = D.ndef[t];
  int rank = 0;
  for (int u = 0; u < ntask; u++) {
    const int v = D.ndef[u];
    rank += (v > mine || (v == mine && u < t)) ? 1 : 0;
  }
  D.task_order[rank] = t;
}

__global__ void __launch_bounds__(32) k_lsd_grow(LineDev D) {
  __shared__ int sel[32];
  const int task = D.task_order[blockIdx.x], f = task / D.nl, o = task % D.nl;
  lsd::Task T;
  make_task(D, f, o, T);
  lsd::grow_task(T, D.serial_mode == 1, sel);                      // 3: 32-seed lock-step waves, 1: one seed at a time
}

// default schedule: waves of 32*NW seeds, one CTA of NW warps per task.  MINB = CTAs per SM the register budget is cut for.
template <int NW, int MINB>
__global__ void __launch_bounds__(32 * NW, MINB) k_lsd_grow_block(LineDev D) {
  __shared__ lsd::BlockShared S;
  const int task = D.task_order[blockIdx.x], f = task / D.nl, o = task % D.nl;
  lsd::Task T;
  make_task(D, f, o, T);
  lsd::grow_task_block(T, S);
}

// round-2 schedule (lsd_grow2.cuh): phase A per lane, phase B per warp from a CTA-wide queue
template <int NW, int MINB>
__global__ void __launch_bounds__(32 * NW, MINB) k_lsd_grow2(LineDev D) {
  __shared__ lsd::Block2Shared<NW> S;
  const int task = D.task_order[blockIdx.x], f = task / D.nl, o = task % D.nl;
  lsd::Task T;
  make_task(D, f, o, T);
  lsd::grow_task_block2<NW>(T, D.ctx + (size_t)task * lsd::kCtxSlots, D.grow_ta, S);
}

constexpr int kLgamN = 32768;
static double host_log_gamma(double x) {
  if (x > 15.0) return 0.918938533204673 + (x - 0.5) * std::log(x) - x + 0.5 * x * std::log(x * std::sinh(1 / x) + 1 / (810.0 * std::pow(x, 6.0)));
  static const double q[7] = {75122.6331530, 80916.6278952, 36308.2951477, 8687.24529705, 1168.92649479, 83.8676043424, 2.50662827511};
  double a = (x + 0.5) * std::log(x + 5.5) - (x + 5.5), b = 0;
  for (int n = 0; n < 7; ++n) { a -= std::log(x + (double)n); b += q[n] * std::pow(x, (double)n); }
  return a + std::log(b);
}

// the NFA table of lsd::nfa_lookup: entry (octave, halvings j, n, k) = nfa(n, k, p / 2^j), one thread per entry
__global__ void __launch_bounds__(128) k_nfa_table(LineDev D, double* __restrict__ tab) {
  const int o = blockIdx.z, j = blockIdx.y, e = blockIdx.x * 128 + threadIdx.x;
  if (e >= lsd::kNfaTabTri) return;
  int n = (int)((sqrt(8.0 * (double)e + 1.0) - 1.0) / 2.0);
  while (n * (n + 1) / 2 > e) --n;
  while ((n + 1) * (n + 2) / 2 <= e) ++n;
  const int k = e - n * (n + 1) / 2;
  lsd::Task T;
  T.log_nt = D.O[o].log_nt; T.lgam = D.lgam; T.lgam_n = D.lgam_n;
  double p = D.p;
  for (int i = 0; i < j; i++) p /= 2;
  tab[((size_t)o * lsd::kNfaTabLevels + j) * lsd::kNfaTabTri + e] = lsd::nfa(T, n, k, p);
}

// small rectangles: one thread each; rectangles whose scan visits more than kBigRect pixels are queued for the warp kernel
constexpr double kBigRect = 384.0;
template <int MINB>
__global__ void __launch_bounds__(64, MINB) k_lsd_nfa(LineDev D) {
  const int task = blockIdx.y;
  const int f = task / D.nl, o = task % D.nl;
  const int np = D.npend[task];
  if ((int)(blockIdx.x * blockDim.x) >= np) return;
  lsd::Task T;
  make_task(D, f, o, T);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < np; i += gridDim.x * blockDim.x) {
    if (D.refine >= 2 && lsd::rect_area_bound(T.pend[i].rec) > kBigRect) {
      const int k = atomicAdd(D.nbig + task, 1);
      D.bigidx[(size_t)task * D.pend_cap + k] = i;                           // big list grows from the front
    } else if (!lsd::validate_first(T, i)) {
      const int k = atomicAdd(D.nbig + (size_t)D.nl * D.B + task, 1);
      D.bigidx[(size_t)task * D.pend_cap + D.pend_cap - 1 - k] = i;          // retry list grows from the back
    }
  }
}

// second pass over the small rectangles whose first evaluation failed: up to 25 more evaluations each, five at a time
// (lsd::validate_rest_warp).  A warp wants a dozen rectangles to keep its six groups busy: the number of warps that work on a
// task follows the length of its list
constexpr int kRestBlocks = 16;
template <int MINB>
__global__ void __launch_bounds__(128, MINB) k_lsd_nfa_rest(LineDev D) {
  const int task = blockIdx.y;
  const int f = task / D.nl, o = task % D.nl;
  const int nf = D.nbig[(size_t)D.nl * D.B + task];
  const int nwarps = min(max((nf + 11) / 12, 1), (int)gridDim.x * 4);
  const int wg = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (nf == 0 || (int)blockIdx.x * 4 >= nwarps) return;
  lsd::Task T;
  make_task(D, f, o, T);
  if (wg < nwarps) lsd::validate_rest_warp(T, D.bigidx + (size_t)task * D.pend_cap + D.pend_cap - 1, nf, wg, nwarps);
}

__global__ void __launch_bounds__(128) k_lsd_nfa_big(LineDev D) {
  const int task = blockIdx.y;
  const int f = task / D.nl, o = task % D.nl;
  const int nb = D.nbig[task];
  if ((int)blockIdx.x * 4 >= nb) return;
  lsd::Task T;
  make_task(D, f, o, T);
  for (int k = blockIdx.x * 4 + (threadIdx.x >> 5); k < nb; k += gridDim.x * 4)
    lsd::validate_pending<true>(T, D.bigidx[(size_t)task * D.pend_cap + k]);
}

// ------------------------------------------------------------------------------------------------
// L7: KeyLine construction (LSDDetectorC::detectImpl tail, LSDDetector_custom.cpp:311-367; checkLineExtremes :112-138)
//     + optional top-N by response (Lineextractor.cc:73-82).  One CTA per frame; ordered compaction by block scan.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_keylines(LineDev D, sdpl_keyline* __restrict__ kls, int capacity, int* __restrict__ n_out,
                                                  sdpl_keyline* __restrict__ tmp) {
  __shared__ int wsum[8];
  __shared__ int sh_base;
  const int f = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // with a top-N filter the unfiltered list goes to tmp first
  const bool topn = D.nfeatures != 0;
  sdpl_keyline* out = topn ? tmp + (size_t)f * D.pend_cap * D.nl : kls + (size_t)f * capacity;
  const int out_cap = topn ? D.pend_cap * D.nl : capacity;
  if (tid == 0) sh_base = 0;
  __syncthreads();
  for (int o = 0; o < D.nl; o++) {
    const OctDev& O = D.O[o];
    const int task = f * D.nl + o;
    const lsd::Pending* P = D.pend + (size_t)task * D.pend_cap;
    const int np = D.npend[task];
    for (int i0 = 0; i0 < np; i0 += 256) {
      const int i = i0 + tid;
      bool keep = false;
      float e0 = 0, e1 = 0, e2 = 0, e3 = 0, length = 0;
      if (i < np && P[i].accepted) {
        e0 = P[i].seg[0]; e1 = P[i].seg[1]; e2 = P[i].seg[2]; e3 = P[i].seg[3];
        if (e0 < 0) e0 = 0;
        if (e0 >= O.w) e0 = (float)O.w - 1.0f;
        if (e2 < 0) e2 = 0;
        if (e2 >= O.w) e2 = (float)O.w - 1.0f;
        if (e1 < 0) e1 = 0;
        if (e1 >= O.h) e1 = (float)O.h - 1.0f;
        if (e3 < 0) e3 = 0;
        if (e3 >= O.h) e3 = (float)O.h - 1.0f;
        const float ddx = __fsub_rn(e0, e2), ddy = __fsub_rn(e1, e3);
        const double l = (double)(float)sqrt((double)ddx * (double)ddx + (double)ddy * (double)ddy);
        length = (float)l;
        keep = l > D.min_length;
      }
      const uint32_t m = __ballot_sync(0xffffffffu, keep);
      if (lane == 0) wsum[warp] = __popc(m);
      __syncthreads();
      int before = sh_base;
      for (int k = 0; k < warp; k++) before += wsum[k];
      if (keep) {
        const int idx = before + __popc(m & ((1u << lane) - 1u));
        if (idx < out_cap) {
          sdpl_keyline kl;
          kl.sx = __fmul_rn(e0, O.oct_scale); kl.sy = __fmul_rn(e1, O.oct_scale);
          kl.ex = __fmul_rn(e2, O.oct_scale); kl.ey = __fmul_rn(e3, O.oct_scale);
          kl.sx_oct = e0; kl.sy_oct = e1; kl.ex_oct = e2; kl.ey_oct = e3;
          kl.length = length;
          const int x0 = cv_round_f(e0), y0 = cv_round_f(e1), x1 = cv_round_f(e2), y1 = cv_round_f(e3);
          kl.num_pixels = max(abs(x1 - x0), abs(y1 - y0)) + 1;
          const float ay = __fsub_rn(kl.ey, kl.sy), ax = __fsub_rn(kl.ex, kl.sx);
          kl.angle = (float)atan2((double)ay, (double)ax);
          kl.class_id = idx;
          kl.octave = o;
          kl.size = __fmul_rn(ax, ay);
          kl.response = __fdiv_rn(length, (float)max(O.w, O.h));
          kl.pt_x = __fdiv_rn(__fadd_rn(kl.ex, kl.sx), 2.f);
          kl.pt_y = __fdiv_rn(__fadd_rn(kl.ey, kl.sy), 2.f);
          out[idx] = kl;
        }
      }
      __syncthreads();
      if (tid == 0) { int t = 0; for (int k = 0; k < 8; k++) t += wsum[k]; sh_base += t; }
      __syncthreads();
    }
  }
  const int n = sh_base;
  if (!topn || n <= D.nfeatures) {
    if (topn) {   // fewer lines than the cap: copy through unchanged
      for (int i = tid; i < min(n, capacity); i += 256) kls[(size_t)f * capacity + i] = out[i];
    }
    if (tid == 0) n_out[f] = n;   // n > capacity is reported by the host wrapper (SDPL_ERR_CAPACITY)
    return;
  }
  // top-N by response, descending, ties by original order (oracle decision; the reference uses unstable std::sort)
  for (int i = tid; i < min(n, out_cap); i += 256) {
    const float r = out[i].response;
    int rank = 0;
    for (int j = 0; j < min(n, out_cap); j++) {
      const float rj = out[j].response;
      rank += (rj > r) || (rj == r && j < i);
    }
    if (rank < D.nfeatures && rank < capacity) {
      sdpl_keyline kl = out[i];
      kl.class_id = rank;
      kls[(size_t)f * capacity + rank] = kl;
    }
  }
  if (tid == 0) n_out[f] = D.nfeatures;
}

// ------------------------------------------------------------------------------------------------
// L8: LBD inputs (BinaryDescriptor::computeGaussianPyramid / computeSobel, binary_descriptor_custom.cpp:350-398):
//     octave 0 = GaussianBlur 5x5 sigma 1 (Q8.8 [14,62,104,62,14]); octave k = pyrDown(prev, (w/2,h/2)); Sobel 3x3 -> s16.
//     Octave 0 is ONE kernel: a CTA stages a (128+8) x (32+6) window of the frame in shared memory as aligned 32-bit words
//     (the frame's rows are not word aligned -- 1242 is no multiple of 4 --, so every staged word is cut from two aligned global
//     words with a funnel shift; rows and the three columns beyond an image edge are reflected while staging), runs the
//     horizontal 5-tap pass as two dp4a per pixel on byte windows of (previous | own | next) word, the vertical pass on the
//     16-bit row sums, keeps the blurred tile (+1 pixel all round) in shared memory, writes its interior once (source of octave 1)
//     and takes the Sobel derivatives from it.  The blurred image mirrors about the image edges exactly as the frame does
//     (symmetric kernel, reflect-101 on both), so Sobel's own reflect-101 border is what the halo already holds.
// ------------------------------------------------------------------------------------------------
constexpr int kLT_W = 128, kLT_H = 32;          // output tile
constexpr int kLT_WW = kLT_W / 4 + 2;           // staged words per row: columns -4 .. kLT_W + 3
constexpr int kLT_IR = kLT_H + 6;               // staged rows: -3 .. kLT_H + 2
constexpr int kLT_GR = kLT_H + 2;               // blurred rows: -1 .. kLT_H
__global__ void __launch_bounds__(256) k_lbd_blur_sobel(LineDev D) {
  __shared__ uint32_t s_in[kLT_IR][kLT_WW];
  __shared__ uint2 s_hz[kLT_IR][kLT_WW];
  __shared__ uint32_t s_g[kLT_GR][kLT_WW];
  const int x0 = blockIdx.x * kLT_W, y0 = blockIdx.y * kLT_H, f = blockIdx.z;
  const int W = D.in_w, H = D.in_h;
  const OctDev& O = D.O[0];
  {
    const uintptr_t lo = (uintptr_t)D.in & ~(uintptr_t)3;
    const uintptr_t hi = ((uintptr_t)D.in + (size_t)(D.B - 1) * D.in_frame + (size_t)(H - 1) * D.in_stride + W + 3) & ~(uintptr_t)3;
    stage_tile<kLT_WW>(s_in, kLT_IR, kLT_WW, D.in + (size_t)f * D.in_frame, D.in_stride, W, H, x0, y0 - 3, min(x0 + kLT_W + 2, W + 2), lo, hi);
  }
  __syncthreads();
  hz5_tile<kLT_WW>(s_in, s_hz, kLT_IR, kLT_WW, 14u, 62u, 104u);       // word m = columns x0 + 4m - 4 ..
  __syncthreads();
  // vertical pass: blurred row gr = image row y0 - 1 + gr from staged rows gr .. gr + 4
  {
    uint8_t* gdst = D.g + (size_t)f * D.g_frame + O.g_off;
    for (int i = threadIdx.x; i < kLT_GR * kLT_WW; i += 256) {
      const int gr = i / kLT_WW, m = i - gr * kLT_WW;
      const uint32_t g4 = vt5_word<kLT_WW>(s_hz, gr, m, 14u, 62u, 104u);
      s_g[gr][m] = g4;
      const int y = y0 - 1 + gr, x = x0 + 4 * (m - 1);
      if (gr >= 1 && gr <= kLT_H && m >= 1 && m <= kLT_W / 4 && y < H && x < W) *(uint32_t*)(gdst + (size_t)y * O.gstride + x) = g4;
    }
  }
  __syncthreads();
  // Sobel 3x3 on the blurred tile: dx(c) = A(c+1) - A(c-1), A = r0 + 2 r1 + r2;  dy(c) = B(c-1) + 2 B(c) + B(c+1), B = r2 - r0
  {
    short2* sdo = D.sd + (size_t)f * D.lbd_frame + O.lbd_off;
    for (int i = threadIdx.x; i < kLT_H * (kLT_W / 4); i += 256) {
      const int r = i / (kLT_W / 4), m = 1 + (i - r * (kLT_W / 4));
      const int y = y0 + r, x = x0 + 4 * (m - 1);
      if (y >= H || x >= W) continue;
      int A[6], Bv[6];
#pragma unroll
      for (int k = 0; k < 6; k++) {
        // column x - 1 + k: byte 3 of word m-1, bytes 0..3 of word m, byte 0 of word m+1
        const int wi = k == 0 ? m - 1 : (k == 5 ? m + 1 : m), bi = k == 0 ? 3 : (k == 5 ? 0 : k - 1);
        const int t0 = (int)((s_g[r][wi] >> (8 * bi)) & 0xffu), t1 = (int)((s_g[r + 1][wi] >> (8 * bi)) & 0xffu),
                  t2 = (int)((s_g[r + 2][wi] >> (8 * bi)) & 0xffu);
        A[k] = t0 + 2 * t1 + t2; Bv[k] = t2 - t0;
      }
      short gx[4], gy[4];
#pragma unroll
      for (int e = 0; e < 4; e++) { gx[e] = (short)(A[e + 2] - A[e]); gy[e] = (short)(Bv[e] + 2 * Bv[e + 1] + Bv[e + 2]); }
      const size_t q = (size_t)y * W + x;
      uint32_t pk[4];
#pragma unroll
      for (int e = 0; e < 4; e++) pk[e] = (uint32_t)(uint16_t)gx[e] | ((uint32_t)(uint16_t)gy[e] << 16);
      if (x + 3 < W && ((q & 1) == 0)) {                      // 8-byte aligned: two pixel pairs
        *(uint2*)(sdo + q) = make_uint2(pk[0], pk[1]);
        *(uint2*)(sdo + q + 2) = make_uint2(pk[2], pk[3]);
      } else {
#pragma unroll
        for (int e = 0; e < 4; e++) if (x + e < W) *(uint32_t*)(sdo + q + e) = pk[e];
      }
    }
  }
}

constexpr int kLbdTiles = 4;
__global__ void __launch_bounds__(256) k_lbd_pyrdown(LineDev D, int o) {
  const OctDev& Od = D.O[o];
  const OctDev& Os = D.O[o - 1];
  const int x = blockIdx.x * 64 + (threadIdx.x & 63), f = blockIdx.z;
  if (x >= Od.lw) return;
  const uint8_t* src = D.g + (size_t)f * D.g_frame + Os.g_off;
  const int sw = Os.lw, sh = Os.lh, sp = Os.gstride;
  int cx[5];
#pragma unroll
  for (int k = 0; k < 5; k++) cx[k] = reflect101(2 * x + k - 2, sw);
  // kLbdTiles 64 x 4 tiles per block (one-tile blocks are too short-lived, see k_lsd_grad)
#pragma unroll 1
  for (int t = 0; t < kLbdTiles; t++) {
  const int y = (blockIdx.y * kLbdTiles + t) * 4 + (threadIdx.x >> 6);
  if (y >= Od.lh) break;
  int rows[5];
#pragma unroll
  for (int k = 0; k < 5; k++) {
    const uint8_t* s = src + (size_t)reflect101(2 * y + k - 2, sh) * sp;
    rows[k] = s[cx[0]] + s[cx[4]] + 4 * (s[cx[1]] + s[cx[3]]) + 6 * s[cx[2]];
  }
  const int v = rows[0] + rows[4] + 4 * (rows[1] + rows[3]) + 6 * rows[2];
  D.g[(size_t)f * D.g_frame + Od.g_off + (size_t)y * Od.gstride + x] = (uint8_t)((v + 128) >> 8);
  }
}

__global__ void __launch_bounds__(256) k_lbd_sobel(LineDev D, int o) {
  const OctDev& O = D.O[o];
  const int x = blockIdx.x * 64 + (threadIdx.x & 63), f = blockIdx.z;
  if (x >= O.lw) return;
  const int w = O.lw, h = O.lh;
  const uint8_t* src = D.g + (size_t)f * D.g_frame + O.g_off;
#pragma unroll 1
  for (int t = 0; t < kLbdTiles; t++) {
  const int y = (blockIdx.y * kLbdTiles + t) * 4 + (threadIdx.x >> 6);
  if (y >= h) break;
  const uint8_t* r0 = src + (size_t)reflect101(y - 1, h) * O.gstride;
  const uint8_t* r1 = src + (size_t)y * O.gstride;
  const uint8_t* r2 = src + (size_t)reflect101(y + 1, h) * O.gstride;
  const int xm = reflect101(x - 1, w), xp = reflect101(x + 1, w);
  const int gx = ((int)r0[xp] - r0[xm]) + 2 * ((int)r1[xp] - r1[xm]) + ((int)r2[xp] - r2[xm]);
  const int gy = ((int)r2[xm] + 2 * r2[x] + r2[xp]) - ((int)r0[xm] + 2 * r0[x] + r0[xp]);
  const size_t q = (size_t)f * D.lbd_frame + O.lbd_off + (size_t)y * w + x;
  D.sd[q] = make_short2((short)gx, (short)gy);
  }
}

// ------------------------------------------------------------------------------------------------
// L9: LBD descriptor (BinaryDescriptor::computeLBD :1026-1372, binaryConversion :401-412, combinations :74-107).
//     One CTA of 64 threads per line: thread h < 63 accumulates row h of the line support region sequentially (fp32
//     order preserved), thread b < 9 then accumulates band b over its rows in row order, thread 0 normalises,
//     32 threads emit the 32 descriptor bytes.
// ------------------------------------------------------------------------------------------------
__constant__ unsigned char c_comb[32][2] = {{0, 1}, {0, 2}, {0, 3}, {0, 4}, {0, 5}, {0, 6}, {1, 2}, {1, 3}, {1, 4}, {1, 5}, {1, 6},
                                            {2, 3}, {2, 4}, {2, 5}, {2, 6}, {2, 7}, {2, 8}, {3, 4}, {3, 5}, {3, 6}, {3, 7}, {3, 8},
                                            {4, 5}, {4, 6}, {4, 7}, {4, 8}, {5, 6}, {5, 7}, {5, 8}, {6, 7}, {6, 8}, {7, 8}};

template <int MINB>
__global__ void __launch_bounds__(64, MINB) k_lbd(LineDev D, const sdpl_keyline* __restrict__ kls, int capacity, const int* __restrict__ n_arr,
                                            uint8_t* __restrict__ desc, float* __restrict__ fdesc) {
  __shared__ float rows[63][8];
  __shared__ float des[72];
  const int f = blockIdx.y, tid = threadIdx.x;
  const int nlines = min(n_arr[f], capacity);
  for (int li = blockIdx.x; li < nlines; li += gridDim.x) {
  const sdpl_keyline kl = kls[(size_t)f * capacity + li];
  const int o = min(max(kl.octave, 0), D.nl - 1);
  const OctDev& O = D.O[o];
  const int realWidth = O.lw, realHeight = O.lh;
  const short2* sdImg = D.sd + (size_t)f * D.lbd_frame + O.lbd_off;
  const int NB = 9, WB = 7;
  const short imageWidth = (short)(realWidth - 1), imageHeight = (short)(realHeight - 1);
  const short lengthOfLSP = (short)kl.num_pixels;
  const short halfHeight = (short)((WB * NB - 1) / 2), halfWidth = (short)((lengthOfLSP - 1) / 2);
  const float midX = (float)(0.5 * (double)__fadd_rn(kl.sx_oct, kl.ex_oct)), midY = (float)(0.5 * (double)__fadd_rn(kl.sy_oct, kl.ey_oct));
  double dsn_, dcs_; sdpl_sincos((double)kl.angle, &dsn_, &dcs_);
  const float dL0 = (float)dcs_, dL1 = (float)dsn_;
  const float dO0 = -dL1, dO1 = dL0;
  if (tid < 63) {
    float t0 = __fmul_rn(-dL0, (float)halfWidth), t1 = __fmul_rn(dL1, (float)halfHeight);
    t0 = __fadd_rn(t0, t1);
    float sCorX0 = __fadd_rn(t0, midX);
    t0 = __fmul_rn(-dL1, (float)halfWidth); t1 = __fmul_rn(dL0, (float)halfHeight);
    t0 = __fsub_rn(t0, t1);
    float sCorY0 = __fadd_rn(t0, midY);
    for (int k = 0; k < tid; k++) { sCorX0 = __fsub_rn(sCorX0, dL1); sCorY0 = __fadd_rn(sCorY0, dL0); }
    float sCorX = sCorX0, sCorY = sCorY0;
    float pgdLRowSum = 0, ngdLRowSum = 0, pgdORowSum = 0, ngdORowSum = 0;
    for (short wID = 0; wID < lengthOfLSP; wID++) {
      short tempCor = (short)roundf(sCorX);
      const short xCor = (tempCor < 0) ? 0 : (tempCor > imageWidth) ? imageWidth : tempCor;
      tempCor = (short)roundf(sCorY);
      const short yCor = (tempCor < 0) ? 0 : (tempCor > imageHeight) ? imageHeight : tempCor;
      const int q = yCor * realWidth + xCor;
      const short2 sdv = __ldg(sdImg + q);
      const float dx = (float)sdv.x, dy = (float)sdv.y;
      const float gDL = __fadd_rn(__fmul_rn(dx, dL0), __fmul_rn(dy, dL1));
      const float gDO = __fadd_rn(__fmul_rn(dx, dO0), __fmul_rn(dy, dO1));
      if (gDL > 0) pgdLRowSum = __fadd_rn(pgdLRowSum, gDL); else ngdLRowSum = __fsub_rn(ngdLRowSum, gDL);
      if (gDO > 0) pgdORowSum = __fadd_rn(pgdORowSum, gDO); else ngdORowSum = __fsub_rn(ngdORowSum, gDO);
      sCorX = __fadd_rn(sCorX, dL0);
      sCorY = __fadd_rn(sCorY, dL1);
    }
    const float coef = D.gaussG[tid];
    pgdLRowSum = __fmul_rn(coef, pgdLRowSum); ngdLRowSum = __fmul_rn(coef, ngdLRowSum);
    pgdORowSum = __fmul_rn(coef, pgdORowSum); ngdORowSum = __fmul_rn(coef, ngdORowSum);
    rows[tid][0] = pgdLRowSum; rows[tid][1] = ngdLRowSum;
    rows[tid][2] = __fmul_rn(pgdLRowSum, pgdLRowSum); rows[tid][3] = __fmul_rn(ngdLRowSum, ngdLRowSum);
    rows[tid][4] = pgdORowSum; rows[tid][5] = ngdORowSum;
    rows[tid][6] = __fmul_rn(pgdORowSum, pgdORowSum); rows[tid][7] = __fmul_rn(ngdORowSum, ngdORowSum);
  }
  __syncthreads();
  if (tid < NB) {
    const int b = tid;
    float s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const int h0 = max(0, (b - 1) * WB), h1 = min(NB * WB, (b + 2) * WB);
    for (int h = h0; h < h1; h++) {
      const int hb = h / WB, r = h % WB;
      // row h adds to its own band with gaussL[r+7], to band hb-1 with gaussL[r+14], to band hb+1 with gaussL[r]
      const float c = hb == b ? D.gaussL[r + WB] : (hb == b + 1 ? D.gaussL[r + 2 * WB] : D.gaussL[r]);
      const float c2 = __fmul_rn(c, c);
      s[0] = __fadd_rn(s[0], __fmul_rn(c, rows[h][0])); s[1] = __fadd_rn(s[1], __fmul_rn(c, rows[h][1]));
      s[2] = __fadd_rn(s[2], __fmul_rn(c2, rows[h][2])); s[3] = __fadd_rn(s[3], __fmul_rn(c2, rows[h][3]));
      s[4] = __fadd_rn(s[4], __fmul_rn(c, rows[h][4])); s[5] = __fadd_rn(s[5], __fmul_rn(c, rows[h][5]));
      s[6] = __fadd_rn(s[6], __fmul_rn(c2, rows[h][6])); s[7] = __fadd_rn(s[7], __fmul_rn(c2, rows[h][7]));
    }
    const float invN = (b == 0 || b == NB - 1) ? (float)(1.0 / (WB * 2.0)) : (float)(1.0 / (WB * 3.0));
    float* d = des + b * 8;
    float t;
    t = __fmul_rn(s[0], invN); d[0] = t; d[4] = sqrtf(__fsub_rn(__fmul_rn(s[2], invN), __fmul_rn(t, t)));
    t = __fmul_rn(s[1], invN); d[1] = t; d[5] = sqrtf(__fsub_rn(__fmul_rn(s[3], invN), __fmul_rn(t, t)));
    t = __fmul_rn(s[4], invN); d[2] = t; d[6] = sqrtf(__fsub_rn(__fmul_rn(s[6], invN), __fmul_rn(t, t)));
    t = __fmul_rn(s[5], invN); d[3] = t; d[7] = sqrtf(__fsub_rn(__fmul_rn(s[7], invN), __fmul_rn(t, t)));
  }
  __syncthreads();
  if (tid == 0) {
    float tempM = 0, tempS = 0;
    for (int base = 0; base < NB; base++) {
      const float* v = des + base * 8;
      for (int k = 0; k < 4; k++) tempM = __fadd_rn(tempM, __fmul_rn(v[k], v[k]));
      for (int k = 4; k < 8; k++) tempS = __fadd_rn(tempS, __fmul_rn(v[k], v[k]));
    }
    tempM = __fdiv_rn(1.f, sqrtf(tempM));
    tempS = __fdiv_rn(1.f, sqrtf(tempS));
    for (int base = 0; base < NB; base++) {
      float* v = des + base * 8;
      for (int k = 0; k < 4; k++) v[k] = __fmul_rn(v[k], tempM);
      for (int k = 4; k < 8; k++) v[k] = __fmul_rn(v[k], tempS);
    }
    for (int i = 0; i < NB * 8; i++) if ((double)des[i] > 0.4) des[i] = (float)0.4;
    float temp = 0;
    for (int i = 0; i < NB * 8; i++) temp = __fadd_rn(temp, __fmul_rn(des[i], des[i]));
    temp = __fdiv_rn(1.f, sqrtf(temp));
    for (int i = 0; i < NB * 8; i++) des[i] = __fmul_rn(des[i], temp);
  }
  __syncthreads();
  if (fdesc) for (int i = tid; i < 72; i += 64) fdesc[((size_t)f * capacity + li) * 72 + i] = des[i];
  if (tid < 32) {
    const float* f1 = des + 8 * c_comb[tid][0];
    const float* f2 = des + 8 * c_comb[tid][1];
    unsigned r = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) r |= (unsigned)(f1[k] > f2[k]) << k;
    desc[((size_t)f * capacity + li) * 32 + tid] = (uint8_t)r;
  }
  __syncthreads();
  }
}

}  // namespace sdpl

// =====================================================================================================
// Host side
// =====================================================================================================
using namespace sdpl;

struct sdpl_line {
  int nfeatures, refine, nlevels, extractor, device;
  float lsd_scale, scale;
  std::vector<float> sf, isf;
  cudaStream_t own_stream = nullptr, stream = nullptr;
  cudaStream_t side = nullptr;                 // the warp-per-rectangle NFA kernel runs here, beside the per-thread second pass
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  int gw = 0, gh = 0, gB = 0;
  LineDev D;
  DevBuf lvl, scaled, px, ang, g2, order, hist, maxg2, ndef, torder, reg, pend, npend, g, sd, err, tables, tmpkl;
  DevBuf in_stage, out_kls, out_desc, out_n, prof, lgam, nbig, bigidx, ctx, nfatab;
  DevBuf ed_img, ed_work, ed_nanch, ed_tab;    // EDLines back-end (extractor == 1): images, per-task scratch, anchor counts, host tables
  EdDev E;
  int ed_w = 0, ed_h = 0;
  int grow_warps = 0 /* auto */, sm_count = 148, grow_minb = 0 /* auto */;
  int grow_warps_small = 16, grow_ta_small = -1;  // small batches (tasks <= SMs): warps per task, phase-A cap (-1: grow_ta)
  int grow_forced = 0;   // SDPL_GROW given: use it for big batches only (single frames keep the 8-warp variant)
  int grow_smem = 0;     // tuning: dynamic shared memory requested by the <4,5> variant (caps its CTAs per SM without touching registers)
  int prof_detail = 0;   // sdpl_line_debug_grow_detail: thread 0 of every task accumulates prof[8..15] (costs ~3 % of the grow kernel)
  int nfa_minb = 10;      // second NFA pass compiled for 10 CTAs per SM (48 registers): 6.11 ms at 16, 5.33 at 12, 5.24 at 10, 5.37 at 8 (512 frames)
  int grow_legacy = 0, grow_ta = 16;   // phase-A cap: 16 measured best at 512 frames (4: 51.2, 8: 49.9, 16: 47.6 ms)
  void* h_stage = nullptr; size_t h_stage_bytes = 0;
  int pend_cap = 4096;
  int last_B = 0, launches = 0, serial_mode = 0;
  StageTimer timer;
};

namespace {
struct ExactCoef { int ofs; unsigned short c1; };
// coefficients of cv::resize INTER_LINEAR_EXACT (oracle/cvprim.cpp exact_coeffs); edge cases folded into (ofs, c1):
// left of the first source sample -> (0, 0); at / beyond the last -> (src-1, 0)
void exact_coeffs(double inv_scale, int srcsize, int dstsize, std::vector<ExactCoef>& out) {
  volatile double scale = 1.0 / inv_scale;
  out.resize(dstsize);
  for (int val = 0; val < dstsize; val++) {
    volatile double t = scale * ((double)val + 0.5);
    const double fval = t - 0.5;
    const int ival = h_cv_floor(fval);
    out[val].ofs = 0; out[val].c1 = 0;
    if (ival >= 0 && srcsize > 1) {
      if (ival < srcsize - 1) {
        out[val].ofs = ival;
        const double fr = fval - (double)ival;
        out[val].c1 = (unsigned short)(fr < 0 ? 0 : h_cv_round(fr * 256.0));
      } else {
        out[val].ofs = srcsize - 1;
      }
    }
  }
}
}  // namespace

static int line_setup(sdpl_line* o, int w, int h, int B) {
  if (o->gw == w && o->gh == h && o->gB >= B) { o->D.B = B; return SDPL_OK; }
  if (o->gw == w && o->gh == h) B = std::max(B, o->gB);
  // a new geometry or a bigger batch rewrites tables, arenas and the overflow flag: nothing of this handle may be in flight
  if (o->gw) cudaStreamSynchronize(o->stream);
  const int nl = o->nlevels;
  LineDev& D = o->D;
  memset(&D, 0, sizeof(D));
  D.nl = nl;
  // LSD constants, Lineextractor.cc:54-70
  const double ang_th = 22.5, quant = 2.0;
  D.prec = lsd::kPI * ang_th / 180; D.p = ang_th / 180; D.rho = quant / sin(D.prec);
  {
    int g = std::max(0, (int)(4.0 * D.rho * D.rho) - 4);
    while (!(sqrt((double)g / 4.0) > D.rho)) g++;
    D.g2_min = g;
  }
  D.density_th = 0.8; D.log_eps = 0.0; D.scale = (double)o->lsd_scale; D.refine = o->refine; D.serial_mode = o->serial_mode;
  D.min_length = 0.02 * (double)std::min(w, h);
  D.nfeatures = o->nfeatures;
  {  // BinaryDescriptor::BinaryDescriptor weights, binary_descriptor_custom.cpp:217-259 (integer divisions as written there)
    const int WB = 7, NB = 9;
    double u = (WB * 3 - 1) / 2, sigma = (WB * 2 + 1) / 2, inv = -1 / (2 * sigma * sigma);
    for (int i = 0; i < WB * 3; i++) { const double d = i - u; D.gaussL[i] = (float)exp(d * d * inv); }
    u = (NB * WB - 1) / 2; sigma = u; inv = -1 / (2 * sigma * sigma);
    for (int i = 0; i < NB * WB; i++) { const double d = i - u; D.gaussG[i] = (float)exp(d * d * inv); }
  }
  std::vector<unsigned short> t_u16; std::vector<short> t_s16; std::vector<int> t_i32;
  std::vector<size_t> xofs_at(nl), yofs_at(nl), xa_at(nl), ya_at(nl), exo_at(nl), eyo_at(nl), exc_at(nl), eyc_at(nl);
  size_t lvl_off = 0, px_off = 0, lbd_off = 0, g_off = 0, sc_off = 0, hist_off = 0, reg_off = 0;
  int lw = w, lh = h;
  for (int l = 0; l < nl; l++) {
    OctDev& O = D.O[l];
    O.w = h_cv_round((float)w * o->isf[l]);
    O.h = h_cv_round((float)h * o->isf[l]);
    if (O.w < 8 || O.h < 8 || O.w > 16000 || O.h > 16000) { set_last_error("line pyramid octave size out of range"); return SDPL_ERR_ARG; }
    O.sw = h_cv_round((double)O.w * D.scale); O.sh = h_cv_round((double)O.h * D.scale);
    if (O.sw < 2 || O.sh < 2) { set_last_error("LSD working image too small"); return SDPL_ERR_ARG; }
    O.npx = O.sw * O.sh;
    if (l > 0) { lw /= 2; lh /= 2; }
    O.lw = lw; O.lh = lh;
    if (lw < 1 || lh < 1 || lw > 32767 || lh > 32767) { set_last_error("LBD octave size out of range"); return SDPL_ERR_ARG; }
    O.lvl_off = lvl_off; if (l > 0) lvl_off += align_up((size_t)O.w * O.h, 16);
    O.px_off = px_off; px_off += align_up((size_t)O.npx, 16);
    O.sc_stride = (int)align_up((size_t)O.sw, 16); O.sc_off = sc_off; sc_off += (size_t)O.sc_stride * O.sh;
    O.lbd_off = lbd_off; lbd_off += align_up((size_t)lw * lh, 16);
    O.gstride = (int)align_up((size_t)lw, 16); O.g_off = g_off; g_off += (size_t)O.gstride * lh;
    O.nchunks = div_up(O.npx, kSortChunk);
    O.hist_off = hist_off; hist_off += (size_t)O.nchunks * kBins;
    O.lane_cap = 4096;                                   // per-lane list ring, power of two >= npx/32
    while (O.lane_cap < O.npx / 32) O.lane_cap *= 2;
    O.reg_off = reg_off; reg_off += (size_t)32 * O.lane_cap + O.npx;
    O.log_nt = 5 * (log10((double)O.sw) + log10((double)O.sh)) / 2 + log10(11.0);
    O.min_reg = (int)(size_t)(-O.log_nt / log10(D.p));
    O.oct_scale = (float)pow((double)o->scale, (double)l);     // pow(float(scale), octaveIdx)
    // INTER_LINEAR tables octave l-1 -> l (same arithmetic as the ORB pyramid)
    O.area_fast = 0;
    if (l > 0) {
      const int sw = D.O[l - 1].w, sh = D.O[l - 1].h;
      const double scale_x = 1. / ((double)O.w / sw), scale_y = 1. / ((double)O.h / sh);
      const int isx = h_cv_round(scale_x), isy = h_cv_round(scale_y);
      if (fabs(scale_x - isx) < 2.220446049250313e-16 && fabs(scale_y - isy) < 2.220446049250313e-16 && isx == 2 && isy == 2) O.area_fast = 1;
      xofs_at[l] = t_u16.size(); xa_at[l] = t_s16.size();
      for (int dx = 0; dx < O.w; dx++) {
        float fx = (float)((dx + 0.5) * scale_x - 0.5);
        int sx = h_cv_floor(fx);
        fx -= sx;
        if (sx < 0) { fx = 0; sx = 0; }
        if (sx >= sw - 1) { fx = 0; sx = sw - 1; }
        t_u16.push_back((unsigned short)sx);
        t_s16.push_back((short)std::min(32767, h_cv_round((1.f - fx) * 2048)));
        t_s16.push_back((short)std::min(32767, h_cv_round(fx * 2048)));
      }
      yofs_at[l] = t_u16.size(); ya_at[l] = t_s16.size();
      for (int dy = 0; dy < O.h; dy++) {
        float fy = (float)((dy + 0.5) * scale_y - 0.5);
        const int sy = h_cv_floor(fy);
        fy -= sy;
        t_u16.push_back((unsigned short)(short)sy);
        t_s16.push_back((short)std::min(32767, h_cv_round((1.f - fy) * 2048)));
        t_s16.push_back((short)std::min(32767, h_cv_round(fy * 2048)));
      }
    }
    // INTER_LINEAR_EXACT tables octave -> 0.8x
    std::vector<ExactCoef> cx, cy;
    exact_coeffs(D.scale, O.w, O.sw, cx);
    exact_coeffs(D.scale, O.h, O.sh, cy);
    exo_at[l] = t_i32.size(); for (auto& c : cx) t_i32.push_back(c.ofs);
    eyo_at[l] = t_i32.size(); for (auto& c : cy) t_i32.push_back(c.ofs);
    exc_at[l] = t_u16.size(); for (auto& c : cx) t_u16.push_back(c.c1);
    eyc_at[l] = t_u16.size(); for (auto& c : cy) t_u16.push_back(c.c1);
    // the shared-memory tile of k_lsd_scale must cover the source span of a 64x16 output tile
    for (int x0 = 0; x0 < O.sw; x0 += kST_W) {
      const int x1 = std::min(x0 + kST_W, O.sw) - 1;
      if (std::min(cx[x1].ofs + 1, O.w - 1) - cx[x0].ofs + 1 + 6 > kSS_W) { set_last_error("unsupported LSD scale (tile span)"); return SDPL_ERR_UNSUPPORTED; }
    }
    for (int y0 = 0; y0 < O.sh; y0 += kST_H) {
      const int y1 = std::min(y0 + kST_H, O.sh) - 1;
      if (std::min(cy[y1].ofs + 1, O.h - 1) - cy[y0].ofs + 1 + 6 > kSS_H) { set_last_error("unsupported LSD scale (tile span)"); return SDPL_ERR_UNSUPPORTED; }
    }
  }
  D.lvl_frame = align_up(std::max<size_t>(lvl_off, 16), 256); D.px_frame = align_up(px_off, 256); D.lbd_frame = align_up(lbd_off, 256); D.g_frame = align_up(g_off, 256); D.sc_frame = align_up(sc_off, 256);
  D.hist_frame = hist_off; D.reg_frame = reg_off;
  // rectangles per (frame, octave): scales with the working image (1242x375 -> 6400; an overflow is reported, never silent)
  o->pend_cap = std::max(4096, (int)align_up((size_t)D.O[0].npx / 48, 256));
  D.pend_cap = o->pend_cap;
  D.in_w = w; D.in_h = h;
  int rc;
  if ((rc = o->lvl.reserve(D.lvl_frame * B))) return rc;
  if ((rc = o->scaled.reserve(D.sc_frame * B))) return rc;
  if ((rc = o->px.reserve(sizeof(lsd::PxRec) * D.px_frame * B))) return rc;
  if ((rc = o->ang.reserve(sizeof(double) * D.px_frame * B))) return rc;
  if ((rc = o->g2.reserve(sizeof(int) * D.px_frame * B))) return rc;
  if ((rc = o->order.reserve(sizeof(uint32_t) * D.px_frame * B))) return rc;
  if ((rc = o->hist.reserve(sizeof(uint32_t) * D.hist_frame * B))) return rc;
  if ((rc = o->maxg2.reserve(sizeof(int) * nl * B))) return rc;
  if ((rc = o->ndef.reserve(sizeof(int) * nl * B))) return rc;
  if ((rc = o->torder.reserve(sizeof(int) * nl * B))) return rc;
  if ((rc = o->reg.reserve(sizeof(int) * D.reg_frame * B))) return rc;
  if ((rc = o->pend.reserve(sizeof(lsd::Pending) * (size_t)D.pend_cap * nl * B))) return rc;
  if ((rc = o->npend.reserve(sizeof(int) * nl * B))) return rc;
  if ((rc = o->g.reserve(D.g_frame * B))) return rc;
  if ((rc = o->sd.reserve(sizeof(short2) * D.lbd_frame * B))) return rc;
  if ((rc = o->tmpkl.reserve(sizeof(sdpl_keyline) * (size_t)D.pend_cap * nl * (o->nfeatures ? B : 1)))) return rc;
  if ((rc = o->err.reserve(sizeof(int)))) return rc;
  if ((rc = o->prof.reserve(sizeof(long long) * 16 * nl * B))) return rc;
  if ((rc = o->nbig.reserve(sizeof(int) * 2 * nl * B))) return rc;
  if ((rc = o->ctx.reserve(sizeof(lsd::SlotCtx) * (size_t)lsd::kCtxSlots * nl * B))) return rc;
  if ((rc = o->bigidx.reserve(sizeof(int) * (size_t)D.pend_cap * nl * B))) return rc;
  if (!o->lgam.p) {
    if ((rc = o->lgam.reserve(sizeof(double) * kLgamN))) return rc;
    // log_gamma(i) as the reference's NFA evaluates it (Lanczos / Windschitl formulas, descriptor_custom.hpp:687-760 show the
    // same ones), tabulated once per handle ON THE HOST: the table then holds exactly the doubles the C library gives the CPU
    // path (CUDA's log / pow / sinh differ from glibc's in the last bit for some arguments)
    std::vector<double> tab(kLgamN, 0.0);
    for (int i = 1; i < kLgamN; i++) tab[i] = host_log_gamma((double)i);
    SDPL_CUDA(cudaMemcpy(o->lgam.p, tab.data(), sizeof(double) * kLgamN, cudaMemcpyHostToDevice));
  }
  const size_t u16_bytes = align_up(t_u16.size() * 2 + 2, 16), s16_bytes = align_up(t_s16.size() * 2 + 2, 16);
  if ((rc = o->tables.reserve(u16_bytes + s16_bytes + t_i32.size() * 4 + 16))) return rc;
  char* tb = (char*)o->tables.p;
  if (!t_u16.empty()) SDPL_CUDA(cudaMemcpy(tb, t_u16.data(), t_u16.size() * 2, cudaMemcpyHostToDevice));
  if (!t_s16.empty()) SDPL_CUDA(cudaMemcpy(tb + u16_bytes, t_s16.data(), t_s16.size() * 2, cudaMemcpyHostToDevice));
  if (!t_i32.empty()) SDPL_CUDA(cudaMemcpy(tb + u16_bytes + s16_bytes, t_i32.data(), t_i32.size() * 4, cudaMemcpyHostToDevice));
  SDPL_CUDA(cudaMemset(o->err.p, 0, sizeof(int)));
  for (int l = 0; l < nl; l++) {
    OctDev& O = D.O[l];
    const unsigned short* u16 = (const unsigned short*)tb;
    const short* s16 = (const short*)(tb + u16_bytes);
    const int* i32 = (const int*)(tb + u16_bytes + s16_bytes);
    if (l > 0) {
      O.xofs = u16 + xofs_at[l]; O.yofs = u16 + yofs_at[l];
      O.xa = (const short2*)(s16 + xa_at[l]); O.ya = (const short2*)(s16 + ya_at[l]);
    }
    O.ex_ofs = i32 + exo_at[l]; O.ey_ofs = i32 + eyo_at[l];
    O.ex_c1 = u16 + exc_at[l]; O.ey_c1 = u16 + eyc_at[l];
  }
  D.lvl = o->lvl.as<uint8_t>(); D.scaled = o->scaled.as<uint8_t>(); D.px = o->px.as<lsd::PxRec>(); D.ang = o->ang.as<double>(); D.g2 = o->g2.as<int>();
  D.order = o->order.as<uint32_t>(); D.hist = o->hist.as<uint32_t>();
  D.maxg2 = o->maxg2.as<int>(); D.ndef = o->ndef.as<int>(); D.task_order = o->torder.as<int>(); D.reg = o->reg.as<int>(); D.pend = o->pend.as<lsd::Pending>();
  D.npend = o->npend.as<int>(); D.g = o->g.as<uint8_t>(); D.sd = o->sd.as<short2>();
  D.err = o->err.as<int>(); D.prof = o->prof.as<long long>(); D.lgam = o->lgam.as<double>(); D.lgam_n = kLgamN;
  D.ctx = o->ctx.as<lsd::SlotCtx>(); D.grow_ta = o->grow_ta;
  D.nbig = o->nbig.as<int>(); D.bigidx = o->bigidx.as<int>();
  D.B = B;
  {
    // NFA values of every small rectangle this geometry can produce (lsd::nfa_lookup); asynchronous on the handle's stream
    if ((rc = o->nfatab.reserve(sizeof(double) * (size_t)nl * lsd::kNfaTabLevels * lsd::kNfaTabTri))) return rc;
    D.nfa_tab = nullptr;
    k_nfa_table<<<dim3(div_up(lsd::kNfaTabTri, 128), lsd::kNfaTabLevels, nl), 128, 0, o->stream>>>(D, o->nfatab.as<double>());
    SDPL_LAUNCH_CHECK();
    SDPL_CUDA(cudaStreamSynchronize(o->stream));      // set-up is the rare path; the caller may switch streams before the next call
    D.nfa_tab = o->nfatab.as<double>();
  }
  o->gw = w; o->gh = h; o->gB = B;
  return SDPL_OK;
}

// LSD detection + KeyLine construction for B frames resident on the device (asynchronous on o->stream)
// EDLines back-end: buffers and host tables for B frames of the current geometry (after line_setup)
static int ed_setup(sdpl_line* o, int B) {
  LineDev& D = o->D;
  EdDev& E = o->E;
  const int nl = o->nlevels;
  size_t img_off = 0, work_off = 0;
  for (int l = 0; l < nl; l++) {
    EdOct& O = E.O[l];
    O.w = D.O[l].w; O.h = D.O[l].h; O.npx = O.w * O.h;
    if (O.w < 8 || O.h < 8 || O.w > 32767 || O.h > 32767) { set_last_error("EDLines: image size out of range"); return SDPL_ERR_UNSUPPORTED; }
    O.anchors_cap = O.npx / 4 + 1024; O.pixels_cap = O.npx / 4 + 1024; O.stack_cap = O.npx / 16 + 1024; O.chains_cap = O.npx / 16 + 1024;
    O.chain_nos_cap = O.npx / 16 + 1024; O.seg_px_cap = O.npx / 2 + 1024; O.seg_cap = O.npx / 64 + 256; O.lines_cap = O.npx / 40 + 256;
    O.min_line_len = sdpl_ed::host::min_line_len(O.w, O.h);
    O.img_off = img_off; img_off += (ed_img_bytes(O.npx) + 255) & ~(size_t)255;
    O.work_off = work_off; work_off += (ed_work_bytes(O) + 255) & ~(size_t)255;
  }
  E.img_frame = img_off; E.work_frame = work_off;
  int rc;
  if ((rc = o->ed_img.reserve(E.img_frame * B))) return rc;
  if ((rc = o->ed_work.reserve(E.work_frame * B))) return rc;
  if ((rc = o->ed_nanch.reserve(sizeof(int) * 2 * nl * B))) return rc;
  if ((rc = o->ed_tab.reserve(sizeof(double) * (sdpl_ed::kAtanLut + 1) + sizeof(int) * (size_t)kEdNfaN * kMaxOct))) return rc;
  E.img = o->ed_img.as<uint8_t>(); E.work = o->ed_work.as<uint8_t>(); E.n_anchors = o->ed_nanch.as<int>(); E.nseg = E.n_anchors + (size_t)nl * B;
  E.atan_lut = o->ed_tab.as<double>();
  E.nfa_min_k = (const int*)(o->ed_tab.as<uint8_t>() + sizeof(double) * (sdpl_ed::kAtanLut + 1));
  if (o->ed_w != D.in_w || o->ed_h != D.in_h) {
    // the C library's atan / log / exp / pow of the validation tables are evaluated on the host (as in the oracle)
    std::vector<double> lut(sdpl_ed::kAtanLut + 1);
    sdpl_ed::host::atan_table(lut.data());
    std::vector<int> mk((size_t)kEdNfaN * kMaxOct, 0);
    for (int l = 0; l < nl; l++)
      if (!sdpl_ed::host::nfa_table(E.O[l].w, E.O[l].h, kEdNfaN, mk.data() + (size_t)l * kEdNfaN)) { set_last_error("EDLines: NFA table is not monotone"); return SDPL_ERR_UNSUPPORTED; }
    SDPL_CUDA(cudaMemcpyAsync((void*)E.atan_lut, lut.data(), sizeof(double) * lut.size(), cudaMemcpyHostToDevice, o->stream));
    SDPL_CUDA(cudaMemcpyAsync((void*)E.nfa_min_k, mk.data(), sizeof(int) * mk.size(), cudaMemcpyHostToDevice, o->stream));
    SDPL_CUDA(cudaStreamSynchronize(o->stream));
    o->ed_w = D.in_w; o->ed_h = D.in_h;
  }
  return SDPL_OK;
}

static int line_detect_dev(sdpl_line* o, const uint8_t* d_imgs, int B, int w, int h, int stride, size_t frame_stride,
                           sdpl_keyline* d_kls, int capacity, int* d_n_out) {
  int rc = line_setup(o, w, h, B);
  if (rc) return rc;
  if (o->extractor == 1 && (rc = ed_setup(o, B))) return rc;
  LineDev& D = o->D;
  D.B = B; D.in = d_imgs; D.in_stride = stride; D.in_frame = frame_stride; D.serial_mode = o->serial_mode;
  D.prof_detail = o->prof_detail;
  if (D.prof_detail) SDPL_CUDA(cudaMemsetAsync(D.prof, 0, sizeof(long long) * 16 * o->nlevels * B, o->stream));
  cudaStream_t st = o->stream;
  const int nl = o->nlevels;
  o->timer.begin(st);
  for (int l = 1; l < nl; l++) {
    k_line_resize<<<dim3(div_up(D.O[l].w, 128), div_up(D.O[l].h, kResizeRows), B), 128, 0, st>>>(D, l);
    SDPL_LAUNCH_CHECK();
  }
  o->timer.mark(st, "lsd_pyramid");
  if (o->extractor == 1) {
    // EDLines back-end (edlines.cuh): the lines land in the Pending slots, key lines / top-N / LBD are shared with the LSD back-end
    const EdDev& E = o->E;
    for (int l = 0; l < nl; l++) {
      const dim3 g(div_up(E.O[l].w, 64), div_up(E.O[l].h, 4), B);
      k_ed_smooth<<<g, 256, 0, st>>>(D, E, l);
      SDPL_LAUNCH_CHECK();
      k_ed_grad<<<g, 256, 0, st>>>(E, l);
      SDPL_LAUNCH_CHECK();
      k_ed_anchor<<<g, 256, 0, st>>>(E, l);
      SDPL_LAUNCH_CHECK();
    }
    o->timer.mark(st, "ed_gradient");
    k_ed_sort<<<nl * B, 32, 0, st>>>(D, E);
    SDPL_LAUNCH_CHECK();
    o->timer.mark(st, "ed_sort");
    k_ed_link<<<nl * B, 32, 0, st>>>(D, E);
    SDPL_LAUNCH_CHECK();
    o->timer.mark(st, "ed_link");
    k_ed_fit<<<nl * B, kEdFitThreads, 0, st>>>(D, E);
    SDPL_LAUNCH_CHECK();
    o->timer.mark(st, "ed_fit_validate");
    k_keylines<<<B, 256, 0, st>>>(D, d_kls, capacity, d_n_out, o->tmpkl.as<sdpl_keyline>());
    SDPL_LAUNCH_CHECK();
    o->timer.mark(st, "keylines");
    o->last_B = B;
    return SDPL_OK;
  }
  SDPL_CUDA(cudaMemsetAsync(D.maxg2, 0, sizeof(int) * nl * B, st));
  for (int l = 0; l < nl; l++) {
    k_lsd_scale<<<dim3(div_up(D.O[l].sw, kST_W), div_up(D.O[l].sh, kST_H), B), 256, 0, st>>>(D, l);
    SDPL_LAUNCH_CHECK();
  }
  o->timer.mark(st, "lsd_scale");
  for (int l = 0; l < nl; l++) {
    k_lsd_grad<<<dim3(div_up(D.O[l].sw, 64), div_up(D.O[l].sh, 4 * kGradTiles), B), 256, 0, st>>>(D, l);
    SDPL_LAUNCH_CHECK();
  }
  o->timer.mark(st, "lsd_gradient");
  for (int l = 0; l < nl; l++) {
    k_lsd_sort<false><<<dim3(div_up(D.O[l].nchunks, kSortWarps), B), kSortWarps * 32, 0, st>>>(D, l);
    SDPL_LAUNCH_CHECK();
  }
  k_lsd_sort_scan<<<dim3(nl, B), kBins, 0, st>>>(D);
  SDPL_LAUNCH_CHECK();
  for (int l = 0; l < nl; l++) {
    k_lsd_sort<true><<<dim3(div_up(D.O[l].nchunks, kSortWarps), B), kSortWarps * 32, 0, st>>>(D, l);
    SDPL_LAUNCH_CHECK();
  }
  k_lsd_task_rank<<<div_up(nl * B, 128), 128, 0, st>>>(D, nl * B);
  SDPL_LAUNCH_CHECK();
  o->timer.mark(st, "lsd_sort");
  if (o->serial_mode == 0 && !o->grow_legacy) {
    // two-phase waves (lsd_grow2.cuh).  Warps per task / CTAs per SM as below
    const bool small = nl * B <= o->sm_count;
    const bool forced = o->grow_warps > 0 && !(o->grow_forced && small);
    const int nw = forced ? o->grow_warps : (small ? o->grow_warps_small : 4);
    const int mb = forced && o->grow_minb > 0 ? o->grow_minb : (small ? 1 : 4);
    D.grow_ta = (small && !forced && o->grow_ta_small >= 0) ? o->grow_ta_small : o->grow_ta;
    if (nw >= 16) k_lsd_grow2<16, 1><<<nl * B, 512, 0, st>>>(D);
    else if (nw >= 8 && mb >= 2) k_lsd_grow2<8, 2><<<nl * B, 256, 0, st>>>(D);
    else if (nw >= 8) k_lsd_grow2<8, 1><<<nl * B, 256, 0, st>>>(D);
    else if (nw >= 4 && mb <= 2) k_lsd_grow2<4, 2><<<nl * B, 128, 0, st>>>(D);
    else if (nw >= 4 && mb == 3) k_lsd_grow2<4, 3><<<nl * B, 128, 0, st>>>(D);
    else if (nw >= 4 && mb <= 4) k_lsd_grow2<4, 4><<<nl * B, 128, o->grow_smem, st>>>(D);
    else if (nw >= 4 && mb == 5) k_lsd_grow2<4, 5><<<nl * B, 128, o->grow_smem, st>>>(D);
    else if (nw >= 4 && mb <= 6) k_lsd_grow2<4, 6><<<nl * B, 128, 0, st>>>(D);
    else if (nw >= 4) k_lsd_grow2<4, 8><<<nl * B, 128, 0, st>>>(D);
    else if (nw >= 2 && mb <= 4) k_lsd_grow2<2, 4><<<nl * B, 64, 0, st>>>(D);
    else if (nw >= 2 && mb <= 8) k_lsd_grow2<2, 8><<<nl * B, 64, 0, st>>>(D);
    else if (nw >= 2) k_lsd_grow2<2, 12><<<nl * B, 64, 0, st>>>(D);
    else k_lsd_grow2<1, 16><<<nl * B, 32, 0, st>>>(D);
  }
  else if (o->serial_mode == 0) {
    // legacy round-1 schedule (per-lane speculation): warps per task / CTAs per SM: 8 warps and one CTA per SM (256-seed waves, shortest latency) while every task gets an
    // SM to itself; 4 warps and 4 CTAs per SM (127 registers) for big batches (measured best at 512 frames: 134 ms vs 158 ms
    // for 8x2 and 250 ms for 8x1); sdpl_line_set_serial can pin both
    const bool small = nl * B <= o->sm_count;
    const int nw = o->grow_warps > 0 ? o->grow_warps : (small ? 8 : 4);
    const int mb = o->grow_minb > 0 ? o->grow_minb : (small ? 1 : 4);
    if (nw >= 8 && mb >= 2) k_lsd_grow_block<8, 2><<<nl * B, 256, 0, st>>>(D);
    else if (nw >= 8) k_lsd_grow_block<8, 1><<<nl * B, 256, 0, st>>>(D);
    else if (nw >= 4 && mb == 3) k_lsd_grow_block<4, 3><<<nl * B, 128, 0, st>>>(D);
    else if (nw >= 4 && mb == 4) k_lsd_grow_block<4, 4><<<nl * B, 128, 0, st>>>(D);
    else if (nw >= 4 && mb == 5) k_lsd_grow_block<4, 5><<<nl * B, 128, 0, st>>>(D);
    else if (nw >= 4 && mb >= 6) k_lsd_grow_block<4, 6><<<nl * B, 128, 0, st>>>(D);
    else if (nw >= 4) k_lsd_grow_block<4, 2><<<nl * B, 128, 0, st>>>(D);
    else if (nw >= 2 && mb >= 8) k_lsd_grow_block<2, 8><<<nl * B, 64, 0, st>>>(D);
    else if (nw >= 2 && mb >= 6) k_lsd_grow_block<2, 6><<<nl * B, 64, 0, st>>>(D);
    else if (nw >= 2) k_lsd_grow_block<2, 4><<<nl * B, 64, 0, st>>>(D);
    else k_lsd_grow_block<1, 8><<<nl * B, 32, 0, st>>>(D);
  }
  else k_lsd_grow<<<nl * B, 32, 0, st>>>(D);
  SDPL_LAUNCH_CHECK();
  o->timer.mark(st, "lsd_grow");
  SDPL_CUDA(cudaMemsetAsync(D.nbig, 0, sizeof(int) * 2 * nl * B, st));
  {
    const int nb = getenv("SDPL_NFA1_MINB") ? atoi(getenv("SDPL_NFA1_MINB")) : 24;   // 80 registers at 12 CTAs/SM: 4.61 ms NFA stage; 64 at 16: 4.51; 40 at 24: 4.42
    if (nb >= 24) k_lsd_nfa<24><<<dim3(div_up(D.pend_cap, 64), nl * B), 64, 0, st>>>(D);
    else if (nb >= 16) k_lsd_nfa<16><<<dim3(div_up(D.pend_cap, 64), nl * B), 64, 0, st>>>(D);
    else k_lsd_nfa<12><<<dim3(div_up(D.pend_cap, 64), nl * B), 64, 0, st>>>(D);
  }
  SDPL_LAUNCH_CHECK();
  // the big rectangles (one warp each, few and long: a latency-bound kernel at 6 % occupancy) and the second pass over the small
  // ones work on disjoint lists: side by side, unless the stages are being timed one after another
  const bool fork = !o->timer.enabled;
  cudaStream_t sb = fork ? o->side : st;
  if (fork) { SDPL_CUDA(cudaEventRecord(o->ev_fork, st)); SDPL_CUDA(cudaStreamWaitEvent(sb, o->ev_fork, 0)); }
  k_lsd_nfa_big<<<dim3(64, nl * B), 128, 0, sb>>>(D);
  SDPL_LAUNCH_CHECK();
  if (fork) SDPL_CUDA(cudaEventRecord(o->ev_join, sb));
  if (o->nfa_minb >= 16) k_lsd_nfa_rest<16><<<dim3(kRestBlocks, nl * B), 128, 0, st>>>(D);
  else if (o->nfa_minb >= 12) k_lsd_nfa_rest<12><<<dim3(kRestBlocks, nl * B), 128, 0, st>>>(D);
  else if (o->nfa_minb >= 10) k_lsd_nfa_rest<10><<<dim3(kRestBlocks, nl * B), 128, 0, st>>>(D);
  else k_lsd_nfa_rest<8><<<dim3(kRestBlocks, nl * B), 128, 0, st>>>(D);
  SDPL_LAUNCH_CHECK();
  if (fork) SDPL_CUDA(cudaStreamWaitEvent(st, o->ev_join, 0));
  o->timer.mark(st, "lsd_nfa");
  k_keylines<<<B, 256, 0, st>>>(D, d_kls, capacity, d_n_out, o->tmpkl.as<sdpl_keyline>());
  SDPL_LAUNCH_CHECK();
  o->timer.mark(st, "keylines");
  o->last_B = B;
  return SDPL_OK;
}

// LBD on device keylines of B frames (uses D.in as set by the caller)
static int line_lbd_dev(sdpl_line* o, const uint8_t* d_imgs, int B, int w, int h, int stride, size_t frame_stride,
                        const sdpl_keyline* d_kls, int capacity, const int* d_n, int max_lines, uint8_t* d_desc, float* d_fdesc) {
  int rc = line_setup(o, w, h, B);
  if (rc) return rc;
  LineDev& D = o->D;
  D.B = B; D.in = d_imgs; D.in_stride = stride; D.in_frame = frame_stride;
  cudaStream_t st = o->stream;
  const int nl = o->nlevels;
  k_lbd_blur_sobel<<<dim3(div_up(w, kLT_W), div_up(h, kLT_H), B), 256, 0, st>>>(D);
  SDPL_LAUNCH_CHECK();
  for (int l = 1; l < nl; l++) {
    k_lbd_pyrdown<<<dim3(div_up(D.O[l].lw, 64), div_up(D.O[l].lh, 4 * kLbdTiles), B), 256, 0, st>>>(D, l);
    SDPL_LAUNCH_CHECK();
  }
  for (int l = 1; l < nl; l++) {
    k_lbd_sobel<<<dim3(div_up(D.O[l].lw, 64), div_up(D.O[l].lh, 4 * kLbdTiles), B), 256, 0, st>>>(D, l);
    SDPL_LAUNCH_CHECK();
  }
  o->timer.mark(st, "lbd_sobel");
  if (max_lines > 0) {
    // CTAs per SM the band kernel is compiled for: 96 registers at 10 (2.39 ms per 512 frames), 64 at 16 (1.92 ms: the gathers want
    // resident warps more than the cold normalisation code wants registers), 40 at 24 (1.95 ms)
    const int lb = getenv("SDPL_LBD_MINB") ? atoi(getenv("SDPL_LBD_MINB")) : 16;
    if (lb >= 24) k_lbd<24><<<dim3(std::min(max_lines, 256), B), 64, 0, st>>>(D, d_kls, capacity, d_n, d_desc, d_fdesc);
    else if (lb >= 16) k_lbd<16><<<dim3(std::min(max_lines, 256), B), 64, 0, st>>>(D, d_kls, capacity, d_n, d_desc, d_fdesc);
    else k_lbd<10><<<dim3(std::min(max_lines, 256), B), 64, 0, st>>>(D, d_kls, capacity, d_n, d_desc, d_fdesc);
    SDPL_LAUNCH_CHECK();
  }
  o->timer.mark(st, "lbd_bands");
  return SDPL_OK;
}

static int line_check_err(sdpl_line* o) {
  int e = 0;
  SDPL_CUDA(cudaMemcpyAsync(&e, o->err.p, sizeof(int), cudaMemcpyDeviceToHost, o->stream));
  SDPL_CUDA(cudaStreamSynchronize(o->stream));
  if (e) {
    cudaMemsetAsync(o->err.p, 0, sizeof(int), o->stream);
    set_last_error("device buffer overflow in the line pipeline (pending rectangles / keylines)");
    return e;
  }
  return SDPL_OK;
}

extern "C" {

int sdpl_line_create(sdpl_line** out, int nfeatures, int refine, float lsd_scale, int nlevels, float scale, int extractor, int device) {
  if (!out || nfeatures < 0 || refine < 0 || refine > 2 || nlevels < 1 || !(scale > 1.0f)) {
    set_last_error("sdpl_line_create: bad argument");
    return SDPL_ERR_ARG;
  }
  if (extractor != 0 && extractor != 1) { set_last_error("sdpl_line_create: extractor must be 0 (LSD) or 1 (EDLines)"); return SDPL_ERR_UNSUPPORTED; }
  if (nlevels > kMaxOct) { set_last_error("sdpl_line_create: at most 4 octaves"); return SDPL_ERR_UNSUPPORTED; }
  {  // only the reference's pre-scaling kernel is implemented: sigma = 0.6/0.8 = 0.75 -> 7 taps [0,4,56,136,56,4,0]
    const double sigma = 0.6 / (double)lsd_scale;
    if (!(lsd_scale < 1.0f) || fabs(sigma - 0.75) > 1e-6) { set_last_error("sdpl_line_create: only lsd_scale = 0.8 is implemented"); return SDPL_ERR_UNSUPPORTED; }
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
    set_last_error("sdpl_line_create: no such CUDA device (this library has no CPU fallback)");
    return SDPL_ERR_CUDA;
  }
  SDPL_CUDA(cudaSetDevice(device));
  sdpl_line* o = new sdpl_line;
  o->nfeatures = nfeatures; o->refine = refine; o->nlevels = nlevels; o->extractor = extractor; o->device = device;
  o->lsd_scale = lsd_scale; o->scale = scale;
  // LSDDetectorC::ComputePyramid scale tables, LSDDetector_custom.cpp:79-90
  o->sf.resize(nlevels); o->isf.resize(nlevels);
  o->sf[0] = 1.0f;
  for (int l = 0; l < nlevels; l++) {
    if (l > 0) o->sf[l] = o->sf[l - 1] * scale;
    o->isf[l] = 1.0f / o->sf[l];
  }
  cudaDeviceGetAttribute(&o->sm_count, cudaDevAttrMultiProcessorCount, device);
  SDPL_CUDA(cudaStreamCreateWithFlags(&o->own_stream, cudaStreamNonBlocking));
  o->stream = o->own_stream;
  SDPL_CUDA(cudaStreamCreateWithFlags(&o->side, cudaStreamNonBlocking));
  SDPL_CUDA(cudaEventCreateWithFlags(&o->ev_fork, cudaEventDisableTiming));
  SDPL_CUDA(cudaEventCreateWithFlags(&o->ev_join, cudaEventDisableTiming));
  if (const char* g = getenv("SDPL_GROW")) {
    // tuning knob read at creation: "warps,ctas_per_sm,phaseA_cap[,dynamic_smem_bytes]" of the region-growing kernel (big batches)
    int nw = 0, mb = 0, ta = -1, sm = 0;
    if (sscanf(g, "%d,%d,%d,%d", &nw, &mb, &ta, &sm) >= 2) {
      o->grow_warps = std::min(std::max(nw, 1), lsd::kMaxGrowWarps2); o->grow_minb = std::max(mb, 1);
      if (ta >= 0) o->grow_ta = ta;
      o->grow_smem = std::max(sm, 0);
      o->grow_forced = 1;
    }
  }
  if (const char* g = getenv("SDPL_GROW_SMALL")) {      // the same for small batches (every task has an SM to itself): "warps,phaseA_cap"
    int nw = 0, ta = -1;
    if (sscanf(g, "%d,%d", &nw, &ta) >= 1) { o->grow_warps_small = std::min(std::max(nw, 1), lsd::kMaxGrowWarps2); if (ta >= 0) o->grow_ta_small = ta; }
  }
  if (o->grow_smem > 0) {
    SDPL_CUDA(cudaFuncSetAttribute(k_lsd_grow2<4, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, o->grow_smem));
    SDPL_CUDA(cudaFuncSetAttribute(k_lsd_grow2<4, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, o->grow_smem));
  }
  *out = o;
  return SDPL_OK;
}

void sdpl_line_destroy(sdpl_line* o) {
  if (!o) return;
  cudaSetDevice(o->device);
  cudaStreamSynchronize(o->stream);
  for (DevBuf* b : {&o->lvl, &o->scaled, &o->px, &o->ang, &o->g2, &o->order, &o->ed_img, &o->ed_work, &o->ed_nanch, &o->ed_tab, &o->hist, &o->maxg2, &o->ndef, &o->torder, &o->reg, &o->pend, &o->npend,
                    &o->g, &o->sd, &o->err, &o->tables, &o->tmpkl, &o->in_stage, &o->out_kls, &o->out_desc, &o->out_n, &o->prof, &o->lgam, &o->nbig, &o->bigidx, &o->ctx, &o->nfatab})
    b->release();
  if (o->h_stage) cudaFreeHost(o->h_stage);
  o->timer.release();
  if (o->own_stream) cudaStreamDestroy(o->own_stream);
  if (o->side) cudaStreamDestroy(o->side);
  if (o->ev_fork) cudaEventDestroy(o->ev_fork);
  if (o->ev_join) cudaEventDestroy(o->ev_join);
  delete o;
}

int sdpl_line_set_stream(sdpl_line* o, void* s) { if (!o) return SDPL_ERR_ARG; o->stream = s ? (cudaStream_t)s : o->own_stream; return SDPL_OK; }
int sdpl_line_levels(const sdpl_line* o) { return o ? o->nlevels : 0; }
// mvScaleFactor_l / mvInvScaleFactor_l / mvLevelSigma2_l / mvInvLevelSigma2_l, Lineextractor.cc:84-96
int sdpl_line_tables(const sdpl_line* o, float* sf, float* isf, float* s2, float* is2) {
  if (!o) return SDPL_ERR_ARG;
  for (int i = 0; i < o->nlevels; i++) {
    const float sig = i == 0 ? 1.0f : o->sf[i] * o->sf[i];
    if (sf) sf[i] = o->sf[i];
    if (isf) isf[i] = o->isf[i];
    if (s2) s2[i] = sig;
    if (is2) is2[i] = 1.0f / sig;
  }
  return SDPL_OK;
}
int sdpl_line_last_launches(const sdpl_line* o) { return o ? o->launches : 0; }
int sdpl_line_peek_error_async(sdpl_line* o, void* stream, int* host_flag) {
  if (!o || !host_flag) return SDPL_ERR_ARG;
  if (!o->err.p) { *host_flag = 0; return SDPL_OK; }
  SDPL_CUDA(cudaMemcpyAsync(host_flag, o->err.p, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  return SDPL_OK;
}
// the overflow flag of everything enqueued so far on the handle's stream, copied to *dst (pinned host or device memory) and
// cleared, both in stream order: the flag then belongs to the batch that raised it
int sdpl_line_take_error_async(sdpl_line* o, int* dst) {
  if (!o || !dst) return SDPL_ERR_ARG;
  if (!o->err.p) { *dst = 0; return SDPL_OK; }
  SDPL_CUDA(cudaMemcpyAsync(dst, o->err.p, sizeof(int), cudaMemcpyDefault, o->stream));
  SDPL_CUDA(cudaMemsetAsync(o->err.p, 0, sizeof(int), o->stream));
  return SDPL_OK;
}
int sdpl_line_check(sdpl_line* o) {
  if (!o) return SDPL_ERR_ARG;
  if (!o->err.p) return SDPL_OK;
  SDPL_CUDA(cudaSetDevice(o->device));
  return line_check_err(o);
}
int sdpl_line_set_profiling(sdpl_line* o, int on) { if (!o) return SDPL_ERR_ARG; o->timer.enabled = on != 0; return SDPL_OK; }
int sdpl_line_stage_times(sdpl_line* o, float* ms, const char** names, int* launches, int cap) {
  if (!o) return 0;
  cudaSetDevice(o->device);
  return o->timer.read(ms, names, launches, cap);
}
// test / tuning knob: region-growing schedule (bits 0-1).  0 = speculative waves, one CTA of NW warps per task (default: the two-phase
// schedule of lsd_grow2.cuh; bit 2 selects the round-1 per-lane schedule instead), 1 = strictly one seed at a time (no speculation),
// 3 = single-warp waves of 32 seeds.  Bits 8-11 / 12-15: warps per task / CTAs per SM of mode 0; bits 3-7: CTAs per SM the second
// NFA pass is compiled for; bits 24-30: phase-A expansion cap + 1.  All give identical results.  (Mode 2, a re-order-buffer schedule
// of round 1, was measured 3x slower than the waves and has been removed.)
int sdpl_line_set_extractor(sdpl_line* o, int extractor) {
  if (!o || (extractor != 0 && extractor != 1)) { set_last_error("sdpl_line_set_extractor: extractor must be 0 (LSD) or 1 (EDLines)"); return SDPL_ERR_ARG; }
  o->extractor = extractor;
  return SDPL_OK;
}

int sdpl_line_set_serial(sdpl_line* o, int on) {
  if (!o || on < 0 || (on & 3) == 2) return SDPL_ERR_ARG;
  o->serial_mode = on & 3;
  if ((on >> 3) & 31) o->nfa_minb = (on >> 3) & 31;    // bits 3-7: CTAs per SM the second NFA pass is compiled for (tuning)
  o->grow_legacy = (on >> 2) & 1;                    // bit 2: the round-1 per-lane schedule instead of the two-phase one
  if ((on >> 24) & 0x7f) o->grow_ta = ((on >> 24) & 0x7f) - 1;   // bits 24-30: phase-A expansion cap + 1
  const int w = (on >> 8) & 0xffff;
  if (w && o->serial_mode == 0) { o->grow_warps = std::min(w & 15, lsd::kMaxGrowWarps); if (w >> 4) o->grow_minb = w >> 4; }
  return SDPL_OK;
}

int sdpl_line_extract_batch_dev(sdpl_line* o, const uint8_t* d_imgs, int nframes, int w, int h, int stride, size_t frame_stride,
                                sdpl_keyline* d_kls, uint8_t* d_desc, int capacity, int* d_n_out, int sync) {
  if (!o || !d_imgs || nframes < 1 || w < 1 || h < 1 || stride < w || !d_kls || !d_desc || capacity < 1 || !d_n_out) {
    set_last_error("sdpl_line_extract_batch_dev: bad argument");
    return SDPL_ERR_ARG;
  }
  SDPL_CUDA(cudaSetDevice(o->device));
  g_launches = 0;
  int rc = line_detect_dev(o, d_imgs, nframes, w, h, stride, frame_stride, d_kls, capacity, d_n_out);
  if (!rc) rc = line_lbd_dev(o, d_imgs, nframes, w, h, stride, frame_stride, d_kls, capacity, d_n_out,
                             std::min(capacity, o->nfeatures ? o->nfeatures : o->pend_cap * o->nlevels), d_desc, nullptr);
  o->launches = g_launches;
  if (rc) return rc;
  if (sync) return line_check_err(o);
  return SDPL_OK;
}

int sdpl_line_extract_batch(sdpl_line* o, const uint8_t* imgs, int nframes, int w, int h, int stride, size_t frame_stride,
                            sdpl_keyline* kls, uint8_t* desc, int capacity, int* n_out) {
  if (!o || !n_out || nframes < 0) { set_last_error("sdpl_line_extract_batch: bad argument"); return SDPL_ERR_ARG; }
  if (!imgs || w <= 0 || h <= 0 || nframes == 0) {
    for (int f = 0; f < nframes; f++) n_out[f] = 0;
    return SDPL_OK;
  }
  if (stride < w || !kls || !desc || capacity < 1) { set_last_error("sdpl_line_extract_batch: bad argument"); return SDPL_ERR_ARG; }
  SDPL_CUDA(cudaSetDevice(o->device));
  int rc;
  if ((rc = line_setup(o, w, h, nframes))) return rc;     // fixes pend_cap for this geometry
  const int cap_dev = o->nfeatures ? std::max(o->nfeatures, 1) : o->pend_cap * o->nlevels;
  const size_t in_bytes = (size_t)w * h * nframes;
  if ((rc = o->in_stage.reserve(in_bytes))) return rc;
  if ((rc = o->out_kls.reserve(sizeof(sdpl_keyline) * (size_t)cap_dev * nframes))) return rc;
  if ((rc = o->out_desc.reserve((size_t)32 * cap_dev * nframes))) return rc;
  if ((rc = o->out_n.reserve(sizeof(int) * nframes))) return rc;
  const size_t out_bytes = (sizeof(sdpl_keyline) + 32) * (size_t)cap_dev * nframes + sizeof(int) * nframes;
  if (o->h_stage_bytes < out_bytes) {
    if (o->h_stage) cudaFreeHost(o->h_stage);
    o->h_stage = nullptr; o->h_stage_bytes = 0;
    SDPL_CUDA(cudaMallocHost(&o->h_stage, out_bytes));
    o->h_stage_bytes = out_bytes;
  }
  cudaStream_t st = o->stream;
  if (stride == w && frame_stride == (size_t)w * h) {
    SDPL_CUDA(cudaMemcpyAsync(o->in_stage.p, imgs, in_bytes, cudaMemcpyHostToDevice, st));
  } else {
    for (int f = 0; f < nframes; f++)
      SDPL_CUDA(cudaMemcpy2DAsync((char*)o->in_stage.p + (size_t)f * w * h, w, imgs + (size_t)f * frame_stride, stride, w, h,
                                  cudaMemcpyHostToDevice, st));
  }
  g_launches = 0;
  rc = line_detect_dev(o, o->in_stage.as<uint8_t>(), nframes, w, h, w, (size_t)w * h, o->out_kls.as<sdpl_keyline>(), cap_dev,
                       o->out_n.as<int>());
  if (rc) { o->launches = g_launches; return rc; }
  // the line counts decide the LBD grid: fetch them (small D2H) before launching the descriptor kernel
  char* hs = (char*)o->h_stage;
  sdpl_keyline* h_k = (sdpl_keyline*)hs;
  uint8_t* h_d = (uint8_t*)(hs + sizeof(sdpl_keyline) * (size_t)cap_dev * nframes);
  int* h_n = (int*)(hs + (sizeof(sdpl_keyline) + 32) * (size_t)cap_dev * nframes);
  SDPL_CUDA(cudaMemcpyAsync(h_n, o->out_n.p, sizeof(int) * nframes, cudaMemcpyDeviceToHost, st));
  if ((rc = line_check_err(o))) { o->launches = g_launches; return rc; }
  int max_lines = 0;
  for (int f = 0; f < nframes; f++) max_lines = std::max(max_lines, std::min(h_n[f], cap_dev));
  rc = line_lbd_dev(o, o->in_stage.as<uint8_t>(), nframes, w, h, w, (size_t)w * h, o->out_kls.as<sdpl_keyline>(), cap_dev,
                    o->out_n.as<int>(), max_lines, o->out_desc.as<uint8_t>(), nullptr);
  o->launches = g_launches;
  if (rc) return rc;
  if (max_lines > 0) {
    for (int f = 0; f < nframes; f++) {
      const int n = std::min(h_n[f], cap_dev);
      if (n <= 0) continue;
      SDPL_CUDA(cudaMemcpyAsync(h_k + (size_t)f * cap_dev, o->out_kls.as<sdpl_keyline>() + (size_t)f * cap_dev, sizeof(sdpl_keyline) * n,
                                cudaMemcpyDeviceToHost, st));
      SDPL_CUDA(cudaMemcpyAsync(h_d + (size_t)f * cap_dev * 32, o->out_desc.as<uint8_t>() + (size_t)f * cap_dev * 32, (size_t)32 * n,
                                cudaMemcpyDeviceToHost, st));
    }
  }
  if ((rc = line_check_err(o))) return rc;
  int status = SDPL_OK;
  for (int f = 0; f < nframes; f++) {
    int n = h_n[f];
    n_out[f] = n;
    if (n > cap_dev) { n = cap_dev; status = SDPL_ERR_OVERFLOW; }
    if (n > capacity) { status = SDPL_ERR_CAPACITY; n = capacity; }
    memcpy(kls + (size_t)f * capacity, h_k + (size_t)f * cap_dev, sizeof(sdpl_keyline) * n);
    memcpy(desc + (size_t)f * capacity * 32, h_d + (size_t)f * cap_dev * 32, (size_t)32 * n);
  }
  if (status == SDPL_ERR_CAPACITY) set_last_error("sdpl_line_extract: output capacity too small");
  if (status == SDPL_ERR_OVERFLOW) set_last_error("sdpl_line_extract: more keylines than the internal capacity");
  return status;
}

int sdpl_line_extract(sdpl_line* o, const uint8_t* img, int w, int h, int stride, sdpl_keyline* kls, uint8_t* desc, int capacity,
                      int* n_out) {
  return sdpl_line_extract_batch(o, img, 1, w, h, stride, (size_t)stride * (h > 0 ? h : 0), kls, desc, capacity, n_out);
}

int sdpl_line_lbd_compute(sdpl_line* o, const uint8_t* img, int w, int h, int stride, const sdpl_keyline* kls, int n, uint8_t* desc) {
  if (!o || !img || w < 1 || h < 1 || stride < w || n < 0) { set_last_error("sdpl_line_lbd_compute: bad argument"); return SDPL_ERR_ARG; }
  if (n == 0) return SDPL_OK;   // "Error: keypoint list is empty": descriptors untouched, binary_descriptor_custom.cpp:556-560
  if (!kls || !desc) { set_last_error("sdpl_line_lbd_compute: null pointer"); return SDPL_ERR_ARG; }
  for (int i = 0; i < n; i++)
    if (kls[i].octave < 0 || kls[i].octave >= o->nlevels) { set_last_error("sdpl_line_lbd_compute: keyline octave outside the pyramid"); return SDPL_ERR_ARG; }
  SDPL_CUDA(cudaSetDevice(o->device));
  int rc;
  if ((rc = o->in_stage.reserve((size_t)w * h))) return rc;
  if ((rc = o->out_kls.reserve(sizeof(sdpl_keyline) * (size_t)n))) return rc;
  if ((rc = o->out_desc.reserve((size_t)32 * n))) return rc;
  if ((rc = o->out_n.reserve(sizeof(int)))) return rc;
  cudaStream_t st = o->stream;
  SDPL_CUDA(cudaMemcpy2DAsync(o->in_stage.p, w, img, stride, w, h, cudaMemcpyHostToDevice, st));
  SDPL_CUDA(cudaMemcpyAsync(o->out_kls.p, kls, sizeof(sdpl_keyline) * n, cudaMemcpyHostToDevice, st));
  SDPL_CUDA(cudaMemcpyAsync(o->out_n.p, &n, sizeof(int), cudaMemcpyHostToDevice, st));
  g_launches = 0;
  o->timer.begin(st);
  rc = line_lbd_dev(o, o->in_stage.as<uint8_t>(), 1, w, h, w, (size_t)w * h, o->out_kls.as<sdpl_keyline>(), n, o->out_n.as<int>(), n,
                    o->out_desc.as<uint8_t>(), nullptr);
  o->launches = g_launches;
  if (rc) return rc;
  SDPL_CUDA(cudaMemcpyAsync(desc, o->out_desc.p, (size_t)32 * n, cudaMemcpyDeviceToHost, st));
  SDPL_CUDA(cudaStreamSynchronize(st));
  return SDPL_OK;
}

// introspection: rectangles handed to the NFA stage, in seed order: 8 doubles each {x1,y1,x2,y2,width,seed,npix,tag}
int sdpl_line_debug_pending(sdpl_line* o, int frame, int octave, double* out, int capacity, int* n_out) {
  if (!o || o->gw == 0 || frame < 0 || frame >= o->last_B || octave < 0 || octave >= o->nlevels || !n_out) return SDPL_ERR_ARG;
  SDPL_CUDA(cudaSetDevice(o->device));
  SDPL_CUDA(cudaStreamSynchronize(o->stream));
  const int task = frame * o->nlevels + octave;
  int np = 0;
  SDPL_CUDA(cudaMemcpy(&np, o->D.npend + task, sizeof(int), cudaMemcpyDeviceToHost));
  std::vector<lsd::Pending> P(np);
  if (np) SDPL_CUDA(cudaMemcpy(P.data(), o->D.pend + (size_t)task * o->D.pend_cap, sizeof(lsd::Pending) * np, cudaMemcpyDeviceToHost));
  for (int i = 0; i < np && i < capacity && out; i++) {
    double* d = out + 8 * i;
    d[0] = P[i].rec.x1; d[1] = P[i].rec.y1; d[2] = P[i].rec.x2; d[3] = P[i].rec.y2; d[4] = P[i].rec.width; d[5] = P[i].rec.p;
    d[6] = P[i].accepted; d[7] = P[i].tag;
  }
  *n_out = np;
  return SDPL_OK;
}

// introspection: grow-kernel cycle counters of one task: {select, speculate, evaluate+commit, re-run, waves, re-runs, dead, seeds}
int sdpl_line_debug_grow_detail(sdpl_line* o, int on) { if (!o) return SDPL_ERR_ARG; o->prof_detail = on; return SDPL_OK; }
int sdpl_line_debug_grow_profile(sdpl_line* o, int frame, int octave, long long* out8) {
  if (!o || o->gw == 0 || frame < 0 || frame >= o->last_B || octave < 0 || octave >= o->nlevels || !out8) return SDPL_ERR_ARG;
  SDPL_CUDA(cudaSetDevice(o->device));
  SDPL_CUDA(cudaStreamSynchronize(o->stream));
  SDPL_CUDA(cudaMemcpy(out8, o->D.prof + (size_t)(frame * o->nlevels + octave) * 16, sizeof(long long) * 16, cudaMemcpyDeviceToHost));
  return SDPL_OK;
}

int sdpl_line_lsd_segments(sdpl_line* o, int frame, int octave, float* xyxy, int capacity, int* n_out) {
  if (!o || o->gw == 0 || frame < 0 || frame >= o->last_B || octave < 0 || octave >= o->nlevels || !n_out) return SDPL_ERR_ARG;
  SDPL_CUDA(cudaSetDevice(o->device));
  SDPL_CUDA(cudaStreamSynchronize(o->stream));
  const int task = frame * o->nlevels + octave;
  int np = 0;
  SDPL_CUDA(cudaMemcpy(&np, o->D.npend + task, sizeof(int), cudaMemcpyDeviceToHost));
  std::vector<lsd::Pending> P(np);
  if (np) SDPL_CUDA(cudaMemcpy(P.data(), o->D.pend + (size_t)task * o->D.pend_cap, sizeof(lsd::Pending) * np, cudaMemcpyDeviceToHost));
  int n = 0;
  for (int i = 0; i < np; i++) {
    if (!P[i].accepted) continue;
    if (xyxy && n < capacity) memcpy(xyxy + 4 * n, P[i].seg, 16);
    n++;
  }
  *n_out = n;
  return SDPL_OK;
}

}  // extern "C"
