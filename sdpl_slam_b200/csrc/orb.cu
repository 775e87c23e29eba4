// orb.cu -- sm_100a CUDA implementation of the SDPL-SLAM ORB extractor (reference: src/ORBextractor.cc).
//
// Batched over frames: every kernel takes a frame index in the grid so that one launch processes all
// frames (and, where possible, all pyramid levels) of a chunk.  Device data layout per chunk of B frames:
//   pyr   [B][sum_l (h_l+38)*pstride_l]   u8   padded pyramid planes (19 px reflect-101 border)      (K1)
//   score [B][sum_l h_l*sstride_l]        u8   FAST-9/16 corner score, 0 where < minThFAST           (K2)
//   blur  [B][sum_l (h_l+38)*pstride_l]   u8   7x7 sigma-2 blurred levels, same padded geometry as pyr (K6)
//   cell  [B][ncells]                     u32  per 30-px cell: (threshold<<16) | keypoint count       (K3a)
//   cand  [B][sum_l cand_cap_l]           u32 (y<<16|x) + u8 response, cell-major/row-major order     (K3b)
//   kp    [B][sum_l kp_cap_l]             u32 (y<<16|x) + u8 response in quadtree list order          (K4)
//   out   [B][capacity] sdpl_keypoint + [B][capacity][32] descriptors + [B] counts                    (K5/K7)
// All stages are integer/byte work (HBM/L2 bound); no tensor cores are involved.
#include "common.cuh"
#include <math.h>
#include <string.h>
#include <algorithm>

namespace sdpl {

constexpr int kMaxLevels = 16;
constexpr int kBorder = 19;       // EDGE_THRESHOLD, src/ORBextractor.cc:63
constexpr int kCellApron = 66;    // max (interior + 2) per cell side handled by the NMS kernel

struct LvlDev {
  int w, h, pstride, sstride;
  unsigned long long pyr_off, s_off;       // byte offsets inside one frame's pyr / score(blur) region
  int cell_base, ncells, quota, cand_cap, cand_off, kp_off, kp_cap;
  int fast_tile_base, fast_tiles_x, blur_tile_base, blur_tiles_x;
  int nIni; float hX;
  int area_fast;                           // 1: exact 2x2 decimation (INTER_AREA path of cv::resize)
  const unsigned short* xofs; const short2* xa; const unsigned short* yofs; const short2* ya;  // resize tables
};

struct CellDev { short level, x0, y0, x1, y1, sx, sy, pad; };  // window [x0,x1)x[y0,y1) in level coords, shift

struct OrbDev {
  int nl, B, ini_th, min_th;
  unsigned long long pyr_frame, s_frame;
  int blur_tiles;
  int cells_per_frame, cand_per_frame, kp_per_frame;
  uint8_t *pyr, *score, *blur;
  uint32_t* cellinfo;
  uint2* cellmask; int mask_words;     // per cell: (maximum, maximum >= iniThFAST) bit masks, 32 pixels per entry
  uint32_t* cand_xy; uint8_t* cand_resp; unsigned short* cand_node;
  int* n_cand; int* n_kp;
  uint32_t* kp_xy; uint8_t* kp_resp;
  const CellDev* cells;
  int* err;
  LvlDev L[kMaxLevels];
  float sf[kMaxLevels];
  int umax[16];
};

// ------------------------------------------------------------------------------------------------
// K1: pyramid.  ComputePyramid, src/ORBextractor.cc:1112-1137.
// ------------------------------------------------------------------------------------------------
constexpr int kBaseRows = 8;
__global__ void __launch_bounds__(128) k_pyr_base(const uint8_t* __restrict__ in, int in_stride, size_t in_frame,
                                                   uint8_t* __restrict__ out, int w, int h, int pstride, size_t out_frame) {
  int x4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (x4 >= pstride) return;
  // kBaseRows rows per thread: the reflected source columns are computed once (and a one-row block was too short-lived)
  int sx[4];
#pragma unroll
  for (int k = 0; k < 4; k++) sx[k] = x4 + k < w + 2 * kBorder ? reflect101(x4 + k - kBorder, w) : -1;
  const int rows = h + 2 * kBorder;
#pragma unroll 1
  for (int y = blockIdx.y * kBaseRows; y < min(rows, (int)(blockIdx.y + 1) * kBaseRows); y++) {
    const uint8_t* src = in + (size_t)blockIdx.z * in_frame + (size_t)reflect101(y - kBorder, h) * in_stride;
    uint32_t v = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const uint32_t b = sx[k] >= 0 ? src[sx[k]] : 0;
      v |= b << (8 * k);
    }
    *(uint32_t*)(out + (size_t)blockIdx.z * out_frame + (size_t)y * pstride + x4) = v;
  }
}

// level l from level l-1: cv::resize INTER_LINEAR (Q11 fixed point) + reflect-101 border in one pass.  A thread owns four
// output columns (one aligned word) and kPyrRows consecutive rows: the column part of the bilinear set-up (reflected
// column, source offset, Q11 weights from the tables) is done once and reused down the rows.
constexpr int kPyrRows = 8;
__global__ void __launch_bounds__(128) k_pyr_resize(OrbDev D, int l) {
  const LvlDev& Ld = D.L[l];
  const LvlDev& Ls = D.L[l - 1];
  const int x4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (x4 >= Ld.pstride) return;
  uint8_t* frame = D.pyr + (size_t)blockIdx.z * D.pyr_frame;
  const uint8_t* src = frame + Ls.pyr_off + (size_t)kBorder * Ls.pstride + kBorder;
  const int rows = Ld.h + 2 * kBorder;
  const int y_first = blockIdx.y * kPyrRows;
  int sx0[4], sx1[4];
  short2 aa[4];
  bool in[4];
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const int x = x4 + k;
    in[k] = x < Ld.w + 2 * kBorder;
    const int ix = reflect101(min(x, Ld.w + 2 * kBorder - 1) - kBorder, Ld.w);
    if (Ld.area_fast) { sx0[k] = 2 * ix; sx1[k] = 2 * ix + 1; aa[k] = make_short2(0, 0); }
    else { sx0[k] = Ld.xofs[ix]; sx1[k] = min(sx0[k] + 1, Ls.w - 1); aa[k] = Ld.xa[ix]; }
  }
  int h0[4], h1[4] = {0, 0, 0, 0}, prev_sy1 = -1;
#pragma unroll 1
  for (int r = 0; r < kPyrRows; r++) {
    const int y = y_first + r;
    if (y >= rows) break;
    const int iy = reflect101(y - kBorder, Ld.h);
    uint32_t v = 0;
    if (Ld.area_fast) {
      const uint8_t* s0 = src + (size_t)(2 * iy) * Ls.pstride;
      const uint8_t* s1 = s0 + Ls.pstride;
#pragma unroll
      for (int k = 0; k < 4; k++) {
        uint32_t b = 0;
        if (in[k]) b = (s0[sx0[k]] + s0[sx1[k]] + s1[sx0[k]] + s1[sx1[k]] + 2) >> 2;
        v |= b << (8 * k);
      }
    } else {
      const int sy = (short)Ld.yofs[iy];
      const short2 bb = Ld.ya[iy];
      const int sy0 = min(max(sy, 0), Ls.h - 1), sy1 = min(max(sy + 1, 0), Ls.h - 1);
      // the rows of a block are uniform across its threads: when this output row's upper source row is the previous output
      // row's lower one (4 rows of 5 at scale 1.2) its horizontal pass is already in registers
      if (sy0 == prev_sy1) {
#pragma unroll
        for (int k = 0; k < 4; k++) h0[k] = h1[k];
      } else {
        const uint8_t* s0 = src + (size_t)sy0 * Ls.pstride;
#pragma unroll
        for (int k = 0; k < 4; k++) h0[k] = s0[sx0[k]] * aa[k].x + s0[sx1[k]] * aa[k].y;
      }
      const uint8_t* s1 = src + (size_t)sy1 * Ls.pstride;
#pragma unroll
      for (int k = 0; k < 4; k++) h1[k] = s1[sx0[k]] * aa[k].x + s1[sx1[k]] * aa[k].y;
      prev_sy1 = sy1;
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const int o = (((bb.x * (h0[k] >> 4)) >> 16) + ((bb.y * (h1[k] >> 4)) >> 16) + 2) >> 2;
        const uint32_t b = in[k] ? (uint32_t)min(max(o, 0), 255) : 0u;
        v |= b << (8 * k);
      }
    }
    *(uint32_t*)(frame + Ld.pyr_off + (size_t)y * Ld.pstride + x4) = v;
  }
}

// ------------------------------------------------------------------------------------------------
// K2: FAST-9/16 corner score (cv::FAST + cornerScore<16>), threshold independent:
//     score = max over the 16 arcs of 9 contiguous ring pixels of min(signed difference) - 1;
//     a pixel is a corner at threshold t  <=>  score >= t.  Stored as u8, 0 where score < minThFAST.
// ------------------------------------------------------------------------------------------------
constexpr int kFT_W = 64, kFT_H = 32;  // FAST tile

__device__ __forceinline__ int arc9_max_of_min(const int (&d)[16]) {
  int m2[16], m4[16], m8[16];
#pragma unroll
  for (int k = 0; k < 16; k++) m2[k] = min(d[k], d[(k + 1) & 15]);
#pragma unroll
  for (int k = 0; k < 16; k++) m4[k] = min(m2[k], m2[(k + 2) & 15]);
#pragma unroll
  for (int k = 0; k < 16; k++) m8[k] = min(m4[k], m4[(k + 4) & 15]);
  int best = -256;
#pragma unroll
  for (int k = 0; k < 16; k++) best = max(best, min(m8[k], d[(k + 8) & 15]));
  return best;
}

__device__ __forceinline__ bool has_arc9(uint32_t m16) {
  uint32_t m = m16 | (m16 << 16);
  m &= m >> 1; m &= m >> 2; m &= m >> 4;  // runs of 8
  m &= m >> 1;                            // runs of 9
  return m != 0;
}

// Two horizontally adjacent pixels per step as packed s16x2 lanes: the 16 ring differences d[k] = ring[k] - centre are one
// VIADD.16x2 each, the "minimum over every arc of 9" is two levels of the native three-input VIMNMX3.S16x2
// (min3 of min3's), the dark side is the same with max3 on the same differences (min over the arc of (c - r) = -max(r - c)).
// Shared memory holds the tile twice, the second copy shifted by one byte, so every ring pair is an aligned 16-bit load.
__device__ __forceinline__ unsigned expand2(unsigned v16) { return __byte_perm(v16, 0u, 0x4140); }   // (b0, b1) -> s16x2

__global__ void __launch_bounds__(256) k_fast_score(OrbDev D, int total_tiles) {
  constexpr int S = kFT_W + 12;                                 // row stride: 19 words
  __shared__ __align__(4) uint8_t tileA[(kFT_H + 6) * S];
  __shared__ __align__(4) uint8_t tileB[(kFT_H + 6) * S];       // tileB[i] == tileA[i + 1]
  int t = blockIdx.x;
  int l = 0;
#pragma unroll 1
  for (int i = 1; i < D.nl; i++) if (t >= D.L[i].fast_tile_base) l = i;
  const LvlDev& L = D.L[l];
  t -= L.fast_tile_base;
  const int tx = t % L.fast_tiles_x, ty = t / L.fast_tiles_x;
  const int x0 = kBorder + tx * kFT_W, y0 = kBorder + ty * kFT_H;   // level (interior) coordinates of the tile
  // stage tile + 3 px halo as aligned 32-bit words of the padded plane.  The tile starts at padded column 2 * kBorder - 3 + 64 tx
  // = 35 + 64 tx: the word-aligned copy starts 3 bytes earlier (kShift), so tile column xx sits at byte xx + kShift of a
  // staged row.  Words past the end of a padded row / rows past the plane are clamped: they only feed pixels that are
  // not written (a written pixel's ring stays inside the plane).  tileB = tileA shifted by one byte (funnel shift).
  constexpr int kShift = (2 * kBorder - 3) & 3;
  static_assert(kShift == 3 && (kFT_W & 3) == 0, "tile origin alignment");
  constexpr int kWords = (kFT_W + 6 + kShift + 1 + 3) / 4;      // words per staged row (the byte after the last one feeds tileB)
  static_assert(kWords * 4 <= S, "staged row fits the tile stride");
  {
    const uint32_t* plane = (const uint32_t*)(D.pyr + (size_t)blockIdx.y * D.pyr_frame + L.pyr_off);
    const int ws = L.pstride / 4, last_row = L.h + 2 * kBorder - 1;
    const int w0 = (2 * kBorder - 3 + tx * kFT_W) / 4;          // first word (exact: the byte offset is 32 + 64 tx)
    const int r0 = kBorder - 3 + y0;                             // padded row of the tile's first staged row
    for (int i = threadIdx.x; i < (kFT_H + 6) * kWords; i += 256) {
      const int yy = i / kWords, j = i - yy * kWords;
      const uint32_t* row = plane + (size_t)min(r0 + yy, last_row) * ws;
      const uint32_t a = row[min(w0 + j, ws - 1)], b = row[min(w0 + j + 1, ws - 1)];
      *(uint32_t*)(tileA + yy * S + 4 * j) = a;
      *(uint32_t*)(tileB + yy * S + 4 * j) = __funnelshift_r(a, b, 8);
    }
  }
  __syncthreads();
  uint8_t* sc = D.score + (size_t)blockIdx.y * D.s_frame + L.s_off;
  const int xe = L.w - kBorder, ye = L.h - kBorder;
  const int tmin = D.min_th;
  // ring offsets (dx, dy), k = 0..15 (cv::FAST order)
  constexpr int RX[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
  constexpr int RY[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};
  // Pass 1 -- quick reject (cv::FAST's own pre-test, generalised): every arc of 9 ring pixels holds two NEIGHBOURING compass
  // points (k = 0, 4, 8, 12), so a pair of pixels neither of which has two neighbouring compass points brighter than
  // centre + minThFAST, nor two darker than centre - minThFAST, scores below minThFAST: zeros are stored at once.  The
  // surviving pairs (about a third on densely textured frames, far fewer on smooth ones) are compacted into a shared list
  // and pass 2 runs the full 16-arc evaluation on the list with all lanes busy.
  __shared__ unsigned short surv[(kFT_W / 2) * kFT_H];
  __shared__ int warp_surv[8];
  int ns;
  {
    // thread -> fixed pair column, rows tid/32 + 8 it: offsets advance by constants, no division in the loop
    const int lane = threadIdx.x & 31, xx = 2 * lane, gx = x0 + xx;
    int yy = threadIdx.x >> 5;
    int base = (yy + 3) * S + xx + 3 + kShift;
    uint8_t* o = sc + (size_t)(y0 + yy) * L.sstride + gx;
    const size_t ostep = (size_t)8 * L.sstride;
    unsigned kept = 0;                                             // bit it: the pair of row tid / 32 + 8 it survives
#pragma unroll
    for (int it = 0; it < kFT_H / 8; it++, yy += 8, base += 8 * S, o += ostep) {
      bool keep = false;
      if (gx < xe && y0 + yy < ye) {
        const unsigned c2 = expand2(*(const unsigned short*)(tileA + base));
        unsigned d[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const int off = base + RY[4 * j] * S + RX[4 * j];
          const unsigned short v = (RX[4 * j] & 1) ? *(const unsigned short*)(tileB + off - 1) : *(const unsigned short*)(tileA + off);
          d[j] = __vsub2(expand2(v), c2);
        }
        // upper bound of score + 1 from the four compass points: max over neighbouring pairs of their minimum (bright side),
        // minus the min over neighbouring pairs of their maximum (dark side)
        const unsigned br = __vimax3_s16x2(__vmins2(d[0], d[1]), __vmins2(d[1], d[2]), __vmaxs2(__vmins2(d[2], d[3]), __vmins2(d[3], d[0])));
        const unsigned dk = __vimin3_s16x2(__vmaxs2(d[0], d[1]), __vmaxs2(d[1], d[2]), __vmins2(__vmaxs2(d[2], d[3]), __vmaxs2(d[3], d[0])));
        const unsigned q = __vmaxs2(br, __vsub2(0u, dk));
        keep = (int)(short)(q & 0xffffu) > tmin || (int)(short)(q >> 16) > tmin;
        if (!keep) {
          o[0] = 0;
          if (gx + 1 < xe) o[1] = 0;
        }
      }
      kept |= (unsigned)keep << it;
    }
    // ordered-by-warp compaction without atomics: per-warp counts, one barrier, every warp sums the counts before it
    unsigned m[kFT_H / 8];
    int cnt = 0;
#pragma unroll
    for (int it = 0; it < kFT_H / 8; it++) { m[it] = __ballot_sync(0xffffffffu, (kept >> it) & 1u); cnt += __popc(m[it]); }
    if (lane == 0) warp_surv[threadIdx.x >> 5] = cnt;
    __syncthreads();
    int at = 0; ns = 0;
#pragma unroll
    for (int w = 0; w < 8; w++) { const int c = warp_surv[w]; if (w < (int)(threadIdx.x >> 5)) at += c; ns += c; }
    const unsigned below = (1u << lane) - 1u;
#pragma unroll
    for (int it = 0; it < kFT_H / 8; it++) {
      if ((kept >> it) & 1u) surv[at + __popc(m[it] & below)] = (unsigned short)(((threadIdx.x >> 5) + 8 * it) * (kFT_W / 2) + lane);
      at += __popc(m[it]);
    }
  }
  __syncthreads();
#pragma unroll 1
  for (int j = threadIdx.x; j < ns; j += 256) {
    const int i = surv[j];
    const int yy = i / (kFT_W / 2), xx = (i % (kFT_W / 2)) * 2;   // even column inside the tile
    const int gx = x0 + xx, gy = y0 + yy;
    const int base = (yy + 3) * S + xx + 3 + kShift;               // even byte offset of the centre pair (tile column xx + 3)
    const unsigned c2 = expand2(*(const unsigned short*)(tileA + base));
    unsigned d[16];
#pragma unroll
    for (int k = 0; k < 16; k++) {
      const int off = base + RY[k] * S + RX[k];                    // parity known at compile time: RX odd -> odd offset
      const unsigned short v = (RX[k] & 1) ? *(const unsigned short*)(tileB + off - 1) : *(const unsigned short*)(tileA + off);
      d[k] = __vsub2(expand2(v), c2);
    }
    unsigned m3[16], M3[16];
#pragma unroll
    for (int k = 0; k < 16; k++) {
      m3[k] = __vimin3_s16x2(d[k], d[(k + 1) & 15], d[(k + 2) & 15]);
      M3[k] = __vimax3_s16x2(d[k], d[(k + 1) & 15], d[(k + 2) & 15]);
    }
    unsigned bb = 0x80008000u, dd = 0x7fff7fffu;                   // max of arc-minima (bright), min of arc-maxima (dark)
#pragma unroll
    for (int k = 0; k < 16; k += 2) {
      const unsigned a0 = __vimin3_s16x2(m3[k], m3[(k + 3) & 15], m3[(k + 6) & 15]);
      const unsigned a1 = __vimin3_s16x2(m3[k + 1], m3[(k + 4) & 15], m3[(k + 7) & 15]);
      bb = __vimax3_s16x2(bb, a0, a1);
      const unsigned b0 = __vimax3_s16x2(M3[k], M3[(k + 3) & 15], M3[(k + 6) & 15]);
      const unsigned b1 = __vimax3_s16x2(M3[k + 1], M3[(k + 4) & 15], M3[(k + 7) & 15]);
      dd = __vimin3_s16x2(dd, b0, b1);
    }
    const unsigned best = __vmaxs2(bb, __vsub2(0u, dd));           // per lane: max(bright, dark)
    int s0 = (int)(short)(best & 0xffffu) - 1, s1 = (int)(short)(best >> 16) - 1;
    s0 = s0 < tmin ? 0 : min(s0, 255);
    s1 = s1 < tmin ? 0 : min(s1, 255);
    uint8_t* o = sc + (size_t)gy * L.sstride + gx;
    o[0] = (uint8_t)s0;
    if (gx + 1 < xe) o[1] = (uint8_t)s1;
  }
}

// ------------------------------------------------------------------------------------------------
// K3: per-cell non-max suppression + threshold fallback + ordered compaction.
//     ComputeKeyPointsOctTree cell loop, src/ORBextractor.cc:778-818.  One warp per cell.
//     Keypoint at threshold t  <=>  score >= t and score > all 8 neighbours inside the cell interior.
//     Cell uses iniThFAST unless that yields zero keypoints, then minThFAST.
// ------------------------------------------------------------------------------------------------
//     Two launches: k_cell_nms evaluates the maxima once and keeps them as bit masks (32 pixels per word, one word for
//     "local maximum", one for "local maximum with score >= iniThFAST"); k_cell_emit turns the masks of the threshold the
//     cell ended up with into the ordered candidate list.
constexpr int kNmsRows = 8;   // rows of a cell loaded ahead of the window (the loop is bound by the latency of these loads)
__global__ void __launch_bounds__(256) k_cell_nms(OrbDev D) {
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int c = blockIdx.x * 8 + warp;
  int f = blockIdx.y;
  if (c >= D.cells_per_frame) return;
  CellDev cd = D.cells[c];
  const LvlDev& L = D.L[cd.level];
  int ix0 = cd.x0 + 3, iy0 = cd.y0 + 3, iw = cd.x1 - cd.x0 - 6, ih = cd.y1 - cd.y0 - 6;
  uint32_t* info = D.cellinfo + (size_t)f * D.cells_per_frame + c;
  if (iw <= 0 || ih <= 0) { if (lane == 0) *info = (uint32_t)D.ini_th << 16; return; }
  // The score plane is sparse (0 below minThFAST): every lane reads its own pixel, and only the few non-zero ones look at
  // their eight neighbours (straight from the plane; neighbours outside the cell interior count as 0).
  // (row, column) of the linear index advance incrementally: no integer division by the run-time cell width
  const uint8_t* sc = D.score + (size_t)f * D.s_frame + L.s_off + (size_t)iy0 * L.sstride + ix0;
  const int ss = L.sstride;
  int cnt_hi = 0, cnt_lo = 0;
  uint2* masks = D.cellmask + ((size_t)f * D.cells_per_frame + c) * D.mask_words;
  if (iw <= 32) {
    // lane == column of the cell interior: a row is one coalesced byte load, the horizontal neighbours come through two
    // shuffles, the rows above and below from the sliding window (hm = max of left / right, h3 = max of all three), and the
    // row's ballot is appended to the linear bit stream (pixel index = row * iw + column, 32 pixels per mask word) --
    // about 25 instructions per row of up to 32 pixels, against 65 per 32 pixels for the pixel-per-lane loop below.
    // Columns >= iw and rows outside the interior read as 0, as neighbours outside the interior do there.
    const bool incol = lane < iw;
    const uint8_t* col = sc + lane;
    const int ini = D.ini_th;
    auto horiz = [&](int v, int& hm, int& h3) {
      int l = __shfl_up_sync(0xffffffffu, v, 1), r = __shfl_down_sync(0xffffffffu, v, 1);
      if (lane == 0) l = 0;
      if (lane == 31) r = 0;
      hm = max(l, r); h3 = max(hm, v);
    };
    int v = incol ? (int)col[0] : 0, hm, h3, up3 = 0;
    horiz(v, hm, h3);
    unsigned long long lo_buf = 0, hi_buf = 0;
    int pos = 0, widx = 0;
    for (int y0 = 0; y0 < ih; y0 += kNmsRows) {
      int nv[kNmsRows];
#pragma unroll
      for (int u = 0; u < kNmsRows; u++) nv[u] = (incol && y0 + 1 + u < ih) ? (int)col[(size_t)(y0 + 1 + u) * ss] : 0;
#pragma unroll
      for (int u = 0; u < kNmsRows; u++) {
        if (y0 + u >= ih) break;
        int nhm, nh3;
        horiz(nv[u], nhm, nh3);
        const bool ismax = v > max(hm, max(up3, nh3));
        const uint32_t m_lo = __ballot_sync(0xffffffffu, ismax), m_hi = __ballot_sync(0xffffffffu, ismax && v >= ini);
        cnt_lo += __popc(m_lo);
        cnt_hi += __popc(m_hi);
        lo_buf |= (unsigned long long)m_lo << pos;
        hi_buf |= (unsigned long long)m_hi << pos;
        pos += iw;
        if (pos >= 32) {
          if (lane == 0) masks[widx] = make_uint2((uint32_t)lo_buf, (uint32_t)hi_buf);
          widx++; lo_buf >>= 32; hi_buf >>= 32; pos -= 32;
        }
        up3 = h3; v = nv[u]; hm = nhm; h3 = nh3;
      }
    }
    if (pos > 0 && lane == 0) masks[widx] = make_uint2((uint32_t)lo_buf, (uint32_t)hi_buf);
  } else {
  const int npx = iw * ih;
  int yy = lane / iw, xx = lane - yy * iw;
  for (int i0 = 0; i0 < npx; i0 += 32) {
    int i = i0 + lane;
    bool ismax = false; int v = 0;
    if (i < npx) {
      const uint8_t* p = sc + (size_t)yy * ss + xx;
      v = p[0];
      if (v > 0) {
        const bool l = xx > 0, r = xx < iw - 1, u = yy > 0, d = yy < ih - 1;
        int m = 0;
        if (l) m = max(m, (int)p[-1]);
        if (r) m = max(m, (int)p[1]);
        if (u) {
          m = max(m, (int)p[-ss]);
          if (l) m = max(m, (int)p[-ss - 1]);
          if (r) m = max(m, (int)p[-ss + 1]);
        }
        if (d) {
          m = max(m, (int)p[ss]);
          if (l) m = max(m, (int)p[ss - 1]);
          if (r) m = max(m, (int)p[ss + 1]);
        }
        ismax = v > m;
      }
    }
    const uint32_t m_lo = __ballot_sync(0xffffffffu, ismax), m_hi = __ballot_sync(0xffffffffu, ismax && v >= D.ini_th);
    cnt_lo += __popc(m_lo);
    cnt_hi += __popc(m_hi);
    if (lane == 0) masks[i0 >> 5] = make_uint2(m_lo, m_hi);
    xx += 32;
    while (xx >= iw) { xx -= iw; yy++; }
  }
  }
  if (lane == 0) {
    int t = cnt_hi > 0 ? D.ini_th : D.min_th;
    int n = cnt_hi > 0 ? cnt_hi : cnt_lo;
    if (n > 0xFFFF) { n = 0xFFFF; atomicExch(D.err, SDPL_ERR_OVERFLOW); }
    *info = ((uint32_t)t << 16) | (uint32_t)n;
  }
}

__global__ void __launch_bounds__(256) k_cell_emit(OrbDev D) {
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int c = blockIdx.x * 8 + warp;
  int f = blockIdx.y;
  if (c >= D.cells_per_frame) return;
  CellDev cd = D.cells[c];
  const LvlDev& L = D.L[cd.level];
  const int ix0 = cd.x0 + 3, iy0 = cd.y0 + 3, iw = cd.x1 - cd.x0 - 6, ih = cd.y1 - cd.y0 - 6;
  if (iw <= 0 || ih <= 0) return;
  const uint32_t info = D.cellinfo[(size_t)f * D.cells_per_frame + c];
  if ((info & 0xFFFFu) == 0) return;
  const bool hi = (int)(info >> 16) == D.ini_th;
  // offset of this cell = sum of the counts of the preceding cells of the same level
  int acc = 0;
  const uint32_t* lv = D.cellinfo + (size_t)f * D.cells_per_frame + L.cell_base;
  for (int i = lane; i < c - L.cell_base; i += 32) acc += (int)(lv[i] & 0xFFFF);
#pragma unroll
  for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  uint32_t base = (uint32_t)acc;
  const uint8_t* sc = D.score + (size_t)f * D.s_frame + L.s_off;
  uint32_t* oxy = D.cand_xy + (size_t)f * D.cand_per_frame + L.cand_off;
  uint8_t* ors = D.cand_resp + (size_t)f * D.cand_per_frame + L.cand_off;
  const uint2* masks = D.cellmask + ((size_t)f * D.cells_per_frame + c) * D.mask_words;
  const int nwords = (iw * ih + 31) >> 5;
  for (int g0 = 0; g0 < nwords; g0 += 32) {
    const int g = g0 + lane;
    uint32_t m = 0;
    if (g < nwords) { const uint2 mm = masks[g]; m = hi ? mm.y : mm.x; }
    // exclusive prefix of the set bits over the 32 words of this trip
    const int n = __popc(m);
    int pre = n;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, pre, o); if (lane >= o) pre += t; }
    uint32_t o = base + (uint32_t)(pre - n);
    base += (uint32_t)__shfl_sync(0xffffffffu, pre, 31);
    while (m) {
      const int bit = __ffs(m) - 1;
      m &= m - 1;
      const int i = g * 32 + bit;
      const int yy = i / iw, xx = i - yy * iw;
      if (o < (uint32_t)L.cand_cap) {
        // cell-local FAST coordinate + cell shift (j*wCell, i*hCell): border-relative coordinates
        oxy[o] = ((uint32_t)((yy + 3) + cd.sy) << 16) | (uint32_t)((xx + 3) + cd.sx);
        ors[o] = sc[(size_t)(iy0 + yy) * L.sstride + ix0 + xx];
      } else {
        atomicExch(D.err, SDPL_ERR_OVERFLOW);
      }
      ++o;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// K4: quadtree keypoint distribution.  DistributeOctTree / DivideNode, src/ORBextractor.cc:470-752.
//     One CTA per (frame, level).  The std::list of nodes is an array in list order (node id == list
//     position); every sweep rebuilds the array exactly as push_front / erase would leave the list:
//     children of the nodes split in this sweep in REVERSE creation order, then the surviving nodes in
//     their old order.  Candidates carry their node position; child sizes come from shared-memory atomics.
//     The "largest node first" phase orders (size, creation seq) descending (oracle decision i) and stops
//     at the first split that reaches N nodes.
// ------------------------------------------------------------------------------------------------
constexpr int kQT = 256;  // threads per quadtree CTA

struct QtSmem {          // carved from dynamic shared memory, capacity `cap` nodes
  short *x0[2], *x1[2], *y0[2], *y1[2];
  int* cnt[2];
  int* ccnt;             // [4*cap] tentative child counts
  int* childpos;         // [4*cap]
  int* keeppos;          // [cap]
  int* sa;               // [cap] scan A
  int* sb;               // [cap] scan B
  int* flag;             // [cap] split flag / rank
  int* order;            // [cap] rank -> position
};

// exclusive scan of a[0..n) in place (shared memory), returns total; all kQT threads call.
__device__ int block_excl_scan(int* a, int n, int* warp_tot /*[8+1]*/) {
  int per = (n + kQT - 1) / kQT;
  int b = threadIdx.x * per, e = min(b + per, n);
  int s = 0;
  for (int i = b; i < e; i++) s += a[i];
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int inc = s;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
  if (lane == 31) warp_tot[w] = inc;
  __syncthreads();
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int i = 0; i < kQT / 32; i++) { int t = warp_tot[i]; warp_tot[i] = acc; acc += t; }
    warp_tot[kQT / 32] = acc;
  }
  __syncthreads();
  int run = warp_tot[w] + inc - s;
  for (int i = b; i < e; i++) { int t = a[i]; a[i] = run; run += t; }
  int total = warp_tot[kQT / 32];
  __syncthreads();
  return total;
}

__device__ __forceinline__ int child_of(int x, int y, int x0, int x1, int y0, int y1) {
  int xm = x0 + ((x1 - x0 + 1) >> 1), ym = y0 + ((y1 - y0 + 1) >> 1);   // ceil(float(d)/2)
  return (x < xm ? 0 : 1) + (y < ym ? 0 : 2);
}

__global__ void __launch_bounds__(kQT) k_quadtree(OrbDev D, int cap) {
  extern __shared__ __align__(16) unsigned char smraw[];
  __shared__ int warp_tot[kQT / 32 + 1];
  __shared__ int sh_n, sh_misc[4];
  const int l = blockIdx.x, f = blockIdx.y, tid = threadIdx.x;
  const LvlDev& L = D.L[l];
  QtSmem S;
  {
    unsigned char* p = smraw;
    auto take = [&](size_t bytes) { unsigned char* r = p; p += (bytes + 15) / 16 * 16; return r; };
    for (int k = 0; k < 2; k++) {
      S.x0[k] = (short*)take(cap * 2); S.x1[k] = (short*)take(cap * 2);
      S.y0[k] = (short*)take(cap * 2); S.y1[k] = (short*)take(cap * 2);
      S.cnt[k] = (int*)take(cap * 4);
    }
    S.ccnt = (int*)take(cap * 16); S.childpos = (int*)take(cap * 16);
    S.keeppos = (int*)take(cap * 4); S.sa = (int*)take(cap * 4); S.sb = (int*)take(cap * 4);
    S.flag = (int*)take(cap * 4); S.order = (int*)take(cap * 4);
  }
  // number of candidates of this (frame, level) = sum of cell counts
  {
    const uint32_t* ci = D.cellinfo + (size_t)f * D.cells_per_frame + L.cell_base;
    int acc = 0;
    for (int i = tid; i < L.ncells; i += kQT) acc += (int)(ci[i] & 0xFFFF);
#pragma unroll
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((tid & 31) == 0) warp_tot[tid >> 5] = acc;
    __syncthreads();
    if (tid == 0) { int t = 0; for (int i = 0; i < kQT / 32; i++) t += warp_tot[i]; sh_n = min(t, L.cand_cap); }
    __syncthreads();
  }
  const int n = sh_n;
  const uint32_t* cxy = D.cand_xy + (size_t)f * D.cand_per_frame + L.cand_off;
  const uint8_t* crs = D.cand_resp + (size_t)f * D.cand_per_frame + L.cand_off;
  unsigned short* node = D.cand_node + (size_t)f * D.cand_per_frame + L.cand_off;
  uint32_t* oxy = D.kp_xy + (size_t)f * D.kp_per_frame + L.kp_off;
  uint8_t* ors = D.kp_resp + (size_t)f * D.kp_per_frame + L.kp_off;
  if (tid == 0) { D.n_cand[f * D.nl + l] = n; }
  if (n == 0 || L.nIni < 1) { if (tid == 0) D.n_kp[f * D.nl + l] = 0; return; }
  const int N = L.quota;
  const int W = L.w - 32, H = L.h - 32;
  int cur = 0;
  // ---- roots (ORBextractor.cc:532-575) ----
  const int nIni = L.nIni; const float hX = L.hX;
  for (int i = tid; i < nIni; i += kQT) {
    S.x0[0][i] = (short)__float2int_rz(__fmul_rn(hX, (float)i));
    S.x1[0][i] = (short)__float2int_rz(__fmul_rn(hX, (float)(i + 1)));
    S.y0[0][i] = 0; S.y1[0][i] = (short)H;
    S.cnt[0][i] = 0;
  }
  __syncthreads();
  for (int i = tid; i < n; i += kQT) {
    int x = cxy[i] & 0xFFFF;
    int r = __float2int_rz(__fdiv_rn((float)x, hX));
    r = min(max(r, 0), nIni - 1);
    node[i] = (unsigned short)r;
    atomicAdd(&S.cnt[0][r], 1);
  }
  __syncthreads();
  // drop empty roots (list erase), keep order
  for (int i = tid; i < nIni; i += kQT) S.sa[i] = S.cnt[0][i] > 0;
  __syncthreads();
  int Lsz = block_excl_scan(S.sa, nIni, warp_tot);
  for (int i = tid; i < nIni; i += kQT) {
    if (S.cnt[0][i] > 0) {
      int p = S.sa[i];
      S.x0[1][p] = S.x0[0][i]; S.x1[1][p] = S.x1[0][i]; S.y0[1][p] = S.y0[0][i]; S.y1[1][p] = S.y1[0][i];
      S.cnt[1][p] = S.cnt[0][i];
    }
    S.keeppos[i] = S.sa[i];
  }
  __syncthreads();
  for (int i = tid; i < n; i += kQT) node[i] = (unsigned short)S.keeppos[node[i]];
  cur = 1;
  __syncthreads();

  bool phase2 = false, finish = false;
  int Qprev = 0;   // nodes [0,Qprev) of the current list were created in the previous round
  int guard = 0;
  while (!finish && guard++ < 64) {
    const int prevL = Lsz;
    short *x0 = S.x0[cur], *x1 = S.x1[cur], *y0 = S.y0[cur], *y1 = S.y1[cur];
    int* cnt = S.cnt[cur];
    // which nodes are split candidates this round
    for (int p = tid; p < Lsz; p += kQT) {
      bool cand = cnt[p] > 1 && (!phase2 || p < Qprev);
      S.flag[p] = cand ? 1 : 0;
      S.ccnt[4 * p] = S.ccnt[4 * p + 1] = S.ccnt[4 * p + 2] = S.ccnt[4 * p + 3] = 0;
    }
    __syncthreads();
    for (int i = tid; i < n; i += kQT) {
      int p = node[i];
      if (S.flag[p]) {
        uint32_t xy = cxy[i];
        int c = child_of(xy & 0xFFFF, xy >> 16, x0[p], x1[p], y0[p], y1[p]);
        atomicAdd(&S.ccnt[4 * p + c], 1);
      }
    }
    __syncthreads();
    int nsplit_total = 0;
    if (!phase2) {
      // every multi-point node is split, in list order
      for (int p = tid; p < Lsz; p += kQT) {
        int nc = 0;
        if (S.flag[p]) for (int c = 0; c < 4; c++) nc += S.ccnt[4 * p + c] > 0;
        S.sa[p] = nc;                 // children created by p
        S.sb[p] = S.flag[p] ? 0 : 1;  // survives
        S.order[p] = p;               // processing order == list order
      }
      __syncthreads();
      nsplit_total = Lsz;             // all flagged nodes processed (rank == position)
    } else {
      // expandable nodes ordered by (size, creation seq) descending; creation seq = Qprev-1-p
      for (int p = tid; p < Lsz; p += kQT) {
        int r = -1;
        if (S.flag[p]) {
          int cp = cnt[p];
          r = 0;
          for (int q = 0; q < Qprev; q++) {
            if (!S.flag[q]) continue;
            int cq = cnt[q];
            // q precedes p if (cq, seq_q) > (cp, seq_p); seq = Qprev-1-pos  => larger seq == smaller pos
            if (cq > cp || (cq == cp && q < p)) r++;
          }
        }
        S.keeppos[p] = r;             // rank among expandables (temporarily)
      }
      __syncthreads();
      if (tid == 0) sh_misc[0] = 0;
      __syncthreads();
      for (int p = tid; p < Lsz; p += kQT) if (S.flag[p]) { S.order[S.keeppos[p]] = p; atomicAdd(&sh_misc[0], 1); }
      __syncthreads();
      const int E = sh_misc[0];
      // gain per rank, inclusive prefix, first rank reaching N
      for (int r = tid; r < E; r += kQT) {
        int p = S.order[r], nc = 0;
        for (int c = 0; c < 4; c++) nc += S.ccnt[4 * p + c] > 0;
        S.sb[r] = nc - 1;
      }
      __syncthreads();
      block_excl_scan(S.sb, E, warp_tot);   // sb[r] = gain of ranks < r
      if (tid == 0) sh_misc[1] = E;         // number of ranks processed
      __syncthreads();
      for (int r = tid; r < E; r += kQT) {
        int p = S.order[r], nc = 0;
        for (int c = 0; c < 4; c++) nc += S.ccnt[4 * p + c] > 0;
        int after = Lsz + S.sb[r] + nc - 1;  // list size after processing rank r
        if (after >= N) atomicMin(&sh_misc[1], r + 1);
      }
      __syncthreads();
      nsplit_total = sh_misc[1];
      // un-flag expandables that are not reached; build sa (children per processed rank) and sb (survivor flags)
      for (int p = tid; p < Lsz; p += kQT) {
        if (S.flag[p] && S.keeppos[p] >= nsplit_total) S.flag[p] = 0;
      }
      __syncthreads();
      for (int r = tid; r < Lsz; r += kQT) {
        int nc = 0;
        if (r < nsplit_total) { int p = S.order[r]; for (int c = 0; c < 4; c++) nc += S.ccnt[4 * p + c] > 0; }
        S.sa[r] = nc;
      }
      for (int p = tid; p < Lsz; p += kQT) S.sb[p] = S.flag[p] ? 0 : 1;
      __syncthreads();
    }
    // sa is indexed by processing rank (phase 1: rank == position); sb by position
    int Q = block_excl_scan(S.sa, phase2 ? nsplit_total : Lsz, warp_tot);
    int U = block_excl_scan(S.sb, Lsz, warp_tot);
    const int nxt = cur ^ 1;
    if (tid == 0) sh_misc[2] = 0;
    __syncthreads();
    int nexp_local = 0;
    const int nrank = phase2 ? nsplit_total : Lsz;
    for (int r = tid; r < nrank; r += kQT) {
      int p = S.order[r];
      if (!S.flag[p]) continue;
      int q = S.sa[r];
      int px0 = x0[p], px1 = x1[p], py0 = y0[p], py1 = y1[p];
      int xm = px0 + ((px1 - px0 + 1) >> 1), ym = py0 + ((py1 - py0 + 1) >> 1);
      for (int c = 0; c < 4; c++) {
        int cc = S.ccnt[4 * p + c];
        if (cc == 0) { S.childpos[4 * p + c] = -1; continue; }
        int pos = Q - 1 - q; q++;
        S.childpos[4 * p + c] = pos;
        S.x0[nxt][pos] = (short)((c & 1) ? xm : px0); S.x1[nxt][pos] = (short)((c & 1) ? px1 : xm);
        S.y0[nxt][pos] = (short)((c & 2) ? ym : py0); S.y1[nxt][pos] = (short)((c & 2) ? py1 : ym);
        S.cnt[nxt][pos] = cc;
        if (cc > 1) nexp_local++;
      }
    }
    for (int p = tid; p < Lsz; p += kQT) {
      if (S.flag[p]) continue;
      int pos = Q + S.sb[p];
      S.keeppos[p] = pos;
      S.x0[nxt][pos] = x0[p]; S.x1[nxt][pos] = x1[p]; S.y0[nxt][pos] = y0[p]; S.y1[nxt][pos] = y1[p];
      S.cnt[nxt][pos] = cnt[p];
    }
    if (nexp_local) atomicAdd(&sh_misc[2], nexp_local);
    __syncthreads();
    for (int i = tid; i < n; i += kQT) {
      int p = node[i];
      if (S.flag[p]) {
        uint32_t xy = cxy[i];
        int c = child_of(xy & 0xFFFF, xy >> 16, x0[p], x1[p], y0[p], y1[p]);
        node[i] = (unsigned short)S.childpos[4 * p + c];
      } else {
        node[i] = (unsigned short)S.keeppos[p];
      }
    }
    const int nExpand = sh_misc[2];
    __syncthreads();
    Lsz = Q + U;
    Qprev = Q;
    cur = nxt;
    if (Lsz > cap - 4) { if (tid == 0) atomicExch(D.err, SDPL_ERR_OVERFLOW); break; }
    if (Lsz >= N || Lsz == prevL) finish = true;
    else if (!phase2 && Lsz + nExpand * 3 > N) phase2 = true;
  }
  // ---- best response per node, first maximum wins (ORBextractor.cc:733-749) ----
  unsigned int* best = (unsigned int*)S.ccnt;
  for (int p = tid; p < Lsz; p += kQT) best[p] = 0;
  __syncthreads();
  for (int i = tid; i < n; i += kQT) {
    unsigned int key = ((unsigned int)crs[i] << 24) | (0xFFFFFFu - (unsigned int)i);
    atomicMax(&best[node[i]], key);
  }
  __syncthreads();
  for (int p = tid; p < Lsz; p += kQT) {
    if (p >= L.kp_cap) { atomicExch(D.err, SDPL_ERR_OVERFLOW); continue; }
    unsigned int i = 0xFFFFFFu - (best[p] & 0xFFFFFFu);
    oxy[p] = cxy[i];
    ors[p] = crs[i];
  }
  if (tid == 0) D.n_kp[f * D.nl + l] = min(Lsz, L.kp_cap);
}

// ------------------------------------------------------------------------------------------------
// K6: GaussianBlur 7x7 sigma 2, BORDER_REFLECT_101 on the un-padded level (ORBextractor.cc:1083-1084).
//     Fixed point: Q8.8 kernel [18,34,48,56,48,34,18], horizontal Q8.8, vertical Q16.16, (v + 2^15) >> 16.
// ------------------------------------------------------------------------------------------------
//     The padded pyramid plane already holds the reflect-101 border the blur needs (19 >= 3 px), so the kernel reads it as
//     aligned 32-bit words with no border arithmetic, and the blurred plane uses the same padded geometry (aligned word
//     stores).  A thread owns one word = four columns and walks down kBlurRows rows: the seven taps of the horizontal pass
//     are two dp4a on byte windows cut from (previous | own | next) word -- the neighbours' words come by shuffle --, the
//     seven horizontal results of each column slide through registers for the vertical pass.  One warp = 128 columns.
constexpr int kBlurRows = 32;
__global__ void __launch_bounds__(128) k_blur7(OrbDev D) {
  const int lane = threadIdx.x & 31;
  int t = blockIdx.x * 4 + (threadIdx.x >> 5), l = 0;
  if (t >= D.blur_tiles) return;
#pragma unroll 1
  for (int i = 1; i < D.nl; i++) if (t >= D.L[i].blur_tile_base) l = i;
  const LvlDev& L = D.L[l];
  t -= L.blur_tile_base;
  const int tx = t % L.blur_tiles_x, ty = t / L.blur_tiles_x;
  // words of a padded row that overlap the interior: px 16.. (word 4) to px kBorder + w - 1
  const int w_first = kBorder / 4, w_last = (kBorder + L.w - 1) / 4;
  const int word = min(w_first + tx * 32 + lane, w_last + 1);              // lanes past the end feed their left neighbour only
  const bool store = w_first + tx * 32 + lane <= w_last;
  const int y0 = ty * kBlurRows, nrows = min(kBlurRows, L.h - y0);
  const size_t plane = (size_t)blockIdx.y * D.pyr_frame + L.pyr_off;
  const int ws = L.pstride / 4;                                            // words per row
  const uint32_t* in = (const uint32_t*)(D.pyr + plane) + (size_t)(kBorder + y0 - 3) * ws + word;
  uint32_t* out = (uint32_t*)(D.blur + plane) + (size_t)(kBorder + y0) * ws + word;
  const uint32_t K_lo = 18u | (34u << 8) | (48u << 16) | (56u << 24), K_hi = 48u | (34u << 8) | (18u << 16);
  uint32_t hz[7][4];
#pragma unroll
  for (int k = 0; k < 7; k++) { hz[k][0] = hz[k][1] = hz[k][2] = hz[k][3] = 0; }
#pragma unroll 1
  for (int i0 = 0; i0 < nrows + 6; i0 += 7) {
#pragma unroll
    for (int k = 0; k < 7; k++) {
      const int i = i0 + k;
      if (i < nrows + 6) {                                                 // warp-uniform
        const uint32_t cur = in[(size_t)i * ws];
        uint32_t prev = __shfl_up_sync(0xffffffffu, cur, 1), next = __shfl_down_sync(0xffffffffu, cur, 1);
        if (lane == 0) prev = in[(size_t)i * ws - 1];
        if (lane == 31) next = in[(size_t)i * ws + 1];
        hz[k][0] = __dp4a(__funnelshift_r(prev, cur, 8), K_lo, __dp4a(__funnelshift_r(cur, next, 8), K_hi, 0u));
        hz[k][1] = __dp4a(__funnelshift_r(prev, cur, 16), K_lo, __dp4a(__funnelshift_r(cur, next, 16), K_hi, 0u));
        hz[k][2] = __dp4a(__funnelshift_r(prev, cur, 24), K_lo, __dp4a(__funnelshift_r(cur, next, 24), K_hi, 0u));
        hz[k][3] = __dp4a(cur, K_lo, __dp4a(next, K_hi, 0u));
        if (i >= 6) {
          uint32_t v[4];
#pragma unroll
          for (int c = 0; c < 4; c++) {
            v[c] = 56u * hz[(k + 4) % 7][c] + (1u << 15);
            v[c] += 18u * (hz[(k + 1) % 7][c] + hz[k][c]);
            v[c] += 34u * (hz[(k + 2) % 7][c] + hz[(k + 6) % 7][c]);
            v[c] += 48u * (hz[(k + 3) % 7][c] + hz[(k + 5) % 7][c]);
          }
          // byte 2 of every sum (sum of the weights is 2^16: the result cannot exceed 255)
          const uint32_t lo = __byte_perm(v[0], v[1], 0x0062), hi = __byte_perm(v[2], v[3], 0x0062);
          if (store) out[(size_t)(i - 6) * ws] = __byte_perm(lo, hi, 0x5410);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// K5 + K7: IC_Angle orientation (ORBextractor.cc:66-93) and rotated-BRIEF descriptor (:97-136), one warp per
//          keypoint; writes the final cv::KeyPoint record (ORBextractor.cc:826-836, 1099-1105).
// ------------------------------------------------------------------------------------------------
__device__ __align__(16) const signed char g_pattern[1024] = {
#include "../../include/sdpl_orb_pattern.inc"
};

__global__ void __launch_bounds__(256) k_orient_describe(OrbDev D, sdpl_keypoint* __restrict__ kps, uint8_t* __restrict__ desc,
                                                          int capacity, int* __restrict__ n_out) {
  const int f = blockIdx.y, lane = threadIdx.x & 31;
  int slot = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int* nk = D.n_kp + f * D.nl;
  if (slot == 0 && lane == 0) {
    int t = 0;
    for (int i = 0; i < D.nl; i++) t += nk[i];
    n_out[f] = t;
  }
  if (slot >= D.kp_per_frame) return;
  int l = 0;
#pragma unroll 1
  for (int i = 1; i < D.nl; i++) if (slot >= D.L[i].kp_off) l = i;
  const LvlDev& L = D.L[l];
  int idx = slot - L.kp_off;
  if (idx >= nk[l]) return;
  int obase = 0;
  for (int i = 0; i < l; i++) obase += nk[i];
  const int o = obase + idx;
  if (o >= capacity) return;  // host reports SDPL_ERR_CAPACITY from n_out
  uint32_t xy = D.kp_xy[(size_t)f * D.kp_per_frame + slot];
  int resp = D.kp_resp[(size_t)f * D.kp_per_frame + slot];
  const int kx = (int)(xy & 0xFFFF) + 16, ky = (int)(xy >> 16) + 16;   // + minBorderX/Y
  // ---- orientation on the un-blurred level ----
  const uint8_t* c = D.pyr + (size_t)f * D.pyr_frame + L.pyr_off + (size_t)(kBorder + ky) * L.pstride + kBorder + kx;
  int m10 = 0, m01 = 0;
  if (lane < 31) {
    // lanes over the 31 columns of the patch, rows in sequence: every load of the warp is one contiguous 31-byte segment
    // (integer moments: any summation order gives the reference's m10 / m01)
    const int u = lane - 15, au = u < 0 ? -u : u;
#pragma unroll
    for (int v = -15; v <= 15; v++) {
      if (au <= D.umax[v < 0 ? -v : v]) {
        const int p = c[(ptrdiff_t)v * L.pstride + u];
        m10 += u * p; m01 += v * p;
      }
    }
  }
#pragma unroll
  for (int o2 = 16; o2; o2 >>= 1) { m10 += __shfl_xor_sync(0xffffffffu, m10, o2); m01 += __shfl_xor_sync(0xffffffffu, m01, o2); }
  const float angle = fast_atan2_deg((float)m01, (float)m10);
  // ---- descriptor on the blurred level ----
  const float factorPI = (float)(3.14159265358979323846 / 180.f);
  const float ar = __fmul_rn(angle, factorPI);
  const float a = (float)cos((double)ar), b = (float)sin((double)ar);
  const uint8_t* cb = D.blur + (size_t)f * D.pyr_frame + L.pyr_off + (size_t)(kBorder + ky) * L.pstride + kBorder + kx;
  // the lane's eight tests = 32 signed bytes of the pattern, fetched as two 16-byte loads
  uint32_t pw[8];
  {
    const uint4* p4 = (const uint4*)(g_pattern + lane * 32);
    const uint4 p0 = p4[0], p1 = p4[1];
    pw[0] = p0.x; pw[1] = p0.y; pw[2] = p0.z; pw[3] = p0.w; pw[4] = p1.x; pw[5] = p1.y; pw[6] = p1.z; pw[7] = p1.w;
  }
  int val = 0;
#pragma unroll
  for (int t = 0; t < 8; t++) {
    const int px0 = (int)(signed char)(pw[t] & 0xffu), py0 = (int)(signed char)((pw[t] >> 8) & 0xffu);
    const int px1 = (int)(signed char)((pw[t] >> 16) & 0xffu), py1 = (int)(signed char)(pw[t] >> 24);
    int r0 = cv_round_f(__fadd_rn(__fmul_rn((float)px0, b), __fmul_rn((float)py0, a)));
    int c0 = cv_round_f(__fsub_rn(__fmul_rn((float)px0, a), __fmul_rn((float)py0, b)));
    int r1 = cv_round_f(__fadd_rn(__fmul_rn((float)px1, b), __fmul_rn((float)py1, a)));
    int c1 = cv_round_f(__fsub_rn(__fmul_rn((float)px1, a), __fmul_rn((float)py1, b)));
    int t0 = cb[(ptrdiff_t)r0 * L.pstride + c0], t1 = cb[(ptrdiff_t)r1 * L.pstride + c1];
    val |= (t0 < t1) << t;
  }
  desc[((size_t)f * capacity + o) * 32 + lane] = (uint8_t)val;
  if (lane == 0) {
    sdpl_keypoint k;
    float fx = (float)kx, fy = (float)ky;
    if (l != 0) { fx = __fmul_rn(fx, D.sf[l]); fy = __fmul_rn(fy, D.sf[l]); }
    k.x = fx; k.y = fy;
    k.size = (float)(int)__fmul_rn(31.f, D.sf[l]);
    k.angle = angle; k.response = (float)resp; k.octave = l; k.class_id = -1;
    kps[(size_t)f * capacity + o] = k;
  }
}

}  // namespace sdpl

// =====================================================================================================
// Host side
// =====================================================================================================
using namespace sdpl;

struct sdpl_orb {
  int nfeatures, nlevels, ini_th, min_th, device;
  float scale;
  std::vector<float> sf, isf, s2, is2;
  std::vector<int> quota;
  int umax[16];
  cudaStream_t own_stream = nullptr, stream = nullptr;
  // geometry-dependent state
  int gw = 0, gh = 0, gB = 0;
  OrbDev D;
  std::vector<CellDev> cells;
  DevBuf pyr, score, blur, cellinfo, cellmask, cand_xy, cand_resp, cand_node, n_cand, n_kp, kp_xy, kp_resp, cellsdev, tables, err;
  DevBuf in_stage, out_kps, out_desc, out_n;
  void* h_stage = nullptr; size_t h_stage_bytes = 0;   // pinned
  int fast_tiles = 0, blur_tiles = 0, max_quota = 0, kp_cap_total = 0, qt_cap = 64;
  int last_B = 0, launches = 0;
  int last_capacity = 0;
  StageTimer timer;
};

static int orb_setup(sdpl_orb* o, int w, int h, int B) {
  if (o->gw == w && o->gh == h && o->gB >= B) { o->D.B = B; return SDPL_OK; }
  if (o->gw == w && o->gh == h) B = std::max(B, o->gB);
  // a new geometry or a bigger batch rewrites tables, arenas and the overflow flag: nothing of this handle may be in flight
  if (o->gw) cudaStreamSynchronize(o->stream);
  const int nl = o->nlevels;
  OrbDev& D = o->D;
  memset(&D, 0, sizeof(D));
  D.nl = nl; D.ini_th = o->ini_th; D.min_th = o->min_th;
  memcpy(D.umax, o->umax, sizeof(D.umax));
  for (int l = 0; l < nl; l++) D.sf[l] = o->sf[l];
  size_t pyr_off = 0, s_off = 0;
  int cell_base = 0, cand_off = 0, kp_off = 0, ft = 0, bt = 0, mask_words = 1;
  o->cells.clear();
  std::vector<unsigned short> tab_u16;   // all resize tables packed: per level xofs,yofs (u16) then xa,ya (short2)
  std::vector<short> tab_s16;
  std::vector<size_t> xofs_at(nl), yofs_at(nl), xa_at(nl), ya_at(nl);
  o->max_quota = 0;
  for (int l = 0; l < nl; l++) {
    LvlDev& L = D.L[l];
    L.w = h_cv_round((float)w * o->isf[l]);
    L.h = h_cv_round((float)h * o->isf[l]);
    if (L.w < 1 || L.h < 1 || L.w > 32000 || L.h > 32000) { set_last_error("pyramid level size out of range"); return SDPL_ERR_ARG; }
    L.pstride = (int)align_up(L.w + 2 * kBorder, 4);
    L.sstride = (int)align_up(L.w, 4);
    L.pyr_off = pyr_off; pyr_off += align_up((size_t)L.pstride * (L.h + 2 * kBorder), 16);
    L.s_off = s_off; s_off += align_up((size_t)L.sstride * L.h, 16);
    L.quota = o->quota[l];
    o->max_quota = std::max(o->max_quota, L.quota);
    // cells (ComputeKeyPointsOctTree, ORBextractor.cc:762-796)
    const int minB = kBorder - 3, maxBX = L.w - kBorder + 3, maxBY = L.h - kBorder + 3;
    const float width = (float)(maxBX - minB), height = (float)(maxBY - minB);
    const int nCols = (int)(width / 30.f), nRows = (int)(height / 30.f);
    L.cell_base = cell_base; L.ncells = 0;
    L.nIni = 0; L.hX = 1.f;
    if (nCols >= 1 && nRows >= 1 && width > 0 && height > 0) {
      const int wCell = (int)ceilf(width / nCols), hCell = (int)ceilf(height / nRows);
      if (wCell + 2 > kCellApron || hCell + 2 > kCellApron) { set_last_error("cell larger than NMS tile"); return SDPL_ERR_UNSUPPORTED; }
      for (int i = 0; i < nRows; i++) {
        const float iniY = (float)(minB + i * hCell);
        float maxY = iniY + hCell + 6;
        if (iniY >= maxBY - 3) continue;
        if (maxY > maxBY) maxY = (float)maxBY;
        for (int j = 0; j < nCols; j++) {
          const float iniX = (float)(minB + j * wCell);
          float maxX = iniX + wCell + 6;
          if (iniX >= maxBX - 6) continue;
          if (maxX > maxBX) maxX = (float)maxBX;
          CellDev c;
          c.level = (short)l; c.x0 = (short)(int)iniX; c.y0 = (short)(int)iniY; c.x1 = (short)(int)maxX; c.y1 = (short)(int)maxY;
          c.sx = (short)(j * wCell); c.sy = (short)(i * hCell); c.pad = 0;
          o->cells.push_back(c);
          mask_words = std::max(mask_words, (((int)maxX - (int)iniX - 6) * ((int)maxY - (int)iniY - 6) + 31) / 32);
          L.ncells++;
        }
      }
      L.nIni = (int)roundf((float)(maxBX - minB) / (float)(maxBY - minB));
      if (L.nIni < 1) { set_last_error("image aspect ratio gives zero quadtree roots (reference divides by zero)"); return SDPL_ERR_UNSUPPORTED; }
      L.hX = (float)(maxBX - minB) / (float)L.nIni;
    }
    cell_base += L.ncells;
    // worst case after NMS: no two keypoints are 8-adjacent -> <= ceil(w/2)*ceil(h/2); budget area/6 (+slack)
    L.cand_cap = (int)std::min<size_t>((size_t)L.w * L.h / 6 + 1024, (size_t)0xFFFFFF);
    L.cand_off = cand_off; cand_off += (int)align_up(L.cand_cap, 4);
    L.kp_cap = L.quota + 8 + 4 * std::max(L.nIni, 1);
    L.kp_off = kp_off; kp_off += L.kp_cap;
    int dw = L.w - 2 * kBorder, dh = L.h - 2 * kBorder;
    L.fast_tiles_x = dw > 0 ? div_up(dw, kFT_W) : 0;
    int fty = dh > 0 ? div_up(dh, kFT_H) : 0;
    L.fast_tile_base = ft; ft += L.fast_tiles_x * fty;
    L.blur_tiles_x = div_up((kBorder + L.w - 1) / 4 - kBorder / 4 + 1, 32);
    L.blur_tile_base = bt; bt += L.blur_tiles_x * div_up(L.h, kBlurRows);
    // resize tables for level l from level l-1 (cv::resize INTER_LINEAR, fixed-point Q11)
    L.area_fast = 0;
    if (l > 0) {
      const int sw = D.L[l - 1].w, sh = D.L[l - 1].h;
      double scale_x = 1. / ((double)L.w / sw), scale_y = 1. / ((double)L.h / sh);
      int isx = h_cv_round(scale_x), isy = h_cv_round(scale_y);
      if (fabs(scale_x - isx) < 2.220446049250313e-16 && fabs(scale_y - isy) < 2.220446049250313e-16 && isx == 2 && isy == 2)
        L.area_fast = 1;
      xofs_at[l] = tab_u16.size();
      xa_at[l] = tab_s16.size();
      for (int dx = 0; dx < L.w; dx++) {
        float fx = (float)((dx + 0.5) * scale_x - 0.5);
        int sx = h_cv_floor(fx);
        fx -= sx;
        if (sx < 0) { fx = 0; sx = 0; }
        if (sx >= sw - 1) { fx = 0; sx = sw - 1; }
        tab_u16.push_back((unsigned short)sx);
        tab_s16.push_back((short)std::min(32767, h_cv_round((1.f - fx) * 2048)));
        tab_s16.push_back((short)std::min(32767, h_cv_round(fx * 2048)));
      }
      yofs_at[l] = tab_u16.size();
      ya_at[l] = tab_s16.size();
      for (int dy = 0; dy < L.h; dy++) {
        float fy = (float)((dy + 0.5) * scale_y - 0.5);
        int sy = h_cv_floor(fy);
        fy -= sy;
        tab_u16.push_back((unsigned short)(short)sy);
        tab_s16.push_back((short)std::min(32767, h_cv_round((1.f - fy) * 2048)));
        tab_s16.push_back((short)std::min(32767, h_cv_round(fy * 2048)));
      }
    }
  }
  D.pyr_frame = align_up(pyr_off, 256); D.s_frame = align_up(s_off, 256);
  D.cells_per_frame = cell_base; D.cand_per_frame = cand_off; D.kp_per_frame = kp_off; D.mask_words = mask_words;
  o->fast_tiles = ft; o->blur_tiles = bt; D.blur_tiles = bt; o->kp_cap_total = kp_off;
  o->qt_cap = 64;
  for (int l = 0; l < nl; l++) o->qt_cap = std::max(o->qt_cap, D.L[l].kp_cap + 8);
  int rc;
  if ((rc = o->pyr.reserve(D.pyr_frame * B))) return rc;
  if ((rc = o->score.reserve(D.s_frame * B))) return rc;
  if ((rc = o->blur.reserve(D.pyr_frame * B))) return rc;
  if ((rc = o->cellinfo.reserve(sizeof(uint32_t) * (size_t)std::max(1, cell_base) * B))) return rc;
  if ((rc = o->cellmask.reserve(sizeof(uint2) * (size_t)std::max(1, cell_base) * std::max(1, D.mask_words) * B))) return rc;
  if ((rc = o->cand_xy.reserve(sizeof(uint32_t) * (size_t)cand_off * B))) return rc;
  if ((rc = o->cand_resp.reserve((size_t)cand_off * B))) return rc;
  if ((rc = o->cand_node.reserve(sizeof(unsigned short) * (size_t)cand_off * B))) return rc;
  if ((rc = o->n_cand.reserve(sizeof(int) * nl * B))) return rc;
  if ((rc = o->n_kp.reserve(sizeof(int) * nl * B))) return rc;
  if ((rc = o->kp_xy.reserve(sizeof(uint32_t) * (size_t)kp_off * B))) return rc;
  if ((rc = o->kp_resp.reserve((size_t)kp_off * B))) return rc;
  if ((rc = o->cellsdev.reserve(sizeof(CellDev) * std::max<size_t>(1, o->cells.size())))) return rc;
  if ((rc = o->err.reserve(sizeof(int)))) return rc;
  size_t u16_bytes = align_up(tab_u16.size() * 2, 16);
  if ((rc = o->tables.reserve(u16_bytes + tab_s16.size() * 2 + 16))) return rc;
  if (!o->cells.empty()) SDPL_CUDA(cudaMemcpy(o->cellsdev.p, o->cells.data(), sizeof(CellDev) * o->cells.size(), cudaMemcpyHostToDevice));
  if (!tab_u16.empty()) {
    SDPL_CUDA(cudaMemcpy(o->tables.p, tab_u16.data(), tab_u16.size() * 2, cudaMemcpyHostToDevice));
    SDPL_CUDA(cudaMemcpy((char*)o->tables.p + u16_bytes, tab_s16.data(), tab_s16.size() * 2, cudaMemcpyHostToDevice));
  }
  SDPL_CUDA(cudaMemset(o->err.p, 0, sizeof(int)));
  for (int l = 1; l < nl; l++) {
    LvlDev& L = D.L[l];
    L.xofs = o->tables.as<unsigned short>() + xofs_at[l];
    L.yofs = o->tables.as<unsigned short>() + yofs_at[l];
    L.xa = (const short2*)((char*)o->tables.p + u16_bytes + xa_at[l] * 2);
    L.ya = (const short2*)((char*)o->tables.p + u16_bytes + ya_at[l] * 2);
  }
  D.pyr = o->pyr.as<uint8_t>(); D.score = o->score.as<uint8_t>(); D.blur = o->blur.as<uint8_t>();
  D.cellinfo = o->cellinfo.as<uint32_t>(); D.cellmask = o->cellmask.as<uint2>(); D.cand_xy = o->cand_xy.as<uint32_t>(); D.cand_resp = o->cand_resp.as<uint8_t>();
  D.cand_node = o->cand_node.as<unsigned short>(); D.n_cand = o->n_cand.as<int>(); D.n_kp = o->n_kp.as<int>();
  D.kp_xy = o->kp_xy.as<uint32_t>(); D.kp_resp = o->kp_resp.as<uint8_t>(); D.cells = o->cellsdev.as<CellDev>();
  D.err = o->err.as<int>();
  D.B = B;
  o->gw = w; o->gh = h; o->gB = B;
  return SDPL_OK;
}

static size_t qt_smem_bytes(int cap) {
  auto a16 = [](size_t b) { return (b + 15) / 16 * 16; };
  return 2 * (4 * a16((size_t)cap * 2) + a16((size_t)cap * 4)) + 2 * a16((size_t)cap * 16) + 5 * a16((size_t)cap * 4);
}

// Runs the whole pipeline for B frames resident on the device.  Asynchronous on o->stream.
static int orb_run_dev(sdpl_orb* o, const uint8_t* d_imgs, int B, int w, int h, int stride, size_t frame_stride,
                       sdpl_keypoint* d_kps, uint8_t* d_desc, int capacity, int* d_n_out) {
  int rc = orb_setup(o, w, h, B);
  if (rc) return rc;
  OrbDev D = o->D;
  D.B = B;
  cudaStream_t st = o->stream;
  const int nl = o->nlevels;
  o->timer.begin(st);
  {
    const LvlDev& L = D.L[0];
    dim3 g(div_up(L.pstride / 4, 128), div_up(L.h + 2 * kBorder, kBaseRows), B);
    k_pyr_base<<<g, 128, 0, st>>>(d_imgs, stride, frame_stride, D.pyr + L.pyr_off, L.w, L.h, L.pstride, D.pyr_frame);
    SDPL_LAUNCH_CHECK();
  }
  for (int l = 1; l < nl; l++) {
    const LvlDev& L = D.L[l];
    dim3 g(div_up(L.pstride / 4, 128), div_up(L.h + 2 * kBorder, kPyrRows), B);
    k_pyr_resize<<<g, 128, 0, st>>>(D, l);
    SDPL_LAUNCH_CHECK();
  }
  o->timer.mark(st, "pyramid");
  if (o->fast_tiles > 0) {
    k_fast_score<<<dim3(o->fast_tiles, B), 256, 0, st>>>(D, o->fast_tiles);
    SDPL_LAUNCH_CHECK();
  }
  o->timer.mark(st, "fast_score");
  if (D.cells_per_frame > 0) {
    dim3 g(div_up(D.cells_per_frame, 8), B);
    k_cell_nms<<<g, 256, 0, st>>>(D);
    SDPL_LAUNCH_CHECK();
    k_cell_emit<<<g, 256, 0, st>>>(D);
    SDPL_LAUNCH_CHECK();
  }
  o->timer.mark(st, "cell_nms");
  {
    int cap = o->qt_cap;
    size_t smem = qt_smem_bytes(cap);
    if (smem > 200 * 1024) { set_last_error("per-level feature quota too large for the quadtree kernel"); return SDPL_ERR_UNSUPPORTED; }
    if (smem > 48 * 1024) SDPL_CUDA(cudaFuncSetAttribute(k_quadtree, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_quadtree<<<dim3(nl, B), kQT, smem, st>>>(D, cap);
    SDPL_LAUNCH_CHECK();
  }
  o->timer.mark(st, "quadtree");
  k_blur7<<<dim3(div_up(o->blur_tiles, 4), B), 128, 0, st>>>(D);
  SDPL_LAUNCH_CHECK();
  o->timer.mark(st, "blur7");
  k_orient_describe<<<dim3(div_up(D.kp_per_frame, 8), B), 256, 0, st>>>(D, d_kps, d_desc, capacity, d_n_out);
  SDPL_LAUNCH_CHECK();
  o->timer.mark(st, "orient_describe");
  o->last_B = B; o->last_capacity = capacity;
  return SDPL_OK;
}

static int orb_check_err(sdpl_orb* o) {
  int e = 0;
  SDPL_CUDA(cudaMemcpyAsync(&e, o->err.p, sizeof(int), cudaMemcpyDeviceToHost, o->stream));
  SDPL_CUDA(cudaStreamSynchronize(o->stream));
  if (e) {
    cudaMemsetAsync(o->err.p, 0, sizeof(int), o->stream);
    set_last_error("device buffer overflow in ORB pipeline (candidates / quadtree nodes)");
    return e;
  }
  return SDPL_OK;
}

extern "C" {

int sdpl_orb_create(sdpl_orb** out, int nfeatures, float scale, int nlevels, int ini_th, int min_th, int device) {
  if (!out || nfeatures < 1 || nlevels < 1 || nlevels > kMaxLevels || !(scale > 1.0f) || ini_th < 1 || min_th < 1 ||
      min_th > ini_th || ini_th > 254) {
    set_last_error("sdpl_orb_create: bad argument");
    return SDPL_ERR_ARG;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
    set_last_error("sdpl_orb_create: no such CUDA device (this library has no CPU fallback)");
    return SDPL_ERR_CUDA;
  }
  SDPL_CUDA(cudaSetDevice(device));
  sdpl_orb* o = new sdpl_orb;
  o->nfeatures = nfeatures; o->nlevels = nlevels; o->ini_th = ini_th; o->min_th = min_th; o->device = device; o->scale = scale;
  // scale tables, ORBextractor.cc:404-421 (scaleFactor member is a double, include/ORBextractor.h:85)
  const double sfd = (double)scale;
  o->sf.resize(nlevels); o->isf.resize(nlevels); o->s2.resize(nlevels); o->is2.resize(nlevels); o->quota.resize(nlevels);
  o->sf[0] = 1.f; o->s2[0] = 1.f;
  for (int i = 1; i < nlevels; i++) { o->sf[i] = (float)(o->sf[i - 1] * sfd); o->s2[i] = o->sf[i] * o->sf[i]; }
  for (int i = 0; i < nlevels; i++) { o->isf[i] = 1.0f / o->sf[i]; o->is2[i] = 1.0f / o->s2[i]; }
  // per-level quota, ORBextractor.cc:425-435
  float factor = (float)(1.0f / sfd);
  float want = nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)nlevels));
  int sum = 0;
  for (int l = 0; l < nlevels - 1; l++) { o->quota[l] = h_cv_round(want); sum += o->quota[l]; want *= factor; }
  o->quota[nlevels - 1] = std::max(nfeatures - sum, 0);
  // umax, ORBextractor.cc:443-458
  {
    const int HP = 15;
    int vmax = h_cv_floor(HP * sqrtf(2.f) / 2 + 1), vmin = (int)ceilf(HP * sqrtf(2.f) / 2);
    for (int v = 0; v <= vmax; ++v) o->umax[v] = h_cv_round(sqrt((double)HP * HP - v * v));
    for (int v = HP, v0 = 0; v >= vmin; --v) {
      while (o->umax[v0] == o->umax[v0 + 1]) ++v0;
      o->umax[v] = v0;
      ++v0;
    }
  }
  SDPL_CUDA(cudaStreamCreateWithFlags(&o->own_stream, cudaStreamNonBlocking));
  o->stream = o->own_stream;
  *out = o;
  return SDPL_OK;
}

void sdpl_orb_destroy(sdpl_orb* o) {
  if (!o) return;
  cudaSetDevice(o->device);
  cudaStreamSynchronize(o->stream);
  for (DevBuf* b : {&o->pyr, &o->score, &o->blur, &o->cellinfo, &o->cellmask, &o->cand_xy, &o->cand_resp, &o->cand_node, &o->n_cand, &o->n_kp,
                    &o->kp_xy, &o->kp_resp, &o->cellsdev, &o->tables, &o->err, &o->in_stage, &o->out_kps, &o->out_desc, &o->out_n})
    b->release();
  if (o->h_stage) cudaFreeHost(o->h_stage);
  o->timer.release();
  if (o->own_stream) cudaStreamDestroy(o->own_stream);
  delete o;
}

int sdpl_orb_set_stream(sdpl_orb* o, void* s) { if (!o) return SDPL_ERR_ARG; o->stream = s ? (cudaStream_t)s : o->own_stream; return SDPL_OK; }
int sdpl_orb_levels(const sdpl_orb* o) { return o ? o->nlevels : 0; }
int sdpl_orb_tables(const sdpl_orb* o, float* a, float* b, float* c, float* d) {
  if (!o) return SDPL_ERR_ARG;
  for (int i = 0; i < o->nlevels; i++) {
    if (a) a[i] = o->sf[i];
    if (b) b[i] = o->isf[i];
    if (c) c[i] = o->s2[i];
    if (d) d[i] = o->is2[i];
  }
  return SDPL_OK;
}
int sdpl_orb_quota(const sdpl_orb* o, int* q, int* umax) {
  if (!o) return SDPL_ERR_ARG;
  if (q) for (int i = 0; i < o->nlevels; i++) q[i] = o->quota[i];
  if (umax) memcpy(umax, o->umax, sizeof(o->umax));
  return SDPL_OK;
}
int sdpl_orb_max_keypoints(const sdpl_orb* o) {
  if (!o) return 0;
  int t = 64;
  for (int q : o->quota) t += q + 3;
  return t;
}
int sdpl_orb_last_launches(const sdpl_orb* o) { return o ? o->launches : 0; }
int sdpl_orb_peek_error_async(sdpl_orb* o, void* stream, int* host_flag) {
  if (!o || !host_flag) return SDPL_ERR_ARG;
  if (!o->err.p) { *host_flag = 0; return SDPL_OK; }
  SDPL_CUDA(cudaMemcpyAsync(host_flag, o->err.p, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  return SDPL_OK;
}
// the overflow flag of everything enqueued so far on the handle's stream, copied to *dst (pinned host or device memory) and
// cleared, both in stream order: the flag then belongs to the batch that raised it
int sdpl_orb_take_error_async(sdpl_orb* o, int* dst) {
  if (!o || !dst) return SDPL_ERR_ARG;
  if (!o->err.p) { *dst = 0; return SDPL_OK; }
  SDPL_CUDA(cudaMemcpyAsync(dst, o->err.p, sizeof(int), cudaMemcpyDefault, o->stream));
  SDPL_CUDA(cudaMemsetAsync(o->err.p, 0, sizeof(int), o->stream));
  return SDPL_OK;
}
int sdpl_orb_check(sdpl_orb* o) {
  if (!o) return SDPL_ERR_ARG;
  if (!o->err.p) return SDPL_OK;
  SDPL_CUDA(cudaSetDevice(o->device));
  return orb_check_err(o);
}
int sdpl_orb_set_profiling(sdpl_orb* o, int on) { if (!o) return SDPL_ERR_ARG; o->timer.enabled = on != 0; return SDPL_OK; }
int sdpl_orb_stage_times(sdpl_orb* o, float* ms, const char** names, int* launches, int cap) {
  if (!o) return 0;
  cudaSetDevice(o->device);
  return o->timer.read(ms, names, launches, cap);
}

int sdpl_orb_extract_batch_dev(sdpl_orb* o, const uint8_t* d_imgs, int nframes, int w, int h, int stride, size_t frame_stride,
                               sdpl_keypoint* d_kps, uint8_t* d_desc, int capacity, int* d_n_out, int sync) {
  if (!o || !d_imgs || nframes < 1 || w < 1 || h < 1 || stride < w || !d_kps || !d_desc || capacity < 1 || !d_n_out) {
    set_last_error("sdpl_orb_extract_batch_dev: bad argument");
    return SDPL_ERR_ARG;
  }
  SDPL_CUDA(cudaSetDevice(o->device));
  g_launches = 0;
  int rc = orb_run_dev(o, d_imgs, nframes, w, h, stride, frame_stride, d_kps, d_desc, capacity, d_n_out);
  o->launches = g_launches;
  if (rc) return rc;
  if (sync) return orb_check_err(o);
  return SDPL_OK;
}

int sdpl_orb_extract_batch(sdpl_orb* o, const uint8_t* imgs, int nframes, int w, int h, int stride, size_t frame_stride,
                           sdpl_keypoint* kps, uint8_t* desc, int capacity, int* n_out) {
  if (!o || !n_out || nframes < 0) { set_last_error("sdpl_orb_extract_batch: bad argument"); return SDPL_ERR_ARG; }
  if (!imgs || w <= 0 || h <= 0 || nframes == 0) {   // empty image: silent return, ORBextractor.cc:1038
    for (int f = 0; f < nframes; f++) n_out[f] = 0;
    return SDPL_OK;
  }
  if (stride < w || !kps || !desc || capacity < 1) { set_last_error("sdpl_orb_extract_batch: bad argument"); return SDPL_ERR_ARG; }
  SDPL_CUDA(cudaSetDevice(o->device));
  int rc;
  if ((rc = orb_setup(o, w, h, nframes))) return rc;
  const int cap_dev = o->kp_cap_total;
  const size_t in_bytes = (size_t)w * h * nframes;
  if ((rc = o->in_stage.reserve(in_bytes))) return rc;
  if ((rc = o->out_kps.reserve(sizeof(sdpl_keypoint) * (size_t)cap_dev * nframes))) return rc;
  if ((rc = o->out_desc.reserve((size_t)32 * cap_dev * nframes))) return rc;
  if ((rc = o->out_n.reserve(sizeof(int) * nframes))) return rc;
  const size_t out_bytes = (sizeof(sdpl_keypoint) + 32) * (size_t)cap_dev * nframes + sizeof(int) * nframes;
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
  rc = orb_run_dev(o, o->in_stage.as<uint8_t>(), nframes, w, h, w, (size_t)w * h, o->out_kps.as<sdpl_keypoint>(),
                   o->out_desc.as<uint8_t>(), cap_dev, o->out_n.as<int>());
  o->launches = g_launches;
  if (rc) return rc;
  char* hs = (char*)o->h_stage;
  sdpl_keypoint* h_k = (sdpl_keypoint*)hs;
  uint8_t* h_d = (uint8_t*)(hs + sizeof(sdpl_keypoint) * (size_t)cap_dev * nframes);
  int* h_n = (int*)(hs + (sizeof(sdpl_keypoint) + 32) * (size_t)cap_dev * nframes);
  SDPL_CUDA(cudaMemcpyAsync(h_k, o->out_kps.p, sizeof(sdpl_keypoint) * (size_t)cap_dev * nframes, cudaMemcpyDeviceToHost, st));
  SDPL_CUDA(cudaMemcpyAsync(h_d, o->out_desc.p, (size_t)32 * cap_dev * nframes, cudaMemcpyDeviceToHost, st));
  SDPL_CUDA(cudaMemcpyAsync(h_n, o->out_n.p, sizeof(int) * nframes, cudaMemcpyDeviceToHost, st));
  if ((rc = orb_check_err(o))) return rc;
  int status = SDPL_OK;
  for (int f = 0; f < nframes; f++) {
    int n = h_n[f];
    n_out[f] = n;
    if (n > capacity) { status = SDPL_ERR_CAPACITY; n = capacity; }
    memcpy(kps + (size_t)f * capacity, h_k + (size_t)f * cap_dev, sizeof(sdpl_keypoint) * n);
    memcpy(desc + (size_t)f * capacity * 32, h_d + (size_t)f * cap_dev * 32, (size_t)32 * n);
  }
  if (status == SDPL_ERR_CAPACITY) set_last_error("sdpl_orb_extract: output capacity too small");
  return status;
}

int sdpl_orb_extract(sdpl_orb* o, const uint8_t* img, int w, int h, int stride, sdpl_keypoint* kps, uint8_t* desc, int capacity,
                     int* n_out) {
  return sdpl_orb_extract_batch(o, img, 1, w, h, stride, (size_t)stride * (h > 0 ? h : 0), kps, desc, capacity, n_out);
}

int sdpl_orb_pyramid_level(sdpl_orb* o, int frame, int level, uint8_t* out, int out_stride, int* w_out, int* h_out) {
  if (!o || o->gw == 0 || level < 0 || level >= o->nlevels || frame < 0 || frame >= o->last_B) return SDPL_ERR_ARG;
  const LvlDev& L = o->D.L[level];
  if (w_out) *w_out = L.w;
  if (h_out) *h_out = L.h;
  if (!out) return SDPL_OK;
  if (out_stride < L.w + 2 * kBorder) return SDPL_ERR_ARG;
  SDPL_CUDA(cudaSetDevice(o->device));
  SDPL_CUDA(cudaStreamSynchronize(o->stream));
  SDPL_CUDA(cudaMemcpy2D(out, out_stride, o->D.pyr + (size_t)frame * o->D.pyr_frame + L.pyr_off, L.pstride, L.w + 2 * kBorder,
                         L.h + 2 * kBorder, cudaMemcpyDeviceToHost));
  return SDPL_OK;
}

int sdpl_orb_blurred_level(sdpl_orb* o, int frame, int level, uint8_t* out, int out_stride) {
  if (!o || o->gw == 0 || level < 0 || level >= o->nlevels || frame < 0 || frame >= o->last_B || !out) return SDPL_ERR_ARG;
  const LvlDev& L = o->D.L[level];
  if (out_stride < L.w) return SDPL_ERR_ARG;
  SDPL_CUDA(cudaSetDevice(o->device));
  SDPL_CUDA(cudaStreamSynchronize(o->stream));
  SDPL_CUDA(cudaMemcpy2D(out, out_stride, o->D.blur + (size_t)frame * o->D.pyr_frame + L.pyr_off + (size_t)kBorder * L.pstride + kBorder, L.pstride, L.w, L.h,
                         cudaMemcpyDeviceToHost));
  return SDPL_OK;
}

int sdpl_orb_candidates(sdpl_orb* o, int frame, int level, int* xs, int* ys, int* resp, int capacity, int* n_out) {
  if (!o || o->gw == 0 || level < 0 || level >= o->nlevels || frame < 0 || frame >= o->last_B || !n_out) return SDPL_ERR_ARG;
  const LvlDev& L = o->D.L[level];
  SDPL_CUDA(cudaSetDevice(o->device));
  SDPL_CUDA(cudaStreamSynchronize(o->stream));
  int n = 0;
  SDPL_CUDA(cudaMemcpy(&n, o->D.n_cand + frame * o->nlevels + level, sizeof(int), cudaMemcpyDeviceToHost));
  *n_out = n;
  int m = std::min(n, capacity);
  if (m <= 0 || !xs) return SDPL_OK;
  std::vector<uint32_t> xy(m);
  std::vector<uint8_t> r(m);
  SDPL_CUDA(cudaMemcpy(xy.data(), o->D.cand_xy + (size_t)frame * o->D.cand_per_frame + L.cand_off, sizeof(uint32_t) * m, cudaMemcpyDeviceToHost));
  SDPL_CUDA(cudaMemcpy(r.data(), o->D.cand_resp + (size_t)frame * o->D.cand_per_frame + L.cand_off, m, cudaMemcpyDeviceToHost));
  for (int i = 0; i < m; i++) { xs[i] = xy[i] & 0xFFFF; ys[i] = xy[i] >> 16; resp[i] = r[i]; }
  return SDPL_OK;
}

int sdpl_orb_level_counts(sdpl_orb* o, int frame, int* per_level) {
  if (!o || o->gw == 0 || frame < 0 || frame >= o->last_B || !per_level) return SDPL_ERR_ARG;
  SDPL_CUDA(cudaSetDevice(o->device));
  SDPL_CUDA(cudaStreamSynchronize(o->stream));
  SDPL_CUDA(cudaMemcpy(per_level, o->D.n_kp + frame * o->nlevels, sizeof(int) * o->nlevels, cudaMemcpyDeviceToHost));
  return SDPL_OK;
}

}  // extern "C"
