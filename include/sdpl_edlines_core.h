/*
 * include/sdpl_edlines_core.h -- the sequential half of the EDLines back-end (Lineextractor extractor == 1), written once for the host
 * (oracle/ed_oracle.cpp) and for the device (sdpl_slam_b200/csrc/edlines.cu, one thread per (frame, octave) task): anchor linking
 * (ED::JoinAnchorPointsUsingSortedAnchors, 3rdparty/line_descriptor/src/ED_Lib/ED.cpp:400-963), line fitting
 * (EDLines::SplitSegment2Lines, EDLines.cpp:270-358), joining (JoinCollinearLines / TryToJoinTwoLineSegments, :362-400, :775-925) and
 * validation (ValidateLineSegments / ValidateLineSegmentRect / EnumerateRectPoints, :405-566, :988-1155; NFALUT, NFA.cpp).
 * Restated, not copied: one table-driven walk instead of four unrolled ones, index arithmetic instead of pointers, explicit
 * capacities.  Both sides are pinned to the reference's own ED_Lib sources compiled unmodified (oracle/refshim,
 * tests/test_oracle_vs_ref.py).  Only + - * / sqrt fabs ceil on doubles and integer arithmetic: the same bits on x86-64 without
 * FMA contraction and on sm_100a with --fmad=false; the two transcendental tables (atan, NFA thresholds) come from the host.
 *
 * Three properties of the reference that this code reproduces on purpose: the validation reads the SOURCE
 * image through `srcImg[r * width + c]` although the Mat it was given is a ROI of LSDDetectorC's 19-pixel padded buffer (row stride
 * width + 38, EDLines.cpp:462-466, 551-555); the per-pixel validation of medium lines starts at the segment's pixel 0 instead of the
 * line's first pixel and swaps x / y (:432, :449-450); and atan(1.0 / b) is evaluated for b == 0 (:426, :520).
 */
#ifndef SDPL_EDLINES_CORE_H
#define SDPL_EDLINES_CORE_H
#include <math.h>
#include <stdint.h>
#ifdef __CUDACC__
#define SDPL_ED_HD __host__ __device__
#else
#define SDPL_ED_HD
#endif

namespace sdpl_ed {

enum { kVertical = 1, kHorizontal = 2, kAnchor = 254, kEdge = 255 };
enum { kLeft = 1, kRight = 2, kUp = 3, kDown = 4 };
enum { kGradThresh = 36, kAnchorThresh = 8, kMinPathLen = 10 };       // EDLines.cpp:9: ED(srcImage, SOBEL_OPERATOR, 36, 8), ED.h defaults
enum { kErrStack = 1, kErrChains = 2, kErrPixels = 4, kErrSegPixels = 8, kErrSegments = 16, kErrLines = 32, kErrNfaTable = 64, kErrShortSegment = 128 };
constexpr int kAtanLut = 1024;                                          // NFA.h: MAX_LUT_SIZE
constexpr double kPi = 3.14159265358979323846;

struct Chain { int dir, len, parent, child0, child1, pix; };           // pix = offset of the chain's first pixel in Work::pixels
struct Node { int r, c, parent, dir; };                                 // walk stack entry / scratch of the tree walks
struct Line {                                                           // EDLines.h: LineSegment
  double a, b; int invert;
  double sx, sy, ex, ey;
  int segmentNo, firstPixelIndex, len;
};

struct Work {                      // one (frame, octave) task
  int w, h;
  const int16_t* grad;             // |gx| + |gy| of the smoothed image, border = kGradThresh - 1
  const uint8_t* dir;              // kVertical / kHorizontal where grad >= kGradThresh
  uint8_t* edge;                   // 0 / kAnchor on entry; kEdge on the drawn pixels on exit
  const int* anchors; int n_anchors;        // pixel offsets in processing order (descending gradient, row-major inside a value)
  int* pixels; int pixels_cap;     // pixels of the current walk, packed (y << 16) | x
  Node* stack; int stack_cap;
  Chain* chains; int chains_cap;
  int* chain_nos; int chain_nos_cap;
  int* seg_px; int seg_px_cap;     // all segments' pixels, packed, segment s = seg_px[seg_off[s], seg_off[s + 1])
  int* seg_off; int seg_cap; int nseg;
  Line* lines; int lines_cap; int nlines;
  // validation
  const uint8_t* src; int src_stride;       // the level image the reference's ED object was given (ROI origin, real row stride)
  unsigned long long src_magic;             // ceil(2^35 / (w + 38)) when w + 38 < 2^12, else 0 (src_flat)
  const double* atan_lut;          // atan(i / 1024), i = 0..1024
  const int* nfa_min_k; int nfa_n; // smallest k with NFALUT::checkValidationByNFA(n, k) for n < nfa_n (n + 1 = never)
  int min_line_len;
  int err;
};

SDPL_ED_HD inline unsigned long long src_magic_of(int w) {
  const unsigned long long S = (unsigned long long)(w + 38);
  return S < (1ull << 12) ? ((1ull << 35) + S - 1) / S : 0ull;
}
SDPL_ED_HD inline int px_pack(int x, int y) { return (y << 16) | x; }
SDPL_ED_HD inline int px_x(int p) { return p & 0xffff; }
SDPL_ED_HD inline int px_y(int p) { return (int)((unsigned)p >> 16); }
SDPL_ED_HD inline int iabs(int v) { return v < 0 ? -v : v; }

// ------------------------------------------------------------------------------------------------
// ED::LongestChain (ED.cpp:1049-1071): length of the longest path below `root`; prunes the shorter child of every chain it passes.
// The reference recurses; a device thread has no stack for that, so the post-order walk uses the (idle) walk stack.  Node::c holds the
// phase (0 = enter, 1 = children done), Node::parent / Node::dir the results of the two children.
// ------------------------------------------------------------------------------------------------
SDPL_ED_HD inline int longest_chain(Work& W, int root) {
  Chain* ch = W.chains;
  if (root == -1 || ch[root].len == 0) return 0;
  Node* st = W.stack;                 // r = chain, c = phase (0 enter, 1 first child done, 2 second child done), parent = len0
  int top = 0, ret = 0;
  st[0].r = root; st[0].c = 0; st[0].parent = 0;
  while (top >= 0) {
    const int id = st[top].r;
    if (st[top].c == 0) {
      st[top].c = 1;
      const int c0 = ch[id].child0;
      ret = 0;
      if (c0 != -1 && ch[c0].len != 0) {
        if (top + 1 >= W.stack_cap) { W.err |= kErrStack; return 0; }
        ++top; st[top].r = c0; st[top].c = 0; st[top].parent = 0;
        continue;
      }
    }
    if (st[top].c == 1) {
      st[top].parent = ret;             // length below the first child
      st[top].c = 2;
      const int c1 = ch[id].child1;
      ret = 0;
      if (c1 != -1 && ch[c1].len != 0) {
        if (top + 1 >= W.stack_cap) { W.err |= kErrStack; return 0; }
        ++top; st[top].r = c1; st[top].c = 0; st[top].parent = 0;
        continue;
      }
    }
    const int len0 = st[top].parent, len1 = ret;
    int mx;
    if (len0 >= len1) { mx = len0; ch[id].child1 = -1; } else { mx = len1; ch[id].child0 = -1; }
    ret = ch[id].len + mx;
    --top;
  }
  return ret;
}

SDPL_ED_HD inline int retrieve_chain_nos(Work& W, int root) {           // ED.cpp:1073-1086
  int count = 0;
  while (root != -1) {
    if (count >= W.chain_nos_cap) { W.err |= kErrChains; break; }
    W.chain_nos[count++] = root;
    if (W.chains[root].child0 != -1) root = W.chains[root].child0; else root = W.chains[root].child1;
  }
  return count;
}

// ------------------------------------------------------------------------------------------------
// ED::JoinAnchorPointsUsingSortedAnchors (ED.cpp:400-963)
// ------------------------------------------------------------------------------------------------
SDPL_ED_HD inline void seg_push(Work& W, int& n_total, int p) {
  if (n_total >= W.seg_px_cap) { W.err |= kErrSegPixels; return; }
  W.seg_px[n_total++] = p;
}

SDPL_ED_HD inline void link_anchors(Work& W) {
  const int w = W.w;
  const int16_t* grad = W.grad; const uint8_t* dirm = W.dir; uint8_t* edge = W.edge;
  Chain* ch = W.chains; int* px = W.pixels; Node* st = W.stack;
  int seg_total = 0;                // pixels stored so far = start of the open segment
  W.nseg = 0;
  if (W.seg_cap > 0) W.seg_off[0] = 0;
  for (int k = 0; k < W.n_anchors; k++) {
    const int off = W.anchors[k];
#ifdef __CUDA_ARCH__
    // most anchors have been erased or drawn over by the time their turn comes: the byte that says so is a scattered read, asked for a
    // few anchors ahead (a cache hint only; the value is read when the anchor's turn comes)
    if (k + 8 < W.n_anchors) asm volatile("prefetch.global.L2 [%0];" :: "l"(edge + W.anchors[k + 8]));
#endif
    if (edge[off] != kAnchor) continue;
    const int ai = off / w, aj = off % w;
    ch[0].len = 0; ch[0].parent = -1; ch[0].dir = 0; ch[0].child0 = ch[0].child1 = -1; ch[0].pix = 0;
    int no_chains = 1, len = 0, dup = 0, top = -1;
    if (dirm[off] == kVertical) {
      ++top; st[top].r = ai; st[top].c = aj; st[top].dir = kDown; st[top].parent = 0;
      ++top; st[top].r = ai; st[top].c = aj; st[top].dir = kUp; st[top].parent = 0;
    } else {
      ++top; st[top].r = ai; st[top].c = aj; st[top].dir = kRight; st[top].parent = 0;
      ++top; st[top].r = ai; st[top].c = aj; st[top].dir = kLeft; st[top].parent = 0;
    }
    while (top >= 0) {
      int r = st[top].r, c = st[top].c;
      const int dir = st[top].dir, parent = st[top].parent;
      --top;
      if (edge[r * w + c] != kEdge) dup++;
      if (no_chains >= W.chains_cap || len + 2 >= W.pixels_cap || top + 3 >= W.stack_cap) { W.err |= kErrChains; top = -1; break; }
      ch[no_chains].dir = dir; ch[no_chains].parent = parent; ch[no_chains].child0 = ch[no_chains].child1 = -1;
      int chain_len = 0;
      ch[no_chains].pix = len;
      px[len++] = px_pack(c, r); chain_len++;
      // geometry of the walk: forward step (fr, fc), lateral axis (lr, lc), which lateral neighbour is looked at first
      const bool horiz = dir == kLeft || dir == kRight;
      const int fr = dir == kUp ? -1 : (dir == kDown ? 1 : 0), fc = dir == kLeft ? -1 : (dir == kRight ? 1 : 0);
      const int lr = horiz ? 1 : 0, lc = horiz ? 0 : 1;
      const int first_lat = (dir == kLeft || dir == kUp) ? -1 : 1;
      const int want = horiz ? kHorizontal : kVertical;
      const bool to_child0 = dir == kLeft || dir == kUp;
      bool stopped = false;
      // the walk in linear offsets: p = r * w + c, one step ahead = p + df, one pixel to either side = -+ dl (adds only: the walk is
      // bound by the instruction count of its one thread on the device)
      const int df = fr * w + fc, dl = lr * w + lc;
      int p = r * w + c;
      int cur_dir = dirm[p];
      while (cur_dir == want) {
        // every value the step can need is loaded up front, independently of the decisions below (one memory round trip per step)
        const int oB = p + df, oA = oB - dl, oC = oB + dl, o1 = p - dl, o2 = p + dl;
        const int eA = edge[oA], eB = edge[oB], eC = edge[oC], e1 = edge[o1], e2 = edge[o2];
        const int gA = grad[oA], gB = grad[oB], gC = grad[oC];
        const int dA = dirm[oA], dB = dirm[oB], dC = dirm[oC];
        edge[p] = kEdge;
        // clean up the anchors beside the path
        if (e1 == kAnchor) edge[o1] = 0;
        if (e2 == kAnchor) edge[o2] = 0;
        // the neighbour ahead that is already an anchor / edge pixel, looked for in the reference's order; else the largest gradient
        const int eF = first_lat < 0 ? eA : eC, eS = first_lat < 0 ? eC : eA;
        int lat;
        if (eB >= kAnchor) lat = 0;
        else if (eF >= kAnchor) lat = first_lat;
        else if (eS >= kAnchor) lat = -first_lat;
        else { lat = 0; if (gA > gB) { lat = gA > gC ? -1 : 1; } else if (gC > gB) lat = 1; }
        r += fr + lat * lr; c += fc + lat * lc;
        p = lat == 0 ? oB : (lat < 0 ? oA : oC);
        const int eN = lat == 0 ? eB : (lat < 0 ? eA : eC), gN = lat == 0 ? gB : (lat < 0 ? gA : gC);
        cur_dir = lat == 0 ? dB : (lat < 0 ? dA : dC);
        if (eN == kEdge || gN < kGradThresh) {
          ch[no_chains].len = chain_len;
          if (to_child0) ch[parent].child0 = no_chains; else ch[parent].child1 = no_chains;
          no_chains++;
          stopped = true;
          break;
        }
        if (len + 2 >= W.pixels_cap) { W.err |= kErrPixels; stopped = true; break; }
        px[len++] = px_pack(c, r); chain_len++;
      }
      if (stopped) continue;
      // the gradient direction changed: continue in both perpendicular directions from here
      ++top; st[top].r = r; st[top].c = c; st[top].dir = horiz ? kDown : kRight; st[top].parent = no_chains;
      ++top; st[top].r = r; st[top].c = c; st[top].dir = horiz ? kUp : kLeft; st[top].parent = no_chains;
      len--; chain_len--;
      ch[no_chains].len = chain_len;
      if (to_child0) ch[parent].child0 = no_chains; else ch[parent].child1 = no_chains;
      no_chains++;
    }
    if (W.err) return;
    if (len - dup < kMinPathLen) {
      for (int q = 0; q < len; q++) edge[px_y(px[q]) * w + px_x(px[q])] = 0;
      continue;
    }
    // ---- the walk becomes one segment (longest path through the anchor) plus one per remaining long chain ----
    if (W.nseg + 1 >= W.seg_cap) { W.err |= kErrSegments; return; }
    int* seg = W.seg_px + seg_total;          // open segment
    int ns = 0;
    int total = longest_chain(W, ch[0].child1);
    if (total > 0) {
      const int count = retrieve_chain_nos(W, ch[0].child1);
      for (int q = count - 1; q >= 0; q--) {
        const int cn = W.chain_nos[q];
        const int* cp = px + ch[cn].pix;
        if (ch[cn].pix + ch[cn].len - 1 < 0) { W.err |= kErrShortSegment; return; }
        int fr_ = px_y(cp[ch[cn].len - 1]), fc_ = px_x(cp[ch[cn].len - 1]);
        int index = ns - 2;
        while (index >= 0) {
          if (iabs(fr_ - px_y(seg[index])) <= 1 && iabs(fc_ - px_x(seg[index])) <= 1) { ns--; index--; } else break;
        }
        if (ch[cn].len > 1 && ns > 0) {
          fr_ = px_y(cp[ch[cn].len - 2]); fc_ = px_x(cp[ch[cn].len - 2]);
          if (iabs(fr_ - px_y(seg[ns - 1])) <= 1 && iabs(fc_ - px_x(seg[ns - 1])) <= 1) ch[cn].len--;
        }
        for (int l = ch[cn].len - 1; l >= 0; l--) { if (seg_total + ns >= W.seg_px_cap) { W.err |= kErrSegPixels; return; } seg[ns++] = cp[l]; }
        ch[cn].len = 0;
      }
    }
    total = longest_chain(W, ch[0].child0);
    if (total > 1) {
      const int count = retrieve_chain_nos(W, ch[0].child0);
      const int last = W.chain_nos[0];
      ch[last].pix++; ch[last].len--;
      for (int q = 0; q < count; q++) {
        const int cn = W.chain_nos[q];
        const int* cp = px + ch[cn].pix;
        int fr_ = px_y(cp[0]), fc_ = px_x(cp[0]);
        int index = ns - 2;
        while (index >= 0) {
          if (iabs(fr_ - px_y(seg[index])) <= 1 && iabs(fc_ - px_x(seg[index])) <= 1) { ns--; index--; } else break;
        }
        int start = 0;
        if (ch[cn].len > 1 && ns > 0) {
          fr_ = px_y(cp[1]); fc_ = px_x(cp[1]);
          if (iabs(fr_ - px_y(seg[ns - 1])) <= 1 && iabs(fc_ - px_x(seg[ns - 1])) <= 1) start = 1;
        }
        for (int l = start; l < ch[cn].len; l++) { if (seg_total + ns >= W.seg_px_cap) { W.err |= kErrSegPixels; return; } seg[ns++] = cp[l]; }
        ch[cn].len = 0;
      }
    }
    // first pixel clean-up (the reference indexes element 1 unconditionally: a segment shorter than two pixels is undefined there)
    if (ns < 2) { W.err |= kErrShortSegment; return; }
    if (iabs(px_y(seg[1]) - px_y(seg[ns - 1])) <= 1 && iabs(px_x(seg[1]) - px_x(seg[ns - 1])) <= 1) {
      for (int q = 1; q < ns; q++) seg[q - 1] = seg[q];
      ns--;
    }
    seg_total += ns;
    W.nseg++;
    W.seg_off[W.nseg] = seg_total;
    // the remaining long chains
    for (int c2 = 2; c2 < no_chains; c2++) {
      if (ch[c2].len < 2) continue;
      total = longest_chain(W, c2);
      if (total >= 10) {
        if (W.nseg + 1 >= W.seg_cap) { W.err |= kErrSegments; return; }
        const int count = retrieve_chain_nos(W, c2);
        seg = W.seg_px + seg_total; ns = 0;
        for (int q = 0; q < count; q++) {
          const int cn = W.chain_nos[q];
          const int* cp = px + ch[cn].pix;
          int fr_ = px_y(cp[0]), fc_ = px_x(cp[0]);
          int index = ns - 2;
          while (index >= 0) {
            if (iabs(fr_ - px_y(seg[index])) <= 1 && iabs(fc_ - px_x(seg[index])) <= 1) { ns--; index--; } else break;
          }
          int start = 0;
          if (ch[cn].len > 1 && ns > 0) {
            fr_ = px_y(cp[1]); fc_ = px_x(cp[1]);
            if (iabs(fr_ - px_y(seg[ns - 1])) <= 1 && iabs(fc_ - px_x(seg[ns - 1])) <= 1) start = 1;
          }
          for (int l = start; l < ch[cn].len; l++) { if (seg_total + ns >= W.seg_px_cap) { W.err |= kErrSegPixels; return; } seg[ns++] = cp[l]; }
          ch[cn].len = 0;
        }
        seg_total += ns;
        W.nseg++;
        W.seg_off[W.nseg] = seg_total;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// EDLines geometry helpers (EDLines.cpp:573-771, 931-984)
// ------------------------------------------------------------------------------------------------
SDPL_ED_HD inline void closest_point(double x1, double y1, double a, double b, int invert, double& xo, double& yo) {
  double x2, y2;
  if (invert == 0) {
    if (b == 0) { x2 = x1; y2 = a; }
    else { const double d = -1.0 / b; const double c = y1 - d * x1; x2 = (a - c) / (d - b); y2 = a + b * x2; }
  } else {
    if (b == 0) { x2 = a; y2 = y1; }
    else { const double d = -1.0 / b; const double c = x1 - d * y1; y2 = (a - c) / (d - b); x2 = a + b * y2; }
  }
  xo = x2; yo = y2;
}
SDPL_ED_HD inline double min_distance(double x1, double y1, double a, double b, int invert) {
  double x2, y2;
  closest_point(x1, y1, a, b, invert, x2, y2);
  return sqrt((x1 - x2) * (x1 - x2) + (y1 - y2) * (y1 - y2));
}
// pixel coordinates of a segment as the doubles the reference copies into its x / y buffers
struct SegXY {
  const int* p;
  SDPL_ED_HD double x(int i) const { return (double)px_x(p[i]); }
  SDPL_ED_HD double y(int i) const { return (double)px_y(p[i]); }
};
// LineFit with a known orientation (EDLines.cpp:662-692)
SDPL_ED_HD inline void line_fit_known(const SegXY& s, int o, int count, double& a, double& b, int invert) {
  if (count < 2) return;
  const double S = count;
  double Sx = 0.0, Sy = 0.0, Sxx = 0.0, Sxy = 0.0;
  for (int i = 0; i < count; i++) { Sx += s.x(o + i); Sy += s.y(o + i); }
  if (invert) { const double d = Sx; Sx = Sy; Sy = d; }
  for (int i = 0; i < count; i++) {
    const double xi = invert ? s.y(o + i) : s.x(o + i), yi = invert ? s.x(o + i) : s.y(o + i);
    Sxx += xi * xi; Sxy += xi * yi;
  }
  const double D = S * Sxx - Sx * Sx;
  a = (Sxx * Sy - Sx * Sxy) / D;
  b = (S * Sxy - Sx * Sy) / D;
}
// LineFit that also chooses the orientation and returns the fit error (EDLines.cpp:698-771)
SDPL_ED_HD inline void line_fit(const SegXY& s, int o, int count, double& a, double& b, double& e, int& invert) {
  if (count < 2) return;
  const double S = count;
  double Sx = 0.0, Sy = 0.0, Sxx = 0.0, Sxy = 0.0;
  for (int i = 0; i < count; i++) { Sx += s.x(o + i); Sy += s.y(o + i); }
  const double mx = Sx / count, my = Sy / count;
  double dx = 0.0, dy = 0.0;
  for (int i = 0; i < count; i++) { dx += (s.x(o + i) - mx) * (s.x(o + i) - mx); dy += (s.y(o + i) - my) * (s.y(o + i) - my); }
  if (dx < dy) { invert = 1; const double d = Sx; Sx = Sy; Sy = d; } else invert = 0;
  for (int i = 0; i < count; i++) {
    const double xi = invert ? s.y(o + i) : s.x(o + i), yi = invert ? s.x(o + i) : s.y(o + i);
    Sxx += xi * xi; Sxy += xi * yi;
  }
  const double D = S * Sxx - Sx * Sx;
  a = (Sxx * Sy - Sx * Sxy) / D;
  b = (S * Sxy - Sx * Sy) / D;
  if (b == 0.0) {
    double error = 0.0;
    for (int i = 0; i < count; i++) { const double yi = invert ? s.x(o + i) : s.y(o + i); error += fabs(a - yi); }
    e = error / count;
  } else {
    double error = 0.0;
    for (int i = 0; i < count; i++) {
      const double xi = invert ? s.y(o + i) : s.x(o + i), yi = invert ? s.x(o + i) : s.y(o + i);
      const double d = -1.0 / b;
      const double c = yi - d * xi;
      const double x2 = (a - c) / (d - b);
      const double y2 = a + b * x2;
      error += (xi - x2) * (xi - x2) + (yi - y2) * (yi - y2);
    }
    e = sqrt(error / count);
  }
}

// EDLines::SplitSegment2Lines (EDLines.cpp:270-358); line_error = 1.0.  The lines of segment seg_no go to out[0, cap) (out == null: they
// are only counted); returns their number.  Segments are independent of each other, so a caller may work on many at once.
SDPL_ED_HD inline int split_segment(Work& W, int seg_no, Line* out, int cap) {
  int n_out = 0;
  SegXY s; s.p = W.seg_px + W.seg_off[seg_no];
  int no_pixels = W.seg_off[seg_no + 1] - W.seg_off[seg_no];
  const int mll = W.min_line_len;
  const double line_error = 1.0;
  int o = 0;                        // x / y pointer advance
  int first_pixel_index = 0;
  while (no_pixels >= mll) {
    bool valid = false;
    double lastA = 0, lastB = 0, error = 0;
    int lastInvert = 0;
    while (no_pixels >= mll) {
      line_fit(s, o, mll, lastA, lastB, error, lastInvert);
      if (error <= 0.5) { valid = true; break; }
      no_pixels -= 1; o += 1; first_pixel_index += 1;
    }
    if (!valid) return n_out;
    int index = mll, len = mll;
    while (index < no_pixels) {
      const int start_index = index;
      int last_good = index - 1, good = 0, bad = 0;
      while (index < no_pixels) {
        const double d = min_distance(s.x(o + index), s.y(o + index), lastA, lastB, lastInvert);
        if (d <= line_error) { last_good = index; good++; bad = 0; }
        else { bad++; if (bad >= 5) break; }
        index++;
      }
      if (good >= 2) {
        len += last_good - start_index + 1;
        line_fit_known(s, o, len, lastA, lastB, lastInvert);
        index = last_good + 1;
      }
      if (good < 2 || index >= no_pixels) {
        double sx, sy, ex, ey;
        int idx = 0;
        while (min_distance(s.x(o + idx), s.y(o + idx), lastA, lastB, lastInvert) > line_error) idx++;
        closest_point(s.x(o + idx), s.y(o + idx), lastA, lastB, lastInvert, sx, sy);
        const int skipped = idx;
        idx = last_good;
        while (min_distance(s.x(o + idx), s.y(o + idx), lastA, lastB, lastInvert) > line_error) idx--;
        closest_point(s.x(o + idx), s.y(o + idx), lastA, lastB, lastInvert, ex, ey);
        if (out) {
          if (n_out >= cap) { W.err |= kErrLines; return n_out; }
          Line& L = out[n_out];
          L.a = lastA; L.b = lastB; L.invert = lastInvert; L.sx = sx; L.sy = sy; L.ex = ex; L.ey = ey;
          L.segmentNo = seg_no; L.firstPixelIndex = first_pixel_index + skipped; L.len = idx - skipped + 1;
        }
        n_out++;
        len = idx + 1;
        break;
      }
    }
    no_pixels -= len; o += len; first_pixel_index += len;
  }
  return n_out;
}

SDPL_ED_HD inline void update_line_parameters(Line& l) {               // EDLines.cpp:962-984
  const double dx = l.ex - l.sx, dy = l.ey - l.sy;
  if (fabs(dx) >= fabs(dy)) {
    l.invert = 0;
    if (fabs(dy) < 1e-3) { l.b = 0; l.a = (l.sy + l.ey) / 2; }
    else { l.b = dy / dx; l.a = l.sy - l.b * l.sx; }
  } else {
    l.invert = 1;
    if (fabs(dx) < 1e-3) { l.b = 0; l.a = (l.sx + l.ex) / 2; }
    else { l.b = dx / dy; l.a = l.sx - l.b * l.sy; }
  }
}
// EDLines::TryToJoinTwoLineSegments (EDLines.cpp:775-925): max distance 6.0, max error 1.3; ls1 is updated in place
SDPL_ED_HD inline bool try_join(Line& l1, const Line& l2) {
  double dx = l1.sx - l2.sx, dy = l1.sy - l2.sy;
  double d = sqrt(dx * dx + dy * dy), mn = d;
  dx = l1.sx - l2.ex; dy = l1.sy - l2.ey; d = sqrt(dx * dx + dy * dy); if (d < mn) mn = d;
  dx = l1.ex - l2.sx; dy = l1.ey - l2.sy; d = sqrt(dx * dx + dy * dy); if (d < mn) mn = d;
  dx = l1.ex - l2.ex; dy = l1.ey - l2.ey; d = sqrt(dx * dx + dy * dy); if (d < mn) mn = d;
  if (mn > 6.0) return false;
  dx = l1.sx - l1.ex; dy = l1.sy - l1.ey;
  const double prev_len = sqrt(dx * dx + dy * dy);
  dx = l2.sx - l2.ex; dy = l2.sy - l2.ey;
  const double next_len = sqrt(dx * dx + dy * dy);
  const Line* shorter = &l1; const Line* longer = &l2;
  if (prev_len > next_len) { shorter = &l2; longer = &l1; }
  double dist = min_distance(shorter->sx, shorter->sy, longer->a, longer->b, longer->invert);
  dist += min_distance((shorter->sx + shorter->ex) / 2.0, (shorter->sy + shorter->ey) / 2.0, longer->a, longer->b, longer->invert);
  dist += min_distance(shorter->ex, shorter->ey, longer->a, longer->b, longer->invert);
  dist /= 3.0;
  if (dist > 1.3) return false;
  // the pair of end points that are farthest apart become the new end points
  dx = fabs(l1.sx - l2.sx); dy = fabs(l1.sy - l2.sy); d = dx + dy;
  double mx = d; int which = 1;
  dx = fabs(l1.sx - l2.ex); dy = fabs(l1.sy - l2.ey); d = dx + dy; if (d > mx) { mx = d; which = 2; }
  dx = fabs(l1.ex - l2.sx); dy = fabs(l1.ey - l2.sy); d = dx + dy; if (d > mx) { mx = d; which = 3; }
  dx = fabs(l1.ex - l2.ex); dy = fabs(l1.ey - l2.ey); d = dx + dy; if (d > mx) { mx = d; which = 4; }
  if (which == 1) { l1.ex = l2.sx; l1.ey = l2.sy; }
  else if (which == 2) { l1.ex = l2.ex; l1.ey = l2.ey; }
  else if (which == 3) { l1.sx = l2.sx; l1.sy = l2.sy; }
  else { l1.sx = l1.ex; l1.sy = l1.ey; l1.ex = l2.ex; l1.ey = l2.ey; }
  if (l1.firstPixelIndex + l1.len + 5 >= l2.firstPixelIndex) l1.len += l2.len;
  else if (l2.len > l1.len) { l1.firstPixelIndex = l2.firstPixelIndex; l1.len = l2.len; }
  update_line_parameters(l1);
  return true;
}
// the lines L[0, n) of ONE segment: JoinCollinearLines' inner loop (EDLines.cpp:375-396); returns how many remain, packed at L[0, ..)
SDPL_ED_HD inline int join_segment(Line* L, int n) {
  if (n <= 0) return 0;
  int last = 0;
  for (int j = 1; j < n; j++) {
    if (!try_join(L[last], L[j])) {
      last++;
      if (last != j) L[last] = L[j];
    }
  }
  if (last != 0) {
    if (try_join(L[0], L[last])) last--;
  }
  return last + 1;
}
// EDLines::JoinCollinearLines (EDLines.cpp:362-400)
SDPL_ED_HD inline void join_collinear(Work& W) {
  Line* L = W.lines;
  int last = -1, i = 0;
  const int n = W.nlines;
  while (i < n) {
    const int seg_no = L[i].segmentNo;
    last++;
    if (last != i) L[last] = L[i];
    const int first = last;
    int count = 1;
    for (int j = i + 1; j < n; j++) {
      if (L[j].segmentNo != seg_no) break;
      if (!try_join(L[last], L[j])) {
        last++;
        if (last != j) L[last] = L[j];
      }
      count++;
    }
    if (first != last) {
      if (try_join(L[first], L[last])) last--;
    }
    i += count;
  }
  W.nlines = last + 1;
}

// ------------------------------------------------------------------------------------------------
// validation (EDLines.cpp:405-566, NFA.cpp:38-104)
// ------------------------------------------------------------------------------------------------
SDPL_ED_HD inline double my_atan2(const Work& W, double yy, double xx) {
  double y = fabs(yy), x = fabs(xx);
  bool invert = false;
  if (y > x) { const double t = x; x = y; y = t; invert = true; }
  if (x == 0) x = 0.000001;
  const double ratio = y / x;
  double angle = W.atan_lut[(int)(ratio * kAtanLut)];
  if (xx >= 0) {
    if (yy >= 0) { if (invert) angle = kPi / 2 - angle; }
    else { if (!invert) angle = kPi - angle; else angle = kPi / 2 + angle; }
  } else {
    if (yy >= 0) { if (!invert) angle = kPi - angle; else angle = kPi / 2 + angle; }
    else { if (invert) angle = kPi / 2 - angle; }
  }
  return angle;
}
SDPL_ED_HD inline bool nfa_ok(Work& W, int n, int k) {
  if (n >= W.nfa_n) { W.err |= kErrNfaTable; return false; }
  return k >= W.nfa_min_k[n];
}
// The source pixel the reference reads for "srcImg[flat]": its Mat is a ROI of LSDDetectorC's padded level buffer (19-pixel
// reflect-101 border, rows of width + 38 bytes, LSDDetector_custom.cpp:76-110), and `data + flat` is simply `flat` bytes past the ROI
// origin of that buffer.  Layout-independent form: padded position -> level coordinates -> reflect-101.
SDPL_ED_HD inline int reflect101(int v, int n) { if (v < 0) v = -v; if (v >= n) v = 2 * n - 2 - v; return v; }
SDPL_ED_HD inline int src_flat(const Work& W, int flat) {
  const int S = W.w + 38;
  const int P = 19 * S + 19 + flat;
  // P / S by multiplication: exact for P < 2^23 and S < 2^12 with a 35-bit reciprocal (sixteen of these per tested pixel)
  int q;
  if (W.src_magic && P < (1 << 23)) q = (int)(((unsigned long long)(unsigned)P * W.src_magic) >> 35); else q = P / S;
  const int y = reflect101(q - 19, W.h), x = reflect101(P - q * S - 19, W.w);
  return (int)W.src[y * W.src_stride + x];
}
SDPL_ED_HD inline bool pixel_aligned(const Work& W, int r, int c, double line_angle, double prec) {
  const int w = W.w;
  const int com1 = src_flat(W, (r + 1) * w + c + 1) - src_flat(W, (r - 1) * w + c - 1);
  const int com2 = src_flat(W, (r - 1) * w + c + 1) - src_flat(W, (r + 1) * w + c - 1);
  const int gx = com1 + com2 + src_flat(W, r * w + c + 1) - src_flat(W, r * w + c - 1);
  const int gy = com1 - com2 + src_flat(W, (r + 1) * w + c) - src_flat(W, (r - 1) * w + c);
  const double pixel_angle = my_atan2(W, (double)gx, (double)-gy);
  const double diff = fabs(line_angle - pixel_angle);
  return diff <= prec || diff >= kPi - prec;
}
SDPL_ED_HD inline double line_angle_of(const Line& l) {
  double a = l.invert == 0 ? atan(l.b) : atan(1.0 / l.b);
  if (a < 0) a += kPi;
  return a;
}
// ValidateLineSegmentRect + EnumerateRectPoints (EDLines.cpp:507-566, 988-1155): the points of the 2-pixel-wide rectangle around the
// line are visited in the reference's order and tested as they come (the reference stores them first; count / aligned are sums)
SDPL_ED_HD inline bool validate_rect(Work& W, const Line& l, double prec) {
  const double line_angle = line_angle_of(l);
  const double x1 = l.sx, y1 = l.sy, x2 = l.ex, y2 = l.ey;
  const double width = 2;
  double dx = x2 - x1, dy = y2 - y1;
  const double vlen = sqrt(dx * dx + dy * dy);
  // A line whose two end points coincide (TryToJoinTwoLineSegments can produce one) has no rectangle: the reference divides by zero
  // here, its enumeration never terminates and it writes past its point buffers (it crashes on frame 2822 of the synthetic sequence).
  // Decision: such a line is not validated.
  if (!(vlen > 0.0)) return false;
  dx = dx / vlen; dy = dy / vlen;
  double vxt[4], vyt[4], vx[4], vy[4];
  vxt[0] = x1 - dy * width / 2.0; vyt[0] = y1 + dx * width / 2.0;
  vxt[1] = x2 - dy * width / 2.0; vyt[1] = y2 + dx * width / 2.0;
  vxt[2] = x2 + dy * width / 2.0; vyt[2] = y2 - dx * width / 2.0;
  vxt[3] = x1 + dy * width / 2.0; vyt[3] = y1 - dx * width / 2.0;
  int offset;
  if (x1 < x2 && y1 <= y2) offset = 0;
  else if (x1 >= x2 && y1 < y2) offset = 1;
  else if (x1 > x2 && y1 >= y2) offset = 2;
  else offset = 3;
  for (int n = 0; n < 4; n++) { vx[n] = vxt[(offset + n) % 4]; vy[n] = vyt[(offset + n) % 4]; }
  int x = (int)ceil(vx[0]) - 1;
  int y = (int)ceil(vy[0]);
  double ys = -1.7976931348623157e308, ye = -1.7976931348623157e308;
  int count = 0, aligned = 0;
  long guard = 0;
  while (true) {
    y++;
    while ((double)y > ye && (double)x <= vx[2]) {
      x++;
      if ((double)x > vx[2]) break;
      if ((double)x < vx[3]) {
        if (fabs(vx[0] - vx[3]) <= 0.01) {
          if (vy[0] < vy[3]) ys = vy[0];
          else if (vy[0] > vy[3]) ys = vy[3];
          else ys = vy[0] + (x - vx[0]) * (vy[3] - vy[0]) / (vx[3] - vx[0]);
        } else ys = vy[0] + (x - vx[0]) * (vy[3] - vy[0]) / (vx[3] - vx[0]);
      } else {
        if (fabs(vx[3] - vx[2]) <= 0.01) {
          if (vy[3] < vy[2]) ys = vy[3];
          else if (vy[3] > vy[2]) ys = vy[2];
          else ys = vy[3] + (x - vx[3]) * (y2 - vy[3]) / (vx[2] - vx[3]);
        } else ys = vy[3] + (x - vx[3]) * (vy[2] - vy[3]) / (vx[2] - vx[3]);
      }
      if ((double)x < vx[1]) {
        if (fabs(vx[0] - vx[1]) <= 0.01) {
          if (vy[0] < vy[1]) ye = vy[1];
          else if (vy[0] > vy[1]) ye = vy[0];
          else ye = vy[0] + (x - vx[0]) * (vy[1] - vy[0]) / (vx[1] - vx[0]);
        } else ye = vy[0] + (x - vx[0]) * (vy[1] - vy[0]) / (vx[1] - vx[0]);
      } else {
        if (fabs(vx[1] - vx[2]) <= 0.01) {
          if (vy[1] < vy[2]) ye = vy[2];
          else if (vy[1] > vy[2]) ye = vy[1];
          else ye = vy[1] + (x - vx[1]) * (vy[2] - vy[1]) / (vx[2] - vx[1]);
        } else ye = vy[1] + (x - vx[1]) * (vy[2] - vy[1]) / (vx[2] - vx[1]);
      }
      y = (int)ceil(ys);
    }
    if ((double)x > vx[2]) break;
    if (++guard > 1000000) { W.err |= kErrNfaTable; break; }
    // the point (x, y)
    const int r = y, c = x;
    if (r <= 0 || r >= W.h - 1 || c <= 0 || c >= W.w - 1) continue;
    count++;
    if (pixel_aligned(W, r, c, line_angle, prec)) aligned++;
  }
  return nfa_ok(W, count, aligned);
}
// one line of ValidateLineSegments (EDLines.cpp:413-493)
SDPL_ED_HD inline bool validate_one(Work& W, const Line& l) {
  const double prec = (22.5 / 180) * kPi;
  bool valid = false;
  if (l.len >= 80) valid = true;
  else if (l.len <= 25) valid = validate_rect(W, l, prec);
  else {
    const double line_angle = line_angle_of(l);
    const int* pixels = W.seg_px + W.seg_off[l.segmentNo];      // the segment's pixel 0 (sic), x and y swapped (sic)
    int aligned = 0, count = 0;
    for (int j = 0; j < l.len; j++) {
      const int r = px_x(pixels[j]), c = px_y(pixels[j]);
      if (r <= 0 || r >= W.h - 1 || c <= 0 || c >= W.w - 1) continue;
      count++;
      if (pixel_aligned(W, r, c, line_angle, prec)) aligned++;
    }
    valid = nfa_ok(W, count, aligned);
    if (!valid) valid = validate_rect(W, l, prec);
  }
  return valid;
}
// EDLines::ValidateLineSegments (EDLines.cpp:405-505)
SDPL_ED_HD inline void validate_lines(Work& W) {
  int n_valid = 0;
  for (int i = 0; i < W.nlines; i++) {
    const bool valid = validate_one(W, W.lines[i]);
    if (valid) { if (i != n_valid) W.lines[n_valid] = W.lines[i]; n_valid++; }
  }
  W.nlines = n_valid;
}
// everything after the anchors: the sequential part of one task
SDPL_ED_HD inline void run_task(Work& W) {
  W.err = 0; W.nlines = 0; W.nseg = 0;
  link_anchors(W);
  if (W.err) return;
  for (int s = 0; s < W.nseg && !W.err; s++) W.nlines += split_segment(W, s, W.lines + W.nlines, W.lines_cap - W.nlines);
  if (W.err) return;
  join_collinear(W);
  validate_lines(W);
}

// ------------------------------------------------------------------------------------------------
// Host-side tables (the C library's atan / log / exp / pow / sinh / log10 are evaluated here, once, on the host -- by the oracle and
// by the product library alike -- and the results are what the sequential code above reads).
// ------------------------------------------------------------------------------------------------
namespace host {
inline double log_gamma(double x) {                                     // NFA.cpp:191-217
  if (x > 15) return 0.918938533204673 + (x - 0.5) * log(x) - x + 0.5 * x * log(x * sinh(1 / x) + 1 / (810.0 * pow(x, 6.0)));
  static const double q[7] = {75122.6331530, 80916.6278952, 36308.2951477, 8687.24529705, 1168.92649479, 83.8676043424, 2.50662827511};
  double a = (x + 0.5) * log(x + 5.5) - (x + 5.5), b = 0.0;
  for (int n = 0; n < 7; n++) { a -= log(x + (double)n); b += q[n] * pow(x, (double)n); }
  return a + log(b);
}
inline bool double_equal(double a, double b) {                          // NFA.cpp:219-240
  if (a == b) return true;
  const double abs_diff = fabs(a - b), aa = fabs(a), bb = fabs(b);
  double abs_max = aa > bb ? aa : bb;
  if (abs_max < 2.2250738585072014e-308) abs_max = 2.2250738585072014e-308;
  return (abs_diff / abs_max) <= (100.0 * 2.220446049250313e-16);
}
inline double nfa(int n, int k, double prob, double logNT) {            // NFALUT::nfa, NFA.cpp:106-189
  const double tolerance = 0.1;
  if (n < 0 || k < 0 || k > n || prob <= 0.0 || prob >= 1.0) return -1.0;
  if (n == 0 || k == 0) return -logNT;
  if (n == k) return -logNT - (double)n * log10(prob);
  const double p_term = prob / (1.0 - prob);
  const double log1term = log_gamma((double)n + 1.0) - log_gamma((double)k + 1.0) - log_gamma((double)(n - k) + 1.0) + (double)k * log(prob) +
                          (double)(n - k) * log(1.0 - prob);
  double term = exp(log1term);
  if (double_equal(term, 0.0)) {
    if ((double)k > (double)n * prob) return -log1term / 2.30258509299404568402 - logNT;
    return -logNT;
  }
  double bin_tail = term;
  for (int i = k + 1; i <= n; i++) {
    const double bin_term = (double)(n - i + 1) * (1.0 / (double)i);
    const double mult_term = bin_term * p_term;
    term *= mult_term;
    bin_tail += term;
    if (bin_term < 1.0) {
      const double err = term * ((1.0 - pow(mult_term, (double)(n - i + 1))) / (1.0 - mult_term) - 1.0);
      if (err < tolerance * fabs(-log10(bin_tail) - logNT) * bin_tail) break;
    }
  }
  return -log10(bin_tail) - logNT;
}
// min_k[n], n < n_max: the smallest k for which NFALUT::checkValidationByNFA(n, k) holds for an image of w x h (n + 1 = never).
// n < lutSize = (w + h) / 8 goes through the reference's LUT (NFALUT::NFALUT, NFA.cpp:5-31), larger n through nfa(n, k) >= 0.
// Returns false when the direct branch is not monotone in k (then a threshold cannot represent it; never observed).
inline bool nfa_table(int w, int h, int n_max, int* min_k) {
  const double prob = 0.125, logNT = 2.0 * (log10((double)w) + log10((double)h));
  const int lut_size = (w + h) / 8;
  int j = 1;
  bool ok = true;
  for (int i = 0; i < n_max; i++) min_k[i] = i + 1;
  if (lut_size > 0 && n_max > 0) min_k[0] = 1;                          // LUT[0] = 1: k >= 1 can never hold for n = 0
  for (int i = 1; i < lut_size && i < n_max; i++) {
    int lut = lut_size + 1;
    double ret = nfa(i, j, prob, logNT);
    bool found = true;
    if (ret < 0) {
      while (j < i) { j++; ret = nfa(i, j, prob, logNT); if (ret >= 0) break; }
      if (ret < 0) found = false;
    }
    if (found) lut = j;
    min_k[i] = lut > i ? i + 1 : lut;
  }
  // direct branch: nfa(n, k) >= 0 is an up-set in k whose lower end does not fall as n grows, so the threshold is followed from n to
  // n + 1 instead of scanning every k (a full scan of n_max^2 / 2 tail sums takes seconds); the two neighbours of the threshold are
  // evaluated to confirm it
  int k0 = 0;
  for (int n = lut_size > 0 ? lut_size : 0; n < n_max; n++) {
    while (k0 > 0 && nfa(n, k0 - 1, prob, logNT) >= 0.0) k0--;
    while (k0 <= n && !(nfa(n, k0, prob, logNT) >= 0.0)) k0++;
    min_k[n] = k0 <= n ? k0 : n + 1;
    if (k0 < n && !(nfa(n, k0 + 1, prob, logNT) >= 0.0)) ok = false;
    if (k0 > n) k0 = n + 1;
  }
  return ok;
}
inline int min_line_len(int w, int h) {                                 // EDLines::ComputeMinLineLength + the floor of 9, EDLines.cpp:16-20, 255-264
  const double logNT = 2.0 * (log10((double)w) + log10((double)h));
  int m = (int)round((-logNT / log10(0.125)) * 0.5);
  return m < 9 ? 9 : m;
}
inline void atan_table(double* lut) { for (int i = 0; i <= kAtanLut; i++) lut[i] = atan((double)i / kAtanLut); }
}  // namespace host

}  // namespace sdpl_ed
#endif
