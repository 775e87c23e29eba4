// oracle/cvprim.cpp -- ORACLE (test infrastructure): bit-exact CPU restatements of the OpenCV
// primitives the reference front-end calls.  OpenCV is an un-vendored dependency of the reference
// ("tested with OpenCV 3.4", README.md:12); each routine below restates the published integer /
// float arithmetic of that primitive and is pinned against python cv2 in tests/test_oracle_vs_cv2.py.
//
// Call sites in the reference:
//   resize INTER_LINEAR      src/ORBextractor.cc:1125, 3rdparty/line_descriptor/src/LSDDetector_custom.cpp:98
//   copyMakeBorder REFLECT101 src/ORBextractor.cc:1127,1132, LSDDetector_custom.cpp:100,105
//   FAST(img,kps,th,true)    src/ORBextractor.cc:798,803
//   GaussianBlur             src/ORBextractor.cc:1084, binary_descriptor_custom.cpp:358, (LSD internal)
//   fastAtan2                src/ORBextractor.cc:92 (and LSD internal)
//   pyrDown / Sobel          binary_descriptor_custom.cpp:366,395-396
//   resize INTER_LINEAR_EXACT (inside cv::LineSegmentDetector, LSDDetector_custom.cpp:307)
#include "oracle_internal.h"
#include "../include/sdpl_trig.h"
#include <cmath>
#include <cfloat>
#include <cstring>
#include <vector>
#include <algorithm>

namespace orc {

int cv_round(float v) { return (int)lrintf(v); }    // cvRound: round-half-even (SSE cvtss2si)
int cv_round(double v) { return (int)lrint(v); }
int cv_floor(double v) { int i = (int)v; return i - (i > v); }
int cv_floor(float v) { int i = (int)v; return i - (i > v); }
int cv_ceil(double v) { int i = (int)v; return i + (i < v); }

int reflect101(int p, int len) {
  if (len == 1) return 0;
  while (p < 0 || p >= len) {
    if (p < 0) p = -p;
    else p = 2 * (len - 1) - p;
  }
  return p;
}

static inline short sat_short_from_float(float v) {
  int i = cv_round(v);
  return (short)std::min(32767, std::max(-32768, i));
}

// cv::resize(..., INTER_LINEAR) on 8UC1.  Fixed point Q11 coefficients, int horizontal pass,
// vertical pass ((b0*(S0>>4))>>16 + (b1*(S1>>4))>>16 + 2)>>2.  Exact 2x2 decimation switches to
// INTER_AREA ((a+b+c+d+2)>>2) as cv::resize does.
void resize_linear_u8(const uint8_t* src, int sw, int sh, int sstride, uint8_t* dst, int dw, int dh, int dstride) {
  double inv_scale_x = (double)dw / sw, inv_scale_y = (double)dh / sh;
  double scale_x = 1. / inv_scale_x, scale_y = 1. / inv_scale_y;
  int iscale_x = cv_round(scale_x), iscale_y = cv_round(scale_y);
  bool area_fast = std::abs(scale_x - iscale_x) < DBL_EPSILON && std::abs(scale_y - iscale_y) < DBL_EPSILON;
  if (area_fast && iscale_x == 2 && iscale_y == 2) {
    for (int y = 0; y < dh; y++) {
      const uint8_t* s0 = src + (size_t)(2 * y) * sstride;
      const uint8_t* s1 = s0 + sstride;
      uint8_t* d = dst + (size_t)y * dstride;
      for (int x = 0; x < dw; x++) d[x] = (uint8_t)((s0[2 * x] + s0[2 * x + 1] + s1[2 * x] + s1[2 * x + 1] + 2) >> 2);
    }
    return;
  }
  std::vector<int> xofs(dw), yofs(dh);
  std::vector<short> ia(2 * dw), ib(2 * dh);
  for (int dx = 0; dx < dw; dx++) {
    float fx = (float)((dx + 0.5) * scale_x - 0.5);
    int sx = cv_floor(fx);
    fx -= sx;
    if (sx < 0) { fx = 0; sx = 0; }
    if (sx >= sw - 1) { fx = 0; sx = sw - 1; }
    xofs[dx] = sx;
    ia[2 * dx] = sat_short_from_float((1.f - fx) * 2048);
    ia[2 * dx + 1] = sat_short_from_float(fx * 2048);
  }
  for (int dy = 0; dy < dh; dy++) {
    float fy = (float)((dy + 0.5) * scale_y - 0.5);
    int sy = cv_floor(fy);
    fy -= sy;
    yofs[dy] = sy;
    ib[2 * dy] = sat_short_from_float((1.f - fy) * 2048);
    ib[2 * dy + 1] = sat_short_from_float(fy * 2048);
  }
  std::vector<int> r0(dw), r1(dw);
  auto hrow = [&](int sy, std::vector<int>& out) {
    sy = std::min(std::max(sy, 0), sh - 1);
    const uint8_t* s = src + (size_t)sy * sstride;
    for (int dx = 0; dx < dw; dx++) {
      int sx = xofs[dx];
      int v = s[sx] * ia[2 * dx];
      if (sx + 1 < sw) v += s[sx + 1] * ia[2 * dx + 1];
      out[dx] = v;
    }
  };
  for (int dy = 0; dy < dh; dy++) {
    hrow(yofs[dy], r0);
    hrow(yofs[dy] + 1, r1);
    int b0 = ib[2 * dy], b1 = ib[2 * dy + 1];
    uint8_t* d = dst + (size_t)dy * dstride;
    for (int dx = 0; dx < dw; dx++) {
      int v = (((b0 * (r0[dx] >> 4)) >> 16) + ((b1 * (r1[dx] >> 4)) >> 16) + 2) >> 2;
      d[dx] = (uint8_t)std::min(255, std::max(0, v));
    }
  }
}

// cv::resize(..., INTER_LINEAR_EXACT) on 8UC1 (resize_bitExact): Q8.8 coefficients from IEEE double
// (softdouble) arithmetic, horizontal pass to Q8.8, vertical to Q16.16, (v + 2^15) >> 16.
namespace {
struct ExactCoef { int ofs; uint16_t c0, c1; };
void exact_coeffs(double inv_scale, int srcsize, int dstsize, std::vector<ExactCoef>& out, int& minofst, int& maxofst) {
  volatile double scale = 1.0 / inv_scale;
  out.resize(dstsize);
  minofst = 0; maxofst = dstsize;
  for (int val = 0; val < dstsize; val++) {
    volatile double t = scale * ((double)val + 0.5);
    double fval = t - 0.5;
    int ival = cv_floor(fval);
    out[val].ofs = 0; out[val].c0 = 256; out[val].c1 = 0;
    if (ival >= 0 && srcsize > 1) {
      if (ival < srcsize - 1) {
        out[val].ofs = ival;
        double fr = fval - (double)ival;
        int c1 = fr < 0 ? 0 : cv_round(fr * 256.0);
        out[val].c1 = (uint16_t)c1;
        out[val].c0 = (uint16_t)(256 - c1);
      } else {
        out[val].ofs = srcsize - 1;
        maxofst = std::min(maxofst, val);
      }
    } else {
      minofst = std::max(minofst, val + 1);
    }
  }
}
}  // namespace

void resize_linear_exact_u8(const uint8_t* src, int sw, int sh, int sstride, uint8_t* dst, int dw, int dh, int dstride,
                            double inv_scale_x, double inv_scale_y) {
  std::vector<ExactCoef> cx, cy;
  int xmin, xmax, ymin, ymax;
  exact_coeffs(inv_scale_x, sw, dw, cx, xmin, xmax);
  exact_coeffs(inv_scale_y, sh, dh, cy, ymin, ymax);
  std::vector<uint16_t> r0(dw), r1(dw);
  auto hrow = [&](int sy, std::vector<uint16_t>& out) {
    const uint8_t* s = src + (size_t)sy * sstride;
    for (int x = 0; x < dw; x++) {
      if (x < xmin) out[x] = (uint16_t)(s[0] << 8);
      else if (x >= xmax) out[x] = (uint16_t)(s[sw - 1] << 8);
      else out[x] = (uint16_t)(cx[x].c0 * s[cx[x].ofs] + cx[x].c1 * s[cx[x].ofs + 1]);
    }
  };
  for (int y = 0; y < dh; y++) {
    uint8_t* d = dst + (size_t)y * dstride;
    if (y < ymin || y >= ymax) {
      hrow(y < ymin ? 0 : sh - 1, r0);
      for (int x = 0; x < dw; x++) d[x] = (uint8_t)((r0[x] + 128) >> 8);
      continue;
    }
    hrow(cy[y].ofs, r0);
    hrow(cy[y].ofs + 1, r1);
    uint32_t m0 = cy[y].c0, m1 = cy[y].c1;
    for (int x = 0; x < dw; x++) {
      uint32_t v = m0 * r0[x] + m1 * r1[x];
      uint32_t o = (v + (1u << 15)) >> 16;
      d[x] = (uint8_t)std::min(255u, o);
    }
  }
}

void border_reflect101_u8(const uint8_t* src, int w, int h, int sstride, uint8_t* dst, int border, int dstride) {
  for (int y = -border; y < h + border; y++) {
    const uint8_t* s = src + (size_t)reflect101(y, h) * sstride;
    uint8_t* d = dst + (size_t)(y + border) * dstride;
    for (int x = -border; x < w + border; x++) d[x + border] = s[reflect101(x, w)];
  }
}

// GaussianBlur on 8UC1, fixed-point path (Q8.8 kernel, Q16.16 vertical, round half up), BORDER_REFLECT_101.
static const int kG7s2[7] = {18, 34, 48, 56, 48, 34, 18};   // 7x7 sigma 2      (src/ORBextractor.cc:1084)
static const int kG5s1[5] = {14, 62, 104, 62, 14};          // 5x5 sigma 1      (binary_descriptor_custom.cpp:358)
static const int kG7s075[7] = {0, 4, 56, 136, 56, 4, 0};    // 7x7 sigma 0.75   (LSD: sigma_scale/scale = 0.6/0.8)
void gaussian_blur_u8(const uint8_t* src, int w, int h, int sstride, uint8_t* dst, int dstride, int kind) {
  const int* k = kind == 0 ? kG7s2 : kind == 1 ? kG5s1 : kG7s075;
  const int n = kind == 1 ? 5 : 7, r = n / 2;
  // horizontal pass into a u16 plane (Q8.8 sums fit 16 bits); each row is first extended by reflect-101 so that the
  // tap loop has no border logic (same values, same order of additions as the per-tap reflect)
  std::vector<uint16_t> tmp((size_t)w * h);
  std::vector<uint8_t> ext((size_t)w + 2 * r);
  for (int y = 0; y < h; y++) {
    const uint8_t* sr = src + (size_t)y * sstride;
    for (int i = 0; i < r; i++) { ext[i] = sr[reflect101(i - r, w)]; ext[r + w + i] = sr[reflect101(w + i, w)]; }
    memcpy(&ext[r], sr, w);
    uint16_t* t = &tmp[(size_t)y * w];
    const uint8_t* e = ext.data();
    if (n == 7) {
      for (int x = 0; x < w; x++)
        t[x] = (uint16_t)((uint32_t)k[0] * e[x] + (uint32_t)k[1] * e[x + 1] + (uint32_t)k[2] * e[x + 2] + (uint32_t)k[3] * e[x + 3] +
                          (uint32_t)k[4] * e[x + 4] + (uint32_t)k[5] * e[x + 5] + (uint32_t)k[6] * e[x + 6]);
    } else {
      for (int x = 0; x < w; x++)
        t[x] = (uint16_t)((uint32_t)k[0] * e[x] + (uint32_t)k[1] * e[x + 1] + (uint32_t)k[2] * e[x + 2] + (uint32_t)k[3] * e[x + 3] +
                          (uint32_t)k[4] * e[x + 4]);
    }
  }
  // vertical pass (Q16.16, round half up); rows addressed through reflect-101 row pointers
  std::vector<uint8_t> out((size_t)w * h);
  for (int y = 0; y < h; y++) {
    const uint16_t* rows[7];
    for (int i = 0; i < n; i++) rows[i] = &tmp[(size_t)reflect101(y + i - r, h) * w];
    uint8_t* o = &out[(size_t)y * w];
    if (n == 7) {
      for (int x = 0; x < w; x++) {
        const uint32_t acc = (uint32_t)k[0] * rows[0][x] + (uint32_t)k[1] * rows[1][x] + (uint32_t)k[2] * rows[2][x] + (uint32_t)k[3] * rows[3][x] +
                             (uint32_t)k[4] * rows[4][x] + (uint32_t)k[5] * rows[5][x] + (uint32_t)k[6] * rows[6][x];
        o[x] = (uint8_t)std::min(255u, (acc + (1u << 15)) >> 16);
      }
    } else {
      for (int x = 0; x < w; x++) {
        const uint32_t acc = (uint32_t)k[0] * rows[0][x] + (uint32_t)k[1] * rows[1][x] + (uint32_t)k[2] * rows[2][x] + (uint32_t)k[3] * rows[3][x] +
                             (uint32_t)k[4] * rows[4][x];
        o[x] = (uint8_t)std::min(255u, (acc + (1u << 15)) >> 16);
      }
    }
  }
  for (int y = 0; y < h; y++) memcpy(dst + (size_t)y * dstride, &out[(size_t)y * w], w);
}

// cv::pyrDown 8UC1: separable [1 4 6 4 1], (v + 128) >> 8, BORDER_REFLECT_101, dst size given.
void pyrdown_u8(const uint8_t* src, int w, int h, int sstride, uint8_t* dst, int dw, int dh, int dstride) {
  std::vector<int> rows((size_t)5 * dw);
  for (int y = 0; y < dh; y++) {
    for (int k = 0; k < 5; k++) {
      const uint8_t* s = src + (size_t)reflect101(2 * y + k - 2, h) * sstride;
      int* r = &rows[(size_t)k * dw];
      for (int x = 0; x < dw; x++) {
        int c = 2 * x;
        r[x] = s[reflect101(c - 2, w)] + s[reflect101(c + 2, w)] + 4 * (s[reflect101(c - 1, w)] + s[reflect101(c + 1, w)]) +
               6 * s[reflect101(c, w)];
      }
    }
    uint8_t* d = dst + (size_t)y * dstride;
    for (int x = 0; x < dw; x++) {
      int v = rows[x] + rows[4 * dw + x] + 4 * (rows[dw + x] + rows[3 * dw + x]) + 6 * rows[2 * dw + x];
      d[x] = (uint8_t)((v + 128) >> 8);
    }
  }
}

// cv::Sobel(src, CV_16S, 1,0,3) and (0,1,3), BORDER_REFLECT_101.
void sobel3_s16(const uint8_t* src, int w, int h, int sstride, int16_t* dx, int16_t* dy) {
  for (int y = 0; y < h; y++) {
    const uint8_t* r0 = src + (size_t)reflect101(y - 1, h) * sstride;
    const uint8_t* r1 = src + (size_t)y * sstride;
    const uint8_t* r2 = src + (size_t)reflect101(y + 1, h) * sstride;
    for (int x = 0; x < w; x++) {
      int xm = reflect101(x - 1, w), xp = reflect101(x + 1, w);
      int gx = (r0[xp] - r0[xm]) + 2 * (r1[xp] - r1[xm]) + (r2[xp] - r2[xm]);
      int gy = (r2[xm] + 2 * r2[x] + r2[xp]) - (r0[xm] + 2 * r0[x] + r0[xp]);
      dx[(size_t)y * w + x] = (int16_t)gx;
      dy[(size_t)y * w + x] = (int16_t)gy;
    }
  }
}

// cv::fastAtan2: 7th-order odd polynomial in strict fp32 (no FMA contraction), degrees in [0,360).
float fast_atan2(float y, float x) {
  static const float scale = (float)(180.0 / 3.141592653589793238462643383279502884);
  static const float p1 = 0.9997878412794807f * scale, p3 = -0.3258083974640975f * scale,
                     p5 = 0.1555786518463281f * scale, p7 = -0.04432655554792128f * scale;
  volatile float ax = std::fabs(x), ay = std::fabs(y);
  volatile float a, c, c2, t;
  if (ax >= ay) {
    t = ax + (float)DBL_EPSILON;
    c = ay / t;
    c2 = c * c;
    t = p7 * c2; t = t + p5; t = t * c2; t = t + p3; t = t * c2; t = t + p1;
    a = t * c;
  } else {
    t = ay + (float)DBL_EPSILON;
    c = ax / t;
    c2 = c * c;
    t = p7 * c2; t = t + p5; t = t * c2; t = t + p3; t = t * c2; t = t + p1;
    t = t * c;
    a = 90.f - t;
  }
  if (x < 0) a = 180.f - a;
  if (y < 0) a = 360.f - a;
  return a;
}

// FAST-9/16 corner score of a pixel: max over the 16 arcs of 9 contiguous ring pixels of the minimum
// signed difference (bright or dark), minus 1.  A pixel is a corner at threshold t  <=>  score >= t.
static const int kRing[16][2] = {{0, 3},  {1, 3},   {2, 2},   {3, 1},   {3, 0},  {3, -1}, {2, -2}, {1, -3},
                                 {0, -3}, {-1, -3}, {-2, -2}, {-3, -1}, {-3, 0}, {-3, 1}, {-2, 2}, {-1, 3}};
int fast_score(const uint8_t* p, int stride) {
  int d[25];
  int v = p[0];
  for (int k = 0; k < 16; k++) d[k] = v - p[kRing[k][1] * stride + kRing[k][0]];
  for (int k = 16; k < 25; k++) d[k] = d[k - 16];
  int best = -256;
  for (int k = 0; k < 16; k++) {
    int mn = d[k], mx = d[k];
    for (int j = 1; j < 9; j++) { mn = std::min(mn, d[k + j]); mx = std::max(mx, d[k + j]); }
    best = std::max(best, std::max(mn, -mx));
  }
  return best - 1;
}

// cv::FAST(img, kps, threshold, nonmaxSuppression=true) on a stand-alone image: interior [3,w-3)x[3,h-3),
// keep iff score strictly greater than the 8 neighbours (non-corners count as 0), row-major order.
int fast9_nms(const uint8_t* img, int w, int h, int stride, int threshold, std::vector<int>& xs, std::vector<int>& ys,
              std::vector<int>& sc) {
  xs.clear(); ys.clear(); sc.clear();
  if (w < 7 || h < 7) return 0;
  std::vector<int> s((size_t)w * h, 0);
  // An arc of 9 contiguous ring pixels contains at least one pixel of every opposite pair (k, k+8): a corner at `threshold`
  // needs, for all 8 pairs, one member brighter than centre+threshold (bright arc) or, for all 8 pairs, one member darker
  // than centre-threshold (dark arc).  Pixels failing this necessary test have score < threshold (stored as 0 either
  // way), so the full score is only computed for the survivors -- the same early-out cv::FAST uses.
  int ofs[16];
  for (int k = 0; k < 16; k++) ofs[k] = kRing[k][1] * stride + kRing[k][0];
  for (int y = 3; y < h - 3; y++)
    for (int x = 3; x < w - 3; x++) {
      const uint8_t* p = img + (size_t)y * stride + x;
      const int hi = p[0] + threshold, lo = p[0] - threshold;
      bool bright = true, dark = true;
      for (int k = 0; k < 8 && (bright || dark); k++) {
        const int a = p[ofs[k]], b = p[ofs[k + 8]];
        bright = bright && (a > hi || b > hi);
        dark = dark && (a < lo || b < lo);
      }
      int v = 0;
      if (bright || dark) {
        v = fast_score(p, stride);
        if (v < threshold) v = 0;
      }
      s[(size_t)y * w + x] = v;
    }
  for (int y = 3; y < h - 3; y++)
    for (int x = 3; x < w - 3; x++) {
      int v = s[(size_t)y * w + x];
      if (v <= 0) continue;
      bool ok = true;
      for (int dy = -1; dy <= 1 && ok; dy++)
        for (int dx = -1; dx <= 1; dx++) {
          if (!dx && !dy) continue;
          if (!(v > s[(size_t)(y + dy) * w + x + dx])) { ok = false; break; }
        }
      if (ok) { xs.push_back(x); ys.push_back(y); sc.push_back(v); }
    }
  return (int)xs.size();
}

}  // namespace orc

extern "C" {
void orc_resize_linear_u8(const uint8_t* s, int sw, int sh, int ss, uint8_t* d, int dw, int dh, int ds) {
  orc::resize_linear_u8(s, sw, sh, ss, d, dw, dh, ds);
}
void orc_resize_linear_exact_u8(const uint8_t* s, int sw, int sh, int ss, uint8_t* d, int dw, int dh, int ds, double fx,
                                double fy) {
  orc::resize_linear_exact_u8(s, sw, sh, ss, d, dw, dh, ds, fx > 0 ? fx : (double)dw / sw, fy > 0 ? fy : (double)dh / sh);
}
void orc_border_reflect101_u8(const uint8_t* s, int w, int h, int ss, uint8_t* d, int b, int ds) {
  orc::border_reflect101_u8(s, w, h, ss, d, b, ds);
}
void orc_gaussian_blur_u8(const uint8_t* s, int w, int h, int ss, uint8_t* d, int ds, int kind) {
  orc::gaussian_blur_u8(s, w, h, ss, d, ds, kind);
}
void orc_pyrdown_u8(const uint8_t* s, int w, int h, int ss, uint8_t* d, int dw, int dh, int ds) {
  orc::pyrdown_u8(s, w, h, ss, d, dw, dh, ds);
}
void orc_sobel3_s16(const uint8_t* s, int w, int h, int ss, int16_t* dx, int16_t* dy) { orc::sobel3_s16(s, w, h, ss, dx, dy); }
float orc_fast_atan2(float y, float x) { return orc::fast_atan2(y, x); }
int orc_fast9_nms(const uint8_t* img, int w, int h, int stride, int th, int* xs, int* ys, int* sc, int cap) {
  std::vector<int> a, b, c;
  int n = orc::fast9_nms(img, w, h, stride, th, a, b, c);
  for (int i = 0; i < n && i < cap; i++) { xs[i] = a[i]; ys[i] = b[i]; sc[i] = c[i]; }
  return n;
}
int orc_cv_round_f(float v) { return orc::cv_round(v); }
int orc_cv_round_d(double v) { return orc::cv_round(v); }
void orc_sdpl_sincos(double x, double* s, double* c) { sdpl_sincos(x, s, c); }
}
