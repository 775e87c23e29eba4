// oracle/lsd_oracle.cpp -- ORACLE (test infrastructure): CPU restatement of cv::LineSegmentDetector
// (OpenCV imgproc/lsd.cpp, a port of von Gioi's LSD), which the reference calls through
// cv::createLineSegmentDetector(...)->detect(level) at 3rdparty/line_descriptor/src/LSDDetector_custom.cpp:291-309.
// OpenCV is an un-vendored dependency of the reference ("tested with OpenCV 3.4"); this file restates the published
// algorithm and is pinned black-box against python cv2's LineSegmentDetector in
// tests/test_oracle_vs_cv2.py.  Strict IEEE double/float, no FMA contraction.
//
// Oracle decisions:
//  (vi) seed order: descending gradient bin; tie_mode 0 = row-major (stable) inside a bin -- what the GPU path
//       implements; tie_mode 1 = the order libstdc++'s std::sort leaves (what an OpenCV binary built with libstdc++
//       produces), kept only to quantify the difference against cv2.
//  trig in region growing: OpenCV calls sincosf; the oracle evaluates cos/sin in double on the float-rounded angle,
//       rounds to float and accumulates in float (identical except where glibc's sincosf is not correctly rounded).
//  (ix) every double cos / sin of the LSD path (region2rect's rectangle axes, the per-pixel unit vectors) is
//       include/sdpl_trig.h -- strict-IEEE, < 0.6 ulp, the SAME source the CUDA kernels compile -- instead of the C
//       library's: glibc and CUDA differ in the last bit for ~0.1 % of the arguments (glibc's own FMA / non-FMA builds
//       differ as well), and one flipped edge pixel of a rotated rectangle changes an NFA verdict about once per 3000
//       rectangles.  The cv2 pin (tests/test_oracle_vs_cv2.py) holds with either.
//  rect_nfa: follows the OpenCV 4.x implementation (double vertices, ceil/int column limits), recovered from the
//       cv2 4.13 binary and verified bit-exact against it; the OpenCV 3.4-era integer-division scan is kept behind
//       orc_lsd_set_nfa_variant(4) only to document the difference (the reference README pins "OpenCV 3.4").
#include "oracle_internal.h"
#include "../include/sdpl_trig.h"
#include <cmath>
#include <cfloat>
#include <cstring>
#include <algorithm>

namespace orc {
namespace {

const double kPI = 3.1415926535897932384626433832795;
const double kNotDef = -1024.0;
const double kDegToRad = kPI / 180;
const double k3_2PI = (3 * kPI) / 2;
const double k2PI = 2 * kPI;
const double kLn10 = 2.30258509299404568402;

int g_trig_mode = 0;
int g_nfa_variant = 16 + 4;  // bit 16: OpenCV 4.x rect scan (else the 3.4-era integer scan); bit 4: log_gamma(n+1)
std::vector<double> g_dbg;  // per output line: width, p, log_nfa, n, k
int g_last_n = 0, g_last_k = 0;
std::vector<uint8_t> g_last_scaled;
int g_last_w = 0, g_last_h = 0;
std::vector<double> g_cand;  // per rectangle handed to rect_improve: x1,y1,x2,y2,width, first rect_nfa value, final log_nfa
std::vector<int> g_trace;   // per processed seed: position in the seed order, pixels expanded (all growths), first region size
int g_trace_on = 0, g_expanded = 0;
std::vector<int> g_trace_n;  // list length after every expansion, -1 between growths (frontier model for the GPU batch size)

struct RegPoint { int x, y; double angle, modgrad; };
struct Rect { double x1, y1, x2, y2, width, x, y, theta, dx, dy, prec, p; };
struct NormPoint { int x, y, norm; };

inline double dist_sq(double x1, double y1, double x2, double y2) { return (x2 - x1) * (x2 - x1) + (y2 - y1) * (y2 - y1); }
inline double dist(double x1, double y1, double x2, double y2) { return std::sqrt(dist_sq(x1, y1, x2, y2)); }
inline double angle_diff_signed(double a, double b) {
  double d = a - b;
  while (d <= -kPI) d += k2PI;
  while (d > kPI) d -= k2PI;
  return d;
}
inline double angle_diff(double a, double b) { return std::fabs(angle_diff_signed(a, b)); }
inline bool double_equal(double a, double b) {
  if (a == b) return true;
  double ad = std::fabs(a - b), aa = std::fabs(a), bb = std::fabs(b);
  double mx = aa > bb ? aa : bb;
  if (mx < DBL_MIN) mx = DBL_MIN;
  return (ad / mx) <= (100.0 * DBL_EPSILON);
}
inline double log_gamma_windschitl(double x) {
  return 0.918938533204673 + (x - 0.5) * std::log(x) - x + 0.5 * x * std::log(x * std::sinh(1 / x) + 1 / (810.0 * std::pow(x, 6.0)));
}
inline double log_gamma_lanczos(double x) {
  static const double q[7] = {75122.6331530, 80916.6278952, 36308.2951477, 8687.24529705, 1168.92649479, 83.8676043424, 2.50662827511};
  double a = (x + 0.5) * std::log(x + 5.5) - (x + 5.5);
  double b = 0;
  for (int n = 0; n < 7; ++n) {
    a -= std::log(x + double(n));
    b += q[n] * std::pow(x, double(n));
  }
  return a + std::log(b);
}
inline double log_gamma(double x) { return x > 15.0 ? log_gamma_windschitl(x) : log_gamma_lanczos(x); }

struct Lsd {
  int w = 0, h = 0;
  double log_nt = 0;
  std::vector<uint8_t> img;       // scaled image
  std::vector<double> angles, modgrad;
  std::vector<uint8_t> used;
  std::vector<NormPoint> ordered;

  bool aligned(int x, int y, double theta, double prec) const {
    if (x < 0 || y < 0 || x >= w || y >= h) return false;
    double a = angles[(size_t)y * w + x];
    if (a == kNotDef) return false;
    double n = theta - a;
    if (n < 0) n = -n;
    if (n > k3_2PI) {
      n -= k2PI;
      if (n < 0) n = -n;
    }
    return n <= prec;
  }

  void ll_angle(double threshold, int n_bins, int tie_mode) {
    angles.assign((size_t)w * h, kNotDef);
    modgrad.assign((size_t)w * h, 0.0);
    double max_grad = -1;
    for (int y = 0; y < h - 1; ++y) {
      const uint8_t* r0 = &img[(size_t)y * w];
      const uint8_t* r1 = r0 + w;
      for (int x = 0; x < w - 1; ++x) {
        int DA = r1[x + 1] - r0[x], BC = r0[x + 1] - r1[x];
        int gx = DA + BC, gy = DA - BC;
        double norm = std::sqrt((gx * gx + gy * gy) / 4.0);
        modgrad[(size_t)y * w + x] = norm;
        if (norm <= threshold) angles[(size_t)y * w + x] = kNotDef;
        else {
          angles[(size_t)y * w + x] = fast_atan2((float)gx, (float)-gy) * kDegToRad;
          if (norm > max_grad) max_grad = norm;
        }
      }
    }
    double bin_coef = (max_grad > 0) ? double(n_bins - 1) / max_grad : 0;
    ordered.clear();
    ordered.reserve((size_t)(w - 1) * (h - 1));
    for (int y = 0; y < h - 1; ++y)
      for (int x = 0; x < w - 1; ++x) ordered.push_back(NormPoint{x, y, int(modgrad[(size_t)y * w + x] * bin_coef)});
    auto cmp = [](const NormPoint& a, const NormPoint& b) { return a.norm > b.norm; };
    if (tie_mode == 1) std::sort(ordered.begin(), ordered.end(), cmp);
    else std::stable_sort(ordered.begin(), ordered.end(), cmp);
  }

  void region_grow(int sx, int sy, std::vector<RegPoint>& reg, double& reg_angle, double prec) {
    reg.clear();
    reg_angle = angles[(size_t)sy * w + sx];
    reg.push_back(RegPoint{sx, sy, reg_angle, modgrad[(size_t)sy * w + sx]});
    float sumdx = float(sdpl_cos(reg_angle)), sumdy = float(sdpl_sin(reg_angle));
    used[(size_t)sy * w + sx] = 1;
    for (size_t i = 0; i < reg.size(); i++) {
      ++g_expanded;
      const int px = reg[i].x, py = reg[i].y;
      int xx_min = std::max(px - 1, 0), xx_max = std::min(px + 1, w - 1);
      int yy_min = std::max(py - 1, 0), yy_max = std::min(py + 1, h - 1);
      for (int yy = yy_min; yy <= yy_max; ++yy)
        for (int xx = xx_min; xx <= xx_max; ++xx) {
          uint8_t& u = used[(size_t)yy * w + xx];
          if (u != 1 && aligned(xx, yy, reg_angle, prec)) {
            const double angle = angles[(size_t)yy * w + xx];
            u = 1;
            reg.push_back(RegPoint{xx, yy, angle, modgrad[(size_t)yy * w + xx]});
            if (g_trig_mode == 0) {          // double cos of the float-rounded angle, rounded to float, float accumulate
              volatile float c = (float)sdpl_cos((double)(float)angle), s = (float)sdpl_sin((double)(float)angle);
              sumdx = sumdx + c; sumdy = sumdy + s;
            } else if (g_trig_mode == 1) {   // cosf / sinf, float accumulate
              volatile float c = cosf((float)angle), s = sinf((float)angle);
              sumdx = sumdx + c; sumdy = sumdy + s;
            } else {                         // double cos added in double, then rounded to the float accumulator
              sumdx = (float)((double)sumdx + std::cos((double)(float)angle));
              sumdy = (float)((double)sumdy + std::sin((double)(float)angle));
            }
            reg_angle = fast_atan2(sumdy, sumdx) * kDegToRad;
          }
        }
      if (g_trace_on == 2) g_trace_n.push_back((int)reg.size());
    }
    if (g_trace_on == 2) g_trace_n.push_back(-1);
  }

  double get_theta(const std::vector<RegPoint>& reg, double x, double y, double reg_angle, double prec) const {
    double Ixx = 0, Iyy = 0, Ixy = 0;
    for (size_t i = 0; i < reg.size(); ++i) {
      double dx = (double)reg[i].x - x, dy = (double)reg[i].y - y, wgt = reg[i].modgrad;
      Ixx += dy * dy * wgt;
      Iyy += dx * dx * wgt;
      Ixy -= dx * dy * wgt;
    }
    double lambda = 0.5 * (Ixx + Iyy - std::sqrt((Ixx - Iyy) * (Ixx - Iyy) + 4.0 * Ixy * Ixy));
    double theta = (std::fabs(Ixx) > std::fabs(Iyy)) ? double(fast_atan2(float(lambda - Ixx), float(Ixy)))
                                                     : double(fast_atan2(float(Ixy), float(lambda - Iyy)));
    theta *= kDegToRad;
    if (angle_diff(theta, reg_angle) > prec) theta += kPI;
    return theta;
  }

  void region2rect(const std::vector<RegPoint>& reg, double reg_angle, double prec, double p, Rect& rec) const {
    double x = 0, y = 0, sum = 0;
    for (size_t i = 0; i < reg.size(); ++i) {
      double wgt = reg[i].modgrad;
      x += double(reg[i].x) * wgt;
      y += double(reg[i].y) * wgt;
      sum += wgt;
    }
    x /= sum; y /= sum;
    double theta = get_theta(reg, x, y, reg_angle, prec);
    double dx = sdpl_cos(theta), dy = sdpl_sin(theta);
    double l_min = 0, l_max = 0, w_min = 0, w_max = 0;
    for (size_t i = 0; i < reg.size(); ++i) {
      double rdx = double(reg[i].x) - x, rdy = double(reg[i].y) - y;
      double l = rdx * dx + rdy * dy;
      double ww = -rdx * dy + rdy * dx;
      if (l > l_max) l_max = l; else if (l < l_min) l_min = l;
      if (ww > w_max) w_max = ww; else if (ww < w_min) w_min = ww;
    }
    rec.x1 = x + l_min * dx; rec.y1 = y + l_min * dy;
    rec.x2 = x + l_max * dx; rec.y2 = y + l_max * dy;
    rec.width = w_max - w_min;
    rec.x = x; rec.y = y; rec.theta = theta; rec.dx = dx; rec.dy = dy; rec.prec = prec; rec.p = p;
    if (rec.width < 1.0) rec.width = 1.0;
  }

  bool reduce_region_radius(std::vector<RegPoint>& reg, double reg_angle, double prec, double p, Rect& rec, double density,
                            double density_th) {
    double xc = double(reg[0].x), yc = double(reg[0].y);
    double r1 = dist_sq(xc, yc, rec.x1, rec.y1), r2 = dist_sq(xc, yc, rec.x2, rec.y2);
    double rad_sq = r1 > r2 ? r1 : r2;
    while (density < density_th) {
      rad_sq *= 0.75 * 0.75;
      for (size_t i = 0; i < reg.size(); ++i) {
        if (dist_sq(xc, yc, double(reg[i].x), double(reg[i].y)) > rad_sq) {
          used[(size_t)reg[i].y * w + reg[i].x] = 0;
          std::swap(reg[i], reg[reg.size() - 1]);
          reg.pop_back();
          --i;
        }
      }
      if (reg.size() < 2) return false;
      region2rect(reg, reg_angle, prec, p, rec);
      density = double(reg.size()) / (dist(rec.x1, rec.y1, rec.x2, rec.y2) * rec.width);
    }
    return true;
  }

  bool refine(std::vector<RegPoint>& reg, double reg_angle, double prec, double p, Rect& rec, double density_th) {
    double density = double(reg.size()) / (dist(rec.x1, rec.y1, rec.x2, rec.y2) * rec.width);
    if (density >= density_th) return true;
    double xc = double(reg[0].x), yc = double(reg[0].y);
    const double ang_c = reg[0].angle;
    double sum = 0, s_sum = 0;
    int n = 0;
    for (size_t i = 0; i < reg.size(); ++i) {
      used[(size_t)reg[i].y * w + reg[i].x] = 0;
      if (dist(xc, yc, reg[i].x, reg[i].y) < rec.width) {
        double d = angle_diff_signed(reg[i].angle, ang_c);
        sum += d;
        s_sum += d * d;
        ++n;
      }
    }
    double mean_angle = sum / double(n);
    double tau = 2.0 * std::sqrt((s_sum - 2.0 * mean_angle * sum) / double(n) + mean_angle * mean_angle);
    region_grow(reg[0].x, reg[0].y, reg, reg_angle, tau);
    if (reg.size() < 2) return false;
    region2rect(reg, reg_angle, prec, p, rec);
    density = double(reg.size()) / (dist(rec.x1, rec.y1, rec.x2, rec.y2) * rec.width);
    if (density < density_th) return reduce_region_radius(reg, reg_angle, prec, p, rec, density, density_th);
    return true;
  }

  double nfa(int n, int k, double p) const {
    if (n == 0 || k == 0) return -log_nt;
    if (n == k) return -log_nt - double(n) * std::log10(p);
    double p_term = p / (1 - p);
    double log1term = ((g_nfa_variant & 4) ? log_gamma(double(n) + 1) : (double(n) + 1)) - log_gamma(double(k) + 1) - log_gamma(double(n - k) + 1) + double(k) * std::log(p) +
                      double(n - k) * std::log(1.0 - p);
    double term = std::exp(log1term);
    if (double_equal(term, 0)) {
      if (k > n * p) return -log1term / kLn10 - log_nt;
      return -log_nt;
    }
    double bin_tail = term, tolerance = 0.1;
    for (int i = k + 1; i <= n; ++i) {
      double bin_term = double(n - i + 1) / double(i);
      double mult_term = bin_term * p_term;
      term *= mult_term;
      bin_tail += term;
      if (bin_term < 1) {
        double err = term * ((1 - std::pow(mult_term, double(n - i + 1))) / (1 - mult_term) - 1);
        if (err < tolerance * std::fabs(-std::log10(bin_tail) - log_nt) * bin_tail) break;
      }
    }
    return -std::log10(bin_tail) - log_nt;
  }

  static double dyhw0(const Rect& r) { return r.dy * (r.width / 2.0); }
  static double dxhw0(const Rect& r) { return r.dx * (r.width / 2.0); }
  double rect_nfa(const Rect& rec) const {
    int total_pts = 0, alg_pts = 0;
    if (g_nfa_variant & 16) {  // default
      // OpenCV >= 4.5.x rect_nfa: double vertices rotated to start at the min-y (then min-x) vertex, counter-clockwise;
      // rows ceil(top) .. ceil(bottom); columns ceil(left limit) .. int(right limit).
      double vx[4] = {rec.x1 - dyhw0(rec), rec.x2 - dyhw0(rec), rec.x2 + dyhw0(rec), rec.x1 + dyhw0(rec)};
      double vy[4] = {rec.y1 + dxhw0(rec), rec.y2 + dxhw0(rec), rec.y2 - dxhw0(rec), rec.y1 - dxhw0(rec)};
      int off = 0;
      for (int i = 1; i < 4; ++i)
        if (vy[i] < vy[off] || (vy[i] == vy[off] && vx[i] < vx[off])) off = i;
      double px[4], py[4];
      for (int i = 0; i < 4; ++i) { px[i] = vx[(i + off) % 4]; py[i] = vy[(i + off) % 4]; }
      auto slope = [&](int a, int b) { return ((int)std::ceil(py[b]) != (int)std::ceil(py[a])) ? (px[b] - px[a]) / (py[b] - py[a]) : 0.0; };
      const double flstep = slope(0, 1), slstep = slope(1, 2), frstep = slope(0, 3), srstep = slope(3, 2);
      const int y_end = (int)std::ceil(py[2]), c1 = (int)std::ceil(py[1]), c3 = (int)std::ceil(py[3]);
      for (int y = (int)std::ceil(py[0]); y <= y_end; ++y) {
        if (y < 0 || y >= h) continue;
        double left = (y <= c1) ? px[0] + (y - py[0]) * flstep : px[1] + (y - py[1]) * slstep;
        double right = (y < c3) ? px[0] + (y - py[0]) * frstep : px[3] + (y - py[3]) * srstep;
        // same pixel set as `for (x = ceil(left); x <= int(right); ++x) if (0 <= x < w)`, without walking the
        // out-of-image part of the span (near-horizontal edges give spans of 1e5+ columns)
        if (!(right >= 0) || !(left <= (double)(w - 1))) continue;
        const int xb = (int)std::ceil(left > 0 ? left : 0.0), xe = (int)(right < (double)(w - 1) ? right : (double)(w - 1));
        for (int x = xb; x <= xe; ++x) {
          ++total_pts;
          if (aligned(x, y, rec.theta, rec.prec)) ++alg_pts;
        }
      }
      g_last_n = total_pts; g_last_k = alg_pts;
      return nfa(total_pts, alg_pts, rec.p);
    }
    double half_width = rec.width / 2.0;
    double dyhw = rec.dy * half_width, dxhw = rec.dx * half_width;
    struct Edge { int x, y; bool taken; } e[4];
    e[0] = Edge{int(rec.x1 - dyhw), int(rec.y1 + dxhw), false};
    e[1] = Edge{int(rec.x2 - dyhw), int(rec.y2 + dxhw), false};
    e[2] = Edge{int(rec.x2 + dyhw), int(rec.y2 - dxhw), false};
    e[3] = Edge{int(rec.x1 + dyhw), int(rec.y1 - dxhw), false};
    std::sort(e, e + 4, [](const Edge& a, const Edge& b) { return a.x == b.x ? a.y < b.y : a.x < b.x; });
    Edge *min_y = &e[0], *max_y = &e[0];
    for (int i = 1; i < 4; ++i) {
      if (min_y->y > e[i].y) min_y = &e[i];
      if (max_y->y < e[i].y) max_y = &e[i];
    }
    min_y->taken = true;
    Edge* leftmost = nullptr;
    for (int i = 0; i < 4; ++i)
      if (!e[i].taken) { if (!leftmost) leftmost = &e[i]; else if (leftmost->x > e[i].x) leftmost = &e[i]; }
    leftmost->taken = true;
    Edge* rightmost = nullptr;
    for (int i = 0; i < 4; ++i)
      if (!e[i].taken) { if (!rightmost) rightmost = &e[i]; else if (rightmost->x < e[i].x) rightmost = &e[i]; }
    rightmost->taken = true;
    Edge* tailp = nullptr;
    for (int i = 0; i < 4; ++i)
      if (!e[i].taken) { if (!tailp) tailp = &e[i]; else if (tailp->x > e[i].x) tailp = &e[i]; }
    tailp->taken = true;
    // integer divisions and the p.x/p.y mix-up below are those of the OpenCV source
    double flstep, slstep, frstep, srstep;
    {
      const bool dd = g_nfa_variant & 1;   // double division instead of integer division
      const int ty = (g_nfa_variant & 2) ? tailp->y : tailp->x;
      auto dv = [&](int a, int b) { return dd ? (double)a / (double)b : (double)(a / b); };
      flstep = (min_y->y != leftmost->y) ? dv(min_y->x - leftmost->x, min_y->y - leftmost->y) : 0;
      slstep = (leftmost->y != ty) ? dv(leftmost->x - tailp->x, leftmost->y - ty) : 0;
      frstep = (min_y->y != rightmost->y) ? dv(min_y->x - rightmost->x, min_y->y - rightmost->y) : 0;
      srstep = (rightmost->y != ty) ? dv(rightmost->x - tailp->x, rightmost->y - ty) : 0;
    }
    double lstep = flstep, rstep = frstep;
    double left_x = min_y->x, right_x = min_y->x;
    int min_iter = min_y->y, max_iter = max_y->y;
    for (int y = min_iter; y <= max_iter; ++y) {
      if (y < 0 || y >= h) continue;
      for (int x = int(left_x); x <= int(right_x); ++x) {
        if (x < 0 || x >= w) continue;
        ++total_pts;
        if (aligned(x, y, rec.theta, rec.prec)) ++alg_pts;
      }
      if (y >= leftmost->y) lstep = slstep;
      if (y >= rightmost->y) rstep = srstep;
      left_x += lstep;
      right_x += rstep;
    }
    g_last_n = total_pts; g_last_k = alg_pts;
    return nfa(total_pts, alg_pts, rec.p);
  }

  double rect_improve(Rect& rec, double log_eps) const {
    const double delta = 0.5, delta_2 = delta / 2.0;
    double log_nfa = rect_nfa(rec);
    if (log_nfa > log_eps) return log_nfa;
    Rect r = rec;
    for (int n = 0; n < 5; ++n) {
      r.p /= 2; r.prec = r.p * kPI;
      double v = rect_nfa(r);
      if (v > log_nfa) { log_nfa = v; rec = r; }
    }
    if (log_nfa > log_eps) return log_nfa;
    r = rec;
    for (int n = 0; n < 5; ++n) {
      if ((r.width - delta) >= 0.5) {
        r.width -= delta;
        double v = rect_nfa(r);
        if (v > log_nfa) { rec = r; log_nfa = v; }
      }
    }
    if (log_nfa > log_eps) return log_nfa;
    r = rec;
    for (int n = 0; n < 5; ++n) {
      if ((r.width - delta) >= 0.5) {
        r.x1 += -r.dy * delta_2; r.y1 += r.dx * delta_2; r.x2 += -r.dy * delta_2; r.y2 += r.dx * delta_2;
        r.width -= delta;
        double v = rect_nfa(r);
        if (v > log_nfa) { rec = r; log_nfa = v; }
      }
    }
    if (log_nfa > log_eps) return log_nfa;
    r = rec;
    for (int n = 0; n < 5; ++n) {
      if ((r.width - delta) >= 0.5) {
        r.x1 -= -r.dy * delta_2; r.y1 -= r.dx * delta_2; r.x2 -= -r.dy * delta_2; r.y2 -= r.dx * delta_2;
        r.width -= delta;
        double v = rect_nfa(r);
        if (v > log_nfa) { rec = r; log_nfa = v; }
      }
    }
    if (log_nfa > log_eps) return log_nfa;
    r = rec;
    for (int n = 0; n < 5; ++n) {
      if ((r.width - delta) >= 0.5) {
        r.p /= 2; r.prec = r.p * kPI;
        double v = rect_nfa(r);
        if (v > log_nfa) { rec = r; log_nfa = v; }
      }
    }
    return log_nfa;
  }
};

}  // namespace

int lsd_detect(const uint8_t* src, int sw, int sh, int sstride, int refine_mode, double scale, double sigma_scale, double quant,
               double ang_th, double log_eps, double density_th, int n_bins, int tie_mode, std::vector<float>& lines) {
  lines.clear();
  g_dbg.clear();
  g_trace.clear();
  g_cand.clear();
  Lsd L;
  const double prec = kPI * ang_th / 180, p = ang_th / 180;
  const double rho = quant / std::sin(prec);
  if (scale != 1) {
    const double sigma = (scale < 1) ? (sigma_scale / scale) : sigma_scale;
    const unsigned int hk = (unsigned int)std::ceil(sigma * std::sqrt(2 * 3.0 * std::log(10.0)));
    if (!(hk == 3 && std::fabs(sigma - 0.75) < 1e-6)) return -1;  // only the reference's 0.6/0.8 kernel is restated
    std::vector<uint8_t> blurred((size_t)sw * sh);
    gaussian_blur_u8(src, sw, sh, sstride, blurred.data(), sw, 2);
    L.w = cv_round(sw * scale); L.h = cv_round(sh * scale);
    L.img.resize((size_t)L.w * L.h);
    resize_linear_exact_u8(blurred.data(), sw, sh, sw, L.img.data(), L.w, L.h, L.w, scale, scale);
  } else {
    L.w = sw; L.h = sh;
    L.img.resize((size_t)sw * sh);
    for (int y = 0; y < sh; y++) memcpy(&L.img[(size_t)y * sw], src + (size_t)y * sstride, sw);
  }
  g_last_scaled = L.img; g_last_w = L.w; g_last_h = L.h;
  L.ll_angle(rho, n_bins, tie_mode);
  L.log_nt = 5 * (std::log10(double(L.w)) + std::log10(double(L.h))) / 2 + std::log10(11.0);
  const size_t min_reg_size = size_t(-L.log_nt / std::log10(p));
  L.used.assign((size_t)L.w * L.h, 0);
  std::vector<RegPoint> reg;
  for (size_t i = 0; i < L.ordered.size(); ++i) {
    const int px = L.ordered[i].x, py = L.ordered[i].y;
    if (L.used[(size_t)py * L.w + px] == 0 && L.angles[(size_t)py * L.w + px] != kNotDef) {
      double reg_angle;
      g_expanded = 0;
      L.region_grow(px, py, reg, reg_angle, prec);
      const int n_first = (int)reg.size();
      struct Tr { int i, n1; ~Tr() { if (g_trace_on) { g_trace.push_back(i); g_trace.push_back(g_expanded); g_trace.push_back(n1); } } } tr{(int)i, n_first};
      if (reg.size() < min_reg_size) continue;
      Rect rec;
      L.region2rect(reg, reg_angle, prec, p, rec);
      double log_nfa = -1;
      if (refine_mode > 0) {
        if (!L.refine(reg, reg_angle, prec, p, rec, density_th)) continue;
        if (refine_mode >= 2) {
          const double v_first = L.rect_nfa(rec);
          const Rect rin = rec;
          log_nfa = L.rect_improve(rec, log_eps);
          if (g_trace_on) { g_cand.push_back(rin.x1); g_cand.push_back(rin.y1); g_cand.push_back(rin.x2); g_cand.push_back(rin.y2); g_cand.push_back(rin.width); g_cand.push_back(v_first); g_cand.push_back(log_nfa); g_cand.push_back((double)g_last_n); g_cand.push_back((double)g_last_k); }
          if (log_nfa <= log_eps) continue;
        }
      }
      Rect rec0 = rec;
      rec.x1 += 0.5; rec.y1 += 0.5; rec.x2 += 0.5; rec.y2 += 0.5;
      if (scale != 1) { rec.x1 /= scale; rec.y1 /= scale; rec.x2 /= scale; rec.y2 /= scale; rec.width /= scale; }
      { double v = L.rect_nfa(rec0); g_dbg.push_back(rec.width); g_dbg.push_back(rec.p); g_dbg.push_back(log_nfa); g_dbg.push_back(g_last_n); g_dbg.push_back(g_last_k); g_dbg.push_back(rec0.x1); g_dbg.push_back(rec0.y1); g_dbg.push_back(rec0.x2); g_dbg.push_back(rec0.y2); g_dbg.push_back(rec0.width); g_dbg.push_back(rec0.dx); g_dbg.push_back(rec0.dy); g_dbg.push_back(rec0.theta); (void)v; }
      lines.push_back(float(rec.x1)); lines.push_back(float(rec.y1)); lines.push_back(float(rec.x2)); lines.push_back(float(rec.y2));
    }
  }
  return (int)lines.size() / 4;
}

}  // namespace orc

extern "C" {
void orc_lsd_set_trig_mode(int m) { orc::g_trig_mode = m; }
void orc_lsd_set_nfa_variant(int m) { orc::g_nfa_variant = m; }
void orc_lsd_trace(int on) { orc::g_trace_on = on; orc::g_trace.clear(); orc::g_trace_n.clear(); orc::g_cand.clear(); }
int orc_lsd_trace_cand(double* out, int cap) { int n = (int)orc::g_cand.size(); for (int i = 0; i < n && i < cap; i++) out[i] = orc::g_cand[i]; return n; }
int orc_lsd_trace_n(int* out, int cap) { int n = (int)orc::g_trace_n.size(); for (int i = 0; i < n && i < cap; i++) out[i] = orc::g_trace_n[i]; return n; }
int orc_lsd_trace_get(int* out, int cap) { int n = (int)orc::g_trace.size(); for (int i = 0; i < n && i < cap; i++) out[i] = orc::g_trace[i]; return n; }
int orc_lsd_debug(double* out, int cap) { int n = (int)orc::g_dbg.size(); for (int i = 0; i < n && i < cap; i++) out[i] = orc::g_dbg[i]; return n; }
int orc_lsd_detect(const uint8_t* img, int w, int h, int stride, int refine, double scale, double sigma_scale, double quant,
                   double ang_th, double log_eps, double density_th, int n_bins, int tie_mode, float* lines, int cap) {
  std::vector<float> v;
  int n = orc::lsd_detect(img, w, h, stride, refine, scale, sigma_scale, quant, ang_th, log_eps, density_th, n_bins, tie_mode, v);
  if (n < 0) return n;
  for (int i = 0; i < n && i < cap; i++) memcpy(lines + 4 * i, &v[4 * i], 16);
  return n;
}
int orc_lsd_last_scaled(uint8_t* dst, int cap, int* w, int* h) {
  *w = orc::g_last_w; *h = orc::g_last_h;
  int n = orc::g_last_w * orc::g_last_h;
  if (dst && cap >= n) memcpy(dst, orc::g_last_scaled.data(), n);
  return n;
}
}
