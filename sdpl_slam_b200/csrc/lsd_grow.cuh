// lsd_grow.cuh -- device code of the LSD region stage: greedy region growing, region2rect, refine, and the
// rectangle NFA validation, i.e. the part of cv::LineSegmentDetector::detect (OpenCV imgproc/lsd.cpp, reached from
// 3rdparty/line_descriptor/src/LSDDetector_custom.cpp:291-309) that is an ORDER-DEPENDENT sequential algorithm on the CPU:
// seeds are visited by descending gradient bin, a pixel claimed by an earlier seed is unavailable to later ones, and
// `refine` un-marks and re-grows.
//
// B200 design ("speculate / validate / commit"):
//   * one warp per (frame, pyramid level) task; many tasks in flight per SM.
//   * a WAVE takes the next 32 still-free seeds in seed order; lane k runs the complete per-seed pipeline (grow ->
//     region2rect -> refine) SPECULATIVELY against the committed `used` map, without writing it.  Speculative ownership
//     is tracked in the same per-pixel state word with atomicMax of a stamp = (wave, 31-lane, phase): a lower lane
//     (= earlier seed) always wins a pixel, a lane that meets a pixel stamped by an earlier seed of its wave aborts.
//   * commit walks the wave in seed order: a lane whose seed is still free, that finished, whose examined-and-joined set E
//     is still entirely its own (no earlier seed of the wave took a pixel) and free in the committed map is provably
//     identical to what the sequential algorithm would have produced, so all such lanes below the first doubtful lane
//     commit in parallel (their sets are disjoint).  The first doubtful lane is re-run NON-speculatively (exact
//     sequential semantics against the committed map); the remaining lanes are then re-validated.
//   * rectangles that survive go to a pending list in seed order; their NFA validation (rect_improve: up to 26 scans of
//     the rotated rectangle) never touches `used`, so it runs afterwards with one warp per rectangle.
// The result is bit-identical to the sequential order of the oracle (oracle/lsd_oracle.cpp) for every decision that
// depends on `used`; double-precision transcendental calls (cos/sin/log/exp) are CUDA's, not glibc's (<= 2 ulp apart).
#pragma once
#include "common.cuh"

namespace sdpl {
namespace lsd {

constexpr double kPI = 3.1415926535897932384626433832795;
constexpr double kNotDef = -1024.0;
constexpr double kDegToRad = kPI / 180;
constexpr double k3_2PI = (3 * kPI) / 2;
constexpr double k2PI = 2 * kPI;
constexpr double kLn10 = 2.30258509299404568402;
constexpr uint32_t kUsed = 0x80000000u;

struct __align__(16) PxA { double ang; float c, s; };   // level-line angle (rad) or kNotDef; cos/sin of float(angle)

struct Rect { double x1, y1, x2, y2, width, x, y, theta, dx, dy, prec, p; };

struct __align__(8) Pending {      // a rectangle waiting for / after NFA validation, in seed order
  Rect rec;
  float seg[4];                    // x1,y1,x2,y2 in level pixels (after +0.5 and /scale), valid when accepted
  int accepted;
  int tag;                         // (wave << 8) | (lane << 1) | re-run flag  (introspection)
  int seed, npix;                  // seed pixel index and number of pixels finally marked (introspection)
};

struct Task {                      // one (frame, level)
  int w, h, npx;
  const PxA* px;
  const int* g2;                   // gx^2+gy^2 ; modgrad = sqrt(g2/4.0)
  uint32_t* state;                 // bit31 = used (committed) ; low bits = speculative stamp
  const uint32_t* order;           // defined pixels by descending bin, row-major inside a bin
  int ndef;
  int* reg_spec; int lane_cap;     // 32 lane segments of lane_cap ints
  int* reg_serial;                 // npx ints (non-speculative re-run)
  Pending* pend; int pend_cap; int* npend;
  double prec, p, log_nt, density_th, log_eps, scale /* lsd scale as a double, (double)0.8f */;
  int min_reg, refine;
  int* err;
};

__device__ __forceinline__ double dist_sq(double x1, double y1, double x2, double y2) { return (x2 - x1) * (x2 - x1) + (y2 - y1) * (y2 - y1); }
__device__ __forceinline__ double dist(double x1, double y1, double x2, double y2) { return sqrt(dist_sq(x1, y1, x2, y2)); }
__device__ __forceinline__ double angle_diff_signed(double a, double b) {
  double d = a - b;
  while (d <= -kPI) d += k2PI;
  while (d > kPI) d -= k2PI;
  return d;
}
__device__ __forceinline__ double angle_diff(double a, double b) { return fabs(angle_diff_signed(a, b)); }
__device__ __forceinline__ bool aligned_angle(double a, double theta, double prec) {
  if (a == kNotDef) return false;
  double n = theta - a;
  if (n < 0) n = -n;
  if (n > k3_2PI) {
    n -= k2PI;
    if (n < 0) n = -n;
  }
  return n <= prec;
}
__device__ __forceinline__ double modgrad_of(int g2) { return sqrt((double)g2 / 4.0); }
__device__ __forceinline__ uint32_t ld_state(const uint32_t* p) { return *(const volatile uint32_t*)p; }

// ------------------------------------------------------------------------------------------------
// region_grow.  SPEC=false: exact sequential semantics, marks `used` (state = kUsed).
//               SPEC=true : reads the committed map, claims pixels with atomicMax(stamp); returns false on abort
//                           (met a pixel stamped by an earlier seed of this wave, or ran out of list space).
// ------------------------------------------------------------------------------------------------
template <bool SPEC>
__device__ bool region_grow(const Task& T, int seed, int* reg, int cap, int& n_out, double& reg_angle, double prec, uint32_t stamp) {
  const int w = T.w, h = T.h;
  int n = 1;
  reg[0] = seed;
  double ra = T.px[seed].ang;
  float sumdx = (float)cos(ra), sumdy = (float)sin(ra);
  if (SPEC) {
    uint32_t old = atomicMax(&T.state[seed], stamp);
    if (old > stamp) { n_out = n; return false; }
  } else {
    T.state[seed] = kUsed;
  }
  for (int i = 0; i < n; i++) {
    const int p = reg[i];
    const int py = p / w, px = p - py * w;
    const int x0 = max(px - 1, 0), x1 = min(px + 1, w - 1), y0 = max(py - 1, 0), y1 = min(py + 1, h - 1);
    // issue all neighbour loads first (independent), then run the sequential tests
    uint32_t st[9];
    PxA pa[9];
#pragma unroll
    for (int k = 0; k < 9; k++) {
      const int yy = py - 1 + k / 3, xx = px - 1 + k % 3;
      const bool in = yy >= y0 && yy <= y1 && xx >= x0 && xx <= x1 && k != 4;
      const int q = yy * w + xx;
      st[k] = in ? ld_state(T.state + q) : kUsed;
      if (in) pa[k] = T.px[q]; else pa[k].ang = kNotDef;
    }
#pragma unroll
    for (int k = 0; k < 9; k++) {
      if (k == 4) continue;
      const uint32_t s = st[k];
      if (s & kUsed) continue;
      if (SPEC && s == stamp) continue;                     // already in my region (this phase)
      if (!aligned_angle(pa[k].ang, ra, prec)) continue;
      const int q = (py - 1 + k / 3) * w + (px - 1 + k % 3);
      if (SPEC) {
        uint32_t old = atomicMax(&T.state[q], stamp);
        if (old > stamp) { n_out = n; return false; }      // an earlier seed of this wave owns it (or it got committed)
        if (n >= cap) { n_out = n; return false; }
      } else {
        T.state[q] = kUsed;
      }
      reg[n++] = q;
      sumdx = __fadd_rn(sumdx, pa[k].c);
      sumdy = __fadd_rn(sumdy, pa[k].s);
      ra = (double)fast_atan2_deg(sumdy, sumdx) * kDegToRad;
    }
  }
  n_out = n;
  reg_angle = ra;
  return true;
}

__device__ double get_theta(const Task& T, const int* reg, int n, double x, double y, double reg_angle, double prec) {
  double Ixx = 0, Iyy = 0, Ixy = 0;
  for (int i = 0; i < n; ++i) {
    const int q = reg[i];
    const int qy = q / T.w, qx = q - qy * T.w;
    const double dx = (double)qx - x, dy = (double)qy - y, wgt = modgrad_of(T.g2[q]);
    Ixx += dy * dy * wgt;
    Iyy += dx * dx * wgt;
    Ixy -= dx * dy * wgt;
  }
  const double lambda = 0.5 * (Ixx + Iyy - sqrt((Ixx - Iyy) * (Ixx - Iyy) + 4.0 * Ixy * Ixy));
  double theta = (fabs(Ixx) > fabs(Iyy)) ? (double)fast_atan2_deg((float)(lambda - Ixx), (float)Ixy)
                                         : (double)fast_atan2_deg((float)Ixy, (float)(lambda - Iyy));
  theta *= kDegToRad;
  if (angle_diff(theta, reg_angle) > prec) theta += kPI;
  return theta;
}

__device__ void region2rect(const Task& T, const int* reg, int n, double reg_angle, double prec, double p, Rect& rec) {
  double x = 0, y = 0, sum = 0;
  for (int i = 0; i < n; ++i) {
    const int q = reg[i];
    const int qy = q / T.w, qx = q - qy * T.w;
    const double wgt = modgrad_of(T.g2[q]);
    x += (double)qx * wgt;
    y += (double)qy * wgt;
    sum += wgt;
  }
  x /= sum; y /= sum;
  const double theta = get_theta(T, reg, n, x, y, reg_angle, prec);
  const double dx = cos(theta), dy = sin(theta);
  double l_min = 0, l_max = 0, w_min = 0, w_max = 0;
  for (int i = 0; i < n; ++i) {
    const int q = reg[i];
    const int qy = q / T.w, qx = q - qy * T.w;
    const double rdx = (double)qx - x, rdy = (double)qy - y;
    const double l = rdx * dx + rdy * dy;
    const double ww = -rdx * dy + rdy * dx;
    if (l > l_max) l_max = l; else if (l < l_min) l_min = l;
    if (ww > w_max) w_max = ww; else if (ww < w_min) w_min = ww;
  }
  rec.x1 = x + l_min * dx; rec.y1 = y + l_min * dy;
  rec.x2 = x + l_max * dx; rec.y2 = y + l_max * dy;
  rec.width = w_max - w_min;
  rec.x = x; rec.y = y; rec.theta = theta; rec.dx = dx; rec.dy = dy; rec.prec = prec; rec.p = p;
  if (rec.width < 1.0) rec.width = 1.0;
}

// reduce_region_radius: swap-removal keeps every removed pixel in the array tail [n, n_orig) (needed as part of E)
template <bool SPEC>
__device__ bool reduce_region_radius(const Task& T, int* reg, int& n, double reg_angle, double prec, double p, Rect& rec, double density) {
  const int s0 = reg[0];
  const double xc = (double)(s0 % T.w), yc = (double)(s0 / T.w);
  const double r1 = dist_sq(xc, yc, rec.x1, rec.y1), r2 = dist_sq(xc, yc, rec.x2, rec.y2);
  double rad_sq = r1 > r2 ? r1 : r2;
  while (density < T.density_th) {
    rad_sq *= 0.75 * 0.75;
    for (int i = 0; i < n; ++i) {
      const int q = reg[i];
      if (dist_sq(xc, yc, (double)(q % T.w), (double)(q / T.w)) > rad_sq) {
        if (!SPEC) T.state[q] = 0;
        reg[i] = reg[n - 1];
        reg[n - 1] = q;
        --n;
        --i;
      }
    }
    if (n < 2) return false;
    region2rect(T, reg, n, reg_angle, prec, p, rec);
    density = (double)n / (dist(rec.x1, rec.y1, rec.x2, rec.y2) * rec.width);
  }
  return true;
}

struct SeedResult {
  int ok;          // speculation finished (always 1 for SPEC=false)
  int n1;          // first region size, list at reg[0, n1)
  int n2_orig;     // re-grown region slots at reg[n1, n1+n2_orig) (0 when refine did not re-grow)
  int nf;          // pixels finally marked: reg[foff, foff+nf)
  int foff;
  int has_rect;
  Rect rec;
};

// The per-seed pipeline of lsd_detect's main loop (oracle/lsd_oracle.cpp lsd_detect), up to the rectangle handed to
// rect_improve.  reg must hold cap ints.
template <bool SPEC>
__device__ void process_seed(const Task& T, int seed, int* reg, int cap, uint32_t stamp, SeedResult& R) {
  R.ok = 1; R.n2_orig = 0; R.has_rect = 0; R.foff = 0; R.nf = 0;
  double reg_angle = 0;
  int n1 = 0;
  if (!region_grow<SPEC>(T, seed, reg, cap, n1, reg_angle, T.prec, stamp)) { R.ok = 0; R.n1 = n1; return; }
  R.n1 = n1; R.nf = n1;
  if (n1 < T.min_reg) return;
  region2rect(T, reg, n1, reg_angle, T.prec, T.p, R.rec);
  if (T.refine <= 0) { R.has_rect = 1; return; }
  // ---- refine ----
  double density = (double)n1 / (dist(R.rec.x1, R.rec.y1, R.rec.x2, R.rec.y2) * R.rec.width);
  if (density >= T.density_th) { R.has_rect = 1; return; }
  const double xc = (double)(seed % T.w), yc = (double)(seed / T.w);
  const double ang_c = T.px[seed].ang;
  double sum = 0, s_sum = 0;
  int cnt = 0;
  for (int i = 0; i < n1; ++i) {
    const int q = reg[i];
    if (!SPEC) T.state[q] = 0;
    if (dist(xc, yc, (double)(q % T.w), (double)(q / T.w)) < R.rec.width) {
      const double d = angle_diff_signed(T.px[q].ang, ang_c);
      sum += d;
      s_sum += d * d;
      ++cnt;
    }
  }
  const double mean_angle = sum / (double)cnt;
  const double tau = 2.0 * sqrt((s_sum - 2.0 * mean_angle * sum) / (double)cnt + mean_angle * mean_angle);
  int* reg2 = SPEC ? reg + n1 : reg;           // the speculative run keeps the first region (it is part of E)
  const int cap2 = SPEC ? cap - n1 : cap;
  int n2 = 0;
  if (SPEC && cap2 < 1) { R.ok = 0; return; }
  if (!region_grow<SPEC>(T, seed, reg2, cap2, n2, reg_angle, tau, stamp | 1u)) { R.ok = 0; R.n2_orig = n2; return; }
  R.n2_orig = SPEC ? n2 : 0;
  R.foff = SPEC ? n1 : 0;
  R.nf = n2;
  if (n2 < 2) return;
  region2rect(T, reg2, n2, reg_angle, T.prec, T.p, R.rec);
  density = (double)n2 / (dist(R.rec.x1, R.rec.y1, R.rec.x2, R.rec.y2) * R.rec.width);
  if (density < T.density_th) {
    const bool keep = reduce_region_radius<SPEC>(T, reg2, n2, reg_angle, T.prec, T.p, R.rec, density);
    R.nf = n2;
    if (!keep) return;
  }
  R.has_rect = 1;
}

// ------------------------------------------------------------------------------------------------
// The grow kernel: one warp per task.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void append_rect(const Task& T, int idx, const Rect& rec, int tag, int seed, int npix) {
  if (idx < T.pend_cap) { T.pend[idx].rec = rec; T.pend[idx].accepted = 0; T.pend[idx].tag = tag; T.pend[idx].seed = seed; T.pend[idx].npix = npix; }
  else atomicExch(T.err, SDPL_ERR_OVERFLOW);
}

__device__ void grow_task(const Task& T, int serial_mode, int* sel /* 32 ints of shared memory */) {
  const int lane = threadIdx.x & 31;
  const uint32_t lt = (1u << lane) - 1u;
  int cursor = 0;
  int npend = 0;
  uint32_t wave = 0;
  int* my_reg = T.reg_spec + (size_t)lane * T.lane_cap;
  const int max_sel = serial_mode ? 1 : 32;
  while (true) {
    // ---- select the next (up to) 32 free seeds in order ----
    int nsel = 0;
    while (nsel < max_sel && cursor < T.ndef) {
      const int idx = cursor + lane;
      const int p = idx < T.ndef ? (int)T.order[idx] : -1;
      const bool fr = p >= 0 && !(ld_state(T.state + (p >= 0 ? p : 0)) & kUsed);
      const uint32_t m = __ballot_sync(0xffffffffu, fr);
      const int cnt = __popc(m);
      const int take = min(cnt, max_sel - nsel);
      const int rank = __popc(m & lt);
      if (fr && rank < take) sel[nsel + rank] = p;     // k-th selected seed goes to lane k
      if (cnt > take) cursor += (int)__fns(m, 0, take) + 1;   // stop right after the last taken one
      else cursor += 32;
      nsel += take;
    }
    __syncwarp();
    const int my_seed = lane < nsel ? sel[lane] : -1;
    __syncwarp();
    if (nsel == 0) break;
    wave++;
    SeedResult R;
    R.ok = 0; R.n1 = 0; R.n2_orig = 0; R.nf = 0; R.foff = 0; R.has_rect = 0;
    const uint32_t stamp = (wave << 7) | ((uint32_t)(31 - lane) << 1);   // bit0 = phase
    if (!serial_mode && lane < nsel) process_seed<true>(T, my_seed, my_reg, T.lane_cap, stamp, R);
    __syncwarp();
    uint32_t pending = nsel >= 32 ? 0xffffffffu : ((1u << nsel) - 1u);
    while (pending) {
      const bool mine = (pending >> lane) & 1u;
      bool dead = false, good = false;
      if (mine) {
        dead = (ld_state(T.state + my_seed) & kUsed) != 0;
        if (!dead && R.ok && !serial_mode) {
          // E = first region + re-grown region slots: every pixel must still carry my stamp (either phase) => no earlier
          // seed of the wave touched it and it is free in the committed map
          good = true;
          const int nE = R.n1 + R.n2_orig;
          const uint32_t key = stamp >> 1;
          for (int i = 0; i < nE; i++) {
            if ((ld_state(T.state + my_reg[i]) >> 1) != key) { good = false; break; }
          }
        }
      }
      const uint32_t deadm = __ballot_sync(0xffffffffu, dead), goodm = __ballot_sync(0xffffffffu, good);
      const uint32_t bad = pending & ~deadm & ~goodm;
      const int kstar = bad ? __ffs(bad) - 1 : 32;
      const uint32_t below = kstar >= 32 ? 0xffffffffu : ((1u << kstar) - 1u);
      const uint32_t commit = pending & goodm & below;
      const bool do_commit = (commit >> lane) & 1u;
      if (do_commit) {
        // un-claimed leftovers of E (first region pixels dropped by refine, radius-reduced pixels) stay free: their
        // stale stamps are harmless.  Mark the final set.
        const int* f = my_reg + R.foff;
        for (int i = 0; i < R.nf; i++) T.state[f[i]] = kUsed;
      }
      const uint32_t rectm = __ballot_sync(0xffffffffu, do_commit && R.has_rect);
      if (do_commit && R.has_rect) append_rect(T, npend + __popc(rectm & lt), R.rec, (int)((wave << 8) | (lane << 1) | 0), my_seed, R.nf);
      npend += __popc(rectm);
      __syncwarp();
      if (kstar < 32) {
        int has = 0;
        // the commits just made (lanes below kstar) may have taken kstar's seed: then the sequential algorithm skips it
        if (lane == kstar && !(ld_state(T.state + my_seed) & kUsed)) {
          SeedResult S;
          process_seed<false>(T, my_seed, T.reg_serial, T.npx, 0u, S);
          has = S.has_rect;
          if (has) append_rect(T, npend, S.rec, (int)((wave << 8) | (lane << 1) | 1), my_seed, S.nf);
        }
        has = __shfl_sync(0xffffffffu, has, kstar);
        npend += has;
        __syncwarp();
        pending &= ~below & ~(1u << kstar);
      } else {
        pending = 0;
      }
    }
  }
  if (lane == 0) *T.npend = min(npend, T.pend_cap);
}

// ------------------------------------------------------------------------------------------------
// NFA validation of one rectangle by one warp (rect_nfa / nfa / rect_improve of the oracle).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double log_gamma_windschitl(double x) {
  return 0.918938533204673 + (x - 0.5) * log(x) - x + 0.5 * x * log(x * sinh(1 / x) + 1 / (810.0 * pow(x, 6.0)));
}
__device__ __forceinline__ double log_gamma_lanczos(double x) {
  const double q[7] = {75122.6331530, 80916.6278952, 36308.2951477, 8687.24529705, 1168.92649479, 83.8676043424, 2.50662827511};
  double a = (x + 0.5) * log(x + 5.5) - (x + 5.5);
  double b = 0;
  for (int n = 0; n < 7; ++n) {
    a -= log(x + (double)n);
    b += q[n] * pow(x, (double)n);
  }
  return a + log(b);
}
__device__ __forceinline__ double log_gamma(double x) { return x > 15.0 ? log_gamma_windschitl(x) : log_gamma_lanczos(x); }
__device__ __forceinline__ bool double_equal(double a, double b) {
  if (a == b) return true;
  const double ad = fabs(a - b), aa = fabs(a), bb = fabs(b);
  double mx = aa > bb ? aa : bb;
  if (mx < 2.2250738585072014e-308) mx = 2.2250738585072014e-308;
  return (ad / mx) <= (100.0 * 2.220446049250313e-16);
}

__device__ double nfa(int n, int k, double p, double log_nt) {
  if (n == 0 || k == 0) return -log_nt;
  if (n == k) return -log_nt - (double)n * log10(p);
  const double p_term = p / (1 - p);
  const double log1term = log_gamma((double)n + 1) - log_gamma((double)k + 1) - log_gamma((double)(n - k) + 1) + (double)k * log(p) +
                          (double)(n - k) * log(1.0 - p);
  double term = exp(log1term);
  if (double_equal(term, 0)) {
    if (k > n * p) return -log1term / kLn10 - log_nt;
    return -log_nt;
  }
  double bin_tail = term;
  const double tolerance = 0.1;
  for (int i = k + 1; i <= n; ++i) {
    const double bin_term = (double)(n - i + 1) / (double)i;
    const double mult_term = bin_term * p_term;
    term *= mult_term;
    bin_tail += term;
    if (bin_term < 1) {
      const double err = term * ((1 - pow(mult_term, (double)(n - i + 1))) / (1 - mult_term) - 1);
      if (err < tolerance * fabs(-log10(bin_tail) - log_nt) * bin_tail) break;
    }
  }
  return -log10(bin_tail) - log_nt;
}

// counts the pixels of the rotated rectangle and those aligned with it; warp-cooperative, all lanes get the totals
__device__ double rect_nfa(const Task& T, const Rect& rec) {
  const int lane = threadIdx.x & 31;
  const double hw = rec.width / 2.0;
  const double dyhw = rec.dy * hw, dxhw = rec.dx * hw;
  double vx[4] = {rec.x1 - dyhw, rec.x2 - dyhw, rec.x2 + dyhw, rec.x1 + dyhw};
  double vy[4] = {rec.y1 + dxhw, rec.y2 + dxhw, rec.y2 - dxhw, rec.y1 - dxhw};
  int off = 0;
#pragma unroll
  for (int i = 1; i < 4; ++i)
    if (vy[i] < vy[off] || (vy[i] == vy[off] && vx[i] < vx[off])) off = i;
  double px[4], py[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) { px[i] = vx[(i + off) & 3]; py[i] = vy[(i + off) & 3]; }
  const int c0 = (int)ceil(py[0]), c1 = (int)ceil(py[1]), c2 = (int)ceil(py[2]), c3 = (int)ceil(py[3]);
  const double flstep = (c1 != c0) ? (px[1] - px[0]) / (py[1] - py[0]) : 0.0;
  const double slstep = (c2 != c1) ? (px[2] - px[1]) / (py[2] - py[1]) : 0.0;
  const double frstep = (c3 != c0) ? (px[3] - px[0]) / (py[3] - py[0]) : 0.0;
  const double srstep = (c2 != c3) ? (px[2] - px[3]) / (py[2] - py[3]) : 0.0;
  const int ya = max(c0, 0), yb = min(c2, T.h - 1);
  int total = 0, alg = 0;
  const int nrows = yb - ya + 1;
  const bool by_rows = nrows >= 24;
  for (int yy = ya + (by_rows ? lane : 0); yy <= yb; yy += (by_rows ? 32 : 1)) {
    const double left = (yy <= c1) ? px[0] + ((double)yy - py[0]) * flstep : px[1] + ((double)yy - py[1]) * slstep;
    const double right = (yy < c3) ? px[0] + ((double)yy - py[0]) * frstep : px[3] + ((double)yy - py[3]) * srstep;
    if (!(right >= 0) || !(left <= (double)(T.w - 1))) continue;
    const int xb = (int)ceil(left > 0 ? left : 0.0), xe = (int)(right < (double)(T.w - 1) ? right : (double)(T.w - 1));
    const PxA* row = T.px + (size_t)yy * T.w;
    for (int x = xb + (by_rows ? 0 : lane); x <= xe; x += (by_rows ? 1 : 32)) {
      ++total;
      if (aligned_angle(row[x].ang, rec.theta, rec.prec)) ++alg;
    }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) { total += __shfl_xor_sync(0xffffffffu, total, o); alg += __shfl_xor_sync(0xffffffffu, alg, o); }
  return nfa(total, alg, rec.p, T.log_nt);
}

__device__ double rect_improve(const Task& T, Rect& rec) {
  const double delta = 0.5, delta_2 = delta / 2.0;
  const double log_eps = T.log_eps;
  double log_nfa = rect_nfa(T, rec);
  if (log_nfa > log_eps) return log_nfa;
  Rect r = rec;
  for (int n = 0; n < 5; ++n) {
    r.p /= 2; r.prec = r.p * kPI;
    const double v = rect_nfa(T, r);
    if (v > log_nfa) { log_nfa = v; rec = r; }
  }
  if (log_nfa > log_eps) return log_nfa;
  r = rec;
  for (int n = 0; n < 5; ++n) {
    if ((r.width - delta) >= 0.5) {
      r.width -= delta;
      const double v = rect_nfa(T, r);
      if (v > log_nfa) { rec = r; log_nfa = v; }
    }
  }
  if (log_nfa > log_eps) return log_nfa;
  r = rec;
  for (int n = 0; n < 5; ++n) {
    if ((r.width - delta) >= 0.5) {
      r.x1 += -r.dy * delta_2; r.y1 += r.dx * delta_2; r.x2 += -r.dy * delta_2; r.y2 += r.dx * delta_2;
      r.width -= delta;
      const double v = rect_nfa(T, r);
      if (v > log_nfa) { rec = r; log_nfa = v; }
    }
  }
  if (log_nfa > log_eps) return log_nfa;
  r = rec;
  for (int n = 0; n < 5; ++n) {
    if ((r.width - delta) >= 0.5) {
      r.x1 -= -r.dy * delta_2; r.y1 -= r.dx * delta_2; r.x2 -= -r.dy * delta_2; r.y2 -= r.dx * delta_2;
      r.width -= delta;
      const double v = rect_nfa(T, r);
      if (v > log_nfa) { rec = r; log_nfa = v; }
    }
  }
  if (log_nfa > log_eps) return log_nfa;
  r = rec;
  for (int n = 0; n < 5; ++n) {
    if ((r.width - delta) >= 0.5) {
      r.p /= 2; r.prec = r.p * kPI;
      const double v = rect_nfa(T, r);
      if (v > log_nfa) { rec = r; log_nfa = v; }
    }
  }
  return log_nfa;
}

// one warp: validate pending rectangle i of task T and write its segment
__device__ void validate_pending(const Task& T, int i) {
  Rect rec = T.pend[i].rec;
  bool acc = true;
  if (T.refine >= 2) {
    const double v = rect_improve(T, rec);
    acc = v > T.log_eps;
  }
  if ((threadIdx.x & 31) == 0) {
    Pending& P = T.pend[i];
    P.accepted = acc ? 1 : 0;
    if (acc) {
      double x1 = rec.x1 + 0.5, y1 = rec.y1 + 0.5, x2 = rec.x2 + 0.5, y2 = rec.y2 + 0.5;
      if (T.scale != 1.0) { x1 /= T.scale; y1 /= T.scale; x2 /= T.scale; y2 /= T.scale; }
      P.seg[0] = (float)x1; P.seg[1] = (float)y1; P.seg[2] = (float)x2; P.seg[3] = (float)y2;
    }
  }
}

}  // namespace lsd
}  // namespace sdpl
