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
//     the rotated rectangle) never touches `used`, so it runs afterwards with one thread per rectangle.
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
// What the kernels keep per pixel, one 16-byte record = one load per neighbour test: the state word (bit31 = used, low bits =
// speculative stamp), the level-line angle as the FLOAT DEGREES fastAtan2 returned (the double radians the algorithm works with
// are (double)deg * kDegToRad, the very product ll_angle forms -- recomputed on load, bit-identical) or kNotDefDeg, and the
// cos / sin of float(angle).  Round 1 / 2 had a 16-byte {double, float, float} record plus a separate state array: two sectors
// per test instead of one.
struct __align__(16) PxRec { uint32_t state; float deg; float c, s; };
constexpr float kNotDefDeg = -1024.f;
__device__ __forceinline__ double rec_angle(float deg) { return deg == kNotDefDeg ? kNotDef : (double)deg * kDegToRad; }
struct PxArr {                     // T.px[q] -> PxA by value (angle rebuilt); T.px.cs(q) -> cos / sin
  const PxRec* p;
  __device__ __forceinline__ PxA operator[](int q) const {
    const PxRec* r = p + q;
    const float deg = r->deg;
    const float2 cs = *reinterpret_cast<const float2*>(&r->c);
    PxA a; a.ang = rec_angle(deg); a.c = cs.x; a.s = cs.y;
    return a;
  }
  __device__ __forceinline__ double ang(int q) const { return rec_angle(p[q].deg); }
  __device__ __forceinline__ float2 cs(int q) const { return *reinterpret_cast<const float2*>(&p[q].c); }
};
struct StateArr {                  // T.state + q / T.state[q]: the state word of record q
  PxRec* p;
  __device__ __forceinline__ uint32_t* operator+(int q) const { return &p[q].state; }
  __device__ __forceinline__ uint32_t& operator[](int q) const { return p[q].state; }
};

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
  PxArr px;                        // the per-pixel records (see PxRec): T.px[q] / T.state[q] are views of the same array
  const double* ang;               // the angles alone, densely packed: what the NFA scans read
  const int* g2;                   // gx^2+gy^2 ; modgrad = sqrt(g2/4.0)
  StateArr state;                  // bit31 = used (committed) ; low bits = speculative stamp
  const uint32_t* order;           // defined pixels by descending bin, row-major inside a bin
  int ndef;
  int* reg_spec; int lane_cap;     // 32 lane segments of lane_cap ints
  int* reg_serial;                 // npx ints (non-speculative re-run)
  Pending* pend; int pend_cap; int* npend;
  double prec, p, log_nt, density_th, log_eps, scale /* lsd scale as a double, (double)0.8f */;
  int min_reg, refine;
  int* err;
  const double* lgam; int lgam_n;  // log_gamma(i) for integer i < lgam_n (same formulas, tabulated once per handle)
  const double* nfa_tab;           // nfa(n, k, p / 2^j) of this octave for n <= kNfaTabN, j < kNfaTabLevels (or null): see nfa_lookup
  int prof_detail;                 // != 0: warp 0 also accumulates prof[8..15] (phase A / B split, growth, rectangle fit, refine cycles, steps)
  long long* prof;                 // optional [16]: [0..7] = cycles in select / speculate / evaluate+commit / re-run, waves, re-runs, dead, seeds
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
  // isAligned of the reference, branch-free: |theta - a| <= prec, or past 3/2 pi and |(|theta - a|) - 2 pi| <= prec -- the same
  // doubles as the if-chain.  An undefined angle (kNotDef = -1024) fails both tests for any tolerance below 1000 rad, so the
  // reference's explicit test for it is implied.
  const double t = fabs(theta - a);
  const double u = fabs(t - k2PI);
  return (t <= prec) | ((t > k3_2PI) & (u <= prec));
}
__device__ __forceinline__ double modgrad_of(int g2) { return sqrt((double)g2 / 4.0); }
// The state words of a task are only ever touched by the ONE CTA that grows that task, so CTA scope is all the coherence the
// speculative stamps need: relaxed.cta loads may be served by the SM's L1 (a volatile / .sys load goes to L2 every time, and
// the neighbourhood loads are the latency chain of the whole kernel), relaxed.cta RED.MAX claims a pixel.
__device__ __forceinline__ uint32_t ld_state(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.relaxed.cta.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// state word + pixel record in one 16-byte load / state word + angle in one 8-byte load (the state word with ld_state's semantics)
__device__ __forceinline__ void ld_rec(const PxRec* p, uint32_t& st, PxA& a) {
  uint32_t r0, r1, r2, r3;
  asm volatile("ld.relaxed.cta.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "l"(p) : "memory");
  st = r0; a.ang = rec_angle(__uint_as_float(r1)); a.c = __uint_as_float(r2); a.s = __uint_as_float(r3);
}
__device__ __forceinline__ void ld_rec_angle(const PxRec* p, uint32_t& st, double& ang) {
  uint32_t r0, r1;
  asm volatile("ld.relaxed.cta.global.v2.u32 {%0, %1}, [%2];" : "=r"(r0), "=r"(r1) : "l"(p) : "memory");
  st = r0; ang = rec_angle(__uint_as_float(r1));
}
__device__ __forceinline__ void claim_max(uint32_t* p, uint32_t v) {
  asm volatile("red.relaxed.cta.global.max.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}

// ------------------------------------------------------------------------------------------------
// region_grow.  SPEC=false: exact sequential semantics, marks `used` (state = kUsed).
//               SPEC=true : reads the committed map, claims pixels with atomicMax(stamp); returns false on abort
//                           (met a pixel stamped by an earlier seed of this wave, or ran out of list space).
// ------------------------------------------------------------------------------------------------
template <bool SPEC>
__device__ bool region_grow(const Task& T, int seed, int* reg, int cap, int& n_out, double& reg_angle, double prec, uint32_t stamp,
                            int& bx0, int& by0, int& bx1, int& by1) {
  const int w = T.w, h = T.h;
  int n = 1;
  reg[0] = seed;
  double ra = T.px.ang(seed);
  float sumdx = (float)sdpl_cos(ra), sumdy = (float)sdpl_sin(ra);
  if (SPEC) {
    if (ld_state(T.state + seed) > stamp) { n_out = n; return false; }
    claim_max(&T.state[seed], stamp);
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
        // s is either free, a stale stamp, or a later seed's stamp (all < mine): claim it.  An earlier seed's stamp (> mine)
        // means a conflict: stop here, the commit re-runs this seed.  The claim is a fire-and-forget RED.MAX (no round trip);
        // if an earlier seed claims the same pixel concurrently its stamp wins and the commit-time ownership check sees it.
        if (s > stamp) { n_out = n; return false; }
        if (n >= cap) { n_out = n; return false; }
        claim_max(&T.state[q], stamp);
      } else {
        T.state[q] = kUsed;
      }
      reg[n++] = q;
      bx0 = min(bx0, px - 1 + k % 3); bx1 = max(bx1, px - 1 + k % 3); by0 = min(by0, py - 1 + k / 3); by1 = max(by1, py - 1 + k / 3);
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
  const double dx = sdpl_cos(theta), dy = sdpl_sin(theta);
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
  int bx0, by0, bx1, by1;   // bounding box of every pixel the seed touched (speculative runs and the re-run)
  Rect rec;
};

// The per-seed pipeline of lsd_detect's main loop (oracle/lsd_oracle.cpp lsd_detect), up to the rectangle handed to
// rect_improve.  reg must hold cap ints.
template <bool SPEC>
__device__ void process_seed(const Task& T, int seed, int* reg, int cap, uint32_t stamp, SeedResult& R) {
  R.ok = 1; R.n2_orig = 0; R.has_rect = 0; R.foff = 0; R.nf = 0;
  R.bx0 = R.bx1 = seed % T.w; R.by0 = R.by1 = seed / T.w;
  double reg_angle = 0;
  int n1 = 0;
  if (!region_grow<SPEC>(T, seed, reg, cap, n1, reg_angle, T.prec, stamp, R.bx0, R.by0, R.bx1, R.by1)) { R.ok = 0; R.n1 = n1; return; }
  R.n1 = n1; R.nf = n1;
  if (n1 < T.min_reg) return;
  region2rect(T, reg, n1, reg_angle, T.prec, T.p, R.rec);
  if (T.refine <= 0) { R.has_rect = 1; return; }
  // ---- refine ----
  double density = (double)n1 / (dist(R.rec.x1, R.rec.y1, R.rec.x2, R.rec.y2) * R.rec.width);
  if (density >= T.density_th) { R.has_rect = 1; return; }
  const double xc = (double)(seed % T.w), yc = (double)(seed / T.w);
  const double ang_c = T.px.ang(seed);
  double sum = 0, s_sum = 0;
  int cnt = 0;
  for (int i = 0; i < n1; ++i) {
    const int q = reg[i];
    if (!SPEC) T.state[q] = 0;
    if (dist(xc, yc, (double)(q % T.w), (double)(q / T.w)) < R.rec.width) {
      const double d = angle_diff_signed(T.px.ang(q), ang_c);
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
  if (!region_grow<SPEC>(T, seed, reg2, cap2, n2, reg_angle, tau, stamp | 1u, R.bx0, R.by0, R.bx1, R.by1)) { R.ok = 0; R.n2_orig = n2; return; }
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


// Does every pixel of list[0,n) still carry the stamp key `key` (i.e. is it still exclusively mine and uncommitted)?
// Four independent list/state loads in flight per step.
__device__ __forceinline__ bool owns_all(const Task& T, const int* list, int n, uint32_t key) {
  int ok = 1, i = 0;
  for (; i + 4 <= n && ok; i += 4) {
    const int q0 = list[i], q1 = list[i + 1], q2 = list[i + 2], q3 = list[i + 3];
    const uint32_t s0 = ld_state(T.state + q0), s1 = ld_state(T.state + q1), s2 = ld_state(T.state + q2), s3 = ld_state(T.state + q3);
    ok = ((s0 >> 1) == key) & ((s1 >> 1) == key) & ((s2 >> 1) == key) & ((s3 >> 1) == key);
  }
  for (; i < n && ok; i++) ok = (ld_state(T.state + list[i]) >> 1) == key;
  return ok != 0;
}
__device__ __forceinline__ void mark_used(const Task& T, const int* list, int n) {
  int i = 0;
  for (; i + 4 <= n; i += 4) {
    const int q0 = list[i], q1 = list[i + 1], q2 = list[i + 2], q3 = list[i + 3];
    T.state[q0] = kUsed; T.state[q1] = kUsed; T.state[q2] = kUsed; T.state[q3] = kUsed;
  }
  for (; i < n; i++) T.state[list[i]] = kUsed;
}

// Warp-cooperative version of mark_used: lanes with `active` own a list; lists longer than 32 entries are marked by the
// whole warp (32 list loads in flight instead of one dependent chain), short ones by their owner.
__device__ __forceinline__ void mark_used_coop(const Task& T, bool active, const int* list, int n) {
  const int lane = threadIdx.x & 31;
  uint32_t longm = __ballot_sync(0xffffffffu, active && n > 32);
  while (longm) {
    const int src = __ffs(longm) - 1;
    longm &= longm - 1;
    const unsigned long long p = __shfl_sync(0xffffffffu, (unsigned long long)(size_t)list, src);
    const int m = __shfl_sync(0xffffffffu, n, src);
    const int* l = (const int*)(size_t)p;
    for (int i = lane; i < m; i += 32) T.state[l[i]] = kUsed;
  }
  if (active && n <= 32) mark_used(T, list, n);
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
  long long t_sel = 0, t_spec = 0, t_commit = 0, t_redo = 0, n_redo = 0, n_dead = 0, n_seed = 0;
  while (true) {
    long long c0 = clock64();
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
    n_seed += nsel;
    long long c1 = clock64(); t_sel += c1 - c0;
    SeedResult R;
    R.ok = 0; R.n1 = 0; R.n2_orig = 0; R.nf = 0; R.foff = 0; R.has_rect = 0;
    const uint32_t stamp = (wave << 7) | ((uint32_t)(31 - lane) << 1);   // bit0 = phase
    if (!serial_mode && lane < nsel) process_seed<true>(T, my_seed, my_reg, T.lane_cap, stamp, R);
    __syncwarp();
    long long c2 = clock64(); t_spec += c2 - c1;
    uint32_t pending = nsel >= 32 ? 0xffffffffu : ((1u << nsel) - 1u);
    while (pending) {
      const bool mine = (pending >> lane) & 1u;
      bool dead = false, good = false;
      if (mine) {
        dead = (ld_state(T.state + my_seed) & kUsed) != 0;
        if (!dead && R.ok && !serial_mode) {
          // E = first region + re-grown region slots: every pixel must still carry my stamp (either phase) => no earlier
          // seed of the wave touched it and it is free in the committed map
          good = owns_all(T, my_reg, R.n1 + R.n2_orig, stamp >> 1);
        }
      }
      const uint32_t deadm = __ballot_sync(0xffffffffu, dead), goodm = __ballot_sync(0xffffffffu, good);
      n_dead += __popc(pending & deadm);
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
        long long c3 = clock64();
        n_redo++;
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
        t_redo += clock64() - c3;
        pending &= ~below & ~(1u << kstar);
      } else {
        pending = 0;
      }
    }
    t_commit += clock64() - c2;
  }
  if (lane == 0) *T.npend = min(npend, T.pend_cap);
  if (lane == 0 && T.prof) {
    T.prof[0] = t_sel; T.prof[1] = t_spec; T.prof[2] = t_commit - t_redo; T.prof[3] = t_redo; T.prof[4] = wave; T.prof[5] = n_redo;
    T.prof[6] = n_dead; T.prof[7] = n_seed;
  }
}

// ------------------------------------------------------------------------------------------------
// Block-level waves: the same speculate / validate / commit scheme with NW warps per task, i.e. waves of K = 32*NW seeds.
// Thread t of the CTA owns the t-th seed of the wave (stamp priority K-1-t).  Fewer waves per task, and a long region only
// keeps its own warp busy while the other warps of the CTA idle at the barrier without consuming issue slots.
// ------------------------------------------------------------------------------------------------
constexpr int kMaxGrowWarps = 8;
struct BlockShared {
  int sel[32 * kMaxGrowWarps];
  uint32_t deadm[kMaxGrowWarps], goodm[kMaxGrowWarps], rectm[kMaxGrowWarps], pend[kMaxGrowWarps];
  int nsel, cursor, npend, has;
  int rb[4];                    // bounding box written by the last re-run
};


// The block kernel keeps region lists as packed coordinates (y << 16 | x) instead of linear indices: no integer division
// by the (run-time) image width anywhere in the per-pixel loops.
__device__ __forceinline__ int xy_pack(int x, int y) { return (y << 16) | x; }
__device__ __forceinline__ int xy_x(int v) { return v & 0xffff; }
__device__ __forceinline__ int xy_y(int v) { return (int)((unsigned)v >> 16); }
__device__ __forceinline__ int xy_lin(int v, int w) { return xy_y(v) * w + xy_x(v); }

__device__ __forceinline__ bool owns_all_xy(const Task& T, const int* list, int n, uint32_t key) {
  // eight list entries, then their eight state words, in flight per trip: the walk is two dependent round trips per trip
  const int w = T.w;
  int ok = 1;
  for (int i = 0; i < n && ok; i += 8) {
    int q[8];
    uint32_t sv[8];
#pragma unroll
    for (int j = 0; j < 8; j++) q[j] = i + j < n ? xy_lin(list[i + j], w) : -1;
#pragma unroll
    for (int j = 0; j < 8; j++) sv[j] = q[j] >= 0 ? ld_state(T.state + q[j]) : (key << 1);
#pragma unroll
    for (int j = 0; j < 8; j++) ok &= (sv[j] >> 1) == key;
  }
  return ok != 0;
}
// warp-cooperative ownership check: lists longer than 64 entries are walked by the whole warp (128 loads in flight per
// step instead of one lane's dependent chain), short ones by their owner.  `active` lanes get their verdict back.
__device__ __forceinline__ bool owns_all_coop_xy(const Task& T, bool active, const int* list, int n, uint32_t key) {
  const int lane = threadIdx.x & 31, w = T.w;
  bool res = false;
  uint32_t longm = __ballot_sync(0xffffffffu, active && n > 64);
  while (longm) {
    const int src = __ffs(longm) - 1;
    longm &= longm - 1;
    const unsigned long long p = __shfl_sync(0xffffffffu, (unsigned long long)(size_t)list, src);
    const int m = __shfl_sync(0xffffffffu, n, src);
    const uint32_t k = __shfl_sync(0xffffffffu, key, src);
    const int* l = (const int*)(size_t)p;
    int ok = 1;
    for (int i = lane * 4; i < m; i += 128) {
      const int c = min(4, m - i);
      int q[4]; uint32_t sv[4];
#pragma unroll
      for (int j = 0; j < 4; j++) q[j] = j < c ? xy_lin(l[i + j], w) : -1;
#pragma unroll
      for (int j = 0; j < 4; j++) sv[j] = j < c ? ld_state(T.state + q[j]) : (k << 1);
#pragma unroll
      for (int j = 0; j < 4; j++) ok &= (sv[j] >> 1) == k;
    }
    const bool all_ok = __all_sync(0xffffffffu, ok);
    if (lane == src) res = all_ok;
  }
  if (active && n <= 64) res = owns_all_xy(T, list, n, key);
  return res;
}
__device__ __forceinline__ void mark_used_coop_xy(const Task& T, bool active, const int* list, int n) {
  const int lane = threadIdx.x & 31, w = T.w;
  uint32_t longm = __ballot_sync(0xffffffffu, active && n > 32);
  while (longm) {
    const int src = __ffs(longm) - 1;
    longm &= longm - 1;
    const unsigned long long p = __shfl_sync(0xffffffffu, (unsigned long long)(size_t)list, src);
    const int m = __shfl_sync(0xffffffffu, n, src);
    const int* l = (const int*)(size_t)p;
    // four list blocks in flight per trip (a plain loop waits for every load before it issues the next one)
    for (int i = lane; i < m; i += 128) {
      int q[4];
#pragma unroll
      for (int j = 0; j < 4; j++) q[j] = i + 32 * j < m ? l[i + 32 * j] : -1;
#pragma unroll
      for (int j = 0; j < 4; j++) if (q[j] >= 0) T.state[xy_lin(q[j], w)] = kUsed;
    }
  }
  if (active && n <= 32) {
    for (int i = 0; i < n; i += 8) {
      int q[8];
#pragma unroll
      for (int j = 0; j < 8; j++) q[j] = i + j < n ? list[i + j] : -1;
#pragma unroll
      for (int j = 0; j < 8; j++) if (q[j] >= 0) T.state[xy_lin(q[j], w)] = kUsed;
    }
  }
}

// Warp-cooperative version of the SEQUENTIAL per-seed pipeline (exactly the semantics and arithmetic order of
// process_seed<false>): the whole warp works on ONE seed.  Used for the re-runs, which otherwise leave 127 of 128 threads
// idle behind one dependent-instruction chain.  All scalar state is warp-uniform (every lane holds the same value); lanes
// differ only in the element they fetch:
//   * growth: the 8 neighbours of the expanded pixel are loaded and tested by 8 lanes at once; joins are resolved in scan
//     order with ballots (after a join only the later neighbours are re-tested against the updated region angle);
//   * rectangle fit / refine statistics: 32 list elements per step are loaded and transformed (sqrt, products) in parallel
//     and then accumulated by every lane in list order through shuffles, so the floating-point sums are bit-identical to
//     the sequential loop.
__device__ __forceinline__ double shfl_d(double v, int src) {
  int lo = __double2loint(v), hi = __double2hiint(v);
  lo = __shfl_sync(0xffffffffu, lo, src); hi = __shfl_sync(0xffffffffu, hi, src);
  return __hiloint2double(hi, lo);
}

// region2rect (+ get_theta) by a whole warp: list / n / ra are warp-uniform, the result is returned in every lane.
// 32 elements per step are fetched and transformed in parallel, then accumulated by every lane in list order through
// shuffles, so the double sums are bit-identical to the sequential loops of the reference.
__device__ __forceinline__ void coop_region2rect(const Task& T, const int* list, int n, double ra, Rect& rec) {
  const int lane = threadIdx.x & 31, w = T.w;
  double x = 0, y = 0, sum = 0;
  for (int b = 0; b < n; b += 32) {
    const int e = b + lane;
    double wgt = 0, qxw = 0, qyw = 0;
    if (e < n) {
      const int q = list[e];
      const int qy = xy_y(q), qx = xy_x(q);
      wgt = modgrad_of(T.g2[qy * w + qx]);
      qxw = (double)qx * wgt; qyw = (double)qy * wgt;
    }
    const int m = min(32, n - b);
    for (int j = 0; j < m; j++) { x += shfl_d(qxw, j); y += shfl_d(qyw, j); sum += shfl_d(wgt, j); }
  }
  x /= sum; y /= sum;
  double Ixx = 0, Iyy = 0, Ixy = 0;
  for (int b = 0; b < n; b += 32) {
    const int e = b + lane;
    double t0 = 0, t1 = 0, t2 = 0;
    if (e < n) {
      const int q = list[e];
      const int qy = xy_y(q), qx = xy_x(q);
      const double dx = (double)qx - x, dy = (double)qy - y, wgt = modgrad_of(T.g2[qy * w + qx]);
      t0 = dy * dy * wgt; t1 = dx * dx * wgt; t2 = dx * dy * wgt;
    }
    const int m = min(32, n - b);
    for (int j = 0; j < m; j++) { Ixx += shfl_d(t0, j); Iyy += shfl_d(t1, j); Ixy -= shfl_d(t2, j); }
  }
  const double lambda = 0.5 * (Ixx + Iyy - sqrt((Ixx - Iyy) * (Ixx - Iyy) + 4.0 * Ixy * Ixy));
  double theta = (fabs(Ixx) > fabs(Iyy)) ? (double)fast_atan2_deg((float)(lambda - Ixx), (float)Ixy)
                                         : (double)fast_atan2_deg((float)Ixy, (float)(lambda - Iyy));
  theta *= kDegToRad;
  if (angle_diff(theta, ra) > T.prec) theta += kPI;
  double dy_, dx_;
  sdpl_sincos(theta, &dy_, &dx_);
  double l_min = 0, l_max = 0, w_min = 0, w_max = 0;      // min / max are order-independent: plain warp reductions
  for (int e = lane; e < n; e += 32) {
    const int q = list[e];
    const int qy = xy_y(q), qx = xy_x(q);
    const double rdx = (double)qx - x, rdy = (double)qy - y;
    const double l = rdx * dx_ + rdy * dy_;
    const double ww = -rdx * dy_ + rdy * dx_;
    l_max = fmax(l_max, l); l_min = fmin(l_min, l);
    w_max = fmax(w_max, ww); w_min = fmin(w_min, ww);
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    l_max = fmax(l_max, shfl_d(l_max, lane ^ o)); l_min = fmin(l_min, shfl_d(l_min, lane ^ o));
    w_max = fmax(w_max, shfl_d(w_max, lane ^ o)); w_min = fmin(w_min, shfl_d(w_min, lane ^ o));
  }
  rec.x1 = x + l_min * dx_; rec.y1 = y + l_min * dy_;
  rec.x2 = x + l_max * dx_; rec.y2 = y + l_max * dy_;
  rec.width = w_max - w_min;
  rec.x = x; rec.y = y; rec.theta = theta; rec.dx = dx_; rec.dy = dy_; rec.prec = T.prec; rec.p = T.p;
  if (rec.width < 1.0) rec.width = 1.0;
}

// refine: the re-growth tolerance from the angle spread of the region pixels within `width` of the seed (warp-cooperative,
// sums in list order); UNMARK = sequential semantics (the region's pixels are released before the re-growth)
template <bool UNMARK>
__device__ __forceinline__ double coop_refine_tau(const Task& T, const int* list, int n, int sx, int sy, double seed_ang, double width) {
  const int lane = threadIdx.x & 31, w = T.w;
  double sm = 0, s_sum = 0;
  int cnt = 0;
  for (int b = 0; b < n; b += 32) {
    const int e = b + lane;
    double d = 0; int use = 0;
    if (e < n) {
      const int qp = list[e];
      const int q = xy_lin(qp, w);
      if (UNMARK) T.state[q] = 0;
      if (dist((double)sx, (double)sy, (double)xy_x(qp), (double)xy_y(qp)) < width) { d = angle_diff_signed(T.px.ang(q), seed_ang); use = 1; }
    }
    const uint32_t um = __ballot_sync(0xffffffffu, use);
    const int m = min(32, n - b);
    for (int j = 0; j < m; j++) {
      if ((um >> j) & 1u) { const double dj = shfl_d(d, j); sm += dj; s_sum += dj * dj; ++cnt; }
    }
  }
  const double mean_angle = sm / (double)cnt;
  return 2.0 * sqrt((s_sum - 2.0 * mean_angle * sm) / (double)cnt + mean_angle * mean_angle);
}

__device__ void process_seed_coop(const Task& T, const int seed, int* const reg, SeedResult& R) {
  const int lane = threadIdx.x & 31;
  const int w = T.w, h = T.h;
  R.ok = 1; R.n1 = 0; R.n2_orig = 0; R.has_rect = 0; R.foff = 0; R.nf = 0;
  const int sx = seed % w, sy = seed / w;
  int bx0 = sx, bx1 = sx, by0 = sy, by1 = sy;
  int state = 0, n = 0;
  double prec = T.prec, ra = 0, rad_sq = 0;
  const double seed_ang = T.px.ang(seed);
  // neighbour handled by this lane (lanes 0..8 except 4)
  const int ndx = (lane % 3) - 1, ndy = (lane / 3) - 1;
  const bool nlane = lane < 9 && lane != 4;
#pragma unroll 1
  while (true) {
    if (state <= 1) {
      // ---------------- region_grow ----------------
      if (lane == 0) { T.state[seed] = kUsed; reg[0] = xy_pack(sx, sy); }
      n = 1;
      ra = seed_ang;
      double sn, cs;
      sdpl_sincos(ra, &sn, &cs);
      float sumdx = (float)cs, sumdy = (float)sn;
      int nxt = xy_pack(sx, sy);
      __syncwarp();
#pragma unroll 1
      for (int i = 0; i < n; i++) {
        const int p = nxt;
        const int n_start = n;
        const int py = xy_y(p), px = xy_x(p);
        const int yy = py + ndy, xx = px + ndx;
        const bool in = nlane && yy >= 0 && yy < h && xx >= 0 && xx < w;
        const int q = yy * w + xx;
        uint32_t st = kUsed;
        PxA pa; pa.ang = kNotDef; pa.c = 0.f; pa.s = 0.f;
        if (in) { st = ld_state(T.state + q); pa = T.px[q]; }
        if (i + 1 < n_start) nxt = reg[i + 1];
        const bool cand = in && !(st & kUsed) && pa.ang != kNotDef;
        uint32_t remaining = __ballot_sync(0xffffffffu, cand);
        while (remaining) {
          const bool al = ((remaining >> lane) & 1u) && aligned_angle(pa.ang, ra, prec);
          const uint32_t m = __ballot_sync(0xffffffffu, al);
          if (!m) break;
          const int k = __ffs(m) - 1;                     // first aligned neighbour in scan order
          const int qk = __shfl_sync(0xffffffffu, q, k);
          const int qx = __shfl_sync(0xffffffffu, xx, k), qy = __shfl_sync(0xffffffffu, yy, k);
          const float ck = __shfl_sync(0xffffffffu, pa.c, k), sk = __shfl_sync(0xffffffffu, pa.s, k);
          if (lane == 0) { T.state[qk] = kUsed; reg[n] = xy_pack(qx, qy); }
          if (n == i + 1) nxt = xy_pack(qx, qy);
          n++;
          bx0 = min(bx0, qx); bx1 = max(bx1, qx); by0 = min(by0, qy); by1 = max(by1, qy);
          sumdx = __fadd_rn(sumdx, ck);
          sumdy = __fadd_rn(sumdy, sk);
          ra = (double)fast_atan2_deg(sumdy, sumdx) * kDegToRad;
          remaining &= ~((2u << k) - 1u);                 // earlier neighbours were already decided
        }
        __syncwarp();                                      // lane 0's list / state writes before the next loads
      }
      if (state == 0) { R.n1 = n; R.nf = n; } else { R.nf = n; }
      if (state == 0 ? (n < T.min_reg) : (n < 2)) break;
    } else {
      // ---------------- one pass of reduce_region_radius (rare; order-dependent swap removal, done uniformly) ----------------
      rad_sq *= 0.75 * 0.75;
      for (int i = 0; i < n; ++i) {
        const int q = reg[i];
        if (dist_sq((double)sx, (double)sy, (double)xy_x(q), (double)xy_y(q)) > rad_sq) {
          const int last = reg[n - 1];
          __syncwarp();
          if (lane == 0) { T.state[xy_lin(q, w)] = 0; reg[i] = last; reg[n - 1] = q; }
          __syncwarp();
          --n;
          --i;
        }
      }
      R.nf = n;
      if (n < 2) break;
    }
    coop_region2rect(T, reg, n, ra, R.rec);
    const double density = (double)n / (dist(R.rec.x1, R.rec.y1, R.rec.x2, R.rec.y2) * R.rec.width);
    if (state == 0) {
      if (T.refine <= 0 || density >= T.density_th) { R.has_rect = 1; break; }
      prec = coop_refine_tau<true>(T, reg, n, sx, sy, seed_ang, R.rec.width);
      __syncwarp();
      state = 1;
      continue;
    }
    if (density >= T.density_th) { R.has_rect = 1; break; }
    if (state == 1) {
      const double r1 = dist_sq((double)sx, (double)sy, R.rec.x1, R.rec.y1), r2 = dist_sq((double)sx, (double)sy, R.rec.x2, R.rec.y2);
      rad_sq = r1 > r2 ? r1 : r2;
      state = 2;
    }
  }
  R.bx0 = bx0; R.bx1 = bx1; R.by0 = by0; R.by1 = by1;
  __syncwarp();
}

// The speculative pass of one wave, warp-synchronous: all 32 lanes of a warp call this together, `active` lanes own a seed.
// Region growth runs per lane (divergent, each lane its own region); everything that is a loop over a finished region --
// the rectangle fit and the refine statistics -- is done by the WHOLE warp for one lane's region at a time
// (coop_region2rect / coop_refine_tau): ncu showed a third of the kernel's active time in those loops with exactly one
// thread per warp running.  Arithmetic and its order are those of process_seed<true>.
__device__ void speculate_wave(const Task& T, const bool active, const int seed, int* const reg, const int cap, const uint32_t stamp0,
                               SeedResult& R) {
  const int lane = threadIdx.x & 31;
  const int w = T.w, h = T.h;
  enum { P_GROW = 0, P_RECT, P_REFSTAT, P_REDUCE, P_DONE };
  R.ok = 1; R.n1 = 0; R.n2_orig = 0; R.has_rect = 0; R.foff = 0; R.nf = 0;
  const int sx = active ? seed % w : 0, sy = active ? seed / w : 0;
  R.bx0 = R.bx1 = sx; R.by0 = R.by1 = sy;
  int phase = active ? P_GROW : P_DONE;
  int state = 0;                 // 0: first growth, 1: re-growth with the refined tolerance, 2: radius reduction passes
  int* cur = reg;
  int n = 0, capc = cap;
  double prec = T.prec, ra = 0, rad_sq = 0;
  uint32_t stamp = stamp0;
  const double seed_ang = active ? T.px.ang(seed) : 0.0;
#pragma unroll 1
  while (__any_sync(0xffffffffu, phase != P_DONE)) {
    if (phase == P_GROW) {
      // ---------------- region_grow (per lane) ----------------
      bool aborted = false;
      if (capc < 1 || ld_state(T.state + seed) > stamp) aborted = true;
      else {
        claim_max(&T.state[seed], stamp);
        cur[0] = xy_pack(sx, sy); n = 1;
        ra = seed_ang;
        double sn, cs;
        sdpl_sincos(ra, &sn, &cs);
        float sumdx = (float)cs, sumdy = (float)sn;
        int nxt = xy_pack(sx, sy);
#pragma unroll 1
        for (int i = 0; i < n && !aborted; i++) {
          const int p = nxt;
          const int n_start = n;
          const int py = xy_y(p), px = xy_x(p);
          uint32_t st[9];
          const int rofs[3] = {max(py - 1, 0) * w, py * w, min(py + 1, h - 1) * w};
          const int cofs[3] = {max(px - 1, 0), px, min(px + 1, w - 1)};
          uint32_t vmask = 0x1EFu;                                  // bits 0..8 without the centre
          if (px == 0) vmask &= ~0x049u;
          if (px == w - 1) vmask &= ~0x124u;
          if (py == 0) vmask &= ~0x007u;
          if (py == h - 1) vmask &= ~0x1C0u;
          // only the angle of the eight neighbours is fetched up front; cos / sin are read for the one that joins
          double ang[9];
#pragma unroll
          for (int k = 0; k < 9; k++) {
            if (k == 4) continue;
            const int q = rofs[k / 3] + cofs[k % 3];
            st[k] = ld_state(T.state + q);
            ang[k] = T.px.ang(q);            // (not the dense T.ang: this load also brings the cos/sin of the neighbour into L1)
          }
          if (i + 1 < n_start) nxt = cur[i + 1];
          // candidates: free, not mine yet, gradient defined; `foreign` = carries the stamp of an earlier seed of the wave
          uint32_t cand = 0, foreign = 0;
#pragma unroll
          for (int k = 0; k < 9; k++) {
            if (k == 4) continue;
            const uint32_t sv = st[k];
            if (!(sv & kUsed) && sv != stamp && ang[k] != kNotDef) cand |= 1u << k;
            if (sv > stamp) foreign |= 1u << k;
          }
          cand &= vmask;
          while (cand) {
            // alignment of all eight slots against the running angle, branch-free (aligned_angle without the undefined test:
            // cand excludes those pixels; |t| <= prec, or past 3/2 pi and 2 pi - |t| <= prec -- the same doubles)
            uint32_t al = 0;
#pragma unroll
            for (int k = 0; k < 9; k++) {
              if (k == 4) continue;
              const double t = fabs(ra - ang[k]);
              const double u = fabs(t - k2PI);
              const bool ok = (t <= prec) | ((t > k3_2PI) & (u <= prec));
              al |= (ok ? 1u : 0u) << k;
            }
            al &= cand;
            if (!al) break;
            const int k = __ffs(al) - 1;
            const int qx = px - 1 + k % 3, qy = py - 1 + k / 3;
            if (((foreign >> k) & 1u) || n >= capc) { aborted = true; break; }
            const int q = qy * w + qx;
            claim_max(&T.state[q], stamp);
            const float2 cs = T.px.cs(q);
            const int qp = xy_pack(qx, qy);
            if (n == i + 1) nxt = qp;
            cur[n++] = qp;
            R.bx0 = min(R.bx0, qx); R.bx1 = max(R.bx1, qx);
            R.by0 = min(R.by0, qy); R.by1 = max(R.by1, qy);
            sumdx = __fadd_rn(sumdx, cs.x);
            sumdy = __fadd_rn(sumdy, cs.y);
            ra = (double)fast_atan2_deg(sumdy, sumdx) * kDegToRad;
            cand &= ~((2u << k) - 1u);
          }
        }
      }
      if (state == 0) { R.n1 = n; R.nf = n; } else { R.n2_orig = n; R.nf = n; }
      if (aborted) { R.ok = 0; phase = P_DONE; }
      else if (state == 0 ? (n < T.min_reg) : (n < 2)) phase = P_DONE;
      else phase = P_RECT;
    } else if (phase == P_REDUCE) {
      // ---------------- one pass of reduce_region_radius (per lane; order-dependent swap removal) ----------------
      rad_sq *= 0.75 * 0.75;
      for (int i = 0; i < n; ++i) {
        const int q = cur[i];
        if (dist_sq((double)sx, (double)sy, (double)xy_x(q), (double)xy_y(q)) > rad_sq) {
          cur[i] = cur[n - 1];
          cur[n - 1] = q;
          --n;
          --i;
        }
      }
      R.nf = n;
      phase = n < 2 ? P_DONE : P_RECT;
    }
    __syncwarp();
    // ---------------- rectangle fits, one region at a time by the whole warp ----------------
    uint32_t m = __ballot_sync(0xffffffffu, phase == P_RECT);
    while (m) {
      const int L = __ffs(m) - 1;
      m &= m - 1;
      const int* lp = (const int*)(size_t)__shfl_sync(0xffffffffu, (unsigned long long)(size_t)cur, L);
      const int ln = __shfl_sync(0xffffffffu, n, L);
      const double lra = shfl_d(ra, L);
      Rect rc;
      coop_region2rect(T, lp, ln, lra, rc);
      if (lane == L) R.rec = rc;
    }
    if (phase == P_RECT) {
      const double density = (double)n / (dist(R.rec.x1, R.rec.y1, R.rec.x2, R.rec.y2) * R.rec.width);
      if (state == 0) {
        if (T.refine <= 0 || density >= T.density_th) { R.has_rect = 1; phase = P_DONE; }
        else phase = P_REFSTAT;
      } else if (density >= T.density_th) { R.has_rect = 1; phase = P_DONE; }
      else {
        if (state == 1) {
          const double r1 = dist_sq((double)sx, (double)sy, R.rec.x1, R.rec.y1), r2 = dist_sq((double)sx, (double)sy, R.rec.x2, R.rec.y2);
          rad_sq = r1 > r2 ? r1 : r2;
          state = 2;
        }
        phase = P_REDUCE;
      }
    }
    // ---------------- refine statistics (speculative: nothing is un-marked), again one region at a time ----------------
    m = __ballot_sync(0xffffffffu, phase == P_REFSTAT);
    while (m) {
      const int L = __ffs(m) - 1;
      m &= m - 1;
      const int* lp = (const int*)(size_t)__shfl_sync(0xffffffffu, (unsigned long long)(size_t)cur, L);
      const int ln = __shfl_sync(0xffffffffu, n, L);
      const int lsx = __shfl_sync(0xffffffffu, sx, L), lsy = __shfl_sync(0xffffffffu, sy, L);
      const double lsa = shfl_d(seed_ang, L), lw = shfl_d(R.rec.width, L);
      const double tau = coop_refine_tau<false>(T, lp, ln, lsx, lsy, lsa, lw);
      if (lane == L) {
        prec = tau;
        cur = reg + n; capc = cap - n; R.foff = n; stamp = stamp0 | 1u;   // keep the first region: it is part of E
        state = 1;
        phase = P_GROW;
      }
    }
  }
}

__device__ void grow_task_block(const Task& T, BlockShared& S) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int NW = blockDim.x >> 5, K = blockDim.x;
  const uint32_t lt = (1u << lane) - 1u;
  const int cap = (32 * T.lane_cap) / K;
  int* my_reg = T.reg_spec + (size_t)tid * cap;
  if (tid == 0) { S.cursor = 0; }
  int npend = 0;          // rectangles appended so far (uniform across the CTA)
  uint32_t wave = 0;
  long long t_sel = 0, t_spec = 0, t_commit = 0, t_redo = 0, n_redo = 0, n_round = 0, n_seed = 0;
  __syncthreads();
  while (true) {
    long long c0 = clock64();
    // ---- warp 0 selects the next (up to) K free seeds in order ----
    if (warp == 0) {
      int cursor = S.cursor, nsel = 0;
      // eight blocks of 32 list entries per trip: the two dependent loads (list entry, its state word) of all eight are in
      // flight together; the blocks are then consumed in order exactly as one at a time
      constexpr int kSelBlocks = 8;
      while (nsel < K && cursor < T.ndef) {
        int p[kSelBlocks];
        bool fr[kSelBlocks];
#pragma unroll
        for (int j = 0; j < kSelBlocks; j++) {
          const int idx = cursor + j * 32 + lane;
          p[j] = idx < T.ndef ? (int)T.order[idx] : -1;
        }
#pragma unroll
        for (int j = 0; j < kSelBlocks; j++) fr[j] = !(ld_state(T.state + (p[j] >= 0 ? p[j] : 0)) & kUsed) && p[j] >= 0;
        bool stop = false;
#pragma unroll
        for (int j = 0; j < kSelBlocks; j++) {
          if (stop) continue;
          const uint32_t m = __ballot_sync(0xffffffffu, fr[j]);
          const int c = __popc(m);
          const int take = min(c, K - nsel);
          const int rank = __popc(m & lt);
          if (fr[j] && rank < take) S.sel[nsel + rank] = p[j];
          nsel += take;
          if (c > take) { cursor += (int)__fns(m, 0, take) + 1; stop = true; }
          else { cursor += 32; stop = nsel >= K || cursor >= T.ndef; }
        }
      }
      if (lane == 0) { S.cursor = cursor; S.nsel = nsel; }
    }
    __syncthreads();
    const int nsel = S.nsel;
    if (nsel == 0) break;
    wave++;
    n_seed += nsel;
    const int my_seed = tid < nsel ? S.sel[tid] : -1;
    long long c1 = clock64(); t_sel += c1 - c0;
    SeedResult R;
    R.ok = 0; R.n1 = 0; R.n2_orig = 0; R.nf = 0; R.foff = 0; R.has_rect = 0;
    const uint32_t stamp = (wave << 11) | ((uint32_t)(K - 1 - tid) << 1);   // bit0 = phase, 10 bits of seed priority
    int verdict = 0;        // cached verdict: 0 unknown, 1 good (invalidated only by a re-run that wrote near me), 2 dead, 3 must re-run
    int kstar = -1;         // -1: the speculative pass of the wave; otherwise the seed that is re-run sequentially this round
    int settled = 0;        // seeds [0, settled) of the wave are committed, dropped or re-run (uniform across the CTA)
    long long c2 = c1;
    // every round starts with ONE call site of the per-seed pipeline (all seeds speculatively in the first round) or the
    // warp-cooperative sequential re-run of the first doubtful seed; three CTA barriers per round
    while (true) {
      n_round++;
      long long cr = clock64();
      if (kstar < 0) {
        if (warp * 32 < nsel) speculate_wave(T, tid < nsel, my_seed, my_reg, cap, stamp, R);
      } else if (warp == (kstar >> 5)) {
        // the whole warp of kstar re-runs that seed sequentially (exact sequential semantics).  The commits of the previous
        // round may have taken the seed: then the sequential algorithm skips it
        const int ks_lane = kstar & 31;
        const int seed_k = __shfl_sync(0xffffffffu, my_seed, ks_lane);
        const bool taken = (ld_state(T.state + seed_k) & kUsed) != 0;
        int has = 0;
        if (!taken) {
          SeedResult Q;
          process_seed_coop(T, seed_k, T.reg_serial, Q);
          has = Q.has_rect;
          if (lane == ks_lane) {
            if (has) append_rect(T, npend, Q.rec, (int)((wave << 11) | (tid << 1) | 1), my_seed, Q.nf);
            S.rb[0] = Q.bx0; S.rb[1] = Q.by0; S.rb[2] = Q.bx1; S.rb[3] = Q.by1;
          }
        } else if (lane == ks_lane) { S.rb[0] = 1; S.rb[2] = 0; }
        if (lane == ks_lane) S.has = has;
      }
      __syncthreads();                                                                     // (1) run / re-run finished
      if (kstar < 0) { c2 = clock64(); t_spec += c2 - c1; }
      else {
        npend += S.has;
        settled = kstar + 1;
        // the re-run wrote `used` inside its bounding box only: cached "good" verdicts elsewhere stay valid
        if (verdict == 1 && !(S.rb[0] > R.bx1 || S.rb[2] < R.bx0 || S.rb[1] > R.by1 || S.rb[3] < R.by0)) verdict = 0;
        t_redo += clock64() - cr;
      }
      if (settled >= nsel) break;
      const bool mine = tid >= settled && tid < nsel;
      bool dead = false, good = false;
      {
        // seeds without a verdict: seed still free?  then the ownership walk (warp-cooperative for long lists)
        const bool need = mine && verdict == 0;
        bool d0 = false;
        if (need) d0 = (ld_state(T.state + my_seed) & kUsed) != 0;
        const bool own = owns_all_coop_xy(T, need && !d0 && R.ok, my_reg, R.n1 + R.n2_orig, stamp >> 1);
        if (need) verdict = d0 ? 2 : (own ? 1 : 3);
      }
      if (mine) {
        if (verdict == 3) {
          // lost a pixel (or never finished): that is permanent; only "my seed was taken meanwhile" can still change
          dead = (ld_state(T.state + my_seed) & kUsed) != 0;
          if (dead) verdict = 2;
        }
        dead = verdict == 2; good = verdict == 1;
      }
      const uint32_t dm = __ballot_sync(0xffffffffu, dead), gm = __ballot_sync(0xffffffffu, good);
      const uint32_t rm = __ballot_sync(0xffffffffu, good && R.has_rect);
      if (lane == 0) { S.deadm[warp] = dm; S.goodm[warp] = gm; S.rectm[warp] = rm; }
      __syncthreads();                                                                     // (2) verdicts published
      // first unsettled seed that is neither dead nor provably good; rectangles of the good seeds before it, in seed order
      int ks = nsel, before = 0, total = 0;
      for (int v = 0; v < NW; v++) {
        const int first = v * 32;
        if (first >= nsel || first + 32 <= settled) continue;
        uint32_t pm = 0xffffffffu;
        if (settled > first) pm &= ~((1u << (settled - first)) - 1u);
        if (nsel < first + 32) pm &= (1u << (nsel - first)) - 1u;
        if (ks == nsel) {
          const uint32_t bad = pm & ~S.deadm[v] & ~S.goodm[v];
          if (bad) { ks = first + __ffs(bad) - 1; pm &= (1u << (ks - first)) - 1u; }
          const uint32_t rr = S.rectm[v] & pm;
          total += __popc(rr);
          if (v < warp) before += __popc(rr);
          else if (v == warp) before += __popc(rr & lt);
        }
      }
      const bool below = tid < ks;
      const bool do_commit = mine && good && below;
      mark_used_coop_xy(T, do_commit, my_reg + R.foff, R.nf);
      if (do_commit && R.has_rect) append_rect(T, npend + before, R.rec, (int)((wave << 11) | (tid << 1) | 0), my_seed, R.nf);
      npend += total;
      settled = ks;
      if (ks >= nsel) break;
      kstar = ks;
      n_redo++;
      __syncthreads();                                                                     // (3) commits visible to the re-run
    }
    t_commit += clock64() - c2;
    __syncthreads();
  }
  if (tid == 0) *T.npend = min(npend, T.pend_cap);
  if (tid == 0 && T.prof) {
    T.prof[0] = t_sel; T.prof[1] = t_spec; T.prof[2] = t_commit - t_redo; T.prof[3] = t_redo; T.prof[4] = wave; T.prof[5] = n_redo;
    T.prof[6] = n_round; T.prof[7] = n_seed;
  }
}

// ------------------------------------------------------------------------------------------------
// NFA validation of one rectangle by one thread (rect_nfa / nfa / rect_improve of the oracle).
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

// everything nfa() needs from the task goes by value: a `const Task&` parameter of an out-of-line function would force the whole task
// descriptor of the calling kernel into local memory
__device__ __noinline__ double nfa_v(const double log_nt, const double* __restrict__ lgam, const int lgam_n, int n, int k, double p) {
  auto log_gamma_int = [&](int i) { return i < lgam_n ? lgam[i] : log_gamma((double)i); };
  if (n == 0 || k == 0) return -log_nt;
  if (n == k) return -log_nt - (double)n * log10(p);
  const double p_term = p / (1 - p);
  const double log1term = log_gamma_int(n + 1) - log_gamma_int(k + 1) - log_gamma_int(n - k + 1) + (double)k * log(p) +
                          (double)(n - k) * log(1.0 - p);
  double term = exp(log1term);
  if (double_equal(term, 0)) {
    if (k > n * p) return -log1term / kLn10 - log_nt;
    return -log_nt;
  }
  double bin_tail = term;
  const double tolerance = 0.1;
  int i0 = k + 1;
  {
    // while i <= (n+1)/2 the binomial ratio (n-i+1)/i is >= 1 and the reference loop only multiplies and accumulates: run
    // that part with the (independent) divisions of four steps in flight; the products and the sum stay strictly in order
    const int m = min(n, (n + 1) / 2);
    for (; i0 + 3 <= m; i0 += 4) {
      const double b0 = (double)(n - i0 + 1) / (double)i0, b1 = (double)(n - i0) / (double)(i0 + 1);
      const double b2 = (double)(n - i0 - 1) / (double)(i0 + 2), b3 = (double)(n - i0 - 2) / (double)(i0 + 3);
      const double m0 = b0 * p_term, m1 = b1 * p_term, m2 = b2 * p_term, m3 = b3 * p_term;
      term *= m0; bin_tail += term;
      term *= m1; bin_tail += term;
      term *= m2; bin_tail += term;
      term *= m3; bin_tail += term;
    }
  }
  for (int i = i0; i <= n; ++i) {
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
__device__ __forceinline__ double nfa(const Task& T, int n, int k, double p) { return nfa_v(T.log_nt, T.lgam, T.lgam_n, n, k, p); }

// nfa(n, k, p) depends on two small integers, on p -- which rect_improve only ever halves, starting from T.p, at most ten times --
// and on the octave's log_nt.  The values for n <= kNfaTabN are therefore tabulated once per handle and geometry BY THIS VERY
// FUNCTION (k_nfa_table in line.cu), so a look-up returns bit for bit what the call would: the validation kernels then consist of
// the pixel scans alone -- no log / exp / pow / division chains, no divergent binomial-tail loops, a fraction of the code footprint
// (the second NFA pass was instruction-fetch bound with them).
constexpr int kNfaTabN = 512;
constexpr int kNfaTabLevels = 11;
constexpr int kNfaTabTri = (kNfaTabN + 1) * (kNfaTabN + 2) / 2;
__device__ __forceinline__ double nfa_lookup(const Task& T, int n, int k, double p) {
  if (T.nfa_tab && n <= kNfaTabN) {
    const int j = ((__double2hiint(T.p) >> 20) & 0x7ff) - ((__double2hiint(p) >> 20) & 0x7ff);   // number of halvings
    if ((unsigned)j < (unsigned)kNfaTabLevels && p * (double)(1 << j) == T.p)
      return T.nfa_tab[(size_t)j * kNfaTabTri + ((n * (n + 1)) >> 1) + k];
  }
  return nfa(T, n, k, p);
}

// counts the pixels of the rotated rectangle and those aligned with it.  Integer counting is order-independent, so any
// traversal gives the oracle's totals.  COOP = false: one thread per rectangle; COOP = true: one warp per rectangle (lanes
// over rows for tall rectangles, over columns otherwise; the NFA itself is evaluated by lane 0 and broadcast).
template <bool COOP>
__device__ __forceinline__ double rect_nfa_body(const Task& T, const Rect& rec) {
  const int lane = COOP ? (threadIdx.x & 31) : 0;
  const double hw = rec.width / 2.0;
  const double dyhw = rec.dy * hw, dxhw = rec.dx * hw;
  // the four corners, rotated so that the one with the smallest y (then x) comes first.  Everything stays in scalars: an array
  // indexed by the rotation would live in local memory (the second pass wrote 2.7 MB of it per frame to DRAM)
  const double vx0 = rec.x1 - dyhw, vx1 = rec.x2 - dyhw, vx2 = rec.x2 + dyhw, vx3 = rec.x1 + dyhw;
  const double vy0 = rec.y1 + dxhw, vy1 = rec.y2 + dxhw, vy2 = rec.y2 - dxhw, vy3 = rec.y1 - dxhw;
  int off = 0;
  {
    double by = vy0, bx = vx0;
    if (vy1 < by || (vy1 == by && vx1 < bx)) { off = 1; by = vy1; bx = vx1; }
    if (vy2 < by || (vy2 == by && vx2 < bx)) { off = 2; by = vy2; bx = vx2; }
    if (vy3 < by || (vy3 == by && vx3 < bx)) { off = 3; }
  }
  const bool o1 = off & 1, o2 = (off & 2) != 0;
  // element (i + off) & 3: first rotate by one if bit 0 is set, then by two if bit 1 is set
  const double ax0 = o1 ? vx1 : vx0, ax1 = o1 ? vx2 : vx1, ax2 = o1 ? vx3 : vx2, ax3 = o1 ? vx0 : vx3;
  const double ay0 = o1 ? vy1 : vy0, ay1 = o1 ? vy2 : vy1, ay2 = o1 ? vy3 : vy2, ay3 = o1 ? vy0 : vy3;
  const double px[4] = {o2 ? ax2 : ax0, o2 ? ax3 : ax1, o2 ? ax0 : ax2, o2 ? ax1 : ax3};
  const double py[4] = {o2 ? ay2 : ay0, o2 ? ay3 : ay1, o2 ? ay0 : ay2, o2 ? ay1 : ay3};
  const int c0 = (int)ceil(py[0]), c1 = (int)ceil(py[1]), c2 = (int)ceil(py[2]), c3 = (int)ceil(py[3]);
  const double flstep = (c1 != c0) ? (px[1] - px[0]) / (py[1] - py[0]) : 0.0;
  const double slstep = (c2 != c1) ? (px[2] - px[1]) / (py[2] - py[1]) : 0.0;
  const double frstep = (c3 != c0) ? (px[3] - px[0]) / (py[3] - py[0]) : 0.0;
  const double srstep = (c2 != c3) ? (px[2] - px[3]) / (py[2] - py[3]) : 0.0;
  const int ya = max(c0, 0), yb = min(c2, T.h - 1);
  int total = 0, alg = 0;
  const bool by_rows = COOP && (yb - ya + 1) >= 24;
  const int ystep = by_rows ? 32 : 1, xstep = (COOP && !by_rows) ? 32 : 1;
  for (int yy = ya + (by_rows ? lane : 0); yy <= yb; yy += ystep) {
    const double left = (yy <= c1) ? px[0] + ((double)yy - py[0]) * flstep : px[1] + ((double)yy - py[1]) * slstep;
    const double right = (yy < c3) ? px[0] + ((double)yy - py[0]) * frstep : px[3] + ((double)yy - py[3]) * srstep;
    if (!(right >= 0) || !(left <= (double)(T.w - 1))) continue;
    const int xb = (int)ceil(left > 0 ? left : 0.0), xe = (int)(right < (double)(T.w - 1) ? right : (double)(T.w - 1));
    const double* row = T.ang + (size_t)yy * T.w;
    for (int x = xb + ((COOP && !by_rows) ? lane : 0); x <= xe; x += xstep) {
      ++total;
      if (aligned_angle(row[x], rec.theta, rec.prec)) ++alg;
    }
  }
  if (!COOP) return nfa_lookup(T, total, alg, rec.p);
#pragma unroll
  for (int o = 16; o; o >>= 1) { total += __shfl_xor_sync(0xffffffffu, total, o); alg += __shfl_xor_sync(0xffffffffu, alg, o); }
  double v = 0;
  if (lane == 0) v = nfa_lookup(T, total, alg, rec.p);
  return __shfl_sync(0xffffffffu, v, 0);
}

// out-of-line copy for the callers with many call sites (rect_improve_rest); the per-thread passes, which call it from ONE place,
// use the body directly so that neither the rectangle nor the task descriptor has to live in local memory
template <bool COOP>
__device__ __noinline__ double rect_nfa(const Task& T, const Rect& rec) { return rect_nfa_body<COOP>(T, rec); }

// rect_improve after its first evaluation (log_nfa = rect_nfa(rec) <= log_eps): the 25 refinement attempts
template <bool COOP>
__device__ double rect_improve_rest(const Task& T, Rect& rec, double log_nfa) {
  const double delta = 0.5, delta_2 = delta / 2.0;
  const double log_eps = T.log_eps;
  Rect r = rec;
  for (int n = 0; n < 5; ++n) {
    r.p /= 2; r.prec = r.p * kPI;
    const double v = rect_nfa<COOP>(T, r);
    if (v > log_nfa) { log_nfa = v; rec = r; }
  }
  if (log_nfa > log_eps) return log_nfa;
  r = rec;
  for (int n = 0; n < 5; ++n) {
    if ((r.width - delta) >= 0.5) {
      r.width -= delta;
      const double v = rect_nfa<COOP>(T, r);
      if (v > log_nfa) { rec = r; log_nfa = v; }
    }
  }
  if (log_nfa > log_eps) return log_nfa;
  r = rec;
  for (int n = 0; n < 5; ++n) {
    if ((r.width - delta) >= 0.5) {
      r.x1 += -r.dy * delta_2; r.y1 += r.dx * delta_2; r.x2 += -r.dy * delta_2; r.y2 += r.dx * delta_2;
      r.width -= delta;
      const double v = rect_nfa<COOP>(T, r);
      if (v > log_nfa) { rec = r; log_nfa = v; }
    }
  }
  if (log_nfa > log_eps) return log_nfa;
  r = rec;
  for (int n = 0; n < 5; ++n) {
    if ((r.width - delta) >= 0.5) {
      r.x1 -= -r.dy * delta_2; r.y1 -= r.dx * delta_2; r.x2 -= -r.dy * delta_2; r.y2 -= r.dx * delta_2;
      r.width -= delta;
      const double v = rect_nfa<COOP>(T, r);
      if (v > log_nfa) { rec = r; log_nfa = v; }
    }
  }
  if (log_nfa > log_eps) return log_nfa;
  r = rec;
  for (int n = 0; n < 5; ++n) {
    if ((r.width - delta) >= 0.5) {
      r.p /= 2; r.prec = r.p * kPI;
      const double v = rect_nfa<COOP>(T, r);
      if (v > log_nfa) { rec = r; log_nfa = v; }
    }
  }
  return log_nfa;
}

template <bool COOP>
__device__ double rect_improve(const Task& T, Rect& rec) {
  const double log_nfa = rect_nfa<COOP>(T, rec);
  if (log_nfa > T.log_eps) return log_nfa;
  return rect_improve_rest<COOP>(T, rec, log_nfa);
}

// upper bound of the number of pixels one scan of the rectangle visits (width only shrinks during rect_improve)
__device__ __forceinline__ double rect_area_bound(const Rect& r) { return (dist(r.x1, r.y1, r.x2, r.y2) + 2.0) * (r.width + 2.0); }

__device__ __forceinline__ void write_segment(const Task& T, int i, const Rect& rec, bool acc);

// small rectangles, pass 1: the first NFA evaluation.  Returns true when the rectangle is settled (accepted); otherwise the
// value is parked in the (not yet used) segment slot and the rectangle goes to the second pass, where every thread of a warp
// has the same 25 evaluations ahead of it instead of a mix of 1 and 26
__device__ bool validate_first(const Task& T, int i) {
  const Rect rec = T.pend[i].rec;
  if (T.refine < 2) { write_segment(T, i, rec, true); return true; }
  const double v = rect_nfa_body<false>(T, rec);
  if (v > T.log_eps) { write_segment(T, i, rec, true); return true; }
  *reinterpret_cast<double*>(T.pend[i].seg) = v;
  return false;
}
// Variation `n` (0..4) of refinement stage `stage` (0..4) of rect_improve, built from the stage's starting rectangle with the
// same repeated operations as the sequential loops (so the doubles are identical).  Inside a stage the rectangle tried at
// iteration n does not depend on the outcome of iterations 0..n-1 -- only `rec`/`log_nfa` do -- which is what lets five lanes
// evaluate a stage at once.  Returns false when the reference loop skips that iteration (width cannot shrink further).
__device__ __forceinline__ bool rect_variation(int stage, int n, Rect& r) {
  const double delta = 0.5, delta_2 = delta / 2.0;
  if (stage == 0) {
    for (int j = 0; j <= n; j++) r.p /= 2;
    r.prec = r.p * kPI;
    return true;
  }
  if (stage == 4) {
    if (!((r.width - delta) >= 0.5)) return false;
    for (int j = 0; j <= n; j++) r.p /= 2;
    r.prec = r.p * kPI;
    return true;
  }
  for (int j = 0; j <= n; j++) {
    if (!((r.width - delta) >= 0.5)) return false;
    if (stage == 2) { r.x1 += -r.dy * delta_2; r.y1 += r.dx * delta_2; r.x2 += -r.dy * delta_2; r.y2 += r.dx * delta_2; }
    if (stage == 3) { r.x1 -= -r.dy * delta_2; r.y1 -= r.dx * delta_2; r.x2 -= -r.dy * delta_2; r.y2 -= r.dx * delta_2; }
    r.width -= delta;
  }
  return true;
}

// Pass 2 by one warp: six rectangles at a time, five lanes each (lanes 30/31 idle).  Every trip of the loop is one stage of
// rect_improve for the six rectangles: lane (g, n) evaluates variation n, the five values are folded in iteration order
// (strict >, first maximum wins), the group moves to the next stage or settles the rectangle and takes the next one of the
// warp's share (positions first, first + stride, ... of the failed list, counted from the back of `back`).
__device__ void validate_rest_warp(const Task& T, const int* back, int nf, int first, int stride) {
  const int lane = threadIdx.x & 31, grp = lane / 5, var = lane - grp * 5;
  const bool worker = grp < 6;
  const int gl0 = worker ? grp * 5 : 25;
  int next = first, idx = -1, stage = 0;
  Rect rec;
  double log_nfa = 0;
#pragma unroll 1
  while (true) {
#pragma unroll 1
    for (int g = 0; g < 6; g++) {
      const bool idle = __shfl_sync(0xffffffffu, idx, g * 5) < 0;
      if (idle && next < nf) {
        if (grp == g) {
          idx = back[-next];
          rec = T.pend[idx].rec;
          log_nfa = *reinterpret_cast<const double*>(T.pend[idx].seg);
          stage = 0;
        }
        next += stride;
      }
    }
    const bool busy = worker && idx >= 0;
    if (!__any_sync(0xffffffffu, busy)) break;
    double v = 0;
    int have = 0;
    if (busy) {
      Rect r = rec;
      have = rect_variation(stage, var, r) ? 1 : 0;
      if (have) v = rect_nfa_body<false>(T, r);
    }
    int best = -1;
    double bv = log_nfa;
#pragma unroll
    for (int n = 0; n < 5; n++) {
      const double vn = shfl_d(v, gl0 + n);
      const int hn = __shfl_sync(0xffffffffu, have, gl0 + n);
      if (hn && vn > bv) { bv = vn; best = n; }
    }
    if (busy) {
      if (best >= 0) { rect_variation(stage, best, rec); log_nfa = bv; }
      ++stage;
      if (log_nfa > T.log_eps || stage == 5) {
        if (var == 0) write_segment(T, idx, rec, log_nfa > T.log_eps);
        idx = -1;
      }
    }
  }
}

__device__ __forceinline__ void write_segment(const Task& T, int i, const Rect& rec, bool acc) {
  Pending& P = T.pend[i];
  P.accepted = acc ? 1 : 0;
  if (acc) {
    double x1 = rec.x1 + 0.5, y1 = rec.y1 + 0.5, x2 = rec.x2 + 0.5, y2 = rec.y2 + 0.5;
    if (T.scale != 1.0) { x1 /= T.scale; y1 /= T.scale; x2 /= T.scale; y2 /= T.scale; }
    P.seg[0] = (float)x1; P.seg[1] = (float)y1; P.seg[2] = (float)x2; P.seg[3] = (float)y2;
  }
}

// validate pending rectangle i of task T and write its segment (COOP = false: by one thread; true: by one warp)
template <bool COOP>
__device__ void validate_pending(const Task& T, int i) {
  Rect rec = T.pend[i].rec;
  bool acc = true;
  if (T.refine >= 2) {
    const double v = rect_improve<COOP>(T, rec);
    acc = v > T.log_eps;
  }
  if (!COOP || (threadIdx.x & 31) == 0) {
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
