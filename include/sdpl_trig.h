/*
 * sdpl_trig.h -- double-precision sin / cos in STRICT IEEE arithmetic (+, -, * on doubles, no fused multiply-add, no libm),
 * shared verbatim by the CUDA kernels (csrc/*.cu, compiled with --fmad=false) and by the CPU oracle (oracle/*.cpp, compiled
 * with -ffp-contract=off), so that both sides get bit-identical values.
 *
 * Why: the LSD rectangle of a region is built from cos(theta), sin(theta) (cv::LineSegmentDetector region2rect) and then
 * scanned pixel by pixel (rect_nfa: ceil / int of the rotated edges).  CUDA's and glibc's sin/cos differ in the last bit for
 * a fraction of the arguments; one flipped edge pixel changes the NFA count and, about once per 3000 rectangles, the verdict
 * on a line.  The reference leaves these functions to the C library of the machine it runs on (glibc itself picks an FMA or a
 * non-FMA build of s_sin.c at load time, with different last bits), so the choice is pinned here instead
 * (oracle decision 9).
 *
 * Method (the classical table + polynomial scheme, as in glibc's dbl-64/s_sin.c): Cody-Waite reduction by pi/2 with a
 * 4-part constant, a 128-entry table of double-double sin/cos at k/128 (sdpl_trig_table.inc, generated with mpmath) and
 * short Taylor polynomials for the remainder.  Error below 0.6 ulp; verified against mpmath and against libm in
 * tests/test_trig.py.  Valid for |x| < 1e5 (the front-end's angles are within a few turns); larger arguments and
 * non-finite values are not supported.
 */
#ifndef SDPL_TRIG_H
#define SDPL_TRIG_H

#if defined(__CUDACC__)
#define SDPL_HD __host__ __device__ __forceinline__
#define SDPL_TRIG_CONST __device__ __constant__
#else
#define SDPL_HD static inline
#endif

#if defined(__CUDACC__)
__device__ static const double sdpl_trig_tab_dev[512] = {
#include "sdpl_trig_table.inc"
};
#endif
static const double sdpl_trig_tab_host[512] = {
#include "sdpl_trig_table.inc"
};

#if defined(__CUDA_ARCH__)
#define SDPL_TRIG_TAB sdpl_trig_tab_dev
#else
#define SDPL_TRIG_TAB sdpl_trig_tab_host
#endif

/* sin(x + dx) for |x| <= 0.87, |dx| << |x| */
SDPL_HD double sdpl_k_sin(double x, double dx) {
  const double ax = x < 0 ? -x : x;
  if (ax < 0.126) {
    const double xx = x * x;
    const double s1 = -0x1.5555555555555p-3, s2 = 0x1.1111111110ecep-7, s3 = -0x1.a01a019db08b8p-13, s4 = 0x1.71de27b9a7ed9p-19,
                 s5 = -0x1.addffc2fcdf59p-26;
    const double poly = (((s5 * xx + s4) * xx + s3) * xx + s2) * xx + s1;
    const double t = (poly * x - 0.5 * dx) * xx + dx;
    return x + t;
  }
  if (x <= 0) dx = -dx;
  const double big = 0x1.8p45;                     /* ax + big rounds ax to a multiple of 1/128 */
  const double u = big + ax;
  const double r = ax - (u - big);
  const int k = (int)((u - big) * 128.0);
  const double xx = r * r;
  const double sn3 = -1.66666666666664880952546298448555E-01, sn5 = 8.33333214285722277379541354343671E-03;
  const double cs2 = 4.99999999999999999999950396842453E-01, cs4 = -4.16666666666664434524222570944589E-02,
               cs6 = 1.38888874007937613028114285595617E-03;
  const double s = r + (dx + r * xx * (sn3 + xx * sn5));
  const double c = r * dx + xx * (cs2 + xx * (cs4 + xx * cs6));
  const double sn = SDPL_TRIG_TAB[4 * k], ssn = SDPL_TRIG_TAB[4 * k + 1], cs = SDPL_TRIG_TAB[4 * k + 2], ccs = SDPL_TRIG_TAB[4 * k + 3];
  const double cor = (ssn + s * ccs - sn * c) + cs * s;
  const double v = sn + cor;
  return x < 0 ? -v : v;
}

/* cos(x + dx) for |x| <= 0.87 */
SDPL_HD double sdpl_k_cos(double x, double dx) {
  if (x < 0) dx = -dx;
  const double ax = x < 0 ? -x : x;
  const double big = 0x1.8p45;
  const double u = big + ax;
  const double r = ax - (u - big) + dx;
  const int k = (int)((u - big) * 128.0);
  const double xx = r * r;
  const double sn3 = -1.66666666666664880952546298448555E-01, sn5 = 8.33333214285722277379541354343671E-03;
  const double cs2 = 4.99999999999999999999950396842453E-01, cs4 = -4.16666666666664434524222570944589E-02,
               cs6 = 1.38888874007937613028114285595617E-03;
  const double s = r + r * xx * (sn3 + xx * sn5);
  const double c = xx * (cs2 + xx * (cs4 + xx * cs6));
  const double sn = SDPL_TRIG_TAB[4 * k], ssn = SDPL_TRIG_TAB[4 * k + 1], cs = SDPL_TRIG_TAB[4 * k + 2], ccs = SDPL_TRIG_TAB[4 * k + 3];
  const double cor = (ccs - s * ssn - cs * c) - sn * s;
  return cs + cor;
}

/* x = n * pi/2 + (a + da), |a| <= pi/4; returns n mod 4.  Valid for |x| < 1e5. */
SDPL_HD int sdpl_reduce_pio2(double x, double* a, double* da) {
  const double hpinv = 0x1.45f306dc9c883p-1;       /* 2/pi */
  const double toint = 0x1.8p52;
  const double mp1 = 0x1.921fb58000000p0, mp2 = -0x1.dde973c000000p-27, pp3 = -0x1.cb3b398000000p-55, pp4 = -0x1.d747f23e32ed7p-83;
  const double t = x * hpinv + toint;
  const double xn = t - toint;
  const double y = (x - xn * mp1) - xn * mp2;
  const int n = (int)xn & 3;
  double t1 = xn * pp3;
  const double t2 = y - t1;
  double db = (y - t2) - t1;
  t1 = xn * pp4;
  const double b = t2 - t1;
  db += (t2 - b) - t1;
  *a = b; *da = db;
  return n;
}

SDPL_HD void sdpl_sincos(double x, double* sn, double* cs) {
  const double ax = x < 0 ? -x : x;
  if (ax < 0.855469) {
    *sn = ax < 0x1p-26 ? x : sdpl_k_sin(x, 0.0);
    *cs = ax < 0x1p-27 ? 1.0 : sdpl_k_cos(x, 0.0);
    return;
  }
  double a, da;
  const int n = sdpl_reduce_pio2(x, &a, &da);
  const double s = sdpl_k_sin(a, da), c = sdpl_k_cos(a, da);
  /* sin(x) = {s, c, -s, -c}[n], cos(x) = {c, -s, -c, s}[n] */
  const double sv = (n & 1) ? c : s, cv = (n & 1) ? s : c;
  *sn = (n & 2) ? -sv : sv;
  *cs = ((n + 1) & 2) ? -cv : cv;
}
SDPL_HD double sdpl_sin(double x) { double s, c; sdpl_sincos(x, &s, &c); return s; }
SDPL_HD double sdpl_cos(double x) { double s, c; sdpl_sincos(x, &s, &c); return c; }

#endif
