// lsd_grow2.cuh -- the region-growing schedule of round 2: "two-phase waves".
//
// Same speculate / validate / commit scheme as lsd_grow.cuh: a wave takes the next K still-free seeds in
// seed order, every seed is grown speculatively against the committed `used` map with a stamp whose priority is the seed's
// rank, seeds are then validated and committed in rank order and the first doubtful one is re-run with exact sequential
// semantics.  What changed is WHO grows a region.  ncu on the round-1 kernel (profiles/r1_lsd_kernels_final_*): 60 % of the
// issued instructions sat in the per-lane growth loop with 3.8 of 32 lanes active -- a warp spends most of a wave waiting
// for the one or two lanes that own a long region, and the CTA waits for that warp (39 % of the stall samples at the wave
// barrier).  Now
//   phase A: every lane grows its own seed for at most T_A expansions (all lanes busy: most seeds give regions of a few
//            pixels and end here);
//   phase B: the seeds that are not finished are queued for the whole CTA and each is taken over by a WARP
//            (seed_pipeline_coop): four list pixels are expanded per step, lane (j, s) tests neighbour s of pixel j.  The
//            sequential algorithm updates the region angle after every join, so the warp first decides all 32 tests
//            against the angle at the start of the step, forms the float prefix sums of the predicted joins in scan
//            order, recomputes the exact angle every test would have seen and compares: the first test whose exact decision
//            differs ends the accepted prefix and the rest of the step is redone from there (coop_grow).  The result is the
//            sequential result by induction over the scan order; in the common case one step costs one round of loads and one
//            polynomial arctangent for ~3 expansions instead of ~3 dependent rounds.  Rectangle fits and refine statistics were
//            already warp-cooperative; the sequential re-runs use the same routine without stamps.
// (Of the ~1000 re-run rounds of a frame ~730 find their seed taken by the commits of the same round; the ~280 real re-runs grow 20 pixels
// on average and cost 40 k cycles each -- profiles/r2_rerun_statistics.txt.  Letting ONE lane grow such a region before the warp takes
// over was measured slower, 8.0 -> 8.7 ms per frame of re-runs, and was removed.)
// Results of a seed live in a per-slot context in global memory (SlotCtx) instead of the registers of "its" thread, so any
// warp can finish any seed and the commit rounds read them back.
#pragma once
#include "lsd_grow.cuh"

namespace sdpl {
namespace lsd {

constexpr int kCtxSlots = 512;              // seed slots per task (>= seeds per wave)
constexpr int kMaxGrowWarps2 = 16;          // warps per task of the two-phase schedule (512-seed waves)

// position of the (k+1)-th set bit of m (k < popc(m)): five popc steps instead of the software loop behind __fns
__device__ __forceinline__ int nth_set_bit(uint32_t m, int k) {
  int pos = 0;
#pragma unroll
  for (int wdt = 16; wdt; wdt >>= 1) {
    const int c = __popc(m & ((1u << wdt) - 1u));
    if (k >= c) { k -= c; m >>= wdt; pos += wdt; }
  }
  return pos;
}
// q / w and q % w for 0 <= q < 2^31 with a per-task reciprocal (one real division per CTA, none per seed)
struct DivW { uint32_t w, magic; };
__device__ __forceinline__ DivW make_divw(int w) { DivW d; d.w = (uint32_t)w; d.magic = (uint32_t)(0x100000000ull / (uint32_t)w); return d; }
__device__ __forceinline__ void divmod_w(const DivW d, int q, int& y, int& x) {
  uint32_t qy = __umulhi((uint32_t)q, d.magic);      // floor(2^32 / w) under-estimates: at most a few units too small
  uint32_t r = (uint32_t)q - qy * d.w;
  while (r >= d.w) { r -= d.w; ++qy; }
  y = (int)qy; x = (int)r;
}
__device__ __noinline__ void sincos_call(double a, double* sn, double* cs) { sdpl_sincos(a, sn, cs); }
__device__ __forceinline__ double dmax_(double a, double b) { return a > b ? a : b; }   // operands are never NaN here
__device__ __forceinline__ double dmin_(double a, double b) { return a < b ? a : b; }

struct __align__(16) SlotCtx {
  Rect rec;                                 // rectangle of the seed when flags has kCtxRect
  double ra;                                // hand-over state of an unfinished first growth: region angle ...
  float sumdx, sumdy;                       // ... unit-vector sums ...
  int n, i;                                 // ... list length and expansion cursor
  int n1, n2o, nf, foff;                    // first region size, re-grown slots, pixels finally marked: list[foff, foff + nf)
  int flags;
  int bx0, by0, bx1, by1;                   // bounding box of every pixel the seed touched
  int pad[3];
};
enum { kCtxOk = 1, kCtxRect = 2 };

// ------------------------------------------------------------------------------------------------
// Warp-cooperative region growth from expansion cursor i to completion.  Warp-uniform: n, i, sumdx, sumdy, ra.
// SPEC: speculative (stamps, abort on a pixel of an earlier seed or a full list -> returns false); otherwise exact sequential
// semantics (marks `used`).
// ------------------------------------------------------------------------------------------------
__device__ __noinline__ bool coop_grow(const Task& T, const bool SPEC, int* const list, const int cap, int& n, int i, float& sumdx, float& sumdy,
                                       double& ra, const double prec, const uint32_t stamp, int& bx0, int& by0, int& bx1, int& by1) {
  const uint32_t FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31, w = T.w, h = T.h;
  const uint32_t lt = (1u << lane) - 1u;
  const int j = lane >> 3, s = lane & 7;
  const int k9 = s + (s >> 2);                       // neighbour slots 0..8 without the centre, scan order
  const int ndx = k9 % 3 - 1, ndy = k9 / 3 - 1;
  int lx0 = bx0, ly0 = by0, lx1 = bx1, ly1 = by1;
  bool aborted = false;
#pragma unroll 1
  while (i < n) {
    const int m = min(4, n - i);
    if (T.prof_detail == 1 && threadIdx.x == 0) T.prof[14] += 1;
    const bool act = j < m;
    const int p = list[i + (act ? j : 0)];
    const int qx = xy_x(p) + ndx, qy = xy_y(p) + ndy;
    const bool inb = act && (unsigned)qx < (unsigned)w && (unsigned)qy < (unsigned)h;
    const int q = inb ? qy * w + qx : 0;
    uint32_t sv = kUsed;
    PxA pa; pa.ang = kNotDef; pa.c = 0.f; pa.s = 0.f;
    if (inb) ld_rec(T.px.p + q, sv, pa);
    i += m;
    bool cand = inb && !(sv & kUsed) && pa.ang != kNotDef;
    if (SPEC) cand = cand && sv != stamp;
    uint32_t und = __ballot_sync(FULL, cand);
    if (!und) continue;                              // nothing was written: no ordering needed before the next loads
    const bool foreign = SPEC && sv > stamp;         // carries the stamp of an earlier seed of this wave
    // lanes that test the same pixel (an arithmetic version -- four broadcasts of the list pixels and neighbourhood tests -- was
    // 6 % slower than this one instruction: 47.4 -> 50.2 ms per 512 frames)
    const uint32_t grp = __match_any_sync(FULL, cand ? q : ~lane);
    const uint32_t grp_before = grp & lt;
#pragma unroll 1
    while (und) {
      const bool mine = (und >> lane) & 1u;
      // decision of every open test against the angle at the start of this round
      const bool d0 = mine && aligned_angle(pa.ang, ra, prec);
      const uint32_t d0m = __ballot_sync(FULL, d0);
      if (!d0m) break;                               // no join at all: the angle does not move, every decision is exact
      const bool j0 = d0 && !(grp_before & d0m);     // predicted joins: the first test of each pixel
      const uint32_t j0m = __ballot_sync(FULL, j0);
      // unit-vector sums after the predicted joins, in scan order (strictly sequential float additions); psx/psy = the sums
      // this lane's test sees (joins before it), runx/runy = after all of them
      float runx = sumdx, runy = sumdy, psx = sumdx, psy = sumdy;
      for (uint32_t mm = j0m; mm; mm &= mm - 1) {
        const int L = __ffs(mm) - 1;
        runx = __fadd_rn(runx, __shfl_sync(FULL, pa.c, L));
        runy = __fadd_rn(runy, __shfl_sync(FULL, pa.s, L));
        if (L < lane) { psx = runx; psy = runy; }
      }
      const bool moved = (j0m & lt) != 0;
      const double ra_me = moved ? (double)fast_atan2_deg(psy, psx) * kDegToRad : ra;
      const double ra_end = (double)fast_atan2_deg(runy, runx) * kDegToRad;
      // exact decisions; a test shadowed by an earlier predicted join of the same pixel is skipped by the sequential
      // algorithm as well (the pixel is used by then), provided that join is confirmed -- which the prefix rule below ensures
      const bool shadowed = (grp_before & j0m) != 0;
      const bool e = mine && !shadowed && aligned_angle(pa.ang, ra_me, prec);
      const uint32_t em = __ballot_sync(FULL, e);
      const uint32_t mism = __ballot_sync(FULL, mine && !shadowed && (e != d0));
      const int F = mism ? __ffs(mism) - 1 : 32;     // first test whose exact decision differs from the prediction
      const uint32_t belowF = F >= 32 ? FULL : ((1u << F) - 1u);
      uint32_t acc = j0m & belowF;                   // confirmed joins: everything before F ...
      const bool f_joins = F < 32 && ((em >> F) & 1u);
      if (f_joins) acc |= 1u << F;                   // ... and F itself by its exact decision
      if (F < 32) {
        float fx = __shfl_sync(FULL, psx, F), fy = __shfl_sync(FULL, psy, F);
        const float cF = __shfl_sync(FULL, pa.c, F), sF = __shfl_sync(FULL, pa.s, F);
        if (f_joins) { fx = __fadd_rn(fx, cF); fy = __fadd_rn(fy, sF); }
        sumdx = fx; sumdy = fy;
        if (acc) ra = (double)fast_atan2_deg(sumdy, sumdx) * kDegToRad;
      } else {
        sumdx = runx; sumdy = runy; ra = ra_end;
      }
      const int cnt = __popc(acc);
      const bool ja = (acc >> lane) & 1u;
      if (SPEC) {
        if (__any_sync(FULL, ja && foreign) || n + cnt > cap) { aborted = true; break; }
      }
      if (ja) {
        list[n + __popc(acc & lt)] = xy_pack(qx, qy);
        if (SPEC) claim_max(&T.state[q], stamp); else T.state[q] = kUsed;
        lx0 = min(lx0, qx); lx1 = max(lx1, qx); ly0 = min(ly0, qy); ly1 = max(ly1, qy);
      }
      n += cnt;
      if (F >= 32) break;                            // every open test is decided
      // still open: the tests after F whose pixel has not joined meanwhile
      const uint32_t gone = __ballot_sync(FULL, (grp & acc) != 0);
      und &= ~(belowF | (1u << F) | gone);
    }
    if (aborted) break;
    __syncwarp();                                    // list / state writes of this step before the next step's loads
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    lx0 = min(lx0, __shfl_xor_sync(FULL, lx0, o)); ly0 = min(ly0, __shfl_xor_sync(FULL, ly0, o));
    lx1 = max(lx1, __shfl_xor_sync(FULL, lx1, o)); ly1 = max(ly1, __shfl_xor_sync(FULL, ly1, o));
  }
  bx0 = lx0; by0 = ly0; bx1 = lx1; by1 = ly1;
  return !aborted;
}

// ------------------------------------------------------------------------------------------------
// region2rect (+ get_theta) by a whole warp, as coop_region2rect of lsd_grow.cuh, but the sums that have to be accumulated in
// list order (bit-identical to the sequential loops of the reference) are staged in shared memory: every lane computes the
// three terms of its element, then lanes 0..2 each add up one of the three sums over the 32 staged values.  The shuffle
// version spent 18 warp instructions per element on moving the terms to every lane (a quarter of the kernel's instructions,
// profiles/r2_grow2_v1); this one spends two.  sb = 3 x 32 doubles of this warp.
// ------------------------------------------------------------------------------------------------
__device__ __noinline__ void coop_region2rect_smem(const Task& T, double* const sb, const int* list, int n, double ra, Rect& rec) {
  const int lane = threadIdx.x & 31, w = T.w;
  const int row = lane < 3 ? lane : 0;
  double acc = 0;
  for (int b = 0; b < n; b += 32) {
    const int e = b + lane;
    double wgt = 0, qxw = 0, qyw = 0;
    if (e < n) {
      const int q = list[e];
      const int qy = xy_y(q), qx = xy_x(q);
      wgt = modgrad_of(T.g2[qy * w + qx]);
      qxw = (double)qx * wgt; qyw = (double)qy * wgt;
    }
    sb[lane] = qxw; sb[32 + lane] = qyw; sb[64 + lane] = wgt;
    __syncwarp();
    const int m = min(32, n - b);
    if (lane < 3) {
      const double* src = sb + 32 * row;
#pragma unroll 8
      for (int j = 0; j < m; j++) acc += src[j];
    }
    __syncwarp();
  }
  double x = shfl_d(acc, 0), y = shfl_d(acc, 1);
  const double sum = shfl_d(acc, 2);
  x /= sum; y /= sum;
  acc = 0;
  for (int b = 0; b < n; b += 32) {
    const int e = b + lane;
    double t0 = 0, t1 = 0, t2 = 0;
    if (e < n) {
      const int q = list[e];
      const int qy = xy_y(q), qx = xy_x(q);
      const double dx = (double)qx - x, dy = (double)qy - y, wgt = modgrad_of(T.g2[qy * w + qx]);
      t0 = dy * dy * wgt; t1 = dx * dx * wgt; t2 = dx * dy * wgt;
    }
    sb[lane] = t0; sb[32 + lane] = t1; sb[64 + lane] = -t2;         // Ixy -= t  ==  Ixy += -t in IEEE arithmetic
    __syncwarp();
    const int m = min(32, n - b);
    if (lane < 3) {
      const double* src = sb + 32 * row;
#pragma unroll 8
      for (int j = 0; j < m; j++) acc += src[j];
    }
    __syncwarp();
  }
  const double Ixx = shfl_d(acc, 0), Iyy = shfl_d(acc, 1), Ixy = shfl_d(acc, 2);
  const double lambda = 0.5 * (Ixx + Iyy - sqrt((Ixx - Iyy) * (Ixx - Iyy) + 4.0 * Ixy * Ixy));
  double theta = (fabs(Ixx) > fabs(Iyy)) ? (double)fast_atan2_deg((float)(lambda - Ixx), (float)Ixy)
                                         : (double)fast_atan2_deg((float)Ixy, (float)(lambda - Iyy));
  theta *= kDegToRad;
  if (angle_diff(theta, ra) > T.prec) theta += kPI;
  double dy_, dx_;
  sincos_call(theta, &dy_, &dx_);
  double l_min = 0, l_max = 0, w_min = 0, w_max = 0;      // min / max are order-independent: plain warp reductions
#pragma unroll 1
  for (int e = lane; e < n; e += 32) {
    const int q = list[e];
    const int qy = xy_y(q), qx = xy_x(q);
    const double rdx = (double)qx - x, rdy = (double)qy - y;
    const double l = rdx * dx_ + rdy * dy_;
    const double ww = -rdx * dy_ + rdy * dx_;
    l_max = dmax_(l_max, l); l_min = dmin_(l_min, l);
    w_max = dmax_(w_max, ww); w_min = dmin_(w_min, ww);
  }
#pragma unroll 1
  for (int o = 16; o; o >>= 1) {
    l_max = dmax_(l_max, shfl_d(l_max, lane ^ o)); l_min = dmin_(l_min, shfl_d(l_min, lane ^ o));
    w_max = dmax_(w_max, shfl_d(w_max, lane ^ o)); w_min = dmin_(w_min, shfl_d(w_min, lane ^ o));
  }
  rec.x1 = x + l_min * dx_; rec.y1 = y + l_min * dy_;
  rec.x2 = x + l_max * dx_; rec.y2 = y + l_max * dy_;
  rec.width = w_max - w_min;
  rec.x = x; rec.y = y; rec.theta = theta; rec.dx = dx_; rec.dy = dy_; rec.prec = T.prec; rec.p = T.p;
  if (rec.width < 1.0) rec.width = 1.0;
}

// ------------------------------------------------------------------------------------------------
// The per-seed pipeline of lsd_detect's main loop by one warp (all scalars warp-uniform): growth -> region2rect -> refine
// (tolerance from the angle spread, re-growth, radius reduction).  n > 0 resumes a first growth handed over by phase A.
// SPEC = false is the exact sequential algorithm (marks / un-marks `used`, one list re-used); SPEC = true keeps the first
// region in front of the re-grown one (both are part of the set E the commit validates) and writes stamps only.
// ------------------------------------------------------------------------------------------------
__device__ __noinline__ void seed_pipeline_coop(const Task& T, const bool SPEC, double* const sb, const int seed, const int sx, const int sy,
                                                int* const reg, const int cap,
                                                const uint32_t stamp0, int n, int i, float sumdx, float sumdy, double ra, int bx0, int by0,
                                                int bx1, int by1, SeedResult& R) {
  const int lane = threadIdx.x & 31, w = T.w;
  R.ok = 1; R.n1 = 0; R.n2_orig = 0; R.has_rect = 0; R.foff = 0; R.nf = 0;
  if (n == 0) { bx0 = bx1 = sx; by0 = by1 = sy; }
  int state = 0;                 // 0: first growth, 1: re-growth with the refined tolerance, 2: radius reduction passes
  int* cur = reg;
  int capc = cap;
  double prec = T.prec, rad_sq = 0;
  uint32_t stamp = stamp0;
  const double seed_ang = T.px.ang(seed);
#pragma unroll 1
  while (true) {
    if (state <= 1) {
      if (n == 0) {
        if (SPEC) {
          if (capc < 1 || ld_state(T.state + seed) > stamp) { R.ok = 0; break; }
        }
        if (lane == 0) {
          if (SPEC) claim_max(&T.state[seed], stamp); else T.state[seed] = kUsed;
          cur[0] = xy_pack(sx, sy);
        }
        n = 1; i = 0;
        ra = seed_ang;
        double sn, cs;
        sincos_call(ra, &sn, &cs);
        sumdx = (float)cs; sumdy = (float)sn;
        __syncwarp();
      }
      const long long pg0 = clock64();
      const bool ok = coop_grow(T, SPEC, cur, capc, n, i, sumdx, sumdy, ra, prec, stamp, bx0, by0, bx1, by1);
      if (T.prof_detail == 1 && threadIdx.x == 0) T.prof[11] += clock64() - pg0;
      if (T.prof_detail == 2 && !SPEC && lane == 0) atomicAdd((unsigned long long*)&T.prof[10], (unsigned long long)(clock64() - pg0));
      if (state == 0) { R.n1 = n; R.nf = n; } else { if (SPEC) R.n2_orig = n; R.nf = n; }
      if (!ok) { R.ok = 0; break; }
      if (state == 0 ? (n < T.min_reg) : (n < 2)) break;
    } else {
      // one pass of reduce_region_radius: order-dependent swap removal, done by one lane (rare)
      rad_sq *= 0.75 * 0.75;
      if (lane == 0) {
        for (int e = 0; e < n; ++e) {
          const int q = cur[e];
          if (dist_sq((double)sx, (double)sy, (double)xy_x(q), (double)xy_y(q)) > rad_sq) {
            if (!SPEC) T.state[xy_lin(q, w)] = 0;
            cur[e] = cur[n - 1];
            cur[n - 1] = q;
            --n;
            --e;
          }
        }
      }
      n = __shfl_sync(0xffffffffu, n, 0);
      __syncwarp();
      R.nf = n;
      if (n < 2) break;
    }
    const long long pr0 = clock64();
    coop_region2rect_smem(T, sb, cur, n, ra, R.rec);
    if (T.prof_detail == 1 && threadIdx.x == 0) T.prof[12] += clock64() - pr0;
    const double density = (double)n / (dist(R.rec.x1, R.rec.y1, R.rec.x2, R.rec.y2) * R.rec.width);
    if (state == 0) {
      if (T.refine <= 0 || density >= T.density_th) { R.has_rect = 1; break; }
      const long long pt0 = clock64();
      prec = SPEC ? coop_refine_tau<false>(T, cur, n, sx, sy, seed_ang, R.rec.width) : coop_refine_tau<true>(T, cur, n, sx, sy, seed_ang, R.rec.width);
      if (T.prof_detail == 1 && threadIdx.x == 0) T.prof[13] += clock64() - pt0;
      __syncwarp();
      state = 1;
      if (SPEC) { cur = reg + n; capc = cap - n; R.foff = n; stamp = stamp0 | 1u; }   // the first region stays: it is part of E
      n = 0;
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

// ------------------------------------------------------------------------------------------------
// Phase A: every lane grows ONE seed at a time for at most `ta` expansions, speculatively, with the arithmetic and claim rules
// of the per-lane growth of round 1; a lane whose seed is settled (finished below the minimum region size, aborted, or
// handed over to phase B) writes the seed's context and takes the next seed slot of the wave from a CTA-wide counter, so the
// warp keeps all its lanes busy as long as the wave has seeds left (the static one-seed-per-thread version ran at 7 of 32
// lanes: most seeds end after one or two expansions).
// ------------------------------------------------------------------------------------------------
template <int NW>
struct Block2Shared {
  int sel[32 * NW];
  int q[32 * NW];                        // seed slots waiting for phase B
  uint32_t deadm[NW], goodm[NW], rectm[NW], simplem[NW], cdeadm[NW];
  int nsel, cursor, has, qn, qhead;
  int anext;                             // phase A: next seed slot to hand to an idle lane
  int rb[4];                             // bounding box written by the last re-run
  double sb[NW][96];                     // per warp: staged terms of the ordered rectangle-fit sums
};

template <int NW>
__device__ __forceinline__ void phase_a(const Task& T, SlotCtx* const ctx, Block2Shared<NW>& S, const int nsel, const int cap, const int ta,
                                        const uint32_t wave) {
  const uint32_t FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31, K = 32 * NW;
  const uint32_t lt = (1u << lane) - 1u;
  const int w = T.w, h = T.h;
  const double prec = T.prec;
  const DivW dw = make_divw(w);
  int slot = -1;                      // the seed slot this lane is growing (-1: idle)
  int* cur = nullptr;
  int n = 0, i = 0, nxt = 0, bx0 = 0, by0 = 0, bx1 = 0, by1 = 0;
  uint32_t stamp = 0;
  float sumdx = 0.f, sumdy = 0.f;
  double ra = 0;
  bool more = true;                   // the wave may still have unassigned seeds (warp-uniform)
#pragma unroll 1
  while (true) {
    // ---- idle lanes take the next seed slots, in order ----
    const uint32_t idlem = __ballot_sync(FULL, slot < 0);
    if (idlem && more) {
      int base = 0;
      if (lane == 0) base = atomicAdd(&S.anext, __popc(idlem));
      base = __shfl_sync(FULL, base, 0);
      if (base + __popc(idlem) >= nsel) more = false;
      if (slot < 0) {
        const int sidx = base + __popc(idlem & lt);
        if (sidx < nsel) {
          slot = sidx;
          const int seed = S.sel[slot];
          stamp = (wave << 11) | ((uint32_t)(K - 1 - slot) << 1);
          cur = T.reg_spec + (size_t)slot * cap;
          int sy, sx;
          divmod_w(dw, seed, sy, sx);
          bx0 = bx1 = sx; by0 = by1 = sy;
          n = 0; i = 0;
          if (ta <= 0) {
            // no per-lane growth at all: hand the untouched seed to phase B
            SlotCtx& c = ctx[slot];
            c.n = 0; c.i = 0; c.sumdx = 0.f; c.sumdy = 0.f; c.ra = 0.0; c.bx0 = sx; c.by0 = sy; c.bx1 = sx; c.by1 = sy; c.flags = -1;
            slot = -1;
          } else if (cap < 1 || ld_state(T.state + seed) > stamp) {
            SlotCtx& c = ctx[slot];                             // dead on arrival: an earlier seed of the wave owns the pixel
            c.n1 = 0; c.n2o = 0; c.nf = 0; c.foff = 0; c.flags = 0; c.bx0 = sx; c.by0 = sy; c.bx1 = sx; c.by1 = sy;
            slot = -1;
          } else {
            claim_max(&T.state[seed], stamp);
            nxt = xy_pack(sx, sy);
            cur[0] = nxt; n = 1;
            ra = T.px.ang(seed);
            double sn, cs;
            sincos_call(ra, &sn, &cs);
            sumdx = (float)cs; sumdy = (float)sn;
          }
        }
      }
    }
    if (!__any_sync(FULL, slot >= 0)) {
      if (!more) break;
      continue;
    }
    // ---- one expansion per active lane ----
    int settled = 0;                  // 1: finished, 2: aborted, 3: hand over to phase B
    if (slot >= 0) {
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
      double ang[9];
#pragma unroll
      for (int k = 0; k < 9; k++) {
        if (k == 4) continue;
        const int q = rofs[k / 3] + cofs[k % 3];
        ld_rec_angle(T.px.p + q, st[k], ang[k]);
      }
      if (i + 1 < n_start) nxt = cur[i + 1];
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
        if (((foreign >> k) & 1u) || n >= cap) { settled = 2; break; }
        const int q = qy * w + qx;
        claim_max(&T.state[q], stamp);
        const float2 csq = T.px.cs(q);
        const int qp = xy_pack(qx, qy);
        if (n == i + 1) nxt = qp;
        cur[n++] = qp;
        bx0 = min(bx0, qx); bx1 = max(bx1, qx);
        by0 = min(by0, qy); by1 = max(by1, qy);
        sumdx = __fadd_rn(sumdx, csq.x);
        sumdy = __fadd_rn(sumdy, csq.y);
        ra = (double)fast_atan2_deg(sumdy, sumdx) * kDegToRad;
        cand &= ~((2u << k) - 1u);
      }
      i++;
      if (!settled) {
        if (i >= n) settled = (n < T.min_reg) ? 1 : 3;          // a finished region that is big enough still needs its rectangle
        else if (i >= ta) settled = 3;
      }
      if (settled) {
        SlotCtx& c = ctx[slot];
        c.bx0 = bx0; c.by0 = by0; c.bx1 = bx1; c.by1 = by1;
        if (settled == 3) { c.n = n; c.i = i; c.sumdx = sumdx; c.sumdy = sumdy; c.ra = ra; c.flags = i < n ? -1 : -2; }
        else { c.n1 = n; c.n2o = 0; c.nf = settled == 1 ? n : 0; c.foff = 0; c.flags = settled == 1 ? kCtxOk : 0; }
        slot = -1;
      }
    }
  }
}

template <int NW>
__device__ void grow_task_block2(const Task& T, SlotCtx* const ctx, const int ta, Block2Shared<NW>& S) {
  const uint32_t FULL = 0xffffffffu;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int K = 32 * NW;
  const uint32_t lt = (1u << lane) - 1u;
  const int cap = (32 * T.lane_cap) / K;
  const DivW dw = make_divw(T.w);
  int* const my_reg = T.reg_spec + (size_t)tid * cap;
  if (tid == 0) { S.cursor = 0; S.qn = 0; S.qhead = 0; S.anext = 0; }
  int npend = 0;          // rectangles appended so far (uniform across the CTA)
  uint32_t wave = 0;
  long long t_sel = 0, t_spec = 0, t_commit = 0, t_redo = 0, n_redo = 0, n_seed = 0, n_b = 0;
  __syncthreads();
  while (true) {
    long long c0 = clock64();
    // ---- warp 0 selects the next (up to) K free seeds in order (eight list blocks in flight per trip) ----
    if (warp == 0) {
      int cursor = S.cursor, nsel = 0;
      constexpr int kSelBlocks = 8;
      while (nsel < K && cursor < T.ndef) {
        int p[kSelBlocks];
        bool fr[kSelBlocks];
#pragma unroll
        for (int jb = 0; jb < kSelBlocks; jb++) {
          const int idx = cursor + jb * 32 + lane;
          p[jb] = idx < T.ndef ? (int)T.order[idx] : -1;
        }
#pragma unroll
        for (int jb = 0; jb < kSelBlocks; jb++) fr[jb] = !(ld_state(T.state + (p[jb] >= 0 ? p[jb] : 0)) & kUsed) && p[jb] >= 0;
        uint32_t fm[kSelBlocks];
#pragma unroll
        for (int jb = 0; jb < kSelBlocks; jb++) fm[jb] = __ballot_sync(FULL, fr[jb]);
        bool stop = false;
#pragma unroll
        for (int jb = 0; jb < kSelBlocks; jb++) {
          if (stop) continue;
          const uint32_t m = fm[jb];
          const int c = __popc(m);
          const int take = min(c, K - nsel);
          const int rank = __popc(m & lt);
          if (fr[jb] && rank < take) S.sel[nsel + rank] = p[jb];
          nsel += take;
          if (c > take) { cursor += nth_set_bit(m, take - 1) + 1; stop = true; }     // stop right after the last seed taken
          else { cursor += 32; stop = nsel >= K || cursor >= T.ndef; }
        }
      }
      if (lane == 0) { S.cursor = cursor; S.nsel = nsel; S.qn = 0; S.qhead = 0; S.anext = 0; }
    }
    __syncthreads();
    const int nsel = S.nsel;
    if (nsel == 0) break;
    wave++;
    n_seed += nsel;
    const bool have = tid < nsel;
    const int my_seed = have ? S.sel[tid] : 0;
    long long c1 = clock64(); t_sel += c1 - c0;
    const uint32_t stamp = (wave << 11) | ((uint32_t)(K - 1 - tid) << 1);   // bit0 = phase, 10 bits of seed priority
    // ---- phase A: lanes grow seeds for a few expansions, taking new seeds as they settle ----
    phase_a<NW>(T, ctx, S, nsel, cap, ta, wave);
    const long long ca1 = clock64();
    __syncthreads();                                                                       // (A) every seed has a context
    if (T.prof_detail == 1 && tid == 0) { T.prof[8] += ca1 - c1; T.prof[15] += clock64() - ca1; }
    // ---- the seeds that need phase B, in slot order ----
    {
      // unfinished growths (the long regions) first, then the finished ones that only want their rectangle: the queue is
      // handed out in this order, long jobs should not start last
      const int fl = have ? ctx[tid].flags : 0;
      const uint32_t bm1 = __ballot_sync(FULL, fl == -1), bm2 = __ballot_sync(FULL, fl == -2);
      if (lane == 0) { S.goodm[warp] = bm1; S.rectm[warp] = bm2; }
      __syncthreads();
      int base1 = 0, base2 = 0, tot1 = 0, tot2 = 0;
      for (int v = 0; v < NW; v++) {
        const int c1 = __popc(S.goodm[v]), c2 = __popc(S.rectm[v]);
        if (v < warp) { base1 += c1; base2 += c2; }
        tot1 += c1; tot2 += c2;
      }
      if (fl == -1) S.q[base1 + __popc(bm1 & lt)] = tid;
      if (fl == -2) S.q[tot1 + base2 + __popc(bm2 & lt)] = tid;
      if (tid == 0) S.qn = tot1 + tot2;
    }
    __syncthreads();                                                                       // (A') queue complete
    // ---- phase B: warps take the queued seeds over, one at a time ----
    {
      const int qn = S.qn;
      const long long cb0 = clock64();
      while (true) {
        int k = 0;
        if (lane == 0) k = atomicAdd(&S.qhead, 1);
        k = __shfl_sync(FULL, k, 0);
        if (k >= qn) break;
        n_b++;
        const int slot = S.q[k];
        SlotCtx& c = ctx[slot];
        const uint32_t st_slot = (wave << 11) | ((uint32_t)(K - 1 - slot) << 1);
        SeedResult R;
        int ssy, ssx;
        divmod_w(dw, S.sel[slot], ssy, ssx);
        seed_pipeline_coop(T, true, S.sb[warp], S.sel[slot], ssx, ssy, T.reg_spec + (size_t)slot * cap, cap, st_slot, c.n, c.i, c.sumdx, c.sumdy, c.ra, c.bx0,
                           c.by0, c.bx1, c.by1, R);
        if (lane == 0) {
          c.n1 = R.n1; c.n2o = R.n2_orig; c.nf = R.nf; c.foff = R.foff;
          c.flags = (R.ok ? kCtxOk : 0) | (R.has_rect ? kCtxRect : 0);
          c.bx0 = R.bx0; c.by0 = R.by0; c.bx1 = R.bx1; c.by1 = R.by1;
          if (R.has_rect) c.rec = R.rec;
        }
      }
      if (T.prof_detail == 1 && tid == 0) T.prof[9] += clock64() - cb0;
    }
    const long long cb1 = clock64();
    __syncthreads();                                                                       // (B) every seed of the wave has a result
    if (T.prof_detail == 1 && tid == 0) T.prof[10] += clock64() - cb1;
    int r_ok = 0, r_n1 = 0, r_n2o = 0, r_nf = 0, r_foff = 0, r_rect = 0, r_bx0 = 0, r_by0 = 0, r_bx1 = 0, r_by1 = 0;
    if (have) {
      const SlotCtx& c = ctx[tid];
      r_ok = c.flags & kCtxOk; r_rect = (c.flags & kCtxRect) ? 1 : 0;
      r_n1 = c.n1; r_n2o = c.n2o; r_nf = c.nf; r_foff = c.foff;
      r_bx0 = c.bx0; r_by0 = c.by0; r_bx1 = c.bx1; r_by1 = c.by1;
    }
    long long c2 = clock64(); t_spec += c2 - c1;
    // ---- validate + commit in seed order; the first doubtful seed is re-run sequentially by its warp ----
    int verdict = 0;        // cached verdict: 0 unknown, 1 good (invalidated only by a re-run that wrote near me), 2 dead, 3 must re-run
    int settled = 0;        // seeds [0, settled) of the wave are committed, dropped or re-run (uniform across the CTA)
    while (true) {
      const bool mine = tid >= settled && tid < nsel;
      bool dead = false, good = false;
      {
        const bool need = mine && verdict == 0;
        bool d0 = false;
        if (need) d0 = (ld_state(T.state + my_seed) & kUsed) != 0;
        const bool own = owns_all_coop_xy(T, need && !d0 && r_ok, my_reg, r_n1 + r_n2o, stamp >> 1);
        if (need) verdict = d0 ? 2 : (own ? 1 : 3);
      }
      uint32_t sv_seed = kUsed;
      if (mine) {
        if (verdict == 3) {
          // lost a pixel (or never finished): that is permanent; only "my seed was taken meanwhile" can still change
          sv_seed = ld_state(T.state + my_seed);
          dead = (sv_seed & kUsed) != 0;
          if (dead) verdict = 2;
        }
        dead = verdict == 2; good = verdict == 1;
      }
      // a seed whose own pixel carries the stamp of an earlier seed o of this wave is taken as soon as o commits -- provided o
      // commits every pixel it stamped (no refinement: final list = first region).  Seeing that here, instead of one loop trip
      // later, saves the trip: three of four doubtful seeds are of this kind (profiles/r2_rerun_statistics.txt)
      const uint32_t dm = __ballot_sync(FULL, dead), gm = __ballot_sync(FULL, good);
      const uint32_t rm = __ballot_sync(FULL, good && r_rect);
      const uint32_t sm = __ballot_sync(FULL, good && r_foff == 0 && r_n2o == 0 && r_nf == r_n1);
      if (lane == 0) { S.deadm[warp] = dm; S.goodm[warp] = gm; S.rectm[warp] = rm; S.simplem[warp] = sm; }
      __syncthreads();                                                                     // (1) verdicts published
      {
        bool cd = false;
        if (mine && verdict == 3 && !(sv_seed & kUsed) && (sv_seed >> 11) == wave && !(sv_seed & 1u)) {
          const int o = K - 1 - (int)((sv_seed >> 1) & 0x3ffu);
          cd = o >= 0 && o < tid && ((S.goodm[o >> 5] & S.simplem[o >> 5]) >> (o & 31)) & 1u;
        }
        const uint32_t cm = __ballot_sync(FULL, cd);
        if (lane == 0) S.cdeadm[warp] = cm;
      }
      __syncthreads();                                                                     // (1b) conditionally dead seeds published
      // first unsettled seed that is neither dead nor provably good; rectangles of the good seeds before it, in seed order
      int ks = nsel, before = 0, total = 0;
      for (int v = settled >> 5; v < NW; v++) {
        const int first = v * 32;
        if (first >= nsel) break;
        uint32_t pm = FULL;
        if (settled > first) pm &= ~((1u << (settled - first)) - 1u);
        if (nsel < first + 32) pm &= (1u << (nsel - first)) - 1u;
        const uint32_t bad = pm & ~S.deadm[v] & ~S.goodm[v] & ~S.cdeadm[v];
        if (bad) { ks = first + __ffs(bad) - 1; pm &= (1u << (ks - first)) - 1u; }
        const uint32_t rr = S.rectm[v] & pm;
        total += __popc(rr);
        if (v < warp) before += __popc(rr);
        else if (v == warp) before += __popc(rr & lt);
        if (bad) break;
      }
      const bool do_commit = mine && good && tid < ks;
      mark_used_coop_xy(T, do_commit, my_reg + r_foff, r_nf);
      if (do_commit && r_rect) append_rect(T, npend + before, ctx[tid].rec, (int)((wave << 11) | (tid << 1) | 0), my_seed, r_nf);
      npend += total;
      settled = ks;
      if (ks >= nsel) break;
      n_redo++;
      __syncthreads();                                                                     // (2) commits visible to the re-run
      long long cr = clock64();
      if (warp == (ks >> 5)) {
        // the whole warp of ks re-runs that seed with exact sequential semantics.  The commits just made may have taken the
        // seed: then the sequential algorithm skips it
        const int ks_lane = ks & 31;
        const long long q0 = clock64();
        const int seed_k = __shfl_sync(FULL, my_seed, ks_lane);
        const bool taken = (ld_state(T.state + seed_k) & kUsed) != 0;
        int has = 0;
        if (T.prof_detail == 2 && lane == ks_lane) { atomicAdd((unsigned long long*)&T.prof[8], (unsigned long long)(clock64() - q0)); if (taken) atomicAdd((unsigned long long*)&T.prof[15], 1ull); }
        if (!taken) {
          SeedResult Q;
          int ssy, ssx;
          divmod_w(dw, seed_k, ssy, ssx);
          const long long q1 = clock64();
          seed_pipeline_coop(T, false, S.sb[warp], seed_k, ssx, ssy, T.reg_serial, T.npx, 0u, 0, 0, 0.f, 0.f, 0.0, 0, 0, 0, 0, Q);
          if (T.prof_detail == 2 && lane == ks_lane) atomicAdd((unsigned long long*)&T.prof[9], (unsigned long long)(clock64() - q1));
          has = Q.has_rect;
          if (T.prof_detail == 2 && lane == ks_lane) {          // re-run statistics: pixels grown first / re-grown, rectangles
            atomicAdd((unsigned long long*)&T.prof[11], (unsigned long long)Q.n1);
            atomicAdd((unsigned long long*)&T.prof[12], (unsigned long long)Q.n2_orig + (unsigned long long)Q.nf);
            atomicAdd((unsigned long long*)&T.prof[13], (unsigned long long)Q.has_rect);
            atomicMax((unsigned long long*)&T.prof[14], (unsigned long long)Q.n1);
          }
          if (lane == ks_lane) {
            if (has) append_rect(T, npend, Q.rec, (int)((wave << 11) | (tid << 1) | 1), my_seed, Q.nf);
            S.rb[0] = Q.bx0; S.rb[1] = Q.by0; S.rb[2] = Q.bx1; S.rb[3] = Q.by1;
          }
        } else if (lane == ks_lane) { S.rb[0] = 1; S.rb[2] = 0; }
        if (lane == ks_lane) S.has = has;
      }
      __syncthreads();                                                                     // (3) re-run finished
      npend += S.has;
      settled = ks + 1;
      // the re-run wrote `used` inside its bounding box only: cached "good" verdicts elsewhere stay valid
      if (verdict == 1 && !(S.rb[0] > r_bx1 || S.rb[2] < r_bx0 || S.rb[1] > r_by1 || S.rb[3] < r_by0)) verdict = 0;
      t_redo += clock64() - cr;
      if (settled >= nsel) break;
    }
    t_commit += clock64() - c2;
    __syncthreads();
  }
  if (tid == 0) *T.npend = min(npend, T.pend_cap);
  if (tid == 0 && T.prof) {
    T.prof[0] = t_sel; T.prof[1] = t_spec; T.prof[2] = t_commit - t_redo; T.prof[3] = t_redo; T.prof[4] = wave; T.prof[5] = n_redo;
    T.prof[6] = n_b; T.prof[7] = n_seed;
  }
}

}  // namespace lsd
}  // namespace sdpl
