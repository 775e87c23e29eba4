// oracle/orb_oracle.cpp -- ORACLE (test infrastructure): CPU restatement of the reference ORB extractor,
// argyrissm/SDPL-SLAM src/ORBextractor.cc.  Single thread, strict IEEE (build with -ffp-contract=off).
//
// Oracle decisions where the reference leaves behaviour open:
//  (i)  DistributeOctTree sorts (size, node pointer) (ORBextractor.cc:673-674); pointer order is heap dependent,
//       the oracle orders equal sizes by node creation sequence (later created sorts higher).
//  (ii) no FMA contraction anywhere.
//  (iii) rBRIEF descriptors ARE computed (the call at ORBextractor.cc:1091 is commented out in the reference and
//       its output matrix is uninitialised); they follow computeOrbDescriptor (ORBextractor.cc:97-136) on the
//       7x7 sigma-2 blurred level, with cos/sin evaluated in double and rounded to float.
#include "oracle_internal.h"
#include <cmath>
#include <cstring>
#include <list>
#include <algorithm>

namespace orc {

static const int kPattern[1024] = {
#include "../include/sdpl_orb_pattern.inc"
};

void build_padded_pyramid(const uint8_t* img, int w, int h, int stride, const std::vector<float>& inv_scale,
                          std::vector<PaddedLevel>& levels) {
  // ORBextractor::ComputePyramid, src/ORBextractor.cc:1112-1137 (and LSDDetector_custom.cpp:76-109)
  int n = (int)inv_scale.size();
  levels.resize(n);
  for (int l = 0; l < n; l++) {
    float s = inv_scale[l];
    PaddedLevel& L = levels[l];
    L.w = cv_round((float)w * s);
    L.h = cv_round((float)h * s);
    L.buf.assign((size_t)(L.w + 38) * (L.h + 38), 0);
    if (l == 0) {
      border_reflect101_u8(img, w, h, stride, L.buf.data(), 19, L.stride());
    } else {
      const PaddedLevel& P = levels[l - 1];
      std::vector<uint8_t> tmp((size_t)L.w * L.h);
      resize_linear_u8(P.roi(), P.w, P.h, P.stride(), tmp.data(), L.w, L.h, L.w);
      border_reflect101_u8(tmp.data(), L.w, L.h, L.w, L.buf.data(), 19, L.stride());
    }
  }
}

struct Cand { float x, y; int resp; };

struct QNode {
  int x0, x1, y0, y1;           // [x0,x1) x [y0,y1)  (UL.x, UR.x, UL.y, BL.y)
  std::vector<int> keys;        // candidate indices, inherited order
  bool no_more = false;
  int seq = 0;                  // creation sequence (tie-break, oracle decision i)
  std::list<QNode>::iterator self;
};

// ExtractorNode::DivideNode, src/ORBextractor.cc:470-526
static void divide(const QNode& p, const std::vector<Cand>& c, QNode out[4]) {
  int halfX = (int)std::ceil((float)(p.x1 - p.x0) / 2);
  int halfY = (int)std::ceil((float)(p.y1 - p.y0) / 2);
  int xm = p.x0 + halfX, ym = p.y0 + halfY;
  out[0] = QNode{p.x0, xm, p.y0, ym};
  out[1] = QNode{xm, p.x1, p.y0, ym};
  out[2] = QNode{p.x0, xm, ym, p.y1};
  out[3] = QNode{xm, p.x1, ym, p.y1};
  for (int k : p.keys) {
    const Cand& kp = c[k];
    if (kp.x < xm) {
      if (kp.y < ym) out[0].keys.push_back(k); else out[2].keys.push_back(k);
    } else if (kp.y < ym) out[1].keys.push_back(k);
    else out[3].keys.push_back(k);
  }
  for (int i = 0; i < 4; i++) if (out[i].keys.size() == 1) out[i].no_more = true;
}

// ORBextractor::DistributeOctTree, src/ORBextractor.cc:528-752.  Returns selected candidate indices in lNodes order.
static int distribute(const std::vector<Cand>& c, int minX, int maxX, int minY, int maxY, int N, std::vector<int>& sel) {
  sel.clear();
  int nIni = (int)std::round((float)(maxX - minX) / (maxY - minY));
  if (nIni < 1) return -1;  // reference divides by zero here
  float hX = (float)(maxX - minX) / nIni;
  std::list<QNode> nodes;
  std::vector<QNode*> roots(nIni);
  int seq = 0;
  for (int i = 0; i < nIni; i++) {
    QNode n{(int)(hX * (float)i), (int)(hX * (float)(i + 1)), 0, maxY - minY};
    n.seq = seq++;
    nodes.push_back(n);
    roots[i] = &nodes.back();
  }
  for (size_t i = 0; i < c.size(); i++) {
    int r = (int)(c[i].x / hX);
    if (r < 0 || r >= nIni) return -2;
    roots[r]->keys.push_back((int)i);
  }
  for (auto it = nodes.begin(); it != nodes.end();) {
    if (it->keys.size() == 1) { it->no_more = true; ++it; }
    else if (it->keys.empty()) it = nodes.erase(it);
    else ++it;
  }
  bool finish = false;
  typedef std::pair<int, QNode*> SP;
  std::vector<SP> expand;
  auto by_size_seq = [](const SP& a, const SP& b) {
    if (a.first != b.first) return a.first < b.first;
    return a.second->seq < b.second->seq;
  };
  auto push_children = [&](QNode ch[4], int* n_expand) {
    for (int i = 0; i < 4; i++) {
      if (ch[i].keys.empty()) continue;
      ch[i].seq = seq++;
      nodes.push_front(ch[i]);
      if (ch[i].keys.size() > 1) {
        if (n_expand) (*n_expand)++;
        expand.push_back(SP((int)ch[i].keys.size(), &nodes.front()));
        nodes.front().self = nodes.begin();
      }
    }
  };
  while (!finish) {
    int prev = (int)nodes.size();
    int n_expand = 0;
    expand.clear();
    for (auto it = nodes.begin(); it != nodes.end();) {
      if (it->no_more) { ++it; continue; }
      QNode ch[4];
      divide(*it, c, ch);
      push_children(ch, &n_expand);
      it = nodes.erase(it);
    }
    if ((int)nodes.size() >= N || (int)nodes.size() == prev) {
      finish = true;
    } else if ((int)nodes.size() + n_expand * 3 > N) {
      while (!finish) {
        prev = (int)nodes.size();
        std::vector<SP> todo = expand;
        expand.clear();
        std::sort(todo.begin(), todo.end(), by_size_seq);
        for (int j = (int)todo.size() - 1; j >= 0; j--) {
          QNode ch[4];
          divide(*todo[j].second, c, ch);
          push_children(ch, nullptr);
          nodes.erase(todo[j].second->self);
          if ((int)nodes.size() >= N) break;
        }
        if ((int)nodes.size() >= N || (int)nodes.size() == prev) finish = true;
      }
    }
  }
  for (auto& n : nodes) {
    int best = n.keys[0];
    for (size_t k = 1; k < n.keys.size(); k++)
      if (c[n.keys[k]].resp > c[best].resp) best = n.keys[k];
    sel.push_back(best);
  }
  return (int)sel.size();
}

}  // namespace orc

using namespace orc;

struct orc_orb {
  int nfeatures, nlevels, ini_th, min_th;
  float scale_f;
  std::vector<float> sf, isf, s2, is2;
  std::vector<int> quota;
  int umax[16];
  std::vector<PaddedLevel> levels;
  std::vector<std::vector<uint8_t>> blurred;
  std::vector<std::vector<Cand>> cands;
  std::vector<int> counts;
};

extern "C" {

// ORBextractor::ORBextractor, src/ORBextractor.cc:399-459
orc_orb* orc_orb_create(int nfeatures, float scale, int nlevels, int ini_th, int min_th) {
  if (nlevels < 1 || nfeatures < 1) return nullptr;
  orc_orb* o = new orc_orb;
  o->nfeatures = nfeatures; o->nlevels = nlevels; o->ini_th = ini_th; o->min_th = min_th; o->scale_f = scale;
  double scaleFactor = (double)scale;  // member is a double in the reference header (include/ORBextractor.h:85)
  o->sf.resize(nlevels); o->s2.resize(nlevels); o->isf.resize(nlevels); o->is2.resize(nlevels);
  o->sf[0] = 1.f; o->s2[0] = 1.f;
  for (int i = 1; i < nlevels; i++) {
    o->sf[i] = (float)(o->sf[i - 1] * scaleFactor);
    o->s2[i] = o->sf[i] * o->sf[i];
  }
  for (int i = 0; i < nlevels; i++) { o->isf[i] = 1.0f / o->sf[i]; o->is2[i] = 1.0f / o->s2[i]; }
  o->quota.resize(nlevels);
  float factor = (float)(1.0f / scaleFactor);
  float desired = nfeatures * (1 - factor) / (1 - (float)std::pow((double)factor, (double)nlevels));
  int sum = 0;
  for (int l = 0; l < nlevels - 1; l++) {
    o->quota[l] = cv_round(desired);
    sum += o->quota[l];
    desired *= factor;
  }
  o->quota[nlevels - 1] = std::max(nfeatures - sum, 0);
  // umax, ORBextractor.cc:443-458
  const int HP = 15;
  int vmax = cv_floor(HP * std::sqrt(2.f) / 2 + 1), vmin = cv_ceil(HP * std::sqrt(2.f) / 2);
  const double hp2 = HP * HP;
  for (int v = 0; v <= vmax; ++v) o->umax[v] = cv_round(std::sqrt(hp2 - v * v));
  for (int v = HP, v0 = 0; v >= vmin; --v) {
    while (o->umax[v0] == o->umax[v0 + 1]) ++v0;
    o->umax[v] = v0;
    ++v0;
  }
  return o;
}
void orc_orb_destroy(orc_orb* o) { delete o; }

void orc_orb_tables(const orc_orb* o, float* sf, float* isf, float* s2, float* is2, int* quota, int* umax) {
  for (int i = 0; i < o->nlevels; i++) {
    if (sf) sf[i] = o->sf[i];
    if (isf) isf[i] = o->isf[i];
    if (s2) s2[i] = o->s2[i];
    if (is2) is2[i] = o->is2[i];
    if (quota) quota[i] = o->quota[i];
  }
  if (umax) memcpy(umax, o->umax, sizeof(o->umax));
}

// IC_Angle, src/ORBextractor.cc:66-93
static float ic_angle(const uint8_t* center, int step, const int* umax) {
  int m01 = 0, m10 = 0;
  for (int u = -15; u <= 15; ++u) m10 += u * center[u];
  for (int v = 1; v <= 15; ++v) {
    int vs = 0, d = umax[v];
    for (int u = -d; u <= d; ++u) {
      int p = center[u + v * step], m = center[u - v * step];
      vs += p - m;
      m10 += u * (p + m);
    }
    m01 += v * vs;
  }
  return fast_atan2((float)m01, (float)m10);
}

// computeOrbDescriptor, src/ORBextractor.cc:97-136
static void orb_descriptor(float kp_angle, const uint8_t* center, int step, uint8_t* desc) {
  const float factorPI = (float)(3.141592653589793238462643383279502884 / 180.f);
  float angle = kp_angle * factorPI;
  float a = (float)std::cos((double)angle), b = (float)std::sin((double)angle);
  const int* p = kPattern;
  for (int i = 0; i < 32; i++, p += 32) {
    int val = 0;
    for (int t = 0; t < 8; t++) {
      const int* q = p + 4 * t;  // 4 ints per test: x0,y0,x1,y1
      volatile float xb0 = q[0] * b, ya0 = q[1] * a, xa0 = q[0] * a, yb0 = q[1] * b;
      volatile float xb1 = q[2] * b, ya1 = q[3] * a, xa1 = q[2] * a, yb1 = q[3] * b;
      volatile float r0 = xb0 + ya0, c0 = xa0 - yb0, r1 = xb1 + ya1, c1 = xa1 - yb1;
      int t0 = center[cv_round(r0) * step + cv_round(c0)];
      int t1 = center[cv_round(r1) * step + cv_round(c1)];
      val |= (t0 < t1) << t;
    }
    desc[i] = (uint8_t)val;
  }
}

// ORBextractor::operator(), src/ORBextractor.cc:1035-1110 (+ ComputeKeyPointsOctTree :754-842)
int orc_orb_extract(orc_orb* o, const uint8_t* img, int w, int h, int stride, orc_keypoint* kps, uint8_t* desc, int cap) {
  if (!img || w <= 0 || h <= 0) return 0;  // empty image: silent return (:1038)
  const int nl = o->nlevels;
  build_padded_pyramid(img, w, h, stride, o->isf, o->levels);
  o->cands.assign(nl, {});
  o->blurred.assign(nl, {});
  o->counts.assign(nl, 0);
  std::vector<std::vector<orc_keypoint>> all(nl);
  const float W = 30;
  for (int level = 0; level < nl; level++) {
    const PaddedLevel& L = o->levels[level];
    const int minBX = 16, minBY = 16, maxBX = L.w - 16, maxBY = L.h - 16;
    const float width = (float)(maxBX - minBX), height = (float)(maxBY - minBY);
    const int nCols = (int)(width / W), nRows = (int)(height / W);
    std::vector<Cand>& cand = o->cands[level];
    if (nCols < 1 || nRows < 1 || width <= 0 || height <= 0) continue;  // level too small: reference divides by 0
    const int wCell = (int)std::ceil(width / nCols), hCell = (int)std::ceil(height / nRows);
    std::vector<int> xs, ys, sc;
    for (int i = 0; i < nRows; i++) {
      const float iniY = (float)(minBY + i * hCell);
      float maxY = iniY + hCell + 6;
      if (iniY >= maxBY - 3) continue;
      if (maxY > maxBY) maxY = (float)maxBY;
      for (int j = 0; j < nCols; j++) {
        const float iniX = (float)(minBX + j * wCell);
        float maxX = iniX + wCell + 6;
        if (iniX >= maxBX - 6) continue;
        if (maxX > maxBX) maxX = (float)maxBX;
        int x0 = (int)iniX, y0 = (int)iniY, cw = (int)maxX - x0, chh = (int)maxY - y0;
        const uint8_t* cell = L.roi() + (size_t)y0 * L.stride() + x0;
        int n = fast9_nms(cell, cw, chh, L.stride(), o->ini_th, xs, ys, sc);
        if (n == 0) n = fast9_nms(cell, cw, chh, L.stride(), o->min_th, xs, ys, sc);
        for (int k = 0; k < n; k++) cand.push_back(Cand{(float)xs[k] + j * wCell, (float)ys[k] + i * hCell, sc[k]});
      }
    }
    std::vector<int> sel;
    int r = distribute(cand, minBX, maxBX, minBY, maxBY, o->quota[level], sel);
    if (r < 0) return -1;
    const int patch = (int)(31 * o->sf[level]);
    for (int idx : sel) {
      orc_keypoint k;
      k.x = cand[idx].x + minBX; k.y = cand[idx].y + minBY;
      k.size = (float)patch; k.angle = -1; k.response = (float)cand[idx].resp; k.octave = level; k.class_id = -1;
      all[level].push_back(k);
    }
  }
  for (int level = 0; level < nl; level++) {
    const PaddedLevel& L = o->levels[level];
    for (auto& k : all[level])
      k.angle = ic_angle(L.roi() + (size_t)cv_round(k.y) * L.stride() + cv_round(k.x), L.stride(), o->umax);
  }
  int total = 0;
  for (int level = 0; level < nl; level++) {
    o->counts[level] = (int)all[level].size();
    if (all[level].empty()) continue;
    const PaddedLevel& L = o->levels[level];
    o->blurred[level].resize((size_t)L.w * L.h);
    gaussian_blur_u8(L.roi(), L.w, L.h, L.stride(), o->blurred[level].data(), L.w, 0);
    for (auto& k : all[level]) {
      if (total < cap) {
        if (desc)
          orb_descriptor(k.angle, o->blurred[level].data() + (size_t)cv_round(k.y) * L.w + cv_round(k.x), L.w,
                         desc + (size_t)total * 32);
        orc_keypoint out = k;
        if (level != 0) { out.x = k.x * o->sf[level]; out.y = k.y * o->sf[level]; }
        kps[total] = out;
      }
      total++;
    }
  }
  return total;
}

void orc_orb_level_size(const orc_orb* o, int level, int* w, int* h) { *w = o->levels[level].w; *h = o->levels[level].h; }
const uint8_t* orc_orb_level_padded(const orc_orb* o, int level) { return o->levels[level].buf.data(); }
const uint8_t* orc_orb_level_blurred(const orc_orb* o, int level) {
  return o->blurred[level].empty() ? nullptr : o->blurred[level].data();
}
int orc_orb_level_candidates(const orc_orb* o, int level, int* xs, int* ys, int* resp, int cap) {
  const auto& c = o->cands[level];
  for (size_t i = 0; i < c.size() && (int)i < cap; i++) { xs[i] = (int)c[i].x; ys[i] = (int)c[i].y; resp[i] = c[i].resp; }
  return (int)c.size();
}
int orc_orb_level_count(const orc_orb* o, int level) { return o->counts[level]; }
}
