// oracle/match_oracle.cpp -- ORACLE (test infrastructure): Hamming distance and brute-force matching.
// Metric: cv::line_descriptor::match(P,Q,32), 3rdparty/line_descriptor/src/bitops_custom.hpp:86-99
// (= DBoW2 FORB::distance, 3rdparty/DBoW2/DBoW2/FORB.cpp:81-101).  The reference has no live descriptor
// matcher; semantics are oracle decision (iv): exact brute force, best/second by strict '<'
// scanning train indices in ascending order (ties -> lowest train index), i.e. cv2.BFMatcher(NORM_HAMMING).knnMatch(k=2);
// DMatch layout/field meaning as BinaryDescriptorMatcher::knnMatch (binary_descriptor_matcher.cpp:258-335):
// distance = (float)hamming, imgIdx = 0.
#include "oracle_internal.h"
#include <cstring>
#include <algorithm>

extern "C" {
int orc_hamming256(const uint8_t* a, const uint8_t* b) {
  int d = 0;
  for (int i = 0; i < 4; i++) {
    uint64_t x, y;
    memcpy(&x, a + 8 * i, 8);
    memcpy(&y, b + 8 * i, 8);
    d += __builtin_popcountll(x ^ y);
  }
  return d;
}

void orc_match_knn2(const uint8_t* q, int nq, const uint8_t* t, int nt, orc_dmatch* best, orc_dmatch* second) {
  for (int i = 0; i < nq; i++) {
    int d1 = 257, d2 = 257, i1 = -1, i2 = -1;
    for (int j = 0; j < nt; j++) {
      int d = orc_hamming256(q + 32 * (size_t)i, t + 32 * (size_t)j);
      if (d < d1) { d2 = d1; i2 = i1; d1 = d; i1 = j; }
      else if (d < d2) { d2 = d; i2 = j; }
    }
    best[i] = orc_dmatch{i, i1, 0, (float)d1};
    second[i] = orc_dmatch{i, i2, 0, (float)d2};
  }
}

int orc_match_ratio(const uint8_t* q, int nq, const uint8_t* t, int nt, float ratio, int max_dist, orc_dmatch* out) {
  std::vector<orc_dmatch> b(nq), s(nq);
  orc_match_knn2(q, nq, t, nt, b.data(), s.data());
  int acc = 0;
  for (int i = 0; i < nq; i++) {
    out[i] = b[i];
    bool ok = b[i].train >= 0 && b[i].distance <= (float)max_dist && b[i].distance < ratio * s[i].distance;
    if (!ok) out[i].train = -1; else acc++;
  }
  return acc;
}

void orc_match_radius(const uint8_t* q, int nq, const uint8_t* t, int nt, int radius, int k, int* counts, orc_dmatch* out) {
  std::vector<std::pair<int, int>> v;
  for (int i = 0; i < nq; i++) {
    v.clear();
    for (int j = 0; j < nt; j++) {
      int d = orc_hamming256(q + 32 * (size_t)i, t + 32 * (size_t)j);
      if (d <= radius) v.push_back({d, j});
    }
    std::sort(v.begin(), v.end());
    counts[i] = (int)v.size();
    for (int r = 0; r < k; r++) {
      if (r < (int)v.size()) out[(size_t)i * k + r] = orc_dmatch{i, v[r].second, 0, (float)v[r].first};
      else out[(size_t)i * k + r] = orc_dmatch{i, -1, 0, 257.f};
    }
  }
}
}
