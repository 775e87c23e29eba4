// oracle/refshim/ref_api.cpp -- ORACLE test infrastructure.  C entry points over the COMPILED reference classes
// (SDPL_SLAM::ORBextractor, SDPL_SLAM::Lineextractor, cv::line_descriptor::BinaryDescriptor / BinaryDescriptorMatcher /
// match()) so that tests can pin oracle/*.cpp against the reference's own code (tests/test_oracle_vs_ref.py).
#include "ORBextractor.h"
#include "Lineextractor.h"
#include "../oracle.h"
// bitops_custom.hpp (match()) and the other private headers of line_descriptor
#include "precomp_custom.hpp"

namespace {
// DistributeOctTree, the quota table and umax are protected: a derived class may publish them.
struct OrbProbe : public SDPL_SLAM::ORBextractor {
  OrbProbe(int n, float s, int l, int a, int b) : SDPL_SLAM::ORBextractor(n, s, l, a, b) {}
  using SDPL_SLAM::ORBextractor::DistributeOctTree;
  using SDPL_SLAM::ORBextractor::mnFeaturesPerLevel;
  using SDPL_SLAM::ORBextractor::umax;
};
static_assert(sizeof(cv::KeyPoint) == sizeof(orc_keypoint), "KeyPoint layout");
static_assert(sizeof(cv::line_descriptor::KeyLine) == sizeof(orc_keyline), "KeyLine layout");
}  // namespace

extern "C" {
void ref_orb_arena_reset();

void* ref_orb_create(int nfeatures, float scale, int nlevels, int ini_th, int min_th) {
  return new OrbProbe(nfeatures, scale, nlevels, ini_th, min_th);
}
void ref_orb_destroy(void* h) { delete (OrbProbe*)h; }
void ref_orb_tables(void* h, float* sf, float* isf, float* s2, float* is2, int* quota, int* umax) {
  OrbProbe* o = (OrbProbe*)h;
  std::vector<float> a = o->GetScaleFactors(), b = o->GetInverseScaleFactors(), c = o->GetScaleSigmaSquares(),
                     d = o->GetInverseScaleSigmaSquares();
  for (int i = 0; i < o->GetLevels(); i++) { sf[i] = a[i]; isf[i] = b[i]; s2[i] = c[i]; is2[i] = d[i]; quota[i] = o->mnFeaturesPerLevel[i]; }
  for (int i = 0; i < 16; i++) umax[i] = o->umax[i];
}
int ref_orb_extract(void* h, const uint8_t* img, int w, int hh, int stride, orc_keypoint* kps, uint8_t* desc, int cap) {
  OrbProbe* o = (OrbProbe*)h;
  cv::Mat image(hh, w, CV_8UC1, (void*)img, (size_t)stride), mask, descriptors;
  std::vector<cv::KeyPoint> k;
  ref_orb_arena_reset();
  (*o)(image, mask, k, descriptors);
  int n = (int)k.size();
  for (int i = 0; i < n && i < cap; i++) {
    memcpy(&kps[i], &k[i], sizeof(orc_keypoint));
    if (desc) memcpy(desc + 32 * (size_t)i, descriptors.ptr(i), 32);
  }
  return n;
}
void ref_orb_level_size(void* h, int level, int* w, int* hh) {
  OrbProbe* o = (OrbProbe*)h;
  *w = o->mvImagePyramid[level].cols; *hh = o->mvImagePyramid[level].rows;
}
/* copies the padded plane ((w+38) x (h+38), the buffer mvImagePyramid[level] is a view of) into dst */
void ref_orb_level_padded(void* h, int level, uint8_t* dst) {
  OrbProbe* o = (OrbProbe*)h;
  const cv::Mat& m = o->mvImagePyramid[level];
  const size_t st = m.step;
  const uint8_t* base = m.data - 19 * st - 19;
  for (int y = 0; y < m.rows + 38; y++) memcpy(dst + (size_t)y * (m.cols + 38), base + y * st, m.cols + 38);
}
/* DistributeOctTree (src/ORBextractor.cc:528-752) on caller-provided candidates (border-relative coordinates) */
int ref_orb_distribute(void* h, const float* xs, const float* ys, const float* resp, int n, int minX, int maxX, int minY,
                       int maxY, int N, int level, float* ox, float* oy, float* oresp, int cap) {
  OrbProbe* o = (OrbProbe*)h;
  std::vector<cv::KeyPoint> in;
  for (int i = 0; i < n; i++) in.push_back(cv::KeyPoint(xs[i], ys[i], 7.f, -1, resp[i]));
  ref_orb_arena_reset();
  std::vector<cv::KeyPoint> out = o->DistributeOctTree(in, minX, maxX, minY, maxY, N, level);
  for (int i = 0; i < (int)out.size() && i < cap; i++) { ox[i] = out[i].pt.x; oy[i] = out[i].pt.y; oresp[i] = out[i].response; }
  return (int)out.size();
}

void* ref_line_create(int nfeatures, int refine, float lsd_scale, int nlevels, float scale, int extractor) {
  return new SDPL_SLAM::Lineextractor(nfeatures, refine, lsd_scale, nlevels, scale, extractor);
}
void ref_line_destroy(void* h) { delete (SDPL_SLAM::Lineextractor*)h; }
int ref_line_extract(void* h, const uint8_t* img, int w, int hh, int stride, orc_keyline* kls, uint8_t* desc, int cap) {
  SDPL_SLAM::Lineextractor* l = (SDPL_SLAM::Lineextractor*)h;
  cv::Mat image(hh, w, CV_8UC1, (void*)img, (size_t)stride), mask, descriptors;
  std::vector<cv::line_descriptor::KeyLine> k;
  (*l)(image, mask, k, descriptors);
  // the reference appends two pyramid levels per call (Lineextractor.cc:88-96); drop them, tests reuse the handle
  l->mvImagePyramid_l.clear(); l->mvScaleFactor_l.clear(); l->mvInvScaleFactor_l.clear();
  int n = (int)k.size();
  for (int i = 0; i < n && i < cap; i++) {
    memcpy(&kls[i], &k[i], sizeof(orc_keyline));
    if (desc && !descriptors.empty()) memcpy(desc + 32 * (size_t)i, descriptors.ptr(i), 32);
  }
  return n;
}
/* scale tables as Lineextractor::operator() leaves them after a call (Lineextractor.cc:84-96) */
void ref_line_tables(void* h, const uint8_t* img, int w, int hh, int stride, float* sf, float* isf, float* s2, float* is2) {
  SDPL_SLAM::Lineextractor* l = (SDPL_SLAM::Lineextractor*)h;
  cv::Mat image(hh, w, CV_8UC1, (void*)img, (size_t)stride), mask, descriptors;
  std::vector<cv::line_descriptor::KeyLine> k;
  (*l)(image, mask, k, descriptors);
  for (int i = 0; i < l->nlevels_l; i++) {
    sf[i] = l->mvScaleFactor_l[i]; isf[i] = l->mvInvScaleFactor_l[i]; s2[i] = l->mvLevelSigma2_l[i]; is2[i] = l->mvInvLevelSigma2_l[i];
  }
  l->mvImagePyramid_l.clear(); l->mvScaleFactor_l.clear(); l->mvInvScaleFactor_l.clear();
}
/* BinaryDescriptor::compute (binary_descriptor_custom.cpp:524-687) on caller-provided keylines */
int ref_lbd_compute(const uint8_t* img, int w, int hh, int stride, const orc_keyline* kls, int n, uint8_t* desc) {
  cv::Mat image(hh, w, CV_8UC1, (void*)img, (size_t)stride), descriptors;
  std::vector<cv::line_descriptor::KeyLine> k(n);
  for (int i = 0; i < n; i++) memcpy(&k[i], &kls[i], sizeof(orc_keyline));
  cv::Ptr<cv::line_descriptor::BinaryDescriptor> lbd = cv::line_descriptor::BinaryDescriptor::createBinaryDescriptor();
  lbd->compute(image, k, descriptors);
  if (descriptors.empty()) return 0;
  for (int i = 0; i < n; i++) memcpy(desc + 32 * (size_t)i, descriptors.ptr(i), 32);
  return n;
}
/* match() of bitops_custom.hpp:86-99 */
int ref_hamming256(const uint8_t* a, const uint8_t* b) {
  return cv::line_descriptor::match((UINT8*)a, (UINT8*)b, 32);
}
/* BinaryDescriptorMatcher::knnMatch(query, train, matches, k) (binary_descriptor_matcher.cpp:258-335): per query up to k
   (train index, distance) in the order the reference returns them; counts[i] = how many it returned */
void ref_matcher_knn(const uint8_t* q, int nq, const uint8_t* t, int nt, int k, int32_t* train, float* dist, int32_t* counts) {
  cv::Mat Q(nq, 32, CV_8UC1, (void*)q), T(nt, 32, CV_8UC1, (void*)t);
  cv::Ptr<cv::line_descriptor::BinaryDescriptorMatcher> m = cv::line_descriptor::BinaryDescriptorMatcher::createBinaryDescriptorMatcher();
  std::vector<std::vector<cv::DMatch> > out;
  m->knnMatch(Q, T, out, k);
  for (int i = 0; i < nq; i++) {
    int c = i < (int)out.size() ? (int)out[i].size() : 0;
    counts[i] = c;
    for (int j = 0; j < k; j++) {
      train[(size_t)i * k + j] = j < c ? out[i][j].trainIdx : -1;
      dist[(size_t)i * k + j] = j < c ? out[i][j].distance : -1.f;
    }
  }
}
/* BinaryDescriptorMatcher::match(query, train, matches) (binary_descriptor_matcher.cpp:197-254) */
void ref_matcher_match(const uint8_t* q, int nq, const uint8_t* t, int nt, int32_t* train, float* dist) {
  cv::Mat Q(nq, 32, CV_8UC1, (void*)q), T(nt, 32, CV_8UC1, (void*)t);
  cv::Ptr<cv::line_descriptor::BinaryDescriptorMatcher> m = cv::line_descriptor::BinaryDescriptorMatcher::createBinaryDescriptorMatcher();
  std::vector<cv::DMatch> out;
  m->match(Q, T, out);
  for (int i = 0; i < nq; i++) { train[i] = -1; dist[i] = -1.f; }
  for (size_t i = 0; i < out.size(); i++) { train[out[i].queryIdx] = out[i].trainIdx; dist[out[i].queryIdx] = out[i].distance; }
}
/* add({T1, T2}) / train() / knnMatch(query, matches, k) on the stored set (binary_descriptor_matcher.cpp:127-194, 339-425): what
   the reference puts into trainIdx / imgIdx for a data set of two images */
void ref_matcher_stored_knn(const uint8_t* q, int nq, const uint8_t* t1, int n1, const uint8_t* t2, int n2, int k, int32_t* train,
                            int32_t* img, float* dist) {
  cv::Mat Q(nq, 32, CV_8UC1, (void*)q), T1(n1, 32, CV_8UC1, (void*)t1), T2(n2, 32, CV_8UC1, (void*)t2);
  cv::Ptr<cv::line_descriptor::BinaryDescriptorMatcher> m = cv::line_descriptor::BinaryDescriptorMatcher::createBinaryDescriptorMatcher();
  std::vector<cv::Mat> set; set.push_back(T1); set.push_back(T2);
  m->add(set);
  m->train();
  std::vector<std::vector<cv::DMatch> > out;
  m->knnMatch(Q, out, k);
  for (int i = 0; i < nq; i++)
    for (int j = 0; j < k; j++) {
      const bool have = i < (int)out.size() && j < (int)out[i].size();
      train[(size_t)i * k + j] = have ? out[i][j].trainIdx : -1;
      img[(size_t)i * k + j] = have ? out[i][j].imgIdx : -1;
      dist[(size_t)i * k + j] = have ? out[i][j].distance : -1.f;
    }
}
}  // extern "C"
