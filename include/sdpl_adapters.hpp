// sdpl_adapters.hpp -- header-only C++ adapters with the reference's own class signatures, forwarding to the C ABI of
// libsdpl_frontend.so (include/sdpl_frontend.h).  Needs OpenCV headers (cv::Mat, cv::KeyPoint,
// cv::line_descriptor::KeyLine), which this build image does not have: it is compiled in the reference's tree, not here.
//
//   SDPL_SLAM::ORBextractor   replaces include/ORBextractor.h:33-99  + src/ORBextractor.cc
//   SDPL_SLAM::Lineextractor  replaces include/Lineextractor.h:51-87 + src/Lineextractor.cc
//   sdpl::HammingMatcher      surface of cv::line_descriptor::BinaryDescriptorMatcher::match/knnMatch
//                             (3rdparty/line_descriptor/include/line_descriptor/descriptor_custom.hpp:1015-1126)
//
// Frame.cc / Tracking.cc compile unchanged against these classes (same constructor arguments, operator() signatures,
// getters and public members).  sdpl_keypoint == cv::KeyPoint (28 B POD) and sdpl_keyline == KeyLine (68 B POD) are
// layout-identical, checked by the static_asserts below, so vectors are filled in place without conversion.
#pragma once
#include <stdexcept>
#include <string>
#include <vector>
#include <opencv2/core/core.hpp>
#include <opencv2/features2d/features2d.hpp>
#include <line_descriptor_custom.hpp>
#include "sdpl_frontend.h"

namespace sdpl {
inline void check(int rc) {
  if (rc != SDPL_OK) throw std::runtime_error(std::string("sdpl_frontend: ") + sdpl_strerror(rc) + ": " + sdpl_last_error());
}
}  // namespace sdpl

namespace SDPL_SLAM {

static_assert(sizeof(cv::KeyPoint) == sizeof(sdpl_keypoint), "cv::KeyPoint layout");
static_assert(sizeof(cv::line_descriptor::KeyLine) == sizeof(sdpl_keyline), "KeyLine layout");
static_assert(sizeof(cv::DMatch) == sizeof(sdpl_dmatch), "cv::DMatch layout");

class ORBextractor {
 public:
  enum { HARRIS_SCORE = 0, FAST_SCORE = 1 };
  ORBextractor(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST, int device = 0) {
    sdpl::check(sdpl_orb_create(&h_, nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, device));
    const int n = sdpl_orb_levels(h_);
    mvScaleFactor.resize(n); mvInvScaleFactor.resize(n); mvLevelSigma2.resize(n); mvInvLevelSigma2.resize(n);
    sdpl::check(sdpl_orb_tables(h_, mvScaleFactor.data(), mvInvScaleFactor.data(), mvLevelSigma2.data(), mvInvLevelSigma2.data()));
    scaleFactor_ = scaleFactor;
  }
  ~ORBextractor() { sdpl_orb_destroy(h_); }
  ORBextractor(const ORBextractor&) = delete;
  ORBextractor& operator=(const ORBextractor&) = delete;

  // mask is ignored, as in the reference (src/ORBextractor.cc:1035)
  void operator()(cv::InputArray image, cv::InputArray /*mask*/, std::vector<cv::KeyPoint>& keypoints, cv::OutputArray descriptors) {
    if (image.empty()) return;                                   // :1038
    cv::Mat im = image.getMat();
    CV_Assert(im.type() == CV_8UC1);                             // :1042
    const int cap = sdpl_orb_max_keypoints(h_);
    keypoints.resize(cap);
    cv::Mat desc(cap, 32, CV_8U);
    int n = 0;
    sdpl::check(sdpl_orb_extract(h_, im.data, im.cols, im.rows, (int)im.step, reinterpret_cast<sdpl_keypoint*>(keypoints.data()),
                                 desc.data, cap, &n));
    keypoints.resize(n);
    if (n == 0) descriptors.release(); else desc.rowRange(0, n).copyTo(descriptors);
  }
  int GetLevels() { return sdpl_orb_levels(h_); }
  float GetScaleFactor() { return scaleFactor_; }
  std::vector<float> GetScaleFactors() { return mvScaleFactor; }
  std::vector<float> GetInverseScaleFactors() { return mvInvScaleFactor; }
  std::vector<float> GetScaleSigmaSquares() { return mvLevelSigma2; }
  std::vector<float> GetInverseScaleSigmaSquares() { return mvInvLevelSigma2; }
  // mvImagePyramid (include/ORBextractor.h:71) is read by nobody outside the class; fetch a level on demand instead
  cv::Mat ImagePyramidLevel(int level) {
    int w = 0, h = 0;
    sdpl::check(sdpl_orb_pyramid_level(h_, 0, level, nullptr, 0, &w, &h));
    cv::Mat padded(h + 38, w + 38, CV_8U);
    sdpl::check(sdpl_orb_pyramid_level(h_, 0, level, padded.data, (int)padded.step, &w, &h));
    return padded(cv::Rect(19, 19, w, h));
  }
  sdpl_orb* handle() { return h_; }

 protected:
  sdpl_orb* h_ = nullptr;
  float scaleFactor_ = 1.2f;
  std::vector<float> mvScaleFactor, mvInvScaleFactor, mvLevelSigma2, mvInvLevelSigma2;
};

class Lineextractor {
 public:
  Lineextractor(int lsd_nfeatures, int lsd_refine, float lsd_scale, int nlevels, float scale, int extractor, int device = 0) {
    sdpl::check(sdpl_line_create(&h_, lsd_nfeatures, lsd_refine, lsd_scale, nlevels, scale, extractor, device));
    nlevels_l = nlevels;
    // the reference fills these inside operator() and appends on every call (unbounded growth, Lineextractor.cc:84-96),
    // and Frame reads them BEFORE the first ExtractLines call (Frame.cc:298-302); here they are fixed tables, valid from
    // construction
    mvScaleFactor_l.resize(nlevels); mvInvScaleFactor_l.resize(nlevels); mvLevelSigma2_l.resize(nlevels); mvInvLevelSigma2_l.resize(nlevels);
    sdpl::check(sdpl_line_tables(h_, mvScaleFactor_l.data(), mvInvScaleFactor_l.data(), mvLevelSigma2_l.data(), mvInvLevelSigma2_l.data()));
  }
  ~Lineextractor() { sdpl_line_destroy(h_); }
  Lineextractor(const Lineextractor&) = delete;
  Lineextractor& operator=(const Lineextractor&) = delete;

  void operator()(const cv::Mat& image, const cv::Mat& /*mask*/, std::vector<cv::line_descriptor::KeyLine>& keylines,
                  cv::Mat& descriptors_line) {
    CV_Assert(image.type() == CV_8UC1);
    const int cap = 8192;
    keylines.resize(cap);
    cv::Mat desc(cap, 32, CV_8U);
    int n = 0;
    sdpl::check(sdpl_line_extract(h_, image.data, image.cols, image.rows, (int)image.step,
                                  reinterpret_cast<sdpl_keyline*>(keylines.data()), desc.data, cap, &n));
    keylines.resize(n);
    // BinaryDescriptor::compute leaves `descriptors` untouched when there are no lines (binary_descriptor_custom.cpp:556-560)
    if (n > 0) desc.rowRange(0, n).copyTo(descriptors_line);
  }
  std::vector<cv::Mat> mvImagePyramid_l;      // kept for source compatibility; not filled (read by nobody)
  std::vector<float> mvScaleFactor_l, mvInvScaleFactor_l, mvLevelSigma2_l, mvInvLevelSigma2_l;
  int nlevels_l;
  sdpl_line* handle() { return h_; }

 protected:
  sdpl_line* h_ = nullptr;
};

}  // namespace SDPL_SLAM

namespace sdpl {
// Brute-force 256-bit Hamming search with the surface of cv::line_descriptor::BinaryDescriptorMatcher
// (descriptor_custom.hpp:1015-1126): match / knnMatch against an explicit train matrix or against the set kept by add() / train()
// (resident on the device; trainIdx = row of the concatenated set, imgIdx = image, binary_descriptor_matcher.cpp:381-401).
class HammingMatcher {
 public:
  explicit HammingMatcher(int device = 0) { check(sdpl_matcher_create(&h_, device)); }
  ~HammingMatcher() { sdpl_matcher_destroy(h_); }
  HammingMatcher(const HammingMatcher&) = delete;
  HammingMatcher& operator=(const HammingMatcher&) = delete;
  void add(const std::vector<cv::Mat>& descriptors) {
    for (size_t i = 0; i < descriptors.size(); i++) { check_desc(descriptors[i]); check(sdpl_matcher_add(h_, descriptors[i].data, descriptors[i].rows)); }
  }
  void train() { check(sdpl_matcher_train(h_)); }
  void clear() { check(sdpl_matcher_clear(h_)); }
  void match(const cv::Mat& query, const cv::Mat& train, std::vector<cv::DMatch>& matches) {
    std::vector<std::vector<cv::DMatch> > knn;
    knnMatch(query, train, knn, 1);
    matches.clear();
    for (size_t i = 0; i < knn.size(); i++) if (!knn[i].empty()) matches.push_back(knn[i][0]);
  }
  void match(const cv::Mat& query, std::vector<cv::DMatch>& matches) {
    std::vector<std::vector<cv::DMatch> > knn;
    knnMatch(query, knn, 1);
    matches.clear();
    for (size_t i = 0; i < knn.size(); i++) if (!knn[i].empty()) matches.push_back(knn[i][0]);
  }
  void knnMatch(const cv::Mat& query, const cv::Mat& train, std::vector<std::vector<cv::DMatch> >& matches, int k = 2) {
    check_desc(query);
    if (!train.empty()) check_desc(train);
    CV_Assert(k >= 1);
    std::vector<cv::DMatch> flat((size_t)query.rows * k);
    check(sdpl_match_knn(h_, query.data, query.rows, train.data, train.rows, k, reinterpret_cast<sdpl_dmatch*>(flat.data())));
    unflatten(flat, query.rows, k, matches);
  }
  void knnMatch(const cv::Mat& query, std::vector<std::vector<cv::DMatch> >& matches, int k = 2) {
    check_desc(query);
    CV_Assert(k >= 1);
    std::vector<cv::DMatch> flat((size_t)query.rows * k);
    check(sdpl_matcher_knn(h_, query.data, query.rows, k, reinterpret_cast<sdpl_dmatch*>(flat.data())));
    unflatten(flat, query.rows, k, matches);
  }
 private:
  static void check_desc(const cv::Mat& d) { CV_Assert(d.type() == CV_8U && d.cols == 32 && d.isContinuous()); }
  static void unflatten(const std::vector<cv::DMatch>& flat, int nq, int k, std::vector<std::vector<cv::DMatch> >& matches) {
    matches.resize(nq);
    for (int i = 0; i < nq; i++) {
      matches[i].clear();
      for (int j = 0; j < k; j++) if (flat[(size_t)i * k + j].trainIdx >= 0) matches[i].push_back(flat[(size_t)i * k + j]);
    }
  }
  sdpl_matcher* h_ = nullptr;
};
}  // namespace sdpl
