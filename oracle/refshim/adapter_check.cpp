// oracle/refshim/adapter_check.cpp -- test infrastructure.  Compiles include/sdpl_adapters.hpp -- the C++ classes a
// maintainer drops into the reference tree in place of ORBextractor.{h,cc} / Lineextractor.{h,cc} -- against
// the reference's own line_descriptor headers (KeyLine) and the OpenCV stand-in, links it with the PRODUCT library
// libsdpl_frontend.so and calls it the way Frame::ExtractORB / Frame::ExtractLines do (src/Frame.cc:927-949):
//     (*mpORBextractorLeft)(im, cv::Mat(), mvKeys, mDescriptors);
//     (*mpLineextractorLeft)(im, mask, mvKeys_Line, mDescriptors_Line);
// Usage: adapter_check <w> <h> <frame.u8> <partner.u8> <out.bin>.  No oracle code is linked (shim_core.cpp only).
// out.bin: int32 nkp, nkl, nmatch; cv::KeyPoint[nkp]; u8[nkp][32]; KeyLine[nkl]; u8[nkl][32]; cv::DMatch best[nkp], second[nkp]
#include "sdpl_adapters.hpp"
#include <stdio.h>

static std::vector<unsigned char> slurp(const char* path, size_t n) {
  std::vector<unsigned char> b(n);
  FILE* f = fopen(path, "rb");
  if (!f || fread(b.data(), 1, n, f) != n) { fprintf(stderr, "cannot read %s\n", path); exit(2); }
  fclose(f);
  return b;
}

int main(int argc, char** argv) {
  if (argc != 6) { fprintf(stderr, "usage: adapter_check w h frame.u8 partner.u8 out.bin\n"); return 2; }
  const int w = atoi(argv[1]), h = atoi(argv[2]);
  std::vector<unsigned char> a = slurp(argv[3], (size_t)w * h), b = slurp(argv[4], (size_t)w * h);
  cv::Mat im(h, w, CV_8UC1, a.data()), im2(h, w, CV_8UC1, b.data());
  try {
    // constructor arguments of Tracking::Tracking (src/Tracking.cc:114-126) for KITTI
    SDPL_SLAM::ORBextractor* mpORBextractorLeft = new SDPL_SLAM::ORBextractor(2000, 1.2f, 8, 20, 7);
    SDPL_SLAM::Lineextractor* mpLineextractorLeft = new SDPL_SLAM::Lineextractor(0, 2, 0.8f, 2, 2.0f, 0);
    std::vector<cv::KeyPoint> mvKeys, mvKeys2;
    cv::Mat mDescriptors, mDescriptors2, mask, mDescriptors_Line;
    std::vector<cv::line_descriptor::KeyLine> mvKeys_Line;
    (*mpORBextractorLeft)(im, cv::Mat(), mvKeys, mDescriptors);
    (*mpORBextractorLeft)(im2, cv::Mat(), mvKeys2, mDescriptors2);
    (*mpLineextractorLeft)(im, mask, mvKeys_Line, mDescriptors_Line);
    if (mpORBextractorLeft->GetLevels() != 8 || mpORBextractorLeft->GetScaleFactors().size() != 8 ||
        mpLineextractorLeft->mvScaleFactor_l.size() != 2 || mpLineextractorLeft->nlevels_l != 2) return 3;
    sdpl::HammingMatcher matcher;
    std::vector<std::vector<cv::DMatch> > knn;
    matcher.knnMatch(mDescriptors, mDescriptors2, knn, 2);
    {
      // the stored (device-resident) train set and a general k must agree with the explicit-train call
      std::vector<cv::Mat> set; set.push_back(mDescriptors2);
      matcher.add(set); matcher.train();
      std::vector<std::vector<cv::DMatch> > knn3;
      matcher.knnMatch(mDescriptors, knn3, 3);
      if (knn3.size() != knn.size()) return 4;
      for (size_t i = 0; i < knn.size(); i++)
        for (size_t j = 0; j < knn[i].size(); j++)
          if (knn3[i].size() < knn[i].size() || knn3[i][j].trainIdx != knn[i][j].trainIdx || knn3[i][j].distance != knn[i][j].distance ||
              knn3[i][j].imgIdx != 0) return 4;
      matcher.clear();
    }
    int32_t head[3] = {(int32_t)mvKeys.size(), (int32_t)mvKeys_Line.size(), (int32_t)knn.size()};
    FILE* f = fopen(argv[5], "wb");
    fwrite(head, 4, 3, f);
    fwrite(mvKeys.data(), sizeof(cv::KeyPoint), mvKeys.size(), f);
    for (int i = 0; i < mDescriptors.rows; i++) fwrite(mDescriptors.ptr(i), 1, 32, f);
    fwrite(mvKeys_Line.data(), sizeof(cv::line_descriptor::KeyLine), mvKeys_Line.size(), f);
    for (int i = 0; i < mDescriptors_Line.rows; i++) fwrite(mDescriptors_Line.ptr(i), 1, 32, f);
    for (int j = 0; j < 2; j++)
      for (size_t i = 0; i < knn.size(); i++) {
        cv::DMatch d = j < (int)knn[i].size() ? knn[i][j] : cv::DMatch(-1, -1, -1, 257.f);
        fwrite(&d, sizeof(d), 1, f);
      }
    fclose(f);
    delete mpORBextractorLeft;
    delete mpLineextractorLeft;
  } catch (const std::exception& e) {
    fprintf(stderr, "adapter_check: %s\n", e.what());
    return 1;
  }
  return 0;
}
