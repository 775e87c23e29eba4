// oracle/refshim/ref_frame_api.cpp -- ORACLE test infrastructure.  Flat C interface to the reference's own Frame constructor
// (src/Frame.cc:236-907, compiled UNMODIFIED in ref_frame_tu.cpp) so that tests can pin oracle/post_oracle.cpp against it
// (tests/test_oracle_vs_ref.py): the constructor runs the reference's ORBextractor and Lineextractor on the image and then its own
// loops over the mask / depth / flow planes; this file only builds the cv::Mat arguments and copies the public result vectors out.
#include "sdpl_cvshim.hpp"
#include "sdpl_frameshim.hpp"
#include "Frame.h"
#include "../oracle.h"
#include <string.h>

using SDPL_SLAM::Frame;

struct RefFrame {
  SDPL_SLAM::ORBextractor* orb;
  SDPL_SLAM::Lineextractor* line;
  Frame* f;
};

extern "C" {

void* ref_frame_construct(const uint8_t* gray, const float* depth, const float* flow, const int32_t* mask, int w, int h, int nfeatures,
                          float scale, int nlevels, int ini_th, int min_th, int lsd_nfeatures, int lsd_refine, float lsd_scale, int lsd_levels,
                          float lsd_pyr_scale, float th_depth, float th_depth_obj, int use_sample_fea) {
  RefFrame* R = new RefFrame;
  R->orb = new SDPL_SLAM::ORBextractor(nfeatures, scale, nlevels, ini_th, min_th);
  R->line = new SDPL_SLAM::Lineextractor(lsd_nfeatures, lsd_refine, lsd_scale, lsd_levels, lsd_pyr_scale, 0);
  cv::Mat imGray(h, w, CV_8UC1, (void*)gray);
  cv::Mat imDepth(h, w, CV_32F, (void*)depth);
  cv::Mat imFlow(h, w, CV_32F, (void*)flow, (size_t)w * 8);       // CV_32FC2: rows of w (x, y) pairs, read through at<cv::Vec2f>
  cv::Mat maskSEM(h, w, CV_32S, (void*)mask);
  cv::Mat K(3, 3, CV_32F), dist(4, 1, CV_32F);
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) K.at<float>(i, j) = i == j ? 1.f : 0.f;
  K.at<float>(0, 0) = 718.856f; K.at<float>(1, 1) = 718.856f; K.at<float>(0, 2) = 607.1928f; K.at<float>(1, 2) = 185.2157f;   // KITTI 00-02
  for (int i = 0; i < 4; i++) dist.at<float>(i) = 0.f;
  Frame::mbInitialComputations = true;       // image bounds and grid cell size follow THIS image (Frame.cc:886-903)
  const double ts = 0.0; const float bf = 386.1448f;
  R->f = new Frame(imGray, imDepth, imFlow, maskSEM, ts, R->orb, R->line, K, dist, bf, th_depth, th_depth_obj, use_sample_fea);
  return R;
}
void ref_frame_destroy(void* p) {
  RefFrame* R = (RefFrame*)p;
  if (!R) return;
  delete R->f; delete R->orb; delete R->line; delete R;
}
/* counts: {N, N_l (after the two erase loops), mvStatKeysTmp, mvStatKeysLineTmp, mvObjKeys, mvKeysUn} */
void ref_frame_counts(void* p, int32_t* out6) {
  const Frame& F = *((RefFrame*)p)->f;
  out6[0] = (int)F.mvKeys.size(); out6[1] = (int)F.mvKeys_Line.size(); out6[2] = (int)F.mvStatKeysTmp.size();
  out6[3] = (int)F.mvStatKeysLineTmp.size(); out6[4] = (int)F.mvObjKeys.size(); out6[5] = (int)F.mvKeysUn.size();
}
static void put_kps(const std::vector<cv::KeyPoint>& v, orc_keypoint* dst) { if (!v.empty()) memcpy(dst, v.data(), v.size() * sizeof(orc_keypoint)); }
static void put_kls(const std::vector<cv::line_descriptor::KeyLine>& v, orc_keyline* dst) {
  for (size_t i = 0; i < v.size(); i++) memcpy(&dst[i], &v[i], sizeof(orc_keyline));
}
void ref_frame_points(void* p, orc_keypoint* keys, orc_keypoint* stat, orc_keypoint* corres, float* flow_next, float* stat_depth) {
  const Frame& F = *((RefFrame*)p)->f;
  put_kps(F.mvKeys, keys); put_kps(F.mvStatKeysTmp, stat); put_kps(F.mvCorres, corres);
  for (size_t i = 0; i < F.mvFlowNext.size(); i++) { flow_next[2 * i] = F.mvFlowNext[i].x; flow_next[2 * i + 1] = F.mvFlowNext[i].y; }
  for (size_t i = 0; i < F.mvStatDepthTmp.size(); i++) stat_depth[i] = F.mvStatDepthTmp[i];
}
void ref_frame_lines(void* p, orc_keyline* filtered, orc_keyline* stat, orc_keyline* corres, float* flow_next, double* inf_line, float* stat_depth) {
  const Frame& F = *((RefFrame*)p)->f;
  put_kls(F.mvKeys_Line, filtered); put_kls(F.mvStatKeysLineTmp, stat); put_kls(F.mvCorresLine, corres);
  for (size_t i = 0; i < F.mvFlowNext_Line.size(); i++) {
    flow_next[4 * i] = F.mvFlowNext_Line[i].first.x; flow_next[4 * i + 1] = F.mvFlowNext_Line[i].first.y;
    flow_next[4 * i + 2] = F.mvFlowNext_Line[i].second.x; flow_next[4 * i + 3] = F.mvFlowNext_Line[i].second.y;
  }
  for (size_t i = 0; i < F.mvInfiniteLinesCorr.size(); i++) for (int k = 0; k < 3; k++) inf_line[3 * i + k] = F.mvInfiniteLinesCorr[i](k);
  for (size_t i = 0; i < F.mvStatDepthLineTmp.size(); i++) { stat_depth[2 * i] = F.mvStatDepthLineTmp[i].first; stat_depth[2 * i + 1] = F.mvStatDepthLineTmp[i].second; }
}
void ref_frame_objects(void* p, orc_keypoint* keys, orc_keypoint* corres, float* flow_next, float* depth, int32_t* label) {
  const Frame& F = *((RefFrame*)p)->f;
  put_kps(F.mvObjKeys, keys); put_kps(F.mvObjCorres, corres);
  for (size_t i = 0; i < F.mvObjFlowNext.size(); i++) { flow_next[2 * i] = F.mvObjFlowNext[i].x; flow_next[2 * i + 1] = F.mvObjFlowNext[i].y; }
  for (size_t i = 0; i < F.mvObjDepth.size(); i++) depth[i] = F.mvObjDepth[i];
  for (size_t i = 0; i < F.vSemObjLabel.size(); i++) label[i] = F.vSemObjLabel[i];
}
/* mGrid as the CSR of orc_post_grid: cell c = posX * FRAME_GRID_ROWS + posY */
void ref_frame_grid(void* p, int32_t* cell_start, int32_t* items) {
  const Frame& F = *((RefFrame*)p)->f;
  int k = 0;
  for (int x = 0; x < FRAME_GRID_COLS; x++)
    for (int y = 0; y < FRAME_GRID_ROWS; y++) {
      cell_start[x * FRAME_GRID_ROWS + y] = k;
      for (size_t j = 0; j < F.mGrid[x][y].size(); j++) items[k++] = (int32_t)F.mGrid[x][y][j];
    }
  cell_start[FRAME_GRID_COLS * FRAME_GRID_ROWS] = k;
}
int ref_frame_features_in_area(void* p, float x, float y, float r, int min_level, int max_level, int32_t* out, int cap) {
  const Frame& F = *((RefFrame*)p)->f;
  const std::vector<size_t> v = F.GetFeaturesInArea(x, y, r, min_level, max_level);
  for (size_t i = 0; i < v.size() && (int)i < cap; i++) out[i] = (int32_t)v[i];
  return (int)v.size();
}

}  // extern "C"
