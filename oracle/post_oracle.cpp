// oracle/post_oracle.cpp -- CPU ORACLE (test infrastructure, NOT the product) of the per-frame post-processing that
// Frame::Frame runs on the extractor outputs (reference: src/Frame.cc).  Each function restates
// one loop of the constructor, statement by statement, with std::vector push_back semantics replaced by output arrays in the
// same order.  Inputs are the reference's own per-frame planes: maskSEM (CV_32S), imDepth (CV_32F), imFlow (CV_32FC2).
//
// Parity status: PINNED for every function that restates Frame.cc -- oracle/refshim compiles src/Frame.cc unmodified (Frame.h's
// Eigen / xfeatures2d / flann / Converter includes are satisfied by minimal stand-ins, nothing of them is used on this path) and
// tests/test_oracle_vs_ref.py::test_frame_post_processing_equals_reference_frame_constructor runs the reference's own constructor on
// synthetic image / mask / depth / flow planes and compares all of its output vectors with the functions below, byte for byte.
// The functions restated from src/MapPoint.cc (orc_post_distinctive_descriptors, orc_post_predict_scale) and the window descriptor
// search (orc_post_search_area) stay "parity unpinned": MapPoint.cc is dead code in the reference (not in CMakeLists.txt, includes
// a header that does not exist), so they are checked against an independent numpy restatement (tests/test_oracle_post.py) only.
//
// C++ conversions the code below relies on (all as written in Frame.cc, which has `using namespace std`, :24):
//   int x = kp.pt.x            float -> int truncates toward zero                                   (:353-356, :488-489, :519-522)
//   Mat::at<T>(float, float)   the float indices convert to int the same way                        (:737, :756-757)
//   abs(float)                 std::abs(float) (float result)                                         (:376)
//   pow(int, 2), sqrt(double)  computed in double, then narrowed to float on assignment              (:371)
//   j + flow_x                 int + float -> float                                                   (:792)
#include "oracle.h"
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace {
struct Planes {
  const int32_t* mask; const float* depth; const float* flow; int w, h;
  int m(int y, int x) const { return mask[(size_t)y * w + x]; }
  float d(int y, int x) const { return depth[(size_t)y * w + x]; }
  float fx(int y, int x) const { return flow[((size_t)y * w + x) * 2]; }
  float fy(int y, int x) const { return flow[((size_t)y * w + x) * 2 + 1]; }
};
orc_keypoint make_kp(float x, float y, float size, float angle, float response, int octave, int class_id) {
  orc_keypoint k; k.x = x; k.y = y; k.size = size; k.angle = angle; k.response = response; k.octave = octave; k.class_id = class_id;
  return k;
}
// KeyLine corr_line of Frame.cc:566-582 (and :672-687, :842-856): the fields the reference assigns; the octave-relative end points,
// which it leaves uninitialised, are zero here
orc_keyline make_corr_line(float sx, float sy, float ex, float ey, int octave) {
  orc_keyline c; memset(&c, 0, sizeof(c));
  c.sx = sx; c.sy = sy; c.ex = ex; c.ey = ey; c.octave = octave;
  c.angle = std::atan2(c.ey - c.sy, c.ex - c.sx);                     // std::atan2(float, float)
  c.pt_x = (c.sx + c.ex) / 2; c.pt_y = (c.sy + c.ey) / 2;
  c.size = (c.ex - c.sx) * (c.ey - c.sy);
  c.length = (float)std::sqrt(std::pow(c.ex - c.sx, 2) + std::pow(c.ey - c.sy, 2));
  c.response = 0; c.class_id = -1; c.num_pixels = 0;
  return c;
}
}  // namespace

extern "C" {

/* Semi-dense features on objects, Frame.cc:769-809.  Outputs in scan order; returns the count (may exceed cap: only cap are
   written). */
int orc_post_sample_objects(const int32_t* mask, const float* depth, const float* flow, int w, int h, int step, float th_depth_obj,
                            orc_keypoint* keys, orc_keypoint* corres, float* flow_next, float* depth_out, int32_t* label, int cap) {
  const Planes P{mask, depth, flow, w, h};
  int n = 0;
  for (int i = 0; i < h; i = i + step)
    for (int j = 0; j < w; j = j + step) {
      if (P.m(i, j) != 0 && P.d(i, j) < th_depth_obj && P.d(i, j) > 0) {
        const float flow_x = P.fx(i, j), flow_y = P.fy(i, j);
        if (j + flow_x < w && j + flow_x > 0 && i + flow_y < h && i + flow_y > 0) {
          if (n < cap) {
            flow_next[2 * n] = flow_x; flow_next[2 * n + 1] = flow_y;
            corres[n] = make_kp(j + flow_x, i + flow_y, 0, 0, 0, -1, -1);     // cv::KeyPoint(x, y, size, angle, response, octave)
            keys[n] = make_kp((float)j, (float)i, 0, 0, 0, -1, -1);
            depth_out[n] = P.d(i, j);
            label[n] = P.m(i, j);
          }
          n++;
        }
      }
    }
  return n;
}

/* The two erase loops on mvKeys_Line, Frame.cc:349-389: a line goes when the depth at its mid point departs from the mean of
   the end-point depths by more than 10 * length / 1000, or when its end points carry different mask labels.  keep_idx[k] = index
   (in the input) of the k-th surviving line -- the row of its LBD descriptor (the reference forgets to erase those). */
int orc_post_filter_lines(const orc_keyline* kls, int n, const int32_t* mask, const float* depth, int w, int h, orc_keyline* out,
                          int32_t* keep_idx) {
  const Planes P{mask, depth, nullptr, w, h};
  int k = 0;
  for (int i = 0; i < n; i++) {
    const orc_keyline& l = kls[i];
    const int x1 = (int)l.sx, y1 = (int)l.sy, x2 = (int)l.ex, y2 = (int)l.ey;
    const int xm = (x1 + x2) / 2, ym = (y1 + y2) / 2;
    const float depthStart = P.d(y1, x1), depthMid = P.d(ym, xm), depthEnd = P.d(y2, x2);
    const float depthMidExpected = (depthStart + depthEnd) / 2;
    const float baseThreshold = 10.0;
    const float lineLength = std::sqrt(std::pow(x2 - x1, 2) + std::pow(y2 - y1, 2));
    const float threshold = baseThreshold * (lineLength / 1000);
    if (std::abs(depthMid - depthMidExpected) > threshold) continue;
    if (P.m(y1, x1) != P.m(y2, x2)) continue;
    out[k] = l; keep_idx[k] = i; k++;
  }
  return k;
}

/* Static point correspondences from the detected features (UseSampleFea == 0), Frame.cc:482-512, and their depths, :728-745.
   Returns the count; src_idx[k] = index of the key point in the input. */
int orc_post_point_corres(const orc_keypoint* kps, int n, const int32_t* mask, const float* depth, const float* flow, int w, int h,
                          float th_depth, orc_keypoint* stat, orc_keypoint* corres, float* flow_next, float* stat_depth, int32_t* src_idx) {
  const Planes P{mask, depth, flow, w, h};
  int k = 0;
  for (int i = 0; i < n; ++i) {
    const int x = (int)kps[i].x, y = (int)kps[i].y;
    if (P.m(y, x) != 0) continue;
    if (P.d(y, x) > th_depth || P.d(y, x) <= 0) continue;
    const float flow_xe = P.fx(y, x), flow_ye = P.fy(y, x);
    if (flow_xe != 0 && flow_ye != 0) {
      if (kps[i].x + flow_xe < w && kps[i].y + flow_ye < h && kps[i].x < w && kps[i].y < h) {
        stat[k] = kps[i];
        corres[k] = make_kp(kps[i].x + flow_xe, kps[i].y + flow_ye, 0, 0, 0, kps[i].octave, -1);
        flow_next[2 * k] = flow_xe; flow_next[2 * k + 1] = flow_ye;
        // :728-745: mvStatDepthTmp[i] = d if d > 0 else -1, with d = imDepth.at<float>(kp.pt.y, kp.pt.x)
        const float d = P.d((int)kps[i].y, (int)kps[i].x);
        stat_depth[k] = d > 0 ? d : -1.f;
        src_idx[k] = i;
        k++;
      }
    }
  }
  return k;
}

/* Line correspondences, Frame.cc:513-604 and :746-763.  Lines with both end points on the same object go to obj (n_obj), lines on
   the static background with valid depth and flow to stat / corres / flow_next (4 floats: start x,y, end x,y) / inf_line (3
   doubles: normalised cross product of the homogeneous end points) / stat_depth (2 floats).  Returns the static count. */
int orc_post_line_corres(const orc_keyline* kls, int n, const int32_t* mask, const float* depth, const float* flow, int w, int h,
                         float th_depth, orc_keyline* obj, int32_t* n_obj, orc_keyline* stat, orc_keyline* corres, float* flow_next,
                         double* inf_line, float* stat_depth, int32_t* src_idx) {
  const Planes P{mask, depth, flow, w, h};
  int k = 0, no = 0;
  for (int i = 0; i < n; ++i) {
    const int start_x = (int)kls[i].sx, start_y = (int)kls[i].sy, end_x = (int)kls[i].ex, end_y = (int)kls[i].ey;
    if (P.m(start_y, start_x) != 0 && P.m(end_y, end_x) != 0) {
      if (P.m(start_y, start_x) == P.m(end_y, end_x)) obj[no++] = kls[i];
      continue;
    }
    if (P.m(start_y, start_x) != 0 || P.m(end_y, end_x) != 0) continue;
    if (std::fabs(start_x - end_x) < 1e-6 && std::fabs(start_y - end_y) < 1e-6) continue;
    if (P.d(start_y, start_x) > th_depth || P.d(start_y, start_x) <= 0 || P.d(end_y, end_x) > th_depth || P.d(end_y, end_x) <= 0) continue;
    const float fsx = P.fx(start_y, start_x), fsy = P.fy(start_y, start_x), fex = P.fx(end_y, end_x), fey = P.fy(end_y, end_x);
    if (fsx != 0 && fsy != 0 && fex != 0 && fey != 0) {
      if (start_x + fsx < w && start_y + fsy < h && end_x + fex < w && end_y + fey < h && start_x + fsx > 0 && start_y + fsy > 0 &&
          end_x + fex > 0 && end_y + fey > 0) {
        stat[k] = kls[i];
        const orc_keyline c = make_corr_line(start_x + fsx, start_y + fsy, end_x + fex, end_y + fey, kls[i].octave);
        corres[k] = c;
        flow_next[4 * k] = fsx; flow_next[4 * k + 1] = fsy; flow_next[4 * k + 2] = fex; flow_next[4 * k + 3] = fey;
        // Eigen::Vector3d start(sx, sy, 1), end(ex, ey, 1); inf_line = start.cross(end).normalized()   (:590-594)
        const double a0 = c.sx, a1 = c.sy, a2 = 1, b0 = c.ex, b1 = c.ey, b2 = 1;
        const double c0 = a1 * b2 - a2 * b1, c1 = a2 * b0 - a0 * b2, c2 = a0 * b1 - a1 * b0;
        const double nn = std::sqrt(c0 * c0 + c1 * c1 + c2 * c2);
        inf_line[3 * k] = nn > 0 ? c0 / nn : c0; inf_line[3 * k + 1] = nn > 0 ? c1 / nn : c1; inf_line[3 * k + 2] = nn > 0 ? c2 / nn : c2;
        // :746-763 -- mind the missing braces of the reference: `second` is assigned unconditionally
        const float d_start = P.d((int)kls[i].sy, (int)kls[i].sx), d_end = P.d((int)kls[i].ey, (int)kls[i].ex);
        stat_depth[2 * k] = -1.f; stat_depth[2 * k + 1] = -1.f;
        if (d_start > 0 && d_end > 0) stat_depth[2 * k] = d_start;
        stat_depth[2 * k + 1] = d_end;
        src_idx[k] = i;
        k++;
      }
    }
  }
  *n_obj = no;
  return k;
}

/* AssignFeaturesToGrid / PosInGrid, Frame.cc:910-925, 1023-1035, for undistorted key points (mDistCoef[0] == 0: mvKeysUn = mvKeys,
   :1039-1043; image bounds 0 .. cols / rows, ComputeImageBounds).  cell c = posX * rows + posY; cell_start[c] .. cell_start[c+1]
   index `items`, which lists the key points of a cell in ascending index order (push_back order). */
void orc_post_grid(const orc_keypoint* kps, int n, int w, int h, int grid_cols, int grid_rows, int32_t* cell_start, int32_t* items) {
  const float mnMinX = 0.f, mnMaxX = (float)w, mnMinY = 0.f, mnMaxY = (float)h;
  const float wInv = static_cast<float>(grid_cols) / static_cast<float>(mnMaxX - mnMinX);
  const float hInv = static_cast<float>(grid_rows) / static_cast<float>(mnMaxY - mnMinY);
  std::vector<std::vector<int>> grid((size_t)grid_cols * grid_rows);
  for (int i = 0; i < n; i++) {
    const int posX = (int)std::round((kps[i].x - mnMinX) * wInv), posY = (int)std::round((kps[i].y - mnMinY) * hInv);
    if (posX < 0 || posX >= grid_cols || posY < 0 || posY >= grid_rows) continue;
    grid[(size_t)posX * grid_rows + posY].push_back(i);
  }
  int k = 0;
  for (size_t c = 0; c < grid.size(); c++) {
    cell_start[c] = k;
    for (int v : grid[c]) items[k++] = v;
  }
  cell_start[grid.size()] = k;
}

/* Frame::GetFeaturesInArea, src/Frame.cc:970-1023, on the grid of orc_post_grid: indices of the key points inside the square window
   |dx| < r, |dy| < r around (x, y) with octave in [minLevel, maxLevel] (maxLevel < 0: no upper bound), in the order the reference
   visits them: cells ix-major, then iy, then push_back order inside a cell.  Returns the count (only cap are written). */
int orc_post_features_in_area(const orc_keypoint* kps, int w, int h, int grid_cols, int grid_rows, const int32_t* cell_start,
                              const int32_t* items, float x, float y, float r, int minLevel, int maxLevel, int32_t* out, int cap) {
  const float mnMinX = 0.f, mnMaxX = (float)w, mnMinY = 0.f, mnMaxY = (float)h;
  const float wInv = static_cast<float>(grid_cols) / static_cast<float>(mnMaxX - mnMinX);
  const float hInv = static_cast<float>(grid_rows) / static_cast<float>(mnMaxY - mnMinY);
  int n = 0;
  const int nMinCellX = std::max(0, (int)std::floor((x - mnMinX - r) * wInv));
  if (nMinCellX >= grid_cols) return 0;
  const int nMaxCellX = std::min(grid_cols - 1, (int)std::ceil((x - mnMinX + r) * wInv));
  if (nMaxCellX < 0) return 0;
  const int nMinCellY = std::max(0, (int)std::floor((y - mnMinY - r) * hInv));
  if (nMinCellY >= grid_rows) return 0;
  const int nMaxCellY = std::min(grid_rows - 1, (int)std::ceil((y - mnMinY + r) * hInv));
  if (nMaxCellY < 0) return 0;
  const bool bCheckLevels = (minLevel > 0) || (maxLevel >= 0);
  for (int ix = nMinCellX; ix <= nMaxCellX; ix++)
    for (int iy = nMinCellY; iy <= nMaxCellY; iy++) {
      const int c = ix * grid_rows + iy;
      for (int j = cell_start[c]; j < cell_start[c + 1]; j++) {
        const orc_keypoint& kpUn = kps[items[j]];
        if (bCheckLevels) {
          if (kpUn.octave < minLevel) continue;
          if (maxLevel >= 0)
            if (kpUn.octave > maxLevel) continue;
        }
        const float distx = kpUn.x - x, disty = kpUn.y - y;
        if (std::fabs(distx) < r && std::fabs(disty) < r) { if (n < cap) out[n] = items[j]; n++; }
      }
    }
  return n;
}

/* MapPoint::ComputeDistinctiveDescriptors, src/MapPoint.cc:242-307, for n_points map points whose observed descriptors are rows
   start[p] .. start[p+1] of desc (32 bytes each): all pair distances, per row the median = sorted[(int)(0.5 * (N - 1))] (the row
   includes the zero self distance), the first row with the least median wins (strict <).  best_idx[p] = its index among the point's
   observations (-1 for a point without observations), out_desc = its 32 bytes. */
void orc_post_distinctive_descriptors(const uint8_t* desc, const int32_t* start, int n_points, int32_t* best_idx, uint8_t* out_desc) {
  for (int p = 0; p < n_points; p++) {
    const int N = start[p + 1] - start[p];
    best_idx[p] = -1;
    if (N <= 0) continue;
    const uint8_t* d = desc + (size_t)start[p] * 32;
    std::vector<std::vector<float>> Distances(N, std::vector<float>(N));
    for (int i = 0; i < N; i++) {
      Distances[i][i] = 0;
      for (int j = i + 1; j < N; j++) {
        const int distij = orc_hamming256(d + 32 * (size_t)i, d + 32 * (size_t)j);
        Distances[i][j] = (float)distij; Distances[j][i] = (float)distij;
      }
    }
    int BestMedian = 2147483647, BestIdx = 0;
    for (int i = 0; i < N; i++) {
      std::vector<int> vDists(Distances[i].begin(), Distances[i].end());
      std::sort(vDists.begin(), vDists.end());
      const int median = vDists[(size_t)(0.5 * (N - 1))];
      if (median < BestMedian) { BestMedian = median; BestIdx = i; }
    }
    best_idx[p] = BestIdx;
    memcpy(out_desc + 32 * (size_t)p, d + 32 * (size_t)BestIdx, 32);
  }
}

/* MapPoint::PredictScale, src/MapPoint.cc:385-417: nScale = ceil(log(maxDistance / currentDist) / logScaleFactor) clamped to
   [0, nLevels - 1]; float arithmetic with std::log(float) as written (Frame.cc / MapPoint.cc have `using namespace std`). */
void orc_post_predict_scale(const float* max_distance, const float* current_dist, int n, float log_scale_factor, int n_levels, int32_t* out) {
  for (int i = 0; i < n; i++) {
    const float ratio = max_distance[i] / current_dist[i];
    int nScale = (int)std::ceil(std::log(ratio) / log_scale_factor);
    if (nScale < 0) nScale = 0;
    else if (nScale >= n_levels) nScale = n_levels - 1;
    out[i] = nScale;
  }
}

/* The descriptor search a projection match runs on Frame::GetFeaturesInArea (the reference keeps the window search, src/Frame.cc:
   970-1023, and the descriptor metric; the ORB-SLAM2 loop around them is not in its tree -- SURVEY.md 8f row 4): among the key points
   of the window around (x, y), the best and second-best Hamming distance to `qdesc`, ties to the first visited.  out5 = {best index,
   best distance, best octave, second distance, second octave} (index -1 / distance 256 when missing). */
void orc_post_search_area(const orc_keypoint* kps, const uint8_t* desc, int w, int h, int grid_cols, int grid_rows, const int32_t* cell_start,
                          const int32_t* items, float x, float y, float r, int minLevel, int maxLevel, const uint8_t* qdesc, int32_t* out5) {
  std::vector<int32_t> cand(cell_start[grid_cols * grid_rows] + 1);
  const int n = orc_post_features_in_area(kps, w, h, grid_cols, grid_rows, cell_start, items, x, y, r, minLevel, maxLevel, cand.data(), (int)cand.size());
  int bestDist = 256, bestIdx = -1, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1;
  for (int c = 0; c < n; c++) {
    const int idx = cand[c];
    const int dist = orc_hamming256(qdesc, desc + 32 * (size_t)idx);
    if (dist < bestDist) { bestDist2 = bestDist; bestLevel2 = bestLevel; bestDist = dist; bestLevel = kps[idx].octave; bestIdx = idx; }
    else if (dist < bestDist2) { bestDist2 = dist; bestLevel2 = kps[idx].octave; }
  }
  out5[0] = bestIdx; out5[1] = bestDist; out5[2] = bestLevel; out5[3] = bestDist2; out5[4] = bestLevel2;
}

}  // extern "C"
