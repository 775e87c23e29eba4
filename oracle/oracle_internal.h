// oracle/oracle_internal.h -- ORACLE (test infrastructure) internal C++ declarations.
#ifndef SDPL_ORACLE_INTERNAL_H
#define SDPL_ORACLE_INTERNAL_H
#include "oracle.h"
#include <vector>
#include <cstddef>

namespace orc {
int cv_round(float v);
int cv_round(double v);
int cv_floor(double v);
int cv_floor(float v);
int cv_ceil(double v);
int reflect101(int p, int len);
void resize_linear_u8(const uint8_t* src, int sw, int sh, int sstride, uint8_t* dst, int dw, int dh, int dstride);
void resize_linear_exact_u8(const uint8_t* src, int sw, int sh, int sstride, uint8_t* dst, int dw, int dh, int dstride,
                            double inv_scale_x, double inv_scale_y);
void border_reflect101_u8(const uint8_t* src, int w, int h, int sstride, uint8_t* dst, int border, int dstride);
void gaussian_blur_u8(const uint8_t* src, int w, int h, int sstride, uint8_t* dst, int dstride, int kind);
void pyrdown_u8(const uint8_t* src, int w, int h, int sstride, uint8_t* dst, int dw, int dh, int dstride);
void sobel3_s16(const uint8_t* src, int w, int h, int sstride, int16_t* dx, int16_t* dy);
float fast_atan2(float y, float x);
int fast_score(const uint8_t* p, int stride);
int fast9_nms(const uint8_t* img, int w, int h, int stride, int threshold, std::vector<int>& xs, std::vector<int>& ys,
              std::vector<int>& sc);

// Padded pyramid level shared by the ORB and LSD pyramids (same construction, 19 px reflect-101 border).
struct PaddedLevel {
  int w = 0, h = 0;            // interior size
  std::vector<uint8_t> buf;    // (w+38) x (h+38)
  int stride() const { return w + 38; }
  uint8_t* roi() { return buf.data() + 19 * stride() + 19; }
  const uint8_t* roi() const { return buf.data() + 19 * stride() + 19; }
};
// level 0 = copy + border, level l = INTER_LINEAR resize of level l-1 interior + border (isolated)
void build_padded_pyramid(const uint8_t* img, int w, int h, int stride, const std::vector<float>& inv_scale,
                          std::vector<PaddedLevel>& levels);
// EDLines back-end (ed_oracle.cpp): lines of one padded level, optional intermediates for the tests
struct EdDebug { std::vector<uint8_t> smooth, edge; std::vector<int> seg_px, seg_off; int err = 0; };
int ed_detect(const uint8_t* roi, int w, int h, int stride, std::vector<float>& seg, EdDebug* dbg);


// OpenCV LineSegmentDetector restatement (lsd_oracle.cpp)
int lsd_detect(const uint8_t* img, int w, int h, int stride, int refine, double scale, double sigma_scale, double quant,
               double ang_th, double log_eps, double density_th, int n_bins, int tie_mode, std::vector<float>& lines);
int lbd_compute(const uint8_t* img, int w, int h, int stride, const orc_keyline* kls, int n, uint8_t* desc, float* fdesc);
}  // namespace orc
#endif
