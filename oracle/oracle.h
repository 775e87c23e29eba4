/*
 * oracle/oracle.h -- C interface of the CPU ORACLE (test infrastructure, NOT the product).
 *
 * The oracle is a single-threaded CPU restatement of the reference front-end
 * (argyrissm/SDPL-SLAM: src/ORBextractor.cc, src/Lineextractor.cc,
 * 3rdparty/line_descriptor/src/{LSDDetector_custom,binary_descriptor_custom}.cpp,
 * bitops_custom.hpp, and for Lineextractor's EDLines back-end src/ED_Lib/{ED,EDLines,NFA}.cpp) with the OpenCV 3.4 primitives it calls restated from their
 * bit-exact integer/float formulas.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.
 *
 * Parity status: PINNED for the extractor / matcher path -- against the reference's own sources compiled unmodified over an
 * OpenCV stand-in (oracle/refshim -> oracle/_ref/libsdpl_ref.so, tests/test_oracle_vs_ref.py: byte for byte on every BASELINE
 * geometry) and, at the OpenCV-primitive boundary, against python cv2 4.13 (tests/test_oracle_vs_cv2.py, tests/golden/).
 * The Frame post-processing functions (post_oracle.cpp) are PINNED as well, for everything Frame::Frame itself runs: the same
 * library holds src/Frame.cc compiled unmodified, and test_frame_post_processing_equals_reference_frame_constructor compares every
 * vector the reference's constructor leaves behind (line filters, point / line correspondences, object sampling, grid,
 * GetFeaturesInArea) byte for byte.  Only the three functions restated from src/MapPoint.cc (distinctive descriptor, PredictScale)
 * and the window descriptor search are "parity unpinned": MapPoint.cc is not part of the reference's own build (CMakeLists.txt:62
 * lists no such file, ORBmatcher.h does not exist), so no compiled form of it exists anywhere; they are checked against an
 * independent numpy restatement (tests/test_oracle_post.py).
 */
#ifndef SDPL_ORACLE_H
#define SDPL_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct { float x, y, size, angle, response; int32_t octave, class_id; } orc_keypoint; /* cv::KeyPoint */
typedef struct {
  float angle; int32_t class_id, octave; float pt_x, pt_y, response, size;
  float sx, sy, ex, ey, sx_oct, sy_oct, ex_oct, ey_oct, length; int32_t num_pixels;
} orc_keyline;                                                                                /* KeyLine, 68 B */
typedef struct { int32_t query, train, img; float distance; } orc_dmatch;                     /* cv::DMatch */

/* ---- primitives (each pinned against cv2 in tests/test_oracle_vs_cv2.py) ---- */
void orc_resize_linear_u8(const uint8_t* src, int sw, int sh, int sstride,
                          uint8_t* dst, int dw, int dh, int dstride);
/* fx,fy: the inv_scale passed to cv::resize (0 -> dw/sw, dh/sh as when a dsize is given) */
void orc_resize_linear_exact_u8(const uint8_t* src, int sw, int sh, int sstride,
                                uint8_t* dst, int dw, int dh, int dstride, double fx, double fy);
void orc_border_reflect101_u8(const uint8_t* src, int w, int h, int sstride,
                              uint8_t* dst, int border, int dstride);
/* kind: 0 = 7x7 sigma 2 (ORB), 1 = 5x5 sigma 1 (LBD), 2 = 7x7 sigma 0.75 (LSD 0.6/0.8) */
void orc_gaussian_blur_u8(const uint8_t* src, int w, int h, int sstride, uint8_t* dst, int dstride, int kind);
void orc_pyrdown_u8(const uint8_t* src, int w, int h, int sstride, uint8_t* dst, int dw, int dh, int dstride);
void orc_sobel3_s16(const uint8_t* src, int w, int h, int sstride, int16_t* dx, int16_t* dy);
float orc_fast_atan2(float y, float x);
/* FAST-9/16 with non-max suppression on a stand-alone image; returns count, fills x,y,score (row-major order) */
int orc_fast9_nms(const uint8_t* img, int w, int h, int stride, int threshold,
                  int* xs, int* ys, int* scores, int cap);
int orc_cv_round_f(float v);
int orc_cv_round_d(double v);

/* ---- ORB extractor (src/ORBextractor.cc) ---- */
typedef struct orc_orb orc_orb;
orc_orb* orc_orb_create(int nfeatures, float scale, int nlevels, int ini_th, int min_th);
void orc_orb_destroy(orc_orb*);
/* scale tables + per level quota + umax[16] */
void orc_orb_tables(const orc_orb*, float* sf, float* isf, float* sig2, float* isig2, int* quota, int* umax);
/* full operator(): returns number of keypoints (<= cap) or <0 on error; desc may be NULL */
int orc_orb_extract(orc_orb*, const uint8_t* img, int w, int h, int stride,
                    orc_keypoint* kps, uint8_t* desc, int cap);
/* introspection after extract (for stage-by-stage parity) */
void orc_orb_level_size(const orc_orb*, int level, int* w, int* h);
/* padded level plane: (w+38)x(h+38), row stride = w+38 */
const uint8_t* orc_orb_level_padded(const orc_orb*, int level);
const uint8_t* orc_orb_level_blurred(const orc_orb*, int level); /* w x h, stride w */
/* candidates handed to DistributeOctTree (border-relative coords) */
int orc_orb_level_candidates(const orc_orb*, int level, int* xs, int* ys, int* resp, int cap);
int orc_orb_level_count(const orc_orb*, int level);

/* ---- Line extractor: LSD + KeyLine + LBD ---- */
typedef struct orc_line orc_line;
orc_line* orc_line_create(int nfeatures, int refine, float lsd_scale, int nlevels, float scale, int extractor);
void orc_line_destroy(orc_line*);
int orc_line_extract(orc_line*, const uint8_t* img, int w, int h, int stride,
                     orc_keyline* kls, uint8_t* desc, int cap);
void orc_line_tables(const orc_line*, float* sf, float* isf, float* sig2, float* isig2);
/* stand-alone OpenCV-LSD restatement on one u8 image; lines = x1,y1,x2,y2 floats; tie_mode 0 = row-major stable */
int orc_lsd_detect(const uint8_t* img, int w, int h, int stride, int refine, double scale, double sigma_scale,
                   double quant, double ang_th, double log_eps, double density_th, int n_bins,
                   int tie_mode, float* lines, int cap);
/* stage introspection of the last orc_lsd_detect call (thread-unsafe, test use only) */
int orc_lsd_last_scaled(uint8_t* dst, int cap, int* w, int* h);
/* LBD on given keylines: desc L x 32 ; also float descriptor (L x 72) if fdesc != NULL */
int orc_lbd_compute(const uint8_t* img, int w, int h, int stride, const orc_keyline* kls, int n,
                    uint8_t* desc, float* fdesc);

/* ---- matcher ---- */
int orc_hamming256(const uint8_t* a, const uint8_t* b);
/* brute-force top-2: best/second (train=-1, distance=256.. when missing). strict <, ties -> lowest train index */
void orc_match_knn2(const uint8_t* q, int nq, const uint8_t* t, int nt, orc_dmatch* best, orc_dmatch* second);
/* ratio + max-dist filter on top of knn2: out[i].train = -1 if rejected; returns accepted count */
int orc_match_ratio(const uint8_t* q, int nq, const uint8_t* t, int nt, float ratio, int max_dist, orc_dmatch* out);
/* radius search: per query count of train with dist <= radius, and up to k nearest (ascending dist, then index) */
void orc_match_radius(const uint8_t* q, int nq, const uint8_t* t, int nt, int radius, int k,
                      int* counts, orc_dmatch* out);

/* ---- Frame post-processing (src/Frame.cc), oracle/post_oracle.cpp ---- */
int orc_post_sample_objects(const int32_t* mask, const float* depth, const float* flow, int w, int h, int step, float th_depth_obj,
                            orc_keypoint* keys, orc_keypoint* corres, float* flow_next, float* depth_out, int32_t* label, int cap);
int orc_post_filter_lines(const orc_keyline* kls, int n, const int32_t* mask, const float* depth, int w, int h, orc_keyline* out,
                          int32_t* keep_idx);
int orc_post_point_corres(const orc_keypoint* kps, int n, const int32_t* mask, const float* depth, const float* flow, int w, int h,
                          float th_depth, orc_keypoint* stat, orc_keypoint* corres, float* flow_next, float* stat_depth, int32_t* src_idx);
int orc_post_line_corres(const orc_keyline* kls, int n, const int32_t* mask, const float* depth, const float* flow, int w, int h,
                         float th_depth, orc_keyline* obj, int32_t* n_obj, orc_keyline* stat, orc_keyline* corres, float* flow_next,
                         double* inf_line, float* stat_depth, int32_t* src_idx);
int orc_post_features_in_area(const orc_keypoint* kps, int w, int h, int grid_cols, int grid_rows, const int32_t* cell_start,
                              const int32_t* items, float x, float y, float r, int minLevel, int maxLevel, int32_t* out, int cap);
void orc_post_distinctive_descriptors(const uint8_t* desc, const int32_t* start, int n_points, int32_t* best_idx, uint8_t* out_desc);
void orc_post_predict_scale(const float* max_distance, const float* current_dist, int n, float log_scale_factor, int n_levels, int32_t* out);
void orc_post_search_area(const orc_keypoint* kps, const uint8_t* desc, int w, int h, int grid_cols, int grid_rows, const int32_t* cell_start,
                          const int32_t* items, float x, float y, float r, int minLevel, int maxLevel, const uint8_t* qdesc, int32_t* out5);
void orc_post_grid(const orc_keypoint* kps, int n, int w, int h, int grid_cols, int grid_rows, int32_t* cell_start, int32_t* items);

#ifdef __cplusplus
}
#endif
#endif
