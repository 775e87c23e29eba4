/*
 * sdpl_frontend.h -- C ABI of the B200-native (sm_100a CUDA) SDPL-SLAM feature front-end.
 *
 * This is the drop-in boundary for the per-frame front-end of argyrissm/SDPL-SLAM: every entry point
 * below replaces one C++ interface of the reference (cited as file:line, relative to the reference
 * root).  Plain pointers and sizes only; int status codes; no exceptions cross the boundary.
 * One handle = one CUDA device + one stream + one pre-allocated device arena.  A handle is not
 * thread-safe; different handles are independent.  There is NO CPU fallback: every call needs a
 * CUDA device and returns SDPL_ERR_CUDA otherwise.
 *
 * Host <-> device: the *_extract / *_match entry points take HOST buffers (they are what
 * Frame::ExtractORB / Frame::ExtractLines would call, src/Frame.cc:927-949).  The *_dev entry points
 * take DEVICE pointers and leave results on the device (throughput path, batch of frames resident in HBM).
 */
#ifndef SDPL_FRONTEND_H
#define SDPL_FRONTEND_H
#include <stdint.h>
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

/* == cv::KeyPoint (28 B POD): pt.x, pt.y, size, angle, response, octave, class_id */
typedef struct { float x, y, size, angle, response; int32_t octave, class_id; } sdpl_keypoint;
/* == cv::line_descriptor::KeyLine (68 B POD), 3rdparty/line_descriptor/include/line_descriptor/descriptor_custom.hpp */
typedef struct {
  float angle; int32_t class_id, octave; float pt_x, pt_y, response, size;
  float sx, sy, ex, ey, sx_oct, sy_oct, ex_oct, ey_oct, length; int32_t num_pixels;
} sdpl_keyline;
/* == cv::DMatch (16 B POD): queryIdx, trainIdx, imgIdx, distance */
typedef struct { int32_t query, train, img; float distance; } sdpl_dmatch;

enum {
  SDPL_OK = 0,
  SDPL_ERR_ARG = 1,        /* bad argument (null pointer, non-positive size, unsupported parameter) */
  SDPL_ERR_CUDA = 2,       /* CUDA runtime failure / no device (there is no CPU fallback) */
  SDPL_ERR_CAPACITY = 3,   /* caller-provided output capacity too small (n_out holds the needed count) */
  SDPL_ERR_OVERFLOW = 4,   /* an internal device buffer overflowed (reported, never silently truncated) */
  SDPL_ERR_UNSUPPORTED = 5 /* configuration outside what the kernels implement (e.g. lsd_scale other than 0.8, more than 4 octaves) */
};
const char* sdpl_strerror(int code);
/* last CUDA / internal error text of the calling thread (empty string if none) */
const char* sdpl_last_error(void);
int sdpl_device_count(void);

/* ------------------------------------------------------------------------------------------------
 * ORB extractor -- replaces SDPL_SLAM::ORBextractor (include/ORBextractor.h:33-99, src/ORBextractor.cc)
 * ---------------------------------------------------------------------------------------------- */
typedef struct sdpl_orb sdpl_orb;
/* ORBextractor::ORBextractor(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST), src/ORBextractor.cc:399-459 */
int sdpl_orb_create(sdpl_orb** h, int nfeatures, float scale, int nlevels, int ini_th, int min_th, int device);
void sdpl_orb_destroy(sdpl_orb* h);
/* GetLevels / GetScaleFactors / GetInverseScaleFactors / GetScaleSigmaSquares / GetInverseScaleSigmaSquares
 * (include/ORBextractor.h:49-69); each array has nlevels entries; any pointer may be NULL */
int sdpl_orb_levels(const sdpl_orb* h);
int sdpl_orb_tables(const sdpl_orb* h, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2);
/* mnFeaturesPerLevel (src/ORBextractor.cc:424-435) and umax[16] (:443-458), for parity tests */
int sdpl_orb_quota(const sdpl_orb* h, int* per_level, int* umax16);
/* upper bound of keypoints one frame can produce (sum over levels of quota+3) */
int sdpl_orb_max_keypoints(const sdpl_orb* h);
/* ORBextractor::operator()(image, mask (ignored), keypoints, descriptors), src/ORBextractor.cc:1035-1110.
 * img: HOST u8 grayscale, row stride in bytes.  kps/desc: HOST outputs, capacity rows (desc is capacity x 32).
 * Empty image (w<=0 or h<=0 or img==NULL) -> *n_out = 0, SDPL_OK (the reference returns silently, :1038). */
int sdpl_orb_extract(sdpl_orb* h, const uint8_t* img, int w, int h_, int stride,
                     sdpl_keypoint* kps, uint8_t* desc, int capacity, int* n_out);
/* Batch of nframes same-size frames.  imgs: HOST pointer to frame 0, frames frame_stride bytes apart.
 * Outputs: frame f writes kps[f*capacity ...], desc[(f*capacity)*32 ...], n_out[f]. */
int sdpl_orb_extract_batch(sdpl_orb* h, const uint8_t* imgs, int nframes, int w, int h_, int stride, size_t frame_stride,
                           sdpl_keypoint* kps, uint8_t* desc, int capacity, int* n_out);
/* Same with DEVICE pointers for input and outputs; asynchronous on the handle's stream unless sync != 0.
 * n_out is a DEVICE int array [nframes]. */
int sdpl_orb_extract_batch_dev(sdpl_orb* h, const uint8_t* d_imgs, int nframes, int w, int h_, int stride,
                               size_t frame_stride, sdpl_keypoint* d_kps, uint8_t* d_desc, int capacity, int* d_n_out,
                               int sync);
/* mvImagePyramid (include/ORBextractor.h:71): copies the padded (w+38)x(h+38) plane of `level` of frame `frame` of
 * the last call to a HOST buffer with row stride out_stride; w_out/h_out receive the interior size. out may be NULL. */
int sdpl_orb_pyramid_level(sdpl_orb* h, int frame, int level, uint8_t* out, int out_stride, int* w_out, int* h_out);
/* stage introspection for parity tests: blurred level (w x h), FAST/NMS candidates given to DistributeOctTree */
int sdpl_orb_blurred_level(sdpl_orb* h, int frame, int level, uint8_t* out, int out_stride);
int sdpl_orb_candidates(sdpl_orb* h, int frame, int level, int* xs, int* ys, int* resp, int capacity, int* n_out);
int sdpl_orb_level_counts(sdpl_orb* h, int frame, int* per_level);
/* number of kernels the last extract call launched (for bench.py's gpu_launches) */
int sdpl_orb_last_launches(const sdpl_orb* h);

/* ------------------------------------------------------------------------------------------------
 * Line extractor -- replaces SDPL_SLAM::Lineextractor (include/Lineextractor.h:51-87, src/Lineextractor.cc:42-99)
 * = LSDDetectorC::detect (3rdparty/line_descriptor/src/LSDDetector_custom.cpp:254-369, cv::LineSegmentDetector inside)
 * + BinaryDescriptor::compute (binary_descriptor_custom.cpp:524-687, computeLBD :1026-1372)
 * ---------------------------------------------------------------------------------------------- */
typedef struct sdpl_line sdpl_line;
/* Lineextractor(lsd_nfeatures, lsd_refine, lsd_scale, nlevels, scale, extractor), src/Lineextractor.cc:36-40.
 * extractor 0 = LSD (LSDDetectorC::detect, src/Lineextractor.cc:47-98), 1 = EDLines (LSDDetectorC::detect_ED, :100-135 ->
 * LSDDetector_custom.cpp:372-461 -> 3rdparty/line_descriptor/src/ED_Lib/EDLines.cpp:8-70); anything else: SDPL_ERR_UNSUPPORTED.
 * lsd_refine and lsd_scale are ignored by the EDLines back-end, as in the reference. */
int sdpl_line_create(sdpl_line** h, int nfeatures, int refine, float lsd_scale, int nlevels, float scale, int extractor,
                     int device);
void sdpl_line_destroy(sdpl_line* h);
/* mvScaleFactor_l / mvInvScaleFactor_l / mvLevelSigma2_l / mvInvLevelSigma2_l (include/Lineextractor.h:66-75) */
int sdpl_line_levels(const sdpl_line* h);
int sdpl_line_tables(const sdpl_line* h, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2);
/* Lineextractor::operator()(image, mask (ignored), keylines, descriptors_line) */
int sdpl_line_extract(sdpl_line* h, const uint8_t* img, int w, int h_, int stride,
                      sdpl_keyline* kls, uint8_t* desc, int capacity, int* n_out);
int sdpl_line_extract_batch(sdpl_line* h, const uint8_t* imgs, int nframes, int w, int h_, int stride, size_t frame_stride,
                            sdpl_keyline* kls, uint8_t* desc, int capacity, int* n_out);
int sdpl_line_extract_batch_dev(sdpl_line* h, const uint8_t* d_imgs, int nframes, int w, int h_, int stride,
                                size_t frame_stride, sdpl_keyline* d_kls, uint8_t* d_desc, int capacity, int* d_n_out,
                                int sync);
/* BinaryDescriptor::compute(image, keylines, descriptors) alone, on caller-provided keylines (HOST buffers) */
int sdpl_line_lbd_compute(sdpl_line* h, const uint8_t* img, int w, int h_, int stride,
                          const sdpl_keyline* kls, int n, uint8_t* desc);
/* raw cv::LineSegmentDetector::detect output (x1,y1,x2,y2 per line) of pyramid level `octave` of the last
 * single-frame call, for parity tests */
int sdpl_line_lsd_segments(sdpl_line* h, int frame, int octave, float* xyxy, int capacity, int* n_out);
/* introspection: rectangles handed to the NFA stage in seed order, 8 doubles each {x1,y1,x2,y2,width,p,accepted,tag} */
int sdpl_line_debug_pending(sdpl_line* h, int frame, int octave, double* out, int capacity, int* n_out);
/* introspection: region-growing kernel counters of one (frame, octave): cycles in {select, speculate, evaluate+commit, re-run},
 * then {waves, re-runs, dead seeds, seeds}; after sdpl_line_debug_grow_detail(h, 1), [8..15] = thread 0's cycles in {phase A, phase B busy, phase B
 * barrier wait, region growth, rectangle fits, refine tolerance}, its growth steps and its wait at the phase-A barrier */
int sdpl_line_debug_grow_profile(sdpl_line* h, int frame, int octave, long long* out16);
/* switch the detailed counters [8..15] on / off (off by default: they cost about 3 % of the region-growing kernel) */
int sdpl_line_debug_grow_detail(sdpl_line* h, int on);
int sdpl_line_last_launches(const sdpl_line* h);
/* test / tuning knob, LSD region-growing schedule (bits 0-1; further bits select kernel variants, see line.cu): 0 = speculative waves,
 * one CTA of several warps per (frame, octave) (default), 1 = strictly one seed at a time, 3 = single-warp waves of 32 seeds; all give
 * identical results; 2 is not a schedule any more (SDPL_ERR_ARG) */
int sdpl_line_set_serial(sdpl_line* h, int on);
/* Lineextractor::extractor (include/Lineextractor.h:60) of an existing object: 0 = LSD, 1 = EDLines; takes effect at the next call */
int sdpl_line_set_extractor(sdpl_line* h, int extractor);

/* ------------------------------------------------------------------------------------------------
 * Descriptor matcher -- 256-bit Hamming (cv::line_descriptor::match, bitops_custom.hpp:86-99) brute force;
 * surface of BinaryDescriptorMatcher::match / knnMatch / radiusMatch (binary_descriptor_matcher.cpp:197-504).
 * best/second by strict '<' over ascending train index (ties -> lowest train index). distance = (float)hamming.
 * Missing neighbour (nt < 2): train = -1, distance = 257.
 * ---------------------------------------------------------------------------------------------- */
typedef struct sdpl_matcher sdpl_matcher;
int sdpl_matcher_create(sdpl_matcher** h, int device);
void sdpl_matcher_destroy(sdpl_matcher* h);
int sdpl_match_knn2(sdpl_matcher* h, const uint8_t* q, int nq, const uint8_t* t, int nt,
                    sdpl_dmatch* best, sdpl_dmatch* second);
/* knn2 + ratio test d1 < ratio*d2 and d1 <= max_dist: rejected rows get train = -1; *n_acc = accepted count */
int sdpl_match_ratio(sdpl_matcher* h, const uint8_t* q, int nq, const uint8_t* t, int nt, float ratio, int max_dist,
                     sdpl_dmatch* out, int* n_acc);
/* radius search: counts[i] = #train within radius; out[i*k ..] = k nearest within radius (ascending distance, index) */
int sdpl_match_radius(sdpl_matcher* h, const uint8_t* q, int nq, const uint8_t* t, int nt, int radius, int k,
                      int* counts, sdpl_dmatch* out);
/* general k against an explicit train set: out[i*k + j] = j-th nearest train row of query i, ascending (distance, train index);
 * missing neighbours (nt < k) have train = -1, distance = 257.  BinaryDescriptorMatcher::knnMatch(query, train, matches, k),
 * binary_descriptor_matcher.cpp:258-335 */
int sdpl_match_knn(sdpl_matcher* h, const uint8_t* q, int nq, const uint8_t* t, int nt, int k, sdpl_dmatch* out);
/* The train set kept by the matcher -- BinaryDescriptorMatcher::add / train / clear (descriptor_custom.hpp:1015-1126,
 * binary_descriptor_matcher.cpp:127-194): every add() appends one image's descriptors (HOST pointer; _dev: DEVICE pointer) to a
 * set resident on the device; train() is where the reference builds its hash tables (nothing to do for brute force);
 * queries against the stored set return trainIdx = row in the concatenation of all added images and imgIdx = the image it came
 * from, exactly as the reference composes its DMatch (binary_descriptor_matcher.cpp:161-175, 381-401). */
int sdpl_matcher_add(sdpl_matcher* h, const uint8_t* desc, int n);
int sdpl_matcher_add_dev(sdpl_matcher* h, const uint8_t* d_desc, int n);
int sdpl_matcher_train(sdpl_matcher* h);
int sdpl_matcher_clear(sdpl_matcher* h);
int sdpl_matcher_train_size(const sdpl_matcher* h, int* n_desc, int* n_imgs);
/* match / knnMatch(query, matches, k) against the stored set (binary_descriptor_matcher.cpp:127-194, 339-425): out[i*k + j] */
int sdpl_matcher_knn(sdpl_matcher* h, const uint8_t* q, int nq, int k, sdpl_dmatch* out);
/* radiusMatch(query, matches, maxDistance) against the stored set (binary_descriptor_matcher.cpp:507-590): as sdpl_match_radius */
int sdpl_matcher_radius(sdpl_matcher* h, const uint8_t* q, int nq, int radius, int k, int* counts, sdpl_dmatch* out);
/* DEVICE pointer and row count of the stored set, for the batched device entry points (t_stride = 0: one set for all problems) */
int sdpl_matcher_train_dev(const sdpl_matcher* h, const uint8_t** d_train, int* n_desc);
/* Batched DEVICE variant: npairs problems; problem p matches d_q + p*q_stride (nq[p] rows) against
 * d_t + p*t_stride (nt[p] rows); d_nq/d_nt are DEVICE int arrays; outputs [p*max_q + i]. */
int sdpl_match_knn2_batch_dev(sdpl_matcher* h, const uint8_t* d_q, const int* d_nq, size_t q_stride, const uint8_t* d_t,
                              const int* d_nt, size_t t_stride, int npairs, int max_q, int max_t, sdpl_dmatch* d_best,
                              sdpl_dmatch* d_second, int sync);
/* Batched DEVICE ratio test on knn2 results: d_n_acc[p] = accepted count of problem p; d_out (may be NULL) receives the
 * filtered best matches (train = -1 where rejected), laid out like d_best. */
int sdpl_match_ratio_batch_dev(sdpl_matcher* h, const sdpl_dmatch* d_best, const sdpl_dmatch* d_second, const int* d_nq, int npairs,
                               int max_q, float ratio, int max_dist, sdpl_dmatch* d_out, int* d_n_acc, int sync);
int sdpl_matcher_last_launches(const sdpl_matcher* h);

/* read-and-clear the device-side overflow flag of the last asynchronous (*_dev, sync == 0) calls; synchronises the stream.
 * SDPL_OK, or SDPL_ERR_OVERFLOW when an internal buffer (FAST candidates, quadtree nodes, pending rectangles) overflowed. */
int sdpl_orb_check(sdpl_orb* h);
int sdpl_line_check(sdpl_line* h);
/* the same flag copied to *host_flag (pinned host memory) on `stream` without synchronising or clearing; sticky */
int sdpl_orb_peek_error_async(sdpl_orb* h, void* stream, int* host_flag);
int sdpl_line_peek_error_async(sdpl_line* h, void* stream, int* host_flag);
/* the same flag copied to *dst (pinned host or device memory) AND cleared, both on the handle's own stream: enqueue it right
 * after a batch and the flag belongs to that batch alone (what sdpl_frontend_submit does) */
int sdpl_orb_take_error_async(sdpl_orb* h, int* dst);
int sdpl_line_take_error_async(sdpl_line* h, int* dst);

/* ------------------------------------------------------------------------------------------------
 * The whole per-frame front-end in one call: Frame::Frame's ExtractORB + ExtractLines (src/Frame.cc:314,328 -> :927-949)
 * plus frame-to-frame descriptor association of points and lines (frame t against frame t-1; frame 0 of a call against
 * the last frame of the previous call, none on the very first call / after sdpl_frontend_reset).
 * One upload of the frames; ORB and line pipelines run concurrently on two streams, the two matchers on two more.
 * (Tuning knob read at creation: environment SDPL_FE_LINE_PRIO=high gives the line stream a higher priority than the
 * others; measured 1 % slower than equal priorities on B200, the default.)
 * ---------------------------------------------------------------------------------------------- */
typedef struct sdpl_frontend sdpl_frontend;
typedef struct { int32_t n_kp, n_lines, n_pt_matches, n_ln_matches; } sdpl_frame_stats;
int sdpl_frontend_create(sdpl_frontend** h, int nfeatures, float scale, int nlevels, int ini_th, int min_th, int lsd_nfeatures,
                         int lsd_refine, float lsd_scale, int lsd_levels, float lsd_pyr_scale, float ratio, int max_dist, int device);
void sdpl_frontend_destroy(sdpl_frontend* h);
/* rows per frame of the output arrays of sdpl_frontend_process.  kp_capacity = sdpl_orb_max_keypoints; kl_capacity =
 * lsd_nfeatures when that is > 0, else 2048 by default (the reference has no cap then; a frame with more segments makes
 * collect return SDPL_ERR_CAPACITY with the raw count in stats, its first kl_capacity lines are still matched and returned) */
int sdpl_frontend_capacities(const sdpl_frontend* h, int* kp_capacity, int* kl_capacity);
/* change kl_capacity (rows per frame of the key-line outputs); only while no batch is in flight */
int sdpl_frontend_set_line_capacity(sdpl_frontend* h, int kl_capacity);
/* the line back-end of the front-end's Lineextractor (0 = LSD, the default; 1 = EDLines); only while nothing is in flight */
int sdpl_frontend_set_line_extractor(sdpl_frontend* h, int extractor);
/* forget the previous frame (start of a new sequence) */
int sdpl_frontend_reset(sdpl_frontend* h);
/* imgs: HOST, n frames frame_stride bytes apart.  Outputs (HOST): frame f owns rows [f*cap, f*cap + count):
 * kps/desc/pt_matches with cap = kp_capacity, kls/ldesc/ln_matches with cap = kl_capacity; stats[f] holds the counts.
 * pt_matches[f*cap + i] is the ratio-filtered best match of keypoint i of frame f in frame f-1 (train = -1: rejected). */
int sdpl_frontend_process(sdpl_frontend* h, const uint8_t* imgs, int nframes, int w, int h_, int stride, size_t frame_stride,
                          sdpl_keypoint* kps, uint8_t* desc, sdpl_keyline* kls, uint8_t* ldesc, sdpl_dmatch* pt_matches,
                          sdpl_dmatch* ln_matches, sdpl_frame_stats* stats);
/* Pipelined form: submit enqueues a batch (upload, both pipelines, matching) and returns at once -- `imgs` must stay valid
 * until that batch is collected; collect waits for the OLDEST submitted batch and copies its results to the host on a
 * separate stream, overlapping the kernels of a batch submitted after it.  At most two batches in flight.
 * process == submit + collect. */
int sdpl_frontend_submit(sdpl_frontend* h, const uint8_t* imgs, int nframes, int w, int h_, int stride, size_t frame_stride);
int sdpl_frontend_collect(sdpl_frontend* h, sdpl_keypoint* kps, uint8_t* desc, sdpl_keyline* kls, uint8_t* ldesc, sdpl_dmatch* pt_matches,
                          sdpl_dmatch* ln_matches, sdpl_frame_stats* stats, int* n_frames);
int sdpl_frontend_pending(const sdpl_frontend* h);
int sdpl_frontend_last_launches(const sdpl_frontend* h);

/* Per-stage device timing (CUDA events on the handle's stream).  set_profiling(1) makes every following call record
 * one event per stage; stage_times returns the stages of the LAST call: ms[i], names[i] (static strings), launches[i]
 * = kernels launched in stage i; return value = number of stages written (<= cap).  Synchronises on the last event. */
int sdpl_orb_set_profiling(sdpl_orb* h, int on);
int sdpl_orb_stage_times(sdpl_orb* h, float* ms, const char** names, int* launches, int cap);
int sdpl_line_set_profiling(sdpl_line* h, int on);
int sdpl_line_stage_times(sdpl_line* h, float* ms, const char** names, int* launches, int cap);
int sdpl_matcher_set_profiling(sdpl_matcher* h, int on);
int sdpl_matcher_stage_times(sdpl_matcher* h, float* ms, const char** names, int* launches, int cap);

/* ------------------------------------------------------------------------------------------------
 * Frame post-processing on the device -- the loops Frame::Frame runs on the extractor outputs (src/Frame.cc), SURVEY.md 8f rows
 * 1 and 2.  Inputs are the reference's per-frame planes as they sit in cv::Mat: maskSEM int32 [h][w], imDepth float [h][w],
 * imFlow float [h][w][2]; the *_dev entry points take DEVICE pointers to nframes consecutive planes and per-frame feature blocks
 * of `capacity` rows (the layout the extractors' *_batch_dev calls produce), and are asynchronous on the handle's stream unless
 * sync != 0.  Every output list is in the order the reference's sequential loops push it.  Counts may exceed the capacity (only
 * `capacity` rows are written).
 * ---------------------------------------------------------------------------------------------- */
typedef struct sdpl_post sdpl_post;
int sdpl_post_create(sdpl_post** h, int device);
void sdpl_post_destroy(sdpl_post* h);
int sdpl_post_set_stream(sdpl_post* h, void* cuda_stream);
int sdpl_post_last_launches(const sdpl_post* h);
int sdpl_post_set_profiling(sdpl_post* h, int on);
int sdpl_post_stage_times(sdpl_post* h, float* ms, const char** names, int* launches, int cap);
/* semi-dense features on objects, src/Frame.cc:769-809: every step-th pixel with a non-zero mask label, 0 < depth < th_depth_obj
 * and a flow that stays inside the image.  keys = mvObjKeys (x = column, y = row), corres = mvObjCorres, flow_next = mvObjFlowNext
 * (2 floats), depth_out = mvObjDepth, label = vSemObjLabel */
int sdpl_post_sample_objects_dev(sdpl_post* h, const int32_t* d_mask, const float* d_depth, const float* d_flow, int nframes, int w, int h_,
                                 int step, float th_depth_obj, sdpl_keypoint* d_keys, sdpl_keypoint* d_corres, float* d_flow_next,
                                 float* d_depth_out, int32_t* d_label, int capacity, int* d_n, int sync);
/* the same for ONE frame with HOST buffers (uploads the planes, downloads the lists) */
int sdpl_post_sample_objects(sdpl_post* h, const int32_t* mask, const float* depth, const float* flow, int w, int h_, int step,
                             float th_depth_obj, sdpl_keypoint* keys, sdpl_keypoint* corres, float* flow_next, float* depth_out,
                             int32_t* label, int capacity, int* n_out);
/* the two erase loops on mvKeys_Line, src/Frame.cc:349-389 (depth step at the mid point; end points on different mask labels).
 * keep_idx[k] = row of the k-th surviving line in the input, i.e. the row of its LBD descriptor */
int sdpl_post_filter_lines_dev(sdpl_post* h, const int32_t* d_mask, const float* d_depth, int nframes, int w, int h_, const sdpl_keyline* d_kls,
                               const int* d_n_in, int capacity, sdpl_keyline* d_out, int32_t* d_keep_idx, int* d_n_out, int sync);
/* static point correspondences from the detected features (UseSampleFea == 0), src/Frame.cc:482-512, with their depths (:728-745):
 * stat = mvStatKeysTmp, corres = mvCorres, flow_next = mvFlowNext (2 floats), stat_depth = mvStatDepthTmp, src_idx = row in the input */
int sdpl_post_point_corres_dev(sdpl_post* h, const int32_t* d_mask, const float* d_depth, const float* d_flow, int nframes, int w, int h_,
                               const sdpl_keypoint* d_kps, const int* d_n_in, int capacity, float th_depth, sdpl_keypoint* d_stat,
                               sdpl_keypoint* d_corres, float* d_flow_next, float* d_stat_depth, int32_t* d_src_idx, int* d_n_out, int sync);
/* line correspondences, src/Frame.cc:513-604, with their depths (:746-763): obj = mvObjKeys_Line (both end points on one object),
 * stat = mvStatKeysLineTmp, corres = mvCorresLine, flow_next = mvFlowNext_Line (4 floats: start x, y, end x, y), inf_line =
 * mvInfiniteLinesCorr (3 doubles), stat_depth = mvStatDepthLineTmp (2 floats) */
int sdpl_post_line_corres_dev(sdpl_post* h, const int32_t* d_mask, const float* d_depth, const float* d_flow, int nframes, int w, int h_,
                              const sdpl_keyline* d_kls, const int* d_n_in, int capacity, float th_depth, sdpl_keyline* d_obj, int* d_n_obj,
                              sdpl_keyline* d_stat, sdpl_keyline* d_corres, float* d_flow_next, double* d_inf_line, float* d_stat_depth,
                              int32_t* d_src_idx, int* d_n_out, int sync);
/* AssignFeaturesToGrid / PosInGrid, src/Frame.cc:910-925, 1023-1035 (undistorted key points = key points, mDistCoef[0] == 0):
 * cell c = posX * grid_rows + posY; items[cell_start[c] .. cell_start[c+1]) = key point indices of the cell in ascending order.
 * d_cell_start: [nframes][grid_cols * grid_rows + 1], d_items: [nframes][capacity] */
int sdpl_post_grid_dev(sdpl_post* h, int nframes, int w, int h_, const sdpl_keypoint* d_kps, const int* d_n_in, int capacity, int grid_cols,
                       int grid_rows, int32_t* d_cell_start, int32_t* d_items, int sync);

/* Frame::GetFeaturesInArea, src/Frame.cc:970-1023, on the grid of sdpl_post_grid_dev: d_queries = [nframes][nq][5] floats {x, y, r,
 * minLevel, maxLevel} (maxLevel < 0: no upper bound); d_out[(f*nq + q)*max_out ..] = indices of frame f's key points with |dx| < r,
 * |dy| < r and octave in range, in the order the reference visits them (cells ix-major, iy, then the cell's list);
 * d_counts[f*nq + q] = how many there are (may exceed max_out).  The window search a projection / grid descriptor search starts
 * from. */
int sdpl_post_features_in_area_dev(sdpl_post* h, int nframes, int w, int h_, const sdpl_keypoint* d_kps, int capacity,
                                   const int32_t* d_cell_start, const int32_t* d_items, int grid_cols, int grid_rows, const float* d_queries,
                                   int nq, int32_t* d_out, int max_out, int* d_counts, int sync);

/* The descriptor search a projection match runs on that window (the reference keeps the window search and the
 * Hamming metric, the ORB-SLAM2 loop around them is not in its tree): d_qdesc = [nframes][nq][32] query descriptors; d_out5 =
 * [nframes][nq][5] int32 {best key point index, best distance, its octave, second-best distance, its octave} over the window's key
 * points, ties to the first visited; index -1 / distance 256 when the window is empty. */
int sdpl_post_search_area_dev(sdpl_post* h, int nframes, int w, int h_, const sdpl_keypoint* d_kps, const uint8_t* d_desc, int capacity,
                              const int32_t* d_cell_start, const int32_t* d_items, int grid_cols, int grid_rows, const float* d_queries,
                              const uint8_t* d_qdesc, int nq, int32_t* d_out5, int sync);
/* MapPoint::ComputeDistinctiveDescriptors, src/MapPoint.cc:242-307, for n_points map points: the observed descriptors of point p are
 * rows d_start[p] .. d_start[p+1] of d_desc (at most 64 per point); d_best_idx[p] = the observation with the least median distance to
 * the others (first on ties, -1 without observations), d_out_desc[p] = its 32 bytes */
int sdpl_post_distinctive_descriptors_dev(sdpl_post* h, const uint8_t* d_desc, const int32_t* d_start, int n_points, int32_t* d_best_idx,
                                          uint8_t* d_out_desc, int sync);
/* MapPoint::PredictScale, src/MapPoint.cc:385-417: ceil(log(max_distance / current_dist) / log_scale_factor) clamped to [0, n_levels) */
int sdpl_post_predict_scale_dev(sdpl_post* h, const float* d_max_distance, const float* d_current_dist, int n, float log_scale_factor,
                                int n_levels, int32_t* d_out, int sync);

/* Order-independent 64-bit digest of per-frame result rows resident on the device: adds, for every frame f, the sum of the
 * hashes of rows [0, min(d_n[f], max_rows)) of its block (row_bytes per row, a multiple of 4; blocks frame_stride bytes apart)
 * to d_digest[f] (DEVICE uint64 array the caller zeroes), asynchronously on `stream` (cudaStream_t as void*).  Used to check
 * that a batch sharded over several GPUs gives byte-identical per-frame results. */
int sdpl_rows_digest_dev(const void* d_rows, int row_bytes, size_t frame_stride, const int* d_n, int nframes, int max_rows,
                         unsigned long long salt, unsigned long long* d_digest, void* stream);

/* Bind a handle's work to a caller stream (cudaStream_t as void*; NULL = the handle's own stream). */
int sdpl_orb_set_stream(sdpl_orb* h, void* cuda_stream);
int sdpl_line_set_stream(sdpl_line* h, void* cuda_stream);
int sdpl_matcher_set_stream(sdpl_matcher* h, void* cuda_stream);

#ifdef __cplusplus
}
#endif
#endif
