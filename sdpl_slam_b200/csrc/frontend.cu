// frontend.cu -- the per-frame front-end as ONE call: what Frame::Frame does with its extractors
// (ExtractORB + ExtractLines, src/Frame.cc:314,328 -> :927-949) followed by frame-to-frame descriptor association of
// points and lines (the descriptor search of BASELINE.json's north_star; SURVEY.md F2: the reference itself associates by
// optical flow, so the Hamming search is the intended replacement surface).
//
// sdpl_frontend_process takes a batch of consecutive HOST frames, uploads them once, runs the ORB pipeline and the line
// pipeline concurrently on two streams, matches every frame's ORB / LBD descriptors against the previous frame's on two
// more streams (frame 0 against the last frame of the previous call), and brings everything back to the host.  It is
// built only from the public device-pointer entry points of this library (include/sdpl_frontend.h).
#include "common.cuh"
#include <string.h>
#include <algorithm>

using namespace sdpl;

struct sdpl_frontend {
  int device = 0;
  sdpl_orb* orb = nullptr;
  sdpl_line* line = nullptr;
  sdpl_matcher* pm = nullptr;
  sdpl_matcher* lm = nullptr;
  float ratio = 0.8f; int max_dist = 64;
  int kp_cap = 0, kl_cap = 0;
  cudaStream_t s_io = nullptr, s_orb = nullptr, s_line = nullptr, s_pm = nullptr, s_lm = nullptr;
  cudaEvent_t ev_in = nullptr, ev_orb = nullptr, ev_line = nullptr, ev_pm = nullptr, ev_lm = nullptr;
  // device buffers hold nframes+1 descriptor blocks: slot 0 = last frame of the previous call
  DevBuf d_imgs, d_kps, d_desc, d_nkp, d_kls, d_ldesc, d_nkl, d_pbest, d_psecond, d_pout, d_pacc, d_lbest, d_lsecond, d_lout, d_lacc;
  void* h_stage = nullptr; size_t h_stage_bytes = 0;
  int have_prev = 0;
  int launches = 0;
  int cap_frames = 0;
};

static int fe_reserve(sdpl_frontend* f, int n, int w, int h) {
  int rc;
  const size_t N1 = (size_t)n + 1;
  if ((rc = f->d_imgs.reserve((size_t)w * h * n))) return rc;
  if ((rc = f->d_kps.reserve(sizeof(sdpl_keypoint) * f->kp_cap * N1))) return rc;
  if ((rc = f->d_desc.reserve((size_t)32 * f->kp_cap * N1))) return rc;
  if ((rc = f->d_nkp.reserve(sizeof(int) * N1))) return rc;
  if ((rc = f->d_kls.reserve(sizeof(sdpl_keyline) * f->kl_cap * N1))) return rc;
  if ((rc = f->d_ldesc.reserve((size_t)32 * f->kl_cap * N1))) return rc;
  if ((rc = f->d_nkl.reserve(sizeof(int) * N1))) return rc;
  if ((rc = f->d_pbest.reserve(sizeof(sdpl_dmatch) * f->kp_cap * N1))) return rc;
  if ((rc = f->d_psecond.reserve(sizeof(sdpl_dmatch) * f->kp_cap * N1))) return rc;
  if ((rc = f->d_pout.reserve(sizeof(sdpl_dmatch) * f->kp_cap * N1))) return rc;
  if ((rc = f->d_pacc.reserve(sizeof(int) * N1))) return rc;
  if ((rc = f->d_lbest.reserve(sizeof(sdpl_dmatch) * f->kl_cap * N1))) return rc;
  if ((rc = f->d_lsecond.reserve(sizeof(sdpl_dmatch) * f->kl_cap * N1))) return rc;
  if ((rc = f->d_lout.reserve(sizeof(sdpl_dmatch) * f->kl_cap * N1))) return rc;
  if ((rc = f->d_lacc.reserve(sizeof(int) * N1))) return rc;
  return SDPL_OK;
}

extern "C" {

int sdpl_frontend_create(sdpl_frontend** out, int nfeatures, float scale, int nlevels, int ini_th, int min_th, int lsd_nfeatures,
                         int lsd_refine, float lsd_scale, int lsd_levels, float lsd_pyr_scale, float ratio, int max_dist, int device) {
  if (!out || !(ratio > 0.f) || max_dist < 0) { set_last_error("sdpl_frontend_create: bad argument"); return SDPL_ERR_ARG; }
  sdpl_frontend* f = new sdpl_frontend;
  f->device = device; f->ratio = ratio; f->max_dist = max_dist;
  int rc;
  if ((rc = sdpl_orb_create(&f->orb, nfeatures, scale, nlevels, ini_th, min_th, device)) ||
      (rc = sdpl_line_create(&f->line, lsd_nfeatures, lsd_refine, lsd_scale, lsd_levels, lsd_pyr_scale, 0, device)) ||
      (rc = sdpl_matcher_create(&f->pm, device)) || (rc = sdpl_matcher_create(&f->lm, device))) {
    sdpl_frontend_destroy(f);
    return rc;
  }
  f->kp_cap = sdpl_orb_max_keypoints(f->orb);
  f->kl_cap = lsd_nfeatures > 0 ? lsd_nfeatures : 2048;
  SDPL_CUDA(cudaSetDevice(device));
  {
    // the latency-bound line pipeline runs on the high-priority stream: its CTAs are placed first and the ORB / matching
    // kernels fill the SMs it leaves idle
    int lo = 0, hi = 0;
    SDPL_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    for (cudaStream_t* s : {&f->s_io, &f->s_orb, &f->s_pm, &f->s_lm}) SDPL_CUDA(cudaStreamCreateWithPriority(s, cudaStreamNonBlocking, lo));
    SDPL_CUDA(cudaStreamCreateWithPriority(&f->s_line, cudaStreamNonBlocking, hi));
  }
  for (cudaEvent_t* e : {&f->ev_in, &f->ev_orb, &f->ev_line, &f->ev_pm, &f->ev_lm}) SDPL_CUDA(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
  sdpl_orb_set_stream(f->orb, f->s_orb); sdpl_line_set_stream(f->line, f->s_line);
  sdpl_matcher_set_stream(f->pm, f->s_pm); sdpl_matcher_set_stream(f->lm, f->s_lm);
  *out = f;
  return SDPL_OK;
}

void sdpl_frontend_destroy(sdpl_frontend* f) {
  if (!f) return;
  cudaSetDevice(f->device);
  cudaDeviceSynchronize();
  sdpl_orb_destroy(f->orb); sdpl_line_destroy(f->line); sdpl_matcher_destroy(f->pm); sdpl_matcher_destroy(f->lm);
  for (DevBuf* b : {&f->d_imgs, &f->d_kps, &f->d_desc, &f->d_nkp, &f->d_kls, &f->d_ldesc, &f->d_nkl, &f->d_pbest, &f->d_psecond, &f->d_pout,
                    &f->d_pacc, &f->d_lbest, &f->d_lsecond, &f->d_lout, &f->d_lacc})
    b->release();
  if (f->h_stage) cudaFreeHost(f->h_stage);
  for (cudaStream_t s : {f->s_io, f->s_orb, f->s_line, f->s_pm, f->s_lm}) if (s) cudaStreamDestroy(s);
  for (cudaEvent_t e : {f->ev_in, f->ev_orb, f->ev_line, f->ev_pm, f->ev_lm}) if (e) cudaEventDestroy(e);
  delete f;
}

int sdpl_frontend_capacities(const sdpl_frontend* f, int* kp_capacity, int* kl_capacity) {
  if (!f) return SDPL_ERR_ARG;
  if (kp_capacity) *kp_capacity = f->kp_cap;
  if (kl_capacity) *kl_capacity = f->kl_cap;
  return SDPL_OK;
}
int sdpl_frontend_last_launches(const sdpl_frontend* f) { return f ? f->launches : 0; }
int sdpl_frontend_reset(sdpl_frontend* f) { if (!f) return SDPL_ERR_ARG; f->have_prev = 0; return SDPL_OK; }

int sdpl_frontend_process(sdpl_frontend* f, const uint8_t* imgs, int n, int w, int h, int stride, size_t frame_stride,
                          sdpl_keypoint* kps, uint8_t* desc, sdpl_keyline* kls, uint8_t* ldesc, sdpl_dmatch* pt_matches,
                          sdpl_dmatch* ln_matches, sdpl_frame_stats* stats) {
  if (!f || !imgs || n < 1 || w < 1 || h < 1 || stride < w || !kps || !desc || !kls || !ldesc || !pt_matches || !ln_matches || !stats) {
    set_last_error("sdpl_frontend_process: bad argument");
    return SDPL_ERR_ARG;
  }
  SDPL_CUDA(cudaSetDevice(f->device));
  int rc;
  if ((rc = fe_reserve(f, std::max(n, f->cap_frames), w, h))) return rc;
  f->cap_frames = std::max(n, f->cap_frames);
  const int KC = f->kp_cap, LC = f->kl_cap;
  // slot s = frame s-1 of this call; slot 0 = previous call's last frame
  sdpl_keypoint* dk = f->d_kps.as<sdpl_keypoint>() + KC;
  uint8_t* dd = f->d_desc.as<uint8_t>() + (size_t)32 * KC;
  int* dn = f->d_nkp.as<int>() + 1;
  sdpl_keyline* dl = f->d_kls.as<sdpl_keyline>() + LC;
  uint8_t* dld = f->d_ldesc.as<uint8_t>() + (size_t)32 * LC;
  int* dln = f->d_nkl.as<int>() + 1;
  if (!f->have_prev) {
    SDPL_CUDA(cudaMemsetAsync(f->d_nkp.p, 0, sizeof(int), f->s_io));
    SDPL_CUDA(cudaMemsetAsync(f->d_nkl.p, 0, sizeof(int), f->s_io));
  }
  // ---- one upload shared by both pipelines ----
  if (stride == w && frame_stride == (size_t)w * h) {
    SDPL_CUDA(cudaMemcpyAsync(f->d_imgs.p, imgs, (size_t)w * h * n, cudaMemcpyHostToDevice, f->s_io));
  } else {
    for (int i = 0; i < n; i++)
      SDPL_CUDA(cudaMemcpy2DAsync((char*)f->d_imgs.p + (size_t)i * w * h, w, imgs + (size_t)i * frame_stride, stride, w, h,
                                  cudaMemcpyHostToDevice, f->s_io));
  }
  SDPL_CUDA(cudaEventRecord(f->ev_in, f->s_io));
  int launches = 0;
  // ---- lines (high priority, launched first) and ORB concurrently ----
  SDPL_CUDA(cudaStreamWaitEvent(f->s_line, f->ev_in, 0));
  if ((rc = sdpl_line_extract_batch_dev(f->line, f->d_imgs.as<uint8_t>(), n, w, h, w, (size_t)w * h, dl, dld, LC, dln, 0))) return rc;
  launches += sdpl_line_last_launches(f->line);
  SDPL_CUDA(cudaEventRecord(f->ev_line, f->s_line));
  SDPL_CUDA(cudaStreamWaitEvent(f->s_orb, f->ev_in, 0));
  if ((rc = sdpl_orb_extract_batch_dev(f->orb, f->d_imgs.as<uint8_t>(), n, w, h, w, (size_t)w * h, dk, dd, KC, dn, 0))) return rc;
  launches += sdpl_orb_last_launches(f->orb);
  SDPL_CUDA(cudaEventRecord(f->ev_orb, f->s_orb));
  // ---- frame t against frame t-1 (slot t+1 against slot t), points then lines ----
  SDPL_CUDA(cudaStreamWaitEvent(f->s_pm, f->ev_orb, 0));
  if ((rc = sdpl_match_knn2_batch_dev(f->pm, dd, dn, (size_t)32 * KC, f->d_desc.as<uint8_t>(), f->d_nkp.as<int>(), (size_t)32 * KC, n, KC, KC,
                                      f->d_pbest.as<sdpl_dmatch>(), f->d_psecond.as<sdpl_dmatch>(), 0))) return rc;
  launches += sdpl_matcher_last_launches(f->pm);
  if ((rc = sdpl_match_ratio_batch_dev(f->pm, f->d_pbest.as<sdpl_dmatch>(), f->d_psecond.as<sdpl_dmatch>(), dn, n, KC, f->ratio, f->max_dist,
                                       f->d_pout.as<sdpl_dmatch>(), f->d_pacc.as<int>(), 0))) return rc;
  launches += sdpl_matcher_last_launches(f->pm);
  SDPL_CUDA(cudaStreamWaitEvent(f->s_lm, f->ev_line, 0));
  if ((rc = sdpl_match_knn2_batch_dev(f->lm, dld, dln, (size_t)32 * LC, f->d_ldesc.as<uint8_t>(), f->d_nkl.as<int>(), (size_t)32 * LC, n, LC, LC,
                                      f->d_lbest.as<sdpl_dmatch>(), f->d_lsecond.as<sdpl_dmatch>(), 0))) return rc;
  launches += sdpl_matcher_last_launches(f->lm);
  if ((rc = sdpl_match_ratio_batch_dev(f->lm, f->d_lbest.as<sdpl_dmatch>(), f->d_lsecond.as<sdpl_dmatch>(), dln, n, LC, f->ratio, f->max_dist,
                                       f->d_lout.as<sdpl_dmatch>(), f->d_lacc.as<int>(), 0))) return rc;
  launches += sdpl_matcher_last_launches(f->lm);
  f->launches = launches;
  // ---- results back: counts first (they size the row copies), then the valid rows of every frame ----
  const size_t need = sizeof(int) * 4 * (size_t)n;
  if (f->h_stage_bytes < need) {
    if (f->h_stage) cudaFreeHost(f->h_stage);
    f->h_stage = nullptr; f->h_stage_bytes = 0;
    SDPL_CUDA(cudaMallocHost(&f->h_stage, need));
    f->h_stage_bytes = need;
  }
  int* hn = (int*)f->h_stage;                    // [4][n]: n_kp, n_lines, point matches, line matches
  SDPL_CUDA(cudaStreamWaitEvent(f->s_orb, f->ev_orb, 0));
  SDPL_CUDA(cudaMemcpyAsync(hn, dn, sizeof(int) * n, cudaMemcpyDeviceToHost, f->s_orb));
  SDPL_CUDA(cudaStreamSynchronize(f->s_orb));
  int status = SDPL_OK;
  // ORB rows can go while the line pipeline is still running
  for (int i = 0; i < n; i++) {
    int c = hn[i];
    if (c > KC) { c = KC; status = SDPL_ERR_CAPACITY; }
    if (c > 0) {
      SDPL_CUDA(cudaMemcpyAsync(kps + (size_t)i * KC, dk + (size_t)i * KC, sizeof(sdpl_keypoint) * c, cudaMemcpyDeviceToHost, f->s_orb));
      SDPL_CUDA(cudaMemcpyAsync(desc + (size_t)i * KC * 32, dd + (size_t)i * KC * 32, (size_t)32 * c, cudaMemcpyDeviceToHost, f->s_orb));
    }
  }
  SDPL_CUDA(cudaMemcpyAsync(hn + 2 * n, f->d_pacc.p, sizeof(int) * n, cudaMemcpyDeviceToHost, f->s_pm));
  SDPL_CUDA(cudaStreamSynchronize(f->s_pm));
  for (int i = 0; i < n; i++) {
    const int c = std::min(hn[i], KC);
    if (c > 0) SDPL_CUDA(cudaMemcpyAsync(pt_matches + (size_t)i * KC, f->d_pout.as<sdpl_dmatch>() + (size_t)i * KC, sizeof(sdpl_dmatch) * c,
                                         cudaMemcpyDeviceToHost, f->s_pm));
  }
  SDPL_CUDA(cudaMemcpyAsync(hn + n, dln, sizeof(int) * n, cudaMemcpyDeviceToHost, f->s_line));
  SDPL_CUDA(cudaStreamSynchronize(f->s_line));
  for (int i = 0; i < n; i++) {
    int c = hn[n + i];
    if (c > LC) { c = LC; status = SDPL_ERR_CAPACITY; }
    if (c > 0) {
      SDPL_CUDA(cudaMemcpyAsync(kls + (size_t)i * LC, dl + (size_t)i * LC, sizeof(sdpl_keyline) * c, cudaMemcpyDeviceToHost, f->s_line));
      SDPL_CUDA(cudaMemcpyAsync(ldesc + (size_t)i * LC * 32, dld + (size_t)i * LC * 32, (size_t)32 * c, cudaMemcpyDeviceToHost, f->s_line));
    }
  }
  SDPL_CUDA(cudaMemcpyAsync(hn + 3 * n, f->d_lacc.p, sizeof(int) * n, cudaMemcpyDeviceToHost, f->s_lm));
  SDPL_CUDA(cudaStreamSynchronize(f->s_lm));
  for (int i = 0; i < n; i++) {
    const int c = std::min(hn[n + i], LC);
    if (c > 0) SDPL_CUDA(cudaMemcpyAsync(ln_matches + (size_t)i * LC, f->d_lout.as<sdpl_dmatch>() + (size_t)i * LC, sizeof(sdpl_dmatch) * c,
                                         cudaMemcpyDeviceToHost, f->s_lm));
  }
  // keep the last frame's descriptors as the "previous frame" of the next call (slot n -> slot 0), after the matchers read slot 0
  SDPL_CUDA(cudaEventRecord(f->ev_pm, f->s_pm));
  SDPL_CUDA(cudaEventRecord(f->ev_lm, f->s_lm));
  SDPL_CUDA(cudaStreamWaitEvent(f->s_io, f->ev_pm, 0));
  SDPL_CUDA(cudaStreamWaitEvent(f->s_io, f->ev_lm, 0));
  SDPL_CUDA(cudaMemcpyAsync(f->d_desc.p, dd + (size_t)(n - 1) * KC * 32, (size_t)32 * KC, cudaMemcpyDeviceToDevice, f->s_io));
  SDPL_CUDA(cudaMemcpyAsync(f->d_nkp.p, dn + (n - 1), sizeof(int), cudaMemcpyDeviceToDevice, f->s_io));
  SDPL_CUDA(cudaMemcpyAsync(f->d_ldesc.p, dld + (size_t)(n - 1) * LC * 32, (size_t)32 * LC, cudaMemcpyDeviceToDevice, f->s_io));
  SDPL_CUDA(cudaMemcpyAsync(f->d_nkl.p, dln + (n - 1), sizeof(int), cudaMemcpyDeviceToDevice, f->s_io));
  for (cudaStream_t s : {f->s_orb, f->s_line, f->s_pm, f->s_lm, f->s_io}) SDPL_CUDA(cudaStreamSynchronize(s));
  f->have_prev = 1;
  if ((rc = sdpl_orb_check(f->orb)) || (rc = sdpl_line_check(f->line))) return rc;
  for (int i = 0; i < n; i++) {
    stats[i].n_kp = hn[i]; stats[i].n_lines = hn[n + i]; stats[i].n_pt_matches = hn[2 * n + i]; stats[i].n_ln_matches = hn[3 * n + i];
  }
  if (status == SDPL_ERR_CAPACITY) set_last_error("sdpl_frontend_process: more features than the per-frame capacity");
  return status;
}

}  // extern "C"
