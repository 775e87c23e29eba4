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

struct FeSlot {
  // device buffers hold nframes+1 descriptor blocks: block 0 = last frame of the previous batch
  DevBuf imgs, kps, desc, nkp, kls, ldesc, nkl, pbest, psecond, pout, pacc, lbest, lsecond, lout, lacc;
  cudaEvent_t ev_in = nullptr, ev_orb = nullptr, ev_line = nullptr, ev_pm = nullptr, ev_lm = nullptr;
  cudaEvent_t ev_pread = nullptr, ev_lread = nullptr;   // the NEXT batch has copied this slot's last descriptor block
  bool read_pending = false;
  int* hn = nullptr; size_t hn_bytes = 0;     // pinned: [4][n] counts + 2 error flags
  int n = 0, w = 0, h = 0;
  size_t herr_off = 0;                         // where in hn this batch's two overflow flags were written
  bool busy = false;
};

struct sdpl_frontend {
  int device = 0;
  sdpl_orb* orb = nullptr;
  sdpl_line* line = nullptr;
  sdpl_matcher* pm = nullptr;
  sdpl_matcher* lm = nullptr;
  float ratio = 0.8f; int max_dist = 64;
  int kp_cap = 0, kl_cap = 0;
  cudaStream_t s_io = nullptr, s_orb = nullptr, s_line = nullptr, s_pm = nullptr, s_lm = nullptr, s_out = nullptr;
  FeSlot slot[2];
  int head = 0, count = 0, next = 0;     // FIFO of submitted batches
  int last = -1;                         // slot of the most recently submitted batch (source of the "previous frame")
  int have_prev = 0;
  int launches = 0;
};

static int fe_create_streams(sdpl_frontend* f);

static int fe_reserve(sdpl_frontend* f, FeSlot& S, int n, int w, int h) {
  int rc;
  const size_t N1 = (size_t)n + 1;
  if ((rc = S.imgs.reserve((size_t)w * h * n))) return rc;
  if ((rc = S.kps.reserve(sizeof(sdpl_keypoint) * f->kp_cap * N1))) return rc;
  if ((rc = S.desc.reserve((size_t)32 * f->kp_cap * N1))) return rc;
  if ((rc = S.nkp.reserve(sizeof(int) * N1))) return rc;
  if ((rc = S.kls.reserve(sizeof(sdpl_keyline) * f->kl_cap * N1))) return rc;
  if ((rc = S.ldesc.reserve((size_t)32 * f->kl_cap * N1))) return rc;
  if ((rc = S.nkl.reserve(sizeof(int) * N1))) return rc;
  if ((rc = S.pbest.reserve(sizeof(sdpl_dmatch) * f->kp_cap * N1))) return rc;
  if ((rc = S.psecond.reserve(sizeof(sdpl_dmatch) * f->kp_cap * N1))) return rc;
  if ((rc = S.pout.reserve(sizeof(sdpl_dmatch) * f->kp_cap * N1))) return rc;
  if ((rc = S.pacc.reserve(sizeof(int) * N1))) return rc;
  if ((rc = S.lbest.reserve(sizeof(sdpl_dmatch) * f->kl_cap * N1))) return rc;
  if ((rc = S.lsecond.reserve(sizeof(sdpl_dmatch) * f->kl_cap * N1))) return rc;
  if ((rc = S.lout.reserve(sizeof(sdpl_dmatch) * f->kl_cap * N1))) return rc;
  if ((rc = S.lacc.reserve(sizeof(int) * N1))) return rc;
  const size_t need = sizeof(int) * (4 * (size_t)n + 2);
  if (S.hn_bytes < need) {
    if (S.hn) cudaFreeHost(S.hn);
    S.hn = nullptr; S.hn_bytes = 0;
    SDPL_CUDA(cudaMallocHost((void**)&S.hn, need));
    S.hn_bytes = need;
  }
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
  *out = f;                        // from here on an error path hands the half-built object to the caller's destroy ...
  int rc2 = fe_create_streams(f);
  if (rc2) { sdpl_frontend_destroy(f); *out = nullptr; return rc2; }   // ... or frees it here
  return SDPL_OK;
}

}  // extern "C"

static int fe_create_streams(sdpl_frontend* f) {
  const int device = f->device;
  SDPL_CUDA(cudaSetDevice(device));
  {
    // all streams share one priority by default: a high-priority line stream (SDPL_FE_LINE_PRIO=high) measured 1-5 % slower on
    // B200 -- the region-growing kernel holds every register of an SM while it is resident, so the pipelines take turns anyway
    int lo = 0, hi = 0;
    SDPL_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    for (cudaStream_t* s : {&f->s_io, &f->s_orb, &f->s_pm, &f->s_lm, &f->s_out}) SDPL_CUDA(cudaStreamCreateWithPriority(s, cudaStreamNonBlocking, lo));
    const char* lp = getenv("SDPL_FE_LINE_PRIO");
    SDPL_CUDA(cudaStreamCreateWithPriority(&f->s_line, cudaStreamNonBlocking, (lp && !strcmp(lp, "high")) ? hi : lo));
  }
  for (FeSlot& S : f->slot)
    for (cudaEvent_t* e : {&S.ev_in, &S.ev_orb, &S.ev_line, &S.ev_pm, &S.ev_lm, &S.ev_pread, &S.ev_lread})
      SDPL_CUDA(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
  sdpl_orb_set_stream(f->orb, f->s_orb); sdpl_line_set_stream(f->line, f->s_line);
  sdpl_matcher_set_stream(f->pm, f->s_pm); sdpl_matcher_set_stream(f->lm, f->s_lm);
  return SDPL_OK;
}

extern "C" {

void sdpl_frontend_destroy(sdpl_frontend* f) {
  if (!f) return;
  cudaSetDevice(f->device);
  cudaDeviceSynchronize();
  sdpl_orb_destroy(f->orb); sdpl_line_destroy(f->line); sdpl_matcher_destroy(f->pm); sdpl_matcher_destroy(f->lm);
  for (FeSlot& S : f->slot) {
    for (DevBuf* b : {&S.imgs, &S.kps, &S.desc, &S.nkp, &S.kls, &S.ldesc, &S.nkl, &S.pbest, &S.psecond, &S.pout, &S.pacc, &S.lbest, &S.lsecond,
                      &S.lout, &S.lacc})
      b->release();
    if (S.hn) cudaFreeHost(S.hn);
    for (cudaEvent_t e : {S.ev_in, S.ev_orb, S.ev_line, S.ev_pm, S.ev_lm, S.ev_pread, S.ev_lread}) if (e) cudaEventDestroy(e);
  }
  for (cudaStream_t s : {f->s_io, f->s_orb, f->s_line, f->s_pm, f->s_lm, f->s_out}) if (s) cudaStreamDestroy(s);
  delete f;
}

int sdpl_frontend_capacities(const sdpl_frontend* f, int* kp_capacity, int* kl_capacity) {
  if (!f) return SDPL_ERR_ARG;
  if (kp_capacity) *kp_capacity = f->kp_cap;
  if (kl_capacity) *kl_capacity = f->kl_cap;
  return SDPL_OK;
}
int sdpl_frontend_set_line_capacity(sdpl_frontend* f, int kl_capacity) {
  if (!f || kl_capacity < 1 || f->count > 0) { set_last_error("sdpl_frontend_set_line_capacity: bad argument or batches in flight"); return SDPL_ERR_ARG; }
  cudaSetDevice(f->device);
  cudaDeviceSynchronize();
  if (kl_capacity != f->kl_cap) {
    // the per-slot blocks are laid out by capacity: drop the carried-over "previous frame" with the old layout
    f->kl_cap = kl_capacity; f->have_prev = 0;
    for (FeSlot& S : f->slot) S.n = 0;
  }
  return SDPL_OK;
}
int sdpl_frontend_set_line_extractor(sdpl_frontend* f, int extractor) {
  if (!f || f->count > 0) { set_last_error("sdpl_frontend_set_line_extractor: bad argument or batches in flight"); return SDPL_ERR_ARG; }
  const int rc = sdpl_line_set_extractor(f->line, extractor);
  if (rc == SDPL_OK) f->have_prev = 0;
  return rc;
}
int sdpl_frontend_last_launches(const sdpl_frontend* f) { return f ? f->launches : 0; }
int sdpl_frontend_pending(const sdpl_frontend* f) { return f ? f->count : 0; }
int sdpl_frontend_reset(sdpl_frontend* f) { if (!f) return SDPL_ERR_ARG; f->have_prev = 0; return SDPL_OK; }

// Enqueue one batch: upload, ORB || lines, matching against the previous frame.  Returns without waiting; `imgs` must stay
// valid until the batch has been collected.  At most two batches can be in flight.
int sdpl_frontend_submit(sdpl_frontend* f, const uint8_t* imgs, int n, int w, int h, int stride, size_t frame_stride) {
  if (!f || !imgs || n < 1 || w < 1 || h < 1 || stride < w) { set_last_error("sdpl_frontend_submit: bad argument"); return SDPL_ERR_ARG; }
  if (f->count >= 2) { set_last_error("sdpl_frontend_submit: two batches already in flight, collect one first"); return SDPL_ERR_ARG; }
  SDPL_CUDA(cudaSetDevice(f->device));
  FeSlot& S = f->slot[f->next];
  int rc;
  if ((rc = fe_reserve(f, S, std::max(n, S.n), w, h))) return rc;
  const int KC = f->kp_cap, LC = f->kl_cap;
  sdpl_keypoint* dk = S.kps.as<sdpl_keypoint>() + KC;
  uint8_t* dd = S.desc.as<uint8_t>() + (size_t)32 * KC;
  int* dn = S.nkp.as<int>() + 1;
  sdpl_keyline* dl = S.kls.as<sdpl_keyline>() + LC;
  uint8_t* dld = S.ldesc.as<uint8_t>() + (size_t)32 * LC;
  int* dln = S.nkl.as<int>() + 1;
  // ---- one upload shared by both pipelines ----
  if (stride == w && frame_stride == (size_t)w * h) {
    SDPL_CUDA(cudaMemcpyAsync(S.imgs.p, imgs, (size_t)w * h * n, cudaMemcpyHostToDevice, f->s_io));
  } else {
    for (int i = 0; i < n; i++)
      SDPL_CUDA(cudaMemcpy2DAsync((char*)S.imgs.p + (size_t)i * w * h, w, imgs + (size_t)i * frame_stride, stride, w, h,
                                  cudaMemcpyHostToDevice, f->s_io));
  }
  SDPL_CUDA(cudaEventRecord(S.ev_in, f->s_io));
  int launches = 0;
  // ---- block 0 = last frame of the previous batch (or nothing).  The copies ride on the matcher streams, behind the
  //      previous batch's extraction: the upload above and the extraction below do not wait for the previous batch ----
  if (f->have_prev && f->last >= 0) {
    FeSlot& P = f->slot[f->last];
    const int pl = P.n;   // its last frame sits in block pl
    SDPL_CUDA(cudaStreamWaitEvent(f->s_pm, P.ev_orb, 0));
    SDPL_CUDA(cudaStreamWaitEvent(f->s_lm, P.ev_line, 0));
    SDPL_CUDA(cudaMemcpyAsync(S.desc.p, P.desc.as<uint8_t>() + (size_t)pl * KC * 32, (size_t)32 * KC, cudaMemcpyDeviceToDevice, f->s_pm));
    SDPL_CUDA(cudaMemcpyAsync(S.nkp.p, P.nkp.as<int>() + pl, sizeof(int), cudaMemcpyDeviceToDevice, f->s_pm));
    SDPL_CUDA(cudaMemcpyAsync(S.ldesc.p, P.ldesc.as<uint8_t>() + (size_t)pl * LC * 32, (size_t)32 * LC, cudaMemcpyDeviceToDevice, f->s_lm));
    SDPL_CUDA(cudaMemcpyAsync(S.nkl.p, P.nkl.as<int>() + pl, sizeof(int), cudaMemcpyDeviceToDevice, f->s_lm));
    // the batch after this one extracts into P again: it has to wait for these reads
    SDPL_CUDA(cudaEventRecord(P.ev_pread, f->s_pm));
    SDPL_CUDA(cudaEventRecord(P.ev_lread, f->s_lm));
    P.read_pending = true;
    if (&P == &S) {
      // same buffers (cannot happen while the slots alternate): the extraction below overwrites the block just copied from
      SDPL_CUDA(cudaEventRecord(S.ev_pm, f->s_pm));
      SDPL_CUDA(cudaEventRecord(S.ev_lm, f->s_lm));
      SDPL_CUDA(cudaStreamWaitEvent(f->s_orb, S.ev_pm, 0));
      SDPL_CUDA(cudaStreamWaitEvent(f->s_line, S.ev_lm, 0));
    }
  } else {
    SDPL_CUDA(cudaMemsetAsync(S.nkp.p, 0, sizeof(int), f->s_pm));
    SDPL_CUDA(cudaMemsetAsync(S.nkl.p, 0, sizeof(int), f->s_lm));
  }
  if (S.read_pending) {
    // the previous submit copied this slot's last descriptor block on the matcher streams: do not overwrite it before that
    SDPL_CUDA(cudaStreamWaitEvent(f->s_orb, S.ev_pread, 0));
    SDPL_CUDA(cudaStreamWaitEvent(f->s_line, S.ev_lread, 0));
    S.read_pending = false;
  }
  // ---- lines (launched first) and ORB concurrently ----
  SDPL_CUDA(cudaStreamWaitEvent(f->s_line, S.ev_in, 0));
  if ((rc = sdpl_line_extract_batch_dev(f->line, S.imgs.as<uint8_t>(), n, w, h, w, (size_t)w * h, dl, dld, LC, dln, 0))) return rc;
  launches += sdpl_line_last_launches(f->line);
  if ((rc = sdpl_line_take_error_async(f->line, S.hn + 4 * (size_t)std::max(n, S.n) + 1))) return rc;   // this batch's own overflow flag
  SDPL_CUDA(cudaEventRecord(S.ev_line, f->s_line));
  SDPL_CUDA(cudaStreamWaitEvent(f->s_orb, S.ev_in, 0));
  if ((rc = sdpl_orb_extract_batch_dev(f->orb, S.imgs.as<uint8_t>(), n, w, h, w, (size_t)w * h, dk, dd, KC, dn, 0))) return rc;
  launches += sdpl_orb_last_launches(f->orb);
  if ((rc = sdpl_orb_take_error_async(f->orb, S.hn + 4 * (size_t)std::max(n, S.n)))) return rc;
  SDPL_CUDA(cudaEventRecord(S.ev_orb, f->s_orb));
  // ---- frame t against frame t-1 (block t+1 against block t), points then lines ----
  SDPL_CUDA(cudaStreamWaitEvent(f->s_pm, S.ev_orb, 0));
  if ((rc = sdpl_match_knn2_batch_dev(f->pm, dd, dn, (size_t)32 * KC, S.desc.as<uint8_t>(), S.nkp.as<int>(), (size_t)32 * KC, n, KC, KC,
                                      S.pbest.as<sdpl_dmatch>(), S.psecond.as<sdpl_dmatch>(), 0))) return rc;
  launches += sdpl_matcher_last_launches(f->pm);
  if ((rc = sdpl_match_ratio_batch_dev(f->pm, S.pbest.as<sdpl_dmatch>(), S.psecond.as<sdpl_dmatch>(), dn, n, KC, f->ratio, f->max_dist,
                                       S.pout.as<sdpl_dmatch>(), S.pacc.as<int>(), 0))) return rc;
  launches += sdpl_matcher_last_launches(f->pm);
  SDPL_CUDA(cudaEventRecord(S.ev_pm, f->s_pm));
  SDPL_CUDA(cudaStreamWaitEvent(f->s_lm, S.ev_line, 0));
  if ((rc = sdpl_match_knn2_batch_dev(f->lm, dld, dln, (size_t)32 * LC, S.ldesc.as<uint8_t>(), S.nkl.as<int>(), (size_t)32 * LC, n, LC, LC,
                                      S.lbest.as<sdpl_dmatch>(), S.lsecond.as<sdpl_dmatch>(), 0))) return rc;
  launches += sdpl_matcher_last_launches(f->lm);
  if ((rc = sdpl_match_ratio_batch_dev(f->lm, S.lbest.as<sdpl_dmatch>(), S.lsecond.as<sdpl_dmatch>(), dln, n, LC, f->ratio, f->max_dist,
                                       S.lout.as<sdpl_dmatch>(), S.lacc.as<int>(), 0))) return rc;
  launches += sdpl_matcher_last_launches(f->lm);
  SDPL_CUDA(cudaEventRecord(S.ev_lm, f->s_lm));
  f->launches = launches;
  S.herr_off = 4 * (size_t)std::max(n, S.n);
  S.n = n; S.w = w; S.h = h; S.busy = true;
  f->last = f->next; f->have_prev = 1;
  f->next ^= 1; f->count++;
  return SDPL_OK;
}

// Wait for the oldest submitted batch and bring its results to the host (on a separate stream, so the copies overlap the
// kernels of a batch submitted after it).
int sdpl_frontend_collect(sdpl_frontend* f, sdpl_keypoint* kps, uint8_t* desc, sdpl_keyline* kls, uint8_t* ldesc, sdpl_dmatch* pt_matches,
                          sdpl_dmatch* ln_matches, sdpl_frame_stats* stats, int* n_frames) {
  if (!f || !kps || !desc || !kls || !ldesc || !pt_matches || !ln_matches || !stats) { set_last_error("sdpl_frontend_collect: bad argument"); return SDPL_ERR_ARG; }
  if (f->count < 1) { set_last_error("sdpl_frontend_collect: nothing submitted"); return SDPL_ERR_ARG; }
  SDPL_CUDA(cudaSetDevice(f->device));
  FeSlot& S = f->slot[f->head];
  const int n = S.n, KC = f->kp_cap, LC = f->kl_cap;
  const sdpl_keypoint* dk = S.kps.as<sdpl_keypoint>() + KC;
  const uint8_t* dd = S.desc.as<uint8_t>() + (size_t)32 * KC;
  const int* dn = S.nkp.as<int>() + 1;
  const sdpl_keyline* dl = S.kls.as<sdpl_keyline>() + LC;
  const uint8_t* dld = S.ldesc.as<uint8_t>() + (size_t)32 * LC;
  const int* dln = S.nkl.as<int>() + 1;
  int* hn = S.hn;                       // [4][n]: n_kp, n_lines, point matches, line matches ; then 2 error flags
  int* herr = hn + S.herr_off;          // taken (copied and cleared) in stream order right after this batch's extraction
  cudaStream_t so = f->s_out;
  int status = SDPL_OK;
  // counts size the copies: fetch them as soon as their stage is done, then ONE strided copy per output array -- for every frame
  // its first max-count rows (cudaMemcpy2DAsync: pitch = the frame's capacity-sized block, width = the longest frame's rows), so a
  // batch costs nine copies whatever its size
  auto rows2d = [&](void* dst, const void* src, size_t row_bytes, int cap, int maxc) -> cudaError_t {
    if (maxc <= 0) return cudaSuccess;
    return cudaMemcpy2DAsync(dst, row_bytes * cap, src, row_bytes * cap, row_bytes * maxc, n, cudaMemcpyDeviceToHost, so);
  };
  SDPL_CUDA(cudaStreamWaitEvent(so, S.ev_orb, 0));
  SDPL_CUDA(cudaMemcpyAsync(hn, dn, sizeof(int) * n, cudaMemcpyDeviceToHost, so));
  SDPL_CUDA(cudaStreamSynchronize(so));
  int kmax = 0;
  for (int i = 0; i < n; i++) {
    if (hn[i] > KC) status = SDPL_ERR_CAPACITY;
    kmax = std::max(kmax, std::min(hn[i], KC));
  }
  SDPL_CUDA(rows2d(kps, dk, sizeof(sdpl_keypoint), KC, kmax));
  SDPL_CUDA(rows2d(desc, dd, 32, KC, kmax));
  SDPL_CUDA(cudaStreamWaitEvent(so, S.ev_pm, 0));
  SDPL_CUDA(cudaMemcpyAsync(hn + 2 * n, S.pacc.p, sizeof(int) * n, cudaMemcpyDeviceToHost, so));
  SDPL_CUDA(rows2d(pt_matches, S.pout.p, sizeof(sdpl_dmatch), KC, kmax));
  SDPL_CUDA(cudaStreamWaitEvent(so, S.ev_line, 0));
  SDPL_CUDA(cudaMemcpyAsync(hn + n, dln, sizeof(int) * n, cudaMemcpyDeviceToHost, so));
  SDPL_CUDA(cudaStreamSynchronize(so));
  int lmax = 0;
  for (int i = 0; i < n; i++) {
    if (hn[n + i] > LC) status = SDPL_ERR_CAPACITY;
    lmax = std::max(lmax, std::min(hn[n + i], LC));
  }
  SDPL_CUDA(rows2d(kls, dl, sizeof(sdpl_keyline), LC, lmax));
  SDPL_CUDA(rows2d(ldesc, dld, 32, LC, lmax));
  SDPL_CUDA(cudaStreamWaitEvent(so, S.ev_lm, 0));
  SDPL_CUDA(cudaMemcpyAsync(hn + 3 * n, S.lacc.p, sizeof(int) * n, cudaMemcpyDeviceToHost, so));
  SDPL_CUDA(rows2d(ln_matches, S.lout.p, sizeof(sdpl_dmatch), LC, lmax));
  SDPL_CUDA(cudaStreamSynchronize(so));
  S.busy = false;
  f->head ^= 1; f->count--;
  if (n_frames) *n_frames = n;
  for (int i = 0; i < n; i++) {
    stats[i].n_kp = hn[i]; stats[i].n_lines = hn[n + i]; stats[i].n_pt_matches = hn[2 * n + i]; stats[i].n_ln_matches = hn[3 * n + i];
  }
  if (herr[0] || herr[1]) {
    set_last_error("device buffer overflow in the front-end pipeline (FAST candidates / quadtree nodes / pending rectangles)");
    return SDPL_ERR_OVERFLOW;
  }
  if (status == SDPL_ERR_CAPACITY) set_last_error("sdpl_frontend_collect: more features than the per-frame capacity");
  return status;
}

int sdpl_frontend_process(sdpl_frontend* f, const uint8_t* imgs, int n, int w, int h, int stride, size_t frame_stride,
                          sdpl_keypoint* kps, uint8_t* desc, sdpl_keyline* kls, uint8_t* ldesc, sdpl_dmatch* pt_matches,
                          sdpl_dmatch* ln_matches, sdpl_frame_stats* stats) {
  if (f && f->count > 0) { set_last_error("sdpl_frontend_process: batches are in flight, collect them first"); return SDPL_ERR_ARG; }
  int rc = sdpl_frontend_submit(f, imgs, n, w, h, stride, frame_stride);
  if (rc) return rc;
  return sdpl_frontend_collect(f, kps, desc, kls, ldesc, pt_matches, ln_matches, stats, nullptr);
}

}  // extern "C"
