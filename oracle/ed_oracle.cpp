// oracle/ed_oracle.cpp -- CPU ORACLE (test infrastructure, NOT the product) of the EDLines back-end of Lineextractor (extractor == 1):
// EDLines::EDLines(Mat) of 3rdparty/line_descriptor/src/ED_Lib (ED.cpp:8-62 constructor, :275-362 ComputeGradient, :364-398
// ComputeAnchorPoints, :1000-1047 sortAnchorsByGradValue1; EDLines.cpp:8-70) on one pyramid level.  The image-parallel steps are restated
// here as plain loops; the sequential steps (anchor linking, line fitting, joining, validation) are include/sdpl_edlines_core.h, the
// same source the CUDA kernel compiles.  Parity status: PINNED -- tests/test_oracle_vs_ref.py compares the key lines, and per level the
// smoothed image, the edge map, the segments and the lines, with the reference's own ED_Lib sources compiled unmodified.
#include "oracle.h"
#include "oracle_internal.h"
#include "../include/sdpl_edlines_core.h"
#include <cstring>
#include <map>
#include <utility>
#include <vector>

namespace orc {

// returns the number of lines (x1, y1, x2, y2 appended to seg), < 0 on a capacity error of the sequential part
int ed_detect(const uint8_t* roi, int w, int h, int stride, std::vector<float>& seg, EdDebug* dbg) {
  using namespace sdpl_ed;
  const int npx = w * h;
  // ED.cpp:38-41: GaussianBlur(srcImage, smoothImage, Size(5, 5), 1.0).  The Mat is a ROI of the padded buffer and the blur is not
  // isolated, so OpenCV reads the buffer's border pixels -- which are the reflect-101 extension an isolated blur synthesises.
  std::vector<uint8_t> smooth(npx), dir(npx, 0), edge(npx, 0);
  gaussian_blur_u8(roi, w, h, stride, smooth.data(), w, 1);
  // ComputeGradient with SOBEL_OPERATOR, sumFlag = true
  std::vector<int16_t> grad(npx);
  for (int j = 0; j < w; j++) grad[j] = grad[(h - 1) * w + j] = kGradThresh - 1;
  for (int i = 1; i < h - 1; i++) grad[i * w] = grad[(i + 1) * w - 1] = kGradThresh - 1;
  const uint8_t* s = smooth.data();
  for (int i = 1; i < h - 1; i++)
    for (int j = 1; j < w - 1; j++) {
      const int com1 = s[(i + 1) * w + j + 1] - s[(i - 1) * w + j - 1];
      const int com2 = s[(i - 1) * w + j + 1] - s[(i + 1) * w + j - 1];
      const int gx = std::abs(com1 + com2 + 2 * (s[i * w + j + 1] - s[i * w + j - 1]));
      const int gy = std::abs(com1 - com2 + 2 * (s[(i + 1) * w + j] - s[(i - 1) * w + j]));
      const int sum = gx + gy;
      grad[i * w + j] = (int16_t)sum;
      if (sum >= kGradThresh) dir[i * w + j] = gx >= gy ? kVertical : kHorizontal;
    }
  // ComputeAnchorPoints, scanInterval = 1
  for (int i = 2; i < h - 2; i++)
    for (int j = 2; j < w - 2; j++) {
      const int g = grad[i * w + j];
      if (g < kGradThresh) continue;
      int d1, d2;
      if (dir[i * w + j] == kVertical) { d1 = g - grad[i * w + j - 1]; d2 = g - grad[i * w + j + 1]; }
      else { d1 = g - grad[(i - 1) * w + j]; d2 = g - grad[(i + 1) * w + j]; }
      if (d1 >= kAnchorThresh && d2 >= kAnchorThresh) edge[i * w + j] = kAnchor;
    }
  // sortAnchorsByGradValue1 + the descending loop of JoinAnchorPointsUsingSortedAnchors: highest gradient first, row-major inside a value
  std::vector<int> count(128 * 256 + 1, 0);
  for (int q = 0; q < npx; q++) if (edge[q] == kAnchor) count[grad[q]]++;
  std::vector<int> start(128 * 256 + 1, 0);
  { int acc = 0; for (int g = 128 * 256; g >= 0; g--) { start[g] = acc; acc += count[g]; } }
  int n_anchors = 0; for (int c : count) n_anchors += c;
  std::vector<int> anchors(n_anchors);
  for (int q = 0; q < npx; q++) if (edge[q] == kAnchor) anchors[start[grad[q]]++] = q;

  Work W;
  memset(&W, 0, sizeof(W));
  W.w = w; W.h = h; W.grad = grad.data(); W.dir = dir.data(); W.edge = edge.data();
  W.anchors = anchors.data(); W.n_anchors = n_anchors;
  std::vector<int> pixels(npx + 8), chain_nos(npx / 2 + 64), seg_px(npx + 8), seg_off(npx / 8 + 64);
  std::vector<Node> stack(npx / 2 + 64);
  std::vector<Chain> chains(npx / 2 + 64);
  std::vector<Line> lines(npx / 8 + 64);
  W.pixels = pixels.data(); W.pixels_cap = (int)pixels.size();
  W.stack = stack.data(); W.stack_cap = (int)stack.size();
  W.chains = chains.data(); W.chains_cap = (int)chains.size();
  W.chain_nos = chain_nos.data(); W.chain_nos_cap = (int)chain_nos.size();
  W.seg_px = seg_px.data(); W.seg_px_cap = (int)seg_px.size();
  W.seg_off = seg_off.data(); W.seg_cap = (int)seg_off.size();
  W.lines = lines.data(); W.lines_cap = (int)lines.size();
  W.src = roi; W.src_stride = stride; W.src_magic = src_magic_of(w);
  std::vector<double> lut(kAtanLut + 1);
  host::atan_table(lut.data());
  W.atan_lut = lut.data();
  // validation thresholds for every pixel count a 2-pixel-wide rectangle inside the image can have (a joined line can be much longer
  // than its pixel count says); cached per geometry
  static std::map<std::pair<int, int>, std::vector<int>> cache;
  std::vector<int>& min_k = cache[std::make_pair(w, h)];
  if (min_k.empty()) {
    min_k.resize(8192);
    if (!host::nfa_table(w, h, (int)min_k.size(), min_k.data())) { min_k.clear(); return -2; }
  }
  W.nfa_min_k = min_k.data(); W.nfa_n = (int)min_k.size();
  W.min_line_len = host::min_line_len(w, h);
  run_task(W);
  if (dbg) {
    dbg->smooth = smooth; dbg->edge = edge; dbg->err = W.err;
    dbg->seg_px.assign(seg_px.begin(), seg_px.begin() + (W.nseg ? seg_off[W.nseg] : 0));
    dbg->seg_off.assign(seg_off.begin(), seg_off.begin() + W.nseg + 1);
  }
  if (W.err) return -1;
  // EDLines.cpp:60-66: lines_ED.push_back({sx, sy, ex, ey}) -- Vec4f from doubles
  for (int i = 0; i < W.nlines; i++) {
    seg.push_back((float)W.lines[i].sx); seg.push_back((float)W.lines[i].sy); seg.push_back((float)W.lines[i].ex); seg.push_back((float)W.lines[i].ey);
  }
  return W.nlines;
}

}  // namespace orc

extern "C" {
// test hook: EDLines on ONE image given as LSDDetectorC would hand it over (padded level: roi pointer + stride of the padded buffer)
int orc_ed_level(const uint8_t* img, int w, int h, int stride, float* lines, int line_cap, uint8_t* smooth, uint8_t* edge, int* seg_px, int seg_px_cap,
                 int* seg_off, int seg_cap, int* nseg, int* err) {
  std::vector<orc::PaddedLevel> pyr;
  std::vector<float> isf(1, 1.0f);
  orc::build_padded_pyramid(img, w, h, stride, isf, pyr);
  std::vector<float> seg;
  orc::EdDebug dbg;
  const int n = orc::ed_detect(pyr[0].roi(), w, h, pyr[0].stride(), seg, &dbg);
  if (err) *err = dbg.err;
  if (smooth) memcpy(smooth, dbg.smooth.data(), (size_t)w * h);
  if (edge) memcpy(edge, dbg.edge.data(), (size_t)w * h);
  if (nseg) *nseg = (int)dbg.seg_off.size() - 1;
  if (seg_off) for (size_t i = 0; i < dbg.seg_off.size() && (int)i < seg_cap; i++) seg_off[i] = dbg.seg_off[i];
  if (seg_px) for (size_t i = 0; i < dbg.seg_px.size() && (int)i < seg_px_cap; i++) seg_px[i] = dbg.seg_px[i];
  for (int i = 0; i < n && i < line_cap; i++) memcpy(lines + 4 * i, &seg[4 * i], 16);
  return n;
}
}
