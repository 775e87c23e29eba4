// oracle/line_oracle.cpp -- ORACLE (test infrastructure): CPU restatement of the reference line front-end:
//   Lineextractor::operator()            src/Lineextractor.cc:42-99
//   LSDDetectorC::ComputePyramid/detect  3rdparty/line_descriptor/src/LSDDetector_custom.cpp:76-138, 254-369
//   BinaryDescriptor::compute / computeSobel / computeLBD / binaryConversion
//                                        3rdparty/line_descriptor/src/binary_descriptor_custom.cpp:74-107,217-259,350-412,524-687,1026-1372
// Oracle decisions: no FMA; cos/sin/atan2 evaluated in double and rounded to float; sqrt in float
// (cv::sqrt == std::sqrt(float) inside namespace cv); LBD with more than 2 octaves ignores the fictitious entries.
#include "oracle_internal.h"
#include "../include/sdpl_trig.h"
#include <cmath>
#include <cstring>
#include <algorithm>

namespace orc {

static const int kCombinations[32][2] = {{0, 1}, {0, 2}, {0, 3}, {0, 4}, {0, 5}, {0, 6}, {1, 2}, {1, 3}, {1, 4}, {1, 5}, {1, 6},
                                         {2, 3}, {2, 4}, {2, 5}, {2, 6}, {2, 7}, {2, 8}, {3, 4}, {3, 5}, {3, 6}, {3, 7}, {3, 8},
                                         {4, 5}, {4, 6}, {4, 7}, {4, 8}, {5, 6}, {5, 7}, {5, 8}, {6, 7}, {6, 8}, {7, 8}};

// BinaryDescriptor::computeLBD for one line (binary_descriptor_custom.cpp:1026-1372), 9 bands of width 7.
static void lbd_one(const int16_t* dxImg, const int16_t* dyImg, int realWidth, int realHeight, const orc_keyline& kl,
                    const double* gaussCoefL, const double* gaussCoefG, float* desVec) {
  const int NB = 9, WB = 7;
  const short heightOfLSP = (short)(WB * NB);
  float pgdLBandSum[NB] = {0}, ngdLBandSum[NB] = {0}, pgdL2BandSum[NB] = {0}, ngdL2BandSum[NB] = {0};
  float pgdOBandSum[NB] = {0}, ngdOBandSum[NB] = {0}, pgdO2BandSum[NB] = {0}, ngdO2BandSum[NB] = {0};
  const short imageWidth = (short)(realWidth - 1), imageHeight = (short)(realHeight - 1);
  const short lengthOfLSP = (short)kl.num_pixels;
  const short halfHeight = (short)((heightOfLSP - 1) / 2), halfWidth = (short)((lengthOfLSP - 1) / 2);
  const float midX = (float)(0.5 * (kl.sx_oct + kl.ex_oct)), midY = (float)(0.5 * (kl.sy_oct + kl.ey_oct));
  float dL[2], dO[2];
  dL[0] = (float)sdpl_cos((double)kl.angle);   // include/sdpl_trig.h, shared with the CUDA kernel (oracle decision ix)
  dL[1] = (float)sdpl_sin((double)kl.angle);
  dO[0] = -dL[1];
  dO[1] = dL[0];
  volatile float t0, t1;
  t0 = -dL[0] * halfWidth; t1 = dL[1] * halfHeight; t0 = t0 + t1;
  float sCorX0 = t0 + midX;
  t0 = -dL[1] * halfWidth; t1 = dL[0] * halfHeight; t0 = t0 - t1;
  float sCorY0 = t0 + midY;
  for (short hID = 0; hID < heightOfLSP; hID++) {
    float sCorX = sCorX0, sCorY = sCorY0;
    float pgdLRowSum = 0, ngdLRowSum = 0, pgdORowSum = 0, ngdORowSum = 0;
    for (short wID = 0; wID < lengthOfLSP; wID++) {
      short tempCor = (short)std::round(sCorX);
      short xCor = (tempCor < 0) ? 0 : (tempCor > imageWidth) ? imageWidth : tempCor;
      tempCor = (short)std::round(sCorY);
      short yCor = (tempCor < 0) ? 0 : (tempCor > imageHeight) ? imageHeight : tempCor;
      short dx = dxImg[yCor * realWidth + xCor], dy = dyImg[yCor * realWidth + xCor];
      volatile float a = dx * dL[0], b = dy * dL[1];
      float gDL = a + b;
      a = dx * dO[0]; b = dy * dO[1];
      float gDO = a + b;
      if (gDL > 0) pgdLRowSum += gDL; else ngdLRowSum -= gDL;
      if (gDO > 0) pgdORowSum += gDO; else ngdORowSum -= gDO;
      sCorX += dL[0];
      sCorY += dL[1];
    }
    sCorX0 -= dL[1];
    sCorY0 += dL[0];
    float coef = (float)gaussCoefG[hID];
    pgdLRowSum = coef * pgdLRowSum; ngdLRowSum = coef * ngdLRowSum;
    float pgdL2RowSum = pgdLRowSum * pgdLRowSum, ngdL2RowSum = ngdLRowSum * ngdLRowSum;
    pgdORowSum = coef * pgdORowSum; ngdORowSum = coef * ngdORowSum;
    float pgdO2RowSum = pgdORowSum * pgdORowSum, ngdO2RowSum = ngdORowSum * ngdORowSum;
    auto add = [&](short band, float c) {
      volatile float m;
      volatile float c2 = c * c;
      m = c * pgdLRowSum; pgdLBandSum[band] = pgdLBandSum[band] + m;
      m = c * ngdLRowSum; ngdLBandSum[band] = ngdLBandSum[band] + m;
      m = c2 * pgdL2RowSum; pgdL2BandSum[band] = pgdL2BandSum[band] + m;
      m = c2 * ngdL2RowSum; ngdL2BandSum[band] = ngdL2BandSum[band] + m;
      m = c * pgdORowSum; pgdOBandSum[band] = pgdOBandSum[band] + m;
      m = c * ngdORowSum; ngdOBandSum[band] = ngdOBandSum[band] + m;
      m = c2 * pgdO2RowSum; pgdO2BandSum[band] = pgdO2BandSum[band] + m;
      m = c2 * ngdO2RowSum; ngdO2BandSum[band] = ngdO2BandSum[band] + m;
    };
    short bandID = (short)(hID / WB);
    add(bandID, (float)gaussCoefL[hID % WB + WB]);
    bandID--;
    if (bandID >= 0) add(bandID, (float)gaussCoefL[hID % WB + 2 * WB]);
    bandID = (short)(bandID + 2);
    if (bandID < NB) add(bandID, (float)gaussCoefL[hID % WB]);
  }
  const float invN2 = (float)(1.0 / (WB * 2.0)), invN3 = (float)(1.0 / (WB * 3.0));
  for (short bandID = 0; bandID < NB; bandID++) {
    float invN = (bandID == 0 || bandID == NB - 1) ? invN2 : invN3;
    short d = (short)(bandID * 8);
    volatile float temp, q, r;
    temp = pgdLBandSum[bandID] * invN; desVec[d] = temp;
    q = pgdL2BandSum[bandID] * invN; r = temp * temp; q = q - r; desVec[d + 4] = std::sqrt((float)q);
    temp = ngdLBandSum[bandID] * invN; desVec[d + 1] = temp;
    q = ngdL2BandSum[bandID] * invN; r = temp * temp; q = q - r; desVec[d + 5] = std::sqrt((float)q);
    temp = pgdOBandSum[bandID] * invN; desVec[d + 2] = temp;
    q = pgdO2BandSum[bandID] * invN; r = temp * temp; q = q - r; desVec[d + 6] = std::sqrt((float)q);
    temp = ngdOBandSum[bandID] * invN; desVec[d + 3] = temp;
    q = ngdO2BandSum[bandID] * invN; r = temp * temp; q = q - r; desVec[d + 7] = std::sqrt((float)q);
  }
  volatile float tempM = 0, tempS = 0, m;
  for (int base = 0; base < NB; base++) {
    const float* v = desVec + base * 8;
    for (int k = 0; k < 4; k++) { m = v[k] * v[k]; tempM = tempM + m; }
    for (int k = 4; k < 8; k++) { m = v[k] * v[k]; tempS = tempS + m; }
  }
  tempM = 1 / std::sqrt((float)tempM);
  tempS = 1 / std::sqrt((float)tempS);
  for (int base = 0; base < NB; base++) {
    float* v = desVec + base * 8;
    for (int k = 0; k < 4; k++) v[k] = v[k] * tempM;
    for (int k = 4; k < 8; k++) v[k] = v[k] * tempS;
  }
  for (int i = 0; i < NB * 8; i++) if (desVec[i] > 0.4) desVec[i] = (float)0.4;
  volatile float temp = 0;
  for (int i = 0; i < NB * 8; i++) { m = desVec[i] * desVec[i]; temp = temp + m; }
  temp = 1 / std::sqrt((float)temp);
  for (int i = 0; i < NB * 8; i++) desVec[i] = desVec[i] * temp;
}

// BinaryDescriptor::compute -> computeImpl(returnFloat=false, useDetectionData=false), binary_descriptor_custom.cpp:539-687
int lbd_compute(const uint8_t* img, int w, int h, int stride, const orc_keyline* kls, int n, uint8_t* desc, float* fdesc) {
  if (n == 0) return 0;  // "Error: keypoint list is empty": descriptors left untouched (:556-560)
  int octaveIndex = -1;
  for (int i = 0; i < n; i++) octaveIndex = std::max(octaveIndex, kls[i].octave);
  const int noct = octaveIndex + 1;
  // weights, BinaryDescriptor::BinaryDescriptor, :217-259 (note the integer divisions)
  const int WB = 7, NB = 9;
  double gaussCoefL[WB * 3], gaussCoefG[NB * WB];
  {
    double u = (WB * 3 - 1) / 2;
    double sigma = (WB * 2 + 1) / 2;
    double invsigma2 = -1 / (2 * sigma * sigma);
    for (int i = 0; i < WB * 3; i++) { double dis = i - u; gaussCoefL[i] = std::exp(dis * dis * invsigma2); }
    u = (NB * WB - 1) / 2;
    sigma = u;
    invsigma2 = -1 / (2 * sigma * sigma);
    for (int i = 0; i < NB * WB; i++) { double dis = i - u; gaussCoefG[i] = std::exp(dis * dis * invsigma2); }
  }
  // computeSobel / computeGaussianPyramid, :350-398
  std::vector<std::vector<int16_t>> dxs(noct), dys(noct);
  std::vector<int> ws(noct), hs(noct);
  std::vector<uint8_t> cur((size_t)w * h), nxt;
  gaussian_blur_u8(img, w, h, stride, cur.data(), w, 1);
  int cw = w, ch = h;
  for (int o = 0; o < noct; o++) {
    if (o > 0) {
      int nw = cw / 2, nh = ch / 2;
      if (nw < 1 || nh < 1) return -1;
      nxt.assign((size_t)nw * nh, 0);
      pyrdown_u8(cur.data(), cw, ch, cw, nxt.data(), nw, nh, nw);
      cur.swap(nxt); cw = nw; ch = nh;
    }
    ws[o] = cw; hs[o] = ch;
    dxs[o].resize((size_t)cw * ch); dys[o].resize((size_t)cw * ch);
    sobel3_s16(cur.data(), cw, ch, cw, dxs[o].data(), dys[o].data());
  }
  float f[72];
  for (int i = 0; i < n; i++) {
    const int o = kls[i].octave;
    if (o < 0) return -2;
    lbd_one(dxs[o].data(), dys[o].data(), ws[o], hs[o], kls[i], gaussCoefL, gaussCoefG, f);
    if (fdesc) memcpy(fdesc + (size_t)i * 72, f, sizeof(f));
    if (desc) {
      for (int c = 0; c < 32; c++) {
        const float* f1 = f + 8 * kCombinations[c][0];
        const float* f2 = f + 8 * kCombinations[c][1];
        uint8_t r = 0;
        for (int k = 0; k < 8; k++) if (f1[k] > f2[k]) r = (uint8_t)(r + (1 << k));
        desc[(size_t)i * 32 + c] = r;
      }
    }
  }
  return n;
}

}  // namespace orc

using namespace orc;

struct orc_line {
  int nfeatures, refine, nlevels, extractor;
  float lsd_scale, scale;
  std::vector<float> sf, isf;
  std::vector<std::vector<float>> last_segments;
};

extern "C" {

orc_line* orc_line_create(int nfeatures, int refine, float lsd_scale, int nlevels, float scale, int extractor) {
  if (nlevels < 1 || (extractor != 0 && extractor != 1)) return nullptr;
  orc_line* o = new orc_line{nfeatures, refine, nlevels, extractor, lsd_scale, scale, {}, {}, {}};
  // LSDDetectorC::ComputePyramid scale tables, LSDDetector_custom.cpp:79-90
  o->sf.resize(nlevels); o->isf.resize(nlevels);
  o->sf[0] = 1.0f;
  for (int l = 0; l < nlevels; l++) {
    if (l > 0) o->sf[l] = o->sf[l - 1] * scale;
    o->isf[l] = 1.0f / o->sf[l];
  }
  return o;
}
void orc_line_destroy(orc_line* o) { delete o; }

// Lineextractor.cc:84-96: mvLevelSigma2_l[0]=1, [i]=sf^2 ; mvInvLevelSigma2_l = 1/sigma2
void orc_line_tables(const orc_line* o, float* sf, float* isf, float* s2, float* is2) {
  for (int i = 0; i < o->nlevels; i++) {
    float sig = i == 0 ? 1.0f : o->sf[i] * o->sf[i];
    if (sf) sf[i] = o->sf[i];
    if (isf) isf[i] = o->isf[i];
    if (s2) s2[i] = sig;
    if (is2) is2[i] = 1.0f / sig;
  }
}

int orc_line_extract(orc_line* o, const uint8_t* img, int w, int h, int stride, orc_keyline* kls, uint8_t* desc, int cap) {
  if (!img || w <= 0 || h <= 0) return 0;
  std::vector<PaddedLevel> pyr;
  build_padded_pyramid(img, w, h, stride, o->isf, pyr);
  const double min_length = 0.02 * (std::min(w, h));            // Lineextractor.cc:46,70
  std::vector<orc_keyline> out;
  o->last_segments.assign(o->nlevels, {});
  int class_counter = -1;
  for (int oct = 0; oct < o->nlevels; oct++) {
    const PaddedLevel& L = pyr[oct];
    std::vector<float>& seg = o->last_segments[oct];
    // extractor 0: LSDDetectorC::detectImpl (LSDDetector_custom.cpp:254-369); 1: detectImpl_ED (:386-461), same key-line construction
    int n = o->extractor == 1 ? ed_detect(L.roi(), L.w, L.h, L.stride(), seg, nullptr)
                              : lsd_detect(L.roi(), L.w, L.h, L.stride(), o->refine, (double)o->lsd_scale, 0.6, 2.0, 22.5, 0.0, 0.8, 1024, 0, seg);
    if (n < 0) return -1;
    const float octaveScale = (float)std::pow((double)o->scale, (double)oct);
    for (int k = 0; k < n; k++) {
      float e[4] = {seg[4 * k], seg[4 * k + 1], seg[4 * k + 2], seg[4 * k + 3]};
      // checkLineExtremes, LSDDetector_custom.cpp:112-138
      if (e[0] < 0) e[0] = 0;
      if (e[0] >= L.w) e[0] = (float)L.w - 1.0f;
      if (e[2] < 0) e[2] = 0;
      if (e[2] >= L.w) e[2] = (float)L.w - 1.0f;
      if (e[1] < 0) e[1] = 0;
      if (e[1] >= L.h) e[1] = (float)L.h - 1.0f;
      if (e[3] < 0) e[3] = 0;
      if (e[3] >= L.h) e[3] = (float)L.h - 1.0f;
      volatile float ddx = e[0] - e[2], ddy = e[1] - e[3];
      double length = (float)std::sqrt(std::pow((double)ddx, 2) + std::pow((double)ddy, 2));
      if (!(length > min_length)) continue;
      orc_keyline kl;
      kl.sx = e[0] * octaveScale; kl.sy = e[1] * octaveScale; kl.ex = e[2] * octaveScale; kl.ey = e[3] * octaveScale;
      kl.sx_oct = e[0]; kl.sy_oct = e[1]; kl.ex_oct = e[2]; kl.ey_oct = e[3];
      kl.length = (float)length;
      int x0 = cv_round(e[0]), y0 = cv_round(e[1]), x1 = cv_round(e[2]), y1 = cv_round(e[3]);   // LineIterator, 8-connected
      kl.num_pixels = std::max(std::abs(x1 - x0), std::abs(y1 - y0)) + 1;
      volatile float ay = kl.ey - kl.sy, ax = kl.ex - kl.sx;
      kl.angle = (float)std::atan2((double)ay, (double)ax);
      kl.class_id = ++class_counter;
      kl.octave = oct;
      kl.size = ax * ay;
      kl.response = kl.length / (float)std::max(L.w, L.h);
      volatile float mx = kl.ex + kl.sx, my = kl.ey + kl.sy;
      kl.pt_x = mx / 2; kl.pt_y = my / 2;
      out.push_back(kl);
    }
  }
  if ((int)out.size() > o->nfeatures && o->nfeatures != 0) {
    // Lineextractor.cc:73-82 uses std::sort (unstable); oracle decision: response descending, ties by original order
    std::stable_sort(out.begin(), out.end(), [](const orc_keyline& a, const orc_keyline& b) { return a.response > b.response; });
    out.resize(o->nfeatures);
    for (int i = 0; i < o->nfeatures; i++) out[i].class_id = i;
  }
  int n = (int)out.size();
  if (n > cap) return n;
  for (int i = 0; i < n; i++) kls[i] = out[i];
  if (n > 0 && desc) {
    int r = lbd_compute(img, w, h, stride, out.data(), n, desc, nullptr);
    if (r < 0) return -3;
  }
  return n;
}

int orc_line_last_segments(const orc_line* o, int octave, float* out, int cap) {
  const auto& s = o->last_segments[octave];
  int n = (int)s.size() / 4;
  for (int i = 0; i < n && i < cap; i++) memcpy(out + 4 * i, &s[4 * i], 16);
  return n;
}

int orc_lbd_compute(const uint8_t* img, int w, int h, int stride, const orc_keyline* kls, int n, uint8_t* desc, float* fdesc) {
  return orc::lbd_compute(img, w, h, stride, kls, n, desc, fdesc);
}
}
