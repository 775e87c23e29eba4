// oracle/refshim/shim_prims.cpp -- ORACLE test infrastructure, NOT the product.
// The OpenCV image primitives the reference calls, forwarded to the cv2-pinned restatements in oracle/cvprim.cpp and
// oracle/lsd_oracle.cpp (tests/test_oracle_vs_cv2.py pins each bit-exact against python cv2 4.13).  Anything outside what
// the reference's hot path executes aborts loudly (shim_unsupported).
#include "sdpl_cvshim.hpp"
#include "../oracle_internal.h"

namespace cv {

float fastAtan2(float y, float x) { return orc::fast_atan2(y, x); }

static void need_u8(const Mat& m, const char* who) {
  if (m.type() != CV_8UC1 || m.empty()) shim_unsupported(who);
}

void resize(InputArray _src, OutputArray _dst, Size dsize, double fx, double fy, int interpolation) {
  Mat src = _src.getMat();
  need_u8(src, "resize of a non-8UC1 image");
  if (interpolation != INTER_LINEAR || dsize.width <= 0 || dsize.height <= 0 || fx != 0 || fy != 0)
    shim_unsupported("resize other than INTER_LINEAR to a given dsize");
  _dst.create(dsize, src.type());
  Mat dst = _dst.getMat();
  if (dst.data == src.data) shim_unsupported("in-place resize");
  orc::resize_linear_u8(src.data, src.cols, src.rows, (int)src.step, dst.data, dst.cols, dst.rows, (int)dst.step);
}

void copyMakeBorder(InputArray _src, OutputArray _dst, int top, int bottom, int left, int right, int borderType,
                    const Scalar&) {
  Mat src = _src.getMat();
  need_u8(src, "copyMakeBorder of a non-8UC1 image");
  const bool isolated = (borderType & BORDER_ISOLATED) != 0;
  if ((borderType & ~BORDER_ISOLATED) != BORDER_REFLECT_101 || top != bottom || top != left || top != right)
    shim_unsupported("copyMakeBorder other than a uniform BORDER_REFLECT_101");
  // without BORDER_ISOLATED OpenCV takes border pixels of a ROI from the parent buffer: not modelled, refuse
  if (src.isSubmatrix() && !isolated) shim_unsupported("non-isolated copyMakeBorder of a ROI");
  _dst.create(src.rows + 2 * top, src.cols + 2 * top, src.type());
  Mat dst = _dst.getMat();
  // src may be the interior of dst (ComputePyramid): the interior is only read, row by row in place
  orc::border_reflect101_u8(src.data, src.cols, src.rows, (int)src.step, dst.data, top, (int)dst.step);
}

void GaussianBlur(InputArray _src, OutputArray _dst, Size ksize, double sigmaX, double sigmaY, int borderType) {
  Mat src = _src.getMat();
  need_u8(src, "GaussianBlur of a non-8UC1 image");
  int kind = -1;
  if (sigmaY == 0) sigmaY = sigmaX;
  if (ksize.width == 7 && ksize.height == 7 && sigmaX == 2 && sigmaY == 2) kind = 0;       // ORBextractor.cc:1084
  else if (ksize.width == 5 && ksize.height == 5 && sigmaX == 1 && sigmaY == 1) kind = 1;  // binary_descriptor_custom.cpp:358
  if (kind < 0 || (borderType & ~BORDER_ISOLATED) != BORDER_REFLECT_101) shim_unsupported("GaussianBlur kernel / border");
  if (src.isSubmatrix() && !(borderType & BORDER_ISOLATED)) {
    // OpenCV reads the pixels around a ROI from the parent buffer (ED_Lib smooths LSDDetectorC's pyramid ROI, which sits inside a
    // 19-pixel border): blur the ROI widened by the kernel radius, all of it real pixels, and keep the interior
    const int r = kind == 0 ? 3 : 2;
    if (src.roi_x < r || src.roi_y < r || src.roi_x + src.cols + r > src.whole_cols || src.roi_y + src.rows + r > src.whole_rows)
      shim_unsupported("non-isolated GaussianBlur of a ROI closer to the parent's edge than the kernel radius");
    const int W = src.cols + 2 * r, H = src.rows + 2 * r;
    Mat wide(H, W, CV_8UC1), out(H, W, CV_8UC1);
    for (int y = 0; y < H; y++) memcpy(wide.ptr(y), src.data + (ptrdiff_t)(y - r) * (ptrdiff_t)src.step.p - r, (size_t)W);
    orc::gaussian_blur_u8(wide.data, W, H, (int)wide.step, out.data, (int)out.step, kind);
    _dst.create(src.rows, src.cols, src.type());
    Mat dst = _dst.getMat();
    for (int y = 0; y < src.rows; y++) memcpy(dst.ptr(y), out.ptr(y + r) + r, (size_t)src.cols);
    return;
  }
  Mat in = (_dst.getMat().data == src.data) ? src.clone() : src;
  _dst.create(src.rows, src.cols, src.type());
  Mat dst = _dst.getMat();
  orc::gaussian_blur_u8(in.data, in.cols, in.rows, (int)in.step, dst.data, (int)dst.step, kind);
}

void pyrDown(InputArray _src, OutputArray _dst, const Size& dstsize, int borderType) {
  Mat src = _src.getMat();   // keeps the source buffer alive when dst is the same Mat object
  need_u8(src, "pyrDown of a non-8UC1 image");
  if (borderType != BORDER_DEFAULT || src.isSubmatrix()) shim_unsupported("pyrDown border / ROI");
  Size ds = (dstsize.width > 0 && dstsize.height > 0) ? dstsize : Size((src.cols + 1) / 2, (src.rows + 1) / 2);
  Mat out(ds, src.type());
  orc::pyrdown_u8(src.data, src.cols, src.rows, (int)src.step, out.data, out.cols, out.rows, (int)out.step);
  _dst.mat() = out;
}

void Sobel(InputArray _src, OutputArray _dst, int ddepth, int dx, int dy, int ksize, double scale, double delta,
           int borderType) {
  Mat src = _src.getMat();
  need_u8(src, "Sobel of a non-8UC1 image");
  if (ddepth != CV_16S || ksize != 3 || scale != 1 || delta != 0 || borderType != BORDER_DEFAULT || dx + dy != 1 ||
      src.isSubmatrix())
    shim_unsupported("Sobel other than 3x3 first derivative to CV_16S");
  _dst.create(src.rows, src.cols, CV_16SC1);
  Mat dst = _dst.getMat();
  if (!dst.isContinuous()) shim_unsupported("Sobel into a non-continuous matrix");
  Mat other(src.rows, src.cols, CV_16SC1);
  if (dx == 1) orc::sobel3_s16(src.data, src.cols, src.rows, (int)src.step, dst.ptr<short>(), other.ptr<short>());
  else orc::sobel3_s16(src.data, src.cols, src.rows, (int)src.step, other.ptr<short>(), dst.ptr<short>());
}

void FAST(InputArray _image, std::vector<KeyPoint>& keypoints, int threshold, bool nonmaxSuppression) {
  Mat img = _image.getMat();
  need_u8(img, "FAST of a non-8UC1 image");
  if (!nonmaxSuppression) shim_unsupported("FAST without non-max suppression");
  std::vector<int> xs, ys, sc;
  int n = orc::fast9_nms(img.data, img.cols, img.rows, (int)img.step, threshold, xs, ys, sc);
  keypoints.clear();
  for (int i = 0; i < n; i++) keypoints.push_back(KeyPoint((float)xs[i], (float)ys[i], 7.f, -1, (float)sc[i]));
}

void cvtColor(InputArray, OutputArray, int, int) { shim_unsupported("cvtColor"); }
void Canny(InputArray, OutputArray, double, double, int, bool) { shim_unsupported("Canny"); }
Mat abs(const Mat&) { shim_unsupported("abs(Mat)"); }
void add(InputArray, InputArray, OutputArray) { shim_unsupported("add"); }
double threshold(InputArray, OutputArray, double, double, int) { shim_unsupported("threshold"); }
void compare(InputArray, InputArray, OutputArray, int) { shim_unsupported("compare"); }

LineIterator::LineIterator(const Mat& img, Point pt1, Point pt2, int connectivity, bool) {
  if ((unsigned)pt1.x >= (unsigned)img.cols || (unsigned)pt2.x >= (unsigned)img.cols ||
      (unsigned)pt1.y >= (unsigned)img.rows || (unsigned)pt2.y >= (unsigned)img.rows)
    shim_unsupported("LineIterator with an end point outside the image (clipLine)");
  int dx = pt2.x - pt1.x, dy = pt2.y - pt1.y;
  dx = dx < 0 ? -dx : dx;
  dy = dy < 0 ? -dy : dy;
  count = connectivity == 8 ? (dx > dy ? dx : dy) + 1 : dx + dy + 1;
}

namespace {
class LsdForward : public LineSegmentDetector {
 public:
  int refine, n_bins;
  double scale, sigma_scale, quant, ang_th, log_eps, density_th;
  void detect(const Mat& image, std::vector<Vec4f>& lines) override {
    need_u8(image, "LineSegmentDetector on a non-8UC1 image");
    std::vector<float> out;
    // the ROI of a reflect-101 padded level: the internal blur reads real border pixels = what an isolated reflect-101
    // border gives, so the restatement on the ROI alone is the same computation
    int n = orc::lsd_detect(image.data, image.cols, image.rows, (int)image.step, refine, scale, sigma_scale, quant, ang_th,
                            log_eps, density_th, n_bins, 0, out);
    lines.clear();
    for (int i = 0; i < n; i++) lines.push_back(Vec4f(out[4 * i], out[4 * i + 1], out[4 * i + 2], out[4 * i + 3]));
  }
};
}  // namespace

Ptr<LineSegmentDetector> createLineSegmentDetector(int refine, double scale, double sigma_scale, double quant,
                                                   double ang_th, double log_eps, double density_th, int n_bins) {
  LsdForward* l = new LsdForward();
  l->refine = refine; l->scale = scale; l->sigma_scale = sigma_scale; l->quant = quant; l->ang_th = ang_th;
  l->log_eps = log_eps; l->density_th = density_th; l->n_bins = n_bins;
  return Ptr<LineSegmentDetector>(l);
}

}  // namespace cv
