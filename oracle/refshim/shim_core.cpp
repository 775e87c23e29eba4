// oracle/refshim/shim_core.cpp -- ORACLE test infrastructure, NOT the product.
// Container half of the OpenCV stand-in (include/sdpl_cvshim.hpp): cv::Mat storage, views, copies.  No image arithmetic
// and no dependency on oracle/*.cpp, so the adapter check (adapter_check.cpp) can link it next to the product library.
#include "sdpl_cvshim.hpp"
#include <stdio.h>

namespace cv {

void shim_unsupported(const char* what) {
  fprintf(stderr, "sdpl cv shim: %s is not on the front-end hot path and is not implemented\n", what);
  abort();
}

void Mat::create(int r, int c, int type) {
  if (data && rows == r && cols == c && type_ == type) return;   // cv::Mat::create keeps a matching buffer (ROI included)
  type_ = type;
  rows = r; cols = c;
  step = (size_t)c * elemSize();
  parent_ = false;
  roi_x = roi_y = 0; whole_cols = whole_rows = 0;
  size_t bytes = step.p * (size_t)r;
  buf_ = std::shared_ptr<uchar>((uchar*)malloc(bytes ? bytes : 1), free);
  data = buf_.get();
}

Mat Mat::operator()(const Rect& r) const {
  if (r.x < 0 || r.y < 0 || r.width < 0 || r.height < 0 || r.x + r.width > cols || r.y + r.height > rows)
    shim_unsupported("Mat ROI outside the matrix");
  Mat m(*this);
  m.data = data + step.p * r.y + (size_t)r.x * elemSize();
  m.rows = r.height; m.cols = r.width;
  if (r.width != cols || r.height != rows) {
    m.parent_ = true;
    if (!parent_) { m.whole_cols = cols; m.whole_rows = rows; m.roi_x = 0; m.roi_y = 0; }
    m.roi_x += r.x; m.roi_y += r.y;
  }
  return m;
}

void Mat::copyTo(Mat& dst) const {
  if (empty()) { dst.release(); return; }
  dst.create(rows, cols, type_);
  if (dst.data == data) return;
  for (int y = 0; y < rows; y++) memcpy(dst.ptr(y), ptr(y), (size_t)cols * elemSize());
}
void Mat::copyTo(const _OutputArray& dst) const { copyTo(dst.mat()); }

Mat Mat::clone() const { Mat m; copyTo(m); return m; }

void Mat::push_back(const Mat& m) {
  if (m.empty()) return;
  if (empty()) { *this = m.clone(); return; }
  if (m.cols != cols || m.type() != type_) shim_unsupported("Mat::push_back with a different row type");
  Mat out(rows + m.rows, cols, type_);
  for (int y = 0; y < rows; y++) memcpy(out.ptr(y), ptr(y), (size_t)cols * elemSize());
  for (int y = 0; y < m.rows; y++) memcpy(out.ptr(rows + y), m.ptr(y), (size_t)cols * elemSize());
  *this = out;
}

void Mat::convertTo(Mat& dst, int type) const {
  if (type == type_) { copyTo(dst); return; }
  if (type_ == CV_32S && type == CV_32F) {
    Mat out(rows, cols, type);
    for (int y = 0; y < rows; y++)
      for (int x = 0; x < cols; x++) out.at<float>(y, x) = (float)at<int>(y, x);
    dst = out;
    return;
  }
  shim_unsupported("Mat::convertTo for this depth pair");
}

}  // namespace cv
