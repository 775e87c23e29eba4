// oracle/refshim/include/sdpl_cvshim.hpp -- ORACLE test infrastructure, NOT the product.
//
// A minimal stand-in for the OpenCV 3.x C++ API, just wide enough that the UNMODIFIED reference sources
//   /root/reference/src/ORBextractor.cc, src/Lineextractor.cc,
//   3rdparty/line_descriptor/src/{LSDDetector_custom,binary_descriptor_custom,binary_descriptor_matcher}.cpp
// compile in an image without OpenCV headers (oracle/refshim/Makefile -> oracle/_ref/libsdpl_ref.so).
// Containers (Mat, Point_, KeyPoint ...) are implemented here; the image primitives the reference calls
// (resize, copyMakeBorder, FAST, GaussianBlur, fastAtan2, pyrDown, Sobel, LineSegmentDetector) are declared here and
// forwarded in shim_prims.cpp to oracle/cvprim.cpp / oracle/lsd_oracle.cpp, which are pinned bit-exact against cv2 4.13
// (tests/test_oracle_vs_cv2.py).  Functions the hot path never executes (EDLine detector, FileStorage, drawing,
// cvtColor, Canny) only need to compile: their shim bodies abort with a message.
#ifndef SDPL_CVSHIM_HPP
#define SDPL_CVSHIM_HPP
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <float.h>
#include <limits.h>
#include <assert.h>
#include <time.h>
#include <cmath>
#include <memory>
#include <vector>
#include <string>
#include <iostream>
#include <algorithm>

#define CV_EXPORTS
#define CV_EXPORTS_W
#define CV_WRAP
#define CV_OUT
#define CV_IN_OUT
#define CV_PI 3.1415926535897932384626433832795
#define CV_8U 0
#define CV_8S 1
#define CV_16U 2
#define CV_16S 3
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6
#define CV_8UC1 CV_8U
#define CV_8SC1 CV_8S
#define CV_16SC1 CV_16S
#define CV_32SC1 CV_32S
#define CV_32FC1 CV_32F
#define CV_64FC1 CV_64F

typedef unsigned char uchar;
typedef signed char schar;
typedef unsigned short ushort;

// cvRound = cvtsd2si / cvtss2si under the default rounding mode: round half to even.
inline int cvRound(double v) { return (int)lrint(v); }
inline int cvRound(float v) { return (int)lrintf(v); }
inline int cvRound(int v) { return v; }
inline int cvFloor(double v) { int i = (int)v; return i - (i > v); }
inline int cvFloor(float v) { int i = (int)v; return i - (i > v); }
inline int cvFloor(int v) { return v; }
inline int cvCeil(double v) { int i = (int)v; return i + (i < v); }
inline int cvCeil(float v) { int i = (int)v; return i + (i < v); }
inline int cvCeil(int v) { return v; }

namespace cv {
using std::abs;
using std::exp;
using std::log;
using std::max;
using std::min;
using std::pow;
using std::sqrt;
using std::swap;

[[noreturn]] void shim_unsupported(const char* what);
#define CV_Assert(expr) do { if (!(expr)) cv::shim_unsupported("CV_Assert(" #expr ")"); } while (0)

template <typename T> inline T saturate_cast(float v) { return (T)v; }
template <typename T> inline T saturate_cast(double v) { return (T)v; }
template <typename T> inline T saturate_cast(int v) { return (T)v; }
template <> inline int saturate_cast<int>(float v) { return cvRound(v); }
template <> inline int saturate_cast<int>(double v) { return cvRound(v); }

template <typename T> struct Point_ {
  T x, y;
  Point_() : x(0), y(0) {}
  Point_(T _x, T _y) : x(_x), y(_y) {}
  template <typename U> operator Point_<U>() const { return Point_<U>(saturate_cast<U>(x), saturate_cast<U>(y)); }
};
template <typename T> inline Point_<T>& operator*=(Point_<T>& a, int b) {
  a.x = saturate_cast<T>(a.x * b); a.y = saturate_cast<T>(a.y * b); return a;
}
template <typename T> inline Point_<T>& operator*=(Point_<T>& a, float b) {
  a.x = saturate_cast<T>(a.x * b); a.y = saturate_cast<T>(a.y * b); return a;
}
template <typename T> inline Point_<T>& operator*=(Point_<T>& a, double b) {
  a.x = saturate_cast<T>(a.x * b); a.y = saturate_cast<T>(a.y * b); return a;
}
template <typename T> inline Point_<T> operator+(const Point_<T>& a, const Point_<T>& b) { return Point_<T>(a.x + b.x, a.y + b.y); }
template <typename T> inline Point_<T> operator-(const Point_<T>& a, const Point_<T>& b) { return Point_<T>(a.x - b.x, a.y - b.y); }
template <typename T> inline Point_<T> operator*(const Point_<T>& a, double b) { return Point_<T>(saturate_cast<T>(a.x * b), saturate_cast<T>(a.y * b)); }
typedef Point_<int> Point2i;
typedef Point_<int> Point;
typedef Point_<float> Point2f;
typedef Point_<double> Point2d;

template <typename T> struct Size_ {
  T width, height;
  Size_() : width(0), height(0) {}
  Size_(T w, T h) : width(w), height(h) {}
  bool operator==(const Size_& o) const { return width == o.width && height == o.height; }
  bool operator!=(const Size_& o) const { return !(*this == o); }
};
typedef Size_<int> Size;
template <typename T> inline std::ostream& operator<<(std::ostream& os, const Size_<T>& s) {
  return os << "[" << s.width << " x " << s.height << "]";
}

template <typename T> struct Rect_ {
  T x, y, width, height;
  Rect_() : x(0), y(0), width(0), height(0) {}
  Rect_(T _x, T _y, T w, T h) : x(_x), y(_y), width(w), height(h) {}
};
typedef Rect_<int> Rect;

template <typename T, int N> struct Vec {
  T val[N];
  Vec() { for (int i = 0; i < N; i++) val[i] = T(0); }
  Vec(T a, T b, T c, T d) { static_assert(N == 4, "Vec4 only"); val[0] = a; val[1] = b; val[2] = c; val[3] = d; }
  T& operator[](int i) { return val[i]; }
  const T& operator[](int i) const { return val[i]; }
};
typedef Vec<float, 4> Vec4f;
typedef Vec<int, 4> Vec4i;

template <typename T> struct Scalar_ {
  T val[4];
  Scalar_() { val[0] = val[1] = val[2] = val[3] = 0; }
  Scalar_(T a, T b = 0, T c = 0, T d = 0) { val[0] = a; val[1] = b; val[2] = c; val[3] = d; }
  static Scalar_ all(T v) { return Scalar_(v, v, v, v); }
};
typedef Scalar_<double> Scalar;

struct KeyPoint {
  Point2f pt;
  float size, angle, response;
  int octave, class_id;
  KeyPoint() : pt(0, 0), size(0), angle(-1), response(0), octave(0), class_id(-1) {}
  KeyPoint(Point2f _pt, float _size, float _angle = -1, float _response = 0, int _octave = 0, int _class_id = -1)
      : pt(_pt), size(_size), angle(_angle), response(_response), octave(_octave), class_id(_class_id) {}
  KeyPoint(float x, float y, float _size, float _angle = -1, float _response = 0, int _octave = 0, int _class_id = -1)
      : pt(x, y), size(_size), angle(_angle), response(_response), octave(_octave), class_id(_class_id) {}
};

struct DMatch {
  int queryIdx, trainIdx, imgIdx;
  float distance;
  DMatch() : queryIdx(-1), trainIdx(-1), imgIdx(-1), distance(FLT_MAX) {}
  DMatch(int q, int t, float d) : queryIdx(q), trainIdx(t), imgIdx(-1), distance(d) {}
  DMatch(int q, int t, int i, float d) : queryIdx(q), trainIdx(t), imgIdx(i), distance(d) {}
  bool operator<(const DMatch& m) const { return distance < m.distance; }
};

template <typename T> class Ptr : public std::shared_ptr<T> {
 public:
  Ptr() {}
  Ptr(T* p) : std::shared_ptr<T>(p) {}
  template <typename U> Ptr(const Ptr<U>& o) : std::shared_ptr<T>(o) {}
  bool empty() const { return !this->get(); }
  operator T*() const { return this->get(); }
};

class FileStorage;
class FileNode {
 public:
  FileNode operator[](const char*) const { shim_unsupported("FileNode"); }
  operator int() const { shim_unsupported("FileNode"); }
};
class FileStorage {};
template <typename T> inline FileStorage& operator<<(FileStorage&, const T&) { shim_unsupported("FileStorage"); }

class Algorithm {
 public:
  virtual ~Algorithm() {}
};

enum { BORDER_CONSTANT = 0, BORDER_REPLICATE = 1, BORDER_REFLECT = 2, BORDER_WRAP = 3, BORDER_REFLECT_101 = 4,
       BORDER_REFLECT101 = 4, BORDER_DEFAULT = 4, BORDER_ISOLATED = 16 };
enum { INTER_NEAREST = 0, INTER_LINEAR = 1, INTER_CUBIC = 2, INTER_AREA = 3 };
enum { NORM_HAMMING = 6 };
enum { COLOR_BGR2GRAY = 6 };
enum { THRESH_TOZERO = 3 };
enum { CMP_LT = 3 };
enum { LSD_REFINE_NONE = 0, LSD_REFINE_STD = 1, LSD_REFINE_ADV = 2 };

struct MatStep {
  size_t p;
  MatStep() : p(0) {}
  operator size_t() const { return p; }
  MatStep& operator=(size_t s) { p = s; return *this; }
};

class _OutputArray;
// Mat::zeros() is a matrix EXPRESSION in OpenCV: assigning it to a Mat whose size and type already match writes the zeros
// into the existing buffer (computeDescriptors, src/ORBextractor.cc:1026, relies on this to fill a rowRange() view).
struct MatExprZeros { int rows, cols, type; };
class Mat {
 public:
  int rows, cols;
  uchar* data;
  MatStep step;

  int roi_x = 0, roi_y = 0, whole_cols = 0, whole_rows = 0;   // position of a ROI inside its parent buffer (OpenCV: locateROI)
  Mat() : rows(0), cols(0), data(0), type_(0), parent_(false) {}
  Mat(const MatExprZeros& e) : rows(0), cols(0), data(0), type_(0), parent_(false) { *this = e; }
  Mat& operator=(const MatExprZeros& e) {
    create(e.rows, e.cols, e.type);
    for (int y = 0; y < rows; y++) memset(ptr(y), 0, (size_t)cols * elemSize());
    return *this;
  }
  Mat(int r, int c, int type) : rows(0), cols(0), data(0), type_(0), parent_(false) { create(r, c, type); }
  Mat(Size s, int type) : rows(0), cols(0), data(0), type_(0), parent_(false) { create(s.height, s.width, type); }
  // filled with a scalar (ED_Lib: edge / segment images); only what a one-channel 8- or 16-bit image needs
  Mat(int r, int c, int type, const Scalar& v) : rows(0), cols(0), data(0), type_(0), parent_(false) { create(r, c, type); fill_(v); }
  Mat(Size s, int type, const Scalar& v) : rows(0), cols(0), data(0), type_(0), parent_(false) { create(s.height, s.width, type); fill_(v); }
  void fill_(const Scalar& v) {
    if (v.val[0] == 0) { for (int y = 0; y < rows; y++) memset(ptr(y), 0, (size_t)cols * elemSize()); return; }
    if (depth() != CV_8U) shim_unsupported("Mat(..., Scalar != 0) for a non-8-bit image");
    for (int y = 0; y < rows; y++) memset(ptr(y), (int)v.val[0], (size_t)cols);
  }
  template <typename T> T& at(const Point_<int>& p) { return at<T>(p.y, p.x); }
  // user-owned memory (not freed)
  Mat(int r, int c, int type, void* d, size_t st = 0) : rows(r), cols(c), data((uchar*)d), type_(type), parent_(false) {
    step = st ? st : (size_t)c * elemSize();
  }
  void create(int r, int c, int type);
  void create(Size s, int type) { create(s.height, s.width, type); }
  void release() { buf_.reset(); data = 0; rows = cols = 0; step = 0; parent_ = false; }
  Mat clone() const;
  void copyTo(Mat& dst) const;
  void copyTo(const _OutputArray& dst) const;
  Mat operator()(const Rect& r) const;
  Mat rowRange(int a, int b) const { return (*this)(Rect(0, a, cols, b - a)); }
  Mat colRange(int a, int b) const { return (*this)(Rect(a, 0, b - a, rows)); }
  int type() const { return type_; }
  int depth() const { return type_ & 7; }
  int channels() const { return 1; }
  size_t elemSize() const { static const int sz[8] = {1, 1, 2, 2, 4, 4, 8, 0}; return sz[type_ & 7]; }
  size_t elemSize1() const { return elemSize(); }
  size_t step1() const { return step.p / elemSize1(); }
  bool empty() const { return data == 0 || rows == 0 || cols == 0; }
  bool isContinuous() const { return rows <= 1 || step.p == (size_t)cols * elemSize(); }
  bool isSubmatrix() const { return parent_; }
  Size size() const { return Size(cols, rows); }
  template <typename T> T& at(int y, int x) { return ((T*)(data + step.p * y))[x]; }
  template <typename T> const T& at(int y, int x) const { return ((const T*)(data + step.p * y))[x]; }
  template <typename T> T& at(int i) { return const_cast<T&>(static_cast<const Mat*>(this)->at<T>(i)); }
  template <typename T> const T& at(int i) const {
    if (isContinuous() || rows == 1) return ((const T*)data)[i];
    if (cols == 1) return *(const T*)(data + step.p * i);
    return ((const T*)(data + step.p * (i / cols)))[i % cols];
  }
  uchar* ptr(int y = 0) { return data + step.p * y; }
  const uchar* ptr(int y = 0) const { return data + step.p * y; }
  template <typename T> T* ptr(int y = 0) { return (T*)(data + step.p * y); }
  template <typename T> const T* ptr(int y = 0) const { return (const T*)(data + step.p * y); }
  static MatExprZeros zeros(int r, int c, int type) { MatExprZeros e = {r, c, type}; return e; }
  void push_back(const Mat& m);
  Mat t() const { shim_unsupported("Mat::t"); }
  Mat col(int) const { shim_unsupported("Mat::col"); }
  Mat row(int) const { shim_unsupported("Mat::row"); }
  Mat reshape(int, int = 0) const { shim_unsupported("Mat::reshape"); }
  Mat cross(const Mat&) const { shim_unsupported("Mat::cross"); }
  Mat& setTo(const Scalar&) { shim_unsupported("Mat::setTo"); }
  void convertTo(Mat& dst, int type) const;

 protected:
  int type_;
  bool parent_;                  // true for a view into a larger buffer (ROI)
  std::shared_ptr<uchar> buf_;   // owner of the allocation (empty for user memory)
};
inline Mat operator*(const Mat&, const Mat&) { shim_unsupported("Mat*Mat"); }
inline Mat operator+(const Mat&, const Mat&) { shim_unsupported("Mat+Mat"); }
inline Mat operator/(const Mat&, double) { shim_unsupported("Mat/scalar"); }
inline Mat operator-(const Mat&, const Mat&) { shim_unsupported("Mat-Mat"); }
inline Mat operator-(const Mat&) { shim_unsupported("-Mat"); }

template <typename T> struct DepthOf;
template <> struct DepthOf<uchar> { enum { value = CV_8U }; };
template <> struct DepthOf<short> { enum { value = CV_16S }; };
template <> struct DepthOf<int> { enum { value = CV_32S }; };
template <> struct DepthOf<float> { enum { value = CV_32F }; };
template <> struct DepthOf<double> { enum { value = CV_64F }; };

template <typename T> class Mat_ : public Mat {
 public:
  Mat_() : Mat() { type_ = DepthOf<T>::value; }
  Mat_(int r, int c) : Mat(r, c, DepthOf<T>::value) {}
  Mat_(const Mat& m) : Mat() { *this = m; }
  Mat_& operator=(const Mat& m) {
    if (m.type() == DepthOf<T>::value) Mat::operator=(m);
    else { Mat tmp; m.convertTo(tmp, DepthOf<T>::value); Mat::operator=(tmp); }
    return *this;
  }
  T* operator[](int y) { return (T*)(data + step.p * y); }
  const T* operator[](int y) const { return (const T*)(data + step.p * y); }
};

// The reference only ever passes cv::Mat through these proxies.
class _InputArray {
 public:
  _InputArray() : m_(0) {}
  _InputArray(const Mat& m) : m_(&m) {}
  bool empty() const { return !m_ || m_->empty(); }
  Mat getMat() const { return m_ ? *m_ : Mat(); }
 protected:
  const Mat* m_;
};
class _OutputArray : public _InputArray {
 public:
  _OutputArray() {}
  _OutputArray(Mat& m) : _InputArray(m) {}
  _OutputArray(const Mat& m) : _InputArray(m) {}   // OpenCV has this overload too (fixed-size output)
  void create(int r, int c, int type) const { mat().create(r, c, type); }
  void create(Size s, int type) const { mat().create(s, type); }
  void release() const { mat().release(); }
  Mat& mat() const { return *const_cast<Mat*>(m_); }
};
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;
inline _InputArray noArray() { return _InputArray(); }

// ---- image primitives, implemented in shim_prims.cpp on top of oracle/cvprim.cpp ----
float fastAtan2(float y, float x);
void resize(InputArray src, OutputArray dst, Size dsize, double fx = 0, double fy = 0, int interpolation = INTER_LINEAR);
void copyMakeBorder(InputArray src, OutputArray dst, int top, int bottom, int left, int right, int borderType,
                    const Scalar& value = Scalar());
void GaussianBlur(InputArray src, OutputArray dst, Size ksize, double sigmaX, double sigmaY = 0,
                  int borderType = BORDER_DEFAULT);
void pyrDown(InputArray src, OutputArray dst, const Size& dstsize = Size(), int borderType = BORDER_DEFAULT);
void Sobel(InputArray src, OutputArray dst, int ddepth, int dx, int dy, int ksize = 3, double scale = 1, double delta = 0,
           int borderType = BORDER_DEFAULT);
void FAST(InputArray image, std::vector<KeyPoint>& keypoints, int threshold, bool nonmaxSuppression = true);
void cvtColor(InputArray src, OutputArray dst, int code, int dstCn = 0);
void Canny(InputArray image, OutputArray edges, double t1, double t2, int apertureSize = 3, bool L2gradient = false);
Mat abs(const Mat& m);
void add(InputArray a, InputArray b, OutputArray c);
double threshold(InputArray src, OutputArray dst, double thresh, double maxval, int type);
void compare(InputArray a, InputArray b, OutputArray c, int cmpop);

struct KeyPointsFilter {
  static void retainBest(std::vector<KeyPoint>&, int) { shim_unsupported("KeyPointsFilter::retainBest"); }
};

class LineIterator {
 public:
  // OpenCV 3.4 imgproc/drawing.cpp: points inside the image are not clipped; 8-connected count = max(|dx|,|dy|) + 1
  LineIterator(const Mat& img, Point pt1, Point pt2, int connectivity = 8, bool leftToRight = false);
  int count;
};

class LineSegmentDetector : public Algorithm {
 public:
  virtual void detect(const Mat& image, std::vector<Vec4f>& lines) = 0;
  virtual ~LineSegmentDetector() {}
};
Ptr<LineSegmentDetector> createLineSegmentDetector(int refine = LSD_REFINE_STD, double scale = 0.8,
                                                   double sigma_scale = 0.6, double quant = 2.0, double ang_th = 22.5,
                                                   double log_eps = 0, double density_th = 0.7, int n_bins = 1024);
}  // namespace cv

// ---------------------------------------------------------------------------------------------------------------------
// Enabling the dead computeDescriptors() call of ORBextractor::operator() WITHOUT editing the source.
// src/ORBextractor.cc:1086-1093 reads
//        s_2 = clock();
//        // computeDescriptors(workingMat, keypoints, desc, pattern);
//        e_2 = clock();
// With SDPL_REF_ENABLE_ORB_DESCRIPTORS defined (only for that translation unit), clock() becomes a macro that first hands
// the names `workingMat, keypoints, desc, pattern` -- in scope exactly there -- to a hook defined in ref_orb_tu.cpp, which
// calls the file-static computeDescriptors() of the unmodified source.  At the other clock() call sites of the file those
// names resolve to the dummies below, the hook's overload for them does nothing.
#ifdef SDPL_REF_ENABLE_ORB_DESCRIPTORS
namespace sdpl_ref_hook {
struct Dummy {};
}
static sdpl_ref_hook::Dummy workingMat, keypoints, desc, pattern;
namespace sdpl_ref_hook {
inline void at_clock(Dummy&, Dummy&, Dummy&, Dummy&) {}
template <typename A, typename B, typename C, typename D> inline void at_clock(A&, B&, C&, D&) {}
void at_clock(cv::Mat& working, std::vector<cv::KeyPoint>& kps, cv::Mat& d, std::vector<cv::Point>& pat);
}
#define clock() (sdpl_ref_hook::at_clock(workingMat, keypoints, desc, pattern), ::clock())
#endif

#endif  // SDPL_CVSHIM_HPP
