// oracle/refshim/include/sdpl_frameshim.hpp -- ORACLE test infrastructure, NOT the product.  What src/Frame.cc needs from OpenCV on
// top of sdpl_cvshim.hpp.  Only the Frame CONSTRUCTOR, AssignFeaturesToGrid / PosInGrid and GetFeaturesInArea are ever executed (through
// oracle/refshim/ref_frame_api.cpp): everything they use is implemented here with OpenCV's semantics; what only the other member
// functions of the file use (matrix algebra, RNG, undistortPoints, drawing) is declared so that the file compiles and aborts when called.
#ifndef SDPL_FRAMESHIM_HPP
#define SDPL_FRAMESHIM_HPP
#include "sdpl_cvshim.hpp"
#include <string>
#include <vector>

namespace cv {
typedef Vec<float, 2> Vec2f;
template <typename T> struct Point3_ {
  T x, y, z;
  Point3_() : x(0), y(0), z(0) {}
  Point3_(T a, T b, T c) : x(a), y(b), z(c) {}
};
typedef Point3_<float> Point3f;
typedef Point3_<double> Point3d;

// display / drawing: no-ops (the constructor draws the detected lines into a clone of the image and shows it, Frame.cc:399-408)
inline void line(Mat&, Point2f, Point2f, const Scalar&, int = 1, int = 8, int = 0) {}
inline void line(Mat&, Point_<double>, Point_<double>, const Scalar&, int = 1, int = 8, int = 0) {}   // EDLines::getLineImage / drawOnImage (never called)
enum { LINE_AA = 16, COLOR_GRAY2BGR = 8 };
inline void imshow(const std::string&, const Mat&) {}
inline int waitKey(int = 0) { return -1; }
struct DrawMatchesFlags { enum { DEFAULT = 0 }; };
inline void drawKeypoints(const Mat&, const std::vector<KeyPoint>&, Mat&, const Scalar& = Scalar::all(-1), int = 0) {}

// used by member functions that are never executed here
inline void undistortPoints(const Mat&, Mat&, const Mat&, const Mat&, const Mat& = Mat(), const Mat& = Mat()) { shim_unsupported("undistortPoints"); }
inline double norm(const Mat&) { shim_unsupported("norm"); }
class RNG {
 public:
  RNG() {}
  explicit RNG(unsigned long long) {}
  double gaussian(double) { shim_unsupported("RNG::gaussian"); }
  int uniform(int, int) { shim_unsupported("RNG::uniform"); }
  float uniform(float, float) { shim_unsupported("RNG::uniform"); }
  double uniform(double, double) { shim_unsupported("RNG::uniform"); }
};
inline RNG& theRNG() { static RNG r; return r; }
// cv::Mat_<float>(3, 1) << a, b, c  (comma initialiser): only in functions that are never executed here
template <typename T> struct MatCommaInit {
  template <typename U> MatCommaInit& operator,(U) { return *this; }
  operator Mat() const { shim_unsupported("Mat_ comma initialiser"); }
  operator Mat_<T>() const { shim_unsupported("Mat_ comma initialiser"); }
};
template <typename T, typename U> inline MatCommaInit<T> operator<<(const Mat_<T>&, U) { return MatCommaInit<T>(); }
inline void convertScaleAbs(const Mat&, Mat&, double = 1, double = 0) { shim_unsupported("convertScaleAbs"); }   // ED::getGradImage (never called)
template <typename T, typename P> inline int partition(const std::vector<T>&, std::vector<int>&, P) { shim_unsupported("partition"); }
}  // namespace cv
#endif
