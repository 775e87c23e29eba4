// oracle/refshim/ed_stubs.cpp -- ORACLE test infrastructure.  ED_Lib's ED.cpp / EDLines.cpp (compiled unmodified) contain two constructors
// that take an EDColor object (colour edge detection, 3rdparty/line_descriptor/src/ED_Lib/EDColor.cpp); nothing on the path from
// LSDDetectorC::detect_ED reaches them, so the four EDColor members they call only have to link.
#include "EDColor.h"
int EDColor::getWidth() { cv::shim_unsupported("EDColor"); }
int EDColor::getHeight() { cv::shim_unsupported("EDColor"); }
int EDColor::getSegmentNo() { cv::shim_unsupported("EDColor"); }
std::vector<std::vector<cv::Point>> EDColor::getSegments() { cv::shim_unsupported("EDColor"); }
