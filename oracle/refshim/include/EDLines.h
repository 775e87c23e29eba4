// ORACLE test infrastructure: stands in for 3rdparty/line_descriptor/include/EDLines.h (the ED_Lib back-end,
// extractor == 1) so that the unmodified LSDDetector_custom.cpp compiles without ED_Lib.  The EDLines path is reached only
// through LSDDetectorC::detect_ED, which the LSD hot path (extractor == 0, src/Tracking.cc:110) never calls.
#ifndef SDPL_SHIM_EDLINES_H
#define SDPL_SHIM_EDLINES_H
#include "sdpl_cvshim.hpp"
class EDLines {
 public:
  EDLines() {}
  EDLines(cv::Mat) { cv::shim_unsupported("EDLines"); }
  std::vector<cv::Vec4f> getLines() { cv::shim_unsupported("EDLines::getLines"); }
};
#endif
