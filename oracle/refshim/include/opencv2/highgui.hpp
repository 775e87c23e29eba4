// ORACLE test infrastructure: stands in for the OpenCV header of the same name so that the UNMODIFIED reference
// sources compile here (no OpenCV C++ in this image).  Everything lives in sdpl_cvshim.hpp.
#include "sdpl_cvshim.hpp"
