// oracle/refshim/ref_frame_tu.cpp -- ORACLE test infrastructure.  Compiles the UNMODIFIED /root/reference/src/Frame.cc.
#include "sdpl_cvshim.hpp"
#include "sdpl_frameshim.hpp"
#include "Frame.cc"
