// oracle/refshim: stand-in for the reference's include/Converter.h (g2o / Eigen conversions), which src/Frame.cc includes but does
// not use.  Found before the reference's own header through the include order of oracle/refshim/Makefile.
#ifndef SDPL_REFSHIM_CONVERTER_H
#define SDPL_REFSHIM_CONVERTER_H
#endif
