// OpenCV stand-in (oracle/refshim): see sdpl_cvshim.hpp; the drawing / display extras that Frame.cc and ED_Lib mention are in sdpl_frameshim.hpp
#include "sdpl_cvshim.hpp"
#include "sdpl_frameshim.hpp"
