// OpenCV stand-in (oracle/refshim): see sdpl_cvshim.hpp
#include "sdpl_cvshim.hpp"
