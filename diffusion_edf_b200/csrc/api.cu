#include "../../include/dedf.h"
extern "C" int dedf_build_arch(void) { return 100; }
