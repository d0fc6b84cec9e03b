#include "../../include/scenesplat_b200.h"
extern "C" const char* ss_version(void) { return "scenesplat_b200 0.1 (sm_100a)"; }
