#include "common.cuh"
#include "../../include/scenesplat_b200.h"
namespace ss { unsigned long long g_launch_count = 0; }
extern "C" const char* ss_version(void) { return "scenesplat_b200 0.1 (sm_100a)"; }
extern "C" uint64_t ss_launch_count(void) { return ss::g_launch_count; }
