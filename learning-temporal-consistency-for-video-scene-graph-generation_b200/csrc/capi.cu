// Library-level C-ABI: version string and per-thread error message.
#include <cstring>
#include <cstdio>
#include "../../include/b200vsgg.h"
#include "common.cuh"

namespace vsgg {
static thread_local char g_err[256] = "";
int set_error(int code, const char* msg) {
    snprintf(g_err, sizeof(g_err), "%s", msg ? msg : "");
    return code;
}
}  // namespace vsgg

extern "C" const char* b200vsgg_version(void) { return "b200vsgg 0.1 sm_100a"; }
extern "C" const char* b200vsgg_last_error(void) { return vsgg::g_err; }
