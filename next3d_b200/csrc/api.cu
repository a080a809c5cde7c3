// Library-level entry points: error reporting and version.
#include "common.cuh"
#include "../../include/next3d_b200.h"
#include <stdarg.h>

static thread_local char g_err[512] = "";

void n3d_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char* n3d_last_error(void) { return g_err; }
extern "C" int n3d_version(void) { return 100; }
