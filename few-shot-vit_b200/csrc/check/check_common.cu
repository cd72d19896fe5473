// Error plumbing of the TEST-ONLY checker library (tests/native/libsunb200_check.so).
#include "../common.cuh"

#include <stdarg.h>

static thread_local char g_check_err[512] = "";

void sunb_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_check_err, sizeof(g_check_err), fmt, ap);
    va_end(ap);
}

extern "C" const char* sunb_check_last_error(void) { return g_check_err; }
