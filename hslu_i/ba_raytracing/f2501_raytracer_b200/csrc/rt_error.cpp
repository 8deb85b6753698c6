// rt_error.cpp -- rt_last_error() and the version query of the C ABI (include/rt_api.h).
#include <cstring>

#include "rt_internal.h"

static thread_local char g_err[512] = "";

void rt_set_error_v(const char *fmt, va_list ap) { vsnprintf(g_err, sizeof(g_err), fmt, ap); }

void rt_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    rt_set_error_v(fmt, ap);
    va_end(ap);
}

extern "C" {
const char *rt_last_error(void) { return g_err; }
int rt_api_version(void) { return RT_API_VERSION; }
}
