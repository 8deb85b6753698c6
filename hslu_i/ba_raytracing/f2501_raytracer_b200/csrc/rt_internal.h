// rt_internal.h -- helpers shared by the translation units of librt_b200.so (not installed).
#pragma once
#include <cstdarg>
#include <cstdio>

#include "rt_api.h"

// thread-local message behind rt_last_error()
void rt_set_error(const char *fmt, ...) __attribute__((format(printf, 1, 2)));
void rt_set_error_v(const char *fmt, va_list ap);

// sets the message and returns `code` (so failures read `return rt_fail(RT_ERR_..., "...")`)
inline int rt_fail(int code, const char *fmt, ...) __attribute__((format(printf, 2, 3)));
inline int rt_fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    rt_set_error_v(fmt, ap);
    va_end(ap);
    return code;
}
