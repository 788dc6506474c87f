#include <cstdarg>
#include <cstdio>

#include "mof_b200.h"
#include "mof_error.h"

static thread_local char g_err[512] = "";

extern "C" int mof_set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}

extern "C" const char* mof_last_error_string(void) { return g_err; }
extern "C" int mof_version(void) { return 100; }
