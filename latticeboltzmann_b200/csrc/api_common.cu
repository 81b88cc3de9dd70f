#include "api_common.h"
#include "../../include/lbm_b200.h"

#include <cstdarg>
#include <cstdio>

static thread_local char g_err[512] = "";

extern "C" int lbm_fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

extern "C" const char *lb_last_error(void) { return g_err; }

extern "C" int64_t lb_sizeof_config(void) { return (int64_t)sizeof(lb_config); }
extern "C" int64_t lb_sizeof_export(void) { return (int64_t)sizeof(lb_export); }
