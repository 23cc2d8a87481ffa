// liblina_b200: error reporting + ABI version.  See include/lina_b200.h.
#include <stdarg.h>
#include "common.cuh"

static thread_local char g_err[512] = "";

void lina_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" int lina_abi_version(void) { return 1; }
extern "C" const char *lina_last_error_string(void) { return g_err; }
