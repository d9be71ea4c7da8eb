// Error plumbing of the C ABI (include/occnerf_b200.h).
#include <stdarg.h>
#include "common.cuh"

static thread_local char g_err[512] = "";

void occnerf_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char *occnerf_last_error(void) { return g_err; }
extern "C" int occnerf_abi_version(void) { return 3; }   // 2: cta_pair argument of the canonical MLP entry points, packed warp kernels; 3: of the non-rigid ones
