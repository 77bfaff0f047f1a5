// pb_api.cu — version / error plumbing of libplastid_b200.
#include "pb_common.cuh"
#include <stdarg.h>
#include <stdio.h>

static thread_local char g_err[512] = "";

void pb_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char *pb_version(void) { return "plastid_b200 0.1 (sm_100a)"; }
extern "C" const char *pb_last_error(void) { return g_err; }

extern "C" int pb_device_count(void)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        pb_set_error("cudaGetDeviceCount failed: %s", cudaGetErrorString(e));
        cudaGetLastError();
        return 0;
    }
    return n;
}
