// Library-level entry points: error string, ABI version, device probe.
#include "common.cuh"
#include "../../include/b200ssl.h"

static thread_local char g_err[512] = "";

void b200_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

B200_API const char* b200_last_error(void) { return g_err; }
B200_API int b200_abi_version(void) { return B200_ABI_VERSION; }

B200_API int b200_device_sm(void) {
    int dev = 0, major = 0, minor = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return -1;
    if (cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev) != cudaSuccess) return -1;
    return major * 10 + minor;
}
