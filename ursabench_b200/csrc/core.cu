// Error plumbing, device info and ABI version of libursa_b200.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace ursa {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char *what) {
    set_error("%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
    return URSA_ERR_CUDA;
}

int sm_count() {
    static thread_local int cached_dev = -1, cached = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != cached_dev) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached = n;
        cached_dev = dev;
    }
    return cached;
}

}  // namespace ursa

extern "C" int ursa_abi_version(void) { return URSA_ABI_VERSION; }

extern "C" const char *ursa_last_error(void) { return ursa::g_err; }

extern "C" int ursa_device_info(int *sm, int *major, int *minor) {
    int dev = 0;
    URSA_CUDA(cudaGetDevice(&dev));
    if (sm) URSA_CUDA(cudaDeviceGetAttribute(sm, cudaDevAttrMultiProcessorCount, dev));
    if (major) URSA_CUDA(cudaDeviceGetAttribute(major, cudaDevAttrComputeCapabilityMajor, dev));
    if (minor) URSA_CUDA(cudaDeviceGetAttribute(minor, cudaDevAttrComputeCapabilityMinor, dev));
    return URSA_OK;
}
