// Error plumbing and device checks of the C ABI.
#include <stdarg.h>

#include "common.cuh"

namespace cutdet {

static thread_local char g_error[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
    return code;
}

int cuda_fail(cudaError_t e, const char *what) {
    snprintf(g_error, sizeof(g_error), "CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
    return CUTDET_ECUDA;
}

int sm_count() {
    static int cached = 0;
    if (cached == 0) {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) == cudaSuccess &&
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
            cached = n;
        else
            return 148;
    }
    return cached;
}

}  // namespace cutdet

extern "C" int cutdet_abi_version(void) { return CUTDET_ABI_VERSION; }

extern "C" const char *cutdet_last_error(void) { return cutdet::g_error; }

extern "C" int cutdet_device_check(int *sm_count, int *cc_major, int *cc_minor) {
    int dev = 0;
    CUTDET_CUDA(cudaGetDevice(&dev));
    int major = 0, minor = 0, sms = 0;
    CUTDET_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    CUTDET_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
    CUTDET_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    if (sm_count) *sm_count = sms;
    if (cc_major) *cc_major = major;
    if (cc_minor) *cc_minor = minor;
    if (major != 10)
        return cutdet::fail(CUTDET_EUNSUPPORTED,
                            "libcutdet_b200 is built for sm_100a only; device %d is sm_%d%d", dev, major, minor);
    return CUTDET_OK;
}

extern "C" int cutdet_target_size(int width, int height, int resize, int *new_width, int *new_height) {
    CUTDET_REQUIRE(width > 0 && height > 0 && resize > 0 && new_width && new_height, "cutdet_target_size: bad argument");
    // frameID/data.py:199-202: int(height * (new_width / width)) in Python floats (doubles)
    *new_width = resize;
    *new_height = (int)((double)height * ((double)resize / (double)width));
    return CUTDET_OK;
}
