// Error reporting + device queries behind the C ABI.
#include "common.cuh"

#include <cstring>

namespace stemseg {

static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t err, const char* what, const char* file, int line) {
    set_error("CUDA error %d (%s) in %s at %s:%d", static_cast<int>(err), cudaGetErrorString(err), what, file, line);
    return STEMSEG_ERR_CUDA;
}

int device_sm_count() {
    static thread_local int cached_dev = -1;
    static thread_local int cached = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 1;
    if (dev != cached_dev) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n < 1) n = 1;
        cached = n;
        cached_dev = dev;
    }
    return cached;
}

int require_sm100() {
    int dev = 0, major = 0;
    SS_CUDA_OK(cudaGetDevice(&dev));
    SS_CUDA_OK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    if (major != 10) {
        set_error("stemseg_b200 is built for sm_100a (B200) only; device %d has compute capability major %d", dev,
                  major);
        return STEMSEG_ERR_UNSUPPORTED;
    }
    return STEMSEG_OK;
}

}  // namespace stemseg

extern "C" const char* stemseg_last_error(void) { return stemseg::g_error; }
extern "C" int32_t stemseg_abi_version(void) { return 22; }
extern "C" int32_t stemseg_check_device(void) { return stemseg::require_sm100(); }
