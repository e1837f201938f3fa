// api.cu -- error plumbing and the small non-compute entry points of the C ABI (include/sd3d.h).
#include <cstdarg>
#include <cstdio>

#include "common.cuh"

namespace sd3d {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_launch(const char* what) {
    const cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) return SD3D_OK;
    set_error("%s: CUDA error: %s", what, cudaGetErrorString(e));
    return SD3D_ERR_CUDA;
}

int num_sms() {
    static int cached = 0;  // benign race: same value from every thread
    if (cached > 0) return cached;
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) {
        cudaGetLastError();
        return 148;
    }
    cached = sms;
    return sms;
}

}  // namespace sd3d

extern "C" int sd3d_version(void) { return SD3D_VERSION; }

extern "C" const char* sd3d_last_error(void) { return sd3d::g_err; }

extern "C" int sd3d_device_sms(void) {
    int dev = 0, sms = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) {
        sd3d::set_error("sd3d_device_sms: %s", cudaGetErrorString(e));
        cudaGetLastError();
        return SD3D_ERR_CUDA;
    }
    return sms;
}
