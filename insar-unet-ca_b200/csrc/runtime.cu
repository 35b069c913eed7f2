// runtime.cu — error reporting and device queries shared by every entry point of libunetca_b200.so.
#include "common.cuh"
#include <stdarg.h>
#include <string.h>

namespace unetca {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return UNETCA_ERR_CUDA;
    }
    return UNETCA_OK;
}

// Ordinal of the calling thread's current device, clamped to the size of the per-device caches below.  Everything the
// library remembers (SM count, "kernel attributes set" flags) is a property of a DEVICE: a process that drives several
// GPUs (one model on cuda:0, another on cuda:1) must not reuse what it learnt on the first one.
int device_slot() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0) dev = 0;
    return dev < kMaxDevices ? dev : kMaxDevices - 1;
}

int num_sms() {
    static int n[kMaxDevices] = {};
    const int d = device_slot();
    if (n[d] == 0) {
        if (cudaDeviceGetAttribute(&n[d], cudaDevAttrMultiProcessorCount, d) != cudaSuccess || n[d] <= 0) n[d] = 148;
    }
    return n[d];
}

}  // namespace unetca

extern "C" {
const char* unetca_last_error(void) { return unetca::g_err; }
int unetca_abi_version(void) { return 1; }
int unetca_num_sms(void) { return unetca::num_sms(); }
}
