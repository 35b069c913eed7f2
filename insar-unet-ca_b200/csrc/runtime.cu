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

int num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess ||
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
            n = 148;
    }
    return n;
}

}  // namespace unetca

extern "C" {
const char* unetca_last_error(void) { return unetca::g_err; }
int unetca_abi_version(void) { return 1; }
int unetca_num_sms(void) { return unetca::num_sms(); }
}
