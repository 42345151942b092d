// api.cu - error plumbing and version of the C ABI (include/crw_b200.h).
#include "common.cuh"

#include <stdarg.h>
#include <stdio.h>

namespace crw {

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
        set_error("%s: CUDA error %d (%s)", what, (int)e, cudaGetErrorString(e));
        return CRW_ERR_CUDA;
    }
    return CRW_OK;
}

}  // namespace crw

extern "C" int crw_version(void) { return 100; }   // 0.1.0

extern "C" const char* crw_last_error(void) { return crw::g_err; }
