// Error plumbing and library identification for libamc3d (C-ABI in include/amc3d.h).
#include "common.cuh"
#include <stdarg.h>

namespace amc3d {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_launch(const char *what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: CUDA launch failed: %s", what, cudaGetErrorString(e));
        return (int)e;
    }
    return 0;
}

}  // namespace amc3d

extern "C" int amc3d_version(void) { return AMC3D_VERSION; }
extern "C" const char *amc3d_arch(void) { return "sm_100a"; }
extern "C" const char *amc3d_last_error(void) { return amc3d::g_err; }
