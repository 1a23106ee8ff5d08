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

// FP32-pipe probe: 8 independent FFMA chains per thread, no memory traffic.  bench.py times it with CUDA
// events and reports 2 * 8 * iters * threads FLOP / time as `fp32_tflops` — the measured denominator of the
// search kernels' FP32 roofline (SURVEY.md §8d asks for exactly this next to the nominal 74.4 TFLOP/s).
__global__ void __launch_bounds__(1024) ffma_probe_kernel(int iters, float a, float b, float *out) {
    float x0 = threadIdx.x, x1 = x0 + 1.f, x2 = x0 + 2.f, x3 = x0 + 3.f, x4 = x0 + 4.f, x5 = x0 + 5.f, x6 = x0 + 6.f,
          x7 = x0 + 7.f;
#pragma unroll 4
    for (int i = 0; i < iters; ++i) {
        x0 = __fmaf_rn(x0, a, b); x1 = __fmaf_rn(x1, a, b); x2 = __fmaf_rn(x2, a, b); x3 = __fmaf_rn(x3, a, b);
        x4 = __fmaf_rn(x4, a, b); x5 = __fmaf_rn(x5, a, b); x6 = __fmaf_rn(x6, a, b); x7 = __fmaf_rn(x7, a, b);
    }
    const float s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
    if (s == 12345.678f) out[blockIdx.x * blockDim.x + threadIdx.x] = s;      // keeps the chains alive, never true in practice
}

extern "C" int amc3d_fp32_probe(int iters, int blocks, float *out, double *flop, void *stream) {
    AMC3D_REQUIRE(iters > 0 && blocks > 0 && out != nullptr, AMC3D_EINVAL, "fp32_probe: bad arguments");
    ffma_probe_kernel<<<blocks, 1024, 0, amc3d::as_stream(stream)>>>(iters, 0.999f, 0.001f, out);
    if (flop != nullptr) *flop = 2.0 * 8.0 * (double)iters * 1024.0 * (double)blocks;
    return amc3d::check_launch("fp32_probe");
}

extern "C" int amc3d_version(void) { return AMC3D_VERSION; }
extern "C" const char *amc3d_arch(void) { return "sm_100a"; }
extern "C" const char *amc3d_last_error(void) { return amc3d::g_err; }
