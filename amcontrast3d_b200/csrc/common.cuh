// Shared helpers for the amc3d sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/amc3d.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "amc3d kernels are written for sm_100a (B200) only"
#endif

namespace amc3d {

void set_error(const char *fmt, ...);

inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }

// Return code for a launch: 0 or the cudaError_t, recording the message.
int check_launch(const char *what);

#define AMC3D_REQUIRE(cond, code, ...)            \
    do {                                          \
        if (!(cond)) {                            \
            amc3d::set_error(__VA_ARGS__);        \
            return (code);                        \
        }                                         \
    } while (0)

constexpr int kNumSMs = 148;  // B200

__host__ __device__ inline int div_up(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ inline long long div_up_ll(long long a, long long b) { return (a + b - 1) / b; }

// Squared distance exactly as nvcc -O2 contracts the reference expression
//   (a-x)*(a-x) + (b-y)*(b-y) + (c-z)*(c-z)
// which is FMUL on the *y* term, then FFMA x, then FFMA z:
//     fma(dz,dz, fma(dx,dx, fl(dy*dy)))
// (nvcc fuses the first product of "a*b + c*d" and keeps the second as the FMUL).  Verified
// in the sm_100 SASS of all four recompiled reference kernels (knnquery, ball_query,
// three_nn, furthest_point_sampling) — see DESIGN.md "distance expression"; SURVEY.md
// App. A has the x and y roles swapped.
// Written with explicit intrinsics so that no compiler flag can change the rounding.
__device__ __forceinline__ float dist2_ref(float dx, float dy, float dz) {
    return __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace amc3d
