// AMContrast3D++ masked refinement (RefinementMethod.DualMasks, fusion 'MIN') for sm_100a.
//
// Replaces openpoints/AMContrast3D/MaskedRefine.py:49-119, which materialises an
// [m,K-1,D] neighbour-feature tensor and builds an [m,K-1,D] one-hot by D-2 successive
// torch.cat calls to pick one neighbour row per point.  Here: one argmin per point, then a
// single pass over the (B,D,n) buffer that copies the selected flat D-float chunk.
// Semantics are bug-compatible with the reference: chunks are the raw
// reinterpretation f.view(-1, D) of the CONTIGUOUS (B,D,n) buffer, not a transpose, while
// the mask follows each element's true point index.
#include "common.cuh"

namespace amc3d {

__global__ void __launch_bounds__(256)
refine_select_kernel(int m, int ke, int ld, const int *__restrict__ nbr, const float *__restrict__ a,
                     int *__restrict__ jmin) {
    const int r = blockIdx.x * 256 + threadIdx.x;
    if (r >= m) return;
    const int *row = nbr + (long long)r * ld;
    int best = __ldg(row);
    float bv = __ldg(a + best);
    for (int j = 1; j < ke; ++j) {
        const int nj = __ldg(row + j);
        const float v = __ldg(a + nj);
        if (v < bv) { bv = v; best = nj; }   // strict: first minimum (torch.min over dim)
    }
    jmin[r] = best;
}

// element e of the flat (B,D,n) buffer: point index i = e % n, batch b = e / (D*n);
// chunk row r = e / D, column col = e % D.
__global__ void __launch_bounds__(256)
refine_forward_kernel(long long total, int d, int n, const float *__restrict__ f, const float *__restrict__ a,
                      const int *__restrict__ jmin, float thr, float thr_max, float gamma, float one_m_gamma,
                      float *__restrict__ out, int *__restrict__ update_count) {
    int local = 0;
    for (long long e = (long long)blockIdx.x * 256 + threadIdx.x; e < total; e += (long long)gridDim.x * 256) {
        const long long bd = e / n;            // b*D + dch
        const int i = (int)(e - bd * n);
        const long long b = bd / d;
        const int dch = (int)(bd - b * d);
        const float av = __ldg(a + b * n + i);
        const bool mask = (av <= thr_max) && (av >= thr);
        const float fv = __ldg(f + e);
        float fnew = fv;
        if (mask) {
            const long long r = e / d;
            const int col = (int)(e - r * d);
            fnew = __ldg(f + (long long)__ldg(jmin + r) * d + col);
            if (dch == 0) ++local;
        }
        out[e] = __fadd_rn(__fmul_rn(gamma, fnew), __fmul_rn(one_m_gamma, fv));
    }
    local = __reduce_add_sync(0xffffffffu, local);
    if ((threadIdx.x & 31) == 0 && local) atomicAdd(update_count, local);
}

// direct part: grad_f[e] = (1-gamma)*go[e] + (mask ? 0 : gamma*go[e])
__global__ void __launch_bounds__(256)
refine_backward_direct_kernel(long long total, int d, int n, const float *__restrict__ go,
                              const float *__restrict__ a, float thr, float thr_max, float gamma,
                              float one_m_gamma, float *__restrict__ grad_f) {
    for (long long e = (long long)blockIdx.x * 256 + threadIdx.x; e < total; e += (long long)gridDim.x * 256) {
        const long long bd = e / n;
        const int i = (int)(e - bd * n);
        const long long b = bd / d;
        const float av = __ldg(a + b * n + i);
        const bool mask = (av <= thr_max) && (av >= thr);
        const float g = __ldg(go + e);
        grad_f[e] = one_m_gamma * g + (mask ? 0.f : gamma * g);
    }
}

// cross part: grad_f[jmin[r]*D + col] += gamma*go[e] for masked elements
__global__ void __launch_bounds__(256)
refine_backward_cross_kernel(long long total, int d, int n, const float *__restrict__ go,
                             const float *__restrict__ a, const int *__restrict__ jmin, float thr,
                             float thr_max, float gamma, float *__restrict__ grad_f) {
    for (long long e = (long long)blockIdx.x * 256 + threadIdx.x; e < total; e += (long long)gridDim.x * 256) {
        const long long bd = e / n;
        const int i = (int)(e - bd * n);
        const long long b = bd / d;
        const float av = __ldg(a + b * n + i);
        if ((av <= thr_max) && (av >= thr)) {
            const long long r = e / d;
            const int col = (int)(e - r * d);
            atomicAdd(grad_f + (long long)__ldg(jmin + r) * d + col, gamma * __ldg(go + e));
        }
    }
}

}  // namespace amc3d

using namespace amc3d;

extern "C" int amc3d_refine_select(int m, int ke, int ld, const int *nbr, const float *a, int *jmin, void *stream) {
    AMC3D_REQUIRE(m >= 0 && ke >= 1 && ld >= ke, AMC3D_EINVAL, "refine_select: bad sizes m=%d ke=%d ld=%d", m, ke, ld);
    if (m == 0) return 0;
    refine_select_kernel<<<div_up(m, 256), 256, 0, as_stream(stream)>>>(m, ke, ld, nbr, a, jmin);
    return check_launch("refine_select");
}

static inline int grid_for(long long total) {
    return (int)min(div_up_ll(total, 256), (long long)kNumSMs * 16);
}

extern "C" int amc3d_refine_forward(int b, int d, int n, const float *f, const float *a, const int *jmin,
                                    float thr, float thr_max, float gamma, float *out, int *update_count,
                                    void *stream) {
    AMC3D_REQUIRE(b >= 0 && d >= 1 && n >= 0, AMC3D_EINVAL, "refine_forward: bad sizes b=%d d=%d n=%d", b, d, n);
    const long long total = (long long)b * d * n;
    if (total == 0) return 0;
    // (1 - gamma) is formed in Python double precision and applied as a float32 scalar
    const float omg = (float)(1.0 - (double)gamma);
    refine_forward_kernel<<<grid_for(total), 256, 0, as_stream(stream)>>>(total, d, n, f, a, jmin, thr, thr_max,
                                                                           gamma, omg, out, update_count);
    return check_launch("refine_forward");
}

extern "C" int amc3d_refine_backward(int b, int d, int n, const float *grad_out, const float *a,
                                     const int *jmin, float thr, float thr_max, float gamma, float *grad_f,
                                     void *stream) {
    AMC3D_REQUIRE(b >= 0 && d >= 1 && n >= 0, AMC3D_EINVAL, "refine_backward: bad sizes b=%d d=%d n=%d", b, d, n);
    const long long total = (long long)b * d * n;
    if (total == 0) return 0;
    const float omg = (float)(1.0 - (double)gamma);
    cudaStream_t st = as_stream(stream);
    refine_backward_direct_kernel<<<grid_for(total), 256, 0, st>>>(total, d, n, grad_out, a, thr, thr_max, gamma,
                                                                    omg, grad_f);
    refine_backward_cross_kernel<<<grid_for(total), 256, 0, st>>>(total, d, n, grad_out, a, jmin, thr, thr_max,
                                                                   gamma, grad_f);
    return check_launch("refine_backward");
}
