// The remaining pointops exports on packed (n,3) / (n,c) tensors with cumulative i32 offsets
// (SURVEY.md §8f rank 2): ballquery, furthestsampling, interpolation, subtraction, aggregation.
//
// Replaces openpoints/cpp/pointops/src/{ballquery,sampling,interpolation,subtraction,aggregation}/
// *_cuda_kernel.cu.  None of them has a caller in AMContrast3D (they serve Point Transformer style
// models in OpenPoints); they are here so that openpoints/cpp/pointops/functions/pointops.py imports
// against this library alone.  The packed layout is channel-contiguous, so lane = channel is coalesced
// as it stands; arithmetic follows the reference expressions as nvcc contracts them (accumulations
// `out += a * b` are FMA chains in sample order starting from the caller's initial value).
#include "common.cuh"

namespace amc3d {

// knn_grid.cu / fps.cu
int ball_grid_batched(int nb, int n, int m, float radius, int nsample, const float *xyz, const float *new_xyz,
                      int *idx, cudaStream_t st);
int fps_segments(int count, int n, int m, int log2bs, const float *xyz, float *temp, int *idx, int idx_base,
                 cudaStream_t st);

// smallest s with q < new_offset[s] (reference ballquery_bt_idx)
__device__ __forceinline__ int segment_of(int q, const int *__restrict__ new_offset, int nseg) {
    int s = 0;
    while (s < nseg - 1 && q >= __ldg(new_offset + s)) ++s;
    return s;
}

// One warp per query over its segment [start, end): ballot gives the hits in index order, the scan stops at
// nsample hits; slots beyond the hit count repeat the first hit; rows without a hit are not written
// (ballquery_cuda_kernel.cu:27-76).  Indices are global (into xyz).
__global__ void __launch_bounds__(256)
pointops_ballquery_kernel(int m, int nseg, float radius, int nsample, const float *__restrict__ xyz,
                          const float *__restrict__ new_xyz, const int *__restrict__ offset,
                          const int *__restrict__ new_offset, int *__restrict__ idx) {
    const int lane = threadIdx.x & 31;
    const int q = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (q >= m) return;
    const int seg = segment_of(q, new_offset, nseg);
    const int start = seg == 0 ? 0 : __ldg(offset + seg - 1);
    const int end = __ldg(offset + seg);
    const float *qp = new_xyz + 3ll * q;
    int *row = idx + (long long)q * nsample;
    const float r2 = __fmul_rn(radius, radius);
    const float qx = __ldg(qp), qy = __ldg(qp + 1), qz = __ldg(qp + 2);
    const uint32_t lt = (1u << lane) - 1u;
    int cnt = 0, first = -1;
    for (int k0 = start; k0 < end && cnt < nsample; k0 += 32) {
        const int k = k0 + lane;
        bool hit = false;
        if (k < end) {
            const float *p = xyz + 3ll * k;
            // operand order of the reference: new - x (ballquery_cuda_kernel.cu:59)
            hit = dist2_ref(qx - __ldg(p), qy - __ldg(p + 1), qz - __ldg(p + 2)) < r2;
        }
        const uint32_t mask = __ballot_sync(0xffffffffu, hit);
        if (mask) {
            if (first < 0) first = k0 + __ffs(mask) - 1;
            const int pos = cnt + __popc(mask & lt);
            if (hit && pos < nsample) row[pos] = k;
            cnt += __popc(mask);
        }
    }
    if (first >= 0)
        for (int l = min(cnt, nsample) + lane; l < nsample; l += 32) row[l] = first;
}

// output[i,:] = fma(input[idx[i,j],:], weight[i,j], output[i,:]) for j = 0..k-1  (interpolation_cuda_kernel.cu:5-18)
__global__ void __launch_bounds__(256)
interpolation_fwd_kernel(long long total, int c, int k, const float *__restrict__ input, const int *__restrict__ idx,
                         const float *__restrict__ weight, float *__restrict__ output) {
    for (long long e = (long long)blockIdx.x * 256 + threadIdx.x; e < total; e += (long long)gridDim.x * 256) {
        const long long i = e / c;
        const int ch = (int)(e - i * c);
        float acc = output[e];
        for (int j = 0; j < k; ++j)
            acc = __fmaf_rn(__ldg(input + (long long)__ldg(idx + i * k + j) * c + ch), __ldg(weight + i * k + j), acc);
        output[e] = acc;
    }
}

__global__ void __launch_bounds__(256)
interpolation_bwd_kernel(long long total, int c, int k, const float *__restrict__ grad_output,
                         const int *__restrict__ idx, const float *__restrict__ weight, float *__restrict__ grad_input) {
    for (long long e = (long long)blockIdx.x * 256 + threadIdx.x; e < total; e += (long long)gridDim.x * 256) {
        const long long i = e / c;
        const int ch = (int)(e - i * c);
        const float g = __ldg(grad_output + e);
        for (int j = 0; j < k; ++j)
            atomicAdd(grad_input + (long long)__ldg(idx + i * k + j) * c + ch, __fmul_rn(g, __ldg(weight + i * k + j)));
    }
}

// output[i,s,:] = input1[i,:] - input2[idx[i,s],:]  (subtraction_cuda_kernel.cu:5-16)
__global__ void __launch_bounds__(256)
subtraction_fwd_kernel(long long total, int nsample, int c, const float *__restrict__ in1, const float *__restrict__ in2,
                       const int *__restrict__ idx, float *__restrict__ out) {
    for (long long e = (long long)blockIdx.x * 256 + threadIdx.x; e < total; e += (long long)gridDim.x * 256) {
        const int ch = (int)(e % c);
        const long long is = e / c;                // i * nsample + s
        const long long i = is / nsample;
        out[e] = __fsub_rn(__ldg(in1 + i * c + ch), __ldg(in2 + (long long)__ldg(idx + is) * c + ch));
    }
}

__global__ void __launch_bounds__(256)
subtraction_bwd_kernel(long long total, int nsample, int c, const int *__restrict__ idx,
                       const float *__restrict__ grad_out, float *__restrict__ g1, float *__restrict__ g2) {
    for (long long e = (long long)blockIdx.x * 256 + threadIdx.x; e < total; e += (long long)gridDim.x * 256) {
        const int ch = (int)(e % c);
        const long long is = e / c;
        const long long i = is / nsample;
        const float g = __ldg(grad_out + e);
        atomicAdd(g1 + i * c + ch, g);
        atomicAdd(g2 + (long long)__ldg(idx + is) * c + ch, -g);
    }
}

// output[i,ch] = fma(input[idx[i,s],ch] + position[i,s,ch], weight[i,s,ch % w_c], output[i,ch]) over s
// (aggregation_cuda_kernel.cu:5-21)
__global__ void __launch_bounds__(256)
aggregation_fwd_kernel(long long total, int nsample, int c, int w_c, const float *__restrict__ input,
                       const float *__restrict__ position, const float *__restrict__ weight,
                       const int *__restrict__ idx, float *__restrict__ output) {
    for (long long e = (long long)blockIdx.x * 256 + threadIdx.x; e < total; e += (long long)gridDim.x * 256) {
        const long long i = e / c;
        const int ch = (int)(e - i * c);
        const int wc = ch % w_c;
        float acc = output[e];
        for (int s = 0; s < nsample; ++s) {
            const long long is = i * nsample + s;
            const float v = __fadd_rn(__ldg(input + (long long)__ldg(idx + is) * c + ch), __ldg(position + is * c + ch));
            acc = __fmaf_rn(v, __ldg(weight + is * w_c + wc), acc);
        }
        output[e] = acc;
    }
}

__global__ void __launch_bounds__(256)
aggregation_bwd_kernel(long long total, int nsample, int c, int w_c, const float *__restrict__ input,
                       const float *__restrict__ position, const float *__restrict__ weight,
                       const int *__restrict__ idx, const float *__restrict__ grad_output,
                       float *__restrict__ grad_input, float *__restrict__ grad_position,
                       float *__restrict__ grad_weight) {
    for (long long e = (long long)blockIdx.x * 256 + threadIdx.x; e < total; e += (long long)gridDim.x * 256) {
        const long long i = e / c;
        const int ch = (int)(e - i * c);
        const int wc = ch % w_c;
        const float g = __ldg(grad_output + e);
        for (int s = 0; s < nsample; ++s) {
            const long long is = i * nsample + s;
            const long long src = (long long)__ldg(idx + is) * c + ch;
            const float w = __ldg(weight + is * w_c + wc);
            const float gw = __fmul_rn(g, w);
            atomicAdd(grad_input + src, gw);
            grad_position[is * c + ch] = gw;
            atomicAdd(grad_weight + is * w_c + wc, __fmul_rn(g, __fadd_rn(__ldg(input + src), __ldg(position + is * c + ch))));
        }
    }
}

static inline int blocks_for(long long total) {
    return (int)min(div_up_ll(total, 256), (long long)kNumSMs * 32);
}

}  // namespace amc3d

using namespace amc3d;

extern "C" int amc3d_pointops_ballquery(int n, int m, int nseg, float radius, int nsample, const float *xyz,
                                        const float *new_xyz, const int *offset, const int *new_offset, int *idx,
                                        void *stream) {
    AMC3D_REQUIRE(n >= 0 && m >= 0 && nseg >= 1 && nsample >= 1, AMC3D_EINVAL,
                  "pointops_ballquery: bad sizes n=%d m=%d nseg=%d nsample=%d", n, m, nseg, nsample);
    if (m == 0 || n == 0) return 0;
    cudaStream_t st = as_stream(stream);
    if (nseg == 1 && n >= 2048 && nsample <= 128) {       // one segment: the culled search (identical results)
        const int rc = ball_grid_batched(1, n, m, radius, nsample, xyz, new_xyz, idx, st);
        if (rc != 0) {
            set_error("pointops_ballquery (grid): %s", cudaGetErrorString((cudaError_t)rc));
            return rc;
        }
        return check_launch("pointops_ballquery");
    }
    pointops_ballquery_kernel<<<div_up(m, 8), 256, 0, st>>>(m, nseg, radius, nsample, xyz, new_xyz, offset, new_offset, idx);
    return check_launch("pointops_ballquery");
}

// Per-segment FPS.  h_offset / h_new_offset are HOST copies of the cumulative ends (the reference's Python
// wrapper reads them on the host as well, pointops.py:20-23); n_max = the largest segment, which fixes the
// reference's block size and with it the tie order.  tmp (n) = 1e10 on entry; idx (new_offset[-1]) global indices.
extern "C" int amc3d_pointops_furthestsampling(int nseg, int n_max, const float *xyz, const int *h_offset,
                                               const int *h_new_offset, float *tmp, int *idx, void *stream) {
    AMC3D_REQUIRE(nseg >= 0 && n_max >= 0, AMC3D_EINVAL, "pointops_furthestsampling: bad sizes nseg=%d n_max=%d", nseg, n_max);
    AMC3D_REQUIRE(nseg == 0 || (h_offset != nullptr && h_new_offset != nullptr), AMC3D_EINVAL,
                  "pointops_furthestsampling: host offsets are NULL");
    cudaStream_t st = as_stream(stream);
    // reference block size: largest power of two <= n_max, capped at 1024 (cuda_utils.h opt_n_threads)
    int log2bs = 0;
    while ((2 << log2bs) <= n_max && log2bs < 10) ++log2bs;
    int s = 0;
    while (s < nseg) {
        const int n0 = s == 0 ? 0 : h_offset[s - 1], m0 = s == 0 ? 0 : h_new_offset[s - 1];
        const int n = h_offset[s] - n0, m = h_new_offset[s] - m0;
        AMC3D_REQUIRE(n >= 0 && m >= 0 && n <= n_max, AMC3D_EINVAL, "pointops_furthestsampling: segment %d has n=%d m=%d (n_max=%d)", s, n, m, n_max);
        // a run of consecutive segments of the same shape goes out as one batched launch
        int run = 1;
        while (s + run < nseg && h_offset[s + run] - h_offset[s + run - 1] == n &&
               h_new_offset[s + run] - h_new_offset[s + run - 1] == m)
            ++run;
        if (n > 0 && m > 0) {
            const int rc = fps_segments(run, n, m, log2bs, xyz + 3ll * n0, tmp + n0, idx + m0, n0, st);
            if (rc != 0) return rc;
        }
        s += run;
    }
    return check_launch("pointops_furthestsampling");
}

extern "C" int amc3d_pointops_interpolation_forward(int n, int c, int k, const float *input, const int *idx,
                                                    const float *weight, float *output, void *stream) {
    AMC3D_REQUIRE(n >= 0 && c >= 0 && k >= 0, AMC3D_EINVAL, "pointops_interpolation_forward: negative size");
    const long long total = (long long)n * c;
    if (total == 0) return 0;
    interpolation_fwd_kernel<<<blocks_for(total), 256, 0, as_stream(stream)>>>(total, c, k, input, idx, weight, output);
    return check_launch("pointops_interpolation_forward");
}

extern "C" int amc3d_pointops_interpolation_backward(int n, int c, int k, const float *grad_output, const int *idx,
                                                     const float *weight, float *grad_input, void *stream) {
    AMC3D_REQUIRE(n >= 0 && c >= 0 && k >= 0, AMC3D_EINVAL, "pointops_interpolation_backward: negative size");
    const long long total = (long long)n * c;
    if (total == 0) return 0;
    interpolation_bwd_kernel<<<blocks_for(total), 256, 0, as_stream(stream)>>>(total, c, k, grad_output, idx, weight, grad_input);
    return check_launch("pointops_interpolation_backward");
}

extern "C" int amc3d_pointops_subtraction_forward(int n, int nsample, int c, const float *input1, const float *input2,
                                                  const int *idx, float *output, void *stream) {
    AMC3D_REQUIRE(n >= 0 && nsample >= 0 && c >= 0, AMC3D_EINVAL, "pointops_subtraction_forward: negative size");
    const long long total = (long long)n * nsample * c;
    if (total == 0) return 0;
    subtraction_fwd_kernel<<<blocks_for(total), 256, 0, as_stream(stream)>>>(total, nsample, c, input1, input2, idx, output);
    return check_launch("pointops_subtraction_forward");
}

extern "C" int amc3d_pointops_subtraction_backward(int n, int nsample, int c, const int *idx, const float *grad_output,
                                                   float *grad_input1, float *grad_input2, void *stream) {
    AMC3D_REQUIRE(n >= 0 && nsample >= 0 && c >= 0, AMC3D_EINVAL, "pointops_subtraction_backward: negative size");
    const long long total = (long long)n * nsample * c;
    if (total == 0) return 0;
    subtraction_bwd_kernel<<<blocks_for(total), 256, 0, as_stream(stream)>>>(total, nsample, c, idx, grad_output, grad_input1, grad_input2);
    return check_launch("pointops_subtraction_backward");
}

extern "C" int amc3d_pointops_aggregation_forward(int n, int nsample, int c, int w_c, const float *input,
                                                  const float *position, const float *weight, const int *idx,
                                                  float *output, void *stream) {
    AMC3D_REQUIRE(n >= 0 && nsample >= 0 && c >= 0 && w_c >= 1, AMC3D_EINVAL, "pointops_aggregation_forward: bad sizes");
    const long long total = (long long)n * c;
    if (total == 0) return 0;
    aggregation_fwd_kernel<<<blocks_for(total), 256, 0, as_stream(stream)>>>(total, nsample, c, w_c, input, position, weight, idx, output);
    return check_launch("pointops_aggregation_forward");
}

extern "C" int amc3d_pointops_aggregation_backward(int n, int nsample, int c, int w_c, const float *input,
                                                   const float *position, const float *weight, const int *idx,
                                                   const float *grad_output, float *grad_input, float *grad_position,
                                                   float *grad_weight, void *stream) {
    AMC3D_REQUIRE(n >= 0 && nsample >= 0 && c >= 0 && w_c >= 1, AMC3D_EINVAL, "pointops_aggregation_backward: bad sizes");
    const long long total = (long long)n * c;
    if (total == 0) return 0;
    aggregation_bwd_kernel<<<blocks_for(total), 256, 0, as_stream(stream)>>>(total, nsample, c, w_c, input, position, weight, idx,
                                                                             grad_output, grad_input, grad_position, grad_weight);
    return check_launch("pointops_aggregation_backward");
}
