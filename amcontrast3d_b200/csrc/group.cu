// grouping_operation (gather + scatter-add backward), gather_operation, three_interpolate
// and the packed (n,c) pointops grouping for sm_100a.
//
// Replaces openpoints/cpp/pointnet2_batch/src/group_points_gpu.cu:14-92,
// sampling_gpu.cu:15-90, interpolate_gpu.cu:84-173 and
// openpoints/cpp/pointops/src/grouping/grouping_cuda_kernel.cu:5-25.
//
// The reference gathers 4-byte elements at random positions of a (B,C,N) row, one launch
// thread per output element, re-reading idx once per channel (grid.y = C), and scatters the
// backward with one scalar atomicAdd per element into random addresses.
//
// B200 design (HBM-bound copy, DESIGN.md "grouping"):
//   * the source is first transposed into a channel-contiguous (B,N,C) workspace (a 32x32
//     shared-memory tile transpose; 5% of the output volume), which stays L2-resident
//     (<= 25 MB per call against 126 MB of L2);
//   * a warp owns 32 channels x 32 positions: lane = channel.  Every neighbour index then
//     turns into ONE fully coalesced 128-byte read of the workspace row, and every lane
//     writes 32 contiguous bytes (two STG.128, one full sector) of its (b,c) output row;
//     idx is read once per 32 channels instead of once per channel;
//   * the backward mirrors it: lanes read full 32-byte sectors of grad_out, and the
//     scatter-add becomes a warp-wide RED.ADD.F32 onto 128 contiguous bytes of the
//     L2-resident (B,N,C) workspace (one L2 atomic transaction per 32 elements instead of
//     32), followed by a transpose-accumulate into the caller's (B,C,N) grad buffer.
//   * without a workspace (NULL) or for C < 8 (the xyz grouping, C = 3) a direct kernel
//     keeps idx in registers across channels.
#include "common.cuh"

namespace amc3d {

// ---------------------------------------------------------------------------------------
// (B,R,Cc) -> (B,Cc,R) tile transpose; ACC adds into dst instead of overwriting
// ---------------------------------------------------------------------------------------
template <bool ACC>
__global__ void __launch_bounds__(256)
transpose_kernel(int rows, int cols, const float *__restrict__ src, float *__restrict__ dst) {
    // src (B, rows, cols) -> dst (B, cols, rows)
    __shared__ float t[32][33];
    const int b = blockIdx.z;
    src += (long long)b * rows * cols;
    dst += (long long)b * rows * cols;
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
#pragma unroll
    for (int i = 0; i < 32; i += 8) {
        const int r = r0 + ty + i, c = c0 + tx;
        if (r < rows && c < cols) t[ty + i][tx] = __ldg(src + (long long)r * cols + c);
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 32; i += 8) {
        const int c = c0 + ty + i, r = r0 + tx;
        if (r < rows && c < cols) {
            float *o = dst + (long long)c * rows + r;
            if (ACC) *o += t[tx][ty + i];
            else *o = t[tx][ty + i];
        }
    }
}

template <bool ACC>
static void launch_transpose(int b, int rows, int cols, const float *src, float *dst, cudaStream_t st) {
    dim3 grid(div_up(cols, 32), div_up(rows, 32), b);
    transpose_kernel<ACC><<<grid, 256, 0, st>>>(rows, cols, src, dst);
}

// ---------------------------------------------------------------------------------------
// channel-last gather / scatter: warp = 32 channels x 32 positions
// ---------------------------------------------------------------------------------------
constexpr int GRP_WARPS = 8;  // warps per CTA -> 256 consecutive positions of one channel group

// srcT (B,N,C) channel-contiguous, idx (B,P), out (B,C,P).  Requires P % 8 == 0.
__global__ void __launch_bounds__(GRP_WARPS * 32)
group_fwd_cl_kernel(int C, int N, int P, const float *__restrict__ srcT, const int *__restrict__ idx,
                    float *__restrict__ out) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.z;
    const int c = blockIdx.y * 32 + lane;
    const bool cok = c < C;
    const int cc = cok ? c : C - 1;
    const long long pbase = ((long long)blockIdx.x * GRP_WARPS + warp) * 32;
    const float *src = srcT + (long long)b * N * C + cc;
    const int *ip = idx + (long long)b * P;
    float *orow = out + ((long long)b * C + cc) * P;
#pragma unroll
    for (int o = 0; o < 4; ++o) {
        const long long p0 = pbase + o * 8;
        if (p0 >= P) break;
        const int4 ia = __ldg(reinterpret_cast<const int4 *>(ip + p0));
        const int4 ib = __ldg(reinterpret_cast<const int4 *>(ip + p0 + 4));
        float4 va, vb;
        va.x = __ldg(src + (long long)ia.x * C);
        va.y = __ldg(src + (long long)ia.y * C);
        va.z = __ldg(src + (long long)ia.z * C);
        va.w = __ldg(src + (long long)ia.w * C);
        vb.x = __ldg(src + (long long)ib.x * C);
        vb.y = __ldg(src + (long long)ib.y * C);
        vb.z = __ldg(src + (long long)ib.z * C);
        vb.w = __ldg(src + (long long)ib.w * C);
        if (cok) {
            __stcs(reinterpret_cast<float4 *>(orow + p0), va);
            __stcs(reinterpret_cast<float4 *>(orow + p0 + 4), vb);
        }
    }
}

// grad_out (B,C,P), idx (B,P) -> accT (B,N,C) += ; requires P % 8 == 0
__global__ void __launch_bounds__(GRP_WARPS * 32)
group_bwd_cl_kernel(int C, int N, int P, const float *__restrict__ grad_out, const int *__restrict__ idx,
                    float *__restrict__ accT) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.z;
    const int c = blockIdx.y * 32 + lane;
    if (c >= C) return;
    const long long pbase = ((long long)blockIdx.x * GRP_WARPS + warp) * 32;
    float *acc = accT + (long long)b * N * C + c;
    const int *ip = idx + (long long)b * P;
    const float *grow = grad_out + ((long long)b * C + c) * P;
#pragma unroll
    for (int o = 0; o < 4; ++o) {
        const long long p0 = pbase + o * 8;
        if (p0 >= P) break;
        const int4 ia = __ldg(reinterpret_cast<const int4 *>(ip + p0));
        const int4 ib = __ldg(reinterpret_cast<const int4 *>(ip + p0 + 4));
        const float4 ga = __ldcs(reinterpret_cast<const float4 *>(grow + p0));
        const float4 gb = __ldcs(reinterpret_cast<const float4 *>(grow + p0 + 4));
        atomicAdd(acc + (long long)ia.x * C, ga.x);
        atomicAdd(acc + (long long)ia.y * C, ga.y);
        atomicAdd(acc + (long long)ia.z * C, ga.z);
        atomicAdd(acc + (long long)ia.w * C, ga.w);
        atomicAdd(acc + (long long)ib.x * C, gb.x);
        atomicAdd(acc + (long long)ib.y * C, gb.y);
        atomicAdd(acc + (long long)ib.z * C, gb.z);
        atomicAdd(acc + (long long)ib.w * C, gb.w);
    }
}

// ---------------------------------------------------------------------------------------
// direct kernels (no workspace / tiny C / P not a multiple of 8): one thread per position,
// idx held in a register across the channel loop
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
group_fwd_direct_kernel(int C, int N, int P, int cchunk, const float *__restrict__ points,
                        const int *__restrict__ idx, float *__restrict__ out) {
    const int b = blockIdx.z;
    const long long p = (long long)blockIdx.x * 256 + threadIdx.x;
    if (p >= P) return;
    const int c0 = blockIdx.y * cchunk, c1 = min(C, c0 + cchunk);
    const int i = __ldg(idx + (long long)b * P + p);
    const float *src = points + ((long long)b * C + c0) * N + i;
    float *dst = out + ((long long)b * C + c0) * P + p;
    for (int c = c0; c < c1; ++c, src += N, dst += P) __stcs(dst, __ldg(src));
}

__global__ void __launch_bounds__(256)
group_bwd_direct_kernel(int C, int N, int P, int cchunk, const float *__restrict__ grad_out,
                        const int *__restrict__ idx, float *__restrict__ grad_points) {
    const int b = blockIdx.z;
    const long long p = (long long)blockIdx.x * 256 + threadIdx.x;
    if (p >= P) return;
    const int c0 = blockIdx.y * cchunk, c1 = min(C, c0 + cchunk);
    const int i = __ldg(idx + (long long)b * P + p);
    float *dst = grad_points + ((long long)b * C + c0) * N + i;
    const float *src = grad_out + ((long long)b * C + c0) * P + p;
    for (int c = c0; c < c1; ++c, src += P, dst += N) atomicAdd(dst, __ldcs(src));
}

// three_interpolate: out[b,c,i] = w0*f[i0] + w1*f[i1] + w2*f[i2], contracted exactly like nvcc
// contracts the reference expression (interpolate_gpu.cu:103; verified in its sm_100 SASS):
//     fma(w2,f2, fma(w0,f0, fl(w1*f1)))
__global__ void __launch_bounds__(256)
three_interp_kernel(int C, int M, int N, int cchunk, const float *__restrict__ points,
                    const int *__restrict__ idx, const float *__restrict__ weight,
                    float *__restrict__ out) {
    const int b = blockIdx.z;
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= N) return;
    const int c0 = blockIdx.y * cchunk, c1 = min(C, c0 + cchunk);
    const long long o = 3ll * ((long long)b * N + i);
    const int i0 = __ldg(idx + o), i1 = __ldg(idx + o + 1), i2 = __ldg(idx + o + 2);
    const float w0 = __ldg(weight + o), w1 = __ldg(weight + o + 1), w2 = __ldg(weight + o + 2);
    const float *src = points + ((long long)b * C + c0) * M;
    float *dst = out + ((long long)b * C + c0) * N + i;
    for (int c = c0; c < c1; ++c, src += M, dst += N)
        *dst = __fmaf_rn(w2, __ldg(src + i2), __fmaf_rn(w0, __ldg(src + i0), __fmul_rn(w1, __ldg(src + i1))));
}

__global__ void __launch_bounds__(256)
three_interp_grad_kernel(int C, int N, int M, int cchunk, const float *__restrict__ grad_out,
                         const int *__restrict__ idx, const float *__restrict__ weight,
                         float *__restrict__ grad_points) {
    const int b = blockIdx.z;
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= N) return;
    const int c0 = blockIdx.y * cchunk, c1 = min(C, c0 + cchunk);
    const long long o = 3ll * ((long long)b * N + i);
    const int i0 = __ldg(idx + o), i1 = __ldg(idx + o + 1), i2 = __ldg(idx + o + 2);
    const float w0 = __ldg(weight + o), w1 = __ldg(weight + o + 1), w2 = __ldg(weight + o + 2);
    float *dst = grad_points + ((long long)b * C + c0) * M;
    const float *src = grad_out + ((long long)b * C + c0) * N + i;
    for (int c = c0; c < c1; ++c, src += N, dst += M) {
        const float g = __ldg(src);
        atomicAdd(dst + i0, g * w0);
        atomicAdd(dst + i1, g * w1);
        atomicAdd(dst + i2, g * w2);
    }
}

// packed (n,c) grouping: out[r,:] = in[idx[r],:] for r over m*nsample rows; float4 when c%4==0
template <int VEC>
__global__ void __launch_bounds__(256)
rows_gather_kernel(long long rows, int c, const float *__restrict__ in, const int *__restrict__ idx,
                   float *__restrict__ out) {
    const int cv = c / VEC;
    const long long total = rows * cv;
    for (long long e = (long long)blockIdx.x * 256 + threadIdx.x; e < total; e += (long long)gridDim.x * 256) {
        const long long r = e / cv;
        const int j = (int)(e - r * cv);
        const long long s = __ldg(idx + r);
        if (VEC == 4)
            __stcs(reinterpret_cast<float4 *>(out) + e, __ldg(reinterpret_cast<const float4 *>(in + s * c) + j));
        else
            out[e] = __ldg(in + s * c + j);
    }
}

template <int VEC>
__global__ void __launch_bounds__(256)
rows_scatter_add_kernel(long long rows, int c, const float *__restrict__ grad_out,
                        const int *__restrict__ idx, float *__restrict__ grad_in) {
    const int cv = c / VEC;
    const long long total = rows * cv;
    for (long long e = (long long)blockIdx.x * 256 + threadIdx.x; e < total; e += (long long)gridDim.x * 256) {
        const long long r = e / cv;
        const int j = (int)(e - r * cv);
        const long long s = __ldg(idx + r);
        if (VEC == 4) {
            const float4 g = __ldcs(reinterpret_cast<const float4 *>(grad_out) + e);
            float *p = grad_in + s * c + j * 4;
            asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(g.x), "f"(g.y),
                         "f"(g.z), "f"(g.w)
                         : "memory");
        } else {
            atomicAdd(grad_in + s * c + j, grad_out[e]);
        }
    }
}

static inline int pick_cchunk(int C, long long blocks_xy) {
    // enough CTAs to fill the machine twice, but keep idx reuse across channels
    int chunks = 1;
    while (blocks_xy * chunks < 2 * kNumSMs && chunks < C) chunks *= 2;
    return div_up(C, chunks);
}

}  // namespace amc3d

using namespace amc3d;

static int group_common(bool fwd, int b, int c, int n, long long P, const float *src, const int *idx,
                        float *dst, float *workspace, cudaStream_t st, const char *what) {
    if (b == 0 || c == 0 || P == 0) return 0;
    AMC3D_REQUIRE(b <= 65535, AMC3D_ELIMIT, "%s: batch %d > 65535", what, b);
    AMC3D_REQUIRE(P < (1ll << 31), AMC3D_ELIMIT, "%s: npoints*nsample too large", what);
    const bool aligned = (P % 8 == 0) && ((reinterpret_cast<uintptr_t>(idx) & 15) == 0) &&
                         ((reinterpret_cast<uintptr_t>(fwd ? dst : const_cast<float *>(src)) & 15) == 0);
    if (workspace != nullptr && c >= 8 && aligned && n > 0) {
        dim3 grid((unsigned)div_up_ll(P, GRP_WARPS * 32), div_up(c, 32), b);
        if (fwd) {
            launch_transpose<false>(b, c, n, src, workspace, st);  // (B,C,N) -> (B,N,C)
            group_fwd_cl_kernel<<<grid, GRP_WARPS * 32, 0, st>>>(c, n, (int)P, workspace, idx, dst);
        } else {
            cudaMemsetAsync(workspace, 0, sizeof(float) * (size_t)b * n * c, st);
            group_bwd_cl_kernel<<<grid, GRP_WARPS * 32, 0, st>>>(c, n, (int)P, src, idx, workspace);
            launch_transpose<true>(b, n, c, workspace, dst, st);   // (B,N,C) -> += (B,C,N)
        }
    } else {
        const long long bx = div_up_ll(P, 256);
        const int cchunk = pick_cchunk(c, bx * b);
        dim3 grid((unsigned)bx, div_up(c, cchunk), b);
        if (fwd)
            group_fwd_direct_kernel<<<grid, 256, 0, st>>>(c, n, (int)P, cchunk, src, idx, dst);
        else
            group_bwd_direct_kernel<<<grid, 256, 0, st>>>(c, n, (int)P, cchunk, src, idx, dst);
    }
    return check_launch(what);
}

extern "C" int amc3d_group_points_ws(int b, int c, int n, int npoints, int nsample, const float *points,
                                     const int *idx, float *out, float *workspace, void *stream) {
    AMC3D_REQUIRE(b >= 0 && c >= 0 && n >= 0 && npoints >= 0 && nsample >= 0, AMC3D_EINVAL,
                  "group_points: negative size");
    return group_common(true, b, c, n, (long long)npoints * nsample, points, idx, out, workspace,
                        as_stream(stream), "group_points");
}
extern "C" int amc3d_group_points(int b, int c, int n, int npoints, int nsample, const float *points,
                                  const int *idx, float *out, void *stream) {
    return amc3d_group_points_ws(b, c, n, npoints, nsample, points, idx, out, nullptr, stream);
}

extern "C" int amc3d_group_points_grad_ws(int b, int c, int n, int npoints, int nsample,
                                          const float *grad_out, const int *idx, float *grad_points,
                                          float *workspace, void *stream) {
    AMC3D_REQUIRE(b >= 0 && c >= 0 && n >= 0 && npoints >= 0 && nsample >= 0, AMC3D_EINVAL,
                  "group_points_grad: negative size");
    return group_common(false, b, c, n, (long long)npoints * nsample, grad_out, idx, grad_points,
                        workspace, as_stream(stream), "group_points_grad");
}
extern "C" int amc3d_group_points_grad(int b, int c, int n, int npoints, int nsample,
                                       const float *grad_out, const int *idx, float *grad_points,
                                       void *stream) {
    return amc3d_group_points_grad_ws(b, c, n, npoints, nsample, grad_out, idx, grad_points, nullptr, stream);
}

// gather_points = group_points with nsample = 1
extern "C" int amc3d_gather_points(int b, int c, int n, int npoints, const float *points,
                                   const int *idx, float *out, void *stream) {
    AMC3D_REQUIRE(b >= 0 && c >= 0 && n >= 0 && npoints >= 0, AMC3D_EINVAL, "gather_points: negative size");
    return group_common(true, b, c, n, npoints, points, idx, out, nullptr, as_stream(stream), "gather_points");
}
extern "C" int amc3d_gather_points_grad(int b, int c, int n, int npoints, const float *grad_out,
                                        const int *idx, float *grad_points, void *stream) {
    AMC3D_REQUIRE(b >= 0 && c >= 0 && n >= 0 && npoints >= 0, AMC3D_EINVAL, "gather_points_grad: negative size");
    return group_common(false, b, c, n, npoints, grad_out, idx, grad_points, nullptr, as_stream(stream),
                        "gather_points_grad");
}

extern "C" int amc3d_three_interpolate(int b, int c, int m, int n, const float *points, const int *idx,
                                       const float *weight, float *out, void *stream) {
    AMC3D_REQUIRE(b >= 0 && c >= 0 && m >= 0 && n >= 0, AMC3D_EINVAL, "three_interpolate: negative size");
    AMC3D_REQUIRE(b <= 65535, AMC3D_ELIMIT, "three_interpolate: batch %d > 65535", b);
    if (b == 0 || c == 0 || n == 0) return 0;
    const int bx = div_up(n, 256);
    const int cchunk = pick_cchunk(c, (long long)bx * b);
    dim3 grid(bx, div_up(c, cchunk), b);
    three_interp_kernel<<<grid, 256, 0, as_stream(stream)>>>(c, m, n, cchunk, points, idx, weight, out);
    return check_launch("three_interpolate");
}

extern "C" int amc3d_three_interpolate_grad(int b, int c, int n, int m, const float *grad_out,
                                            const int *idx, const float *weight, float *grad_points,
                                            void *stream) {
    AMC3D_REQUIRE(b >= 0 && c >= 0 && m >= 0 && n >= 0, AMC3D_EINVAL, "three_interpolate_grad: negative size");
    AMC3D_REQUIRE(b <= 65535, AMC3D_ELIMIT, "three_interpolate_grad: batch %d > 65535", b);
    if (b == 0 || c == 0 || n == 0) return 0;
    const int bx = div_up(n, 256);
    const int cchunk = pick_cchunk(c, (long long)bx * b);
    dim3 grid(bx, div_up(c, cchunk), b);
    three_interp_grad_kernel<<<grid, 256, 0, as_stream(stream)>>>(c, n, m, cchunk, grad_out, idx, weight,
                                                                  grad_points);
    return check_launch("three_interpolate_grad");
}

extern "C" int amc3d_grouping_forward(int m, int nsample, int c, const float *input, const int *idx,
                                      float *output, void *stream) {
    AMC3D_REQUIRE(m >= 0 && nsample >= 0 && c >= 0, AMC3D_EINVAL, "grouping_forward: negative size");
    const long long rows = (long long)m * nsample;
    if (rows == 0 || c == 0) return 0;
    const bool v4 = (c % 4 == 0) && ((reinterpret_cast<uintptr_t>(input) & 15) == 0) &&
                    ((reinterpret_cast<uintptr_t>(output) & 15) == 0);
    const long long total = rows * (v4 ? c / 4 : c);
    const int blocks = (int)min(div_up_ll(total, 256), (long long)kNumSMs * 32);
    if (v4) rows_gather_kernel<4><<<blocks, 256, 0, as_stream(stream)>>>(rows, c, input, idx, output);
    else rows_gather_kernel<1><<<blocks, 256, 0, as_stream(stream)>>>(rows, c, input, idx, output);
    return check_launch("grouping_forward");
}

extern "C" int amc3d_grouping_backward(int m, int nsample, int c, const float *grad_output,
                                       const int *idx, float *grad_input, void *stream) {
    AMC3D_REQUIRE(m >= 0 && nsample >= 0 && c >= 0, AMC3D_EINVAL, "grouping_backward: negative size");
    const long long rows = (long long)m * nsample;
    if (rows == 0 || c == 0) return 0;
    const bool v4 = (c % 4 == 0) && ((reinterpret_cast<uintptr_t>(grad_input) & 15) == 0) &&
                    ((reinterpret_cast<uintptr_t>(grad_output) & 15) == 0);
    const long long total = rows * (v4 ? c / 4 : c);
    const int blocks = (int)min(div_up_ll(total, 256), (long long)kNumSMs * 32);
    if (v4) rows_scatter_add_kernel<4><<<blocks, 256, 0, as_stream(stream)>>>(rows, c, grad_output, idx, grad_input);
    else rows_scatter_add_kernel<1><<<blocks, 256, 0, as_stream(stream)>>>(rows, c, grad_output, idx, grad_input);
    return check_launch("grouping_backward");
}
