// ball_query and three_nn on batched (B,N,3) clouds for sm_100a.
//
// Replaces openpoints/cpp/pointnet2_batch/src/ball_query_gpu.cu:15-73 and
// interpolate_gpu.cu:16-81, where every thread streams the whole support cloud from global
// memory with three scalar loads per point.  Here the support cloud of the batch element is
// staged once per CTA through shared memory in the same grouped-SoA tiles as the kNN kernel
// (one broadcast LDS.128 per 1.33 points), 256 queries per CTA.  Distances are the
// reference expression bit for bit (common.cuh dist2_ref, operand order new - support).
#include "common.cuh"
#include <stdlib.h>

namespace amc3d {

constexpr int BQ_THREADS = 256;
constexpr int BQ_TILE = 1024;
constexpr int BQ_GROUPS = BQ_TILE / 4;

__device__ __forceinline__ void load_tile_b(float4 *tile, const float *__restrict__ xyz, int t0,
                                            int cnt) {
    float *tf = reinterpret_cast<float *>(tile);
    const int padded = (cnt + 3) & ~3;
    for (int i = threadIdx.x; i < padded; i += blockDim.x) {
        float x, y, z;
        if (i < cnt) {
            const float *p = xyz + 3ll * (t0 + i);
            x = __ldg(p);
            y = __ldg(p + 1);
            z = __ldg(p + 2);
        } else {
            x = y = z = __int_as_float(0x7f800000);  // +inf: never inside a ball, never nearest
        }
        const int g = i >> 2, l = i & 3;
        tf[g * 12 + l] = x;
        tf[g * 12 + 4 + l] = y;
        tf[g * 12 + 8 + l] = z;
    }
}

// Semantics (ball_query_gpu.cu:29-50): scan k = 0..n-1; on a hit (d2 < r2, strict) with
// cnt == 0 fill the whole row with k; idx[cnt++] = k; stop at nsample hits.  Rows without a
// hit are not written.
__global__ void __launch_bounds__(BQ_THREADS)
ball_query_kernel(int n, int m, float radius, int nsample, const float *__restrict__ new_xyz,
                  const float *__restrict__ xyz, int *__restrict__ idx) {
    __shared__ float4 tile[BQ_GROUPS * 3];
    const int b = blockIdx.y;
    const int q = blockIdx.x * BQ_THREADS + threadIdx.x;
    const bool active = q < m;
    const int qq = active ? q : m - 1;
    xyz += 3ll * b * n;
    const float *qp = new_xyz + 3ll * ((long long)b * m + qq);
    int *row = idx + ((long long)b * m + qq) * nsample;

    const float r2 = __fmul_rn(radius, radius);
    const float qx = __ldg(qp), qy = __ldg(qp + 1), qz = __ldg(qp + 2);
    int cnt = active ? 0 : nsample;  // inactive lanes count as finished
    int first = -1;

    for (int t0 = 0; t0 < n; t0 += BQ_TILE) {
        const int tcnt = min(BQ_TILE, n - t0);
        // the barrier doubles as the early exit: stop when every query of the CTA is full
        if (__syncthreads_and(cnt >= nsample)) break;
        load_tile_b(tile, xyz, t0, tcnt);
        __syncthreads();
        if (cnt >= nsample) continue;
        const int groups = (tcnt + 3) >> 2;
#pragma unroll 2
        for (int g = 0; g < groups; ++g) {
            const float4 X = tile[g * 3], Y = tile[g * 3 + 1], Z = tile[g * 3 + 2];
            const float d0 = dist2_ref(qx - X.x, qy - Y.x, qz - Z.x);
            const float d1 = dist2_ref(qx - X.y, qy - Y.y, qz - Z.y);
            const float d2 = dist2_ref(qx - X.z, qy - Y.z, qz - Z.z);
            const float d3 = dist2_ref(qx - X.w, qy - Y.w, qz - Z.w);
            if (fminf(fminf(d0, d1), fminf(d2, d3)) < r2) {
                const float dd[4] = {d0, d1, d2, d3};
#pragma unroll
                for (int l = 0; l < 4; ++l) {
                    if (dd[l] < r2 && cnt < nsample) {
                        const int k = t0 + g * 4 + l;
                        if (cnt == 0) first = k;
                        row[cnt++] = k;
                    }
                }
                if (cnt >= nsample) break;
            }
        }
    }
    if (active && first >= 0)
        for (int l = cnt; l < nsample; ++l) row[l] = first;
}

// Same semantics, one WARP per query: the 32 lanes test 32 consecutive support points at a time, a
// ballot gives the hits in index order, and the scan stops as soon as nsample hits are in — for the
// small clouds of the deep levels (n <= 2048) this exposes 32x more parallelism than a thread per
// query and makes the early exit effective (a thread-per-query CTA runs until its slowest query).
__global__ void __launch_bounds__(256)
ball_query_warp_kernel(int n, int m, float radius, int nsample, const float *__restrict__ new_xyz,
                       const float *__restrict__ xyz, int *__restrict__ idx) {
    const int lane = threadIdx.x & 31;
    const int b = blockIdx.y;
    const int q = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (q >= m) return;
    xyz += 3ll * b * n;
    const float *qp = new_xyz + 3ll * ((long long)b * m + q);
    int *row = idx + ((long long)b * m + q) * nsample;
    const float r2 = __fmul_rn(radius, radius);
    const float qx = __ldg(qp), qy = __ldg(qp + 1), qz = __ldg(qp + 2);
    const uint32_t lt = (1u << lane) - 1u;
    int cnt = 0, first = -1;
    for (int k0 = 0; k0 < n && cnt < nsample; k0 += 64) {
        bool hit[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int k = k0 + h * 32 + lane;
            hit[h] = false;
            if (k < n) {
                const float *p = xyz + 3ll * k;
                hit[h] = dist2_ref(qx - __ldg(p), qy - __ldg(p + 1), qz - __ldg(p + 2)) < r2;
            }
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const uint32_t mask = __ballot_sync(0xffffffffu, hit[h]);
            if (mask) {
                if (first < 0) first = k0 + h * 32 + __ffs(mask) - 1;
                const int pos = cnt + __popc(mask & lt);
                if (hit[h] && pos < nsample) row[pos] = k0 + h * 32 + lane;
                cnt += __popc(mask);
            }
        }
    }
    if (first >= 0)
        for (int l = min(cnt, nsample) + lane; l < nsample; l += 32) row[l] = first;
}

// Semantics (interpolate_gpu.cu:37-58): ascending k, strict '<' cascade over three bests
// initialised to 1e40 (double) / index 0.  Floats compared as doubles compare identically,
// and (float)1e40 = +inf, so float bests initialised to +inf reproduce it exactly.
__global__ void __launch_bounds__(BQ_THREADS)
three_nn_kernel(int n, int m, const float *__restrict__ unknown, const float *__restrict__ known,
                float *__restrict__ dist2, int *__restrict__ idx) {
    __shared__ float4 tile[BQ_GROUPS * 3];
    const int b = blockIdx.y;
    const int q = blockIdx.x * BQ_THREADS + threadIdx.x;
    const bool active = q < n;
    const int qq = active ? q : n - 1;
    known += 3ll * b * m;
    const float *qp = unknown + 3ll * ((long long)b * n + qq);
    const float ux = __ldg(qp), uy = __ldg(qp + 1), uz = __ldg(qp + 2);

    const float inf = __int_as_float(0x7f800000);
    float b1 = inf, b2 = inf, b3 = inf;
    int i1 = 0, i2 = 0, i3 = 0;

    for (int t0 = 0; t0 < m; t0 += BQ_TILE) {
        const int tcnt = min(BQ_TILE, m - t0);
        __syncthreads();
        load_tile_b(tile, known, t0, tcnt);
        __syncthreads();
        const int groups = (tcnt + 3) >> 2;
#pragma unroll 2
        for (int g = 0; g < groups; ++g) {
            const float4 X = tile[g * 3], Y = tile[g * 3 + 1], Z = tile[g * 3 + 2];
            const float d0 = dist2_ref(ux - X.x, uy - Y.x, uz - Z.x);
            const float d1 = dist2_ref(ux - X.y, uy - Y.y, uz - Z.y);
            const float d2 = dist2_ref(ux - X.z, uy - Y.z, uz - Z.z);
            const float d3 = dist2_ref(ux - X.w, uy - Y.w, uz - Z.w);
            if (fminf(fminf(d0, d1), fminf(d2, d3)) < b3) {
                const float dd[4] = {d0, d1, d2, d3};
#pragma unroll
                for (int l = 0; l < 4; ++l) {
                    const float d = dd[l];
                    const int k = t0 + g * 4 + l;
                    if (d < b1) {
                        b3 = b2; i3 = i2;
                        b2 = b1; i2 = i1;
                        b1 = d;  i1 = k;
                    } else if (d < b2) {
                        b3 = b2; i3 = i2;
                        b2 = d;  i2 = k;
                    } else if (d < b3) {
                        b3 = d;  i3 = k;
                    }
                }
            }
        }
    }
    if (active) {
        const long long o = 3ll * ((long long)b * n + q);
        dist2[o] = b1; dist2[o + 1] = b2; dist2[o + 2] = b3;
        idx[o] = i1;   idx[o + 1] = i2;   idx[o + 2] = i3;
    }
}

// knn_grid.cu: the same searches with spatial culling (identical results)
int knn_grid_batched(int nb, int n, int m, int nsample, const float *xyz, const float *new_xyz, int *idx,
                     float *dist2, cudaStream_t st, int *order_out = nullptr);
int ball_grid_batched(int nb, int n, int m, float radius, int nsample, const float *xyz, const float *new_xyz,
                      int *idx, cudaStream_t st);

// support points per cloud from which the culled search pays for its sort (AMC3D_GRID_MIN overrides;
// 0 disables the culled path)
static int grid_min() {
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("AMC3D_GRID_MIN");
        v = e ? atoi(e) : 2048;
    }
    return v;
}

}  // namespace amc3d

using namespace amc3d;

extern "C" int amc3d_ball_query(int b, int n, int m, float radius, int nsample,
                                const float *new_xyz, const float *xyz, int *idx, void *stream) {
    AMC3D_REQUIRE(b >= 0 && n >= 0 && m >= 0 && nsample >= 1, AMC3D_EINVAL,
                  "ball_query: bad sizes b=%d n=%d m=%d nsample=%d", b, n, m, nsample);
    AMC3D_REQUIRE(b <= 65535, AMC3D_ELIMIT, "ball_query: batch %d > 65535", b);
    if (b == 0 || m == 0 || n == 0) return 0;
    if (grid_min() > 0 && n >= grid_min() && nsample <= 128) {
        const int rc = ball_grid_batched(b, n, m, radius, nsample, xyz, new_xyz, idx, as_stream(stream));
        if (rc != 0) {
            set_error("ball_query (grid): %s", cudaGetErrorString((cudaError_t)rc));
            return rc;
        }
        return check_launch("ball_query");
    }
    static const bool thread_per_query = getenv("AMC3D_BALL_THREAD") != nullptr;   // for measurements
    if (thread_per_query) {
        dim3 grid(div_up(m, BQ_THREADS), b);
        ball_query_kernel<<<grid, BQ_THREADS, 0, as_stream(stream)>>>(n, m, radius, nsample, new_xyz, xyz, idx);
    } else {
        dim3 grid(div_up(m, 8), b);
        ball_query_warp_kernel<<<grid, 256, 0, as_stream(stream)>>>(n, m, radius, nsample, new_xyz, xyz, idx);
    }
    return check_launch("ball_query");
}

extern "C" int amc3d_three_nn(int b, int n, int m, const float *unknown, const float *known,
                              float *dist2, int *idx, void *stream) {
    AMC3D_REQUIRE(b >= 0 && n >= 0 && m >= 0, AMC3D_EINVAL, "three_nn: bad sizes b=%d n=%d m=%d", b, n, m);
    AMC3D_REQUIRE(b <= 65535, AMC3D_ELIMIT, "three_nn: batch %d > 65535", b);
    if (b == 0 || n == 0) return 0;
    if (grid_min() > 0 && m >= grid_min()) {
        // three_nn == exact 3-NN ordered by (d2, index): support = known, queries = unknown
        const int rc = knn_grid_batched(b, m, n, 3, known, unknown, idx, dist2, as_stream(stream));
        if (rc != 0) {
            set_error("three_nn (grid): %s", cudaGetErrorString((cudaError_t)rc));
            return rc;
        }
        return check_launch("three_nn");
    }
    dim3 grid(div_up(n, BQ_THREADS), b);
    three_nn_kernel<<<grid, BQ_THREADS, 0, as_stream(stream)>>>(n, m, unknown, known, dist2, idx);
    return check_launch("three_nn");
}
