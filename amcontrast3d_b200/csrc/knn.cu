// Exact brute-force kNN inside offset segments (pointops.knnquery) for sm_100a.
//
// Replaces openpoints/cpp/pointops/src/knnquery/knnquery_cuda_kernel.cu:65-116 (one thread
// per query, 800-byte local-memory heap, whole segment streamed from global per thread).
//
// Design (DESIGN.md "kNN"):
//   * one thread per query, 256 queries per CTA; the support segment is streamed through
//     shared memory in tiles of 1024 points laid out as groups of 4 points
//     {x0..x3 | y0..y3 | z0..z3} so that one broadcast LDS.128 feeds 4 distance evaluations
//     of every lane (0.75 LDS per pair instead of 3 scalar global loads);
//   * the squared distance is the reference expression bit for bit
//     (FADD x3, FMUL, FFMA, FFMA — common.cuh dist2_ref), operand order new - support;
//   * top-k lives in registers as a sorted list (k <= 32, fully unrolled predicated
//     insertion); the support is scanned in ascending index order with a strict '<'
//     against the current k-th best, so the result is the (d2, index)-lexicographic top-k —
//     identical to the reference heap on tie-free inputs;
//   * 32 < k <= 128 (label vote kr = 64) keeps the sorted list in a thread-private,
//     bank-conflict-free shared-memory column instead.
// FP32-pipe bound: 6 FP32 instructions + ~1.3 compare/select per pair.
#include "common.cuh"
#include <stdlib.h>

namespace amc3d {

constexpr int KNN_THREADS = 256;
constexpr int KNN_TILE = 1024;                 // support points per shared-memory tile
constexpr int KNN_GROUPS = KNN_TILE / 4;       // groups of 4 points
constexpr float KNN_INIT = 1e10f;              // reference initial distance (:89)

// smallest s with q < new_offset[s] (reference get_bt_idx, :51-62), clamped to nseg-1
__device__ __forceinline__ int find_segment(int q, const int *__restrict__ new_offset, int nseg) {
    int s = 0;
    while (s < nseg - 1 && q >= __ldg(new_offset + s)) ++s;
    return s;
}

// Stage support points [t0, t0+cnt) into the grouped SoA tile; pad to a multiple of 4 with
// +inf so that padded lanes can never pass 'd < thr'.
__device__ __forceinline__ void load_tile(float4 *tile, const float *__restrict__ xyz, int t0,
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
            x = y = z = __int_as_float(0x7f800000);
        }
        const int g = i >> 2, l = i & 3;
        tf[g * 12 + l] = x;
        tf[g * 12 + 4 + l] = y;
        tf[g * 12 + 8 + l] = z;
    }
}

// Sorted (ascending) register list of K entries.  When nsample < K the live entries are
// right-aligned in slots [K-nsample, K) and the leading slots hold -inf sentinels that no
// candidate can displace, so the threshold is always the compile-time slot K-1 (a
// run-time slot index would push the whole list into local memory).
template <int K>
struct RegList {
    float d[K];
    int i[K];
    __device__ __forceinline__ void init(int start, int nsample) {
#pragma unroll
        for (int s = 0; s < K; ++s) {
            d[s] = s < K - nsample ? __int_as_float(0xff800000) : KNN_INIT;
            i[s] = start;
        }
    }
    // precondition: nd < d[K-1].  Stable: goes after every entry with distance <= nd.
    __device__ __forceinline__ void insert(float nd, int ni) {
#pragma unroll
        for (int s = K - 1; s > 0; --s) {
            const bool up = nd < d[s - 1];
            const bool here = nd < d[s];
            d[s] = up ? d[s - 1] : (here ? nd : d[s]);
            i[s] = up ? i[s - 1] : (here ? ni : i[s]);
        }
        if (nd < d[0]) {
            d[0] = nd;
            i[0] = ni;
        }
    }
};

// KLIST = register list length; the launcher picks the smallest KLIST >= nsample.
template <int KLIST, bool CHECK_RANGE>
__global__ void __launch_bounds__(KNN_THREADS)
knn_reg_kernel(int n, int m, int nseg, int nsample, const float *__restrict__ xyz,
               const float *__restrict__ new_xyz, const int *__restrict__ offset,
               const int *__restrict__ new_offset, int *__restrict__ idx,
               float *__restrict__ dist2) {
    __shared__ float4 tile[KNN_GROUPS * 3];
    __shared__ int s_range[2];

    const int q = blockIdx.x * KNN_THREADS + threadIdx.x;
    const bool active = q < m;
    const int qq = active ? q : m - 1;

    const int seg = find_segment(qq, new_offset, nseg);
    const int start = seg == 0 ? 0 : __ldg(offset + seg - 1);
    const int end = min(__ldg(offset + seg), n);

    // queries are ordered by segment: the CTA's support range is [start of the first
    // thread's segment, end of the last active thread's segment)
    const int last = min(m, (blockIdx.x + 1) * KNN_THREADS) - 1 - blockIdx.x * KNN_THREADS;
    if (threadIdx.x == 0) s_range[0] = start;
    if (threadIdx.x == last) s_range[1] = end;
    __syncthreads();
    const int lo = s_range[0], hi = s_range[1];

    const float qx = __ldg(new_xyz + 3ll * qq);
    const float qy = __ldg(new_xyz + 3ll * qq + 1);
    const float qz = __ldg(new_xyz + 3ll * qq + 2);

    RegList<KLIST> best;
    best.init(start, nsample);
    float thr = KNN_INIT;

    for (int t0 = lo; t0 < hi; t0 += KNN_TILE) {
        const int cnt = min(KNN_TILE, hi - t0);
        __syncthreads();
        load_tile(tile, xyz, t0, cnt);
        __syncthreads();
        const int groups = (cnt + 3) >> 2;
#pragma unroll 2
        for (int g = 0; g < groups; ++g) {
            const float4 X = tile[g * 3], Y = tile[g * 3 + 1], Z = tile[g * 3 + 2];
            const float d0 = dist2_ref(qx - X.x, qy - Y.x, qz - Z.x);
            const float d1 = dist2_ref(qx - X.y, qy - Y.y, qz - Z.y);
            const float d2 = dist2_ref(qx - X.z, qy - Y.z, qz - Z.z);
            const float d3 = dist2_ref(qx - X.w, qy - Y.w, qz - Z.w);
            if (fminf(fminf(d0, d1), fminf(d2, d3)) < thr) {
                const int s0 = t0 + g * 4;
                const float dd[4] = {d0, d1, d2, d3};
#pragma unroll
                for (int l = 0; l < 4; ++l) {
                    bool ok = dd[l] < thr;
                    if (CHECK_RANGE) ok = ok && (s0 + l >= start) && (s0 + l < end);
                    if (ok) {
                        best.insert(dd[l], s0 + l);
                        thr = best.d[KLIST - 1];
                    }
                }
            }
        }
    }

    if (active) {
        // live entries are slots [KLIST-nsample, KLIST): the shift moves to the address
        int *oi = idx + (long long)q * nsample - (KLIST - nsample);
        float *od = dist2 + (long long)q * nsample - (KLIST - nsample);
#pragma unroll
        for (int s = 0; s < KLIST; ++s)
            if (s >= KLIST - nsample) {
                oi[s] = best.i[s];
                od[s] = best.d[s];
            }
    }
}

// ---- 32 < nsample <= 128: sorted list in a thread-private shared-memory column -----------
constexpr int KNN_BIG_THREADS = 128;

__global__ void __launch_bounds__(KNN_BIG_THREADS)
knn_smem_kernel(int n, int m, int nseg, int nsample, const float *__restrict__ xyz,
                const float *__restrict__ new_xyz, const int *__restrict__ offset,
                const int *__restrict__ new_offset, int *__restrict__ idx,
                float *__restrict__ dist2) {
    extern __shared__ float4 dyn[];
    float4 *tile = dyn;                                                  // KNN_GROUPS*3 float4
    float *ld = reinterpret_cast<float *>(tile + KNN_GROUPS * 3);        // [nsample][threads]
    int *li = reinterpret_cast<int *>(ld + nsample * KNN_BIG_THREADS);   // [nsample][threads]
    __shared__ int s_range[2];

    const int tid = threadIdx.x;
    const int q = blockIdx.x * KNN_BIG_THREADS + tid;
    const bool active = q < m;
    const int qq = active ? q : m - 1;
    const int seg = find_segment(qq, new_offset, nseg);
    const int start = seg == 0 ? 0 : __ldg(offset + seg - 1);
    const int end = min(__ldg(offset + seg), n);
    const int last = min(m, (blockIdx.x + 1) * KNN_BIG_THREADS) - 1 - blockIdx.x * KNN_BIG_THREADS;
    if (tid == 0) s_range[0] = start;
    if (tid == last) s_range[1] = end;
    for (int s = 0; s < nsample; ++s) {
        ld[s * KNN_BIG_THREADS + tid] = KNN_INIT;
        li[s * KNN_BIG_THREADS + tid] = start;
    }
    __syncthreads();
    const int lo = s_range[0], hi = s_range[1];

    const float qx = __ldg(new_xyz + 3ll * qq);
    const float qy = __ldg(new_xyz + 3ll * qq + 1);
    const float qz = __ldg(new_xyz + 3ll * qq + 2);
    float thr = KNN_INIT;

    for (int t0 = lo; t0 < hi; t0 += KNN_TILE) {
        const int cnt = min(KNN_TILE, hi - t0);
        __syncthreads();
        load_tile(tile, xyz, t0, cnt);
        __syncthreads();
        const int groups = (cnt + 3) >> 2;
        for (int g = 0; g < groups; ++g) {
            const float4 X = tile[g * 3], Y = tile[g * 3 + 1], Z = tile[g * 3 + 2];
            const float dd[4] = {dist2_ref(qx - X.x, qy - Y.x, qz - Z.x),
                                 dist2_ref(qx - X.y, qy - Y.y, qz - Z.y),
                                 dist2_ref(qx - X.z, qy - Y.z, qz - Z.z),
                                 dist2_ref(qx - X.w, qy - Y.w, qz - Z.w)};
            if (fminf(fminf(dd[0], dd[1]), fminf(dd[2], dd[3])) < thr) {
#pragma unroll
                for (int l = 0; l < 4; ++l) {
                    const int s = t0 + g * 4 + l;
                    if (dd[l] < thr && s >= start && s < end) {
                        // shift entries greater than dd[l] up by one, from the tail
                        int pos = nsample - 1;
                        while (pos > 0 && dd[l] < ld[(pos - 1) * KNN_BIG_THREADS + tid]) {
                            ld[pos * KNN_BIG_THREADS + tid] = ld[(pos - 1) * KNN_BIG_THREADS + tid];
                            li[pos * KNN_BIG_THREADS + tid] = li[(pos - 1) * KNN_BIG_THREADS + tid];
                            --pos;
                        }
                        ld[pos * KNN_BIG_THREADS + tid] = dd[l];
                        li[pos * KNN_BIG_THREADS + tid] = s;
                        thr = ld[(nsample - 1) * KNN_BIG_THREADS + tid];
                    }
                }
            }
        }
    }
    if (active) {
        for (int s = 0; s < nsample; ++s) {
            idx[(long long)q * nsample + s] = li[s * KNN_BIG_THREADS + tid];
            dist2[(long long)q * nsample + s] = ld[s * KNN_BIG_THREADS + tid];
        }
    }
}

template <int KLIST>
static void launch_reg(bool check, int blocks, cudaStream_t st, int n, int m, int nseg, int nsample,
                       const float *xyz, const float *new_xyz, const int *offset,
                       const int *new_offset, int *idx, float *dist2) {
    if (check)
        knn_reg_kernel<KLIST, true><<<blocks, KNN_THREADS, 0, st>>>(n, m, nseg, nsample, xyz, new_xyz,
                                                                     offset, new_offset, idx, dist2);
    else
        knn_reg_kernel<KLIST, false><<<blocks, KNN_THREADS, 0, st>>>(n, m, nseg, nsample, xyz, new_xyz,
                                                                      offset, new_offset, idx, dist2);
}

int knn_grid_single_segment(int n, int m, int nsample, const float *xyz, const float *new_xyz, int *idx,
                            float *dist2, cudaStream_t st, int *order_out);   // knn_grid.cu

__global__ void iota_kernel(int m, int *out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m) out[i] = i;
}

}  // namespace amc3d

using namespace amc3d;

// AMC3D_KNN_BRUTE=1 forces the brute-force kernels (used to measure them; results are identical)
static bool force_brute() {
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("AMC3D_KNN_BRUTE");
        v = (e && e[0] == '1') ? 1 : 0;
    }
    return v == 1;
}

extern "C" int amc3d_knnquery(int n, int m, int nseg, int nsample, const float *xyz,
                              const float *new_xyz, const int *offset, const int *new_offset,
                              int *idx, float *dist2, void *stream) {
    return amc3d_knnquery_order(n, m, nseg, nsample, xyz, new_xyz, offset, new_offset, idx, dist2, nullptr, stream);
}

extern "C" int amc3d_knnquery_order(int n, int m, int nseg, int nsample, const float *xyz,
                                    const float *new_xyz, const int *offset, const int *new_offset,
                                    int *idx, float *dist2, int *order, void *stream) {
    AMC3D_REQUIRE(n >= 0 && m >= 0 && nseg >= 1 && nsample >= 1, AMC3D_EINVAL,
                  "knnquery: bad sizes n=%d m=%d nseg=%d nsample=%d", n, m, nseg, nsample);
    AMC3D_REQUIRE(nsample <= 128, AMC3D_ELIMIT, "knnquery: nsample=%d > 128", nsample);
    if (m == 0) return 0;
    cudaStream_t st = as_stream(stream);
    // One segment (offset = [n], new_offset = [m] — what AMContrast3D always passes): exact search with
    // spatial culling.  Small problems are not worth the sort.
    if (nseg == 1 && n >= 2048 && !force_brute()) {
        const int rc = knn_grid_single_segment(n, m, nsample, xyz, new_xyz, idx, dist2, st, order);
        if (rc != 0) {
            set_error("knnquery (grid): %s", cudaGetErrorString((cudaError_t)rc));
            return rc;
        }
        return check_launch("knnquery");
    }
    if (order != nullptr) iota_kernel<<<div_up(m, 256), 256, 0, st>>>(m, order);   // brute force visits in index order
    // A single segment (the AMContrast3D case: offset = [B*n]) never needs the per-candidate
    // range check; with several segments a CTA may straddle a boundary, so check.
    const bool check = nseg > 1;
    if (nsample <= 32) {
        const int blocks = div_up(m, KNN_THREADS);
        if (nsample <= 4)
            launch_reg<4>(check, blocks, st, n, m, nseg, nsample, xyz, new_xyz, offset, new_offset, idx, dist2);
        else if (nsample <= 8)
            launch_reg<8>(check, blocks, st, n, m, nseg, nsample, xyz, new_xyz, offset, new_offset, idx, dist2);
        else if (nsample <= 16)
            launch_reg<16>(check, blocks, st, n, m, nseg, nsample, xyz, new_xyz, offset, new_offset, idx, dist2);
        else if (nsample <= 24)
            launch_reg<24>(check, blocks, st, n, m, nseg, nsample, xyz, new_xyz, offset, new_offset, idx, dist2);
        else
            launch_reg<32>(check, blocks, st, n, m, nseg, nsample, xyz, new_xyz, offset, new_offset, idx, dist2);
    } else {
        const int blocks = div_up(m, KNN_BIG_THREADS);
        const size_t smem = sizeof(float4) * KNN_GROUPS * 3 + (size_t)nsample * KNN_BIG_THREADS * 8;
        // the opt-in is per device and a process may drive several: set it on every call (cheap, host-only)
        const cudaError_t ae = cudaFuncSetAttribute(knn_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                    (int)(sizeof(float4) * KNN_GROUPS * 3 + 128 * KNN_BIG_THREADS * 8));
        if (ae != cudaSuccess) {
            set_error("knnquery: cudaFuncSetAttribute: %s", cudaGetErrorString(ae));
            return (int)ae;
        }
        knn_smem_kernel<<<blocks, KNN_BIG_THREADS, smem, st>>>(n, m, nseg, nsample, xyz, new_xyz, offset,
                                                               new_offset, idx, dist2);
    }
    return check_launch("knnquery");
}
