// Furthest point sampling for sm_100a: one thread-block CLUSTER per scene.
//
// Replaces openpoints/cpp/pointnet2_batch/src/sampling_gpu.cu:100-260 (one 1024-thread CTA
// per scene; every round re-reads xyz and the running distance from global memory and runs a
// 10-step shared-memory tree with a __syncthreads per step).
//
// FPS is a chain of m dependent rounds, so the only thing that matters is the latency of a
// round.  Design (DESIGN.md "FPS"):
//   * a cluster of CS CTAs (8, or 16 non-portable for n > 24576) owns a scene; each thread
//     keeps its points' x, y, z and running min-distance in REGISTERS for all m rounds —
//     global memory is touched only at start and end;
//   * per round: FP32 update + thread-local max, then REDUX.MAX/REDUX.MIN warp reductions on
//     the (distance bits, tie key) pair (2 instructions instead of a 5-step shuffle tree),
//     one __syncthreads for the 8 warps, then the CTA winner — including its coordinates —
//     is pushed into every CTA of the cluster with st.async (DSMEM) that completes a
//     transaction on the receiver's mbarrier: one remote store + one mbarrier wait per
//     round, no cluster-wide barrier, no global-memory round trip for the winner's xyz;
//   * the argmax uses the reference's exact tie order (value desc, bit-reversed
//     (k mod bs) asc, k asc with bs = the reference block size for this n; SURVEY.md
//     App. A.1), so the index sequence is identical even on tied inputs.
#include "common.cuh"

namespace amc3d {

constexpr int FPS_THREADS = 256;
constexpr int FPS_WARPS = FPS_THREADS / 32;
constexpr int FPS_MAX_CS = 16;

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_async_v4(uint32_t raddr, uint32_t rbar, uint32_t a, uint32_t b,
                                            uint32_t c, uint32_t d) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1,%2,%3,%4}, [%5];"
                 ::"r"(raddr), "r"(a), "r"(b), "r"(c), "r"(d), "r"(rbar)
                 : "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@!p bra WAIT_LOOP;\n\t}" ::"r"(bar), "r"(parity)
        : "memory");
}

// tie key: lower wins.  (bit-reversed (k mod bs)) << 22 | (k div bs)
__device__ __forceinline__ uint32_t tie_key(int k, int log2bs) {
    const uint32_t low = (uint32_t)k & ((1u << log2bs) - 1u);
    const uint32_t rev = log2bs == 0 ? 0u : (__brev(low) >> (32 - log2bs));
    return (rev << 22) | ((uint32_t)k >> log2bs);
}

struct __align__(16) FpsCand {
    uint32_t v;    // running distance bits (non-negative float: bit order == value order)
    uint32_t tb;   // tie key
    float x, y, z; // coordinates of the candidate (next round's reference point)
    int k;         // its index
    uint32_t pad0, pad1;
};
static_assert(sizeof(FpsCand) == 32, "FpsCand must be 32 bytes");

template <int CS, int PPT>
__global__ void __launch_bounds__(FPS_THREADS)
fps_cluster_kernel(int n, int m, int log2bs, const float *__restrict__ xyz, float *__restrict__ temp,
                   int *__restrict__ idxs) {
    __shared__ FpsCand s_warp[FPS_WARPS];
    __shared__ FpsCand s_exch[2][CS];
    __shared__ __align__(8) uint64_t s_bar[2];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int rank = CS > 1 ? (int)cluster_ctarank() : 0;
    const int batch = blockIdx.x / CS;
    xyz += 3ll * batch * n;
    temp += (long long)batch * n;
    idxs += (long long)batch * m;

    const int chunk = div_up(n, CS);
    const int base = rank * chunk;

    float x[PPT], y[PPT], z[PPT], t[PPT];
    uint32_t tb[PPT];
#pragma unroll
    for (int s = 0; s < PPT; ++s) {
        const int l = s * FPS_THREADS + tid;
        const int k = base + l;
        const bool valid = l < chunk && k < n;
        if (valid) {
            x[s] = __ldg(xyz + 3ll * k);
            y[s] = __ldg(xyz + 3ll * k + 1);
            z[s] = __ldg(xyz + 3ll * k + 2);
            t[s] = temp[k];
            tb[s] = tie_key(k, log2bs);
        } else {
            // +inf coordinates give d = +inf, fminf(inf, 0) = 0: the slot stays at distance 0
            // with the worst tie key and can never beat a real point
            x[s] = y[s] = z[s] = __int_as_float(0x7f800000);
            t[s] = 0.f;
            tb[s] = 0xffffffffu;
        }
    }

    if (CS > 1) {
        if (tid == 0) {
            mbar_init(smem_u32(&s_bar[0]), 1);
            mbar_init(smem_u32(&s_bar[1]), 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        cluster_sync_all();
    }

    float x1 = __ldg(xyz), y1 = __ldg(xyz + 1), z1 = __ldg(xyz + 2);
    if (rank == 0 && tid == 0) idxs[0] = 0;

    uint32_t par = 0, phase = 0;
    for (int j = 1; j < m; ++j) {
        // ---- running distance update + thread-local maximum --------------------------------
        float vmax = 0.f;
#pragma unroll
        for (int s = 0; s < PPT; ++s) {
            const float d = dist2_ref(x[s] - x1, y[s] - y1, z[s] - z1);
            t[s] = fminf(d, t[s]);
            vmax = fmaxf(vmax, t[s]);
        }
        // ---- warp argmax on (value bits desc, tie key asc) via REDUX ------------------------
        const uint32_t vb = __float_as_uint(vmax);
        const uint32_t wv = __reduce_max_sync(0xffffffffu, vb);
        uint32_t mytb = 0xffffffffu;
        if (vb == wv) {
#pragma unroll
            for (int s = 0; s < PPT; ++s)
                if (__float_as_uint(t[s]) == wv) mytb = min(mytb, tb[s]);
        }
        const uint32_t wtb = __reduce_min_sync(0xffffffffu, mytb);
        if (vb == wv && mytb == wtb) {
            FpsCand c;
            c.v = wv; c.tb = wtb; c.x = 0.f; c.y = 0.f; c.z = 0.f; c.k = 0; c.pad0 = 0; c.pad1 = 0;
#pragma unroll
            for (int s = 0; s < PPT; ++s)
                if (tb[s] == wtb) {
                    c.x = x[s]; c.y = y[s]; c.z = z[s];
                    c.k = base + s * FPS_THREADS + tid;
                }
            s_warp[warp] = c;
        }
        __syncthreads();
        // ---- CTA winner (warp 0), pushed to every CTA of the cluster ------------------------
        if (warp == 0) {
            FpsCand c;
            if (lane < FPS_WARPS) c = s_warp[lane];
            else { c.v = 0; c.tb = 0xffffffffu; }
            const uint32_t bv = __reduce_max_sync(0xffffffffu, c.v);
            const uint32_t bt = __reduce_min_sync(0xffffffffu, c.v == bv ? c.tb : 0xffffffffu);
            const bool win = lane < FPS_WARPS && c.v == bv && c.tb == bt;
            if (CS > 1) {
                if (lane == 0) mbar_expect_tx(smem_u32(&s_bar[par]), CS * (uint32_t)sizeof(FpsCand));
                // lanes 0..CS-1 each deliver the winner to one CTA: fetch it by shuffle
                const int wl = __ffs(__ballot_sync(0xffffffffu, win)) - 1;
                const uint32_t w0 = __shfl_sync(0xffffffffu, c.v, wl);
                const uint32_t w1 = __shfl_sync(0xffffffffu, c.tb, wl);
                const uint32_t w2 = __shfl_sync(0xffffffffu, __float_as_uint(c.x), wl);
                const uint32_t w3 = __shfl_sync(0xffffffffu, __float_as_uint(c.y), wl);
                const uint32_t w4 = __shfl_sync(0xffffffffu, __float_as_uint(c.z), wl);
                const uint32_t w5 = __shfl_sync(0xffffffffu, (uint32_t)c.k, wl);
                if (lane < CS) {
                    const uint32_t dst = mapa_u32(smem_u32(&s_exch[par][rank]), lane);
                    const uint32_t bar = mapa_u32(smem_u32(&s_bar[par]), lane);
                    st_async_v4(dst, bar, w0, w1, w2, w3);
                    st_async_v4(dst + 16, bar, w4, w5, 0u, 0u);
                }
            } else if (win) {
                s_exch[par][0] = c;
            }
        }
        if (CS > 1) mbar_wait(smem_u32(&s_bar[par]), phase);
        else __syncthreads();
        // ---- every thread combines the CS candidates (identical result everywhere) ----------
        FpsCand best = s_exch[par][0];
#pragma unroll
        for (int r = 1; r < CS; ++r) {
            const FpsCand c = s_exch[par][r];
            if (c.v > best.v || (c.v == best.v && c.tb < best.tb)) best = c;
        }
        x1 = best.x; y1 = best.y; z1 = best.z;
        if (rank == 0 && tid == 0) idxs[j] = best.k;
        par ^= 1;
        if (par == 0) phase ^= 1;
    }

    // the reference leaves the final running distances in temp
#pragma unroll
    for (int s = 0; s < PPT; ++s) {
        const int l = s * FPS_THREADS + tid;
        const int k = base + l;
        if (l < chunk && k < n) temp[k] = t[s];
    }
    if (CS > 1) cluster_sync_all();  // no CTA may exit while a peer can still address its smem
}

// ---- fallback for any n: one 1024-thread CTA per scene, state in global memory ------------
constexpr int FPS_G_THREADS = 1024;
__global__ void __launch_bounds__(FPS_G_THREADS)
fps_global_kernel(int n, int m, int log2bs, const float *__restrict__ xyz, float *__restrict__ temp,
                  int *__restrict__ idxs) {
    __shared__ uint32_t s_v[32], s_tb[32];
    __shared__ int s_k[32];
    __shared__ int s_old;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    xyz += 3ll * blockIdx.x * n;
    temp += (long long)blockIdx.x * n;
    idxs += (long long)blockIdx.x * m;
    int old = 0;
    if (tid == 0) idxs[0] = 0;
    for (int j = 1; j < m; ++j) {
        const float x1 = __ldg(xyz + 3ll * old), y1 = __ldg(xyz + 3ll * old + 1), z1 = __ldg(xyz + 3ll * old + 2);
        uint32_t bv = 0, bt = 0xffffffffu;
        int bk = 0;
        for (int k = tid; k < n; k += FPS_G_THREADS) {
            const float d = dist2_ref(__ldg(xyz + 3ll * k) - x1, __ldg(xyz + 3ll * k + 1) - y1,
                                      __ldg(xyz + 3ll * k + 2) - z1);
            const float d2 = fminf(d, temp[k]);
            temp[k] = d2;
            const uint32_t v = __float_as_uint(d2), tbk = tie_key(k, log2bs);
            if (v > bv || (v == bv && tbk < bt)) { bv = v; bt = tbk; bk = k; }
        }
        uint32_t wv = __reduce_max_sync(0xffffffffu, bv);
        uint32_t wt = __reduce_min_sync(0xffffffffu, bv == wv ? bt : 0xffffffffu);
        if (bv == wv && bt == wt) { s_v[warp] = wv; s_tb[warp] = wt; s_k[warp] = bk; }
        __syncthreads();
        if (warp == 0) {
            const uint32_t v = s_v[lane], t2 = s_tb[lane];
            wv = __reduce_max_sync(0xffffffffu, v);
            wt = __reduce_min_sync(0xffffffffu, v == wv ? t2 : 0xffffffffu);
            if (v == wv && t2 == wt) s_old = s_k[lane];
        }
        __syncthreads();
        old = s_old;
        if (tid == 0) idxs[j] = old;
        __syncthreads();
    }
}

template <int CS, int PPT>
static cudaError_t launch_cluster(int b, int n, int m, int log2bs, const float *xyz, float *temp, int *idx,
                                  cudaStream_t st) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(b * CS);
    cfg.blockDim = dim3(FPS_THREADS);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CS;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = CS > 1 ? 1 : 0;
    if (CS > 8) {
        cudaError_t e = cudaFuncSetAttribute(fps_cluster_kernel<CS, PPT>,
                                             cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        if (e != cudaSuccess) return e;
    }
    return cudaLaunchKernelEx(&cfg, fps_cluster_kernel<CS, PPT>, n, m, log2bs, xyz, temp, idx);
}

}  // namespace amc3d

using namespace amc3d;

extern "C" int amc3d_furthest_point_sampling(int b, int n, int m, const float *xyz, float *temp,
                                             int *idx, void *stream) {
    AMC3D_REQUIRE(b >= 0 && n >= 1 && m >= 0, AMC3D_EINVAL, "furthest_point_sampling: bad sizes b=%d n=%d m=%d", b, n, m);
    AMC3D_REQUIRE(n <= (1 << 22), AMC3D_ELIMIT, "furthest_point_sampling: n=%d > 4194304", n);
    if (b == 0 || m <= 0) return 0;  // reference kernel returns immediately for m <= 0
    cudaStream_t st = as_stream(stream);
    // reference block size: largest power of two <= n, capped at 1024 (cuda_utils.h:10-14)
    int log2bs = 0;
    while ((2 << log2bs) <= n && log2bs < 10) ++log2bs;

    cudaError_t e;
    if (n <= FPS_THREADS * 2) e = launch_cluster<1, 2>(b, n, m, log2bs, xyz, temp, idx, st);
    else if (n <= FPS_THREADS * 8) e = launch_cluster<1, 8>(b, n, m, log2bs, xyz, temp, idx, st);
    else if (n <= 8 * FPS_THREADS * 4) e = launch_cluster<8, 4>(b, n, m, log2bs, xyz, temp, idx, st);
    else if (n <= 8 * FPS_THREADS * 12) e = launch_cluster<8, 12>(b, n, m, log2bs, xyz, temp, idx, st);
    else if (n <= 16 * FPS_THREADS * 16) {
        e = launch_cluster<16, 16>(b, n, m, log2bs, xyz, temp, idx, st);
        if (e != cudaSuccess) {  // non-portable cluster size refused: fall back
            cudaGetLastError();
            fps_global_kernel<<<b, FPS_G_THREADS, 0, st>>>(n, m, log2bs, xyz, temp, idx);
            e = cudaSuccess;
        }
    } else {
        fps_global_kernel<<<b, FPS_G_THREADS, 0, st>>>(n, m, log2bs, xyz, temp, idx);
        e = cudaSuccess;
    }
    if (e != cudaSuccess) {
        set_error("furthest_point_sampling: launch failed: %s", cudaGetErrorString(e));
        return (int)e;
    }
    return check_launch("furthest_point_sampling");
}
