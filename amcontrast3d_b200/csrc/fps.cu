// Furthest point sampling for sm_100a: one thread-block CLUSTER per scene.
//
// Replaces openpoints/cpp/pointnet2_batch/src/sampling_gpu.cu:100-260 (one 1024-thread CTA
// per scene; every round re-reads xyz and the running distance from global memory and runs a
// 10-step shared-memory tree with a __syncthreads per step).
//
// FPS is a chain of m dependent rounds, so the only thing that matters is the latency of a
// round.  Design (DESIGN.md "FPS"):
//   * a cluster of CS CTAs (up to 16, non-portable size) of 4 warps owns a scene; each thread
//     keeps its points' x, y, z and running min-distance in REGISTERS for all m rounds — global
//     memory is touched only at start and end;
//   * per round each thread updates its points and tracks its best slot; a REDUX.MAX finds the
//     warp's maximum, and the one lane holding it becomes the warp's candidate (value, tie key,
//     xyz, index).  Warps 1..3 hand their candidate to warp 0 through shared memory and a
//     non-blocking bar.arrive; warp 0 reduces the four and pushes the CTA's candidate straight
//     into every CTA of the cluster with st.async (DSMEM) that completes a transaction on the
//     receiver's mbarrier; every warp then waits on its own CTA's mbarrier and reduces the CS
//     candidates itself (one per lane, REDUX again) — one remote store and one mbarrier wait
//     per round, no cluster-wide barrier, no global-memory round trip for the winner's xyz.
//     Measured (tools/micro/exchange.cu): the all-to-all exchange alone costs ~450 cycles at
//     CS = 16 with one sender warp per CTA, ~900 with four — hence the per-CTA pre-reduction;
//   * the argmax uses the reference's exact tie order (value desc, bit-reversed (k mod bs)
//     asc, k asc with bs = the reference block size for this n).  The
//     common tie-free round never looks at tie keys of losing points; a warp that sees two
//     equal maxima falls back to an exact scan, so the index sequence is identical even on
//     lattice inputs.
#include "common.cuh"

namespace amc3d {

constexpr int FPS_WARPS = 4;
constexpr int FPS_THREADS = FPS_WARPS * 32;

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
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
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
    extern __shared__ float4 s_pts[];                // [PPT][FPS_THREADS] copy of this CTA's points
    __shared__ FpsCand s_warp[2][FPS_WARPS];         // per-warp candidates (double-buffered for CS == 1)
    __shared__ FpsCand s_exch[2][CS];                // per-CTA candidates of the whole cluster
    __shared__ __align__(8) uint64_t s_bar[2];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int rank = (int)cluster_ctarank();
    const int batch = blockIdx.x / CS;
    xyz += 3ll * batch * n;
    temp += (long long)batch * n;
    idxs += (long long)batch * m;

    const int chunk = div_up(n, CS);
    const int base = rank * chunk;

    float x[PPT], y[PPT], z[PPT], t[PPT];
#pragma unroll
    for (int s = 0; s < PPT; ++s) {
        const int l = s * FPS_THREADS + tid;
        const int k = base + l;
        if (l < chunk && k < n) {
            x[s] = __ldg(xyz + 3ll * k);
            y[s] = __ldg(xyz + 3ll * k + 1);
            z[s] = __ldg(xyz + 3ll * k + 2);
            t[s] = temp[k];
        } else {
            // +inf coordinates give d = +inf and fminf(inf, 0) = 0: a padding slot stays at
            // distance 0 and is excluded from tie resolution below
            x[s] = y[s] = z[s] = INFINITY;
            t[s] = 0.f;
        }
        s_pts[s * FPS_THREADS + tid] = make_float4(x[s], y[s], z[s], 0.f);
    }

    if (tid == 0) {
        mbar_init(smem_u32(&s_bar[0]), 1);
        mbar_init(smem_u32(&s_bar[1]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    cluster_sync_all();

    float x1 = __ldg(xyz), y1 = __ldg(xyz + 1), z1 = __ldg(xyz + 2);
    if (rank == 0 && tid == 0) idxs[0] = 0;

    // warp 0, lane r delivers this CTA's candidate to CTA r of the cluster
    const uint32_t peer = lane < CS ? lane : 0;
    const uint32_t dst0 = mapa_u32(smem_u32(&s_exch[0][rank]), peer);
    const uint32_t dst1 = mapa_u32(smem_u32(&s_exch[1][rank]), peer);
    const uint32_t rbar0 = mapa_u32(smem_u32(&s_bar[0]), peer);
    const uint32_t rbar1 = mapa_u32(smem_u32(&s_bar[1]), peer);
    const uint32_t lbar0 = smem_u32(&s_bar[0]), lbar1 = smem_u32(&s_bar[1]);

    uint32_t par = 0, phase = 0;
    for (int j = 1; j < m; ++j) {
        // ---- running distance update; every thread tracks its first maximal slot -------------
        float vmax = -1.f;
        int bslot = 0;
#pragma unroll
        for (int s = 0; s < PPT; ++s) {
            const float d = dist2_ref(x[s] - x1, y[s] - y1, z[s] - z1);
            t[s] = fminf(d, t[s]);
            if (t[s] > vmax) { vmax = t[s]; bslot = s; }
        }
        const uint32_t vb = __float_as_uint(vmax);
        const uint32_t wv = __reduce_max_sync(0xffffffffu, vb);
        const bool mine = vb == wv;
        int neq = 0;
        if (mine) {
#pragma unroll
            for (int s = 0; s < PPT; ++s) neq += __float_as_uint(t[s]) == wv ? 1 : 0;
        }
        const uint32_t bal = __ballot_sync(0xffffffffu, mine);
        const uint32_t multi = __ballot_sync(0xffffffffu, mine && neq > 1);
        int src = __ffs(bal) - 1;                        // the lane holding the warp's candidate
        uint32_t mytb = 0xffffffffu;
        if (multi != 0 || (bal & (bal - 1)) != 0) {
            // exact tie resolution (rare): lowest tie key among all slots equal to the maximum
            if (mine) {
#pragma unroll
                for (int s = 0; s < PPT; ++s) {
                    const int l = s * FPS_THREADS + tid;
                    const bool valid = l < chunk && base + l < n;      // padding slots never win a tie
                    const uint32_t tbs = valid ? tie_key(base + l, log2bs) : 0xffffffffu;
                    if (__float_as_uint(t[s]) == wv && tbs < mytb) { mytb = tbs; bslot = s; }
                }
            }
            const uint32_t wtb = __reduce_min_sync(0xffffffffu, mytb);
            src = __ffs(__ballot_sync(0xffffffffu, mine && mytb == wtb)) - 1;
        } else if (mine) {
            mytb = tie_key(base + bslot * FPS_THREADS + tid, log2bs);
        }
        // ---- the candidate lane publishes (value, tie key, xyz, index) for its warp ------------
        if (lane == src) {
            const float4 cp = s_pts[bslot * FPS_THREADS + tid];
            FpsCand c;
            c.v = wv; c.tb = mytb; c.x = cp.x; c.y = cp.y; c.z = cp.z;
            c.k = base + bslot * FPS_THREADS + tid; c.pad0 = 0; c.pad1 = 0;
            s_warp[CS == 1 ? par : 0][warp] = c;
        }
        FpsCand c;
        c.v = 0; c.tb = 0xffffffffu; c.x = c.y = c.z = 0.f; c.k = 0;
        if (CS == 1) {
            // single CTA: one barrier, then every warp reduces the 4 warp candidates itself
            __syncthreads();
            if (lane < FPS_WARPS) c = s_warp[par][lane];
        } else {
            // warps 1..3 only signal; warp 0 collects, reduces and pushes the CTA's candidate into
            // every CTA of the cluster (st.async completes a transaction on the receiver's mbarrier)
            const uint32_t lbar = par ? lbar1 : lbar0;
            if (warp != 0) {
                asm volatile("bar.arrive 1, %0;" ::"r"(FPS_THREADS) : "memory");
            } else {
                asm volatile("bar.sync 1, %0;" ::"r"(FPS_THREADS) : "memory");
                if (lane == 0) mbar_expect_tx(lbar, CS * (uint32_t)sizeof(FpsCand));
                FpsCand w;
                w.v = 0; w.tb = 0xffffffffu; w.x = w.y = w.z = 0.f; w.k = 0;
                if (lane < FPS_WARPS) w = s_warp[0][lane];
                const uint32_t bv = __reduce_max_sync(0xffffffffu, w.v);
                const uint32_t bt = __reduce_min_sync(0xffffffffu, w.v == bv ? w.tb : 0xffffffffu);
                const int wl = __ffs(__ballot_sync(0xffffffffu, w.v == bv && w.tb == bt)) - 1;
                const uint32_t w2 = __shfl_sync(0xffffffffu, __float_as_uint(w.x), wl);
                const uint32_t w3 = __shfl_sync(0xffffffffu, __float_as_uint(w.y), wl);
                const uint32_t w4 = __shfl_sync(0xffffffffu, __float_as_uint(w.z), wl);
                const uint32_t w5 = __shfl_sync(0xffffffffu, (uint32_t)w.k, wl);
                if (lane < CS) {
                    const uint32_t dst = par ? dst1 : dst0, rbar = par ? rbar1 : rbar0;
                    st_async_v4(dst, rbar, bv, bt, w2, w3);
                    st_async_v4(dst + 16, rbar, w4, w5, 0u, 0u);
                }
            }
            mbar_wait(lbar, phase);
            if (lane < CS) c = s_exch[par][lane];
        }
        // ---- every warp reduces the candidates (one per lane) with REDUX ------------------------
        const uint32_t gv = __reduce_max_sync(0xffffffffu, c.v);
        const uint32_t gt = __reduce_min_sync(0xffffffffu, c.v == gv ? c.tb : 0xffffffffu);
        const int gl = __ffs(__ballot_sync(0xffffffffu, c.v == gv && c.tb == gt)) - 1;
        x1 = __shfl_sync(0xffffffffu, c.x, gl);
        y1 = __shfl_sync(0xffffffffu, c.y, gl);
        z1 = __shfl_sync(0xffffffffu, c.z, gl);
        if (rank == 0 && warp == 0) {
            const int kwin = __shfl_sync(0xffffffffu, c.k, gl);
            if (lane == 0) idxs[j] = kwin;
        }
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
    cluster_sync_all();  // no CTA may exit while a peer can still address its shared memory
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
    const size_t smem = sizeof(float4) * PPT * FPS_THREADS;
    cudaError_t e;
    if (smem > 40 * 1024) {
        e = cudaFuncSetAttribute(fps_cluster_kernel<CS, PPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    if (CS > 8) {
        e = cudaFuncSetAttribute(fps_cluster_kernel<CS, PPT>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        if (e != cudaSuccess) return e;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(b * CS);
    cfg.blockDim = dim3(FPS_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CS;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, fps_cluster_kernel<CS, PPT>, n, m, log2bs, xyz, temp, idx);
}

template <int CS>
static cudaError_t launch_small(int ppt, int b, int n, int m, int log2bs, const float *xyz, float *temp, int *idx,
                                cudaStream_t st) {
    if (ppt <= 1) return launch_cluster<CS, 1>(b, n, m, log2bs, xyz, temp, idx, st);
    if (ppt <= 2) return launch_cluster<CS, 2>(b, n, m, log2bs, xyz, temp, idx, st);
    return launch_cluster<CS, 3>(b, n, m, log2bs, xyz, temp, idx, st);
}

template <int CS>
static cudaError_t launch_for_ppt(int ppt, int b, int n, int m, int log2bs, const float *xyz, float *temp, int *idx,
                                  cudaStream_t st) {
    if (ppt <= 1) return launch_cluster<CS, 1>(b, n, m, log2bs, xyz, temp, idx, st);
    if (ppt <= 2) return launch_cluster<CS, 2>(b, n, m, log2bs, xyz, temp, idx, st);
    if (ppt <= 3) return launch_cluster<CS, 3>(b, n, m, log2bs, xyz, temp, idx, st);
    if (ppt <= 4) return launch_cluster<CS, 4>(b, n, m, log2bs, xyz, temp, idx, st);
    if (ppt <= 6) return launch_cluster<CS, 6>(b, n, m, log2bs, xyz, temp, idx, st);
    if (ppt <= 8) return launch_cluster<CS, 8>(b, n, m, log2bs, xyz, temp, idx, st);
    if (ppt <= 12) return launch_cluster<CS, 12>(b, n, m, log2bs, xyz, temp, idx, st);
    if (ppt <= 16) return launch_cluster<CS, 16>(b, n, m, log2bs, xyz, temp, idx, st);
    if (ppt <= 24) return launch_cluster<CS, 24>(b, n, m, log2bs, xyz, temp, idx, st);
    return launch_cluster<CS, 32>(b, n, m, log2bs, xyz, temp, idx, st);
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

    // cluster size: as many CTAs as pay off (each round costs one exchange regardless)
    cudaError_t e = cudaErrorInvalidValue;
    const int cs = n > 1536 ? 16 : (n > 384 ? 4 : 1);
    const int ppt = div_up(div_up(n, cs), FPS_THREADS);
    if (ppt <= 32) {
        if (cs == 16) {
            e = launch_for_ppt<16>(ppt, b, n, m, log2bs, xyz, temp, idx, st);
            if (e != cudaSuccess) {            // non-portable cluster size refused: try 8 CTAs
                cudaGetLastError();
                const int ppt8 = div_up(div_up(n, 8), FPS_THREADS);
                if (ppt8 <= 32) e = launch_for_ppt<8>(ppt8, b, n, m, log2bs, xyz, temp, idx, st);
            }
        } else if (cs == 4) {
            e = launch_small<4>(ppt, b, n, m, log2bs, xyz, temp, idx, st);
        } else {
            e = launch_small<1>(ppt, b, n, m, log2bs, xyz, temp, idx, st);
        }
    }
    if (e != cudaSuccess) {                    // very large scenes (or clusters unavailable)
        cudaGetLastError();
        fps_global_kernel<<<b, FPS_G_THREADS, 0, st>>>(n, m, log2bs, xyz, temp, idx);
    }
    return check_launch("furthest_point_sampling");
}
