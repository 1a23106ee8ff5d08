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
#include <stdlib.h>

namespace amc3d {

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
__device__ __forceinline__ void st_async_b32(uint32_t raddr, uint32_t rbar, uint32_t a) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];"
                 ::"r"(raddr), "r"(a), "r"(rbar)
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
// shared-memory accesses through 32-bit shared addresses computed once outside the round loop
__device__ __forceinline__ uint4 lds_v4(uint32_t a) {
    uint4 r;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(a) : "memory");
    return r;
}
__device__ __forceinline__ uint32_t lds_b32(uint32_t a) {
    uint32_t r;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r) : "r"(a) : "memory");
    return r;
}
__device__ __forceinline__ void sts_v4(uint32_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
__device__ __forceinline__ uint2 lds_v2(uint32_t a) {
    uint2 r;
    asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "r"(a) : "memory");
    return r;
}
__device__ __forceinline__ void sts_v2(uint32_t a, uint32_t x, uint32_t y) {
    asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"(a), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void st_async_v2(uint32_t raddr, uint32_t rbar, uint32_t a, uint32_t b) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.b32 [%0], {%1,%2}, [%3];"
                 ::"r"(raddr), "r"(a), "r"(b), "r"(rbar)
                 : "memory");
}
__device__ __forceinline__ void sts_b32(uint32_t a, uint32_t x) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(x) : "memory");
}

// opaque register copy: the value can no longer be rematerialised from its defining expression
__device__ __forceinline__ uint32_t pin(uint32_t v) {
    uint32_t r;
    asm volatile("mov.u32 %0, %1;" : "=r"(r) : "r"(v));
    return r;
}

// tie key: lower wins.  (bit-reversed (k mod bs)) << 22 | (k div bs);  k < 0 (no valid point) never wins
__device__ __forceinline__ uint32_t tie_key(int k, int log2bs) {
    if (k < 0) return 0xffffffffu;
    const uint32_t low = (uint32_t)k & ((1u << log2bs) - 1u);
    const uint32_t rev = log2bs == 0 ? 0u : (__brev(low) >> (32 - log2bs));
    return (rev << 22) | ((uint32_t)k >> log2bs);
}

// A candidate is 20 bytes in a 32-byte slot: {value bits, index, x, y} {z}.  The value is a
// non-negative float, so its bit pattern orders like the value.
constexpr int CAND_BYTES = 32;
constexpr int CAND_TX = 24;      // bytes actually transferred per candidate: {v,k,x,y} {z,bound}

// argmax over the candidates held one per lane (lanes >= count hold v = 0, k = -1): returns the
// lane of the winner.  The tie key is only evaluated when two candidates share the maximum.
__device__ __forceinline__ int cand_argmax(uint32_t v, int k, int log2bs) {
    const uint32_t gv = __reduce_max_sync(0xffffffffu, v);
    const uint32_t gb = __ballot_sync(0xffffffffu, v == gv);
    if ((gb & (gb - 1)) == 0) return __ffs(gb) - 1;
    const uint32_t tb = v == gv ? tie_key(k, log2bs) : 0xffffffffu;
    const uint32_t gt = __reduce_min_sync(0xffffffffu, tb);
    return __ffs(__ballot_sync(0xffffffffu, v == gv && tb == gt)) - 1;
}

// largest and second-largest of N non-negative values (second = 0 when N == 1), as a merge tree
template <int N>
__device__ __forceinline__ void top2(const float (&t)[N], float &v1, float &v2) {
    float a[N], b[N];
#pragma unroll
    for (int i = 0; i < N; ++i) { a[i] = t[i]; b[i] = 0.f; }
#pragma unroll
    for (int w = N; w > 1; w = (w + 1) / 2) {
#pragma unroll
        for (int i = 0; i < w / 2; ++i) {
            const int o = w - 1 - i;
            const float hi = fmaxf(a[i], a[o]), lo = fminf(a[i], a[o]);
            b[i] = fmaxf(lo, fmaxf(b[i], b[o]));
            a[i] = hi;
        }
    }
    v1 = a[0];
    v2 = b[0];
}

template <int N>
__device__ __forceinline__ float tree_max(const float (&t)[N]) {
    float a[N];
#pragma unroll
    for (int i = 0; i < N; ++i) a[i] = t[i];
#pragma unroll
    for (int w = N; w > 1; w = (w + 1) / 2) {
#pragma unroll
        for (int i = 0; i < w / 2; ++i) a[i] = fmaxf(a[i], a[w - 1 - i]);
    }
    return a[0];
}

// One round costs one dependent chain; everything below is arranged to keep that chain short
// (profiles/r01_fps_before.md: 1970 cycles/round, of which the exchange itself was ~450):
//   * the per-thread maximum is a tree, the slot holding it is recovered afterwards from a
//     bit mask of independent compares (no serial compare/select chain over the PPT slots);
//   * tie keys are evaluated only when two maxima are bit-equal (never, on real scenes);
//   * shared-memory addresses are plain 32-bit registers computed before the loop.
// SX = false: every thread keeps its points' coordinates in registers (the fast layout).  SX = true: they
// stay in shared memory as three planes (12 bytes per point, up to ~13 000 points per CTA) and only the
// running distances live in registers — for scenes of 100k - 210k points, which no longer fit the register
// file of a 16-CTA cluster.
template <int CS, int PPT, int NW, bool SX>
__global__ void __launch_bounds__(NW * 32)
fps_cluster_kernel(int n, int m, int log2bs, const float *__restrict__ xyz, float *__restrict__ temp,
                   int *__restrict__ idxs, int ibase, int istride, int dbg) {
    constexpr int NT = NW * 32;
    constexpr int NC = CS * NW;              // candidates per round: one per warp of the cluster
    constexpr int CPL = (NC + 31) / 32;      // candidates per lane
    static_assert(NC <= 256, "too many candidates");
    extern __shared__ float4 s_pts[];                              // [PPT][NT] copy of this CTA's points
    float *s_soa = reinterpret_cast<float *>(s_pts);               // SX: planes x | y | z, each [PPT][NT]
    constexpr int PL = PPT * NT;                                   // plane length
    __shared__ __align__(16) unsigned char s_warp[2][NW][CAND_BYTES];   // per-warp candidates
    __shared__ __align__(16) unsigned char s_exch[2][NC][CAND_BYTES];   // every warp's candidate, whole cluster
    __shared__ __align__(8) uint64_t s_bar[2];

    const int tid = (int)pin(threadIdx.x), lane = tid & 31, warp = tid >> 5;
    const int rank = (int)cluster_ctarank();
    const int batch = blockIdx.x / CS;
    xyz += 3ll * batch * n;
    temp += (long long)batch * n;
    idxs += (long long)batch * m;
    ibase += batch * istride;     // added to every index written (packed layout: global indices)

    // Point k of the scene belongs to CTA k % CS, warp (k / CS) % NW, lane (k / (CS*NW)) % 32, slot
    // k / (CS*NW*32): consecutive indices land in different warps.  PointNeXt runs FPS on clouds that
    // are already in FPS order (every level below the first), where the next picks ARE the next
    // indices — with a blocked layout they would all sit in one warp, its runner-up bound would
    // equal the next pick and every round would yield a single pick.
    const int kbase = rank + CS * (warp + NW * lane);     // index of slot 0
    constexpr int KSTRIDE = CS * NW * 32;                 // index distance between slots

    float x[SX ? 1 : PPT], y[SX ? 1 : PPT], z[SX ? 1 : PPT], t[PPT];
    bool any_valid = false;
#pragma unroll
    for (int s = 0; s < PPT; ++s) {
        const int k = kbase + s * KSTRIDE;
        // +inf coordinates give d = +inf and fminf(inf, 0) = 0: a padding slot stays at
        // distance 0 and is excluded from tie resolution below
        float px0 = INFINITY, py0 = INFINITY, pz0 = INFINITY;
        t[s] = 0.f;
        if (k < n) {
            px0 = __ldg(xyz + 3ll * k);
            py0 = __ldg(xyz + 3ll * k + 1);
            pz0 = __ldg(xyz + 3ll * k + 2);
            t[s] = temp[k];
            any_valid = true;
        }
        if (SX) {
            s_soa[s * NT + tid] = px0;
            s_soa[PL + s * NT + tid] = py0;
            s_soa[2 * PL + s * NT + tid] = pz0;
        } else {
            x[s] = px0; y[s] = py0; z[s] = pz0;
            s_pts[s * NT + tid] = make_float4(px0, py0, pz0, 0.f);
        }
    }
    const bool warp_valid = __any_sync(0xffffffffu, any_valid);

    if (tid == 0) {
        mbar_init(smem_u32(&s_bar[0]), 1);
        mbar_init(smem_u32(&s_bar[1]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // a warp without a single valid point publishes a losing candidate once and only keeps the barriers
    if (lane == 0) {
        sts_v4(smem_u32(&s_warp[0][warp][0]), 0u, 0xffffffffu, 0u, 0u);
        sts_v4(smem_u32(&s_warp[0][warp][16]), 0u, 0u, 0u, 0u);
        sts_v4(smem_u32(&s_warp[1][warp][0]), 0u, 0xffffffffu, 0u, 0u);
        sts_v4(smem_u32(&s_warp[1][warp][16]), 0u, 0u, 0u, 0u);
    }
    __syncthreads();
    cluster_sync_all();

    float x1 = __ldg(xyz), y1 = __ldg(xyz + 1), z1 = __ldg(xyz + 2);
    if (rank == 0 && tid == 0) idxs[0] = ibase;

    // addresses used every round, pinned in registers: without the opaque moves ptxas re-derives the
    // shared window base (S2UR SR_CgaCtaId + ULEA, ~30 cycles of scoreboard wait each) at every use
    const uint32_t a_pts = pin(smem_u32(&s_pts[tid]));
    const uint32_t a_warp0 = pin(smem_u32(&s_warp[0][0][0]));      // + par*NW*32 + w*32
    const uint32_t a_exch0 = pin(smem_u32(&s_exch[0][0][0]));      // + par*CS*32 + r*32
    const uint32_t lbar0 = pin(smem_u32(&s_bar[0])), lbar1 = pin(smem_u32(&s_bar[1]));
    const int mm = (int)pin((uint32_t)m);
    // warp 0, lane r delivers this CTA's candidate to CTA r of the cluster
    const uint32_t peer = lane < CS ? lane : 0;
    const uint32_t dst0 = mapa_u32(a_exch0 + rank * NW * CAND_BYTES, peer);
    const uint32_t dst1 = mapa_u32(a_exch0 + (NC + rank * NW) * CAND_BYTES, peer);
    const uint32_t rbar0 = mapa_u32(lbar0, peer), rbar1 = mapa_u32(lbar1, peer);

    // running-distance update of this thread's points with one pick
    auto apply_pick = [&](float qx, float qy, float qz) {
#pragma unroll
        for (int s = 0; s < PPT; ++s) {
            if (SX) {
                const float ax = s_soa[s * NT + tid], ay = s_soa[PL + s * NT + tid], az = s_soa[2 * PL + s * NT + tid];
                t[s] = fminf(dist2_ref(ax - qx, ay - qy, az - qz), t[s]);
            } else {
                t[s] = fminf(dist2_ref(x[s] - qx, y[s] - qy, z[s] - qz), t[s]);
            }
        }
    };
    // the first pick (index 0) is applied before the loop, so that t[] is always up to date
    // with every pick made so far when a round starts
    if (m > 1) apply_pick(x1, y1, z1);

    uint32_t par = 0, phase = 0;
    int j = 1, rounds = 0;
    while (j < m) {
        ++rounds;
        // ---- A. this thread's best and second-best, the warp's candidate and its runner-up bound ----
        const uint32_t my_warp_slot = a_warp0 + ((CS == 1 ? par * NW : 0) + warp) * CAND_BYTES;
        if (warp_valid) {
            float v1, v2;
            top2(t, v1, v2);
            const uint32_t vb = __float_as_uint(v1);
            const uint32_t wv = __reduce_max_sync(0xffffffffu, vb);
            const bool mine = vb == wv;
            uint32_t smask = 0;                                   // slots of this thread equal to the warp maximum
#pragma unroll
            for (int s = 0; s < PPT; ++s) smask |= (__float_as_uint(t[s]) == wv ? 1u : 0u) << s;
            const uint32_t bal = __ballot_sync(0xffffffffu, mine);
            const uint32_t multi = __ballot_sync(0xffffffffu, mine && (smask & (smask - 1)) != 0);
            int src = __ffs(bal) - 1;                            // the lane holding the warp's candidate
            int bslot = __ffs(smask) - 1;
            if (multi != 0 || (bal & (bal - 1)) != 0) {
                // exact tie resolution (rare): lowest tie key among all valid slots equal to the maximum
                uint32_t mytb = 0xffffffffu;
                if (mine) {
#pragma unroll
                    for (int s = 0; s < PPT; ++s) {
                        const int k = kbase + s * KSTRIDE;                 // padding slots never win a tie
                        const uint32_t tbs = k < n ? tie_key(k, log2bs) : 0xffffffffu;
                        if (((smask >> s) & 1u) && tbs < mytb) { mytb = tbs; bslot = s; }
                    }
                }
                const uint32_t wtb = __reduce_min_sync(0xffffffffu, mytb);
                src = __ffs(__ballot_sync(0xffffffffu, mine && mytb == wtb)) - 1;
                if (wtb == 0xffffffffu && lane == src) bslot = -1;          // only padding slots: losing candidate
            }
            // upper bound of every point of this warp other than the candidate
            const uint32_t sec = __reduce_max_sync(0xffffffffu, __float_as_uint(lane == src ? v2 : v1));
            if (lane == src) {
                if (bslot >= 0) {
                    uint4 cp;
                    if (SX) {
                        cp.x = __float_as_uint(s_soa[bslot * NT + tid]);
                        cp.y = __float_as_uint(s_soa[PL + bslot * NT + tid]);
                        cp.z = __float_as_uint(s_soa[2 * PL + bslot * NT + tid]);
                    } else {
                        cp = lds_v4(a_pts + bslot * (NT * 16));
                    }
                    sts_v4(my_warp_slot, wv, (uint32_t)(kbase + bslot * KSTRIDE), cp.x, cp.y);
                    sts_v2(my_warp_slot + 16, cp.z, sec);
                } else {
                    sts_v4(my_warp_slot, 0u, 0xffffffffu, 0u, 0u);
                    sts_v2(my_warp_slot + 16, 0u, sec);
                }
            }
        }
        // ---- candidates of the whole scene, CPL per lane: (value, index, xyz) and the bound U ----
        uint32_t cv[CPL], cu = 0;
        int ck[CPL];
        float cx[CPL], cy[CPL], cz[CPL];
        uint32_t a_c;
        if (CS == 1) {
            // single CTA: one barrier, then every warp works on the NW warp candidates itself
            __syncthreads();
            a_c = a_warp0 + par * NW * CAND_BYTES;
        } else {
            // every warp pushes its own candidate straight into every CTA of the cluster (st.async
            // completes a transaction on the receiver's mbarrier): lane r < CS delivers to CTA r
            const uint32_t lbar = par ? lbar1 : lbar0;
            if (tid == 0) mbar_expect_tx(lbar, NC * CAND_TX);
            __syncwarp();
            if (lane < CS) {
                const uint4 w = lds_v4(my_warp_slot);
                const uint2 w2 = lds_v2(my_warp_slot + 16);
                const uint32_t dst = (par ? dst1 : dst0) + warp * CAND_BYTES, rbar = par ? rbar1 : rbar0;
                st_async_v4(dst, rbar, w.x, w.y, w.z, w.w);
                st_async_v2(dst + 16, rbar, w2.x, w2.y);
            }
            mbar_wait(lbar, phase);
            a_c = a_exch0 + par * NC * CAND_BYTES;
        }
#pragma unroll
        for (int c = 0; c < CPL; ++c) {
            cv[c] = 0; ck[c] = -1; cx[c] = cy[c] = cz[c] = 0.f;
            if (c * 32 + lane < NC) {
                const uint4 q = lds_v4(a_c + (c * 32 + lane) * CAND_BYTES);
                const uint2 d = lds_v2(a_c + (c * 32 + lane) * CAND_BYTES + 16);
                cv[c] = q.x; ck[c] = (int)q.y; cx[c] = __uint_as_float(q.z); cy[c] = __uint_as_float(q.w);
                cz[c] = __uint_as_float(d.x);
                cu = max(cu, d.y);
            }
        }
        const uint32_t U = __reduce_max_sync(0xffffffffu, cu);   // no point outside the candidates exceeds U

        // ---- B. exact FPS on the candidate set for as long as its maximum provably beats every other
        //         point: the first pick is the global arg-max; a further pick is valid while its value
        //         is > U, because values only shrink.  The loop is one dependent chain per pick (arg-max
        //         -> winner's xyz -> candidates' new values -> next arg-max).  Applying a pick to this
        //         thread's own points is independent of that chain, so it is software-pipelined: the
        //         body applies the PREVIOUS pick, giving ptxas PPT independent FP32 chains to fill the
        //         chain's shuffle / vote / redux latencies with.  "No previous pick" is (inf,inf,inf),
        //         which leaves every running distance unchanged — no branch in the body.
        float px = INFINITY, py = INFINITY, pz = INFINITY;
        int q = 0;
        uint32_t bv = cv[0];
        int be = 0;
#pragma unroll
        for (int c = 1; c < CPL; ++c)
            if (cv[c] > bv) { bv = cv[c]; be = c; }
        uint32_t val = __reduce_max_sync(0xffffffffu, bv);
        while (true) {
            int neq = 0;                   // entries of this lane equal to the maximum
#pragma unroll
            for (int c = 0; c < CPL; ++c) neq += cv[c] == val ? 1 : 0;
            const uint32_t gb = __ballot_sync(0xffffffffu, neq > 0);
            const uint32_t g2 = __ballot_sync(0xffffffffu, neq > 1);
            int gl = __ffs(gb) - 1;
            if (((gb & (gb - 1)) | g2) != 0) {
                // equal maxima (rare): the reference's tie order decides
                uint32_t tb = 0xffffffffu;
#pragma unroll
                for (int c = 0; c < CPL; ++c) {
                    const uint32_t kc = cv[c] == val ? tie_key(ck[c], log2bs) : 0xffffffffu;
                    if (kc < tb) { tb = kc; be = c; }
                }
                const uint32_t gt = __reduce_min_sync(0xffffffffu, tb);
                gl = __ffs(__ballot_sync(0xffffffffu, tb == gt)) - 1;
            }
            float sx = cx[0], sy = cy[0], sz = cz[0];
            int sk = ck[0];
#pragma unroll
            for (int c = 1; c < CPL; ++c)
                if (be == c) { sx = cx[c]; sy = cy[c]; sz = cz[c]; sk = ck[c]; }
            x1 = __shfl_sync(0xffffffffu, sx, gl);
            y1 = __shfl_sync(0xffffffffu, sy, gl);
            z1 = __shfl_sync(0xffffffffu, sz, gl);
            if (rank == 0 && warp == 0 && lane == gl) idxs[j] = sk + ibase;
            // the previous pick, on this thread's points (independent of everything above and below)
            apply_pick(px, py, pz);
            ++q;
            ++j;
            // like the reference, the last pick of the call is never applied to temp
            const bool last = j >= mm;
            px = last ? INFINITY : x1;
            py = last ? INFINITY : y1;
            pz = last ? INFINITY : z1;
            // candidates' new values and the next arg-max
#pragma unroll
            for (int c = 0; c < CPL; ++c)
                cv[c] = min(cv[c], __float_as_uint(dist2_ref(cx[c] - x1, cy[c] - y1, cz[c] - z1)));
            bv = cv[0];
            be = 0;
#pragma unroll
            for (int c = 1; c < CPL; ++c)
                if (cv[c] > bv) { bv = cv[c]; be = c; }
            val = __reduce_max_sync(0xffffffffu, bv);
            if (last || !(val > U)) break;
        }
        // the round's final pick
        apply_pick(px, py, pz);
        par ^= 1;
        if (par == 0) phase ^= 1;
    }

    // the reference leaves the final running distances in temp
#pragma unroll
    for (int s = 0; s < PPT; ++s) {
        const int k = kbase + s * KSTRIDE;
        if (k < n) temp[k] = t[s];
    }
    if (dbg && rank == 0 && tid == 0) temp[0] = (float)rounds;   // AMC3D_FPS_DEBUG: exchange rounds instead of temp[0]
    cluster_sync_all();  // no CTA may exit while a peer can still address its shared memory
}

// ---- fallback for any n: one 1024-thread CTA per scene, state in global memory ------------
constexpr int FPS_G_THREADS = 1024;
__global__ void __launch_bounds__(FPS_G_THREADS)
fps_global_kernel(int n, int m, int log2bs, const float *__restrict__ xyz, float *__restrict__ temp,
                  int *__restrict__ idxs, int ibase, int istride) {
    __shared__ uint32_t s_v[32], s_tb[32];
    __shared__ int s_k[32];
    __shared__ int s_old;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    xyz += 3ll * blockIdx.x * n;
    temp += (long long)blockIdx.x * n;
    idxs += (long long)blockIdx.x * m;
    ibase += blockIdx.x * istride;
    int old = 0;
    if (tid == 0) idxs[0] = ibase;
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
        if (tid == 0) idxs[j] = old + ibase;
        __syncthreads();
    }
}

template <int CS, int PPT, int NW, bool SX = false>
static cudaError_t launch_cluster(int b, int n, int m, int log2bs, const float *xyz, float *temp, int *idx,
                                  int ibase, int istride, cudaStream_t st) {
    const size_t smem = (SX ? 3 * sizeof(float) : sizeof(float4)) * PPT * NW * 32;
    auto kern = fps_cluster_kernel<CS, PPT, NW, SX>;
    cudaError_t e;
    if (smem > 40 * 1024) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    if (CS > 8) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        if (e != cudaSuccess) return e;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(b * CS);
    cfg.blockDim = dim3(NW * 32);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CS;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    static const int dbg = getenv("AMC3D_FPS_DEBUG") != nullptr;
    return cudaLaunchKernelEx(&cfg, kern, n, m, log2bs, xyz, temp, idx, ibase, istride, dbg);
}

#define FPS_ARGS b, n, m, log2bs, xyz, temp, idx, ibase, istride, st
template <int CS, int NW>
static cudaError_t launch_for_ppt(int ppt, int b, int n, int m, int log2bs, const float *xyz, float *temp, int *idx,
                                  int ibase, int istride, cudaStream_t st) {
    if (ppt <= 1) return launch_cluster<CS, 1, NW>(FPS_ARGS);
    if (ppt <= 2) return launch_cluster<CS, 2, NW>(FPS_ARGS);
    if (ppt <= 3) return launch_cluster<CS, 3, NW>(FPS_ARGS);
    if (ppt <= 4) return launch_cluster<CS, 4, NW>(FPS_ARGS);
    if (ppt <= 6) return launch_cluster<CS, 6, NW>(FPS_ARGS);
    if (ppt <= 8) return launch_cluster<CS, 8, NW>(FPS_ARGS);
    if (ppt <= 12) return launch_cluster<CS, 12, NW>(FPS_ARGS);
    if (ppt <= 16) return launch_cluster<CS, 16, NW>(FPS_ARGS);
    if (NW <= 8) {     // 24+ points per thread only fit the register file with <= 256 threads
        if (ppt <= 24) return launch_cluster<CS, 24, (NW <= 8 ? NW : 8)>(FPS_ARGS);
        if (NW <= 4 && ppt <= 32) return launch_cluster<CS, 32, (NW <= 4 ? NW : 4)>(FPS_ARGS);
    }
    if (NW == 16 && CS == 16) {   // coordinates in shared memory: up to 16 x 512 x 26 = 212 992 points per scene
        if (ppt <= 20) return launch_cluster<16, 20, 16, true>(FPS_ARGS);
        if (ppt <= 26) return launch_cluster<16, 26, 16, true>(FPS_ARGS);
    }
    return cudaErrorInvalidConfiguration;
}

template <int CS>
static cudaError_t launch_for_nw(int nw, int b, int n, int m, int log2bs, const float *xyz, float *temp, int *idx,
                                 int ibase, int istride, cudaStream_t st) {
    const int ppt = div_up(n, CS * nw * 32);
    if (nw == 4) return launch_for_ppt<CS, 4>(ppt, FPS_ARGS);
    if (nw == 8) return launch_for_ppt<CS, 8>(ppt, FPS_ARGS);
    return launch_for_ppt<CS, 16>(ppt, FPS_ARGS);
}

static cudaError_t launch_for_cs(int cs, int nw, int b, int n, int m, int log2bs, const float *xyz, float *temp,
                                 int *idx, int ibase, int istride, cudaStream_t st) {
    if (cs == 16) return launch_for_nw<16>(nw, FPS_ARGS);
    if (cs == 8) return launch_for_nw<8>(nw, FPS_ARGS);
    if (cs == 4) return launch_for_nw<4>(nw, FPS_ARGS);
    return launch_for_nw<1>(nw, FPS_ARGS);
}


// b clouds of n points each, contiguous; log2bs = log2 of the reference's block size (it fixes the tie order);
// every index written is local index + ibase + cloud * istride.  Returns 0 or a cudaError_t.
static int env_int(const char *name, int dflt) {
    const char *e = getenv(name);
    return e ? atoi(e) : dflt;
}

int fps_culled_launch(int b, int n, int m, int log2bs, const float *xyz, float *temp, int *idx, int ibase, int istride,
                      cudaStream_t st);   // knn_grid.cu: exact spatial culling over sorted tiles

int fps_launch(int b, int n, int m, int log2bs, const float *xyz, float *temp, int *idx, int ibase, int istride,
               cudaStream_t st) {
    // Scenes beyond the cluster kernels (n > 212 992: they keep every point in registers / shared memory) take the
    // spatially culled kernel — ~4 tile updates per pick whatever n, 2 - 3 us per pick — instead of the all-points
    // global kernel at the end of this function.  Below that size the cluster kernels are faster (measured:
    // profiles/r02_fps.md).  AMC3D_FPS_CULLED_MIN moves the crossover.
    static const int culled_min = env_int("AMC3D_FPS_CULLED_MIN", 212993);
    if (n >= culled_min && m >= 64) {
        const int rc = fps_culled_launch(b, n, m, log2bs, xyz, temp, idx, ibase, istride, st);
        if (rc == 0) return 0;
        cudaGetLastError();                    // not applicable / failed: the cluster path decides
    }
    // cluster size and warps per CTA: as much parallelism as pays off (each round costs one
    // exchange regardless); AMC3D_FPS_CS / AMC3D_FPS_NW override the choice for experiments
    static const int env_cs = env_int("AMC3D_FPS_CS", 0), env_nw = env_int("AMC3D_FPS_NW", 0);
    // (measured, tools/prof_fps.py: 4 warps beat 8 at equal n — the exchange cost grows with the number
    // of sending warps — so more warps only when the points no longer fit 4 warps' registers)
    int cs = n > 1536 ? 16 : (n > 768 ? 8 : (n > 192 ? 4 : 1));
    int nw = 4;
    if (div_up(n, cs * nw * 32) > 28) {            // measured crossover (tools/diag/fps_nw.py): 40 000 pts 0.29 vs 0.35 us/pick,
                                                   // 64 000 pts 0.40 vs 0.38 us/pick for 4 vs 8 warps
        nw = 8;
        if (div_up(n, cs * nw * 32) > 24) nw = 16;
    }
    if (env_cs == 1 || env_cs == 4 || env_cs == 8 || env_cs == 16) cs = env_cs;
    if (env_nw == 4 || env_nw == 8 || env_nw == 16) nw = env_nw;
    cudaError_t e = launch_for_cs(cs, nw, FPS_ARGS);
    if (e != cudaSuccess && cs == 16) {        // non-portable cluster size refused: try 8 CTAs
        cudaGetLastError();
        e = launch_for_cs(8, 8, FPS_ARGS);
    }
    if (e != cudaSuccess) {                    // very large scenes (or clusters unavailable)
        cudaGetLastError();
        fps_global_kernel<<<b, FPS_G_THREADS, 0, st>>>(n, m, log2bs, xyz, temp, idx, ibase, istride);
    }
    return 0;
}

// `count` consecutive segments of the packed layout, all of n points and m samples (pointops_packed.cu)
int fps_segments(int count, int n, int m, int log2bs, const float *xyz, float *temp, int *idx, int idx_base,
                 cudaStream_t st) {
    if (n > (1 << 22)) {
        set_error("pointops_furthestsampling: segment of %d points > 4194304", n);
        return AMC3D_ELIMIT;
    }
    return fps_launch(count, n, m, log2bs, xyz, temp, idx, idx_base, n, st);
}

}  // namespace amc3d

using namespace amc3d;

extern "C" int amc3d_furthest_point_sampling(int b, int n, int m, const float *xyz, float *temp,
                                             int *idx, void *stream) {
    AMC3D_REQUIRE(b >= 0 && n >= 1 && m >= 0, AMC3D_EINVAL, "furthest_point_sampling: bad sizes b=%d n=%d m=%d", b, n, m);
    AMC3D_REQUIRE(n <= (1 << 22), AMC3D_ELIMIT, "furthest_point_sampling: n=%d > 4194304", n);
    if (b == 0 || m <= 0) return 0;  // reference kernel returns immediately for m <= 0
    // reference block size: largest power of two <= n, capped at 1024 (cuda_utils.h:10-14)
    int log2bs = 0;
    while ((2 << log2bs) <= n && log2bs < 10) ++log2bs;
    fps_launch(b, n, m, log2bs, xyz, temp, idx, 0, 0, as_stream(stream));
    return check_launch("furthest_point_sampling");
}
