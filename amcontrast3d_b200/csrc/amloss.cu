// Adaptive-margin contrastive loss of AMContrast3D, one decoder stage, for sm_100a.
//
// Replaces ~70 ATen kernels per stage plus an O(#boundary points) Python loop:
//   openpoints/AMContrast3D/MarginContrast.py:220-259 (point_contrast_margin),
//   :77-79 (dist_cos), :117-174 (contrast_softnn_margin), :111-115 (posmask_cnt),
//   AEF/utils.py:11-43 (get_subscene_label_CBL), AEF/ambiguity.py:11-93 (ambiguity_function),
//   AEF/function.py:10-39 (inverse_sigmoid_function, square_distance).
// Semantics: SURVEY.md App. A.4.  Nothing of size [m,k,D] or [m,k,ncls] is materialised:
// per point the pipeline keeps the kNN row (4k B), k-1 posmask bits, a count and `a`.
//
// Kernels (all HBM/L2-bound; DESIGN.md "AM loss"):
//   stage_labels   thread/point   integer kNN vote (first-max class)
//   posmask_count  thread/point   label compare -> bitmask, count, global max (atomicMax)
//   ambiguity      thread/point   |cnt-mx|/mx, boundary: n+-, d+- in the reference's FP32
//                                 order, a = 1/(1+pow(e,beta*(cc+ - cc-))), selection stats
//   row_inv_norm   warp/row       1/max(||f||,1e-8)
//   amloss_forward G lanes/anchor cosine to k-1 neighbours with 128-bit coalesced row loads,
//                                 margin, /T, exp, sums, -log, and the analytic backward:
//                                 anchor-role gradient by a second pass over the (L1/L2
//                                 resident) neighbour rows, neighbour-role gradient by
//                                 vector red.global.add.v4.f32 into ghat
//   amloss_reduce  1 CTA          deterministic double-precision mean over selected points
//   amloss_backward G lanes/row   projection through the normalisation, scaled by
//                                 upstream/|sel| read from device memory (no host sync)
#include "common.cuh"
#include <stdlib.h>

namespace amc3d {

// ---------------------------------------------------------------------------------------
// stage labels
// ---------------------------------------------------------------------------------------
constexpr int LBL_THREADS = 128;

__device__ __forceinline__ int map_label(long long t, int ncls, int has_ignore, long long ignore_index) {
    return (has_ignore && t == ignore_index) ? ncls - 1 : (int)t;
}

__global__ void __launch_bounds__(LBL_THREADS)
stage_labels_kernel(int m, int kr, int ncls, int has_ignore, long long ignore_index,
                    const long long *__restrict__ target, const int *__restrict__ nidx,
                    int *__restrict__ cls) {
    extern __shared__ unsigned short s_cnt[];  // [ncls][LBL_THREADS]
    const int tid = threadIdx.x;
    const int i = blockIdx.x * LBL_THREADS + tid;
    if (kr == 0) {
        if (i < m) cls[i] = map_label(__ldg(target + i), ncls, has_ignore, ignore_index);
        return;
    }
    for (int c = 0; c < ncls; ++c) s_cnt[c * LBL_THREADS + tid] = 0;
    if (i >= m) return;
    const int *row = nidx + (long long)i * kr;
    for (int j = 0; j < kr; ++j) {
        const int l = map_label(__ldg(target + __ldg(row + j)), ncls, has_ignore, ignore_index);
        if (l >= 0 && l < ncls) s_cnt[l * LBL_THREADS + tid] += 1;
    }
    int best = 0, bc = s_cnt[tid];
    for (int c = 1; c < ncls; ++c) {
        const int v = s_cnt[c * LBL_THREADS + tid];
        if (v > bc) { bc = v; best = c; }   // strict: first maximum wins (torch.argmax)
    }
    cls[i] = best;
}

// ---------------------------------------------------------------------------------------
// posmask + count + global max
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
posmask_count_kernel(int m, int ke, int ld, const int *__restrict__ nbr, const int *__restrict__ cls,
                     uint32_t *__restrict__ posbits, int *__restrict__ cnt, int *__restrict__ max_cnt) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    int c = 0;
    if (i < m) {
        const int ci = __ldg(cls + i);
        const int *row = nbr + (long long)i * ld;
        uint32_t bits = 0;
        for (int j = 0; j < ke; ++j)
            if (__ldg(cls + __ldg(row + j)) == ci) bits |= 1u << j;
        c = __popc(bits);
        posbits[i] = bits;
        cnt[i] = c;
    }
    const int wmax = __reduce_max_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0 && wmax > 0) atomicMax(max_cnt, wmax);
}

// ---------------------------------------------------------------------------------------
// ambiguity
// ---------------------------------------------------------------------------------------
// square_distance (AEF/function.py:36-38) for one pair.  The expression |a|^2 + |b|^2 - 2ab is ill-conditioned
// in FP32 for close points (d2 ~ 1e-4 from terms ~ 50), so its value — and through cc = n/d the ambiguity —
// depends on the rounding order of the torch backend the reference happens to run on, and on CUDA even on the
// number nb of boundary points, because cuBLAS switches kernels with the batch count of the
// [nb,1,3] x [nb,3,k] product.  All variants are reproduced (V = variant):
//   V 1,2: torch CUDA — what the reference computes where it actually runs; verified bit for bit against
//          torch 2.11 / cuBLAS 12.8 on a B200 for nb = 1 ... 82 307, k = 4 ... 32 (tools/diag/probe*.py):
//       V 1 (nb >= 153):  dot = fl(fma(y1,y2, fl(x1*x2)) + fl(z1*z2))
//       V 2 (nb <  153):  dot = (fl(z1*z2) + fl(x1*x2)) + fl(y1*y2)
//       |p|2 = (x*x + z*z) + y*y     (torch.sum(p ** 2, -1); x*x + (y*y + z*z) for the anchors when nb <= 2)
//   V 0: torch CPU — what tests/golden/loss_golden.npz was generated with:
//       dot  = (fl(x1*x2) + fl(y1*y2)) + fl(z1*z2) ;  |p|2 = (x*x + y*y) + z*z
//   then, in all:  dist = -2*dot ; dist += |p1|2 ; dist += |p2|2
__device__ __forceinline__ float sqnorm_torch(int v, float x, float y, float z) {
    if (v == 0) return __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
    if (v == 3) return __fadd_rn(__fmul_rn(x, x), __fadd_rn(__fmul_rn(y, y), __fmul_rn(z, z)));
    return __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(z, z)), __fmul_rn(y, y));
}

__device__ __forceinline__ float sqdist_torch(int v, float x1, float y1, float z1, float n1, float x2,
                                              float y2, float z2) {
    float dot;
    if (v == 1) dot = __fadd_rn(__fmaf_rn(y1, y2, __fmul_rn(x1, x2)), __fmul_rn(z1, z2));
    else if (v == 2) dot = __fadd_rn(__fadd_rn(__fmul_rn(z1, z2), __fmul_rn(x1, x2)), __fmul_rn(y1, y2));
    else dot = __fadd_rn(__fadd_rn(__fmul_rn(x1, x2), __fmul_rn(y1, y2)), __fmul_rn(z1, z2));
    const float n2 = sqnorm_torch(v, x2, y2, z2);
    float d = __fmul_rn(-2.f, dot);
    d = __fadd_rn(d, n1);
    d = __fadd_rn(d, n2);
    return d;
}

// #boundary points (0 < cnt < max cnt) -> stats[1]; the ambiguity kernel needs the total before it starts
__global__ void __launch_bounds__(256)
boundary_count_kernel(int m, const int *__restrict__ cnt, const int *__restrict__ max_cnt, int *__restrict__ stats) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    const int mx = __ldg(max_cnt);
    const int c = i < m ? __ldg(cnt + i) : 0;
    const unsigned mb = __ballot_sync(0xffffffffu, c > 0 && c < mx);
    if ((threadIdx.x & 31) == 0 && mb) atomicAdd(stats + 1, __popc(mb));
}

template <int BACKEND>
__global__ void __launch_bounds__(256)
ambiguity_kernel(int m, int ke, int ld, const float *__restrict__ p, const int *__restrict__ nbr,
                 const uint32_t *__restrict__ posbits, const int *__restrict__ cnt,
                 const int *__restrict__ max_cnt, int cctype, float beta, float nu_m,
                 float *__restrict__ a_out, int *__restrict__ stats) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    bool sel = false, boundary = false;
    int bin = -1;
    // the rounding variant (see sqdist_torch): uniform over the grid
    int v = 0, v1 = 0;
    if (BACKEND == 1) {
        const int nb = stats[1];
        v = nb >= 153 ? 1 : 2;
        v1 = nb <= 2 ? 3 : v;
    }
    if (i < m) {
        const int mx = __ldg(max_cnt);
        const int c = __ldg(cnt + i);
        // |cnt - mx| / mx, int64 -> float32 true division (ambiguity.py:14)
        float a = __fdiv_rn((float)abs(c - mx), (float)mx);
        boundary = c > 0 && c < mx;
        if (boundary) {
            float dpos, dneg;
            if (cctype == 1) {
                dpos = dneg = 5.0f;
            } else {
                const uint32_t bits = __ldg(posbits + i);
                const float x1 = __ldg(p + 3ll * i), y1 = __ldg(p + 3ll * i + 1), z1 = __ldg(p + 3ll * i + 2);
                const float n1 = sqnorm_torch(v1, x1, y1, z1);
                const int *row = nbr + (long long)i * ld;
                dpos = 0.f; dneg = 0.f;
                for (int j = 0; j < ke; ++j) {
                    const long long nj = __ldg(row + j);
                    float dd = sqdist_torch(v, x1, y1, z1, n1, __ldg(p + 3 * nj), __ldg(p + 3 * nj + 1),
                                            __ldg(p + 3 * nj + 2));
                    if (cctype == 3) dd = __fsqrt_rn(__fadd_rn(fabsf(dd), 1e-12f));
                    if ((bits >> j) & 1u) dpos = __fadd_rn(dpos, dd);
                    else dneg = __fadd_rn(dneg, dd);
                }
            }
            const float ccp = __fdiv_rn((float)c, dpos);
            const float ccn = __fdiv_rn((float)(ke - c), dneg);
            const float e32 = 2.7182817459106445f;  // float32(math.e)
            a = __fdiv_rn(1.f, __fadd_rn(1.f, powf(e32, __fmul_rn(beta, __fsub_rn(ccp, ccn)))));
        }
        a_out[i] = a;
        sel = (0.f < a) && (a <= 1.f);
        // the reference's five report bins (ambiguity.py:79-89)
        const float c10 = ceilf(__fmul_rn(a, 10.f));
        if (a == 0.f) bin = 0;
        else if (0.f < c10 && c10 < nu_m) bin = 1;
        else if (c10 == nu_m) bin = 2;
        else if (nu_m < c10 && c10 < 10.f) bin = 3;
        else if (c10 == 10.f) bin = 4;
    }
    // warp-aggregated counters
    const unsigned ms = __ballot_sync(0xffffffffu, sel);
    unsigned mbin[5];
#pragma unroll
    for (int t = 0; t < 5; ++t) mbin[t] = __ballot_sync(0xffffffffu, bin == t);
    if ((threadIdx.x & 31) == 0) {
        if (ms) atomicAdd(stats + 0, __popc(ms));
#pragma unroll
        for (int t = 0; t < 5; ++t)
            if (mbin[t]) atomicAdd(stats + 2 + t, __popc(mbin[t]));
    }
}

// ---------------------------------------------------------------------------------------
// row norms
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
row_inv_norm_kernel(int m, int d, const float *__restrict__ f, float *__restrict__ inv) {
    const int lane = threadIdx.x & 31;
    const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (r >= m) return;
    const float *row = f + (long long)r * d;
    float s = 0.f;
    if ((d & 3) == 0) {
        for (int c = lane; c < d / 4; c += 32) {
            const float4 v = __ldg(reinterpret_cast<const float4 *>(row) + c);
            s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
        }
    } else {
        for (int c = lane; c < d; c += 32) { const float v = __ldg(row + c); s += v * v; }
    }
    s = warp_sum(s);
    if (lane == 0) inv[r] = 1.f / fmaxf(sqrtf(s), 1e-8f);
}

// ---------------------------------------------------------------------------------------
// fused loss forward + gradient accumulation.  G lanes cooperate on one anchor; each lane
// owns V float4 chunks of the D-float row: chunk index = v*G + g.  D = 4*G*V.
// ---------------------------------------------------------------------------------------
constexpr int KE_MAX = 32;

template <int G>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ void red_add_v4(float *p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d)
                 : "memory");
}

// KEC > 0 (V == 1 only): the ke <= KEC neighbour chunks a lane reads for the cosines stay in registers (4 * KEC
// floats) and feed the gradient pass, so every neighbour row is fetched ONCE; KEC == 0 re-reads them in pass 2.
template <int G, int V, int KEC = 0>
__global__ void __launch_bounds__(256)
amloss_forward_kernel(int m, int ke, int ld, const float *__restrict__ f, const float *__restrict__ inv,
                      const int *__restrict__ nbr, const uint32_t *__restrict__ posbits,
                      const float *__restrict__ a, amc3d_loss_params prm, float *__restrict__ loss_pt,
                      float *__restrict__ ghat, const int *__restrict__ order) {
    constexpr int D4 = G * V;   // float4 chunks per row
    constexpr int D = D4 * 4;
    const int g = threadIdx.x % G;
    const long long slot = ((long long)blockIdx.x * 256 + threadIdx.x) / G;
    // all lanes of a group share i; whole groups exit together.  Shuffles below use the full
    // mask, so a partially filled last warp keeps its idle groups alive until the end.
    // anchors are visited in `order` (spatially coherent): neighbouring groups then gather the same rows.
    // A negative entry is an empty slot: callers may pass the order with the unselected anchors removed and
    // the tail filled with -1 (loss_pt pre-zeroed), so that every live warp is full of selected anchors.
    bool in_range = slot < m;
    long long i = slot;
    if (in_range && order != nullptr) {
        i = __ldg(order + slot);
        in_range = i >= 0;
    }
    const float ai = in_range ? __ldg(a + i) : 0.f;
    const bool sel = in_range && (0.f < ai) && (ai <= 1.f);
    if (in_range && !sel && g == 0) loss_pt[i] = 0.f;
    if (__ballot_sync(0xffffffffu, sel) == 0) return;

    const long long ii = sel ? i : 0;
    const uint32_t bits = sel ? __ldg(posbits + ii) : 0u;
    const float inv_i = __ldg(inv + ii);
    const int *row = nbr + ii * ld;

    float4 fi[V];
#pragma unroll
    for (int v = 0; v < V; ++v) fi[v] = __ldg(reinterpret_cast<const float4 *>(f + ii * D) + v * G + g);

    // pass 1: cosine similarities
    constexpr int KEL = KEC > 0 ? KEC : KE_MAX;       // compile-time bound of the neighbour loops
    float s[KE_MAX];
    float4 wj[KEC > 0 ? KEC : 1];
#pragma unroll
    for (int j = 0; j < KE_MAX; ++j) s[j] = 0.f;
#pragma unroll
    for (int j = 0; j < KEL; ++j) {
        if (j < ke) {
            const long long nj = __ldg(row + j);
            const float4 *fr = reinterpret_cast<const float4 *>(f + nj * D);
            float acc = 0.f;
#pragma unroll
            for (int v = 0; v < V; ++v) {
                const float4 w = __ldg(fr + v * G + g);
                if (KEC > 0) wj[j] = w;
                acc += fi[v].x * w.x + fi[v].y * w.y + fi[v].z * w.z + fi[v].w * w.w;
            }
            acc = group_sum<G>(acc);
            s[j] = acc * inv_i * __ldg(inv + nj);
        }
    }

    // margin, temperature, exp, sums (every lane of the group computes the same scalars)
    const float margin = prm.margin_mode == 1 ? __fadd_rn(__fmul_rn(prm.mu, ai), prm.nu) : prm.nu;
    float e[KE_MAX];
    float P = 0.f, S = 0.f;
    int cnt = 0;
#pragma unroll
    for (int j = 0; j < KE_MAX; ++j) {
        e[j] = 0.f;
        if (j < ke) {
            const bool pos = (bits >> j) & 1u;
            float z = s[j];
            if (prm.db_mode == 1 && pos) z = __fsub_rn(z, margin);
            if (prm.db_mode == 2 && !pos) z = __fadd_rn(z, margin);
            if (prm.has_temperature) z = __fdiv_rn(z, prm.temperature);
            e[j] = expf(z);
            S += e[j];
            if (pos) { P += e[j]; ++cnt; }
        }
    }
    const float invT = prm.has_temperature ? 1.f / prm.temperature : 1.f;
    const float eps = 1e-12f;
    float li;
    float gj[KE_MAX];   // dL_i/ds_ij
    if (prm.cl_method == 1) {
        const float ratio = P / S;
        const float r = ratio + eps;
        li = -logf(r);
        const float c0 = -invT / (r * S);
#pragma unroll
        for (int j = 0; j < KE_MAX; ++j) {
            const float pj = ((bits >> j) & 1u) ? 1.f : 0.f;
            gj[j] = j < ke ? c0 * e[j] * (pj - ratio) : 0.f;
        }
    } else {
        const float Nn = S - P;
        float q = 0.f, w = 0.f;   // q = sum_j loss_ij, w = sum_{pos} e_j/(e_j+Nn)^2
#pragma unroll
        for (int j = 0; j < KE_MAX; ++j) {
            if (j < ke) {
                const bool pos = (bits >> j) & 1u;
                const float pe = pos ? e[j] : 0.f;
                const float den = pe + Nn;
                q += pe / den + eps;
                if (pos) w += e[j] / (den * den);
            }
        }
        const float pn = (float)cnt + eps;
        const float r = q / pn;
        li = -logf(r);
        const float c0 = -invT / (r * pn);
#pragma unroll
        for (int j = 0; j < KE_MAX; ++j) {
            if (j < ke) {
                const bool pos = (bits >> j) & 1u;
                const float den = e[j] + Nn;
                gj[j] = pos ? c0 * e[j] * Nn / (den * den) : -c0 * e[j] * w;
            } else {
                gj[j] = 0.f;
            }
        }
    }
    if (sel && g == 0) loss_pt[i] = li;

    // pass 2: gradients w.r.t. the normalised rows u = f*inv.
    //   anchor role:     ghat[i]  += sum_j gj * u_j      (accumulated in registers)
    //   neighbour role:  ghat[nj] += gj * u_i            (vector reduction into L2)
    float4 acc[V];
#pragma unroll
    for (int v = 0; v < V; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int j = 0; j < KEL; ++j) {
        if (j < ke) {
            const long long nj = __ldg(row + j);
            const float gu = gj[j] * __ldg(inv + nj);   // gj * inv_j  (u_j = f_j * inv_j)
            const float gi = gj[j] * inv_i;             // gj * inv_i  (u_i = f_i * inv_i)
            const float4 *fr = reinterpret_cast<const float4 *>(f + nj * D);
            float *gr = ghat + nj * D;
#pragma unroll
            for (int v = 0; v < V; ++v) {
                const float4 w = KEC > 0 ? wj[j] : __ldg(fr + v * G + g);
                acc[v].x += gu * w.x; acc[v].y += gu * w.y; acc[v].z += gu * w.z; acc[v].w += gu * w.w;
                if (sel)
                    red_add_v4(gr + (v * G + g) * 4, gi * fi[v].x, gi * fi[v].y, gi * fi[v].z, gi * fi[v].w);
            }
        }
    }
    if (sel) {
        float *gr = ghat + i * D;
#pragma unroll
        for (int v = 0; v < V; ++v)
            red_add_v4(gr + (v * G + g) * 4, acc[v].x, acc[v].y, acc[v].z, acc[v].w);
    }
}

// generic-D fallback (D not one of the instantiated sizes): one warp per anchor, scalar lanes
__global__ void __launch_bounds__(256)
amloss_forward_generic_kernel(int m, int d, int ke, int ld, const float *__restrict__ f, const float *__restrict__ inv,
                              const int *__restrict__ nbr, const uint32_t *__restrict__ posbits,
                              const float *__restrict__ a, amc3d_loss_params prm,
                              float *__restrict__ loss_pt, float *__restrict__ ghat) {
    const int lane = threadIdx.x & 31;
    const long long i = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (i >= m) return;
    const float ai = __ldg(a + i);
    const bool sel = (0.f < ai) && (ai <= 1.f);
    if (!sel) { if (lane == 0) loss_pt[i] = 0.f; return; }
    const uint32_t bits = __ldg(posbits + i);
    const float inv_i = __ldg(inv + i);
    const int *row = nbr + i * ld;
    const float *fi = f + i * d;
    const float margin = prm.margin_mode == 1 ? __fadd_rn(__fmul_rn(prm.mu, ai), prm.nu) : prm.nu;
    const float invT = prm.has_temperature ? 1.f / prm.temperature : 1.f;
    const float eps = 1e-12f;

    float e[KE_MAX];
    float P = 0.f, S = 0.f;
    int cnt = 0;
#pragma unroll
    for (int j = 0; j < KE_MAX; ++j) {
        e[j] = 0.f;
        if (j < ke) {
            const long long nj = __ldg(row + j);
            const float *fj = f + nj * d;
            float acc = 0.f;
            for (int c = lane; c < d; c += 32) acc += __ldg(fi + c) * __ldg(fj + c);
            acc = warp_sum(acc);
            const bool pos = (bits >> j) & 1u;
            float z = acc * inv_i * __ldg(inv + nj);
            if (prm.db_mode == 1 && pos) z = __fsub_rn(z, margin);
            if (prm.db_mode == 2 && !pos) z = __fadd_rn(z, margin);
            if (prm.has_temperature) z = __fdiv_rn(z, prm.temperature);
            e[j] = expf(z);
            S += e[j];
            if (pos) { P += e[j]; ++cnt; }
        }
    }
    float li, c0, ratio = 0.f, Nn = 0.f, w = 0.f;
    if (prm.cl_method == 1) {
        ratio = P / S;
        const float r = ratio + eps;
        li = -logf(r);
        c0 = -invT / (r * S);
    } else {
        Nn = S - P;
        float q = 0.f;
#pragma unroll
        for (int j = 0; j < KE_MAX; ++j)
            if (j < ke) {
                const bool pos = (bits >> j) & 1u;
                const float pe = pos ? e[j] : 0.f;
                const float den = pe + Nn;
                q += pe / den + eps;
                if (pos) w += e[j] / (den * den);
            }
        const float pn = (float)cnt + eps;
        const float r = q / pn;
        li = -logf(r);
        c0 = -invT / (r * pn);
    }
    if (lane == 0) loss_pt[i] = li;
#pragma unroll
    for (int j = 0; j < KE_MAX; ++j) {
        if (j < ke) {
            const bool pos = (bits >> j) & 1u;
            float gjv;
            if (prm.cl_method == 1) gjv = c0 * e[j] * ((pos ? 1.f : 0.f) - ratio);
            else {
                const float den = e[j] + Nn;
                gjv = pos ? c0 * e[j] * Nn / (den * den) : -c0 * e[j] * w;
            }
            const long long nj = __ldg(row + j);
            const float gu = gjv * __ldg(inv + nj), gi = gjv * inv_i;
            const float *fj = f + nj * d;
            for (int c = lane; c < d; c += 32) {
                atomicAdd(ghat + i * d + c, gu * __ldg(fj + c));
                atomicAdd(ghat + nj * d + c, gi * __ldg(fi + c));
            }
        }
    }
}

// deterministic mean over the selected points: loss_out[0] (+)= sum(loss_pt) / stats[0]
__global__ void __launch_bounds__(1024)
amloss_reduce_kernel(int m, const float *__restrict__ loss_pt, const int *__restrict__ stats,
                     float *__restrict__ loss_out) {
    __shared__ double s[32];
    double acc = 0.0;
    for (int i = threadIdx.x; i < m; i += 1024) acc += (double)loss_pt[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        acc = s[threadIdx.x];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (threadIdx.x == 0) loss_out[0] = (float)(acc / (double)stats[0]);  // 0/0 = NaN like torch.mean([])
    }
}

// grad_f[r] = scale * (ghat[r] - u_r (u_r . ghat[r])) * inv[r];  G lanes per row, D = 4*G*V
template <int G, int V, bool ACC>
__global__ void __launch_bounds__(256)
amloss_backward_kernel(int m, const float *__restrict__ f, const float *__restrict__ inv,
                       const float *__restrict__ ghat, const float *__restrict__ upstream,
                       const int *__restrict__ stats, float *__restrict__ grad_f) {
    constexpr int D = G * V * 4;
    const int g = threadIdx.x % G;
    const long long r = ((long long)blockIdx.x * 256 + threadIdx.x) / G;
    const bool ok = r < m;
    const long long rr = ok ? r : 0;
    const float scale = __ldg(upstream) / (float)__ldg(stats);
    const float iv = __ldg(inv + rr);
    float4 fv[V], gv[V];
    float dotp = 0.f;
#pragma unroll
    for (int v = 0; v < V; ++v) {
        fv[v] = __ldg(reinterpret_cast<const float4 *>(f + rr * D) + v * G + g);
        gv[v] = __ldg(reinterpret_cast<const float4 *>(ghat + rr * D) + v * G + g);
        dotp += fv[v].x * gv[v].x + fv[v].y * gv[v].y + fv[v].z * gv[v].z + fv[v].w * gv[v].w;
    }
    dotp = group_sum<G>(dotp) * iv;   // u . ghat
    // below the 1e-8 clamp the normalisation is a constant scale: no projection term
    const float proj = (iv >= 1e8f) ? 0.f : dotp * iv;
    if (!ok) return;
    const float sc = scale * iv;
#pragma unroll
    for (int v = 0; v < V; ++v) {
        float4 o;
        o.x = sc * (gv[v].x - proj * fv[v].x);
        o.y = sc * (gv[v].y - proj * fv[v].y);
        o.z = sc * (gv[v].z - proj * fv[v].z);
        o.w = sc * (gv[v].w - proj * fv[v].w);
        float4 *dst = reinterpret_cast<float4 *>(grad_f + r * D) + v * G + g;
        if (ACC) {
            const float4 old = *dst;
            o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
        }
        *dst = o;
    }
}

template <bool ACC>
__global__ void __launch_bounds__(256)
amloss_backward_generic_kernel(int m, int d, const float *__restrict__ f, const float *__restrict__ inv,
                               const float *__restrict__ ghat, const float *__restrict__ upstream,
                               const int *__restrict__ stats, float *__restrict__ grad_f) {
    const int lane = threadIdx.x & 31;
    const long long r = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (r >= m) return;
    const float scale = __ldg(upstream) / (float)__ldg(stats);
    const float iv = __ldg(inv + r);
    float dotp = 0.f;
    for (int c = lane; c < d; c += 32) dotp += __ldg(f + r * d + c) * __ldg(ghat + r * d + c);
    dotp = warp_sum(dotp) * iv;
    const float proj = (iv >= 1e8f) ? 0.f : dotp * iv;
    const float sc = scale * iv;
    for (int c = lane; c < d; c += 32) {
        const float o = sc * (__ldg(ghat + r * d + c) - proj * __ldg(f + r * d + c));
        if (ACC) grad_f[r * d + c] += o;
        else grad_f[r * d + c] = o;
    }
}

// per-class confusion counts of a prediction against the labels (what the trainer's ConfusionMatrix.update
// accumulates and all-reduces every step, main_AA.py:461,496-507): out[0..ncls) = tp, [ncls..2ncls) = #predicted,
// [2ncls..3ncls) = #labelled.  pred == NULL: the prediction is the label of the point's first listed neighbour
// (nbr[i*ld]) — the stand-in the path replay uses.  Block-local shared-memory histograms, one float atomic per
// class and block (exact below 2^24).
__global__ void __launch_bounds__(256)
class_counts_kernel(int m, int ncls, const int *__restrict__ target, const int *__restrict__ pred,
                    const int *__restrict__ nbr, int ld, float *__restrict__ out) {
    extern __shared__ int s_hist[];
    for (int i = threadIdx.x; i < 3 * ncls; i += 256) s_hist[i] = 0;
    __syncthreads();
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < m; i += (long long)gridDim.x * 256) {
        const int t = __ldg(target + i);
        const int p = pred != nullptr ? __ldg(pred + i) : __ldg(target + __ldg(nbr + i * ld));
        if (t >= 0 && t < ncls) {
            atomicAdd(&s_hist[2 * ncls + t], 1);
            if (p == t) atomicAdd(&s_hist[t], 1);
        }
        if (p >= 0 && p < ncls) atomicAdd(&s_hist[ncls + p], 1);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 3 * ncls; i += 256)
        if (s_hist[i]) atomicAdd(out + i, (float)s_hist[i]);
}

}  // namespace amc3d

using namespace amc3d;

extern "C" int amc3d_class_counts(int m, int ncls, const int *target, const int *pred, const int *nbr, int ld,
                                  float *out, void *stream) {
    AMC3D_REQUIRE(m >= 0 && ncls >= 1 && ncls <= 4096, AMC3D_EINVAL, "class_counts: bad sizes m=%d ncls=%d", m, ncls);
    AMC3D_REQUIRE(pred != nullptr || (nbr != nullptr && ld >= 1), AMC3D_EINVAL, "class_counts: neither pred nor nbr given");
    cudaStream_t st = as_stream(stream);
    cudaMemsetAsync(out, 0, sizeof(float) * 3 * (size_t)ncls, st);
    if (m > 0)
        class_counts_kernel<<<min(div_up(m, 256), 2 * kNumSMs), 256, sizeof(int) * 3 * ncls, st>>>(m, ncls, target, pred, nbr, ld, out);
    return check_launch("class_counts");
}

extern "C" int amc3d_stage_labels(int m, int kr, int ncls, int has_ignore, long long ignore_index,
                                  const long long *target, const int *nidx, int *cls, void *stream) {
    AMC3D_REQUIRE(m >= 0 && kr >= 0 && ncls >= 1, AMC3D_EINVAL, "stage_labels: bad sizes m=%d kr=%d ncls=%d", m, kr, ncls);
    AMC3D_REQUIRE(ncls <= 256, AMC3D_ELIMIT, "stage_labels: ncls=%d > 256", ncls);
    if (m == 0) return 0;
    const size_t smem = kr == 0 ? 0 : (size_t)ncls * LBL_THREADS * sizeof(unsigned short);
    if (smem > 48 * 1024)
        cudaFuncSetAttribute(stage_labels_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    stage_labels_kernel<<<div_up(m, LBL_THREADS), LBL_THREADS, smem, as_stream(stream)>>>(
        m, kr, ncls, has_ignore, ignore_index, target, nidx, cls);
    return check_launch("stage_labels");
}

extern "C" int amc3d_posmask_count(int m, int ke, int ld, const int *nbr, const int *cls, uint32_t *posbits,
                                   int *cnt, int *max_cnt, void *stream) {
    AMC3D_REQUIRE(m >= 0 && ke >= 1 && ld >= ke, AMC3D_EINVAL, "posmask_count: bad sizes m=%d ke=%d ld=%d", m, ke, ld);
    AMC3D_REQUIRE(ke <= KE_MAX, AMC3D_ELIMIT, "posmask_count: %d neighbours > 32", ke);
    if (m == 0) return 0;
    posmask_count_kernel<<<div_up(m, 256), 256, 0, as_stream(stream)>>>(m, ke, ld, nbr, cls, posbits, cnt, max_cnt);
    return check_launch("posmask_count");
}

extern "C" int amc3d_ambiguity(int m, int ke, int ld, const float *p, const int *nbr, const uint32_t *posbits,
                               const int *cnt, const int *max_cnt, int cctype, float beta, float nu,
                               float *a, int *stats, void *stream) {
    return amc3d_ambiguity_backend(m, ke, ld, p, nbr, posbits, cnt, max_cnt, cctype, beta, nu, 1, a, stats, stream);
}

extern "C" int amc3d_ambiguity_backend(int m, int ke, int ld, const float *p, const int *nbr,
                                       const uint32_t *posbits, const int *cnt, const int *max_cnt, int cctype,
                                       float beta, float nu, int torch_backend, float *a, int *stats,
                                       void *stream) {
    AMC3D_REQUIRE(torch_backend == 0 || torch_backend == 1, AMC3D_EINVAL, "ambiguity: torch_backend=%d not 0 (cpu) / 1 (cuda)", torch_backend);
    AMC3D_REQUIRE(m >= 0 && ke >= 1 && ke <= KE_MAX && ld >= ke, AMC3D_EINVAL, "ambiguity: bad sizes m=%d ke=%d ld=%d", m, ke, ld);
    AMC3D_REQUIRE(cctype >= 1 && cctype <= 3, AMC3D_EINVAL, "ambiguity: cctype=%d not in 1..3", cctype);
    if (m == 0) return 0;
    // nu_m = nu * 10 in Python double, compared against float32 tensors (ambiguity.py:77)
    const float nu_m = (float)((double)nu * 10.0);
    boundary_count_kernel<<<div_up(m, 256), 256, 0, as_stream(stream)>>>(m, cnt, max_cnt, stats);
    if (torch_backend == 1)
        ambiguity_kernel<1><<<div_up(m, 256), 256, 0, as_stream(stream)>>>(m, ke, ld, p, nbr, posbits, cnt, max_cnt,
                                                                            cctype, beta, nu_m, a, stats);
    else
        ambiguity_kernel<0><<<div_up(m, 256), 256, 0, as_stream(stream)>>>(m, ke, ld, p, nbr, posbits, cnt, max_cnt,
                                                                            cctype, beta, nu_m, a, stats);
    return check_launch("ambiguity");
}

extern "C" int amc3d_row_inv_norm(int m, int d, const float *f, float *inv, void *stream) {
    AMC3D_REQUIRE(m >= 0 && d >= 1, AMC3D_EINVAL, "row_inv_norm: bad sizes m=%d d=%d", m, d);
    if (m == 0) return 0;
    row_inv_norm_kernel<<<div_up(m, 8), 256, 0, as_stream(stream)>>>(m, d, f, inv);
    return check_launch("row_inv_norm");
}

template <int G, int V>
static void launch_fwd(int m, int ke, int ld, const float *f, const float *inv, const int *nbr,
                       const uint32_t *posbits, const float *a, const amc3d_loss_params &prm,
                       float *loss_pt, float *ghat, const int *order, cudaStream_t st) {
    const long long threads = (long long)m * G;
    const unsigned blocks = (unsigned)div_up_ll(threads, 256);
    // The register-cached variants (KEC > 0: every neighbour row fetched once) are built but off: measured at config 2
    // they LOSE, 0.524 ms against 0.424 ms for the four stages — the second read of a row hits L1 and is cheap,
    // 122 registers per thread instead of 65 halve the resident warps, and what bounds the kernel is the
    // neighbour-role reduction traffic into L2 (16 red.v4 per lane and anchor).  AMC3D_AMLOSS_RC=1 enables them.
    static const bool no_rc = getenv("AMC3D_AMLOSS_RC") == nullptr;
    if constexpr (V == 1) {
    if (!no_rc && ke <= 16) {
        amloss_forward_kernel<G, V, 16><<<blocks, 256, 0, st>>>(m, ke, ld, f, inv, nbr, posbits, a, prm, loss_pt, ghat, order);
        return;
    }
    if (!no_rc && ke <= 24) {
        amloss_forward_kernel<G, V, 24><<<blocks, 256, 0, st>>>(m, ke, ld, f, inv, nbr, posbits, a, prm, loss_pt, ghat, order);
        return;
    }
    }
    amloss_forward_kernel<G, V><<<blocks, 256, 0, st>>>(m, ke, ld, f, inv, nbr, posbits, a, prm, loss_pt, ghat, order);
}

extern "C" int amc3d_amloss_forward(int m, int d, int ke, int ld, const float *f, const float *inv,
                                    const int *nbr, const uint32_t *posbits, const float *a,
                                    const amc3d_loss_params *params, float *loss_pt, float *ghat,
                                    void *stream) {
    return amc3d_amloss_forward_order(m, d, ke, ld, f, inv, nbr, posbits, a, params, loss_pt, ghat, nullptr, stream);
}

extern "C" int amc3d_amloss_forward_order(int m, int d, int ke, int ld, const float *f, const float *inv,
                                          const int *nbr, const uint32_t *posbits, const float *a,
                                          const amc3d_loss_params *params, float *loss_pt, float *ghat,
                                          const int *order, void *stream) {
    AMC3D_REQUIRE(m >= 0 && d >= 1 && ke >= 1 && ke <= KE_MAX && ld >= ke, AMC3D_EINVAL, "amloss_forward: bad sizes m=%d d=%d ke=%d ld=%d", m, d, ke, ld);
    AMC3D_REQUIRE(params != nullptr, AMC3D_EINVAL, "amloss_forward: params is NULL");
    AMC3D_REQUIRE(params->cl_method == 1 || params->cl_method == 2, AMC3D_EINVAL, "amloss_forward: cl_method=%d", params->cl_method);
    if (m == 0) return 0;
    cudaStream_t st = as_stream(stream);
    const bool al = ((reinterpret_cast<uintptr_t>(f) | reinterpret_cast<uintptr_t>(ghat)) & 15) == 0;
    if (al && d == 32) launch_fwd<8, 1>(m, ke, ld, f, inv, nbr, posbits, a, *params, loss_pt, ghat, order, st);
    else if (al && d == 64) launch_fwd<16, 1>(m, ke, ld, f, inv, nbr, posbits, a, *params, loss_pt, ghat, order, st);
    else if (al && d == 128) launch_fwd<32, 1>(m, ke, ld, f, inv, nbr, posbits, a, *params, loss_pt, ghat, order, st);
    else if (al && d == 256) launch_fwd<32, 2>(m, ke, ld, f, inv, nbr, posbits, a, *params, loss_pt, ghat, order, st);
    else if (al && d == 512) launch_fwd<32, 4>(m, ke, ld, f, inv, nbr, posbits, a, *params, loss_pt, ghat, order, st);
    else
        amloss_forward_generic_kernel<<<div_up(m, 8), 256, 0, st>>>(m, d, ke, ld, f, inv, nbr, posbits, a, *params,
                                                                    loss_pt, ghat);
    return check_launch("amloss_forward");
}

extern "C" int amc3d_amloss_reduce(int m, const float *loss_pt, const int *stats, float *loss_out,
                                   void *stream) {
    AMC3D_REQUIRE(m >= 0, AMC3D_EINVAL, "amloss_reduce: bad size m=%d", m);
    amloss_reduce_kernel<<<1, 1024, 0, as_stream(stream)>>>(m, loss_pt, stats, loss_out);
    return check_launch("amloss_reduce");
}

template <int G, int V>
static void launch_bwd(int m, const float *f, const float *inv, const float *ghat, const float *upstream,
                       const int *stats, int accumulate, float *grad_f, cudaStream_t st) {
    const unsigned blocks = (unsigned)div_up_ll((long long)m * G, 256);
    if (accumulate) amloss_backward_kernel<G, V, true><<<blocks, 256, 0, st>>>(m, f, inv, ghat, upstream, stats, grad_f);
    else amloss_backward_kernel<G, V, false><<<blocks, 256, 0, st>>>(m, f, inv, ghat, upstream, stats, grad_f);
}

extern "C" int amc3d_amloss_backward(int m, int d, const float *f, const float *inv, const float *ghat,
                                     const float *upstream, const int *stats, int accumulate,
                                     float *grad_f, void *stream) {
    AMC3D_REQUIRE(m >= 0 && d >= 1, AMC3D_EINVAL, "amloss_backward: bad sizes m=%d d=%d", m, d);
    if (m == 0) return 0;
    cudaStream_t st = as_stream(stream);
    const bool al = ((reinterpret_cast<uintptr_t>(f) | reinterpret_cast<uintptr_t>(ghat) |
                      reinterpret_cast<uintptr_t>(grad_f)) & 15) == 0;
    if (al && d == 32) launch_bwd<8, 1>(m, f, inv, ghat, upstream, stats, accumulate, grad_f, st);
    else if (al && d == 64) launch_bwd<16, 1>(m, f, inv, ghat, upstream, stats, accumulate, grad_f, st);
    else if (al && d == 128) launch_bwd<32, 1>(m, f, inv, ghat, upstream, stats, accumulate, grad_f, st);
    else if (al && d == 256) launch_bwd<32, 2>(m, f, inv, ghat, upstream, stats, accumulate, grad_f, st);
    else if (al && d == 512) launch_bwd<32, 4>(m, f, inv, ghat, upstream, stats, accumulate, grad_f, st);
    else if (accumulate)
        amloss_backward_generic_kernel<true><<<div_up(m, 8), 256, 0, st>>>(m, d, f, inv, ghat, upstream, stats, grad_f);
    else
        amloss_backward_generic_kernel<false><<<div_up(m, 8), 256, 0, st>>>(m, d, f, inv, ghat, upstream, stats, grad_f);
    return check_launch("amloss_backward");
}
