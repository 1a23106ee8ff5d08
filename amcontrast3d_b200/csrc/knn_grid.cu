// Exact neighbour search with spatial culling: kNN for the single-segment case (offset = [n]) —
// the case AMContrast3D always produces (pointnext_AA.py:461 flattens the batch into one segment) —
// and, over the same structure built per batch element, three_nn (k = 3) and ball_query on batched
// (B,N,3) clouds.  Everything below is "batched": nb clouds of n points each, cloud b's sorted
// points living at [b*npad, b*npad + n) with npad = n rounded up to a whole tile, so that a tile
// never straddles two clouds.  kNN uses nb = 1.
//
// The brute-force kernel (knn.cu) evaluates all m*n pairs: 3.7e10 for the stage-0 self-kNN of
// BASELINE config 2, ~12 ms at the FP32 issue limit.  The result, however, is fully determined
// by the (d2, index)-lexicographic order, so any evaluation order that provably visits every
// point that can enter a query's top-k returns the identical idx / dist2.  This file does
// that (DESIGN.md "kNN: culled path"):
//
//   1. bbox            one pass, block reduce + ordered-int atomics
//   2. cell sort       points are binned into a 2^b x 2^b x 2^b grid, cells numbered along a
//                      Hilbert curve, and counting-sorted (histogram -> single-CTA scan ->
//                      scatter) into a float4 {x, y, z, original index} array.  The grid is
//                      ONLY a locality heuristic: correctness never depends on it.
//   3. tile AABBs      every 128 consecutive sorted points form a tile with an exact bounding
//                      box computed from the coordinates themselves
//   4. search          one WARP per query, queries visited in sorted (spatially coherent) order so
//                      that neighbouring queries hit the same tiles in L1.  The tile holding the
//                      query's own cell seeds the top-k (a warp radix-select picks the k nearest
//                      of its 128 points); then boxes of 32 tiles and the tile boxes inside the
//                      surviving groups are tested against the current k-th distance, and only
//                      tiles that can still contribute are evaluated — with the reference
//                      distance expression bit for bit, 32 points per warp instruction.  The
//                      sorted top-k is distributed over the lanes (k <= 128 in registers).
//
// Culling is conservative: a tile is skipped only if  lb2 * (1 - 2^-13) > threshold , where lb2
// is the box-to-point squared gap evaluated in FP32; the factor covers the
// rounding of the bound and of the reference expression (relative error < 2^-21 each).
// Candidates arrive out of index order, so the list is ordered by the explicit pair
// (d2, original index) — ties broken by lowest index, independent of visiting order.
#include "common.cuh"

namespace amc3d {

#ifndef AMC3D_GT
#define AMC3D_GT 128
#endif
constexpr int GT = AMC3D_GT;            // points per tile (32, 64 or 128)
constexpr int GR = GT / 32;             // rounds of 32 lanes per tile
constexpr float KG_INIT = 1e10f;
constexpr float KG_SAFE = 1.0f - 1.0f / 8192.0f;

// ---------------------------------------------------------------------------------------
// 1. bounding box (ordered-uint encoding so atomicMin/atomicMax work on floats of any sign)
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t f2ord(float f) {
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t o) {
    return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}

// bb[b*8 + 0..2] = min (ordered), bb[b*8 + 3..5] = max (ordered); initialised to 0xffffffff / 0
__global__ void __launch_bounds__(256)
bbox_kernel(int n, const float *__restrict__ xyz, uint32_t *__restrict__ bb) {
    __shared__ float s_lo[3][8], s_hi[3][8];
    xyz += 3ll * blockIdx.y * n;
    bb += 8 * blockIdx.y;
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float v = __ldg(xyz + 3ll * i + c);
            lo[c] = fminf(lo[c], v);
            hi[c] = fmaxf(hi[c], v);
        }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[c] = fminf(lo[c], __shfl_xor_sync(0xffffffffu, lo[c], o));
            hi[c] = fmaxf(hi[c], __shfl_xor_sync(0xffffffffu, hi[c], o));
        }
        if ((threadIdx.x & 31) == 0) {
            s_lo[c][threadIdx.x >> 5] = lo[c];
            s_hi[c][threadIdx.x >> 5] = hi[c];
        }
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        const int c = threadIdx.x;
        float l = s_lo[c][0], h = s_hi[c][0];
#pragma unroll
        for (int w = 1; w < 8; ++w) {
            l = fminf(l, s_lo[c][w]);
            h = fmaxf(h, s_hi[c][w]);
        }
        if (l <= h) {
            atomicMin(bb + c, f2ord(l));
            atomicMax(bb + 3 + c, f2ord(h));
        }
    }
}

__global__ void bbox_init_kernel(int nb, uint32_t *bb) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nb * 8) bb[i] = (i & 7) < 3 ? 0xffffffffu : 0u;
}

// ---------------------------------------------------------------------------------------
// 2. cell of a point along the space-filling curve (clamped into the grid; NaN -> cell 0)
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t spread3(uint32_t v) {   // 10 bits -> every third bit
    v = (v | (v << 16)) & 0x030000ffu;
    v = (v | (v << 8)) & 0x0300f00fu;
    v = (v | (v << 4)) & 0x030c30c3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}

// Position of grid cell (c[0], c[1], c[2]) along a 3-D Hilbert curve of `bits` levels (Skilling's
// transposition, "Programming the Hilbert curve", 2004).  Unlike the Morton order the Hilbert order has
// no jumps — consecutive cells always share a face — so a run of 128 consecutive sorted points is a
// compact patch and its bounding box is tight; along the Morton curve a run that straddles a jump has a
// box covering everything in between, which every query nearby must then evaluate.  Any bijection of
// the cells onto [0, 8^bits) would give identical search results; this one gives the fewest tile visits.
__device__ __forceinline__ uint32_t hilbert_of(uint32_t c0, uint32_t c1, uint32_t c2, int bits) {
    uint32_t X[3] = {c0, c1, c2};
    const uint32_t M = 1u << (bits - 1);
    for (uint32_t Q = M; Q > 1; Q >>= 1) {
        const uint32_t P = Q - 1;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            if (X[i] & Q) X[0] ^= P;
            else {
                const uint32_t t = (X[0] ^ X[i]) & P;
                X[0] ^= t;
                X[i] ^= t;
            }
        }
    }
    X[1] ^= X[0];
    X[2] ^= X[1];
    uint32_t t = 0;
    for (uint32_t Q = M; Q > 1; Q >>= 1)
        if (X[2] & Q) t ^= Q - 1;
    X[0] ^= t; X[1] ^= t; X[2] ^= t;
    return (spread3(X[0]) << 2) | (spread3(X[1]) << 1) | spread3(X[2]);
}

__device__ __forceinline__ uint32_t cell_of(float x, float y, float z, const uint32_t *__restrict__ bb, int bits) {
    const int G = 1 << bits;
    uint32_t c[3];
    const float p[3] = {x, y, z};
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float lo = ord2f(__ldg(bb + a)), hi = ord2f(__ldg(bb + 3 + a));
        const float ext = fmaxf(hi - lo, 1e-20f);
        float t = (p[a] - lo) / ext * (float)G;
        t = fminf(fmaxf(t, 0.f), (float)(G - 1));      // also maps NaN to 0 (fmaxf(NaN,0) = 0)
        c[a] = (uint32_t)t;
    }
#ifdef AMC3D_MORTON
    return spread3(c[0]) | (spread3(c[1]) << 1) | (spread3(c[2]) << 2);
#else
    return hilbert_of(c[0], c[1], c[2], bits);
#endif
}

__global__ void __launch_bounds__(256)
cell_count_kernel(int n, const float *__restrict__ xyz, const uint32_t *__restrict__ bb, int bits,
                  uint32_t *__restrict__ cell, int *__restrict__ counts) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    const int b = blockIdx.y;
    if (i >= n) return;
    const float *p = xyz + 3ll * ((long long)b * n + i);
    const uint32_t c = ((uint32_t)b << (3 * bits)) | cell_of(__ldg(p), __ldg(p + 1), __ldg(p + 2), bb + 8 * b, bits);
    cell[(long long)b * n + i] = c;
    atomicAdd(counts + c, 1);
}

// exclusive scan of `cells` ints in three small launches: per-CTA scan of 4096 elements + CTA
// totals, scan of the totals by one CTA, add-back
constexpr int SCAN_PER = 16, SCAN_THREADS = 256, SCAN_CHUNK = SCAN_PER * SCAN_THREADS;

__device__ __forceinline__ int block_exclusive(int sum, int *s_warp, int &total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int w = lane < SCAN_THREADS / 32 ? s_warp[lane] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += t;
        }
        s_warp[lane] = w;
    }
    __syncthreads();
    total = s_warp[SCAN_THREADS / 32 - 1];
    return incl - sum + (warp > 0 ? s_warp[warp - 1] : 0);
}

__global__ void __launch_bounds__(SCAN_THREADS)
scan_local_kernel(int cells, int *__restrict__ counts, int *__restrict__ totals) {
    __shared__ int s_warp[32];
    const int base = blockIdx.x * SCAN_CHUNK + threadIdx.x * SCAN_PER;
    int v[SCAN_PER], sum = 0;
#pragma unroll
    for (int j = 0; j < SCAN_PER; ++j) {
        v[j] = base + j < cells ? counts[base + j] : 0;
        sum += v[j];
    }
    int total;
    int excl = block_exclusive(sum, s_warp, total);
#pragma unroll
    for (int j = 0; j < SCAN_PER; ++j) {
        if (base + j < cells) counts[base + j] = excl;
        excl += v[j];
    }
    if (threadIdx.x == 0) totals[blockIdx.x] = total;
}

// one CTA: exclusive scan of up to SCAN_CHUNK totals (enough for 2^24 cells)
__global__ void __launch_bounds__(SCAN_THREADS)
scan_totals_kernel(int nblocks, int *__restrict__ totals) {
    __shared__ int s_warp[32];
    const int base = threadIdx.x * SCAN_PER;
    int v[SCAN_PER], sum = 0;
#pragma unroll
    for (int j = 0; j < SCAN_PER; ++j) {
        v[j] = base + j < nblocks ? totals[base + j] : 0;
        sum += v[j];
    }
    int total;
    int excl = block_exclusive(sum, s_warp, total);
#pragma unroll
    for (int j = 0; j < SCAN_PER; ++j) {
        if (base + j < nblocks) totals[base + j] = excl;
        excl += v[j];
    }
}

__global__ void __launch_bounds__(SCAN_THREADS)
scan_add_kernel(int cells, int *__restrict__ counts, const int *__restrict__ totals, int *__restrict__ copy) {
    const int off = totals[blockIdx.x];
    const int base = blockIdx.x * SCAN_CHUNK + threadIdx.x * SCAN_PER;
#pragma unroll
    for (int j = 0; j < SCAN_PER; ++j)
        if (base + j < cells) {
            const int v = counts[base + j] + off;
            counts[base + j] = v;
            copy[base + j] = v;
        }
}

// scatter into cell order; `starts` is consumed as the running fill pointer of each cell.  The scan runs over
// the cells of all clouds, so cloud b starts at b*n; `pad` = npad - n shifts it to b*npad.
__global__ void __launch_bounds__(256)
scatter_kernel(int n, int pad, const float *__restrict__ xyz, const uint32_t *__restrict__ cell,
               int *__restrict__ starts, float4 *__restrict__ sorted) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    const int b = blockIdx.y;
    if (i >= n) return;
    const long long g = (long long)b * n + i;
    const int pos = atomicAdd(starts + cell[g], 1) + b * pad;
    sorted[pos] = make_float4(__ldg(xyz + 3 * g), __ldg(xyz + 3 * g + 1), __ldg(xyz + 3 * g + 2), __int_as_float(i));
}

// ---------------------------------------------------------------------------------------
// 3. tile boxes: one warp per tile of GT sorted points
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
tile_aabb_kernel(int n, int ntb, int ntiles, const float4 *__restrict__ sorted, float4 *__restrict__ tlo,
                 float4 *__restrict__ thi) {
    const int t = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (t >= ntiles) return;
    const int lt = t % ntb;                       // tile within its cloud
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int j = lane; j < GT; j += 32) {
        if (lt * GT + j < n) {
            const float4 p = __ldg(sorted + (long long)t * GT + j);
            lo[0] = fminf(lo[0], p.x); hi[0] = fmaxf(hi[0], p.x);
            lo[1] = fminf(lo[1], p.y); hi[1] = fmaxf(hi[1], p.y);
            lo[2] = fminf(lo[2], p.z); hi[2] = fmaxf(hi[2], p.z);
        }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[c] = fminf(lo[c], __shfl_xor_sync(0xffffffffu, lo[c], o));
            hi[c] = fmaxf(hi[c], __shfl_xor_sync(0xffffffffu, hi[c], o));
        }
    if (lane == 0) {
        tlo[t] = make_float4(lo[0], lo[1], lo[2], 0.f);
        thi[t] = make_float4(hi[0], hi[1], hi[2], 0.f);
    }
}

// ---------------------------------------------------------------------------------------
// 4. search: one WARP per query
// ---------------------------------------------------------------------------------------
// Why a warp per query and not a thread per query as in the brute-force kernel: once culling
// has cut the candidate set from n to a few hundred points, the cost is no longer the distance
// evaluations but the top-k maintenance, and thread-per-query insertion is divergent (every
// lane's insertion stalls the other 31).  Here the 32 lanes evaluate 32 support points of ONE
// query at a time; the sorted top-k is distributed over the lanes (E entries per lane) and a
// candidate is inserted cooperatively with one ballot-free compare per entry and a one-lane
// shuffle — about 20 instructions per accepted candidate, no divergence.  It also makes sparse
// query sets (the label vote: 3 000 queries against 192 000 points, k = 64) as efficient as
// dense ones, and k up to 128 needs no shared-memory list.
__device__ __forceinline__ bool lex_lt(float ad, int ai, float bd, int bi) {
    return ad < bd || (ad == bd && ai < bi);
}

// squared gap between [alo,ahi] and [blo,bhi] along one axis
__device__ __forceinline__ float gap2(float alo, float ahi, float blo, float bhi) {
    const float g = fmaxf(fmaxf(blo - ahi, alo - bhi), 0.f);
    return g * g;
}
__device__ __forceinline__ float point_box2(float x, float y, float z, const float4 lo, const float4 hi) {
    return gap2(x, x, lo.x, hi.x) + gap2(y, y, lo.y, hi.y) + gap2(z, z, lo.z, hi.z);
}

constexpr int WQ_WARPS = 8;     // warps per CTA
constexpr int WQ_QPW_MAX = 1;   // consecutive sorted queries per warp.  One: measured 16 -> 2 -> 1 at config 2: kNN
                                // 1.37 -> 1.32 -> 1.24 ms, ball queries 0.97 -> 0.86 -> 0.86 ms (shorter tail of the
                                // last wave; neighbouring warps share the tiles through L1 just as well).
                                // AMC3D_KNN_QPW overrides.
constexpr int TG = 32;          // tiles per box group (second culling level)

// Sorted (ascending) list of S = 32*E entries distributed over the warp: lane l holds slots
// l*E .. l*E+E-1.  The nsample live entries are right-aligned (slots S-nsample .. S-1), the
// slots in front hold (-inf, 0) sentinels, so the k-th best is always slot S-1 = lane 31, e = E-1.
template <int E>
struct WarpList {
    float d[E];
    int i[E];
    __device__ __forceinline__ void init(int nsample, int lane) {
#pragma unroll
        for (int e = 0; e < E; ++e) {
            d[e] = (lane * E + e) < 32 * E - nsample ? -INFINITY : KG_INIT;
            i[e] = 0;
        }
    }
    __device__ __forceinline__ void threshold(float &td, int &ti) const {
        td = __shfl_sync(0xffffffffu, d[E - 1], 31);
        ti = __shfl_sync(0xffffffffu, i[E - 1], 31);
    }
    // insert (cd, ci), known to be lex-smaller than the current last entry; warp-uniform arguments
    __device__ __forceinline__ void insert(float cd, int ci, int lane) {
        bool lt[E];
#pragma unroll
        for (int e = 0; e < E; ++e) lt[e] = lex_lt(cd, ci, d[e], i[e]);   // entry e moves up one slot
        const float pd = __shfl_up_sync(0xffffffffu, d[E - 1], 1);
        const int pi = __shfl_up_sync(0xffffffffu, i[E - 1], 1);
        const bool plt = __shfl_up_sync(0xffffffffu, (int)lt[E - 1], 1) && lane > 0;
#pragma unroll
        for (int e = E - 1; e > 0; --e) {
            d[e] = lt[e - 1] ? d[e - 1] : (lt[e] ? cd : d[e]);
            i[e] = lt[e - 1] ? i[e - 1] : (lt[e] ? ci : i[e]);
        }
        d[0] = plt ? pd : (lt[0] ? cd : d[0]);
        i[0] = plt ? pi : (lt[0] ? ci : i[0]);
    }
};

// An upper bound T of the k-th smallest of the NV x 32 distances held by the warp, tight to 2^-7
// relative: radix select on the bit pattern (non-negative floats order like their bits) over bits
// 30..17, two bits per step — the counts for the three candidate prefixes travel in one packed
// REDUX.SUM (each count <= 32 NV <= 1023 fits ten bits) — and the 17 low bits rounded up.  #{d <= T} >= k
// always; the handful of extra candidates a looser T lets through fall off the end of the sorted list.
template <int NV>
__device__ __forceinline__ float warp_kth_bound(const float (&dd)[NV], int k) {
    static_assert(NV * 32 < 1024, "packed counters are ten bits wide");
    uint32_t v[NV];
#pragma unroll
    for (int r = 0; r < NV; ++r) v[r] = __float_as_uint(dd[r]);
    uint32_t T = 0;
#pragma unroll
    for (int b = 29; b >= 17; b -= 2) {
        const uint32_t t1 = T | (1u << b), t2 = T | (2u << b), t3 = T | (3u << b);
        uint32_t c = 0;
#pragma unroll
        for (int r = 0; r < NV; ++r)
            c += (v[r] < t1 ? 1u : 0u) + (v[r] < t2 ? 0x400u : 0u) + (v[r] < t3 ? 0x100000u : 0u);
        c = __reduce_add_sync(0xffffffffu, c);
        const int c1 = c & 0x3ff, c2 = (c >> 10) & 0x3ff, c3 = c >> 20;
        // fewer than k values below a prefix: the k-th is >= that prefix
        T = c3 < k ? t3 : (c2 < k ? t2 : (c1 < k ? t1 : T));
    }
    return __uint_as_float(T | 0x1ffffu);
}

// geometry of a batched, tile-aligned sorted point set
struct Geom {
    int nb;      // clouds
    int n;       // points per cloud
    int npad;    // n rounded up to whole tiles
    int ntb;     // tiles per cloud
    int ngb;     // tile groups per cloud
    int bits;    // grid bits per axis
};

// STATS: count the distance evaluations actually issued (tiles visited x GT) into stats[0..1] — a separate
// instantiation, launched only while amc3d_search_stats() has a counter array registered (bench.py's
// "evaluated pairs" figure), so the production kernels carry no trace of it.
template <int E, bool STATS = false>
__global__ void __launch_bounds__(WQ_WARPS * 32)
knn_wq_kernel(Geom gs, int m, int mpad, int nsample, int self, int qpw, const float4 *__restrict__ sp,
              const float4 *__restrict__ tlo, const float4 *__restrict__ thi, const float4 *__restrict__ glo,
              const float4 *__restrict__ ghi, const float4 *__restrict__ sq, const int *__restrict__ cell_start,
              const uint32_t *__restrict__ bb, int *__restrict__ idx, float *__restrict__ dist2,
              int *__restrict__ order_out, unsigned long long *__restrict__ stats) {
    const int lane = threadIdx.x & 31;
    const int wq0 = (blockIdx.x * WQ_WARPS + (threadIdx.x >> 5)) * qpw;
    const int n = gs.n;

    for (int j = 0; j < qpw; ++j) {
        const int q = wq0 + j;
        if (q >= gs.nb * mpad) return;                        // warp-uniform
        const int b = gs.nb == 1 ? 0 : q / mpad;
        if (q - b * mpad >= m) continue;                      // padding slot of the query layout
        const float4 me = __ldg(sq + q);
        int ntiles = 1;                                       // STATS only: tiles evaluated for this query (seed = 1)
        // the visiting order (spatially coherent) for callers that want to process the queries the same way
        if (order_out != nullptr && lane == 0) order_out[(long long)b * m + (q - b * mpad)] = __float_as_int(me.w);
        const float qx = me.x, qy = me.y, qz = me.z;
        WarpList<E> best;
        best.init(nsample, lane);
        float td = KG_INIT;
        int ti = 0x7fffffff;
        const int tb0 = b * gs.ntb;                           // first tile of the query's cloud

        // distances of this lane's GR points of global tile t (+inf past the end of the cloud or for a tile
        // outside it)
        auto eval_tile = [&](int t, float *dd, int *oi) {
            const float4 *tp = sp + (long long)t * GT + lane;
            const int left = (t >= tb0 && t < tb0 + gs.ntb) ? n - (t - tb0) * GT : 0;
            if (left >= GT) {                                  // a whole tile (all but the last of a cloud)
#pragma unroll
                for (int r = 0; r < GR; ++r) {
                    const float4 p = __ldg(tp + r * 32);
                    dd[r] = dist2_ref(qx - p.x, qy - p.y, qz - p.z);
                    oi[r] = __float_as_int(p.w);
                }
            } else {
#pragma unroll
                for (int r = 0; r < GR; ++r) {
                    dd[r] = INFINITY;
                    oi[r] = 0x7fffffff;
                    if (r * 32 + lane < left) {
                        const float4 p = __ldg(tp + r * 32);
                        dd[r] = dist2_ref(qx - p.x, qy - p.y, qz - p.z);
                        oi[r] = __float_as_int(p.w);
                    }
                }
            }
        };
        // evaluate the 128 points of global tile t against this query
        auto process_tile = [&](int t) {
            float dd[GR];
            int oi[GR];
            eval_tile(t, dd, oi);
            if (STATS) ++ntiles;
#pragma unroll
            for (int r = 0; r < GR; ++r) {
                bool cand = lex_lt(dd[r], oi[r], td, ti);
                uint32_t mask = __ballot_sync(0xffffffffu, cand);
                while (mask) {
                    const int bl = __ffs(mask) - 1;
                    const float cd = __shfl_sync(0xffffffffu, dd[r], bl);
                    const int ci = __shfl_sync(0xffffffffu, oi[r], bl);
                    best.insert(cd, ci, lane);
                    float nd;
                    int ni;
                    best.threshold(nd, ni);
                    td = nd;
                    ti = ni;
                    cand = cand && lane != bl && lex_lt(dd[r], oi[r], td, ti);
                    mask = __ballot_sync(0xffffffffu, cand);
                }
            }
        };

        // ---- seed: the tile at the query's own position; then its two neighbours along the curve ----------
        // A warp radix-select over the own tile's 128 distances gives its k-th smallest up front, so only the
        // ~k nearest go through the insertion, and without the per-insert threshold refresh (an entry beyond
        // the k-th falls off the list).  The neighbours follow unconditionally: skipping them costs 2.4x the
        // search time (a query near the edge of its tile would start the box tests with a bound that lets
        // most of the surrounding tiles through); selecting over all three tiles at once costs 10 % (the
        // select over 12 values per lane is dearer than the insertions it saves).
        int t0;
        if (self) t0 = q / GT;
        else
            t0 = (__ldg(cell_start + (((uint32_t)b << (3 * gs.bits)) | cell_of(qx, qy, qz, bb + 8 * b, gs.bits))) +
                  b * (gs.npad - n)) / GT;
        t0 = min(max(t0, tb0), tb0 + gs.ntb - 1);
        const int r_lo = max(t0 - 1, tb0), r_hi = min(t0 + 1, tb0 + gs.ntb - 1);
        {
            float dd[GR];
            int oi[GR];
            eval_tile(t0, dd, oi);
            td = fminf(warp_kth_bound<GR>(dd, min(nsample, GT)), KG_INIT);
#pragma unroll
            for (int r = 0; r < GR; ++r) {
                uint32_t mask = __ballot_sync(0xffffffffu, dd[r] <= td);
                while (mask) {
                    const int bl = __ffs(mask) - 1;
                    mask &= mask - 1;
                    best.insert(__shfl_sync(0xffffffffu, dd[r], bl), __shfl_sync(0xffffffffu, oi[r], bl), lane);
                }
            }
            best.threshold(td, ti);
        }
        for (int t = r_lo; t <= r_hi; ++t)
            if (t != t0) process_tile(t);

        // ---- everything else through two levels of boxes -----------------------------------
        for (int g0 = 0; g0 < gs.ngb; g0 += 32) {
            const int g = g0 + lane;
            bool gsel = false;
            if (g < gs.ngb)
                gsel = !(point_box2(qx, qy, qz, __ldg(glo + b * gs.ngb + g), __ldg(ghi + b * gs.ngb + g)) * KG_SAFE > td);
            uint32_t gmask = __ballot_sync(0xffffffffu, gsel);
            while (gmask) {
                const int gb = __ffs(gmask) - 1;
                gmask &= gmask - 1;
                const int lt = (g0 + gb) * TG + lane;            // tile within the cloud
                const int t = tb0 + lt;
                float lb = INFINITY;
                if (lt < gs.ntb && (t < r_lo || t > r_hi)) lb = point_box2(qx, qy, qz, __ldg(tlo + t), __ldg(thi + t));
                uint32_t tmask = __ballot_sync(0xffffffffu, !(lb * KG_SAFE > td));
                while (tmask) {
                    const int tbit = __ffs(tmask) - 1;
                    tmask &= tmask - 1;
                    const float lbt = __shfl_sync(0xffffffffu, lb, tbit);
                    if (lbt * KG_SAFE > td) continue;              // threshold tightened meanwhile
                    process_tile(tb0 + (g0 + gb) * TG + tbit);
                }
            }
        }

        // ---- write the live entries (slots S-nsample .. S-1) to the query's original row ----
        const long long qorig = (long long)b * m + __float_as_int(me.w);
        const int shift = 32 * E - nsample;
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int s = lane * E + e - shift;
            if (s >= 0) {
                idx[qorig * nsample + s] = best.i[e];
                dist2[qorig * nsample + s] = best.d[e];
            }
        }
        if (STATS && lane == 0) {
            atomicAdd(stats + 0, (unsigned long long)ntiles * GT);
            atomicAdd(stats + 1, 1ull);
        }
    }
}

// ---------------------------------------------------------------------------------------
// 4b. search for small k with dense queries (three_nn onto a finer level): one THREAD per query, tiles shared
//     by the warp
// ---------------------------------------------------------------------------------------
// With k <= 8 the sorted list fits a few registers of ONE thread and an insertion is a handful of selects,
// so the warp-distributed list above (a shuffle chain per accepted candidate) is the wrong trade.  Here a warp
// takes 32 consecutive sorted queries — neighbours in space — and walks the tiles that ANY of them can still
// need: a tile is skipped only if the gap between its box and the box of the warp's queries, squared, exceeds
// the largest current k-th distance in the warp (same conservative factor as above), which implies the
// per-query condition for every lane.  Every lane evaluates every point of a visited tile (the loads are
// warp-uniform: one L1 transaction serves 32 queries) and keeps its own (d2, index)-ordered list, so the
// result is again exactly the lexicographic top-k, whatever the visiting order.
__device__ __forceinline__ float box_box2(const float (&alo)[3], const float (&ahi)[3], const float4 blo, const float4 bhi) {
    return gap2(alo[0], ahi[0], blo.x, bhi.x) + gap2(alo[1], ahi[1], blo.y, bhi.y) + gap2(alo[2], ahi[2], blo.z, bhi.z);
}

constexpr int TQ_WARPS = 4;

template <int K, bool STATS = false>
__global__ void __launch_bounds__(TQ_WARPS * 32)
knn_tq_kernel(Geom gs, int m, int mpad, int self, const float4 *__restrict__ sp, const float4 *__restrict__ tlo,
              const float4 *__restrict__ thi, const float4 *__restrict__ glo, const float4 *__restrict__ ghi,
              const float4 *__restrict__ sq, const int *__restrict__ cell_start, const uint32_t *__restrict__ bb,
              int *__restrict__ idx, float *__restrict__ dist2, int *__restrict__ order_out,
              unsigned long long *__restrict__ stats) {
    const int lane = threadIdx.x & 31;
    int ntiles = 0;                                   // STATS only
    const long long q0 = ((long long)blockIdx.x * TQ_WARPS + (threadIdx.x >> 5)) * 32;   // first query slot of the warp
    if (q0 >= (long long)gs.nb * mpad) return;
    const int b = (int)(q0 / mpad);                   // mpad is a multiple of 32: one cloud per warp
    const int qpos = (int)(q0 - (long long)b * mpad) + lane;
    if ((int)(q0 - (long long)b * mpad) >= m) return; // only padding slots
    const bool valid = qpos < m;
    const int n = gs.n;
    // padding lanes (end of the cloud's query block) shadow the warp's first query and write nothing
    const float4 me = __ldg(sq + (valid ? q0 + lane : q0));
    const float qx = me.x, qy = me.y, qz = me.z;
    if (order_out != nullptr && valid) order_out[(long long)b * m + qpos] = __float_as_int(me.w);

    float bd[K];
    int bi[K];
#pragma unroll
    for (int e = 0; e < K; ++e) { bd[e] = KG_INIT; bi[e] = 0; }
    const int tb0 = b * gs.ntb;

    // all lanes evaluate all points of tile t.  The tile comes in with ONE round trip to L2 (every lane fetches
    // four of its points, coalesced) and is parked in the warp's shared-memory slot; from there the points are
    // read back warp-uniformly (broadcast), eight at a time, and the insertion code is entered only if some
    // lane has a point inside its bound
    __shared__ float4 s_tile[TQ_WARPS][GT];
    float4 *buf = s_tile[threadIdx.x >> 5];
    auto process_tile = [&](int t) {
        const int cnt = min(GT, n - (t - tb0) * GT);   // warp-uniform
        const float4 *tp = sp + (long long)t * GT;
        if (STATS) ++ntiles;
        float4 v[GR];
#pragma unroll
        for (int r = 0; r < GR; ++r)                   // +inf coordinates: distance +inf, never accepted
            v[r] = r * 32 + lane < cnt ? __ldg(tp + r * 32 + lane) : make_float4(INFINITY, INFINITY, INFINITY, 0.f);
        __syncwarp();                                  // the previous tile has been read by every lane
#pragma unroll
        for (int r = 0; r < GR; ++r) buf[r * 32 + lane] = v[r];
        __syncwarp();
        for (int j0 = 0; j0 < cnt; j0 += 8) {
            float4 p[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) p[u] = buf[j0 + u];
            float d[8];
            bool pass = false;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                d[u] = dist2_ref(qx - p[u].x, qy - p[u].y, qz - p[u].z);
                pass |= d[u] <= bd[K - 1];
            }
            if (!__any_sync(0xffffffffu, pass)) continue;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int oi = __float_as_int(p[u].w);
                const bool cand = lex_lt(d[u], oi, bd[K - 1], bi[K - 1]);
                if (!__any_sync(0xffffffffu, cand)) continue;
                bool lt[K];
#pragma unroll
                for (int e = 0; e < K; ++e) lt[e] = cand && lex_lt(d[u], oi, bd[e], bi[e]);
#pragma unroll
                for (int e = K - 1; e > 0; --e) {
                    bd[e] = lt[e - 1] ? bd[e - 1] : (lt[e] ? d[u] : bd[e]);
                    bi[e] = lt[e - 1] ? bi[e - 1] : (lt[e] ? oi : bi[e]);
                }
                bd[0] = lt[0] ? d[u] : bd[0];
                bi[0] = lt[0] ? oi : bi[0];
            }
        }
    };
    // ---- seeds: the tile at every query's own position (one or two distinct tiles for a coherent warp) ----
    int t_own;
    if (self) t_own = (int)((q0 + (valid ? lane : 0)) / GT);
    else
        t_own = (__ldg(cell_start + (((uint32_t)b << (3 * gs.bits)) | cell_of(qx, qy, qz, bb + 8 * b, gs.bits))) +
                 b * (gs.npad - n)) / GT;
    t_own = min(max(t_own, tb0), tb0 + gs.ntb - 1);
    {
        uint32_t todo = 0xffffffffu;
        while (todo) {
            const int t = __shfl_sync(0xffffffffu, t_own, __ffs(todo) - 1);
            process_tile(t);
            todo &= ~__ballot_sync(0xffffffffu, t_own == t);
        }
    }

    // ---- the box of the warp's queries and the loosest bound among them: a cheap first filter, lanes over
    // boxes; what passes it is tested exactly, lanes over queries (a run of 32 queries along the curve
    // may span a bend of the curve, and then its box says little) -------------------------------------
    float wlo[3] = {qx, qy, qz}, whi[3] = {qx, qy, qz};
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            wlo[c] = fminf(wlo[c], __shfl_xor_sync(0xffffffffu, wlo[c], o));
            whi[c] = fmaxf(whi[c], __shfl_xor_sync(0xffffffffu, whi[c], o));
        }
    for (int g0 = 0; g0 < gs.ngb; g0 += 32) {
        // k-th distances are >= 0 and never NaN: their bit patterns order like the values
        float tdmax = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(bd[K - 1])));
        const int g = g0 + lane;
        bool gsel = false;
        if (g < gs.ngb)
            gsel = !(box_box2(wlo, whi, __ldg(glo + b * gs.ngb + g), __ldg(ghi + b * gs.ngb + g)) * KG_SAFE > tdmax);
        uint32_t gmask = __ballot_sync(0xffffffffu, gsel);
        while (gmask) {
            const int gb = __ffs(gmask) - 1;
            gmask &= gmask - 1;
            const int gi = b * gs.ngb + g0 + gb;
            if (!__any_sync(0xffffffffu, !(point_box2(qx, qy, qz, __ldg(glo + gi), __ldg(ghi + gi)) * KG_SAFE > bd[K - 1])))
                continue;
            tdmax = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(bd[K - 1])));
            const int lt = (g0 + gb) * TG + lane;
            float lb = INFINITY;
            if (lt < gs.ntb) lb = box_box2(wlo, whi, __ldg(tlo + tb0 + lt), __ldg(thi + tb0 + lt));
            uint32_t tmask = __ballot_sync(0xffffffffu, !(lb * KG_SAFE > tdmax));
            while (tmask) {
                const int t = tb0 + (g0 + gb) * TG + __ffs(tmask) - 1;
                tmask &= tmask - 1;
                const bool need = !(point_box2(qx, qy, qz, __ldg(tlo + t), __ldg(thi + t)) * KG_SAFE > bd[K - 1]);
                // a seed tile was evaluated already (evaluating it twice would list its points twice)
                if (!__any_sync(0xffffffffu, need) || __any_sync(0xffffffffu, t_own == t)) continue;
                process_tile(t);
            }
        }
    }

    if (valid) {
        const long long row = ((long long)b * m + __float_as_int(me.w)) * K;
#pragma unroll
        for (int e = 0; e < K; ++e) {
            idx[row + e] = bi[e];
            dist2[row + e] = bd[e];
        }
    }
    if (STATS && lane == 0) {                         // every lane evaluates every point of a visited tile
        atomicAdd(stats + 2, (unsigned long long)ntiles * GT * 32);
        atomicAdd(stats + 3, 32ull);
    }
}

// ---------------------------------------------------------------------------------------
// ball query over the same structure: the nsample SMALLEST original indices with d2 < r^2
// ---------------------------------------------------------------------------------------
// The reference scans the support cloud in index order and keeps the first nsample hits
// (ball_query_gpu.cu:29-50), i.e. the nsample smallest indices inside the ball; slots beyond the hit
// count repeat the first hit, rows without a hit are not written.  Here a warp visits only the
// tiles whose box intersects the ball (conservatively: skipped only if lb2 * (1 - 2^-13) >= r^2) and
// keeps the sorted list of the nsample smallest hit indices distributed over its lanes.
template <int E>
struct WarpIdxList {
    int i[E];
    __device__ __forceinline__ void init(int nsample, int lane) {
#pragma unroll
        for (int e = 0; e < E; ++e) i[e] = (lane * E + e) < 32 * E - nsample ? -1 : 0x7fffffff;
    }
    __device__ __forceinline__ int threshold() const { return __shfl_sync(0xffffffffu, i[E - 1], 31); }
    // insert ci, known to be smaller than the current last entry; warp-uniform argument
    __device__ __forceinline__ void insert(int ci, int lane) {
        bool lt[E];
#pragma unroll
        for (int e = 0; e < E; ++e) lt[e] = ci < i[e];
        const int pi = __shfl_up_sync(0xffffffffu, i[E - 1], 1);
        const bool plt = __shfl_up_sync(0xffffffffu, (int)lt[E - 1], 1) && lane > 0;
#pragma unroll
        for (int e = E - 1; e > 0; --e) i[e] = lt[e - 1] ? i[e - 1] : (lt[e] ? ci : i[e]);
        i[0] = plt ? pi : (lt[0] ? ci : i[0]);
    }
};

template <int E, bool STATS = false>
__global__ void __launch_bounds__(WQ_WARPS * 32)
ball_wq_kernel(Geom gs, int m, int mpad, float radius, int nsample, int self, int qpw, const float4 *__restrict__ sp,
               const float4 *__restrict__ tlo, const float4 *__restrict__ thi, const float4 *__restrict__ glo,
               const float4 *__restrict__ ghi, const float4 *__restrict__ sq, const int *__restrict__ cell_start,
               const uint32_t *__restrict__ bb, int *__restrict__ idx, unsigned long long *__restrict__ stats) {
    const int lane = threadIdx.x & 31;
    const int wq0 = (blockIdx.x * WQ_WARPS + (threadIdx.x >> 5)) * qpw;
    const int n = gs.n;
    const float r2 = __fmul_rn(radius, radius);

    for (int j = 0; j < qpw; ++j) {
        const int q = wq0 + j;
        if (q >= gs.nb * mpad) return;                        // warp-uniform
        const int b = gs.nb == 1 ? 0 : q / mpad;
        if (q - b * mpad >= m) continue;
        const float4 me = __ldg(sq + q);
        const float qx = me.x, qy = me.y, qz = me.z;
        WarpIdxList<E> best;
        best.init(nsample, lane);
        int ti = 0x7fffffff;                                  // current nsample-th smallest hit index
        const int tb0 = b * gs.ntb;
        int ntiles = 0;                                       // STATS only

        auto process_tile = [&](int t) {
            const int left = n - (t - tb0) * GT;               // >= GT for all but the last tile of a cloud
            if (STATS) ++ntiles;
#pragma unroll
            for (int r = 0; r < GR; ++r) {
                bool cand = false;
                int oi = 0x7fffffff;
                if (left >= GT || r * 32 + lane < left) {
                    const float4 p = __ldg(sp + (long long)t * GT + r * 32 + lane);
                    // operand order of the reference: new_xyz - xyz (ball_query_gpu.cu:38-40)
                    const float d = dist2_ref(qx - p.x, qy - p.y, qz - p.z);
                    oi = __float_as_int(p.w);
                    cand = d < r2 && oi < ti;
                }
                uint32_t mask = __ballot_sync(0xffffffffu, cand);
                while (mask) {
                    const int bl = __ffs(mask) - 1;
                    const int ci = __shfl_sync(0xffffffffu, oi, bl);
                    best.insert(ci, lane);
                    ti = best.threshold();
                    cand = cand && lane != bl && oi < ti;
                    mask = __ballot_sync(0xffffffffu, cand);
                }
            }
        };

        int t0;
        if (self) t0 = q / GT;
        else
            t0 = (__ldg(cell_start + (((uint32_t)b << (3 * gs.bits)) | cell_of(qx, qy, qz, bb + 8 * b, gs.bits))) +
                  b * (gs.npad - n)) / GT;
        t0 = min(max(t0, tb0), tb0 + gs.ntb - 1);
        process_tile(t0);

        for (int g0 = 0; g0 < gs.ngb; g0 += 32) {
            const int g = g0 + lane;
            bool gsel = false;
            if (g < gs.ngb)
                gsel = point_box2(qx, qy, qz, __ldg(glo + b * gs.ngb + g), __ldg(ghi + b * gs.ngb + g)) * KG_SAFE < r2;
            uint32_t gmask = __ballot_sync(0xffffffffu, gsel);
            while (gmask) {
                const int gb = __ffs(gmask) - 1;
                gmask &= gmask - 1;
                const int lt = (g0 + gb) * TG + lane;
                const int t = tb0 + lt;
                bool tsel = false;
                if (lt < gs.ntb && t != t0) tsel = point_box2(qx, qy, qz, __ldg(tlo + t), __ldg(thi + t)) * KG_SAFE < r2;
                uint32_t tmask = __ballot_sync(0xffffffffu, tsel);
                while (tmask) {
                    const int tbit = __ffs(tmask) - 1;
                    tmask &= tmask - 1;
                    process_tile(tb0 + (g0 + gb) * TG + tbit);
                }
            }
        }

        // slots S-nsample .. S-1 hold the hits in ascending index order, INT_MAX where there is none
        const int shift = 32 * E - nsample;
        int first = 0x7fffffff;                               // value of slot `shift` (the smallest hit)
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int v = __shfl_sync(0xffffffffu, best.i[e], shift / E);
            if (e == shift % E) first = v;
        }
        if (STATS && lane == 0) {
            atomicAdd(stats + 4, (unsigned long long)ntiles * GT);
            atomicAdd(stats + 5, 1ull);
        }
        if (first == 0x7fffffff) continue;                    // no hit: the row keeps the caller's zeros
        const long long qorig = (long long)b * m + __float_as_int(me.w);
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int s = lane * E + e - shift;
            if (s >= 0) idx[qorig * nsample + s] = best.i[e] == 0x7fffffff ? first : best.i[e];
        }
    }
}

// boxes of TG consecutive tile boxes of one cloud
__global__ void __launch_bounds__(256)
group_aabb_kernel(int ntb, int ngb, int ngroups, const float4 *__restrict__ tlo, const float4 *__restrict__ thi,
                  float4 *__restrict__ glo, float4 *__restrict__ ghi) {
    const int g = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (g >= ngroups) return;
    const int b = g / ngb, lg = g - b * ngb;
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    const int lt = lg * TG + lane;
    if (lt < ntb) {
        const float4 a = __ldg(tlo + b * ntb + lt), c = __ldg(thi + b * ntb + lt);
        lo[0] = a.x; lo[1] = a.y; lo[2] = a.z;
        hi[0] = c.x; hi[1] = c.y; hi[2] = c.z;
    }
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[c] = fminf(lo[c], __shfl_xor_sync(0xffffffffu, lo[c], o));
            hi[c] = fmaxf(hi[c], __shfl_xor_sync(0xffffffffu, hi[c], o));
        }
    if (lane == 0) {
        glo[g] = make_float4(lo[0], lo[1], lo[2], 0.f);
        ghi[g] = make_float4(hi[0], hi[1], hi[2], 0.f);
    }
}

// Scratch comes from a PRIVATE stream-ordered pool per device, owned by this library, never from the
// device's default pool: a pool hands freed memory back to the driver at every synchronisation point
// unless a release threshold is set — a trainer that reads the loss each step (loss.item()) would
// otherwise pay a fresh physical allocation for every scratch buffer of every search call (measured:
// 96 ms per step instead of 22 ms) — and raising the threshold of the default pool would change the
// behaviour of every other cudaMallocAsync user in the host process.  The private pool keeps what the
// largest search needed (bounded by the caller's problem sizes; amc3d_trim_scratch() returns it).
static cudaMemPool_t g_pool[64] = {};
static cudaError_t scratch_pool(cudaMemPool_t *out) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
    if (g_pool[dev] == nullptr) {
        cudaMemPoolProps props = {};
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = dev;
        cudaMemPool_t pool;
        e = cudaMemPoolCreate(&pool, &props);
        if (e != cudaSuccess) return e;
        unsigned long long thr = ~0ull;
        e = cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
        if (e != cudaSuccess) return e;
        g_pool[dev] = pool;
    }
    *out = g_pool[dev];
    return cudaSuccess;
}

// Counter array registered by amc3d_search_stats (device, 8 x u64), or NULL: see knn_wq_kernel<E, STATS>.
static unsigned long long *g_search_stats = nullptr;
extern "C" int amc3d_search_stats(void *counters) {
    g_search_stats = static_cast<unsigned long long *>(counters);
    return 0;
}

// Give the scratch the searches have cached on the current device back to the driver (keeps `keep_bytes`).
extern "C" int amc3d_trim_scratch(size_t keep_bytes) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e == cudaSuccess && (dev < 0 || dev >= 64)) e = cudaErrorInvalidDevice;
    if (e == cudaSuccess && g_pool[dev] != nullptr) e = cudaMemPoolTrimTo(g_pool[dev], keep_bytes);
    if (e != cudaSuccess) set_error("trim_scratch: %s", cudaGetErrorString(e));
    return (int)e;
}

// per-stream scratch, stream-ordered (cudaMallocAsync pools make this cheap after warm-up)
struct Scratch {
    cudaStream_t st;
    void *ptrs[24];
    int count = 0;
    cudaError_t err = cudaSuccess;
    cudaMemPool_t pool = nullptr;
    explicit Scratch(cudaStream_t s) : st(s) { err = scratch_pool(&pool); }
    template <class T>
    T *get(size_t n) {
        void *p = nullptr;
        if (err == cudaSuccess && count >= 24) err = cudaErrorMemoryAllocation;
        if (err == cudaSuccess) err = cudaMallocFromPoolAsync(&p, n * sizeof(T) + 16, pool, st);
        if (err == cudaSuccess) ptrs[count++] = p;
        return reinterpret_cast<T *>(p);
    }
    ~Scratch() {
        for (int i = 0; i < count; ++i) cudaFreeAsync(ptrs[i], st);
    }
};

static int pick_bits(int n) {
    int b = 2;
    while (b < 7 && (1ll << (3 * b)) * 2 < n) ++b;   // about 0.5 - 4 points per cell on average
    return b;
}

// sort nb clouds of `count` points each into curve-cell order, cloud b at [b*cpad, b*cpad + count);
// returns cell_start (exclusive starts over all clouds, unpadded) if wanted
static float4 *sort_points(Scratch &ws, int nb, int count, int cpad, const float *xyz, const uint32_t *bb, int bits,
                           int **cell_start_out) {
    const long long cells = (long long)nb << (3 * bits);
    uint32_t *cell = ws.get<uint32_t>((size_t)nb * count);
    int *counts = ws.get<int>(cells);
    int *fill = ws.get<int>(cells);
    float4 *sorted = ws.get<float4>((size_t)nb * cpad);
    if (ws.err != cudaSuccess) return nullptr;
    cudaMemsetAsync(counts, 0, sizeof(int) * cells, ws.st);
    dim3 grid(div_up(count, 256), nb);
    cell_count_kernel<<<grid, 256, 0, ws.st>>>(count, xyz, bb, bits, cell, counts);
    const int sblocks = (int)div_up_ll(cells, SCAN_CHUNK);
    int *totals = ws.get<int>(sblocks);
    if (ws.err != cudaSuccess) return nullptr;
    scan_local_kernel<<<sblocks, SCAN_THREADS, 0, ws.st>>>((int)cells, counts, totals);
    scan_totals_kernel<<<1, SCAN_THREADS, 0, ws.st>>>(sblocks, totals);
    scan_add_kernel<<<sblocks, SCAN_THREADS, 0, ws.st>>>((int)cells, counts, totals, fill);
    scatter_kernel<<<grid, 256, 0, ws.st>>>(count, cpad - count, xyz, cell, fill, sorted);
    if (cell_start_out) *cell_start_out = counts;
    return sorted;
}

// the structure every search uses: sorted support points + tile / group boxes (+ sorted queries)
struct Built {
    Geom gs;
    uint32_t *bb;
    float4 *sp, *tlo, *thi, *glo, *ghi;
    const float4 *sq;
    int *cell_start;
    int self, mpad;
};

// nb clouds: support xyz (nb,n,3), queries new_xyz (nb,m,3).  Returns a cudaError_t.
static cudaError_t build(Scratch &ws, int nb, int n, int m, const float *xyz, const float *new_xyz, Built &B) {
    cudaStream_t st = ws.st;
    Geom &gs = B.gs;
    gs.nb = nb;
    gs.n = n;
    gs.bits = pick_bits(n);
    gs.ntb = div_up(n, GT);
    gs.npad = gs.ntb * GT;
    gs.ngb = div_up(gs.ntb, TG);
    if (((long long)nb << (3 * gs.bits)) > (1ll << 24)) return cudaErrorInvalidValue;   // scan limit
    const int ntiles = nb * gs.ntb, ngroups = nb * gs.ngb;
    B.bb = ws.get<uint32_t>((size_t)nb * 8);
    B.tlo = ws.get<float4>(ntiles);
    B.thi = ws.get<float4>(ntiles);
    B.glo = ws.get<float4>(ngroups);
    B.ghi = ws.get<float4>(ngroups);
    if (ws.err != cudaSuccess) return ws.err;
    bbox_init_kernel<<<div_up(nb * 8, 256), 256, 0, st>>>(nb, B.bb);
    dim3 bgrid(max(1, min(div_up(n, 1024), div_up(kNumSMs, nb))), nb);
    bbox_kernel<<<bgrid, 256, 0, st>>>(n, xyz, B.bb);
    B.cell_start = nullptr;
    B.sp = sort_points(ws, nb, n, gs.npad, xyz, B.bb, gs.bits, &B.cell_start);
    if (!B.sp) return ws.err;
    tile_aabb_kernel<<<div_up(ntiles, 8), 256, 0, st>>>(n, gs.ntb, ntiles, B.sp, B.tlo, B.thi);
    group_aabb_kernel<<<div_up(ngroups, 8), 256, 0, st>>>(gs.ntb, gs.ngb, ngroups, B.tlo, B.thi, B.glo, B.ghi);
    B.self = (new_xyz == xyz && m == n) ? 1 : 0;
    B.sq = B.sp;
    B.mpad = gs.npad;
    if (!B.self) {
        // queries sorted along the same curve so that consecutive warps touch the same tiles (L1 reuse)
        B.mpad = div_up(m, GT) * GT;
        B.sq = sort_points(ws, nb, m, B.mpad, new_xyz, B.bb, gs.bits, nullptr);
        if (!B.sq) return ws.err;
    }
    return cudaSuccess;
}

// ---------------------------------------------------------------------------------------
// 5. furthest point sampling with exact culling, for scenes too large for the register-resident cluster kernel
//    (fps.cu): one CTA per scene over the sorted tiles.
// ---------------------------------------------------------------------------------------
// A pick can only lower the running distance t[p] of points closer to it than sqrt(t[p]).  Per tile of GT sorted
// points the CTA keeps the tile's maximum of t (with the reference's tie key and that point's coordinates); a
// pick skips every tile whose box is farther from it than that maximum — conservatively, lb2 * (1 - 2^-13) > tmax,
// the same bound as the searches — because none of its points can change; skipped updates would have been
// t = min(t, d) with d > t, so every t, and with it every pick, is bit-identical to the reference's sequential
// scan (sampling_gpu.cu:100-216).  Measured: 4.2 tile updates per pick at 40 000 ... 400 000 points — the work
// per pick no longer grows with n — but a pick is then a chain of three block barriers and two L2 round trips,
// 1.9 - 2.8 us, which loses to the register-resident cluster kernel (0.2 - 1.5 us per pick) wherever that one
// fits (n <= 212 992).  It therefore serves the scenes beyond it, in place of the all-points global kernel.
// Variants measured and dropped: test + update fused per warp without a work list (2.8 us at 64 000 points:
// the hits of one warp serialise), 4 warps with the boxes in shared memory (3.1 us).
constexpr int FC_THREADS = 1024;
constexpr int FC_WARPS = FC_THREADS / 32;
__device__ unsigned long long g_fc_hits[2];     // debug (AMC3D_FPS_DEBUG): tile updates, picks

__device__ __forceinline__ uint32_t fps_tie_key(int k, int log2bs) {       // lower wins; as fps.cu tie_key
    const uint32_t low = (uint32_t)k & ((1u << log2bs) - 1u);
    const uint32_t rev = log2bs == 0 ? 0u : (__brev(low) >> (32 - log2bs));
    return (rev << 22) | ((uint32_t)k >> log2bs);
}

template <bool FC_DEBUG>
__global__ void __launch_bounds__(FC_THREADS)
fps_culled_kernel(int n, int npad, int ntb, int m, int log2bs, const float *__restrict__ xyz,
                  const float4 *__restrict__ sp, const float4 *__restrict__ tlo, const float4 *__restrict__ thi,
                  float *__restrict__ tsort, float *__restrict__ temp, int *__restrict__ idxs, int ibase, int istride) {
    extern __shared__ __align__(16) unsigned char fc_smem[];
    float4 *s_tpt = reinterpret_cast<float4 *>(fc_smem);           // per tile: its farthest point {x, y, z, original index}
    float *s_tmax = reinterpret_cast<float *>(s_tpt + ntb);        //           that point's running distance
    uint32_t *s_ttie = reinterpret_cast<uint32_t *>(s_tmax + ntb); //           and tie key
    int *s_work = reinterpret_cast<int *>(s_ttie + ntb);           // tiles the current pick has to update
    __shared__ uint32_t s_wv[FC_WARPS], s_wt[FC_WARPS];
    __shared__ int s_wi[FC_WARPS];
    __shared__ int s_nwork, s_old;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.x;
    xyz += 3ll * b * n;
    temp += (long long)b * n;
    idxs += (long long)b * m;
    sp += (long long)b * npad;
    tsort += (long long)b * npad;
    tlo += (long long)b * ntb;
    thi += (long long)b * ntb;
    ibase += b * istride;
    // running distances in sorted order, as the caller initialised them (1e10)
    for (int i = tid; i < n; i += FC_THREADS) tsort[i] = temp[__float_as_int(sp[i].w)];
    for (int t = tid; t < ntb; t += FC_THREADS) s_tmax[t] = INFINITY;     // every tile takes part in the first update
    float px = __ldg(xyz), py = __ldg(xyz + 1), pz = __ldg(xyz + 2);       // first pick: index 0
    if (tid == 0) {
        idxs[0] = ibase;
        s_nwork = 0;
    }
    __syncthreads();
    for (int j = 1; j < m; ++j) {
        // A. which tiles can change
        for (int t = tid; t < ntb; t += FC_THREADS) {
            const float lb = point_box2(px, py, pz, __ldg(tlo + t), __ldg(thi + t));
            if (!(lb * KG_SAFE > s_tmax[t])) s_work[atomicAdd(&s_nwork, 1)] = t;
        }
        __syncthreads();
        // B. update them, a warp per tile
        const int nwork = s_nwork;
        if (FC_DEBUG && tid == 0) {
            atomicAdd(&g_fc_hits[0], (unsigned long long)nwork);
            atomicAdd(&g_fc_hits[1], 1ull);
        }
        for (int w = warp; w < nwork; w += FC_WARPS) {
            const int th = s_work[w];
            uint32_t bv = 0, bt = 0xffffffffu;
            float4 bp = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int r = 0; r < GR; ++r) {
                const int i = th * GT + r * 32 + lane;
                if (i < n) {
                    const float4 p = __ldg(sp + i);
                    // operand order of the reference: point - last pick (sampling_gpu.cu:147-150)
                    const float d = dist2_ref(p.x - px, p.y - py, p.z - pz);
                    const float t2 = fminf(d, tsort[i]);
                    tsort[i] = t2;
                    const uint32_t v = __float_as_uint(t2), tk = fps_tie_key(__float_as_int(p.w), log2bs);
                    if (v > bv || (v == bv && tk < bt)) { bv = v; bt = tk; bp = p; }
                }
            }
            const uint32_t wv = __reduce_max_sync(0xffffffffu, bv);
            const uint32_t wt = __reduce_min_sync(0xffffffffu, bv == wv ? bt : 0xffffffffu);
            if (bv == wv && bt == wt) {
                s_tmax[th] = __uint_as_float(wv);
                s_ttie[th] = wt;
                s_tpt[th] = bp;
            }
        }
        __syncthreads();
        // C. argmax over the tile maxima (value desc, tie key asc)
        uint32_t bv = 0, bt = 0xffffffffu;
        int bi = 0;
        for (int t = tid; t < ntb; t += FC_THREADS) {
            const uint32_t v = __float_as_uint(s_tmax[t]), tk = s_ttie[t];
            if (v > bv || (v == bv && tk < bt)) { bv = v; bt = tk; bi = t; }
        }
        uint32_t wv = __reduce_max_sync(0xffffffffu, bv);
        uint32_t wt = __reduce_min_sync(0xffffffffu, bv == wv ? bt : 0xffffffffu);
        if (bv == wv && bt == wt) { s_wv[warp] = wv; s_wt[warp] = wt; s_wi[warp] = bi; }
        __syncthreads();
        if (warp == 0) {
            const uint32_t v = s_wv[lane], t2 = s_wt[lane];
            wv = __reduce_max_sync(0xffffffffu, v);
            wt = __reduce_min_sync(0xffffffffu, v == wv ? t2 : 0xffffffffu);
            if (v == wv && t2 == wt) s_old = s_wi[lane];
            if (lane == 0) s_nwork = 0;
        }
        __syncthreads();
        const float4 pk = s_tpt[s_old];
        px = pk.x; py = pk.y; pz = pk.z;
        if (tid == 0) idxs[j] = __float_as_int(pk.w) + ibase;
    }
    __syncthreads();
    // the caller's temp holds the final running distances, in original order, as the reference leaves them
    for (int i = tid; i < n; i += FC_THREADS) temp[__float_as_int(sp[i].w)] = tsort[i];
}

// b scenes of n points; returns 0, a cudaError_t, or -1 when this path does not apply (caller falls back)
int fps_culled_launch(int b, int n, int m, int log2bs, const float *xyz, float *temp, int *idx, int ibase, int istride,
                      cudaStream_t st) {
    if (n > 1000000 || b > 65535) return -1;         // tile state of a scene lives in shared memory (28 B per 128 points)
    Scratch ws(st);
    Built B;
    cudaError_t e = build(ws, b, n, n, xyz, xyz, B);
    if (e != cudaSuccess) return (int)e;
    float *tsort = ws.get<float>((size_t)b * B.gs.npad);
    if (ws.err != cudaSuccess) return (int)ws.err;
    const size_t smem = 28 * (size_t)B.gs.ntb;
    if (smem > 40 * 1024) {
        e = cudaFuncSetAttribute(fps_culled_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(fps_culled_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    static const bool dbg = getenv("AMC3D_FPS_DEBUG") != nullptr;
    if (dbg) {
        unsigned long long z[2] = {0, 0}, h[2];
        cudaMemcpyToSymbol(g_fc_hits, z, sizeof(z));
        fps_culled_kernel<true><<<b, FC_THREADS, smem, st>>>(n, B.gs.npad, B.gs.ntb, m, log2bs, xyz, B.sp, B.tlo, B.thi, tsort,
                                                            temp, idx, ibase, istride);
        cudaStreamSynchronize(st);
        cudaMemcpyFromSymbol(h, g_fc_hits, sizeof(h));
        fprintf(stderr, "[fps_culled] n=%d m=%d tiles=%d: %.2f tile updates per pick\n", n, m, B.gs.ntb, (double)h[0] / (double)(h[1] ? h[1] : 1));
        return (int)cudaGetLastError();
    }
    fps_culled_kernel<false><<<b, FC_THREADS, smem, st>>>(n, B.gs.npad, B.gs.ntb, m, log2bs, xyz, B.sp, B.tlo, B.thi, tsort,
                                                         temp, idx, ibase, istride);
    return (int)cudaGetLastError();
}

static void search_grid(const Built &B, int m, int &qpw, int &blocks) {
    // enough warps to fill the machine a few times over; long runs of consecutive queries per warp
    // only when there are plenty of queries
    const long long slots = (long long)B.gs.nb * B.mpad;
    static const int qmax = getenv("AMC3D_KNN_QPW") ? atoi(getenv("AMC3D_KNN_QPW")) : WQ_QPW_MAX;
    qpw = (int)max(1ll, min((long long)qmax, slots / (kNumSMs * WQ_WARPS * 8)));
    blocks = (int)div_up_ll(slots, WQ_WARPS * qpw);
}

// exact kNN with culling over nb clouds of n support / m query points; returns 0 or a cudaError_t.
// idx (nb,m,nsample) holds indices local to the cloud, dist2 the squared distances.
int knn_grid_batched(int nb, int n, int m, int nsample, const float *xyz, const float *new_xyz, int *idx,
                     float *dist2, cudaStream_t st, int *order_out) {
    Scratch ws(st);
    Built B;
    cudaError_t e = build(ws, nb, n, m, xyz, new_xyz, B);
    if (e != cudaSuccess) return (int)e;
    int qpw, blocks;
    search_grid(B, m, qpw, blocks);
    unsigned long long *const stats = g_search_stats;
#define KNN_WQ_ARGS B.gs, m, B.mpad, nsample, B.self, qpw, B.sp, B.tlo, B.thi, B.glo, B.ghi, B.sq, B.cell_start, B.bb, idx, dist2, order_out, stats
#define KNN_TQ_ARGS B.gs, m, B.mpad, B.self, B.sp, B.tlo, B.thi, B.glo, B.ghi, B.sq, B.cell_start, B.bb, idx, dist2, order_out, stats
    // Few neighbours and queries at least as dense as the support (three_nn onto the next finer level: 4
    // queries per known point; self searches with k <= 4): thread per query.  Measured at config-2 shapes
    // (Hilbert order): 192 000-query three_nn 0.19 ms against 0.39 ms; self search of 192 000 / 96 000 points
    // with k = 3: 0.26 / 0.15 ms against 0.31 / 0.19 ms.  The warp-per-query kernel wins for queries sparser
    // than the support (k = 4 label vote, 48 000 queries in 192 000 points: 0.19 against 0.30 ms), below
    // ~2 000 query warps (the slowest warp sets the time) and for k > 4 (k = 8: 0.84 against 0.53 ms).
    // AMC3D_KNN_TQ_MAX = 0 switches it off for measurements.
    static const int tq_max = getenv("AMC3D_KNN_TQ_MAX") ? atoi(getenv("AMC3D_KNN_TQ_MAX")) : 4;
    if (nsample <= tq_max && nsample <= 4 && m >= n && (long long)nb * B.mpad >= 64 * 1024) {
        const int tblocks = (int)div_up_ll((long long)nb * B.mpad, TQ_WARPS * 32);
        if (stats != nullptr) {                       // measurement run: the counting instantiations
            switch (nsample) {
                case 1: knn_tq_kernel<1, true><<<tblocks, TQ_WARPS * 32, 0, st>>>(KNN_TQ_ARGS); break;
                case 2: knn_tq_kernel<2, true><<<tblocks, TQ_WARPS * 32, 0, st>>>(KNN_TQ_ARGS); break;
                case 3: knn_tq_kernel<3, true><<<tblocks, TQ_WARPS * 32, 0, st>>>(KNN_TQ_ARGS); break;
                default: knn_tq_kernel<4, true><<<tblocks, TQ_WARPS * 32, 0, st>>>(KNN_TQ_ARGS); break;
            }
            return (int)cudaGetLastError();
        }
        switch (nsample) {
            case 1: knn_tq_kernel<1><<<tblocks, TQ_WARPS * 32, 0, st>>>(KNN_TQ_ARGS); break;
            case 2: knn_tq_kernel<2><<<tblocks, TQ_WARPS * 32, 0, st>>>(KNN_TQ_ARGS); break;
            case 3: knn_tq_kernel<3><<<tblocks, TQ_WARPS * 32, 0, st>>>(KNN_TQ_ARGS); break;
            default: knn_tq_kernel<4><<<tblocks, TQ_WARPS * 32, 0, st>>>(KNN_TQ_ARGS); break;
        }
        return (int)cudaGetLastError();
    }
    if (stats != nullptr) {
        if (nsample <= 32) knn_wq_kernel<1, true><<<blocks, WQ_WARPS * 32, 0, st>>>(KNN_WQ_ARGS);
        else if (nsample <= 64) knn_wq_kernel<2, true><<<blocks, WQ_WARPS * 32, 0, st>>>(KNN_WQ_ARGS);
        else knn_wq_kernel<4, true><<<blocks, WQ_WARPS * 32, 0, st>>>(KNN_WQ_ARGS);
        return (int)cudaGetLastError();
    }
    if (nsample <= 32) knn_wq_kernel<1><<<blocks, WQ_WARPS * 32, 0, st>>>(KNN_WQ_ARGS);
    else if (nsample <= 64) knn_wq_kernel<2><<<blocks, WQ_WARPS * 32, 0, st>>>(KNN_WQ_ARGS);
    else knn_wq_kernel<4><<<blocks, WQ_WARPS * 32, 0, st>>>(KNN_WQ_ARGS);
    return (int)cudaGetLastError();
}

int knn_grid_single_segment(int n, int m, int nsample, const float *xyz, const float *new_xyz, int *idx,
                            float *dist2, cudaStream_t st, int *order_out) {
    return knn_grid_batched(1, n, m, nsample, xyz, new_xyz, idx, dist2, st, order_out);
}

// ball query with culling over nb clouds; idx (nb,m,nsample) pre-zeroed by the caller
int ball_grid_batched(int nb, int n, int m, float radius, int nsample, const float *xyz, const float *new_xyz,
                      int *idx, cudaStream_t st) {
    Scratch ws(st);
    Built B;
    cudaError_t e = build(ws, nb, n, m, xyz, new_xyz, B);
    if (e != cudaSuccess) return (int)e;
    int qpw, blocks;
    search_grid(B, m, qpw, blocks);
    unsigned long long *const stats = g_search_stats;
#define BALL_WQ_ARGS B.gs, m, B.mpad, radius, nsample, B.self, qpw, B.sp, B.tlo, B.thi, B.glo, B.ghi, B.sq, B.cell_start, B.bb, idx, stats
    if (stats != nullptr) {
        if (nsample <= 32) ball_wq_kernel<1, true><<<blocks, WQ_WARPS * 32, 0, st>>>(BALL_WQ_ARGS);
        else if (nsample <= 64) ball_wq_kernel<2, true><<<blocks, WQ_WARPS * 32, 0, st>>>(BALL_WQ_ARGS);
        else ball_wq_kernel<4, true><<<blocks, WQ_WARPS * 32, 0, st>>>(BALL_WQ_ARGS);
        return (int)cudaGetLastError();
    }
    if (nsample <= 32) ball_wq_kernel<1><<<blocks, WQ_WARPS * 32, 0, st>>>(BALL_WQ_ARGS);
    else if (nsample <= 64) ball_wq_kernel<2><<<blocks, WQ_WARPS * 32, 0, st>>>(BALL_WQ_ARGS);
    else ball_wq_kernel<4><<<blocks, WQ_WARPS * 32, 0, st>>>(BALL_WQ_ARGS);
    return (int)cudaGetLastError();
}

}  // namespace amc3d
