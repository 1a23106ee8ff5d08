// Exact kNN with spatial culling for the single-segment case (offset = [n]) — the case
// AMContrast3D always produces (pointnext_AA.py:461 flattens the batch into one segment).
//
// The brute-force kernel (knn.cu) evaluates all m*n pairs: 3.7e10 for the stage-0 self-kNN of
// BASELINE config 2, ~12 ms at the FP32 issue limit.  The result, however, is fully determined
// by the (d2, index)-lexicographic order, so any evaluation order that provably visits every
// point that can enter a query's top-k returns the identical idx / dist2.  This file does
// that (DESIGN.md "kNN: culled path"):
//
//   1. bbox            one pass, block reduce + ordered-int atomics
//   2. cell sort       points are binned into a 2^b x 2^b x 2^b grid, cells numbered along a
//                      Morton curve, and counting-sorted (histogram -> single-CTA scan ->
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

constexpr int GT = 128;                 // threads per CTA == points per tile
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

// bb[0..2] = min (ordered), bb[3..5] = max (ordered); caller initialises to 0xffffffff / 0
__global__ void __launch_bounds__(256)
bbox_kernel(int n, const float *__restrict__ xyz, uint32_t *__restrict__ bb) {
    __shared__ float s_lo[3][8], s_hi[3][8];
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

__global__ void bbox_init_kernel(uint32_t *bb) {
    if (threadIdx.x < 3) bb[threadIdx.x] = 0xffffffffu;
    else if (threadIdx.x < 6) bb[threadIdx.x] = 0u;
}

// ---------------------------------------------------------------------------------------
// 2. Morton cell of a point (clamped into the grid; NaN -> cell 0)
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t spread3(uint32_t v) {   // 10 bits -> every third bit
    v = (v | (v << 16)) & 0x030000ffu;
    v = (v | (v << 8)) & 0x0300f00fu;
    v = (v | (v << 4)) & 0x030c30c3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
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
    return spread3(c[0]) | (spread3(c[1]) << 1) | (spread3(c[2]) << 2);
}

__global__ void __launch_bounds__(256)
cell_count_kernel(int n, const float *__restrict__ xyz, const uint32_t *__restrict__ bb, int bits,
                  uint32_t *__restrict__ cell, int *__restrict__ counts) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const uint32_t c = cell_of(__ldg(xyz + 3ll * i), __ldg(xyz + 3ll * i + 1), __ldg(xyz + 3ll * i + 2), bb, bits);
    cell[i] = c;
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

// scatter into cell order; `starts` is consumed as the running fill pointer of each cell
__global__ void __launch_bounds__(256)
scatter_kernel(int n, const float *__restrict__ xyz, const uint32_t *__restrict__ cell,
               int *__restrict__ starts, float4 *__restrict__ sorted) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const int pos = atomicAdd(starts + cell[i], 1);
    sorted[pos] = make_float4(__ldg(xyz + 3ll * i), __ldg(xyz + 3ll * i + 1), __ldg(xyz + 3ll * i + 2),
                              __int_as_float(i));
}

// ---------------------------------------------------------------------------------------
// 3. tile boxes: one warp per tile of GT sorted points
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
tile_aabb_kernel(int n, int ntiles, const float4 *__restrict__ sorted, float4 *__restrict__ tlo,
                 float4 *__restrict__ thi) {
    const int t = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (t >= ntiles) return;
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int j = lane; j < GT; j += 32) {
        const int i = t * GT + j;
        if (i < n) {
            const float4 p = __ldg(sorted + i);
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
constexpr int WQ_QPW_MAX = 16;  // consecutive sorted queries per warp (tile reuse through L1)
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

// smallest T with  #{distances <= T} >= k  over the 4 x 32 values held by the warp (radix
// select on the bit pattern: non-negative floats order like their bits)
__device__ __forceinline__ float warp_kth_of_128(const float (&dd)[4], int k) {
    uint32_t bitsv[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) bitsv[r] = __float_as_uint(dd[r]);
    uint32_t T = 0;
#pragma unroll 1
    for (int b = 30; b >= 0; --b) {
        const uint32_t cand = T | (1u << b);
        int c = 0;
#pragma unroll
        for (int r = 0; r < 4; ++r) c += bitsv[r] < cand ? 1 : 0;
        c = __reduce_add_sync(0xffffffffu, c);
        if (c < k) T = cand;       // fewer than k values below cand: the k-th is >= cand
    }
    return __uint_as_float(T);
}

template <int E>
__global__ void __launch_bounds__(WQ_WARPS * 32)
knn_wq_kernel(int n, int m, int nsample, int ntiles, int ngroups, int self, int qpw, const float4 *__restrict__ sp,
              const float4 *__restrict__ tlo, const float4 *__restrict__ thi, const float4 *__restrict__ glo,
              const float4 *__restrict__ ghi, const float4 *__restrict__ sq, const int *__restrict__ cell_start,
              const uint32_t *__restrict__ bb, int bits, int *__restrict__ idx, float *__restrict__ dist2) {
    const int lane = threadIdx.x & 31;
    const int wq0 = (blockIdx.x * WQ_WARPS + (threadIdx.x >> 5)) * qpw;

    for (int j = 0; j < qpw; ++j) {
        const int q = wq0 + j;
        if (q >= m) return;                                   // warp-uniform
        const float4 me = __ldg(sq + q);
        const float qx = me.x, qy = me.y, qz = me.z;
        WarpList<E> best;
        best.init(nsample, lane);
        float td = KG_INIT;
        int ti = 0x7fffffff;

        // evaluate the 128 points of tile t against this query; first=true seeds the list
        auto process_tile = [&](int t, bool first) {
            float dd[4];
            int oi[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int s = t * GT + r * 32 + lane;
                if (s < n) {
                    const float4 p = __ldg(sp + s);
                    dd[r] = dist2_ref(qx - p.x, qy - p.y, qz - p.z);
                    oi[r] = __float_as_int(p.w);
                } else {
                    dd[r] = INFINITY;
                    oi[r] = 0x7fffffff;
                }
            }
            if (first) {
                // bulk seed: only the ~k nearest of the first tile go through the insertion
                td = fminf(warp_kth_of_128(dd, min(nsample, GT)), KG_INIT);
                ti = 0x7fffffff;
            }
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                bool cand = dd[r] <= td && lex_lt(dd[r], oi[r], td, ti);
                uint32_t mask = __ballot_sync(0xffffffffu, cand);
                while (mask) {
                    const int b = __ffs(mask) - 1;
                    const float cd = __shfl_sync(0xffffffffu, dd[r], b);
                    const int ci = __shfl_sync(0xffffffffu, oi[r], b);
                    best.insert(cd, ci, lane);
                    float nd;
                    int ni;
                    best.threshold(nd, ni);
                    // while seeding, the provisional (T, INT_MAX) bound stays until the list is full
                    if (!first || nd < KG_INIT) { td = nd; ti = ni; }
                    cand = cand && lane != b && lex_lt(dd[r], oi[r], td, ti);
                    mask = __ballot_sync(0xffffffffu, cand);
                }
            }
            if (first) best.threshold(td, ti);
        };

        // ---- the tile at the query's own position and its two neighbours first -------------
        int t0;
        if (self) t0 = q / GT;
        else t0 = __ldg(cell_start + cell_of(qx, qy, qz, bb, bits)) / GT;
        t0 = min(max(t0, 0), ntiles - 1);
        const int r_lo = max(t0 - 1, 0), r_hi = min(t0 + 1, ntiles - 1);
        process_tile(t0, true);
        for (int t = r_lo; t <= r_hi; ++t)
            if (t != t0) process_tile(t, false);

        // ---- everything else through two levels of boxes -----------------------------------
        for (int g0 = 0; g0 < ngroups; g0 += 32) {
            const int g = g0 + lane;
            bool gs = false;
            if (g < ngroups) gs = !(point_box2(qx, qy, qz, __ldg(glo + g), __ldg(ghi + g)) * KG_SAFE > td);
            uint32_t gmask = __ballot_sync(0xffffffffu, gs);
            while (gmask) {
                const int gb = __ffs(gmask) - 1;
                gmask &= gmask - 1;
                const int t = (g0 + gb) * TG + lane;
                float lb = INFINITY;
                if (t < ntiles && (t < r_lo || t > r_hi)) lb = point_box2(qx, qy, qz, __ldg(tlo + t), __ldg(thi + t));
                uint32_t tmask = __ballot_sync(0xffffffffu, !(lb * KG_SAFE > td));
                while (tmask) {
                    const int tb = __ffs(tmask) - 1;
                    tmask &= tmask - 1;
                    const float lbt = __shfl_sync(0xffffffffu, lb, tb);
                    if (lbt * KG_SAFE > td) continue;              // threshold tightened meanwhile
                    process_tile((g0 + gb) * TG + tb, false);
                }
            }
        }

        // ---- write the live entries (slots S-nsample .. S-1) to the query's original row ----
        const int qorig = __float_as_int(me.w);
        const int shift = 32 * E - nsample;
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int s = lane * E + e - shift;
            if (s >= 0) {
                idx[(long long)qorig * nsample + s] = best.i[e];
                dist2[(long long)qorig * nsample + s] = best.d[e];
            }
        }
    }
}

// boxes of TG consecutive tile boxes
__global__ void __launch_bounds__(256)
group_aabb_kernel(int ntiles, int ngroups, const float4 *__restrict__ tlo, const float4 *__restrict__ thi,
                  float4 *__restrict__ glo, float4 *__restrict__ ghi) {
    const int g = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (g >= ngroups) return;
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    const int t = g * TG + lane;
    if (t < ntiles) {
        const float4 a = __ldg(tlo + t), b = __ldg(thi + t);
        lo[0] = a.x; lo[1] = a.y; lo[2] = a.z;
        hi[0] = b.x; hi[1] = b.y; hi[2] = b.z;
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

// per-stream scratch, stream-ordered (cudaMallocAsync pools make this cheap after warm-up)
struct Scratch {
    cudaStream_t st;
    void *ptrs[16];
    int count = 0;
    cudaError_t err = cudaSuccess;
    explicit Scratch(cudaStream_t s) : st(s) {}
    template <class T>
    T *get(size_t n) {
        void *p = nullptr;
        if (err == cudaSuccess) err = cudaMallocAsync(&p, n * sizeof(T) + 16, st);
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

// sort `count` points into Morton-cell order; returns cell_start (exclusive starts) if wanted
static float4 *sort_points(Scratch &ws, int count, const float *xyz, const uint32_t *bb, int bits,
                           int **cell_start_out) {
    const int cells = 1 << (3 * bits);
    uint32_t *cell = ws.get<uint32_t>(count);
    int *counts = ws.get<int>(cells);
    int *fill = ws.get<int>(cells);
    float4 *sorted = ws.get<float4>(count);
    if (ws.err != cudaSuccess) return nullptr;
    cudaMemsetAsync(counts, 0, sizeof(int) * cells, ws.st);
    cell_count_kernel<<<div_up(count, 256), 256, 0, ws.st>>>(count, xyz, bb, bits, cell, counts);
    const int sblocks = div_up(cells, SCAN_CHUNK);
    int *totals = ws.get<int>(sblocks);
    if (ws.err != cudaSuccess) return nullptr;
    scan_local_kernel<<<sblocks, SCAN_THREADS, 0, ws.st>>>(cells, counts, totals);
    scan_totals_kernel<<<1, SCAN_THREADS, 0, ws.st>>>(sblocks, totals);
    scan_add_kernel<<<sblocks, SCAN_THREADS, 0, ws.st>>>(cells, counts, totals, fill);
    scatter_kernel<<<div_up(count, 256), 256, 0, ws.st>>>(count, xyz, cell, fill, sorted);
    if (cell_start_out) *cell_start_out = counts;
    return sorted;
}

template <int E>
static void launch_wq(cudaStream_t st, int n, int m, int nsample, int ntiles, int ngroups, int self, const float4 *sp,
                      const float4 *tlo, const float4 *thi, const float4 *glo, const float4 *ghi, const float4 *sq,
                      const int *cell_start, const uint32_t *bb, int bits, int *idx, float *dist2) {
    // enough warps to fill the machine a few times over; long runs of consecutive queries per warp
    // only when there are plenty of queries
    const int qpw = max(1, min(WQ_QPW_MAX, m / (kNumSMs * WQ_WARPS * 8)));
    const int blocks = div_up(m, WQ_WARPS * qpw);
    knn_wq_kernel<E><<<blocks, WQ_WARPS * 32, 0, st>>>(n, m, nsample, ntiles, ngroups, self, qpw, sp, tlo, thi, glo, ghi, sq,
                                                       cell_start, bb, bits, idx, dist2);
}

// single-segment exact kNN with culling; returns 0 or a cudaError_t
int knn_grid_single_segment(int n, int m, int nsample, const float *xyz, const float *new_xyz, int *idx,
                            float *dist2, cudaStream_t st) {
    Scratch ws(st);
    const int bits = pick_bits(n);
    const int ntiles = div_up(n, GT);
    const int ngroups = div_up(ntiles, TG);
    uint32_t *bb = ws.get<uint32_t>(8);
    float4 *tlo = ws.get<float4>(ntiles);
    float4 *thi = ws.get<float4>(ntiles);
    float4 *glo = ws.get<float4>(ngroups);
    float4 *ghi = ws.get<float4>(ngroups);
    if (ws.err != cudaSuccess) return (int)ws.err;
    bbox_init_kernel<<<1, 32, 0, st>>>(bb);
    bbox_kernel<<<min(div_up(n, 1024), kNumSMs), 256, 0, st>>>(n, xyz, bb);
    int *cell_start = nullptr;
    float4 *sp = sort_points(ws, n, xyz, bb, bits, &cell_start);
    if (!sp) return (int)ws.err;
    tile_aabb_kernel<<<div_up(ntiles, 8), 256, 0, st>>>(n, ntiles, sp, tlo, thi);
    group_aabb_kernel<<<div_up(ngroups, 8), 256, 0, st>>>(ntiles, ngroups, tlo, thi, glo, ghi);
    const int self = (new_xyz == xyz && m == n) ? 1 : 0;
    const float4 *sq = sp;
    if (!self) {
        sq = sort_points(ws, m, new_xyz, bb, bits, nullptr);
        if (!sq) return (int)ws.err;
    }
    if (nsample <= 32)
        launch_wq<1>(st, n, m, nsample, ntiles, ngroups, self, sp, tlo, thi, glo, ghi, sq, cell_start, bb, bits, idx, dist2);
    else if (nsample <= 64)
        launch_wq<2>(st, n, m, nsample, ntiles, ngroups, self, sp, tlo, thi, glo, ghi, sq, cell_start, bb, bits, idx, dist2);
    else
        launch_wq<4>(st, n, m, nsample, ntiles, ngroups, self, sp, tlo, thi, glo, ghi, sq, cell_start, bb, bits, idx, dist2);
    cudaError_t e = cudaGetLastError();
    return (int)e;
}

}  // namespace amc3d
