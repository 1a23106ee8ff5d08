// Fused  grouping -> 1x1 conv -> BatchNorm (training statistics) -> ReLU -> max over the neighbourhood
// (SURVEY.md §8f rank 1): the operator PointNeXt's SetAbstraction / LocalAggregation compose from
// QueryAndGroup + Conv2d + BatchNorm2d + ReLU + max  (ref: openpoints/models/backbone/pointnext_AA.py:57-63,
// :139-170; grouper openpoints/models/layers/group.py:235-255).  The grouped tensor (B, 3+C, M, ns) — 6.45 GB
// written and re-read per step at BASELINE config 2 — is never materialised.
//
//   forward    y[o, p] = sum_k W'[o, k] x[p, k]        x[p] = [ f[b, idx[p], 0..C) | (xyz[idx[p]] - q) / r | 0 ]
//              as tcgen05 MMAs (kind::tf32, FP32 accumulators in TMEM):  M = 128 output channels, N = 256 grouped
//              positions (= 256 / ns queries; 128 in the 3xTF32 mode), K = C + 8 in 128-byte chunks.  The B operand
//              (positions x K) is GATHERED: 8 lanes fetch one neighbour's 128 contiguous bytes of the
//              channel-contiguous (B, N, C) feature copy with cp.async straight into the 128B-swizzled K-major tile
//              the MMA reads; the A operand (W') arrives as ready-made tiles by bulk copy.  The epilogue thread that
//              owns TMEM lane o holds the ns conv outputs of a query in registers, so max / arg-max over the
//              neighbourhood and the BatchNorm sums are thread-local.  BatchNorm + ReLU are monotone in y
//              (increasing for gamma >= 0, decreasing otherwise), hence  max_s relu(bn(y_s)) = relu(bn(max_s y_s))
//              (min for gamma < 0): ONE pass over the gathered rows gives the pre-normalisation extreme per
//              (query, channel) and the statistics; a small second kernel normalises.
//   precision  X3 = false: operands are read as TF32 by the tensor core (what the reference's cuDNN convolution
//              does on Ampere and later: torch.backends.cudnn.allow_tf32 defaults to True);
//              X3 = true: error-compensated 3 x TF32 (hi*hi + lo*hi + hi*lo, hi = rna(x), lo = x - hi), FP32-faithful
//              to ~1e-6 — what the CPU-generated golden vectors are held to (2e-5).
#include "common.cuh"
#include <stdlib.h>
#include <type_traits>

namespace amc3d {

constexpr int FS_MT = 128;                 // output channels per work item (UMMA M = TMEM lanes)
constexpr int FS_ROWB = 128;               // bytes per tile row = one swizzle span = 32 floats of K
constexpr int FS_PW = 8;                   // producer warps 0 .. FS_PW-1
constexpr int FS_PROD = FS_PW * 32;        // producer threads
constexpr int FS_WARP_MMA = FS_PW;         // issues the MMAs
constexpr int FS_WARP_DP = FS_PW + 1;      // writes the relative-coordinate columns (TF32 path)
constexpr int FS_WARP_EPI = FS_PW + 2;     // first of the 8 epilogue warps
constexpr int FS_EPI = 256;                // epilogue threads: two per TMEM lane (each takes every other query of the tile)
constexpr int FS_RPP = FS_PROD / 8;        // tile rows filled per pass (8 lanes per 128-byte row)
constexpr int FS_THREADS = FS_PROD + 64 + FS_EPI;
constexpr int FS_MAX_STAGES = 8;
constexpr int FS_REPL = 64;                // replicas of the BatchNorm sums (spreads the atomics over 64x the cache lines)

// n / d for n < 2^31 without a division: q = umulhi(n, m) >> s with m = ceil(2^(31+l) / d), l = ceil(log2 d)
struct FastDiv {
    uint32_t m, s, one;
};
static inline FastDiv make_fastdiv(uint32_t d) {
    FastDiv f{0u, 0u, 1u};
    if (d <= 1) return f;
    uint32_t l = 0;
    while ((1ull << l) < d) ++l;
    f.m = (uint32_t)(((1ull << (31 + l)) + d - 1) / d);
    f.s = l - 1;
    f.one = 0;
    return f;
}
__device__ __forceinline__ uint32_t fdiv(uint32_t n, const FastDiv &f) { return f.one ? n : (__umulhi(n, f.m) >> f.s); }

__device__ __forceinline__ uint32_t fs_smem(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Bounded: a protocol error must surface as a failed launch (cudaErrorLaunchFailure through amc3d's error path), never
// as a kernel that spins for ever.  try_wait suspends the thread for a hardware time slice per attempt, so 2^24 failed
// attempts are many seconds — four orders of magnitude beyond the longest legitimate wait in this kernel.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0, tries = 0;
    while (!ok) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
        if (!ok && ++tries > (1u << 24)) __trap();
    }
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem], 128 x N x 8 (TF32), issued by one thread for the CTA
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(acc)
        : "memory");
}
// arrive on an mbarrier once every MMA issued so far has completed (implies fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// K-major operand tile, 128-byte swizzle: rows of 128 bytes, 8-row groups 1024 bytes apart (SBO), tile base
// 1024-byte aligned; advancing along K inside the swizzle span = adding bytes to the start address.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    uint64_t d = (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;                  // leading byte offset (unused for swizzled K-major), 16 B
    d |= (uint64_t)(1024 >> 4) << 32;        // stride byte offset: 8 rows x 128 B
    d |= (uint64_t)1 << 46;                  // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                  // SWIZZLE_128B
    return d;
}
// instruction descriptor: D = F32, A = B = TF32, both K-major, M x N
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// 16-byte slot `c` (0..7) of row `r` inside a 128B-swizzled tile
__device__ __forceinline__ uint32_t sw128_off(int r, int c) { return (uint32_t)(r * FS_ROWB + ((c ^ (r & 7)) << 4)); }

__device__ __forceinline__ float tf32_rna(float x) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return __uint_as_float(u);
}

// 16-byte asynchronous global -> shared copy (LDGSTS); `valid` false writes zeros without reading
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src, bool valid) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(valid ? 16u : 0u) : "memory");
}


template <bool X3>
__device__ __forceinline__ void fs_store(unsigned char *hi, unsigned char *lo, uint32_t off, float4 v) {
    if (X3) {
        float4 h = make_float4(tf32_rna(v.x), tf32_rna(v.y), tf32_rna(v.z), tf32_rna(v.w));
        *reinterpret_cast<float4 *>(hi + off) = h;
        *reinterpret_cast<float4 *>(lo + off) = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
    } else {
        *reinterpret_cast<float4 *>(hi + off) = v;
    }
}

template <int NS>
__device__ __forceinline__ void tmem_ld_query(uint32_t taddr, float (&v)[NS]);

template <>
__device__ __forceinline__ void tmem_ld_query<32>(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

template <>
__device__ __forceinline__ void tmem_ld_query<16>(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// W' (O, Kp) -> operand tiles: block (slice j, chunk kc) holds rows j*128 .. j*128+127, columns kc*32 .. kc*32+31 with the
// 16-byte slot c of row r at r*128 + ((c ^ (r & 7)) << 4), zeros outside the matrix; X3 adds the lo = x - rna(x) tile.
// Rows with gamma < 0 are NEGATED (exact): the accumulators then hold t = sign(gamma) y, whose maximum over the
// neighbourhood is what the pooled output needs for either sign, and the epilogue has no per-element sign to apply.
template <bool X3>
__global__ void __launch_bounds__(256)
fused_sa_wprep_kernel(int O, int Kp, int nchunks, const float *__restrict__ Wp, const float *__restrict__ gamma,
                      float *__restrict__ Wsw) {
    const int blk = blockIdx.x, j = blk / nchunks, kc = blk % nchunks;
    unsigned char *dst = reinterpret_cast<unsigned char *>(Wsw) + (size_t)blk * (X3 ? 2 : 1) * FS_MT * FS_ROWB;
    for (int e = threadIdx.x; e < FS_MT * 8; e += 256) {
        const int r = e >> 3, c = e & 7;
        const int o = j * 128 + r, k0 = kc * 32 + c * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (o < O && k0 < Kp) {
            v = __ldg(reinterpret_cast<const float4 *>(Wp + (long long)o * Kp + k0));
            if (__ldg(gamma + o) < 0.f) v = make_float4(-v.x, -v.y, -v.z, -v.w);
        }
        fs_store<X3>(dst, dst + FS_MT * FS_ROWB, sw128_off(r, c), v);
    }
}

struct FusedFwdArgs {
    const float *fT;        // (B, N, C) channel-contiguous features
    const float *xyz;       // (B, N, 3) support points
    const float *qxyz;      // (B, M, 3) query points
    const int *idx;         // (B, M, NS)
    const float *Wp;        // (O, Kp): [W[:, 3:3+C] | W[:, 0:3] | 0]
    const float *Wsw;       // the same weights as ready-made operand tiles: block (slice, chunk) = 128 rows x 128 B in the
                            // swizzled layout the MMA reads ([hi | lo] halves for 3xTF32), so a chunk is ONE bulk copy
    const float *gamma;     // (O) BatchNorm weight: its sign selects max or min
    float *ysel;            // (B*M, O) pre-normalisation extreme of y over the neighbourhood
    unsigned char *arg;     // (B*M, O) sample index of that extreme (first one)
    double *gsum, *gsumsq;  // FS_REPL x (O) sum y, sum y^2 over all B*M*NS positions (zeroed by the caller)
    long long *dbg;         // AMC3D_FUSED_DBG: per CTA 12 cycle counters (where each role waits), else NULL
    int B, N, M, C, O, Kp;
    int stages;             // shared-memory ring depth (2..FS_MAX_STAGES)
    int w_resident;         // 1: the whole (128 x Kp) weight slice stays in shared memory for the CTA's lifetime
    int nslices;            // ceil(O / 128): output-channel slices per position tile
    int nitems;             // position tiles x slices
    FastDiv div_m, div_s;   // division by M (query -> batch) and by nslices (item -> tile)
    float inv_radius;       // 1/radius with normalize_dp, else 1
};

// Persistent, warp-specialised, one CTA per SM.  A work item = (tile of NT grouped positions, slice of 128 output
// channels); a CTA takes items blockIdx.x, blockIdx.x + gridDim.x, ...  NT = 256 (TF32; the widest UMMA) or 128 (3xTF32).
//   warps 0-7   PRODUCE the gathered operand: a ring of `stages` slots, one K-chunk (32 floats) of the tile per slot
//   warp  8     ISSUES the MMAs into one of two TMEM accumulators (one elected lane)
//   warp  9     TF32 path: computes the relative coordinates (idx -> xyz -> (p - q)/r, a dependent chain of global
//               loads) ahead of the ring and stores them into the item's last chunk
//   warps 10-17 run the EPILOGUE of the previous item out of the other accumulator
// Barriers: full[s] / empty[s] per ring slot (producers <-> MMA), acc_full[b] / acc_empty[b] per accumulator
// (MMA <-> epilogue), w_full for the resident weight slice.  In the TF32 path a producer thread's arrival on full[s] is
// asynchronous (cp.async.mbarrier.arrive.noinc): it fires when the thread's copies of the chunk have landed, so nobody
// waits for them and a chunk reaches the MMA warp as early as it can.
// The single-thread loops (producer set-up, MMA issue) are latency chains of scalar instructions, so they are kept
// short: no divisions (FastDiv), 32-bit element offsets, descriptors advanced by additions.
template <int NS, bool X3>
__global__ void __launch_bounds__(FS_THREADS, 1)
fused_sa_fwd_kernel(const FusedFwdArgs a) {
    constexpr int NT = X3 ? 128 : 256;                    // grouped positions per item (UMMA N)
    constexpr int QPT = NT / NS;                          // queries per tile
    constexpr int NPASS = NT / FS_RPP;                    // passes of the 256 producer threads over the tile rows
    constexpr uint32_t XT_BYTES = (X3 ? 2u : 1u) * NT * FS_ROWB;       // one gathered chunk (hi [+ lo]): NT rows x 128 B
    constexpr uint32_t WT_BYTES = (X3 ? 2u : 1u) * FS_MT * FS_ROWB;    // one weight chunk: 128 rows x 128 B
    extern __shared__ __align__(1024) unsigned char fs_smem_raw[];
    // the runtime only guarantees 16-byte alignment of dynamic shared memory: align by hand
    unsigned char *base = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(fs_smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t bars[2 * FS_MAX_STAGES + 9];
    __shared__ uint32_t tmem_base_s;
    const int S = a.stages;
    const int nchunks = (a.Kp + 31) / 32;
    const bool wres = a.w_resident != 0;
    // layout: [resident W: nchunks tiles] then the ring; a ring slot is [X tile] or [X tile | W tile]
    unsigned char *ring = base + (wres ? (size_t)nchunks * WT_BYTES : 0);
    const uint32_t slot_bytes = wres ? XT_BYTES : XT_BYTES + WT_BYTES;
    uint64_t *full = bars, *empty = bars + FS_MAX_STAGES, *acc_full = bars + 2 * FS_MAX_STAGES,
             *acc_empty = bars + 2 * FS_MAX_STAGES + 2, *w_full = bars + 2 * FS_MAX_STAGES + 4,
             *dp_ready = bars + 2 * FS_MAX_STAGES + 5, *dp_free = bars + 2 * FS_MAX_STAGES + 7;
    // TF32 path: two staging buffers of NT float4 (dp, 0) behind the ring, filled by the dp warp one or two items ahead
    float4 *dp_stage = reinterpret_cast<float4 *>(ring + (size_t)S * slot_bytes);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int Q = a.B * a.M;                               // queries (host checks B*M < 2^25)
    const int nitems = a.nitems, nsl = a.nslices;
    const int kc_dp = a.C >> 5, dp_slot = (a.C & 31) >> 2; // chunk and 16-byte slot that hold (dp, 0): C % 4 == 0

    if (tid == 0) {
        for (int s = 0; s < FS_MAX_STAGES; ++s) {
            mbar_init(fs_smem(&full[s]), FS_PROD);                 // every producer thread arrives
            mbar_init(fs_smem(&empty[s]), 1);                      // one tcgen05.commit
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(fs_smem(&acc_full[b]), 1);                   // one tcgen05.commit
            mbar_init(fs_smem(&acc_empty[b]), FS_EPI);             // every epilogue thread arrives
        }
        mbar_init(fs_smem(w_full), 1);                             // one arrive.expect_tx; the bulk copies complete it
        for (int b = 0; b < 2; ++b) {
            mbar_init(fs_smem(&dp_ready[b]), 32);                  // the dp warp's lanes
            mbar_init(fs_smem(&dp_free[b]), 32);                   // the 32 producer threads of the (dp, 0) column
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == FS_WARP_MMA) {                                     // TMEM: two accumulators of 128 lanes x NT columns
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(fs_smem(&tmem_base_s)),
                     "r"((uint32_t)(2 * NT))
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    if (warp < FS_PW) {
        // ================================================================== producers
        const int slot = tid & 7, rsub = tid >> 3;                 // 8 lanes per 128-byte row, FS_RPP rows per pass
        // rows rsub + 32 i share (row & 7): one swizzled offset, passes 4096 bytes apart
        const uint32_t soff0 = sw128_off(rsub, slot);
        // neighbour index of the positions this thread fetches in an item, -1 past the end (raw loads: nothing depends
        // on them until the item starts, so they are in flight during the previous item's chunks)
        auto load_idx = [&](int item, int (&nidx)[NPASS]) {
            const int q0 = (int)fdiv((uint32_t)item, a.div_s) * QPT;
#pragma unroll
            for (int i = 0; i < NPASS; ++i) {
                const int r = i * FS_RPP + rsub;
                const int qg = q0 + r / NS;
                nidx[i] = qg < Q ? __ldg(a.idx + (long long)qg * NS + (r % NS)) : -1;
            }
        };
        // weights: ready-made tiles in global memory (fused_sa_wprep_kernel) -> one bulk copy per chunk, issued by one
        // thread, completing on the consumer's barrier by byte count: the producer warps never touch them
        auto bulk_w = [&](uint32_t dst, int slice, int kc, uint32_t bar) {
            const unsigned char *src = reinterpret_cast<const unsigned char *>(a.Wsw) + ((size_t)slice * nchunks + kc) * WT_BYTES;
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(dst), "l"(src), "r"(WT_BYTES), "r"(bar)
                         : "memory");
        };
        if (wres && tid == 0) {                                    // the weight slice, once (a.nslices == 1)
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fs_smem(w_full)),
                         "r"((uint32_t)nchunks * WT_BYTES)
                         : "memory");
            for (int kc = 0; kc < nchunks; ++kc) bulk_w(fs_smem(base + (size_t)kc * WT_BYTES), 0, kc, fs_smem(w_full));
        }
        // the thread that announces a chunk's weight bytes (expect_tx) must arrive on the same barrier afterwards, or the
        // phase could complete before the announcement: pick one that is never the (dp, 0) slot
        const int wtid = X3 ? 0 : ((dp_slot + 1) & 7);
        int nxt[NPASS];
        if ((int)blockIdx.x < nitems) load_idx((int)blockIdx.x, nxt);
        int st = 0, pass = 0;                                      // ring slot of the next chunk and the ring's pass count
        long long t_we = 0, t_set = 0;
        const long long t_begin = a.dbg ? clock64() : 0;
        int itl = 0;                                               // items done by this CTA
        for (int item = blockIdx.x; item < nitems; item += (int)gridDim.x, ++itl) {
            const long long ts0 = a.dbg ? clock64() : 0;
            const uint32_t tile = fdiv((uint32_t)item, a.div_s);
            const int sl = item - (int)tile * nsl;
            const int q0 = (int)tile * QPT;
            // element offset of this thread's 16 bytes in each of its rows (host checks B*N*C < 2^31)
            uint32_t xoff[NPASS];
            uint32_t rowv[X3 ? NPASS : 1];                         // support row b*N + n (3xTF32 path: coordinates inline)
            uint32_t livemask = 0;
#pragma unroll
            for (int i = 0; i < NPASS; ++i) {
                const int qg = q0 + (i * FS_RPP + rsub) / NS;
                const uint32_t bb = fdiv((uint32_t)qg, a.div_m);
                const bool live = nxt[i] >= 0;
                const uint32_t row = live ? bb * (uint32_t)a.N + (uint32_t)nxt[i] : 0u;
                xoff[i] = row * (uint32_t)a.C + (uint32_t)(slot * 4);
                if (X3) rowv[i] = row;
                livemask |= (live ? 1u : 0u) << i;
            }
            // TF32 path: global address of this thread's 16 bytes of the NEXT chunk in each of its rows (advanced by 128
            // bytes per chunk) and the copy size that makes a dead row a zero-fill: the chunk loop is copy + add
            unsigned long long gsrc[X3 ? 1 : NPASS];
            uint32_t gsz[X3 ? 1 : NPASS];
            if (!X3) {
#pragma unroll
                for (int i = 0; i < NPASS; ++i) {
                    gsrc[X3 ? 0 : i] = reinterpret_cast<unsigned long long>(a.fT + xoff[i]);
                    gsz[X3 ? 0 : i] = ((livemask >> i) & 1u) ? 16u : 0u;
                }
            }
            if (item + (int)gridDim.x < nitems) load_idx(item + (int)gridDim.x, nxt);   // in flight during this item's chunks
            if (a.dbg) t_set += clock64() - ts0;
            for (int kc = 0; kc < nchunks; ++kc) {
                const long long tw0 = a.dbg ? clock64() : 0;
                if (pass > 0) mbar_wait(fs_smem(&empty[st]), (uint32_t)((pass - 1) & 1));
                if (a.dbg) t_we += clock64() - tw0;
                unsigned char *xt = ring + (size_t)st * slot_bytes;
                const uint32_t xs = fs_smem(xt);
                const int k0 = kc * 32 + slot * 4;
                if (!X3) {
                    // TF32: predicated 16-byte cp.async per row straight into the swizzled tile (zero-fill past the
                    // features and for dead rows); the (dp, 0) slot of the last chunk belongs to the dp warp
                    const bool featk = k0 + 4 <= a.C;
                    if (kc == kc_dp && slot == dp_slot) {
                        const int sb = itl & 1;
                        mbar_wait(fs_smem(&dp_ready[sb]), (uint32_t)((itl >> 1) & 1));
#pragma unroll
                        for (int i = 0; i < NPASS; ++i)
                            *reinterpret_cast<float4 *>(xt + soff0 + (uint32_t)i * (FS_RPP * FS_ROWB)) = dp_stage[sb * NT + i * FS_RPP + rsub];
                        mbar_arrive(fs_smem(&dp_free[sb]));
                    } else if (featk) {
#pragma unroll
                        for (int i = 0; i < NPASS; ++i) {
                            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(xs + soff0 + (uint32_t)i * (FS_RPP * FS_ROWB)),
                                         "l"(gsrc[X3 ? 0 : i]), "r"(gsz[X3 ? 0 : i])
                                         : "memory");
                            gsrc[X3 ? 0 : i] += FS_ROWB;
                        }
                    } else {                                       // past the features: zeros (the source is not read)
#pragma unroll
                        for (int i = 0; i < NPASS; ++i) cp_async16(xs + soff0 + (uint32_t)i * (FS_RPP * FS_ROWB), a.fT, false);
                    }
                } else {
                    // 3xTF32: through registers (split into hi = rna(x) and lo = x - hi), relative coordinates inline
#pragma unroll
                    for (int i = 0; i < NPASS; ++i) {
                        const int r = i * FS_RPP + rsub;
                        const bool live = (livemask >> i) & 1u;
                        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (live && k0 + 4 <= a.C) {
                            v = __ldg(reinterpret_cast<const float4 *>(a.fT + xoff[i] + kc * 32));
                        } else if (live && k0 == a.C) {
                            const float *pp = a.xyz + (size_t)rowv[X3 ? i : 0] * 3;
                            const float *qq = a.qxyz + (size_t)(q0 + r / NS) * 3;
                            v.x = (__ldg(pp) - __ldg(qq)) * a.inv_radius;
                            v.y = (__ldg(pp + 1) - __ldg(qq + 1)) * a.inv_radius;
                            v.z = (__ldg(pp + 2) - __ldg(qq + 2)) * a.inv_radius;
                        }
                        fs_store<X3>(xt, xt + NT * FS_ROWB, sw128_off(r, slot), v);
                    }
                }
                if (!wres && tid == wtid) {
                    asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(fs_smem(&full[st])), "r"(WT_BYTES) : "memory");
                    bulk_w(xs + XT_BYTES, sl, kc, fs_smem(&full[st]));
                }
                if (!X3) {
                    // the arrival fires when this thread's copies have landed; the 32 threads that wrote the (dp, 0) column
                    // with plain stores have none in flight and arrive in the ordinary (release) way
                    if (kc == kc_dp && slot == dp_slot) mbar_arrive(fs_smem(&full[st]));
                    else asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(fs_smem(&full[st])) : "memory");
                } else {
                    fence_async_smem();                            // generic-proxy stores -> visible to the MMA (async proxy)
                    mbar_arrive(fs_smem(&full[st]));
                }
                if (++st == S) { st = 0; ++pass; }
            }
        }
        if (a.dbg && tid == 0) {
            a.dbg[blockIdx.x * 12 + 0] = clock64() - t_begin;
            a.dbg[blockIdx.x * 12 + 1] = t_we;
            a.dbg[blockIdx.x * 12 + 2] = t_set;
        }
    } else if (warp == FS_WARP_MMA) {
        // ================================================================== MMA issuer (one elected lane)
        const uint32_t idesc = umma_idesc_tf32(FS_MT, NT);
        if (wres) {
            mbar_wait(fs_smem(w_full), 0);
            tc_fence_after();
        }
        int it = 0, st = 0;
        uint32_t ph = 0;                                           // ring slot and its phase parity, kept incrementally
        long long t_wf = 0, t_wa = 0, t_iss = 0;
        const long long t_begin = a.dbg ? clock64() : 0;
        const uint32_t ring_s = fs_smem(ring), wres_s = fs_smem(base);
        const uint64_t desc_hi = umma_desc_sw128(0);               // everything but the 14-bit start address
        for (int item = blockIdx.x; item < nitems; item += (int)gridDim.x, ++it) {
            const int buf = it & 1;
            const long long ta0 = a.dbg ? clock64() : 0;
            mbar_wait(fs_smem(&acc_empty[buf]), (uint32_t)(((it >> 1) & 1) ^ 1));   // the epilogue has drained this accumulator
            if (a.dbg) t_wa += clock64() - ta0;
            const uint32_t d = tmem_base + (uint32_t)(buf * NT);
            for (int kc = 0; kc < nchunks; ++kc) {
                const long long tf0 = a.dbg ? clock64() : 0;
                mbar_wait(fs_smem(&full[st]), ph);
                if (a.dbg) t_wf += clock64() - tf0;
                const long long ti0 = a.dbg ? clock64() : 0;
                if (lane == 0) {
                    fence_async_smem();                            // the chunk was written through the generic proxy
                    tc_fence_after();
                    const uint32_t xs = ring_s + (uint32_t)st * slot_bytes;
                    const uint32_t ws = wres ? wres_s + (uint32_t)kc * WT_BYTES : xs + XT_BYTES;
                    // descriptors differ in the start-address field only: (addr >> 4), +2 per k-step of 32 bytes
                    const uint64_t db0 = desc_hi | (uint64_t)((xs & 0x3FFFF) >> 4), da0 = desc_hi | (uint64_t)((ws & 0x3FFFF) >> 4);
                    const int ksteps = min(4, (a.Kp - kc * 32) >> 3);
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        if (ks < ksteps) {
                            const uint32_t acc = (kc | ks) != 0 ? 1u : 0u;
                            const uint64_t da = da0 + (uint64_t)(2 * ks), db = db0 + (uint64_t)(2 * ks);
                            if (X3) {
                                constexpr uint64_t XLO = (uint64_t)((NT * FS_ROWB) >> 4), WLO = (uint64_t)((FS_MT * FS_ROWB) >> 4);
                                umma_tf32(d, da + WLO, db, idesc, acc);                     // lo * hi
                                umma_tf32(d, da, db + XLO, idesc, 1u);                      // hi * lo
                                umma_tf32(d, da, db, idesc, 1u);                            // hi * hi
                            } else {
                                umma_tf32(d, da, db, idesc, acc);
                            }
                        }
                    }
                    umma_commit(fs_smem(&empty[st]));               // frees the ring slot once these MMAs have read it
                    if (kc == nchunks - 1) umma_commit(fs_smem(&acc_full[buf]));
                }
                __syncwarp();
                if (a.dbg) t_iss += clock64() - ti0;
                if (++st == S) { st = 0; ph ^= 1u; }
            }
        }
        if (a.dbg && lane == 0) {
            a.dbg[blockIdx.x * 12 + 4] = clock64() - t_begin;
            a.dbg[blockIdx.x * 12 + 5] = t_wf;
            a.dbg[blockIdx.x * 12 + 6] = t_wa;
            a.dbg[blockIdx.x * 12 + 7] = t_iss;
        }
        tc_fence_before();
    } else if (warp == FS_WARP_DP) {
        // ================================================================== relative coordinates (TF32 path)
        // idx -> xyz is a dependent chain of global loads.  This warp walks it for the item's NT positions — the indices
        // were fetched one item earlier — and leaves (dp, 0) in one of two staging buffers, up to two items ahead of the
        // ring; the 32 producer threads that own that 16-byte column move it into the item's last chunk.
        // (dp_ready / dp_free are an ordinary double-buffer handshake: each side is at most one phase from the other.)
        if (!X3) {
            long long t_wd = 0;
            int nidx[NPASS];
            auto load_idx = [&](int item) {
                const int q0 = (int)fdiv((uint32_t)item, a.div_s) * QPT;
#pragma unroll
                for (int j = 0; j < NPASS; ++j) {
                    const int r = j * 32 + lane;
                    const int qg = q0 + r / NS;
                    nidx[j] = qg < Q ? __ldg(a.idx + (long long)qg * NS + (r % NS)) : -1;
                }
            };
            if ((int)blockIdx.x < nitems) load_idx((int)blockIdx.x);
            int itl = 0;
            for (int item = blockIdx.x; item < nitems; item += (int)gridDim.x, ++itl) {
                const int q0 = (int)fdiv((uint32_t)item, a.div_s) * QPT;
                float4 dp[NPASS];
#pragma unroll
                for (int j = 0; j < NPASS; ++j) {
                    const int qg = q0 + (j * 32 + lane) / NS;
                    dp[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (nidx[j] >= 0) {
                        const uint32_t bb = fdiv((uint32_t)qg, a.div_m);
                        const float *pp = a.xyz + ((size_t)bb * a.N + nidx[j]) * 3;
                        const float *qq = a.qxyz + (size_t)qg * 3;
                        dp[j].x = (__ldg(pp) - __ldg(qq)) * a.inv_radius;
                        dp[j].y = (__ldg(pp + 1) - __ldg(qq + 1)) * a.inv_radius;
                        dp[j].z = (__ldg(pp + 2) - __ldg(qq + 2)) * a.inv_radius;
                    }
                }
                if (item + (int)gridDim.x < nitems) load_idx(item + (int)gridDim.x);
                const int sb = itl & 1;
                const long long td0 = a.dbg ? clock64() : 0;
                if (itl >= 2) mbar_wait(fs_smem(&dp_free[sb]), (uint32_t)(((itl >> 1) - 1) & 1));   // item itl - 2 has been moved out
                if (a.dbg) t_wd += clock64() - td0;
#pragma unroll
                for (int j = 0; j < NPASS; ++j) dp_stage[sb * NT + j * 32 + lane] = dp[j];
                mbar_arrive(fs_smem(&dp_ready[sb]));
            }
            if (a.dbg && lane == 0) a.dbg[blockIdx.x * 12 + 3] = t_wd;
        }
    } else {
        // ================================================================== epilogue: a thread owns one channel lane and
        // every other query of the tile (two threads per lane)
        const int quad = warp & 3;                                 // TMEM lane quadrant this warp may read
        const int half = (warp - FS_WARP_EPI) >> 2;                // which queries of the tile: qi % 2 == half
        const int lane_o = quad * 32 + lane;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(quad * 32) << 16);
        float csum = 0.f, csq = 0.f;                               // BatchNorm sums of channel `co`, flushed when it changes
        int co = -1;
        auto flush = [&]() {
            if (co >= 0 && co < a.O) {
                const long long rep = (long long)((blockIdx.x * 2 + half) % FS_REPL) * 2 * a.O;
                atomicAdd(a.gsum + rep + co, (double)csum);
                atomicAdd(a.gsumsq + rep + co, (double)csq);
            }
            csum = 0.f;
            csq = 0.f;
        };
        int it = 0;
        long long t_wacc = 0;
        const long long t_begin = a.dbg ? clock64() : 0;
        for (int item = blockIdx.x; item < nitems; item += (int)gridDim.x, ++it) {
            const int buf = it & 1;
            const uint32_t tile = fdiv((uint32_t)item, a.div_s);
            const int q0 = (int)tile * QPT;
            const int o = (item - (int)tile * nsl) * FS_MT + lane_o;
            if (o != co) { flush(); co = o; }
            const bool live = o < a.O;
            // gamma < 0: BatchNorm + ReLU decrease in y, the pooled maximum sits at the MINIMUM of y = the maximum of -y
            const float sgn = (live && __ldg(a.gamma + o) < 0.f) ? -1.f : 1.f;
            const long long te0 = a.dbg ? clock64() : 0;
            mbar_wait(fs_smem(&acc_full[buf]), (uint32_t)((it >> 1) & 1));
            if (a.dbg) t_wacc += clock64() - te0;
            tc_fence_after();
#pragma unroll 1
            for (int qi = half; qi < QPT; qi += 2) {
                float v[NS];
                tmem_ld_query<NS>(lane_addr + (uint32_t)(buf * NT + qi * NS), v);   // warp-uniform control flow up to here
                // v = sign(gamma) y (the weight rows carry the sign).  Two independent (max, arg-max) chains over the
                // halves, four partial sums: no branches, short chains
                float m0 = v[0], m1 = v[NS / 2];
                int i0 = 0, i1 = NS / 2;
                float s1a = 0.f, s1b = 0.f, s2a = 0.f, s2b = 0.f;
#pragma unroll
                for (int s = 0; s < NS / 2; ++s) {
                    const float x = v[s], y = v[s + NS / 2];
                    s1a += x;
                    s1b += y;
                    s2a = fmaf(x, x, s2a);
                    s2b = fmaf(y, y, s2b);
                    const bool gx = x > m0, gy = y > m1;
                    m0 = gx ? x : m0;
                    i0 = gx ? s : i0;
                    m1 = gy ? y : m1;
                    i1 = gy ? s + NS / 2 : i1;
                }
                const bool g2 = m1 > m0;                           // ties keep the lower sample index
                const float best = (g2 ? m1 : m0) * sgn;
                const int bi = g2 ? i1 : i0;
                const long long qg = q0 + qi;
                if (live && qg < (long long)Q) {
                    csum += (s1a + s1b) * sgn;
                    csq += s2a + s2b;
                    a.ysel[qg * a.O + o] = best;
                    a.arg[qg * a.O + o] = (unsigned char)bi;
                }
            }
            tc_fence_before();
            mbar_arrive(fs_smem(&acc_empty[buf]));                 // the MMA warp may overwrite this accumulator
        }
        flush();
        if (a.dbg && tid == FS_WARP_EPI * 32) {
            a.dbg[blockIdx.x * 12 + 8] = clock64() - t_begin;
            a.dbg[blockIdx.x * 12 + 9] = t_wacc;
        }
    }
    __syncthreads();
    if (warp == FS_WARP_MMA) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)(2 * NT)) : "memory");
    }
}

// BatchNorm statistics from the sums: mean, biased variance, 1/sqrt(var + eps); one warp per channel over the replicas
__global__ void __launch_bounds__(128)
fused_sa_stats_kernel(int O, double count, float eps, const double *__restrict__ gsum,
                      const double *__restrict__ gsumsq, float *__restrict__ mean,
                      float *__restrict__ var, float *__restrict__ invstd) {
    const int o = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (o >= O) return;
    double s1 = 0.0, s2 = 0.0;
    for (int r = lane; r < FS_REPL; r += 32) {
        s1 += gsum[(long long)r * 2 * O + o];
        s2 += gsumsq[(long long)r * 2 * O + o];
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, d);
        s2 += __shfl_xor_sync(0xffffffffu, s2, d);
    }
    if (lane == 0) {
        const double m = s1 / count;
        double v = s2 / count - m * m;
        if (v < 0.0) v = 0.0;
        mean[o] = (float)m;
        var[o] = (float)v;
        invstd[o] = (float)(1.0 / sqrt(v + (double)eps));
    }
}

// out[b, o, m] = relu((ysel[b*M + m, o] - mean[o]) * invstd[o] * gamma[o] + beta[o]); 32 x 32 transposing tiles
__global__ void __launch_bounds__(256)
fused_sa_finalize_kernel(int B, int M, int O, const float *__restrict__ ysel, const float *__restrict__ mean,
                         const float *__restrict__ invstd, const float *__restrict__ gamma,
                         const float *__restrict__ beta, float *__restrict__ out) {
    __shared__ float t[32][33];
    const int b = blockIdx.z, m0 = blockIdx.x * 32, o0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty + 8 * i, o = o0 + tx;
        float v = 0.f;
        if (m < M && o < O) {
            const float y = __ldg(ysel + ((long long)b * M + m) * O + o);
            v = fmaxf((y - __ldg(mean + o)) * __ldg(invstd + o) * __ldg(gamma + o) + __ldg(beta + o), 0.f);
        }
        t[ty + 8 * i][tx] = v;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int o = o0 + ty + 8 * i, m = m0 + tx;
        if (m < M && o < O) out[((long long)b * O + o) * M + m] = t[tx][ty + 8 * i];
    }
}

template <int NS, bool X3>
static int launch_fused_fwd(FusedFwdArgs &a, cudaStream_t st) {
    static const int env_st = getenv("AMC3D_FUSED_STAGES") ? atoi(getenv("AMC3D_FUSED_STAGES")) : 0;
    static const int env_wres = getenv("AMC3D_FUSED_WRES") ? atoi(getenv("AMC3D_FUSED_WRES")) : 1;
    static const int env_ctas = getenv("AMC3D_FUSED_CTAS") ? atoi(getenv("AMC3D_FUSED_CTAS")) : 0;
    static const int env_kb = getenv("AMC3D_FUSED_KB") ? atoi(getenv("AMC3D_FUSED_KB")) : 212;
    static const bool env_dbg = getenv("AMC3D_FUSED_DBG") != nullptr;
    constexpr int NT = X3 ? 128 : 256;                              // as in the kernel
    const size_t xt = (size_t)(X3 ? 2 : 1) * NT * FS_ROWB, wt = (size_t)(X3 ? 2 : 1) * FS_MT * FS_ROWB;
    const size_t budget = (size_t)env_kb * 1024;                    // of the 227 KB a CTA may own: one persistent CTA per SM
    const int nchunks = (a.Kp + 31) / 32;
    a.nslices = div_up(a.O, FS_MT);
    a.nitems = (int)(div_up_ll((long long)a.B * a.M, NT / NS) * a.nslices);
    a.div_m = make_fastdiv((uint32_t)a.M);
    a.div_s = make_fastdiv((uint32_t)a.nslices);
    // a single output-channel slice whose weights leave room for >= 3 ring slots: keep them resident
    a.w_resident = (env_wres && a.nslices == 1 && nchunks * wt + 3 * xt <= budget) ? 1 : 0;
    const size_t fixed = a.w_resident ? nchunks * wt : 0, slot = a.w_resident ? xt : xt + wt;
    a.stages = (int)min((size_t)FS_MAX_STAGES, (budget - fixed) / slot);
    if (env_st >= 2 && env_st <= a.stages) a.stages = env_st;
    if (a.stages < 2) return (int)cudaErrorInvalidConfiguration;
    const size_t smem = fixed + a.stages * slot + 1024 + (X3 ? 0 : 2 * NT * sizeof(float4));   // + alignment slack + dp staging
    cudaError_t e = cudaFuncSetAttribute(fused_sa_fwd_kernel<NS, X3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    fused_sa_wprep_kernel<X3><<<a.nslices * nchunks, 256, 0, st>>>(a.O, a.Kp, nchunks, a.Wp, a.gamma, const_cast<float *>(a.Wsw));
    static long long *dbg_dev = nullptr;
    a.dbg = nullptr;
    if (env_dbg) {
        if (!dbg_dev) cudaMalloc(&dbg_dev, sizeof(long long) * 12 * 1024);
        cudaMemsetAsync(dbg_dev, 0, sizeof(long long) * 12 * 1024, st);
        a.dbg = dbg_dev;
    }
    const unsigned grid = (unsigned)min(a.nitems, env_ctas > 0 ? env_ctas : kNumSMs);
    fused_sa_fwd_kernel<NS, X3><<<grid, FS_THREADS, smem, st>>>(a);
    if (env_dbg) {                                                  // development aid: cycles per role, averaged over the CTAs
        static long long h[12 * 1024];
        cudaStreamSynchronize(st);
        cudaMemcpy(h, dbg_dev, sizeof(long long) * 12 * grid, cudaMemcpyDeviceToHost);
        double avg[12] = {0};
        for (unsigned b = 0; b < grid; ++b)
            for (int k = 0; k < 12; ++k) avg[k] += (double)h[b * 12 + k] / grid;
        fprintf(stderr, "[fused dbg] C=%d O=%d NT=%d items/CTA=%.1f chunks=%d S=%d wres=%d | producer total %.0f wait_empty %.0f setup %.0f | dp warp wait_free %.0f | "
                        "mma total %.0f wait_full %.0f wait_acc_empty %.0f issue %.0f | epilogue total %.0f wait_acc_full %.0f\n",
                a.C, a.O, NT, (double)a.nitems / grid, nchunks, a.stages, a.w_resident, avg[0], avg[1], avg[2], avg[3], avg[4], avg[5], avg[6],
                avg[7], avg[8], avg[9]);
    }
    return 0;
}

}  // namespace amc3d

using namespace amc3d;

extern "C" int amc3d_fused_sa_forward(int b, int n, int m, int c, int o, int nsample, float radius, int normalize_dp,
                                      int precision, float eps, const float *featT, const float *xyz,
                                      const float *new_xyz, const int *idx, const float *w_packed, float *w_tiles,
                                      const float *gamma, const float *beta, float *ysel, unsigned char *arg,
                                      double *sums, float *mean, float *var, float *invstd, float *out,
                                      void *stream) {
    AMC3D_REQUIRE(b >= 0 && n >= 1 && m >= 0 && c >= 8 && o >= 1, AMC3D_EINVAL, "fused_sa_forward: bad sizes b=%d n=%d m=%d c=%d o=%d", b, n, m, c, o);
    AMC3D_REQUIRE(c % 8 == 0, AMC3D_ELIMIT, "fused_sa_forward: C=%d is not a multiple of 8", c);
    AMC3D_REQUIRE(nsample == 16 || nsample == 32, AMC3D_ELIMIT, "fused_sa_forward: nsample=%d (16 and 32 are built)", nsample);
    AMC3D_REQUIRE(precision == 1 || precision == 3, AMC3D_EINVAL, "fused_sa_forward: precision=%d is not 1 (TF32) / 3 (3xTF32)", precision);
    AMC3D_REQUIRE(b <= 65535 && (long long)b * m < (1ll << 31) / 64, AMC3D_ELIMIT, "fused_sa_forward: batch %d x %d queries too large", b, m);
    AMC3D_REQUIRE((long long)b * n * c < (1ll << 31), AMC3D_ELIMIT, "fused_sa_forward: %d x %d x %d features exceed 2^31 elements", b, n, c);
    if (b == 0 || m == 0) return 0;
    cudaStream_t st = as_stream(stream);
    FusedFwdArgs a;
    a.fT = featT; a.xyz = xyz; a.qxyz = new_xyz; a.idx = idx; a.Wp = w_packed; a.Wsw = w_tiles; a.gamma = gamma;
    a.ysel = ysel; a.arg = arg; a.gsum = sums; a.gsumsq = sums + o;
    a.B = b; a.N = n; a.M = m; a.C = c; a.O = o; a.Kp = c + 8;
    a.inv_radius = normalize_dp ? 1.0f / radius : 1.0f;
    cudaMemsetAsync(sums, 0, sizeof(double) * 2 * (size_t)o * FS_REPL, st);
    int rc;
    if (nsample == 32) rc = precision == 3 ? launch_fused_fwd<32, true>(a, st) : launch_fused_fwd<32, false>(a, st);
    else rc = precision == 3 ? launch_fused_fwd<16, true>(a, st) : launch_fused_fwd<16, false>(a, st);
    if (rc != 0) {
        set_error("fused_sa_forward: %s", cudaGetErrorString((cudaError_t)rc));
        return rc;
    }
    const double count = (double)b * m * nsample;
    fused_sa_stats_kernel<<<div_up(o, 4), 128, 0, st>>>(o, count, eps, sums, sums + o, mean, var, invstd);
    dim3 fgrid(div_up(m, 32), div_up(o, 32), b);
    fused_sa_finalize_kernel<<<fgrid, 256, 0, st>>>(b, m, o, ysel, mean, invstd, gamma, beta, out);
    return check_launch("fused_sa_forward");
}

namespace amc3d {

// =================================================================================================================
// Backward.  With  yhat = (y - mean) * invstd,  z = gamma * yhat + beta,  out = max_s relu(z)  and  G = dL/dout:
//   D[p,o]  = G[q,o] * [out > 0] * [s(p) == arg(q,o)]                (one non-zero per (query, channel))
//   dbeta   = sum_p D            dgamma = sum_p D * yhat
//   dL/dy[p,o] = ghat_o * D[p,o] - c0_o - c1_o * y[p,o]     ghat = gamma*invstd, c1 = ghat*invstd*dgamma/P,
//                                                           c0 = ghat*dbeta/P - c1*mean
// The two terms without D are dense over all P positions but linear in x, so they reduce to the first and second
// moments of the grouped input, which reduce to per-support-point counts because x[p] = [f[idx[p]] | dp[p]]:
// sum_p f[idx[p]] f[idx[p]]^T = sum_n cnt[n] f[n] f[n]^T  (fused_sa_moments_kernel + (B*N) x C x C GEMMs).
// The D term is sparse, and its target only depends on WHICH support point the arg-max position refers to:
//   A[n, o] = sum_{q : idx[q, arg(q,o)] = n} ghat_o * G'[q,o]                      (fused_sa_bwd_scatter_kernel)
//   dW_f = A^T f        df = A W_f        dW_dp[o,:] = sum_q ghat_o G'[q,o] dp[(q, arg(q,o)), :]
// i.e. one scatter of Q*O scalars followed by two (B*N) x O x C GEMMs — 1/nsample of the convolution's work —
// instead of the two convolution-sized products a dense backward would need.  The GEMMs are plain library calls
// on the host side (layers/_fused_backward.py).  Nothing of size P x (3+C) or P x O is ever stored.
// =================================================================================================================

// For a 32 (queries) x 32 (channels) tile: G' = G * [out > 0]; dbeta, dgamma (f64 atomics); the scatter into
// A (B*N, O) (lanes = consecutive channels: 128 contiguous bytes per warp and query); dW of the three
// relative-coordinate columns, wdp (O, 3) f64.
__global__ void __launch_bounds__(256)
fused_sa_bwd_scatter_kernel(int B, int N, int M, int O, int NS, float inv_radius, const float *__restrict__ gout,
                            const float *__restrict__ out, const float *__restrict__ ysel,
                            const unsigned char *__restrict__ arg, const int *__restrict__ idx,
                            const float *__restrict__ xyz, const float *__restrict__ qxyz,
                            const float *__restrict__ mean, const float *__restrict__ invstd,
                            const float *__restrict__ gamma, float *__restrict__ A, int lda, double *__restrict__ dbeta,
                            double *__restrict__ dgamma, double *__restrict__ wdp) {
    __shared__ float t[32][33];
    __shared__ float red[5][8][32];
    const int b = blockIdx.z, m0 = blockIdx.x * 32, o0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int o = o0 + ty + 8 * i, m = m0 + tx;
        float v = 0.f;
        if (m < M && o < O) {
            const long long e = ((long long)b * O + o) * M + m;
            v = __ldg(out + e) > 0.f ? __ldg(gout + e) : 0.f;
        }
        t[ty + 8 * i][tx] = v;                      // [o][m]
    }
    __syncthreads();
    const int o = o0 + tx;
    float sb = 0.f, sg = 0.f, sd[3] = {0.f, 0.f, 0.f};
    if (o < O) {
        const float mu = __ldg(mean + o), is = __ldg(invstd + o), gh = __ldg(gamma + o) * is;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int m = m0 + ty + 8 * i;
            if (m < M) {
                const long long q = (long long)b * M + m;
                const float g = t[tx][ty + 8 * i];
                if (g != 0.f) {
                    const float yh = (__ldg(ysel + q * O + o) - mu) * is;
                    sb += g;
                    sg += g * yh;
                    const int s = __ldg(arg + q * O + o);
                    const long long row = (long long)b * N + __ldg(idx + q * NS + s);
                    const float gg = g * gh;
                    atomicAdd(A + row * lda + o, gg);
#pragma unroll
                    for (int j = 0; j < 3; ++j)
                        sd[j] += gg * ((__ldg(xyz + row * 3 + j) - __ldg(qxyz + q * 3 + j)) * inv_radius);
                }
            }
        }
    }
    red[0][ty][tx] = sb; red[1][ty][tx] = sg; red[2][ty][tx] = sd[0]; red[3][ty][tx] = sd[1]; red[4][ty][tx] = sd[2];
    __syncthreads();
    if (ty < 5 && o < O) {
        float a = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) a += red[ty][i][tx];
        if (a != 0.f) {
            if (ty == 0) atomicAdd(dbeta + o, (double)a);
            else if (ty == 1) atomicAdd(dgamma + o, (double)a);
            else atomicAdd(wdp + o * 3 + (ty - 2), (double)a);
        }
    }
}

// per support point: cnt[b,n] = #grouped positions that reference it, dpsum[b,n,:] = sum of their relative
// coordinates; mom[0..2] = sum_p dp, mom[3..11] = sum_p dp dp^T  (f64)
__global__ void __launch_bounds__(256)
fused_sa_moments_kernel(int B, int N, int M, int NS, float inv_radius, const float *__restrict__ xyz,
                        const float *__restrict__ qxyz, const int *__restrict__ idx, float *__restrict__ cnt, int ldc,
                        float *__restrict__ dpsum, int ldd, double *__restrict__ mom, int packed) {
    const long long P = (long long)B * M * NS;
    const long long p = (long long)blockIdx.x * 256 + threadIdx.x;
    float d[3] = {0.f, 0.f, 0.f};
    if (p < P) {
        const long long q = p / NS;
        const long long b = q / M;
        const long long row = b * N + __ldg(idx + p);
#pragma unroll
        for (int j = 0; j < 3; ++j) d[j] = (__ldg(xyz + row * 3 + j) - __ldg(qxyz + q * 3 + j)) * inv_radius;
        if (packed) {                                  // [dpsum | cnt] are four consecutive, 16-byte aligned floats: one reduction
            asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(dpsum + row * ldd), "f"(d[0]), "f"(d[1]), "f"(d[2]),
                         "f"(1.f)
                         : "memory");
        } else {
            atomicAdd(cnt + row * ldc, 1.f);
#pragma unroll
            for (int j = 0; j < 3; ++j) atomicAdd(dpsum + row * ldd + j, d[j]);
        }
    }
    float v[12] = {d[0], d[1], d[2], d[0] * d[0], d[0] * d[1], d[0] * d[2], d[1] * d[0], d[1] * d[1], d[1] * d[2],
                   d[2] * d[0], d[2] * d[1], d[2] * d[2]};
    __shared__ float red[8][12];
#pragma unroll
    for (int j = 0; j < 12; ++j) {
        const float s = warp_sum(v[j]);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][j] = s;
    }
    __syncthreads();
    if (threadIdx.x < 12) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += red[w][threadIdx.x];
        atomicAdd(mom + threadIdx.x, (double)s);
    }
}

// O-sized coefficient vectors of the dense BatchNorm terms, in FP64 from the FP64 sums:
//   ghat = gamma invstd,  c1 = ghat invstd dgamma / P,  c0 = ghat dbeta / P - c1 mean
// and the row-scaled weights: c1wx (O, Kp) = [ c1 (.) W' | c0 | 0 ]  (one GEMM with W'^T then gives both Q = W'^T diag(c1) W'
// and v = W'^T c0; Kp = C + 8 columns like W' itself, so that every GEMM dimension is a multiple of 8 and the library
// stays on its tensor-core kernels).  Also the FP32 copies of dgamma / dbeta the caller returns.
__global__ void __launch_bounds__(128)
fused_sa_bwd_coef_kernel(int O, int Kq, int Kp, double P, const float *__restrict__ gamma, const float *__restrict__ invstd,
                         const float *__restrict__ mean, const double *__restrict__ red, const float *__restrict__ Wp,
                         float *__restrict__ c1wx, float *__restrict__ dgamma_f, float *__restrict__ dbeta_f) {
    const int o = blockIdx.x;
    const double is = (double)__ldg(invstd + o), ghat = (double)__ldg(gamma + o) * is;
    const double c1d = ghat * is * red[O + o] / P;
    const double c0d = ghat * red[o] / P - c1d * (double)__ldg(mean + o);
    const float c1f = (float)c1d;
    float *row = c1wx + (size_t)o * Kp;
    for (int k = threadIdx.x; k < Kp; k += 128) row[k] = k < Kq ? c1f * __ldg(Wp + (size_t)o * Kp + k) : 0.f;
    __syncthreads();
    if (threadIdx.x == 0) {
        row[Kq] = (float)c0d;
        dgamma_f[o] = (float)red[O + o];
        dbeta_f[o] = (float)red[o];
    }
}

// the small matrices of the backward in one pass (Kq = C + 3, Kp = C + 8, Kc = O + C + 4; g1 = f^T X is (C, Kc),
// qv = W'^T c1wx is (Kp, Kp) with [Q | v] in its first Kq rows and Kq + 1 columns; outputs padded to Kp with zeros):
//   sxx (Kp, Kp) = [[S_ff, S_fd], [S_fd^T, S_dd]]                       second moments of the grouped input
//   wc  (Kc, C)  = [ W_f ; -Q_ff^T ; -Q_fd^T ; -v_f^T ]                  df = X wc
//   dwp (O, Kp)  = [ (A^T f) | wdp ] - c0 (x) S_x                        the arg-max term and the c0 term of dW'
__global__ void __launch_bounds__(256)
fused_sa_bwd_assemble_kernel(int O, int C, const float *__restrict__ g1, const double *__restrict__ mom,
                             const double *__restrict__ wdp, const float *__restrict__ Wp, const float *__restrict__ qv,
                             const float *__restrict__ c1wx, float *__restrict__ sxx, float *__restrict__ wc,
                             float *__restrict__ dwp) {
    const int Kq = C + 3, Kp = C + 8, Kc = O + C + 4;
    const long long n1 = (long long)Kp * Kp, n2 = (long long)Kc * C, n3 = (long long)O * Kp;
    for (long long e = (long long)blockIdx.x * 256 + threadIdx.x; e < n1 + n2 + n3; e += (long long)gridDim.x * 256) {
        if (e < n1) {
            const int i = (int)(e / Kp), j = (int)(e - (long long)i * Kp);
            float x = 0.f;
            if (i < Kq && j < Kq) {
                if (i < C) x = g1[(size_t)i * Kc + O + j];                        // S_ff | S_fd (columns O + C + (j - C))
                else if (j < C) x = g1[(size_t)j * Kc + O + i];                   // S_fd^T
                else x = (float)mom[3 + (i - C) * 3 + (j - C)];
            }
            sxx[e] = x;
        } else if (e < n1 + n2) {
            const long long t = e - n1;
            const int r = (int)(t / C), c = (int)(t - (long long)r * C);
            float x;
            if (r < O) x = __ldg(Wp + (size_t)r * Kp + c);
            else x = -qv[(size_t)c * Kp + (r - O)];                               // -Q[c, j] for j < Kq, then -v[c] (column Kq)
            wc[t] = x;
        } else {
            const long long t = e - n1 - n2;
            const int o = (int)(t / Kp), k = (int)(t - (long long)o * Kp);
            float x = 0.f;
            if (k < Kq) {
                const float base = k < C ? g1[(size_t)k * Kc + o] : (float)wdp[o * 3 + (k - C)];
                const float sx = k < C ? g1[(size_t)k * Kc + O + C + 3] : (float)mom[k - C];
                x = base - c1wx[(size_t)o * Kp + Kq] * sx;
            }
            dwp[t] = x;
        }
    }
}

// zero `ncols` floats at the start of each of `rows` rows that are `ld` floats apart (a 2-D memset of narrow rows
// is far slower than this)
__global__ void __launch_bounds__(256)
fused_sa_zero_cols_kernel(float *__restrict__ base, long long rows, int ld, int ncols) {
    const long long total = rows * ncols;
    for (long long e = (long long)blockIdx.x * 256 + threadIdx.x; e < total; e += (long long)gridDim.x * 256) {
        const long long r = e / ncols;
        base[r * ld + (e - r * ncols)] = 0.f;
    }
}
__global__ void __launch_bounds__(256)
fused_sa_zero_cols4_kernel(float4 *__restrict__ base, long long rows, int ld4, int ncols4) {
    const long long total = rows * ncols4;
    for (long long e = (long long)blockIdx.x * 256 + threadIdx.x; e < total; e += (long long)gridDim.x * 256) {
        const long long r = e / ncols4;
        base[r * ld4 + (e - r * ncols4)] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}
static void zero_cols(float *base, long long rows, int ld, int ncols, cudaStream_t st) {
    if (rows <= 0 || ncols <= 0) return;
    if (ld == ncols) {
        cudaMemsetAsync(base, 0, sizeof(float) * (size_t)rows * ncols, st);
    } else if (ld % 4 == 0 && ncols % 4 == 0 && (reinterpret_cast<uintptr_t>(base) & 15) == 0) {
        const long long total = rows * (ncols / 4);
        fused_sa_zero_cols4_kernel<<<(unsigned)min(div_up_ll(total, 256), (long long)kNumSMs * 16), 256, 0, st>>>(
            reinterpret_cast<float4 *>(base), rows, ld / 4, ncols / 4);
    } else {
        const long long total = rows * ncols;
        fused_sa_zero_cols_kernel<<<(unsigned)min(div_up_ll(total, 256), (long long)kNumSMs * 16), 256, 0, st>>>(base, rows, ld, ncols);
    }
}

}  // namespace amc3d

extern "C" int amc3d_fused_sa_backward_scatter(int b, int n, int m, int o, int nsample, float radius, int normalize_dp,
                                               const float *grad_out, const float *out, const float *ysel,
                                               const unsigned char *arg, const int *idx, const float *xyz,
                                               const float *new_xyz, const float *mean, const float *invstd,
                                               const float *gamma, float *a_scatter, int lda, double *dbeta_dgamma_wdp,
                                               void *stream) {
    AMC3D_REQUIRE(b >= 0 && n >= 1 && m >= 0 && o >= 1 && nsample >= 1, AMC3D_EINVAL, "fused_sa_backward_scatter: bad sizes");
    AMC3D_REQUIRE(b <= 65535, AMC3D_ELIMIT, "fused_sa_backward_scatter: batch %d > 65535", b);
    AMC3D_REQUIRE(lda >= o, AMC3D_EINVAL, "fused_sa_backward_scatter: row stride %d < O = %d", lda, o);
    cudaStream_t st = as_stream(stream);
    cudaMemsetAsync(dbeta_dgamma_wdp, 0, sizeof(double) * 5 * (size_t)o, st);
    zero_cols(a_scatter, (long long)b * n, lda, o, st);
    if (b == 0 || m == 0) return check_launch("fused_sa_backward_scatter");
    dim3 grid(div_up(m, 32), div_up(o, 32), b);
    fused_sa_bwd_scatter_kernel<<<grid, 256, 0, st>>>(b, n, m, o, nsample, normalize_dp ? 1.0f / radius : 1.0f, grad_out, out,
                                                      ysel, arg, idx, xyz, new_xyz, mean, invstd, gamma, a_scatter, lda,
                                                      dbeta_dgamma_wdp, dbeta_dgamma_wdp + o, dbeta_dgamma_wdp + 2 * o);
    return check_launch("fused_sa_backward_scatter");
}

extern "C" int amc3d_fused_sa_moments(int b, int n, int m, int nsample, float radius, int normalize_dp,
                                      const float *xyz, const float *new_xyz, const int *idx, float *cnt, int ld_cnt,
                                      float *dpsum, int ld_dps, double *mom, void *stream) {
    AMC3D_REQUIRE(b >= 0 && n >= 1 && m >= 0 && nsample >= 1, AMC3D_EINVAL, "fused_sa_moments: bad sizes");
    AMC3D_REQUIRE(ld_cnt >= 1 && ld_dps >= 3, AMC3D_EINVAL, "fused_sa_moments: row strides %d, %d", ld_cnt, ld_dps);
    cudaStream_t st = as_stream(stream);
    const bool side_by_side = dpsum + 3 == cnt && ld_cnt == ld_dps;
    const int packed = side_by_side && ld_dps % 4 == 0 && (reinterpret_cast<uintptr_t>(dpsum) & 15) == 0;
    if (side_by_side) {                                             // [dpsum | cnt] side by side: one pass
        zero_cols(dpsum, (long long)b * n, ld_dps, 4, st);
    } else {
        zero_cols(cnt, (long long)b * n, ld_cnt, 1, st);
        zero_cols(dpsum, (long long)b * n, ld_dps, 3, st);
    }
    cudaMemsetAsync(mom, 0, sizeof(double) * 12, st);
    const long long P = (long long)b * m * nsample;
    if (P > 0)
        fused_sa_moments_kernel<<<(unsigned)div_up_ll(P, 256), 256, 0, st>>>(b, n, m, nsample, normalize_dp ? 1.0f / radius : 1.0f,
                                                                              xyz, new_xyz, idx, cnt, ld_cnt, dpsum, ld_dps, mom, packed);
    return check_launch("fused_sa_moments");
}

extern "C" int amc3d_fused_sa_backward_coefs(int c, int o, double positions, const float *gamma, const float *invstd,
                                             const float *mean, const double *dbeta_dgamma_wdp, const float *w_packed,
                                             float *c1wx, float *dgamma, float *dbeta, void *stream) {
    AMC3D_REQUIRE(c >= 1 && o >= 1 && positions > 0, AMC3D_EINVAL, "fused_sa_backward_coefs: bad sizes");
    fused_sa_bwd_coef_kernel<<<o, 128, 0, as_stream(stream)>>>(o, c + 3, c + 8, positions, gamma, invstd, mean, dbeta_dgamma_wdp,
                                                               w_packed, c1wx, dgamma, dbeta);
    return check_launch("fused_sa_backward_coefs");
}

extern "C" int amc3d_fused_sa_backward_assemble(int c, int o, const float *g1, const double *mom,
                                                const double *dbeta_dgamma_wdp, const float *w_packed, const float *qv,
                                                const float *c1wx, float *sxx, float *wc, float *dwp, void *stream) {
    AMC3D_REQUIRE(c >= 1 && o >= 1, AMC3D_EINVAL, "fused_sa_backward_assemble: bad sizes");
    const long long total = (long long)(c + 8) * (c + 8) + (long long)(o + c + 4) * c + (long long)o * (c + 8);
    fused_sa_bwd_assemble_kernel<<<(unsigned)min(div_up_ll(total, 256), (long long)kNumSMs * 8), 256, 0, as_stream(stream)>>>(
        o, c, g1, mom, dbeta_dgamma_wdp + 2 * (size_t)o, w_packed, qv, c1wx, sxx, wc, dwp);
    return check_launch("fused_sa_backward_assemble");
}

// Host-side check of the division-free index arithmetic the fused forward uses on the device (FastDiv): returns
// n / d computed the way the kernel does (multiply-high by the precomputed magic, shift).  Valid for n < 2^31, d >= 1.
extern "C" unsigned int amc3d_debug_fastdiv(unsigned int n, unsigned int d) {
    const FastDiv f = make_fastdiv(d);
    if (f.one) return n;
    return (unsigned int)(((unsigned long long)n * f.m) >> 32) >> f.s;
}
