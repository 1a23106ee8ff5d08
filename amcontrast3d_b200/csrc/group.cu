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
#include <stdlib.h>

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
// TMA-staged variants: a CTA owns a tile of 32 channels x 256 positions held in shared memory.
//
// forward   lane = channel: every neighbour index is ONE coalesced 128-byte read of the (B,N,C)
//           workspace (L2-resident); the lane's 4 consecutive positions go to its tile row with a
//           conflict-free STS.128 (row stride 260 floats = 4 mod 32 banks); after a barrier the 32
//           rows leave as 32 bulk-async (TMA) stores of 1 KB with an L2 evict-first policy, so the
//           write stream neither occupies the LSU nor evicts the workspace from L2.
// backward  32 bulk-async loads bring the 32 x 1 KB grad rows in (mbarrier complete_tx); the
//           scatter-add is issued as red.global.add.v4.f32: a lane owns 4 consecutive channels of
//           one position, a warp instruction adds 4 positions x 128 contiguous bytes into the
//           (B,N,C) accumulator.  Tile rows are laid out at r*288 + 4*(r/4) floats so that the
//           position-major scalar LDS of that mapping is bank-conflict free.
// ---------------------------------------------------------------------------------------
constexpr int TP = 256;          // positions per tile
constexpr int TC = 32;           // channels per tile
constexpr int FWD_LD = TP + 4;   // forward tile row stride (floats)
constexpr int BWD_LD = TP + 32;  // backward tile row stride (floats); row r starts at r*BWD_LD + 4*(r>>2)

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t l2_evict_first_policy() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void bulk_store(void *gdst, uint32_t ssrc, uint32_t bytes, uint64_t pol) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(gdst),
                 "r"(ssrc), "r"(bytes), "l"(pol)
                 : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t sdst, const void *gsrc, uint32_t bytes, uint32_t bar, uint64_t pol) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(sdst),
        "l"(gsrc), "r"(bytes), "r"(bar), "l"(pol)
        : "memory");
}

template <int W>
__global__ void __launch_bounds__(W * 32)
group_fwd_tma_kernel(int C, int N, int P, int nsample, const float *__restrict__ srcT,
                     const int *__restrict__ idx, float *__restrict__ out) {
    constexpr int TPW = W * 32;                      // positions per tile
    constexpr int FWD_LD = TPW + 4;                  // row stride: 4 mod 32 banks -> conflict-free STS.128
    extern __shared__ __align__(128) float tile[];   // [TC][FWD_LD]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.z;
    const int c0 = blockIdx.y * TC;
    const int p0 = blockIdx.x * TPW;
    const int npos = min(TPW, P - p0);                // multiple of 4 (host checks P % 4 == 0)
    const int cc = min(c0 + lane, C - 1);
    const float *src = srcT + (long long)b * N * C + cc;
    const int *ip = idx + (long long)b * P + p0 + warp * 32;
    const int wpos = min(32, npos - warp * 32);       // positions of this warp (may be <= 0)
    float *trow = tile + lane * FWD_LD + warp * 32;
    if (wpos == 32) {
        // full warp tile: two batches of 16 independent gathers in flight per lane.  Slots that repeat
        // the first index of their query (ball_query's padding, ~40 % of all slots) reuse its value
        // instead of issuing another load (nsample % 32 == 0: the warp lies inside one query).
        int i_first = -1;
        float v_first = 0.f;
        if (nsample > 0) {
            const long long gp0 = (long long)p0 + warp * 32;
            i_first = __ldg(idx + (long long)b * P + gp0 / nsample * nsample);
            v_first = __ldg(src + (long long)i_first * C);
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            int4 ii[4];
            float4 v[4];
#pragma unroll
            for (int o = 0; o < 4; ++o) ii[o] = __ldg(reinterpret_cast<const int4 *>(ip + h * 16 + o * 4));
#pragma unroll
            for (int o = 0; o < 4; ++o) {
                v[o].x = ii[o].x == i_first ? v_first : __ldg(src + (long long)ii[o].x * C);
                v[o].y = ii[o].y == i_first ? v_first : __ldg(src + (long long)ii[o].y * C);
                v[o].z = ii[o].z == i_first ? v_first : __ldg(src + (long long)ii[o].z * C);
                v[o].w = ii[o].w == i_first ? v_first : __ldg(src + (long long)ii[o].w * C);
            }
#pragma unroll
            for (int o = 0; o < 4; ++o) *reinterpret_cast<float4 *>(trow + h * 16 + o * 4) = v[o];
        }
    } else {
        for (int o = 0; o * 4 < wpos; ++o) {
            const int4 ia = __ldg(reinterpret_cast<const int4 *>(ip + o * 4));
            float4 v;
            v.x = __ldg(src + (long long)ia.x * C);
            v.y = __ldg(src + (long long)ia.y * C);
            v.z = __ldg(src + (long long)ia.z * C);
            v.w = __ldg(src + (long long)ia.w * C);
            *reinterpret_cast<float4 *>(trow + o * 4) = v;
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> async proxy
    __syncthreads();
    if (warp == 0) {
        if (c0 + lane < C) {
            const uint64_t pol = l2_evict_first_policy();
            bulk_store(out + ((long long)b * C + c0 + lane) * P + p0, smem_addr(tile + lane * FWD_LD),
                       (uint32_t)npos * 4u, pol);
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // smem must outlive the reads
    }
}

__device__ __forceinline__ void red_v4(float *p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// grad_out (B,C,P), idx (B,P) -> accT (B,N,C) +=.  V4: vector reductions (needs C % 4 == 0)
template <bool V4>
__global__ void __launch_bounds__(256)
group_bwd_tma_kernel(int C, int N, int P, int nsample, const float *__restrict__ grad_out,
                     const int *__restrict__ idx, float *__restrict__ accT) {
    extern __shared__ __align__(128) float tile[];   // rows at r*BWD_LD + 4*(r>>2)
    __shared__ __align__(8) uint64_t bar;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.z;
    const int c0 = blockIdx.y * TC;
    const int p0 = blockIdx.x * TP;
    const int npos = min(TP, P - p0);
    const int rows = min(TC, C - c0);
    const uint32_t bar_a = smem_addr(&bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (warp == 0) {
        if (lane == 0)
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a),
                         "r"((uint32_t)(rows * npos * 4))
                         : "memory");
        __syncwarp();
        if (lane < rows) {
            const uint64_t pol = l2_evict_first_policy();
            bulk_load(smem_addr(tile + lane * BWD_LD + 4 * (lane >> 2)),
                      grad_out + ((long long)b * C + c0 + lane) * P + p0, (uint32_t)npos * 4u, bar_a, pol);
        }
    }
    // idx of this warp's 32 positions, one per lane (overlaps the bulk loads)
    const int wbase = warp * 32;
    const int wpos = min(32, npos - wbase);
    int my_idx = 0;
    if (lane < wpos) my_idx = __ldg(idx + (long long)b * P + p0 + wbase + lane);
    {
        uint32_t ok = 0;
        while (!ok) {
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
                "selp.u32 %0, 1, 0, p;\n\t}"
                : "=r"(ok)
                : "r"(bar_a)
                : "memory");
        }
    }
    float *acc = accT + (long long)b * N * C + c0;
    if (V4) {
        // lane -> (position j = lane/8 of a group of 4, channel quad q = lane%8)
        const int j = lane >> 3, q = lane & 7;
        const float *t0 = tile + (4 * q) * BWD_LD + 4 * q + wbase + j;
        const bool cok = 4 * q < rows;                 // C % 4 == 0: whole quads only
        // Positions that repeat the first index of their query (ball_query pads a row with its first
        // hit: ~40 % of all slots at PointNeXt's radii) are summed in registers and leave as ONE
        // reduction — the L2 atomic units are what bounds this kernel.  Valid for any idx: equal
        // indices are the same target.  Needs nsample % 32 == 0 so that a warp stays inside one query.
        uint32_t dup = 0;
        int keep = -1;
        if (nsample > 0) {
            const long long gp0 = (long long)p0 + wbase;
            const long long fp = gp0 / nsample * nsample;
            int first_idx = __shfl_sync(0xffffffffu, my_idx, 0);
            if (fp != gp0) first_idx = __ldg(idx + (long long)b * P + fp);
            const uint32_t same = __ballot_sync(0xffffffffu, lane < wpos && my_idx == first_idx);
            if (same & (same - 1)) {
                keep = __ffs(same) - 1;
                dup = same & ~(1u << keep);
            }
        }
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        if (dup != 0) {
#pragma unroll
            for (int o = 0; o < 8; ++o) {
                if (((dup >> (o * 4 + j)) & 1u) && cok) {
                    a0 += t0[o * 4];
                    a1 += t0[o * 4 + BWD_LD];
                    a2 += t0[o * 4 + 2 * BWD_LD];
                    a3 += t0[o * 4 + 3 * BWD_LD];
                }
            }
#pragma unroll
            for (int x = 8; x <= 16; x <<= 1) {
                a0 += __shfl_xor_sync(0xffffffffu, a0, x);
                a1 += __shfl_xor_sync(0xffffffffu, a1, x);
                a2 += __shfl_xor_sync(0xffffffffu, a2, x);
                a3 += __shfl_xor_sync(0xffffffffu, a3, x);
            }
        }
#pragma unroll
        for (int o = 0; o < 8; ++o) {
            const int pp = o * 4 + j;
            const int pi = __shfl_sync(0xffffffffu, my_idx, pp);
            if (pp < wpos && cok && !((dup >> pp) & 1u)) {
                float g0 = t0[o * 4];
                float g1 = t0[o * 4 + BWD_LD];
                float g2 = t0[o * 4 + 2 * BWD_LD];
                float g3 = t0[o * 4 + 3 * BWD_LD];
                if (pp == keep) { g0 += a0; g1 += a1; g2 += a2; g3 += a3; }
                red_v4(acc + (long long)pi * C + 4 * q, g0, g1, g2, g3);
            }
        }
    } else {
        const float *t0 = tile + lane * BWD_LD + 4 * (lane >> 2) + wbase;
        const bool cok = lane < rows;
#pragma unroll
        for (int o = 0; o < 8; ++o) {
            if (o * 4 < wpos) {
                float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
                if (cok) g = *reinterpret_cast<const float4 *>(t0 + o * 4);
                const int i0 = __shfl_sync(0xffffffffu, my_idx, o * 4);
                const int i1 = __shfl_sync(0xffffffffu, my_idx, o * 4 + 1);
                const int i2 = __shfl_sync(0xffffffffu, my_idx, o * 4 + 2);
                const int i3 = __shfl_sync(0xffffffffu, my_idx, o * 4 + 3);
                if (cok) {
                    atomicAdd(acc + (long long)i0 * C + lane, g.x);
                    atomicAdd(acc + (long long)i1 * C + lane, g.y);
                    atomicAdd(acc + (long long)i2 * C + lane, g.z);
                    atomicAdd(acc + (long long)i3 * C + lane, g.w);
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// three_interpolate through the same machinery: the (B,C,m) source is transposed to (B,m,C), a
// warp reads the three neighbour rows of a position as three coalesced 128-byte loads (lane =
// channel), the weighted sum uses the reference's FMA chain, the 32 x 256 output tile leaves by
// bulk-async stores.  idx / weight of the tile's 256 positions are staged in shared memory once.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
interp_fwd_tma_kernel(int C, int M, int N, const float *__restrict__ srcT, const int *__restrict__ idx,
                      const float *__restrict__ weight, float *__restrict__ out) {
    extern __shared__ __align__(128) float tile[];   // [TC][FWD_LD] then idx[256*3], w[256*3]
    int *s_idx = reinterpret_cast<int *>(tile + TC * FWD_LD);
    float *s_w = reinterpret_cast<float *>(s_idx + TP * 3);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.z;
    const int c0 = blockIdx.y * TC;
    const int p0 = blockIdx.x * TP;
    const int npos = min(TP, N - p0);                 // multiple of 4 (host checks N % 4 == 0)
    const long long o3 = 3ll * ((long long)b * N + p0);
    for (int i = threadIdx.x; i < npos * 3; i += 256) {
        s_idx[i] = __ldg(idx + o3 + i);
        s_w[i] = __ldg(weight + o3 + i);
    }
    __syncthreads();
    const int cc = min(c0 + lane, C - 1);
    const float *src = srcT + (long long)b * M * C + cc;
    const int wpos = min(32, npos - warp * 32);
    float *trow = tile + lane * FWD_LD + warp * 32;
    if (wpos == 32) {
        // full warp tile: four batches of 8 positions, the 24 gathers of a batch in flight together
#pragma unroll
        for (int h = 0; h < 4; ++h) {
            float f[8][3];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int p = warp * 32 + h * 8 + u;
#pragma unroll
                for (int k = 0; k < 3; ++k) f[u][k] = __ldg(src + (long long)s_idx[3 * p + k] * C);
            }
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int p = warp * 32 + h * 8 + u;
                // interpolate_gpu.cu:103 as nvcc contracts it
                v[u] = __fmaf_rn(s_w[3 * p + 2], f[u][2], __fmaf_rn(s_w[3 * p], f[u][0], __fmul_rn(s_w[3 * p + 1], f[u][1])));
            }
            *reinterpret_cast<float4 *>(trow + h * 8) = make_float4(v[0], v[1], v[2], v[3]);
            *reinterpret_cast<float4 *>(trow + h * 8 + 4) = make_float4(v[4], v[5], v[6], v[7]);
        }
    } else {
        for (int o = 0; o * 4 < wpos; ++o) {
            float v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int p = warp * 32 + o * 4 + u;
                const int i0 = s_idx[3 * p], i1 = s_idx[3 * p + 1], i2 = s_idx[3 * p + 2];
                const float w0 = s_w[3 * p], w1 = s_w[3 * p + 1], w2 = s_w[3 * p + 2];
                const float f0 = __ldg(src + (long long)i0 * C), f1 = __ldg(src + (long long)i1 * C),
                            f2 = __ldg(src + (long long)i2 * C);
                v[u] = __fmaf_rn(w2, f2, __fmaf_rn(w0, f0, __fmul_rn(w1, f1)));   // interpolate_gpu.cu:103 as nvcc contracts it
            }
            *reinterpret_cast<float4 *>(trow + o * 4) = make_float4(v[0], v[1], v[2], v[3]);
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        if (c0 + lane < C) {
            const uint64_t pol = l2_evict_first_policy();
            bulk_store(out + ((long long)b * C + c0 + lane) * N + p0, smem_addr(tile + lane * FWD_LD),
                       (uint32_t)npos * 4u, pol);
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
}

// grad_out (B,C,N), idx/weight (B,N,3) -> accT (B,M,C) += ; requires C % 4 == 0, N % 4 == 0
__global__ void __launch_bounds__(256)
interp_bwd_tma_kernel(int C, int N, int M, const float *__restrict__ grad_out, const int *__restrict__ idx,
                      const float *__restrict__ weight, float *__restrict__ accT) {
    extern __shared__ __align__(128) float tile[];   // rows at r*BWD_LD + 4*(r>>2); then idx, w
    __shared__ __align__(8) uint64_t bar;
    int *s_idx = reinterpret_cast<int *>(tile + TC * BWD_LD + 32);
    float *s_w = reinterpret_cast<float *>(s_idx + TP * 3);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.z;
    const int c0 = blockIdx.y * TC;
    const int p0 = blockIdx.x * TP;
    const int npos = min(TP, N - p0);
    const int rows = min(TC, C - c0);
    const uint32_t bar_a = smem_addr(&bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (warp == 0) {
        if (lane == 0)
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a),
                         "r"((uint32_t)(rows * npos * 4))
                         : "memory");
        __syncwarp();
        if (lane < rows) {
            const uint64_t pol = l2_evict_first_policy();
            bulk_load(smem_addr(tile + lane * BWD_LD + 4 * (lane >> 2)),
                      grad_out + ((long long)b * C + c0 + lane) * N + p0, (uint32_t)npos * 4u, bar_a, pol);
        }
    }
    const long long o3 = 3ll * ((long long)b * N + p0);
    for (int i = threadIdx.x; i < npos * 3; i += 256) {
        s_idx[i] = __ldg(idx + o3 + i);
        s_w[i] = __ldg(weight + o3 + i);
    }
    __syncthreads();
    {
        uint32_t ok = 0;
        while (!ok) {
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
                "selp.u32 %0, 1, 0, p;\n\t}"
                : "=r"(ok)
                : "r"(bar_a)
                : "memory");
        }
    }
    float *acc = accT + (long long)b * M * C + c0;
    const int wbase = warp * 32;
    const int wpos = min(32, npos - wbase);
    const int j = lane >> 3, q = lane & 7;
    const float *t0 = tile + (4 * q) * BWD_LD + 4 * q + wbase + j;
    const bool cok = 4 * q < rows;
#pragma unroll
    for (int o = 0; o < 8; ++o) {
        if (o * 4 + j < wpos && cok) {
            const int p = wbase + o * 4 + j;
            const float g0 = t0[o * 4];
            const float g1 = t0[o * 4 + BWD_LD];
            const float g2 = t0[o * 4 + 2 * BWD_LD];
            const float g3 = t0[o * 4 + 3 * BWD_LD];
#pragma unroll
            for (int u = 0; u < 3; ++u) {
                const float w = s_w[3 * p + u];
                red_v4(acc + (long long)s_idx[3 * p + u] * C + 4 * q, __fmul_rn(g0, w), __fmul_rn(g1, w),
                       __fmul_rn(g2, w), __fmul_rn(g3, w));
            }
        }
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

// QueryAndGroup's relative coordinates in one launch (group.py:244-249 of the reference composes them from a
// transpose, a C = 3 grouping, a broadcast subtraction and a division by the radius):
//     out[b,c,j,s] = (xyz[b, idx[b,j,s], c] - query[b,j,c]) * inv_radius
// with inv_radius = 1.0f / (float)radius, which is how torch's CUDA true-division by a Python scalar is
// evaluated (a * reciprocal(b), BinaryDivTrueKernel.cu) — so the result is bit-identical to the composition.
// subtract == 0 leaves out the query, inv_radius == 0 the scaling.  xyz (B,N,3), query (B,M,3), out (B,3,M,ns).
__global__ void __launch_bounds__(256)
group_xyz_rel_kernel(int N, int M, int nsample, int subtract, float inv_radius, const float *__restrict__ xyz,
                     const float *__restrict__ query, const int *__restrict__ idx, float *__restrict__ out) {
    const int b = blockIdx.y;
    const long long P = (long long)M * nsample;
    const long long p = (long long)blockIdx.x * 256 + threadIdx.x;
    if (p >= P) return;
    const int i = __ldg(idx + (long long)b * P + p);
    const float *src = xyz + 3ll * ((long long)b * N + i);
    float v[3] = {__ldg(src), __ldg(src + 1), __ldg(src + 2)};
    if (subtract) {
        const float *q = query + 3ll * ((long long)b * M + p / nsample);
#pragma unroll
        for (int c = 0; c < 3; ++c) v[c] = __fsub_rn(v[c], __ldg(q + c));
    }
    if (inv_radius != 0.f) {
#pragma unroll
        for (int c = 0; c < 3; ++c) v[c] = __fmul_rn(v[c], inv_radius);
    }
    float *o = out + 3ll * b * P + p;
#pragma unroll
    for (int c = 0; c < 3; ++c) o[c * P] = v[c];
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

// AMC3D_GROUP_IMPL: 0 = register-only kernels, 1 = TMA-staged (scalar reductions in the backward),
// 2 = TMA-staged with vector reductions (default)
static int group_impl() {
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("AMC3D_GROUP_IMPL");
        v = e ? atoi(e) : 2;
    }
    return v;
}

static int group_common(bool fwd, int b, int c, int n, long long P, const float *src, const int *idx,
                        float *dst, float *workspace, cudaStream_t st, const char *what, int nsample = 0,
                        bool overwrite = false) {
    if (b == 0 || c == 0 || P == 0) {
        // nothing to scatter: the overwriting variants still owe the caller a zero gradient (torch.empty there)
        if (!fwd && overwrite && (size_t)b * n * c > 0)
            cudaMemsetAsync(dst, 0, sizeof(float) * (size_t)b * n * c, st);
        return check_launch(what);
    }
    AMC3D_REQUIRE(b <= 65535, AMC3D_ELIMIT, "%s: batch %d > 65535", what, b);
    AMC3D_REQUIRE(P < (1ll << 31), AMC3D_ELIMIT, "%s: npoints*nsample too large", what);
    const bool aligned = (P % 8 == 0) && ((reinterpret_cast<uintptr_t>(idx) & 15) == 0) &&
                         ((reinterpret_cast<uintptr_t>(fwd ? dst : const_cast<float *>(src)) & 15) == 0);
    if (workspace != nullptr && c >= 8 && aligned && n > 0) {
        dim3 grid((unsigned)div_up_ll(P, GRP_WARPS * 32), div_up(c, 32), b);
        const int impl = group_impl();
        static const bool no_combine = getenv("AMC3D_GROUP_NOCOMBINE") != nullptr;   // for measurements
        const int combine = (!no_combine && nsample >= 32 && nsample % 32 == 0) ? nsample : 0;
        if (fwd) {
            launch_transpose<false>(b, c, n, src, workspace, st);  // (B,C,N) -> (B,N,C)
            if (impl >= 1) {
                static const int fw = getenv("AMC3D_GROUP_FWD_W") ? atoi(getenv("AMC3D_GROUP_FWD_W")) : 4;   // 4 warps x 128 positions measured best
                if (fw == 4) {
                    dim3 g4((unsigned)div_up_ll(P, 128), div_up(c, 32), b);
                    group_fwd_tma_kernel<4><<<g4, 128, TC * (128 + 4) * sizeof(float), st>>>(c, n, (int)P, combine, workspace, idx, dst);
                } else if (fw == 2) {
                    dim3 g2((unsigned)div_up_ll(P, 64), div_up(c, 32), b);
                    group_fwd_tma_kernel<2><<<g2, 64, TC * (64 + 4) * sizeof(float), st>>>(c, n, (int)P, combine, workspace, idx, dst);
                } else {
                    group_fwd_tma_kernel<8><<<grid, 256, TC * (256 + 4) * sizeof(float), st>>>(c, n, (int)P, combine, workspace, idx, dst);
                }
            }
            else
                group_fwd_cl_kernel<<<grid, GRP_WARPS * 32, 0, st>>>(c, n, (int)P, workspace, idx, dst);
        } else {
            cudaMemsetAsync(workspace, 0, sizeof(float) * (size_t)b * n * c, st);
            const size_t smem = (TC * BWD_LD + 32) * sizeof(float);
            if (impl >= 2 && c % 4 == 0)
                group_bwd_tma_kernel<true><<<grid, 256, smem, st>>>(c, n, (int)P, combine, src, idx, workspace);
            else if (impl >= 1)
                group_bwd_tma_kernel<false><<<grid, 256, smem, st>>>(c, n, (int)P, 0, src, idx, workspace);
            else
                group_bwd_cl_kernel<<<grid, GRP_WARPS * 32, 0, st>>>(c, n, (int)P, src, idx, workspace);
            if (overwrite) launch_transpose<false>(b, n, c, workspace, dst, st);   // (B,N,C) -> (B,C,N)
            else launch_transpose<true>(b, n, c, workspace, dst, st);              // (B,N,C) -> += (B,C,N)
        }
    } else {
        if (!fwd && overwrite) cudaMemsetAsync(dst, 0, sizeof(float) * (size_t)b * n * c, st);
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

// (B, rows, cols) -> (B, cols, rows): the channel-contiguous copy the fused operator gathers from
extern "C" int amc3d_transpose_batched(int b, int rows, int cols, const float *src, float *dst, void *stream) {
    AMC3D_REQUIRE(b >= 0 && rows >= 0 && cols >= 0, AMC3D_EINVAL, "transpose_batched: negative size");
    AMC3D_REQUIRE(b <= 65535, AMC3D_ELIMIT, "transpose_batched: batch %d > 65535", b);
    if (b == 0 || rows == 0 || cols == 0) return 0;
    launch_transpose<false>(b, rows, cols, src, dst, as_stream(stream));
    return check_launch("transpose_batched");
}

extern "C" int amc3d_group_points_ws(int b, int c, int n, int npoints, int nsample, const float *points,
                                     const int *idx, float *out, float *workspace, void *stream) {
    AMC3D_REQUIRE(b >= 0 && c >= 0 && n >= 0 && npoints >= 0 && nsample >= 0, AMC3D_EINVAL,
                  "group_points: negative size");
    return group_common(true, b, c, n, (long long)npoints * nsample, points, idx, out, workspace,
                        as_stream(stream), "group_points", nsample);
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
                        workspace, as_stream(stream), "group_points_grad", nsample);
}
extern "C" int amc3d_group_points_grad_ws_set(int b, int c, int n, int npoints, int nsample,
                                              const float *grad_out, const int *idx, float *grad_points,
                                              float *workspace, void *stream) {
    AMC3D_REQUIRE(b >= 0 && c >= 0 && n >= 0 && npoints >= 0 && nsample >= 0, AMC3D_EINVAL,
                  "group_points_grad: negative size");
    return group_common(false, b, c, n, (long long)npoints * nsample, grad_out, idx, grad_points,
                        workspace, as_stream(stream), "group_points_grad", nsample, true);
}

extern "C" int amc3d_group_points_grad(int b, int c, int n, int npoints, int nsample,
                                       const float *grad_out, const int *idx, float *grad_points,
                                       void *stream) {
    return amc3d_group_points_grad_ws(b, c, n, npoints, nsample, grad_out, idx, grad_points, nullptr, stream);
}

extern "C" int amc3d_group_xyz_relative(int b, int n, int m, int nsample, int subtract, float inv_radius,
                                        const float *xyz, const float *query, const int *idx, float *out,
                                        void *stream) {
    AMC3D_REQUIRE(b >= 0 && n >= 0 && m >= 0 && nsample >= 0, AMC3D_EINVAL, "group_xyz_relative: negative size");
    AMC3D_REQUIRE(b <= 65535, AMC3D_ELIMIT, "group_xyz_relative: batch %d > 65535", b);
    const long long P = (long long)m * nsample;
    if (b == 0 || P == 0) return 0;
    dim3 grid((unsigned)div_up_ll(P, 256), b);
    group_xyz_rel_kernel<<<grid, 256, 0, as_stream(stream)>>>(n, m, nsample, subtract, inv_radius, xyz, query, idx, out);
    return check_launch("group_xyz_relative");
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

static inline bool al16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

extern "C" int amc3d_three_interpolate_ws(int b, int c, int m, int n, const float *points, const int *idx,
                                          const float *weight, float *out, float *workspace, void *stream) {
    AMC3D_REQUIRE(b >= 0 && c >= 0 && m >= 0 && n >= 0, AMC3D_EINVAL, "three_interpolate: negative size");
    AMC3D_REQUIRE(b <= 65535, AMC3D_ELIMIT, "three_interpolate: batch %d > 65535", b);
    if (b == 0 || c == 0 || n == 0) return 0;
    if (workspace != nullptr && c >= 8 && n % 4 == 0 && m > 0 && al16(out) && group_impl() >= 1) {
        cudaStream_t st = as_stream(stream);
        launch_transpose<false>(b, c, m, points, workspace, st);   // (B,C,m) -> (B,m,C)
        dim3 grid(div_up(n, TP), div_up(c, TC), b);
        const size_t smem = (TC * FWD_LD + TP * 6) * sizeof(float);
        interp_fwd_tma_kernel<<<grid, 256, smem, st>>>(c, m, n, workspace, idx, weight, out);
        return check_launch("three_interpolate");
    }
    const int bx = div_up(n, 256);
    const int cchunk = pick_cchunk(c, (long long)bx * b);
    dim3 grid(bx, div_up(c, cchunk), b);
    three_interp_kernel<<<grid, 256, 0, as_stream(stream)>>>(c, m, n, cchunk, points, idx, weight, out);
    return check_launch("three_interpolate");
}

extern "C" int amc3d_three_interpolate(int b, int c, int m, int n, const float *points, const int *idx,
                                       const float *weight, float *out, void *stream) {
    return amc3d_three_interpolate_ws(b, c, m, n, points, idx, weight, out, nullptr, stream);
}

static int interp_grad_common(int b, int c, int n, int m, const float *grad_out, const int *idx,
                              const float *weight, float *grad_points, float *workspace, void *stream,
                              bool overwrite);

extern "C" int amc3d_three_interpolate_grad_ws(int b, int c, int n, int m, const float *grad_out,
                                               const int *idx, const float *weight, float *grad_points,
                                               float *workspace, void *stream) {
    return interp_grad_common(b, c, n, m, grad_out, idx, weight, grad_points, workspace, stream, false);
}

extern "C" int amc3d_three_interpolate_grad_ws_set(int b, int c, int n, int m, const float *grad_out,
                                                   const int *idx, const float *weight, float *grad_points,
                                                   float *workspace, void *stream) {
    return interp_grad_common(b, c, n, m, grad_out, idx, weight, grad_points, workspace, stream, true);
}

static int interp_grad_common(int b, int c, int n, int m, const float *grad_out, const int *idx,
                              const float *weight, float *grad_points, float *workspace, void *stream,
                              bool overwrite) {
    AMC3D_REQUIRE(b >= 0 && c >= 0 && m >= 0 && n >= 0, AMC3D_EINVAL, "three_interpolate_grad: negative size");
    AMC3D_REQUIRE(b <= 65535, AMC3D_ELIMIT, "three_interpolate_grad: batch %d > 65535", b);
    if (b == 0 || c == 0 || n == 0) {
        if (overwrite && (size_t)b * m * c > 0)
            cudaMemsetAsync(grad_points, 0, sizeof(float) * (size_t)b * m * c, as_stream(stream));
        return check_launch("three_interpolate_grad");
    }
    if (workspace != nullptr && c >= 8 && c % 4 == 0 && n % 4 == 0 && m > 0 && al16(grad_out) && group_impl() >= 1) {
        cudaStream_t st = as_stream(stream);
        cudaMemsetAsync(workspace, 0, sizeof(float) * (size_t)b * m * c, st);
        dim3 grid(div_up(n, TP), div_up(c, TC), b);
        const size_t smem = (TC * BWD_LD + 32 + TP * 6) * sizeof(float);
        interp_bwd_tma_kernel<<<grid, 256, smem, st>>>(c, n, m, grad_out, idx, weight, workspace);
        if (overwrite) launch_transpose<false>(b, m, c, workspace, grad_points, st);   // (B,m,C) -> (B,C,m)
        else launch_transpose<true>(b, m, c, workspace, grad_points, st);              // (B,m,C) -> += (B,C,m)
        return check_launch("three_interpolate_grad");
    }
    if (overwrite) cudaMemsetAsync(grad_points, 0, sizeof(float) * (size_t)b * m * c, as_stream(stream));
    const int bx = div_up(n, 256);
    const int cchunk = pick_cchunk(c, (long long)bx * b);
    dim3 grid(bx, div_up(c, cchunk), b);
    three_interp_grad_kernel<<<grid, 256, 0, as_stream(stream)>>>(c, n, m, cchunk, grad_out, idx, weight,
                                                                  grad_points);
    return check_launch("three_interpolate_grad");
}

extern "C" int amc3d_three_interpolate_grad(int b, int c, int n, int m, const float *grad_out,
                                            const int *idx, const float *weight, float *grad_points,
                                            void *stream) {
    return amc3d_three_interpolate_grad_ws(b, c, n, m, grad_out, idx, weight, grad_points, nullptr, stream);
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
