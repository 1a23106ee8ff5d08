// Micro-benchmark: latency of one all-to-all "candidate exchange" round inside a thread-block
// cluster, for the mechanisms considered for FPS (DESIGN.md "FPS").  Prints cycles per round.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void csync() { asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ uint4 ldv4(const void *p) {
    uint4 r;
    asm volatile("ld.volatile.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(smem_u32(p)));
    return r;
}
__device__ __forceinline__ uint32_t mapa(uint32_t a, uint32_t r) { uint32_t o; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(o) : "r"(a), "r"(r)); return o; }

// V1: st.async + mbarrier complete_tx, every warp sends to every CTA
template <int CS, int NW>
__global__ void v1(int rounds, long long *out) {
    constexpr int NE = CS * NW;
    __shared__ __align__(16) uint32_t ex[2][NE][8];
    __shared__ __align__(8) uint64_t bar[2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, rank = ctarank();
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[1])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    csync();
    const uint32_t peer = lane < CS ? lane : 0;
    uint32_t d0 = mapa(smem_u32(&ex[0][rank * NW + warp][0]), peer), d1 = mapa(smem_u32(&ex[1][rank * NW + warp][0]), peer);
    uint32_t b0 = mapa(smem_u32(&bar[0]), peer), b1 = mapa(smem_u32(&bar[1]), peer);
    uint32_t par = 0, phase = 0, acc = 0;
    long long t0 = clock64();
    for (int j = 1; j <= rounds; ++j) {
        const uint32_t lb = smem_u32(&bar[par]);
        if (tid == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(lb), "r"(NE * 32) : "memory");
        if (lane < CS) {
            const uint32_t d = par ? d1 : d0, b = par ? b1 : b0;
            asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1,%2,%3,%4}, [%5];" ::"r"(d), "r"(j + acc), "r"(j), "r"(j), "r"(j), "r"(b) : "memory");
            asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1,%2,%3,%4}, [%5];" ::"r"(d + 16), "r"(j), "r"(j), "r"(j), "r"(j), "r"(b) : "memory");
        }
        asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@!p bra W;\n\t}" ::"r"(lb), "r"(phase) : "memory");
        uint32_t v = lane < NE ? ex[par][lane][0] : 0;
        if (NE > 32 && lane + 32 < NE) v ^= ex[par][lane + 32][0];
        acc = __reduce_max_sync(0xffffffffu, v) & 1;
        par ^= 1; if (par == 0) phase ^= 1;
    }
    long long t1 = clock64();
    csync();
    if (blockIdx.x == 0 && tid == 0) { out[0] = (t1 - t0) / rounds; out[1] = acc; }
}

// V2: plain remote 16-byte stores carrying a round tag, receivers poll their local smem
template <int CS, int NW>
__global__ void v2(int rounds, long long *out) {
    constexpr int NE = CS * NW;
    __shared__ __align__(16) uint32_t ex[2][NE][8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, rank = ctarank();
    for (int i = tid; i < 2 * NE * 8; i += blockDim.x) (&ex[0][0][0])[i] = 0;
    __syncthreads();
    csync();
    const uint32_t peer = lane < CS ? lane : 0;
    uint32_t d0 = mapa(smem_u32(&ex[0][rank * NW + warp][0]), peer), d1 = mapa(smem_u32(&ex[1][rank * NW + warp][0]), peer);
    uint32_t par = 0, acc = 0;
    long long t0 = clock64();
    for (int j = 1; j <= rounds; ++j) {
        if (lane < CS) {
            const uint32_t d = par ? d1 : d0;
            asm volatile("st.shared::cluster.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(d), "r"(j + acc), "r"(j), "r"(j), "r"(j) : "memory");
            asm volatile("st.shared::cluster.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(d + 16), "r"(j), "r"(j), "r"(j), "r"(j) : "memory");
        }
        uint32_t v = 0;
        bool ok;
        do {
            ok = true;
            if (lane < NE) {
                uint4 a = ldv4(&ex[par][lane][0]);
                uint4 b = ldv4(&ex[par][lane][4]);
                ok = a.w == (uint32_t)j && b.w == (uint32_t)j;
                v = a.x;
            }
            if (NE > 32 && lane + 32 < NE) {
                uint4 a = ldv4(&ex[par][lane + 32][0]);
                uint4 b = ldv4(&ex[par][lane + 32][4]);
                ok = ok && a.w == (uint32_t)j && b.w == (uint32_t)j;
                v ^= a.x;
            }
        } while (!__all_sync(0xffffffffu, ok));
        acc = __reduce_max_sync(0xffffffffu, v) & 1;
        par ^= 1;
    }
    long long t1 = clock64();
    csync();
    if (blockIdx.x == 0 && tid == 0) { out[0] = (t1 - t0) / rounds; out[1] = acc; }
}

// V0: no exchange at all (REDUX + loop overhead only)
__global__ void v0(int rounds, long long *out) {
    uint32_t acc = threadIdx.x;
    long long t0 = clock64();
    for (int j = 1; j <= rounds; ++j) acc = __reduce_max_sync(0xffffffffu, acc + j) & 1023;
    long long t1 = clock64();
    if (blockIdx.x == 0 && threadIdx.x == 0) { out[0] = (t1 - t0) / rounds; out[1] = acc; }
}

template <typename K>
static void run(const char *name, K kern, int cs, int nw, int clusters, int rounds, long long *d_out) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cs * clusters);
    cfg.blockDim = dim3(nw * 32);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    if (cs > 8) cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, rounds, d_out);
    cudaError_t e2 = cudaDeviceSynchronize();
    long long h[2] = {0, 0};
    cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
    printf("%-28s CS=%2d warps=%d clusters=%d : %6lld cycles/round  (%s %s)\n", name, cs, nw, clusters, h[0], cudaGetErrorString(e), cudaGetErrorString(e2));
}

int main() {
    long long *d_out; cudaMalloc(&d_out, 16);
    const int R = 20000;
    v0<<<1, 128>>>(R, d_out); cudaDeviceSynchronize();
    long long h[2]; cudaMemcpy(h, d_out, 16, cudaMemcpyDeviceToHost); printf("v0 redux-only loop: %lld cycles/round\n", h[0]);
    for (int clusters : {1, 8}) {
        run("v1 st.async+mbarrier", v1<1, 4>, 1, 4, clusters, R, d_out);
        run("v1 st.async+mbarrier", v1<4, 4>, 4, 4, clusters, R, d_out);
        run("v1 st.async+mbarrier", v1<8, 4>, 8, 4, clusters, R, d_out);
        run("v1 st.async+mbarrier", v1<16, 4>, 16, 4, clusters, R, d_out);
        run("v1 st.async+mbarrier", v1<8, 1>, 8, 1, clusters, R, d_out);
        run("v1 st.async+mbarrier", v1<16, 1>, 16, 1, clusters, R, d_out);
        run("v2 st.cluster+tag poll", v2<1, 4>, 1, 4, clusters, R, d_out);
        run("v2 st.cluster+tag poll", v2<4, 4>, 4, 4, clusters, R, d_out);
        run("v2 st.cluster+tag poll", v2<8, 4>, 8, 4, clusters, R, d_out);
        run("v2 st.cluster+tag poll", v2<16, 4>, 16, 4, clusters, R, d_out);
        run("v2 st.cluster+tag poll", v2<8, 1>, 8, 1, clusters, R, d_out);
        run("v2 st.cluster+tag poll", v2<16, 1>, 16, 1, clusters, R, d_out);
        run("v2 st.cluster+tag poll", v2<16, 2>, 16, 2, clusters, R, d_out);
    }
    return 0;
}
