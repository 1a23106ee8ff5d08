// Input side of the path: the voxel hash of `voxelize` and the crop distances of `crop_pc`
// (ref: openpoints/dataset/data_util.py:92-134 fnv_hash_vec / ravel_hash_vec / voxelize, :137-174 crop_pc), so that
// a batch can be voxel-downsampled and cropped where it already lives instead of in DataLoader workers.
// HBM-bound integer work: one coalesced pass per point.
#include "common.cuh"

namespace amc3d {

// key[i] = FNV64-1A over the three cell coordinates floor(coord / voxel_size) (data_util.py:92-105):
//   h = 14695981039346656037;  for j in x, y, z:  h *= 1099511628211;  h ^= uint64(cell_j)
// The division and floor are FP64, as numpy evaluates `coord / np.array(voxel_size)` for a Python float voxel size.
__global__ void __launch_bounds__(256)
voxel_keys_fnv_kernel(long long n, double voxel_size, const float *__restrict__ coord, unsigned long long *__restrict__ key,
                      long long *__restrict__ cell) {
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    unsigned long long h = 14695981039346656037ull;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const double c = floor((double)__ldg(coord + 3 * i + j) / voxel_size);
        const long long ci = (long long)c;                  // numpy: float64 -> uint64 (two's complement for negatives)
        if (cell != nullptr) cell[3 * i + j] = ci;
        h *= 1099511628211ull;
        h ^= (unsigned long long)ci;
    }
    key[i] = h;
}

// out[i] = sum((coord[i] - coord[init])^2) in FP32, summed (dx2 + dy2) + dz2 as numpy's np.sum over 3 elements
__global__ void __launch_bounds__(256)
crop_dist2_kernel(long long n, const float *__restrict__ coord, long long init, float *__restrict__ out) {
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const float dx = __fsub_rn(__ldg(coord + 3 * i), __ldg(coord + 3 * init));
    const float dy = __fsub_rn(__ldg(coord + 3 * i + 1), __ldg(coord + 3 * init + 1));
    const float dz = __fsub_rn(__ldg(coord + 3 * i + 2), __ldg(coord + 3 * init + 2));
    out[i] = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

}  // namespace amc3d

using namespace amc3d;

extern "C" int amc3d_voxel_keys(long long n, double voxel_size, const float *coord, unsigned long long *keys,
                                long long *cells, void *stream) {
    AMC3D_REQUIRE(n >= 0 && voxel_size > 0.0, AMC3D_EINVAL, "voxel_keys: bad arguments n=%lld voxel_size=%g", n, voxel_size);
    if (n == 0) return 0;
    voxel_keys_fnv_kernel<<<(unsigned)div_up_ll(n, 256), 256, 0, as_stream(stream)>>>(n, voxel_size, coord, keys, cells);
    return check_launch("voxel_keys");
}

extern "C" int amc3d_crop_dist2(long long n, const float *coord, long long init_idx, float *out, void *stream) {
    AMC3D_REQUIRE(n >= 0 && init_idx >= 0 && (n == 0 || init_idx < n), AMC3D_EINVAL, "crop_dist2: bad arguments n=%lld init=%lld", n, init_idx);
    if (n == 0) return 0;
    crop_dist2_kernel<<<(unsigned)div_up_ll(n, 256), 256, 0, as_stream(stream)>>>(n, coord, init_idx, out);
    return check_launch("crop_dist2");
}
