// C-ABI shim over the UNMODIFIED reference launchers, for oracle/_ref/libref_kernels.so.
//
// TEST INFRASTRUCTURE ONLY.  The reference sources are compiled where they lie under
// /root/reference (never copied into this repository); this file only re-exports their
// raw-pointer launchers under extern "C" names so that tests and bench.py can call the
// reference's own CUDA kernels on the B200 through ctypes.  The launchers run on the legacy
// default stream, exactly as in the reference.
//
// Declarations restate the prototypes in the reference headers
// (openpoints/cpp/pointnet2_batch/src/{sampling,ball_query,group_points,interpolate}_gpu.h and
// openpoints/cpp/pointops/src/knnquery/knnquery_cuda_kernel.h) without including them, so the
// shim itself needs no torch headers.
#include <cuda_runtime.h>

void furthest_point_sampling_kernel_launcher(int b, int n, int m, const float *dataset, float *temp, int *idxs);
void gather_points_kernel_launcher_fast(int b, int c, int n, int npoints, const float *points, const int *idx, float *out);
void gather_points_grad_kernel_launcher_fast(int b, int c, int n, int npoints, const float *grad_out, const int *idx, float *grad_points);
void ball_query_kernel_launcher_fast(int b, int n, int m, float radius, int nsample, const float *xyz, const float *new_xyz, int *idx);
void group_points_kernel_launcher_fast(int b, int c, int n, int npoints, int nsample, const float *points, const int *idx, float *out);
void group_points_grad_kernel_launcher_fast(int b, int c, int n, int npoints, int nsample, const float *grad_out, const int *idx, float *grad_points);
void three_nn_kernel_launcher_fast(int b, int n, int m, const float *unknown, const float *known, float *dist2, int *idx);
void three_interpolate_kernel_launcher_fast(int b, int c, int m, int n, const float *points, const int *idx, const float *weight, float *out);
void three_interpolate_grad_kernel_launcher_fast(int b, int c, int n, int m, const float *grad_out, const int *idx, const float *weight, float *grad_points);
extern "C" void knnquery_cuda_launcher(int m, int nsample, const float *xyz, const float *new_xyz, const int *offset, const int *new_offset, int *idx, float *dist2);

extern "C" {
void ref_fps(int b, int n, int m, const float *xyz, float *temp, int *idx) { furthest_point_sampling_kernel_launcher(b, n, m, xyz, temp, idx); }
// NOTE the reference header names these two arguments (xyz, new_xyz) but the definition and
// its caller pass (new_xyz, xyz) — ball_query_gpu.cu:54, ball_query.cpp:40.
void ref_ball_query(int b, int n, int m, float radius, int nsample, const float *new_xyz, const float *xyz, int *idx) { ball_query_kernel_launcher_fast(b, n, m, radius, nsample, new_xyz, xyz, idx); }
void ref_group_points(int b, int c, int n, int npoints, int nsample, const float *points, const int *idx, float *out) { group_points_kernel_launcher_fast(b, c, n, npoints, nsample, points, idx, out); }
void ref_group_points_grad(int b, int c, int n, int npoints, int nsample, const float *grad_out, const int *idx, float *grad_points) { group_points_grad_kernel_launcher_fast(b, c, n, npoints, nsample, grad_out, idx, grad_points); }
void ref_gather_points(int b, int c, int n, int npoints, const float *points, const int *idx, float *out) { gather_points_kernel_launcher_fast(b, c, n, npoints, points, idx, out); }
void ref_gather_points_grad(int b, int c, int n, int npoints, const float *grad_out, const int *idx, float *grad_points) { gather_points_grad_kernel_launcher_fast(b, c, n, npoints, grad_out, idx, grad_points); }
void ref_three_nn(int b, int n, int m, const float *unknown, const float *known, float *dist2, int *idx) { three_nn_kernel_launcher_fast(b, n, m, unknown, known, dist2, idx); }
void ref_three_interpolate(int b, int c, int m, int n, const float *points, const int *idx, const float *weight, float *out) { three_interpolate_kernel_launcher_fast(b, c, m, n, points, idx, weight, out); }
void ref_three_interpolate_grad(int b, int c, int n, int m, const float *grad_out, const int *idx, const float *weight, float *grad_points) { three_interpolate_grad_kernel_launcher_fast(b, c, n, m, grad_out, idx, weight, grad_points); }
void ref_knnquery(int m, int nsample, const float *xyz, const float *new_xyz, const int *offset, const int *new_offset, int *idx, float *dist2) { knnquery_cuda_launcher(m, nsample, xyz, new_xyz, offset, new_offset, idx, dist2); }
int ref_sync(void) { return (int)cudaDeviceSynchronize(); }
}
