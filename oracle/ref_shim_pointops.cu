// C-ABI shim over the UNMODIFIED reference launchers of the remaining pointops operators, for
// oracle/_ref/libref_pointops.so.  TEST INFRASTRUCTURE ONLY — see ref_shim.cu.  A second library because
// pointops/src/sampling/sampling_cuda_kernel.cu and pointnet2_batch/src/sampling_gpu.cu both define a
// global `__update` and cannot be linked together.
#include <cuda_runtime.h>

// openpoints/cpp/pointops/src/{sampling,ballquery,interpolation,subtraction,aggregation}/*_cuda_kernel.h
extern "C" void furthestsampling_cuda_launcher(int b, int n, const float *xyz, const int *offset, const int *new_offset, float *tmp, int *idx);
void ballquery_launcher(int m, float radius, int nsample, const float *xyz, const float *new_xyz, const int *offset, const int *new_offset, int *idx);
extern "C" void interpolation_forward_cuda_launcher(int n, int c, int k, const float *input, const int *idx, const float *weight, float *output);
extern "C" void interpolation_backward_cuda_launcher(int n, int c, int k, const float *grad_output, const int *idx, const float *weight, float *grad_input);
extern "C" void subtraction_forward_cuda_launcher(int n, int nsample, int c, const float *input1, const float *input2, const int *idx, float *output);
extern "C" void subtraction_backward_cuda_launcher(int n, int nsample, int c, const int *idx, const float *grad_output, float *grad_input1, float *grad_input2);
extern "C" void aggregation_forward_cuda_launcher(int n, int nsample, int c, int w_c, const float *input, const float *position, const float *weight, const int *idx, float *output);
extern "C" void aggregation_backward_cuda_launcher(int n, int nsample, int c, int w_c, const float *input, const float *position, const float *weight, const int *idx, const float *grad_output, float *grad_input, float *grad_position, float *grad_weight);


extern "C" {
void ref_pop_fps(int b, int n_max, const float *xyz, const int *offset, const int *new_offset, float *tmp, int *idx) { furthestsampling_cuda_launcher(b, n_max, xyz, offset, new_offset, tmp, idx); }
void ref_pop_ballquery(int m, float radius, int nsample, const float *xyz, const float *new_xyz, const int *offset, const int *new_offset, int *idx) { ballquery_launcher(m, radius, nsample, xyz, new_xyz, offset, new_offset, idx); }
void ref_pop_interpolation_fwd(int n, int c, int k, const float *input, const int *idx, const float *weight, float *output) { interpolation_forward_cuda_launcher(n, c, k, input, idx, weight, output); }
void ref_pop_interpolation_bwd(int n, int c, int k, const float *grad_output, const int *idx, const float *weight, float *grad_input) { interpolation_backward_cuda_launcher(n, c, k, grad_output, idx, weight, grad_input); }
void ref_pop_subtraction_fwd(int n, int nsample, int c, const float *a, const float *b, const int *idx, float *out) { subtraction_forward_cuda_launcher(n, nsample, c, a, b, idx, out); }
void ref_pop_subtraction_bwd(int n, int nsample, int c, const int *idx, const float *g, float *g1, float *g2) { subtraction_backward_cuda_launcher(n, nsample, c, idx, g, g1, g2); }
void ref_pop_aggregation_fwd(int n, int nsample, int c, int w_c, const float *input, const float *position, const float *weight, const int *idx, float *output) { aggregation_forward_cuda_launcher(n, nsample, c, w_c, input, position, weight, idx, output); }
void ref_pop_aggregation_bwd(int n, int nsample, int c, int w_c, const float *input, const float *position, const float *weight, const int *idx, const float *g, float *gi, float *gp, float *gw) { aggregation_backward_cuda_launcher(n, nsample, c, w_c, input, position, weight, idx, g, gi, gp, gw); }
int ref_pop_sync(void) { return (int)cudaDeviceSynchronize(); }
}
