/*
 * amc3d.h — C-ABI of the B200-native (sm_100a) point-grouping operators and the
 * adaptive-margin contrastive loss of AMContrast3D.
 *
 * This is the drop-in boundary (Tier 1).  Every entry point takes plain
 * device pointers, sizes and a CUDA stream (as void*; NULL = legacy default stream, which
 * is what the reference launches on).  No torch types.  Outputs and scratch are allocated
 * by the caller, exactly as in the reference, whose Python wrappers allocate every tensor
 * and whose extension functions only borrow raw pointers for the duration of the launch.
 *
 * Return value: 0 on success, otherwise a negative AMC3D_E* code (argument errors) or the
 * positive cudaError_t of the failed launch.  The reference calls exit(-1) on a launch
 * failure (ball_query_gpu.cu:68-72, sampling_gpu.cu:255-259); a library must not, so the
 * host side (amcontrast3d_b200/_capi.py) turns a non-zero return into a RuntimeError.
 *
 * "ref:" lines cite the reference interface each function replaces, relative to
 * openpoints/cpp/ in YangChenApril/AMContrast3D.
 */
#ifndef AMC3D_H
#define AMC3D_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AMC3D_EINVAL (-1)   /* bad size / unsupported argument */
#define AMC3D_ELIMIT (-2)   /* exceeds a documented limit (e.g. nsample > 128) */

/* library / build identification; amc3d_version() returns AMC3D_VERSION */
#define AMC3D_VERSION 100
int amc3d_version(void);
/* "sm_100a" — the only architecture this library is built for */
const char *amc3d_arch(void);
/* text of the last error on this thread (never NULL) */
const char *amc3d_last_error(void);
/* The culled searches (knnquery, ball_query, three_nn on clouds of >= 2048 points) take their
 * scratch, stream-ordered, from a memory pool PRIVATE to this library (one per device, created on
 * first use); the pool keeps what the largest call needed so that later calls do not allocate.
 * amc3d_trim_scratch returns everything above `keep_bytes` on the current device to the driver.
 * No other global state: nothing here touches the device's default memory pool. */
int amc3d_trim_scratch(size_t keep_bytes);

/* Measurement hooks (bench.py; no caller on the product path).
 * amc3d_fp32_probe launches `blocks` CTAs x 1024 threads of 8 independent FFMA chains of length `iters`
 * and stores the FLOP it executes in *flop (host); timed by the caller -> measured FP32-pipe peak.
 * `out` needs blocks*1024 floats (device; never written in practice).
 * amc3d_search_stats registers a device array of 8 u64 counters (NULL = off, the default).  While
 * registered, the culled searches launch counting instantiations of their kernels which add
 * [0] kNN (warp per query) distance evaluations, [1] its queries, [2] kNN (thread per query) evaluations,
 * [3] its queries, [4] ball-query evaluations, [5] its queries.  Results are unchanged. */
int amc3d_fp32_probe(int iters, int blocks, float *out, double *flop, void *stream);
int amc3d_search_stats(void *counters);

/* ---------------------------------------------------------------------------------------
 * pointnet2_batch family: batched (B,N,3) xyz and (B,C,N) features, all contiguous f32/i32
 * ------------------------------------------------------------------------------------- */

/* Iterative farthest point sampling, first pick = index 0; ties resolved exactly as the
 * reference's shared-memory tree (value desc, bit-reversed (k mod bs) asc, k asc).
 * xyz (B,N,3); temp (B,N) running min-distance, read on entry (caller fills 1e10) and
 * left holding the final distances; idx (B,m) i32.
 * ref: pointnet2_batch/src/sampling.cpp:39 furthest_point_sampling_wrapper,
 *      sampling_gpu.cu:218 furthest_point_sampling_kernel_launcher */
int amc3d_furthest_point_sampling(int b, int n, int m, const float *xyz, float *temp,
                                  int *idx, void *stream);

/* First `nsample` support indices (ascending index) with d2 < radius^2; unfilled slots
 * repeat the first hit; rows with no hit are left untouched (caller zero-fills idx).
 * new_xyz (B,M,3) queries, xyz (B,N,3) support, idx (B,M,nsample) i32.
 * ref: pointnet2_batch/src/ball_query.cpp:29 ball_query_wrapper_fast,
 *      ball_query_gpu.cu:54 ball_query_kernel_launcher_fast */
int amc3d_ball_query(int b, int n, int m, float radius, int nsample, const float *new_xyz,
                     const float *xyz, int *idx, void *stream);

/* out[b,c,j,s] = points[b,c,idx[b,j,s]].  points (B,C,N), idx (B,npoints,nsample),
 * out (B,C,npoints,nsample).
 * ref: group_points.cpp:25 group_points_wrapper_fast, group_points_gpu.cu:75 */
int amc3d_group_points(int b, int c, int n, int npoints, int nsample, const float *points,
                       const int *idx, float *out, void *stream);
/* Same, with a caller-provided workspace of b*n*c floats (device).  With a workspace the
 * gather runs through a channel-contiguous (B,N,C) copy of `points` (coalesced 128-byte
 * reads, HBM-write bound); without one (NULL) a direct kernel is used. */
int amc3d_group_points_ws(int b, int c, int n, int npoints, int nsample, const float *points,
                          const int *idx, float *out, float *workspace, void *stream);

/* grad_points[b,c,idx[b,j,s]] += grad_out[b,c,j,s]; grad_points (B,C,N) pre-zeroed by
 * the caller.  ref: group_points.cpp:13 group_points_grad_wrapper_fast,
 * group_points_gpu.cu:33 */
int amc3d_group_points_grad(int b, int c, int n, int npoints, int nsample,
                            const float *grad_out, const int *idx, float *grad_points,
                            void *stream);
/* Same, with a workspace of b*n*c floats: the scatter-add becomes warp-wide reductions onto
 * 128 contiguous bytes of an L2-resident (B,N,C) accumulator, then a transpose-accumulate. */
int amc3d_group_points_grad_ws(int b, int c, int n, int npoints, int nsample,
                               const float *grad_out, const int *idx, float *grad_points,
                               float *workspace, void *stream);

/* QueryAndGroup's relative coordinates in one launch:
 *   out[b,c,j,s] = (xyz[b,idx[b,j,s],c] - query[b,j,c]) * inv_radius ,  out (B,3,M,nsample)
 * xyz (B,N,3) and query (B,M,3) in their native layout (no transpose), inv_radius = 1.0f/(float)radius —
 * bit-identical to the reference's composition grouping_operation(xyz^T, idx) - query, / radius as torch
 * evaluates it on CUDA.  subtract == 0: no query term; inv_radius == 0: no scaling.
 * ref: models/layers/group.py:244-249 (QueryAndGroup.forward) */
int amc3d_group_xyz_relative(int b, int n, int m, int nsample, int subtract, float inv_radius,
                             const float *xyz, const float *query, const int *idx, float *out,
                             void *stream);

/* amc3d_group_points_grad_ws for a caller that does NOT need the accumulate semantics: grad_points is
 * written, not added to, so it need not be zero-filled (saves the caller's memset and one read of it). */
int amc3d_group_points_grad_ws_set(int b, int c, int n, int npoints, int nsample,
                                   const float *grad_out, const int *idx, float *grad_points,
                                   float *workspace, void *stream);

/* out[b,c,j] = points[b,c,idx[b,j]].  ref: sampling.cpp:16, sampling_gpu.cu:33 */
int amc3d_gather_points(int b, int c, int n, int npoints, const float *points,
                        const int *idx, float *out, void *stream);
/* grad_points[b,c,idx[b,j]] += grad_out[b,c,j].  ref: sampling.cpp:27, sampling_gpu.cu:72 */
int amc3d_gather_points_grad(int b, int c, int n, int npoints, const float *grad_out,
                             const int *idx, float *grad_points, void *stream);

/* 3 nearest `known` points of each `unknown` point, ordered by (d2, index) ascending.
 * unknown (B,n,3), known (B,m,3) -> dist2 (B,n,3) f32 (SQUARED; the Python wrapper takes
 * the sqrt), idx (B,n,3) i32.  Fewer than 3 known points: remaining slots (+inf, 0).
 * ref: interpolate.cpp:20 three_nn_wrapper_fast, interpolate_gpu.cu:62 */
int amc3d_three_nn(int b, int n, int m, const float *unknown, const float *known,
                   float *dist2, int *idx, void *stream);

/* out[b,c,i] = w0*points[b,c,i0] + w1*points[b,c,i1] + w2*points[b,c,i2] (FMA chain as
 * nvcc contracts the reference expression).  points (B,C,m), idx/weight (B,n,3).
 * ref: interpolate.cpp:31, interpolate_gpu.cu:107 */
int amc3d_three_interpolate(int b, int c, int m, int n, const float *points, const int *idx,
                            const float *weight, float *out, void *stream);
/* Same, with a caller-provided workspace of b*m*c floats: the three neighbour rows of a position
 * become three coalesced 128-byte reads of a channel-contiguous copy and the output leaves in
 * bulk-async 1 KB rows.  NULL selects the direct kernel. */
int amc3d_three_interpolate_ws(int b, int c, int m, int n, const float *points, const int *idx,
                               const float *weight, float *out, float *workspace, void *stream);
/* grad_points[b,c,idx[b,i,t]] += grad_out[b,c,i]*weight[b,i,t]; grad_points pre-zeroed.
 * ref: interpolate.cpp:45, interpolate_gpu.cu:152 */
int amc3d_three_interpolate_grad(int b, int c, int n, int m, const float *grad_out,
                                 const int *idx, const float *weight, float *grad_points,
                                 void *stream);
/* Same, with a workspace of b*m*c floats (vector reductions into a channel-contiguous accumulator). */
int amc3d_three_interpolate_grad_ws(int b, int c, int n, int m, const float *grad_out,
                                    const int *idx, const float *weight, float *grad_points,
                                    float *workspace, void *stream);

/* As amc3d_three_interpolate_grad_ws, but grad_points is written, not added to (no zero-fill needed). */
int amc3d_three_interpolate_grad_ws_set(int b, int c, int n, int m, const float *grad_out,
                                        const int *idx, const float *weight, float *grad_points,
                                        float *workspace, void *stream);

/* (B, rows, cols) -> (B, cols, rows), e.g. features (B,C,N) -> the channel-contiguous (B,N,C) copy the fused
 * operator below gathers from (torch's x.transpose(1, 2).contiguous(), as one tiled kernel). */
int amc3d_transpose_batched(int b, int rows, int cols, const float *src, float *dst, void *stream);

/* ---------------------------------------------------------------------------------------
 * Fused  grouping -> 1x1 conv -> BatchNorm (training statistics) -> ReLU -> max over nsample
 * (SURVEY.md §8f rank 1).  Replaces, for one conv layer, the module composition
 *   QueryAndGroup + cat(dp, fj) + Conv2d(3+C, O, 1, bias=False) + BatchNorm2d(O) + ReLU + max(dim=-1)
 * ref: models/backbone/pointnext_AA.py:57-63 LocalAggregation.forward, :139-170 SetAbstraction.forward,
 *      models/layers/group.py:235-255 QueryAndGroup.forward, models/layers/conv.py:24-61 create_convblock2d.
 * The grouped tensor (B, 3+C, M, nsample) is never materialised: tcgen05 MMAs (TF32 operands, FP32
 * accumulators in TMEM) read neighbour rows gathered straight into the swizzled operand tile.
 *   featT (B,N,C)   channel-contiguous features (amc3d_transpose_batched of the module's (B,C,N) input)
 *   xyz (B,N,3) support, new_xyz (B,M,3) queries, idx (B,M,nsample) i32 from amc3d_ball_query
 *   w_packed (O, C+8) = [W[:, 3:3+C] | W[:, 0:3] | 0 0 0 0 0]  (conv weight with the 3 dp columns moved
 *                       behind the features, rows padded to a multiple of 8 floats)
 *   w_tiles         scratch, ceil(O/128) * ceil((C+8)/32) * 4096 floats (x2 for precision 3): the weights as ready-made
 *                   swizzled operand tiles, written here and then fetched by one bulk copy per K chunk
 *   precision 1 = TF32 operands (cuDNN's default for the reference's Conv2d), 3 = 3xTF32 (FP32-faithful)
 * Outputs: out (B,O,M); mean/var (O) the batch statistics (var biased, as normalisation uses it);
 * invstd (O); and, kept for the backward: ysel (B*M,O) the pre-normalisation extreme of each (query,
 * channel), arg (B*M,O) u8 its sample index; sums (128*O) f64 scratch (64 replicas of sum y, sum y^2).
 * Limits: C % 8 == 0, nsample in {16, 32}. */
int amc3d_fused_sa_forward(int b, int n, int m, int c, int o, int nsample, float radius, int normalize_dp,
                           int precision, float eps, const float *featT, const float *xyz,
                           const float *new_xyz, const int *idx, const float *w_packed, float *w_tiles,
                           const float *gamma, const float *beta, float *ysel, unsigned char *arg,
                           double *sums, float *mean, float *var, float *invstd, float *out, void *stream);

/* Backward of amc3d_fused_sa_forward: two device steps around a few small library GEMMs on the host side
 * (amcontrast3d_b200/layers/_fused_backward.py has the algebra; csrc/fused_sa.cu the derivation).
 *  _backward_scatter  G' = grad_out * [out > 0]; dbeta_dgamma_wdp (5*O) f64 = [sum G' | sum G' yhat | dW of the
 *                     three relative-coordinate columns (O,3)]; and a_scatter (B*N rows of O floats, `lda`
 *                     floats apart): A[n,o] = sum over the queries whose arg-max sample of channel o is
 *                     support point n of gamma_o invstd_o G'[q,o].  Then dW_f = A^T f and df = A W_f are
 *                     (B*N) x O x C GEMMs, 1/nsample of the convolution's work.  Zeroes its outputs itself.
 *  _moments           per support point cnt (B*N rows, ld_cnt floats apart) and dpsum (B*N rows of 3 floats,
 *                     ld_dps apart) (how often / with which relative coordinates it is grouped) and mom (12) f64 =
 *                     [sum dp (3) | sum dp dp^T (9)]: with these the dense BatchNorm terms of the gradient
 *                     reduce to (B*N) x C x C GEMMs.
 * The row strides let the host keep A, cnt * f, dpsum and cnt as column blocks of ONE (B*N, O+C+4) matrix, so that
 * the whole feature gradient is one GEMM with it and all weight-gradient moments another. */
int amc3d_fused_sa_backward_scatter(int b, int n, int m, int o, int nsample, float radius, int normalize_dp,
                                    const float *grad_out, const float *out, const float *ysel,
                                    const unsigned char *arg, const int *idx, const float *xyz,
                                    const float *new_xyz, const float *mean, const float *invstd,
                                    const float *gamma, float *a_scatter, int lda, double *dbeta_dgamma_wdp,
                                    void *stream);
int amc3d_fused_sa_moments(int b, int n, int m, int nsample, float radius, int normalize_dp,
                           const float *xyz, const float *new_xyz, const int *idx, float *cnt, int ld_cnt,
                           float *dpsum, int ld_dps, double *mom, void *stream);

/* The small algebra between the two GEMMs of the backward, as two launches instead of ~30 torch ones
 * (Kq = C + 3, Kp = C + 8, Kc = O + C + 4; w_packed (O, Kp) as in the forward; small matrices padded to Kp with zeros
 * so that every library GEMM dimension stays a multiple of 8):
 *  _backward_coefs     c1wx (O, Kp) = [ c1 (.) W' | c0 | 0 ] with ghat = gamma invstd, c1 = ghat invstd dgamma / P,
 *                      c0 = ghat dbeta / P - c1 mean in FP64 (P = positions = B*M*nsample); dgamma, dbeta (O) f32.
 *  _backward_assemble  from g1 = f^T X (C, Kc), mom, qv = W'^T c1wx (Kp, Kp) = [Q | v | 0]:  sxx (Kp, Kp) the second
 *                      moments, wc (Kc, C) = [W_f; -Q_ff^T; -Q_fd^T; -v_f^T]  (df = X wc),
 *                      dwp (O, Kp) = [A^T f | wdp | 0] - c0 (x) S_x. */
int amc3d_fused_sa_backward_coefs(int c, int o, double positions, const float *gamma, const float *invstd,
                                  const float *mean, const double *dbeta_dgamma_wdp, const float *w_packed,
                                  float *c1wx, float *dgamma, float *dbeta, void *stream);
int amc3d_fused_sa_backward_assemble(int c, int o, const float *g1, const double *mom,
                                     const double *dbeta_dgamma_wdp, const float *w_packed, const float *qv,
                                     const float *c1wx, float *sxx, float *wc, float *dwp, void *stream);

/* Test hook (host code only, no launch): n / d evaluated with the multiply-high constants the fused forward uses for
 * its per-item index arithmetic; n < 2^31, d >= 1. */
unsigned int amc3d_debug_fastdiv(unsigned int n, unsigned int d);

/* ---------------------------------------------------------------------------------------
 * Input side: voxel hash and crop distances of the dataset code
 * ------------------------------------------------------------------------------------- */

/* keys[i] = FNV64-1A of floor(coord[i] / voxel_size) (FP64 division and floor, as numpy evaluates it);
 * cells (n,3) i64 receives the integer cell coordinates if not NULL (the 'ravel' hash is formed from them).
 * ref: openpoints/dataset/data_util.py:92-105 fnv_hash_vec, :125-131 voxelize */
int amc3d_voxel_keys(long long n, double voxel_size, const float *coord, unsigned long long *keys,
                     long long *cells, void *stream);
/* out[i] = sum((coord[i] - coord[init_idx])^2), FP32, the sort key of crop_pc.
 * ref: openpoints/dataset/data_util.py:158-160 crop_pc */
int amc3d_crop_dist2(long long n, const float *coord, long long init_idx, float *out, void *stream);

/* ---------------------------------------------------------------------------------------
 * pointops family: packed (n,3) xyz / (n,c) features with cumulative i32 `offset` ends
 * ------------------------------------------------------------------------------------- */

/* Exact k nearest neighbours inside each offset segment, ascending (d2, index).
 * xyz (n,3) support, new_xyz (m,3) queries, offset/new_offset (nseg) device i32 cumulative
 * ends, idx (m,nsample) i32, dist2 (m,nsample) f32 (SQUARED).  Segments shorter than
 * nsample pad with (segment start, 1e10) like the reference.  nsample <= 128 (the
 * reference's hard limit is 100, knnquery_cuda_kernel.cu:86-87).
 * `n` and `nseg` are the sizes of xyz and offset, which the reference launcher does not
 * take (it trusts the offsets); the Python shim passes xyz.shape[0] and offset.shape[0].
 * Contract for nseg == 1 (what AMContrast3D always passes): the one segment IS the whole
 * array, i.e. offset[0] == n and new_offset[0] == m; the culled search does not read the two
 * device arrays in that case (reading them would cost a host synchronisation).  A caller that
 * pads xyz beyond offset[0] must pass n = offset[0].  offset/new_offset must be int32.
 * ref: pointops/src/knnquery/knnquery_cuda.cpp:7 knnquery_cuda,
 *      knnquery_cuda_kernel.cu:111 knnquery_cuda_launcher */
int amc3d_knnquery(int n, int m, int nseg, int nsample, const float *xyz, const float *new_xyz,
                   const int *offset, const int *new_offset, int *idx, float *dist2,
                   void *stream);
/* Same, additionally writing `order` (m) i32 (NULL = not wanted): a permutation of the queries in
 * which consecutive entries are spatially close (the order the culled search visits them in;
 * identity for the brute-force paths).  Feeding it to amc3d_amloss_forward_order makes the
 * neighbour-row gathers of nearby anchors hit L1. */
int amc3d_knnquery_order(int n, int m, int nseg, int nsample, const float *xyz, const float *new_xyz,
                         const int *offset, const int *new_offset, int *idx, float *dist2,
                         int *order, void *stream);

/* out[i,s,:] = in[idx[i,s],:].  in (n,c), idx (m,nsample), out (m,nsample,c).
 * ref: pointops/src/grouping/grouping_cuda.cpp grouping_forward_cuda */
int amc3d_grouping_forward(int m, int nsample, int c, const float *input, const int *idx,
                           float *output, void *stream);
/* grad_in[idx[i,s],:] += grad_out[i,s,:]; grad_in (n,c) pre-zeroed.
 * ref: grouping_cuda.cpp grouping_backward_cuda */
int amc3d_grouping_backward(int m, int nsample, int c, const float *grad_output,
                            const int *idx, float *grad_input, void *stream);

/* The remaining pointops exports (no caller in AMContrast3D; they serve Point-Transformer style models
 * of OpenPoints and complete the `pointops_cuda` module surface, pointops_api.cpp:13-25). */

/* Ball query inside each offset segment: the first nsample support points (index order) with
 * d2 < radius*radius, global indices; slots past the hit count repeat the first hit; rows without a hit
 * are left as the caller filled them (zeros).  Sizes and offsets as for amc3d_knnquery.
 * ref: pointops/src/ballquery/ballquery_cuda.cpp:35, ballquery_cuda_kernel.cu:27-76 */
int amc3d_pointops_ballquery(int n, int m, int nseg, float radius, int nsample, const float *xyz,
                             const float *new_xyz, const int *offset, const int *new_offset, int *idx,
                             void *stream);
/* Furthest point sampling inside each segment: idx[new_offset[s-1] ..) = new_offset[s]-new_offset[s-1]
 * picks of segment s as GLOBAL indices, starting with the segment's first point.  h_offset and
 * h_new_offset are HOST copies of the cumulative ends (the reference's Python wrapper reads them on the
 * host too, pointops.py:20-23); n_max = the largest segment (it fixes the reference's block size and so
 * the order in which equal distances resolve); tmp (n) f32 = 1e10 on entry, clobbered.
 * ref: pointops/src/sampling/sampling_cuda.cpp furthestsampling_cuda, sampling_cuda_kernel.cu:13-133 */
int amc3d_pointops_furthestsampling(int nseg, int n_max, const float *xyz, const int *h_offset,
                                    const int *h_new_offset, float *tmp, int *idx, void *stream);
/* output[i,:] += sum_j input[idx[i,j],:] * weight[i,j]  (an FMA chain in j order from the caller's
 * output).  input (m,c), idx / weight (n,k), output (n,c).
 * ref: pointops/src/interpolation/interpolation_cuda_kernel.cu:5-18 / :20-32 */
int amc3d_pointops_interpolation_forward(int n, int c, int k, const float *input, const int *idx,
                                         const float *weight, float *output, void *stream);
/* grad_input[idx[i,j],:] += grad_output[i,:] * weight[i,j]; grad_input (m,c) pre-zeroed. */
int amc3d_pointops_interpolation_backward(int n, int c, int k, const float *grad_output, const int *idx,
                                          const float *weight, float *grad_input, void *stream);
/* output[i,s,:] = input1[i,:] - input2[idx[i,s],:].  input1, input2 (n,c), idx (n,nsample).
 * ref: pointops/src/subtraction/subtraction_cuda_kernel.cu:5-16 / :18-31 */
int amc3d_pointops_subtraction_forward(int n, int nsample, int c, const float *input1, const float *input2,
                                       const int *idx, float *output, void *stream);
/* grad_input1[i,:] += g, grad_input2[idx[i,s],:] -= g with g = grad_output[i,s,:]; both pre-zeroed. */
int amc3d_pointops_subtraction_backward(int n, int nsample, int c, const int *idx, const float *grad_output,
                                        float *grad_input1, float *grad_input2, void *stream);
/* output[i,ch] += sum_s (input[idx[i,s],ch] + position[i,s,ch]) * weight[i,s,ch % w_c]  (FMA chain in s
 * order).  input (n,c), position (n,nsample,c), weight (n,nsample,w_c), idx (n,nsample), output (n,c).
 * ref: pointops/src/aggregation/aggregation_cuda_kernel.cu:5-21 */
int amc3d_pointops_aggregation_forward(int n, int nsample, int c, int w_c, const float *input,
                                       const float *position, const float *weight, const int *idx,
                                       float *output, void *stream);
/* grad_input (n,c) and grad_weight (n,nsample,w_c) accumulate (pre-zeroed); grad_position (n,nsample,c)
 * is written.  ref: aggregation_cuda_kernel.cu:24-41 */
int amc3d_pointops_aggregation_backward(int n, int nsample, int c, int w_c, const float *input,
                                        const float *position, const float *weight, const int *idx,
                                        const float *grad_output, float *grad_input, float *grad_position,
                                        float *grad_weight, void *stream);

/* ---------------------------------------------------------------------------------------
 * adaptive-margin contrastive loss (one decoder stage)
 * ref: openpoints/AMContrast3D/MarginContrast.py:220-259, AEF/ambiguity.py:11-93,
 *      AEF/utils.py:11-43, AEF/function.py:10-39
 * ------------------------------------------------------------------------------------- */

/* Stage label by kNN vote: cls[i] = first argmax_c #{j<kr : t(nidx[i,j]) == c}, where
 * t(x) = ncls-1 if has_ignore && target[x]==ignore_index else target[x].  kr == 0 means
 * stage 0: cls[i] = t(i) (nidx unused).  target (M0) i64, nidx (m,kr) i32, cls (m) i32.
 * ref: AEF/utils.py:11-43 get_subscene_label_CBL + MarginContrast.py:112 argmax */
int amc3d_stage_labels(int m, int kr, int ncls, int has_ignore, long long ignore_index,
                       const long long *target, const int *nidx, int *cls, void *stream);

/* Per-class confusion counts of a prediction against integer labels: out (3*ncls) f32 = [tp | #predicted |
 * #labelled] (zeroed here), the quantities the trainer all-reduces every step (union = #predicted + #labelled - tp).
 * pred NULL: the prediction is the label of the point's first listed neighbour, target[nbr[i*ld]].
 * ref: utils/metrics.py ConfusionMatrix.update / tp / union / count; examples/segmentation/main_AA.py:461,496-507 */
int amc3d_class_counts(int m, int ncls, const int *target, const int *pred, const int *nbr, int ld,
                       float *out, void *stream);

/* Neighbour lists below are given as (nbr, ld, ke): row i holds its ke neighbour indices at
 * nbr[i*ld .. i*ld+ke).  For a kNN result knn_idx (m,k) whose column 0 is the self match
 * (what the reference drops with [..., 1:], MarginContrast.py:226) pass nbr = knn_idx + 1,
 * ld = k, ke = k-1 — no copy.  ke <= 32. */

/* posmask bits and positive count: posbits (m) u32 (bit j set iff cls[i]==cls[nbr[i,j]]),
 * cnt (m) i32, *max_cnt = max_i cnt[i] (device i32, zeroed by caller).
 * ref: MarginContrast.py:226-231, AEF/ambiguity.py:13-14 */
int amc3d_posmask_count(int m, int ke, int ld, const int *nbr, const int *cls, uint32_t *posbits,
                        int *cnt, int *max_cnt, void *stream);

/* Ambiguity a[i] (App. A.4 step 4).  cctype: 1 = Method1 (d+=d-=5), 2 = Method2 (sum of
 * the reference's square_distance), 3 = Method3 (sum of sqrt(|d2|+1e-12)).
 * stats (device i32[8], zeroed by caller): [0] #selected (0<a<=1), [1] #boundary,
 * [2..6] the reference's five ambiguity bins (a==0, low, semi, high, ceil(10a)==10) for
 * nu_m = nu*10.
 * ref: AEF/ambiguity.py:11-93, AEF/function.py:10-39 */
int amc3d_ambiguity(int m, int ke, int ld, const float *p, const int *nbr, const uint32_t *posbits,
                    const int *cnt, const int *max_cnt, int cctype, float beta, float nu,
                    float *a, int *stats, void *stream);
/* Same with the torch backend whose FP32 rounding of square_distance is reproduced: the reference's
 * |a|^2 + |b|^2 - 2ab is ill-conditioned for close points, and torch's CPU (sgemm, (x+y)+z) and CUDA
 * (cuBLAS fma(y,y',xx') + zz', (x+z)+y) kernels round it differently, so the reference's own `a` differs
 * between its CPU and CUDA runs (up to 3e-2 at BASELINE config 2).  torch_backend 1 = CUDA (what the
 * reference computes where it runs; amc3d_ambiguity uses it; cuBLAS switches kernels at 153 boundary
 * points, which is followed), 0 = CPU (what the golden vectors hold).  Two launches: a boundary count
 * into stats[1], then the ambiguity kernel. */
int amc3d_ambiguity_backend(int m, int ke, int ld, const float *p, const int *nbr,
                            const uint32_t *posbits, const int *cnt, const int *max_cnt, int cctype,
                            float beta, float nu, int torch_backend, float *a, int *stats, void *stream);

/* inv[i] = 1 / max(||f_i||_2, 1e-8).  f (m,d) row-major. */
int amc3d_row_inv_norm(int m, int d, const float *f, float *inv, void *stream);

typedef struct amc3d_loss_params {
    float temperature;   /* used when has_temperature != 0 */
    int has_temperature;
    int margin_mode;     /* 0 constant (nu), 1 adaptive (mu*a+nu) */
    float mu, nu;
    int db_mode;         /* 0 none, 1 '-m' (positives), 2 '+m' (negatives) */
    int cl_method;       /* 1 = supervisedCL Method1, 2 = Method2 */
} amc3d_loss_params;

/* Fused loss forward + gradient accumulation for one stage.
 * For every selected anchor (0<a<=1): cosine similarity to its k-1 neighbours, margin,
 * temperature, exp/sums/log; writes -log r_i to loss_pt[i] (0 for unselected points);
 * accumulates dL_i/du into ghat (m,d) (anchor role and neighbour role; ghat zeroed by the
 * caller) WITHOUT the 1/|sel| factor.  Nothing of size m*k*d is materialised.
 * f (m,d), inv (m), neighbour lists (nbr, ld, ke), posbits/a (m), loss_pt (m).
 * ref: MarginContrast.py:77-79 dist_cos, :117-174 contrast_softnn_margin, :250-257 */
int amc3d_amloss_forward(int m, int d, int ke, int ld, const float *f, const float *inv,
                         const int *nbr, const uint32_t *posbits, const float *a,
                         const amc3d_loss_params *params, float *loss_pt, float *ghat,
                         void *stream);

/* Same, visiting the anchors in the order given (order (m) i32, a permutation of 0..m-1; NULL =
 * index order).  Results are identical up to the summation order of the atomics. */
int amc3d_amloss_forward_order(int m, int d, int ke, int ld, const float *f, const float *inv,
                               const int *nbr, const uint32_t *posbits, const float *a,
                               const amc3d_loss_params *params, float *loss_pt, float *ghat,
                               const int *order, void *stream);

/* loss_out[0] = sum_i loss_pt[i] / stats[0]  (deterministic, double accumulation; the
 * torch.mean over the selected points, MarginContrast.py:257; 0 selected -> NaN) */
int amc3d_amloss_reduce(int m, const float *loss_pt, const int *stats, float *loss_out,
                        void *stream);

/* grad_f[r] = scale * (ghat[r] - u_r (u_r . ghat[r])) * inv[r] (norm >= 1e-8), where
 * scale = upstream[0] / stats[0]  (upstream: device f32 scalar dTotal/dL_s; stats[0] =
 * #selected from amc3d_ambiguity).  Accumulate=0 overwrites grad_f, 1 adds into it. */
int amc3d_amloss_backward(int m, int d, const float *f, const float *inv, const float *ghat,
                          const float *upstream, const int *stats, int accumulate,
                          float *grad_f, void *stream);

/* ---------------------------------------------------------------------------------------
 * AMContrast3D++ masked refinement (RefinementMethod.DualMasks, fusion 'MIN')
 * ref: openpoints/AMContrast3D/MaskedRefine.py:49-119
 * ------------------------------------------------------------------------------------- */

/* jmin[r] = nbr[r, argmin_{j<ke} a[nbr[r,j]]] (first minimum); a (m). */
int amc3d_refine_select(int m, int ke, int ld, const int *nbr, const float *a, int *jmin,
                        void *stream);
/* out = gamma*(f*~mask + cross*mask) + (1-gamma)*f over the (B,D,n) buffer, where
 * cross is the flat D-float chunk jmin[r] of f re-viewed as (B,D,n) and mask[b,0,i] =
 * thr <= a[b,i] <= thr_max.  update_count (device i32, zeroed by caller) += #mask. */
int amc3d_refine_forward(int b, int d, int n, const float *f, const float *a, const int *jmin,
                         float thr, float thr_max, float gamma, float *out,
                         int *update_count, void *stream);
/* grad_f for amc3d_refine_forward; grad_f (B,D,n) written (not accumulated) */
int amc3d_refine_backward(int b, int d, int n, const float *grad_out, const float *a,
                          const int *jmin, float thr, float thr_max, float gamma,
                          float *grad_f, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* AMC3D_H */
