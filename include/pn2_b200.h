/*
 * pn2_b200.h -- C ABI of libpn2_b200.so: the PointNet++ set-abstraction /
 * feature-propagation op path of ero1311/Pointnet2.ScanNet as hand-written
 * sm_100a CUDA.
 *
 * Boundary.  The reference's host bindings (pointnet2/_ext_src/src/*.cpp) call
 * one plain-C++ launcher per kernel, `*_kernel_wrapper(int..., const float*,
 * int*)`, declared at the top of each binding file.  Every entry point below
 * replaces exactly one of those launchers: same scalar arguments in the same
 * order, same pointer meaning, plus an explicit stream (the reference takes
 * at::cuda::getCurrentCUDAStream() inside the launcher) and an int status
 * instead of the reference's fprintf + exit(-1) (include/cuda_utils.h:30-39).
 *
 * Conventions
 *  - all pointers are DEVICE pointers to contiguous fp32 / int32 arrays on the
 *    current device; inputs are borrowed, outputs are caller-allocated.  The
 *    library never allocates, frees or synchronises.
 *  - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *  - return 0 on success; PN2_ERR_* (<0) for rejected arguments; a positive
 *    value is the cudaError_t of a failed launch.  Nothing is written on error.
 *  - outputs are fully written by the kernels: they need NOT be zero-filled by
 *    the caller (the reference's bindings zero-fill because its kernels
 *    accumulate with atomicAdd / leave empty balls untouched).  The results are
 *    the same as the reference's on its zero-filled outputs.
 *  - functions are re-entrant and thread-safe; no global state except the
 *    launch counter.
 *  - arithmetic: fp32, every distance evaluated as ((dx*dx)+(dy*dy))+(dz*dz)
 *    with each operation rounded (no FMA contraction), so sampled / grouped /
 *    neighbour indices are bit-identical to the reference built with
 *    -fmad=false.
 */
#ifndef PN2_B200_H
#define PN2_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define PN2_ABI_VERSION 1

#if defined(__GNUC__)
#define PN2_API __attribute__((visibility("default")))
#else
#define PN2_API
#endif

#define PN2_OK 0
#define PN2_ERR_INVALID_ARG (-1) /* null pointer, negative size */
#define PN2_ERR_UNSUPPORTED (-2) /* size outside what the kernels handle */
#define PN2_ERR_ALIGNMENT (-3)   /* pointer not 4-byte aligned */

/* -- library ------------------------------------------------------------- */
PN2_API int pn2_abi_version(void);
/* kernels launched by this library since load / last reset (all threads) */
PN2_API unsigned long long pn2_launch_count(void);
PN2_API void pn2_reset_launch_count(void);
/* static string for a status returned by any function below */
PN2_API const char *pn2_status_string(int status);

/* -- furthest point sampling ---------------------------------------------
 * replaces furthest_point_sampling_kernel_wrapper (sampling.cpp:11-13,
 * sampling_gpu.cu:175-229).
 *   dataset (b, n, 3) f32   ->   idxs (b, m) i32
 * idxs[.,0] = 0; then m-1 times: every point with |p|^2 > 1e-3 (compared in
 * double, sampling_gpu.cu:100-101) lowers its running minimum squared distance
 * to the last pick, and the next pick is the arg-max of those minima.  Ties
 * are resolved as the reference's strided scan + shared-memory tree resolves
 * them: minimise (bitrev_{log2 bs}(k mod bs), k div bs), bs = min(2^floor(log2
 * n), 512).  If no point qualifies the pick is 0.
 *   temp (b, n) f32: scratch.  The reference requires it pre-filled with 1e10
 * (sampling.cpp:74-76) and streams it through global memory every iteration.
 * Here the running minima live in registers; temp may be NULL and is left
 * untouched when n <= PN2_FPS_MAX_RESIDENT_N.  For larger n it is REQUIRED
 * (any contents; the kernel initialises it) and is used as the reference uses
 * it.
 */
#define PN2_FPS_MAX_RESIDENT_N 8192
PN2_API int pn2_furthest_point_sampling(int b, int n, int m, const float *dataset,
                                float *temp, int *idxs, void *stream);
/* Identical results through the variant that Morton-sorts the cloud inside the kernel and lets a
 * warp skip an iteration when the exactly-rounded lower bound of the distance from the new pick
 * to its bounding box cannot lower any of its minima (applies to 2048 < n <= 8192; otherwise the
 * same as above).  Not the default: see DESIGN.md section 4. */
PN2_API int pn2_furthest_point_sampling_pruned(int b, int n, int m, const float *dataset,
                                       float *temp, int *idxs, void *stream);

/* -- gather ---------------------------------------------------------------
 * replace gather_points(_grad)_kernel_wrapper (sampling.cpp:4-9,
 * sampling_gpu.cu:22-30, 49-57).
 *   fwd : points (b, c, n), idx (b, npoints)       -> out (b, c, npoints)
 *   grad: grad_out (b, c, npoints), idx            -> grad_points (b, c, n)
 *         (scatter-add; untouched entries are written as 0)
 */
PN2_API int pn2_gather_points(int b, int c, int n, int npoints, const float *points,
                      const int *idx, float *out, void *stream);
PN2_API int pn2_gather_points_grad(int b, int c, int n, int npoints,
                           const float *grad_out, const int *idx,
                           float *grad_points, void *stream);

/* -- ball query -------------------------------------------------------------
 * replaces query_ball_point_kernel_wrapper (ball_query.cpp:4-6,
 * ball_query_gpu.cu:46-54).
 *   new_xyz (b, m, 3) centres, xyz (b, n, 3) points -> idx (b, m, nsample)
 * For each centre: the first `nsample` point indices k, ascending, with
 * squared distance < radius*radius (fp32 product, strict); unused slots repeat
 * the first hit; a centre without any hit gets all zeros.
 */
PN2_API int pn2_query_ball_point(int b, int n, int m, float radius, int nsample,
                         const float *new_xyz, const float *xyz, int *idx,
                         void *stream);

/* -- grouping ---------------------------------------------------------------
 * replace group_points(_grad)_kernel_wrapper (group_points.cpp:4-10,
 * group_points_gpu.cu:30-39, 66-75).
 *   fwd : points (b, c, n), idx (b, npoints, nsample)
 *                                   -> out (b, c, npoints, nsample)
 *   grad: grad_out (b, c, npoints, nsample), idx -> grad_points (b, c, n)
 */
PN2_API int pn2_group_points(int b, int c, int n, int npoints, int nsample,
                     const float *points, const int *idx, float *out,
                     void *stream);
PN2_API int pn2_group_points_grad(int b, int c, int n, int npoints, int nsample,
                          const float *grad_out, const int *idx,
                          float *grad_points, void *stream);

/* -- three nearest neighbours ----------------------------------------------
 * replaces three_nn_kernel_wrapper (interpolate.cpp:4-5,
 * interpolate_gpu.cu:61-68).
 *   unknown (b, n, 3), known (b, m, 3) -> dist2 (b, n, 3) f32, idx (b, n, 3)
 * SQUARED distances, ascending; scan order k ascending with strict '<' (the
 * lower index keeps the better slot on ties); with fewer than 3 known points
 * the missing slots are (+inf, 0).
 */
PN2_API int pn2_three_nn(int b, int n, int m, const float *unknown, const float *known,
                 float *dist2, int *idx, void *stream);

/* -- three-point interpolation ----------------------------------------------
 * replace three_interpolate(_grad)_kernel_wrapper (interpolate.cpp:6-12,
 * interpolate_gpu.cu:103-111, 145-154).
 *   fwd : points (b, c, m), idx (b, n, 3), weight (b, n, 3) -> out (b, c, n)
 *         out = ((p[i1]*w1) + (p[i2]*w2)) + (p[i3]*w3)
 *   grad: grad_out (b, c, n), idx, weight -> grad_points (b, c, m)
 */
PN2_API int pn2_three_interpolate(int b, int c, int m, int n, const float *points,
                          const int *idx, const float *weight, float *out,
                          void *stream);
PN2_API int pn2_three_interpolate_grad(int b, int c, int n, int m,
                               const float *grad_out, const int *idx,
                               const float *weight, float *grad_points,
                               void *stream);

/* ==========================================================================
 * Shared-MLP block in position-major ("rows") layout.
 *
 * Replaces what PointnetSAModule(MSG) / PointnetFPModule run around their 1x1
 * convolutions through torch.nn / cuDNN on (B, C, npoint, nsample) tensors
 * (pointnet2_modules.py:60-75, 395-416; pytorch_utils.py:11-36, 87-120):
 * QueryAndGroup's gather + centring + concat, train-mode BatchNorm2d, ReLU and
 * max_pool2d over nsample, forward and backward.  An activation is a row-major
 * [P, C] fp32 matrix (P = b*npoint*nsample grouped positions, or b*n points;
 * C % 4 == 0, C <= 1024, 16-byte aligned), so a 1x1 conv is Z = H * W^T.
 * BatchNorm is carried as the per-channel fold  y = a*z + b  plus the saved
 * (mean, invstd) of the batch.  All reductions are two-stage and
 * order-deterministic.  `workspace`: pn2_bn_stats_workspace_floats(c) + 2*c
 * floats, 16-byte aligned, private to the call while it runs.
 * ========================================================================== */
PN2_API int pn2_bn_stats_workspace_floats(int c);

/* batch statistics of z [p, c]: biased variance; writes the fold (a, b), the saved
 * (mean, invstd), and -- unless running_mean is NULL -- updates the running
 * statistics like torch.nn.BatchNorm2d (momentum, unbiased variance);
 * num_batches_tracked (int64, may be NULL) is incremented by one. */
PN2_API int pn2_bn_stats(int p, int c, const float *z, const float *gamma, const float *beta,
                         float eps, float momentum, float *running_mean, float *running_var,
                         long long *num_batches_tracked, float *a, float *b, float *mean,
                         float *invstd, float *workspace, void *stream);
/* eval mode: the fold from the running statistics */
PN2_API int pn2_bn_fold_eval(int c, const float *gamma, const float *beta,
                             const float *running_mean, const float *running_var, float eps,
                             float *a, float *b, float *mean, float *invstd, void *stream);
/* h = relu ? max(a*z + b, 0) : a*z + b            (h may alias z) */
PN2_API int pn2_bn_relu_apply(int p, int c, const float *z, const float *a, const float *b,
                              float *h, int relu, void *stream);
/* out[g, c] = max over the ns rows of group g of relu(a*z + b); argmax[g, c] = first
 * maximal row (F.max_pool2d's choice)              z [g*ns, c] -> out, argmax [g, c] */
PN2_API int pn2_bn_relu_groupmax(int g, int ns, int c, const float *z, const float *a,
                                 const float *b, float *out, int *argmax, void *stream);
/* adjoint of pn2_bn_relu_groupmax through BatchNorm (training != 0: batch statistics):
 * dout [g, c] -> dense dz [g*ns, c], dgamma [c], dbeta [c] */
PN2_API int pn2_pool_bn_bwd(int g, int ns, int c, int training, const float *dout,
                            const int *argmax, const float *z, const float *a, const float *b,
                            const float *mean, const float *invstd, float *dz, float *dgamma,
                            float *dbeta, float *workspace, void *stream);
/* adjoint of pn2_bn_relu_apply through BatchNorm: dh [p, c] -> dz [p, c] (may alias dh) */
PN2_API int pn2_relu_bn_bwd(int p, int c, int training, int relu, const float *dh, const float *z,
                            const float *a, const float *b, const float *mean,
                            const float *invstd, float *dz, float *dgamma, float *dbeta,
                            float *workspace, void *stream);
/* QueryAndGroup in rows layout (pointnet2_utils.py:317-359): table (b, n, ct) holds
 * [xyz | features] per point (ct % 4 == 0, zero padded); rows[(b, j, s), :] =
 * table[b, idx[b, j, s], :] with the first 3 columns minus new_xyz[b, j].
 * _grad: scatter-add of grad_rows into grad_table (fully written). */
PN2_API int pn2_sa_group_rows(int b, int n, int npoint, int nsample, int ct, const float *table,
                              const float *new_xyz, const int *idx, float *rows, void *stream);
PN2_API int pn2_sa_group_rows_grad(int b, int n, int npoint, int nsample, int ct,
                                   const float *grad_rows, const int *idx, float *grad_table,
                                   void *stream);

/* PointnetFPModule's three_interpolate + concat in rows layout
 * (pointnet2_modules.py:395-412): known (b, m, c2) and skip (b, n, c1) point-major,
 * c2 % 4 == 0; rows[(b, j), :] = [ ((k1*w1)+(k2*w2))+(k3*w3) | skip[b, j, :] | 0 pad ],
 * row width pad4(c2 + c1).  _grad: grad_known (fully written, scatter-add) and
 * grad_skip (may be NULL). */
PN2_API int pn2_fp_interp_rows(int b, int n, int m, int c2, int c1, const float *known,
                               const float *skip, const int *idx, const float *weight,
                               float *rows, void *stream);
PN2_API int pn2_fp_interp_rows_grad(int b, int n, int m, int c2, int c1, const float *grad_rows,
                                    const int *idx, const float *weight, float *grad_known,
                                    float *grad_skip, void *stream);

/* Adjoints of the two row gathers as segmented gather-sums over an INVERSE MAP (csrc/inverse_map.cu): no
 * atomics, no zero fill, fixed summation order (bit-reproducible) -- replacing the atomicAdd scatters of
 * group_points_grad / three_interpolate_grad (group_points_gpu.cu:43-75, interpolate_gpu.cu:116-154).
 * pn2_inverse_map: keys [b, n_items] with values in [0, n_src) (the ball-query idx flattened to
 * [b, npoint * nsample], or the 3-NN idx flattened to [b, 3 * n]) -> offsets [b, n_src + 1] and items
 * [b, n_items]: segment k of cloud i = items[i, offsets[i, k] .. offsets[i, k + 1]) = the item ids whose key
 * is k, ascending (out-of-range keys are dropped; unused tail = -1).  The map depends on coordinates only
 * and is built once per batch beside the ball query / 3-NN.  Supported sizes: pn2_inverse_map_supported. */
PN2_API int pn2_inverse_map_supported(int n_src, int n_items);
PN2_API int pn2_inverse_map(int b, int n_src, int n_items, const int *keys, int *offsets, int *items,
                            void *stream);
PN2_API int pn2_sa_group_rows_grad_inv(int b, int n, int npoint, int nsample, int ct,
                                       const float *grad_rows, const int *offsets, const int *items,
                                       float *grad_table, void *stream);
PN2_API int pn2_fp_interp_rows_grad_inv(int b, int n, int m, int c2, int c1, const float *grad_rows,
                                        const int *offsets, const int *items, const float *weight,
                                        float *grad_known, float *grad_skip, void *stream);

/* ==========================================================================
 * Shared-MLP layer on the tensor cores (tcgen05 / TMEM, TF32 operands, fp32
 * accumulation):  z[p, n] = act(a)[p, k] * w[n, k]^T  with
 *   - act(x) = relu_in ? max(a_in*x + b_in, 0) : a_in*x + b_in  (the previous
 *     layer's BatchNorm fold, applied to the A tile in shared memory; a_in NULL =
 *     identity); h_out (optional, [p, lda]) receives act(a) once, in fp32;
 *   - partials (optional): per-CTA column sums / sums of squares of z,
 *     [pn2_mlp_gemm_grid(p, n)][2][n_tiles * bn] floats, to be finalised with
 *     pn2_bn_finalize (ld = n_tiles * bn).
 * bn = pn2_mlp_gemm_bn(p, n) is the N tile per CTA (narrower when p is small, so that the grid
 * still covers the machine).
 * w_packed comes from pn2_mlp_pack_weights (n_tiles*k_blocks*bn*32 floats: tile
 * images in the UMMA 128-byte-swizzled K-major layout).  a: [p, lda], lda % 4 == 0,
 * columns [k, lda) are read as part of the product (keep them zero); z: [p, ldz].
 * This replaces the cuDNN/cuBLAS 1x1 convolution the reference's SharedMLP calls
 * (pytorch_utils.py:87-95) plus the statistics pass of its BatchNorm2d (:100-104).
 * ========================================================================== */
PN2_API int pn2_mlp_gemm_bn(int p, int n);
PN2_API int pn2_mlp_gemm_n_tiles(int p, int n);
PN2_API int pn2_mlp_gemm_k_blocks(int k);
PN2_API int pn2_mlp_gemm_grid(int p, int n);
PN2_API int pn2_mlp_pack_weights(int p, int n, int k, int ldw, const float *w, float *packed,
                                 void *stream);
PN2_API int pn2_mlp_gemm_pool_supported(int ns);
PN2_API int pn2_mlp_gemm_fwd(int p, int k, int lda, int n, int ldz, const float *a,
                             const float *a_in, const float *b_in, int relu_in, float *h_out,
                             const float *w_packed, float *z, float *partials,
                             const unsigned char *drop, float drop_scale, int pool_ns, float *zmax,
                             float *zmin, int *imax, int *imin, void *stream);
/* pool_ns (0: none; else pn2_mlp_gemm_pool_supported(pool_ns), p % pool_ns == 0): the layer is the pooled
 * last layer of an SA block (max_pool2d over nsample, pointnet2_modules.py:68-70): next to z the kernel
 * leaves, per group of pool_ns rows and channel, the max and the min of z and the first row attaining
 * each (zmax / zmin / imax / imin, [p / pool_ns, n]), read from its staged output blocks -- z is not
 * re-read.  pn2_groupmax_from_extrema turns them into max_s relu(a*z+b) and its arg-max row once the
 * BatchNorm fold (a, b) is known: the sign of a picks max or min.  zsel (optional, [g, c]) receives the
 * picked extremum = z at the arg-max row; pn2_pool_bn_bwd_stats reads it instead of gathering over z. */
PN2_API int pn2_groupmax_from_extrema(int g, int c, const float *zmax, const float *zmin, const int *imax,
                                      const int *imin, const float *a, const float *b, float *out,
                                      int *argmax, float *zsel, void *stream);
/* drop (optional, [p, k] bytes; needs a_in, excludes h_out): keep mask of an nn.Dropout sitting between
 * the previous layer's activation and this layer (pointnet2_semseg.py:156): act(a) is multiplied by
 * drop ? drop_scale : 0 on the tile, so the dropped activation is never materialised. */
PN2_API int pn2_bn_finalize(int p, int c, int ld, int nblk, const float *partials,
                            const float *gamma, const float *beta, float eps, float momentum,
                            float *running_mean, float *running_var, long long *num_batches_tracked,
                            float *a, float *b, float *mean, float *invstd, void *stream);

/* ==========================================================================
 * Whole set-abstraction block in one kernel, inference (csrc/sa_chain.cu): QueryAndGroup ->
 * three Conv2d(1x1) + folded BatchNorm + ReLU -> max over nsample
 * (PointnetSAModule(MSG).forward in eval mode, pointnet2_modules.py:60-75).  table [b, n, ct] =
 * [xyz | features | 0 pad] point rows, new_xyz [b, npoint, 3], idx [b, npoint, ns]; w1p / w2p / w3p:
 * the layers' weights [c_l, c_{l-1}] packed by pn2_mlp_pack_weights_bn (pn2_sa_chain_weight_floats(n, k)
 * floats each; layer 1's k = ct); fold = [a1 | b1 | a2 | b2 | a3 | b3] (pn2_bn_fold_eval); out
 * [b * npoint, c3].  Activations never leave the SM.  Supported shapes: pn2_sa_chain_supported.
 * ========================================================================== */
PN2_API int pn2_sa_chain_supported(int ns, int ct, int c1, int c2, int c3);
PN2_API int pn2_sa_chain_weight_floats(int n, int k);
PN2_API int pn2_mlp_pack_weights_bn(int n, int k, int ldw, const float *w, float *packed, void *stream);
PN2_API int pn2_sa_chain_eval(int b, int n, int npoint, int ns, int ct, const float *table,
                              const float *new_xyz, const int *idx, int c1, int c2, int c3,
                              const float *w1p, const float *w2p, const float *w3p, const float *fold,
                              float *out, void *stream);

/* ==========================================================================
 * Fused backward of one shared-MLP layer on the tensor cores (csrc/mlp_bwd.cu):
 * the adjoint of Conv2d(1x1) -> BatchNorm2d(train) -> ReLU (pytorch_utils.py:11-36)
 * in ONE pass over the rows, replacing autograd's threshold_backward + cudnn
 * batch-norm backward + two cuBLAS GEMMs:
 *     dZ      = a * (dY - c1 - zhat * c2)   dY = relu-masked dh, or (dh == NULL) the
 *                                           arg-max scatter of dout over groups of ns rows
 *     dh_prev = dZ * W                      [p, lda]   (wt_packed / dh_prev NULL: skipped)
 *     dW      = dZ^T * H_prev               per-CTA partials dw_part [pn2_mlp_bwd_grid(p)][nb*32][kb*32],
 *                                           nb = pn2_mlp_bwd_nb(n), kb = pn2_mlp_bwd_kb(lda)
 *                                           (stats_part NULL; sum them with pn2_sum_partials)
 *     stats   (stats_part != NULL: there is a BatchNorm below)  instead of dW the kernel leaves the two
 *               tensor-core products with the ReLU mask M of the layer below,
 *               S = dZ^T M in dw_part and Q = dZ^T (M .* zhat_prev) in stats_part (same shape):
 *               dW = g1 .* Q + g0 .* S and the column sums of dY_prev, dY_prev * zhat_prev
 *               (= sum_n w .* S, sum_n w .* Q) are linear in them, so pn2_mlp_bwd_finalize forms
 *               them once from the sums over all CTAs (w [n, k] fp32 is not read by the layer kernel)
 * with H_prev = relu_in ? max(a_in * a_prev + b_in, 0) : ... recomputed from the previous
 * layer's pre-BN output a_prev [p, lda] (a_in NULL: a_prev is used as it is, e.g. the
 * block's grouped input rows).  z [p, n], a/b/mean/invstd [n] are this layer's saved
 * forward state, c1c2 = [dbeta/P | dgamma/P] from pn2_*_bwd_stats or
 * pn2_bwd_stats_finalize.  Supported: n <= 128, n % 4 == 0, lda <= 160 (128 with
 * statistics), lda % 4 == 0 (pn2_mlp_bwd_supported); wt_packed from pn2_mlp_pack_weights_t (nb * kpad * 32 floats,
 * kpad = pn2_mlp_bwd_kpad(lda)).  TF32 operands, fp32 accumulation.
 * sched: two ints, zero before the first launch (the kernel re-arms them): the tile counter of the
 * dynamic scheduler -- row tiles are handed out by an atomic counter so that the kernel keeps its
 * speed when other streams hold part of the SMs.  One buffer per stream that launches the kernel.
 * ========================================================================== */
PN2_API int pn2_mlp_bwd_supported(int p, int n, int k, int lda, int with_stats);
PN2_API int pn2_mlp_bwd_grid(int p);
PN2_API int pn2_mlp_bwd_nb(int n);
PN2_API int pn2_mlp_bwd_kb(int lda);
PN2_API int pn2_mlp_bwd_kpad(int lda);
PN2_API int pn2_mlp_pack_weights_t(int n, int k, int lda, int ldw, const float *w, float *packed,
                                   void *stream);
PN2_API int pn2_mlp_bwd_layer(int p, int n, int k, int lda, int ns, int relu, int relu_in,
                              const float *dh, const float *dout, const int *argmax, const float *z,
                              const float *a, const float *b, const float *mean, const float *invstd,
                              const float *c1c2, const float *w, const float *a_prev, const float *a_in,
                              const float *b_in, const float *mean_in, const float *invstd_in,
                              const float *wt_packed, float *dh_prev, float *dw_part,
                              float *stats_part, int *sched, const unsigned char *drop_out,
                              float drop_out_scale, const unsigned char *drop_in, float drop_in_scale,
                              void *stream);
/* after pn2_mlp_bwd_layer with stats_part: sums the `parts` = pn2_mlp_bwd_grid(p) images (fixed order) ->
 * dw [nb*32][kb*32] (rows >= n / columns >= k are padding), dgamma / dbeta [k] of the layer below and
 * c1c2 = [dbeta / p | dgamma / p] ([k | k]; zeros in eval mode) for that layer's own pn2_mlp_bwd_layer. */
PN2_API int pn2_mlp_bwd_finalize_scratch_floats(int n, int lda);
PN2_API int pn2_mlp_bwd_finalize(int parts, int n, int k, int lda, const float *s_part, const float *q_part,
                                 const float *w, const float *a_in, const float *b_in, const float *mean_in,
                                 const float *invstd_in, int p, int training, float *dw, float *dgamma,
                                 float *dbeta, float *c1c2, float *scratch, unsigned *counter, void *stream);
/* scratch: pn2_mlp_bwd_finalize_scratch_floats(n, lda) floats; counter: ONE zero-initialised unsigned per
 * launching stream, re-armed by the kernel (its last CTA adds the row groups' column sums, fixed order). */
/* drop_out [p, n] / drop_in [p, k] (optional byte keep masks): dropout applied to this layer's
 * activated output (its adjoint multiplies dh first) / to the activation of the layer below (H_prev
 * and the ReLU mask of the statistics products carry it). */
PN2_API int pn2_pool_bn_bwd_stats(int g, int ns, int c, int training, const float *dout,
                                  const int *argmax, const float *z, const float *zsel /* optional */,
                                  const float *a, const float *b,
                                  const float *mean, const float *invstd, float *dgamma,
                                  float *dbeta, float *c1c2, float *workspace, void *stream);
PN2_API int pn2_relu_bn_bwd_stats(int p, int c, int training, int relu, const float *dh,
                                  const float *z, const float *a, const float *b, const float *mean,
                                  const float *invstd, float *dgamma, float *dbeta, float *c1c2,
                                  float *workspace, void *stream);
PN2_API int pn2_bwd_stats_finalize(int p, int c, int ld, int nblk, int training,
                                   const float *partials, float *dgamma, float *dbeta, float *c1c2,
                                   void *stream);

/* ==========================================================================
 * Wide layers (n > 128 or lda > 160: SA3 / SA4 / FP1-3 of the reference's models,
 * pointnet2_semseg.py:32-34, 112-114).  Their adjoint is split in three: dZ written once by
 * pn2_relu_bn_bwd / pn2_pool_bn_bwd, then two tensor-core products (replacing the two cuBLAS
 * GEMMs autograd runs for Conv2d 1x1, pytorch_utils.py:87-95):
 *   dh_prev[p, n_out] = dZ[p, n] * W[n, k]   pn2_mlp_gemm_fwd with a = dZ and the weight tiles of
 *                                            pn2_mlp_pack_weights_dgrad (columns k..n_out-1 = zeros)
 *   dW[n, k]          = dZ^T * H_prev        pn2_mlp_wgrad (csrc/mlp_wgrad.cu): both operands read
 *                                            MN-major straight from their row-major tensors by TMA box
 *                                            loads; row-split partial images
 *                                            part[pn2_mlp_wgrad_splits(p, n, ldy)][pn2_mlp_wgrad_rows(n)]
 *                                                [pn2_mlp_wgrad_cols(ldy)], summed by pn2_sum_partials;
 *                                            dW is the leading [n, k] block.  x: [p, ldx] (n columns
 *                                            used), y: [p, ldy]; n, ldx, ldy multiples of 4.
 * ========================================================================== */
PN2_API int pn2_mlp_pack_weights_dgrad(int p, int n_out, int n, int k, int ldw, const float *w,
                                       float *packed, void *stream);
/* the data-gradient product with the BatchNorm-adjoint statistics of the layer BELOW from its epilogue (z_prev
 * [p, n_out], fold_prev = [a | b | mean | invstd] of that layer, [4][n_out]; partials as pn2_mlp_gemm_fwd's, finalised
 * by pn2_bwd_stats_finalize), and the dZ half of pn2_relu_bn_bwd for a layer whose c1c2 came that way */
PN2_API int pn2_mlp_dgrad_stats(int p, int n, int n_out, const float *dz, const float *w_packed,
                                float *dh_prev, const float *z_prev, const float *fold_prev, int relu_prev,
                                float *partials, void *stream);
PN2_API int pn2_relu_bn_bwd_dz(int p, int c, int relu, const float *dh, const float *z, const float *a,
                               const float *b, const float *mean, const float *invstd, const float *c1c2,
                               float *dz, void *stream);
PN2_API int pn2_mlp_wgrad_splits(int p, int n, int ldy);
PN2_API int pn2_mlp_wgrad_rows(int n);
PN2_API int pn2_mlp_wgrad_cols(int ldy);
PN2_API int pn2_mlp_wgrad(int p, int n, int ldx, int ldy, const float *x, const float *y, float *part,
                          void *stream);

/* ==========================================================================
 * First layer of a shared-MLP block applied BEFORE the gather (csrc/mlp_lead.cu; training mode).
 * The first Conv2d 1x1 of PointnetSAModule(MSG) / PointnetFPModule acts on gathered rows
 * (pointnet2_modules.py:60-75: QueryAndGroup -> mlp; :395-412: three_interpolate -> cat -> mlp); being linear
 * it commutes with the gather:
 *   SA: z[g, s, :] = u[idx[g, s], :] + W_x (xyz[idx[g, s]] - new_xyz[g]),  u = features W_f^T over the n points
 *   FP: z[r, :]    = sum_t weight[r, t] gk[idx[r, t], :] + W_s skip[r, :], gk = known W_k^T over the m points
 * so the gathered rows (b*npoint*nsample x (3 + C), b*n x (C2 + C1)) are never materialised.  u / gk come from
 * pn2_mlp_gemm_fwd over the source points (weights packed with pn2_mlp_pack_weights_cols: the first k0
 * reduction columns -- the coordinates of a [xyz | features] table -- read as zeros).
 *
 * c = width of the layer (a power of two in [16, 1024]; FP: [64, 1024]); z [P, c] rows; partials [pn2_lead_blocks(P, c, kind)][2][c] = per-CTA column
 * sums of z and z^2 over fixed row ranges (finalise with pn2_bn_finalize, ld = c).  wx = &W[0][0] with row
 * stride ldw (columns 0..2 are the coordinate weights); ws = &W[0][c2] (the c1 <= pn2_fp_lead_max_skip()
 * skip columns).  xyz: point p's coordinates at xyz[p * ldx .. + 3) (ldx = 3 for a (b, n, 3) tensor, the row width
 * when they are read out of a wider point table); skip: row r at skip[r * lds .. + c1).
 * Adjoints: dh [P, c] = gradient wrt the layer's ACTIVATED output, (a, b, mean, invstd, c1c2) its BatchNorm
 * fold and adjoint constants (pn2_relu_bn_bwd_stats / pn2_mlp_bwd_finalize); dZ is formed on the fly.
 *   (zero_fill != 0: the call zero-fills its reduction target first; 0: the caller did, ahead of time)
 *   pn2_sa_lead_bwd: du [b*n, c] += dZ at idx; partials [blocks][3][c] = per-CTA sums
 *     of dZ * (x, y, z of the centred coordinates): dW_x[c, k] = sum over blocks of partials[., k, c].
 *   pn2_fp_lead_bwd: dg [b*m, c] += weight * dZ; partials [blocks][max_skip][c]:
 *     dW_s[c, k] = sum over blocks of partials[., k, c].
 * The remaining products (dW_f = du^T features, dfeatures = du W_f; dW_k = dg^T known, dknown = dg W_k) run over
 * the source points (pn2_mlp_wgrad, pn2_mlp_gemm_fwd with pn2_mlp_pack_weights_dgrad[_cols] tiles).
 * ========================================================================== */
/* out [p, c_out] = [in [p, c_in] | 0 ...] (c_out % 4 == 0): the network's input rows (lib/dataset.py's
 * xyz | features layout, pointnet2_semseg.py:139-146) widened to the level-0 point table in one launch */
PN2_API int pn2_pad_rows(long long p, int c_in, int c_out, const float *in, float *out, void *stream);
/* kind: 0 pn2_sa_lead_fwd, 1 pn2_sa_lead_bwd, 2 pn2_fp_lead_fwd, 3 pn2_fp_lead_bwd (each kernel sizes its grid
 * for its own register budget and for what runs beside it) */
PN2_API int pn2_lead_blocks(int p, int c, int kind);
PN2_API int pn2_fp_lead_max_skip(void);
PN2_API int pn2_mlp_pack_weights_cols(int p, int n, int k, int k0, int ldw, const float *w, float *packed,
                                      void *stream);
PN2_API int pn2_mlp_pack_weights_dgrad_cols(int p, int n_out, int n, int k, int k0, int ldw, const float *w,
                                            float *packed, void *stream);
PN2_API int pn2_sa_lead_fwd(int b, int n, int npoint, int nsample, int c, const float *u, const float *xyz,
                            int ldx, const float *new_xyz, const int *idx, const float *wx, int ldw, float *z,
                            float *partials, void *stream);
PN2_API int pn2_sa_lead_bwd(int b, int n, int npoint, int nsample, int c, int relu, const float *dh,
                            const float *z, const float *a, const float *bb, const float *mean,
                            const float *invstd, const float *c1c2, const float *xyz, int ldx,
                            const float *new_xyz, const int *idx, float *du, int zero_fill, float *partials,
                            void *stream);
PN2_API int pn2_fp_lead_fwd(int b, int n, int m, int c, int c1, const float *gk, const float *skip, int lds,
                            const int *idx, const float *weight, const float *ws, int ldw, float *z,
                            float *partials, void *stream);
PN2_API int pn2_fp_lead_bwd(int b, int n, int m, int c, int c1, int relu, const float *dh, const float *z,
                            const float *a, const float *bb, const float *mean, const float *invstd,
                            const float *c1c2, const float *skip, int lds, const int *idx, const float *weight,
                            float *dg, int zero_fill, float *partials, void *stream);

/* ==========================================================================
 * Training-step utilities on flat fp32 buffers (the step the reference's solver
 * runs after backward, lib/solver.py:171-176, as one launch instead of one per
 * parameter tensor).
 * pn2_sum_partials: out[m] = sum over s of in[s, m] (fixed order; m % 4 == 0).
 * pn2_adam_flat: torch.optim.Adam's update (scripts/train.py:44) over one flat
 * parameter buffer; `step` is a device int64 incremented by the call, so the
 * launch is CUDA-graph replayable; the gradient is multiplied by grad_scale first.
 * ========================================================================== */
PN2_API int pn2_sum_partials(int s, int m, const float *in, float *out, void *stream);
/* the same sum over padded images in[s][rows_in][cols_in] (cols_in % 4 == 0), written as the DENSE leading
 * block out[rows_out][cols_out]: the weight gradient in the parameter's own layout */
PN2_API int pn2_sum_partials_block(int s, int rows_in, int cols_in, int rows_out, int cols_out,
                                   const float *in, float *out, void *stream);
/* pn2_dropout_mask: byte keep mask of n elements, P(keep) = keep (resolution 2^-16), counter-based:
 * state = {seed, launch counter} (device int64[2]); the call bumps the counter, so replays of a captured
 * launch draw fresh masks.  The fused chain applies the mask inside the next layer's kernels. */
PN2_API int pn2_dropout_mask(long long n, float keep, long long *state, unsigned char *mask, void *stream);
PN2_API int pn2_adam_flat(long long n, float *param, const float *grad, float *exp_avg,
                          float *exp_avg_sq, long long *step, float lr, float beta1, float beta2,
                          float eps, float weight_decay, float grad_scale, void *stream);

/* ==========================================================================
 * Weighted cross-entropy of the training step.
 * Replaces F.cross_entropy(reduction="none", ignore_index) -> mean(loss * weights)
 * of lib/loss.py:10-18 (log_softmax, nll_loss, mul, mean and their adjoints: six
 * torch kernels) on the classifier's logit rows.
 *
 * logits [p, c] fp32 rows (c % 4 == 0, 16-byte aligned), targets [p] int64,
 * weights [p] fp32 or NULL.
 * pn2_weighted_ce_fwd: loss[0] = 1/p * sum_i w_i * (lse_i - logits[i, t_i]); rows whose
 *   target equals ignore_index contribute 0 (and still count in 1/p, like the reference's
 *   mean over a reduction="none" loss); lse [p] is kept for the adjoint; `partials`
 *   holds pn2_weighted_ce_partials(p) floats; `counter` is ONE zero-initialised
 *   unsigned that the kernel re-arms itself (last-CTA final sum, fixed order, fp64).
 * pn2_weighted_ce_bwd: dlogits[i, c] = grad_loss[0]/p * w_i * (softmax(logits[i])[c] - [c == t_i]);
 *   grad_loss is a DEVICE scalar (the upstream gradient of the loss).
 * ========================================================================== */
PN2_API int pn2_weighted_ce_partials(int p);
PN2_API int pn2_weighted_ce_fwd(int p, int c, long long ignore_index, const float *logits,
                                const long long *targets, const float *weights, float *lse,
                                float *partials, unsigned *counter, float *loss, void *stream);
PN2_API int pn2_weighted_ce_bwd(int p, int c, long long ignore_index, const float *logits,
                                const long long *targets, const float *weights, const float *lse,
                                const float *grad_loss, float *dlogits, void *stream);

/* ==========================================================================
 * Chunk generation with the scene resident in HBM (csrc/chunker.cu).
 * Replaces the NumPy code of lib/dataset.py that cuts training chunks
 * (ScannetDataset.generate_chunks :204-250, __getitem__ :61-94, _augment :99-200) and
 * whole-scene tiles (ScannetDatasetWholeScene.__getitem__ :293-353) and collate_*'s casts
 * (:379-381).  scene [n, ld] float32 rows in the reference's preprocessed layout
 * (x y z | r g b 0..255 | nx ny nz | instance | label; collect_scannet_scenes.py:64-79 --
 * float64 there, float32-representable values); comparisons and arithmetic run in double
 * exactly as NumPy's; the random draws are made by the caller (np.random) and passed in.
 *
 * pn2_crop_try: one try of generate_chunks (:220-238) around scene row `centre_row`:
 *   sel[0 .. count) = rows inside [curmin - 0.2, curmax + 0.2] in scene order;
 *   stats[4] = {count, rows with label > 0, occupied voxels of the 31 x 31 x 62 grid among the rows
 *   inside [curmin - 0.01, curmax + 0.01], rows whose voxel key is not finite (degenerate box)}.
 * pn2_box_select: nbox boxes [lo, hi] (double [nbox, 3], device): sel [nbox, n], count [nbox] and
 *   (optional) mask [nbox, n] bytes = the scene-length membership mask.
 * pn2_chunk_centre: centre[3] (double, device) = mean of the xyz of rows sel[choice[i]]
 *   (sel NULL: choice[i] are scene rows).
 * pn2_chunk_assemble: out [npoints, ldo] float32 = [xyz | features(feat_first) | rgb / 255 | normal |
 *   features(!feat_first)] of those rows, labels int64, weights = labelweights[label] (* wmask[choice[i]]
 *   when wmask is given: the reference's `mask = mask[choice]`, :334).  aug_ops: bit 0 translate,
 *   bit 1 rotate, bit 2 scale, applied in that order about `centre`; aug_host: 13 HOST doubles
 *   t[3], R[9] row-major, f (read during the call).
 * pn2_chunk_batch: the same for `nchunk` chunks in two launches.  desc: DEVICE array of nchunk records of
 * pn2_chunk_desc_bytes() bytes: { scene, feature, sel, choice, wmask pointers (8 bytes each); int ld, nfeat;
 * int aug_ops, 4 bytes padding; double t[3], R[9], f; double centre[3] (scratch, written by the call) };
 * out [nchunk, npoints, ldo], labels / weights [nchunk, npoints].  The row checks of
 * pn2_chunk_assemble (ld, nfeat against ldo) are the caller's.
 * ========================================================================== */
PN2_API int pn2_chunk_desc_bytes(void);
PN2_API int pn2_chunk_batch(int nchunk, int npoints, void *desc, int feat_first, int use_color,
                            int use_normal, int label_col, const float *labelweights, int num_classes,
                            int ldo, float *out, long long *labels, float *weights, void *stream);
PN2_API int pn2_crop_try(int n, int ld, const float *scene, int label_col, int centre_row, double zmin,
                         double zmax, int *sel, int *stats, void *stream);
PN2_API int pn2_box_select(int n, int ld, const float *scene, int nbox, const double *lo, const double *hi,
                           int *sel, int *count, unsigned char *mask, void *stream);
PN2_API int pn2_chunk_centre(int npoints, int ld, const float *scene, const int *sel, const int *choice,
                             double *centre, void *stream);
PN2_API int pn2_chunk_assemble(int npoints, int ld, const float *scene, const int *sel, const int *choice,
                               const float *feature, int nfeat, int feat_first, int use_color,
                               int use_normal, int label_col, const float *labelweights, int num_classes,
                               const unsigned char *wmask, const double *centre, int aug_ops,
                               const double *aug_host, int ldo, float *out, long long *labels,
                               float *weights, void *stream);

/* ==========================================================================
 * Per-iteration evaluation metrics on the device (csrc/metrics.cu).
 * Replace the host path of compute_acc / compute_miou (scripts/eval.py:39-129, called by
 * lib/solver.py:327-347 every iteration): filter_points' Python loop over all points (:42-43),
 * point_cloud_label_to_surface_voxel_label_fast (lib/pc_util.py:39-51) and the per-class passes.
 * coords [p, 3] float32, preds / targets [p] int64.
 * pn2_first_occurrence_xyz: rep[i] = lowest index with the same (bitwise) coordinates -- the points
 *   filter_points keeps are those with rep[i] == i.  table: pn2_metrics_table_ints(p) ints of scratch.
 * pn2_first_occurrence_voxel: vrep[i] = lowest kept index in i's voxel of edge `res` (the voxel key is
 *   evaluated in float32 exactly as NumPy evaluates it there); -1 for filtered points.  bounds: 6 unsigned,
 *   vkey: p unsigned, table as above (all scratch).
 * pn2_confusion: conf[t * nc + q] += 1 over the points with rep[i] == i (caller zero-fills conf [nc, nc] and
 *   bad [1]; bad counts labels outside [0, nc)); nc <= 64.
 * ========================================================================== */
PN2_API int pn2_metrics_table_ints(int p);
PN2_API int pn2_first_occurrence_xyz(int p, const float *coords, int *table, int *rep, void *stream);
PN2_API int pn2_first_occurrence_voxel(int p, const float *coords, const int *rep, float res, unsigned *bounds,
                                       unsigned *vkey, int *table, int *vrep, void *stream);
PN2_API int pn2_confusion(int p, const int *rep, const long long *preds, const long long *targets, int nc,
                          int *conf, int *bad, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* PN2_B200_H */
