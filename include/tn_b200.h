/*
 * tn_b200.h -- C ABI of libtn_b200.so: the thermal-nerfacto per-ray hot path as sm_100a CUDA kernels.
 *
 * The reference (yvette256/nerfstudio-thermal, a nerfstudio 1.0.2 fork) contains no native code; on this
 * path it either runs torch ops (`implementation="torch"`, the parity target) or calls tiny-cuda-nn.  Each
 * entry point below replaces the torch-op sequence of one reference function; the citation after "replaces:"
 * is relative to /root/reference/nerfstudio/.  INTEGRATION.md shows the ctypes binding a maintainer adds.
 *
 * Conventions (every function):
 *   - plain pointers and sizes only; all data pointers are DEVICE pointers owned by the caller
 *     (PyTorch's caching allocator in the shipped host code) unless the name ends in `_host`;
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*); nothing allocates, synchronises or
 *     keeps global mutable state, so calls are re-entrant, one-process-per-GPU safe and CUDA-graph capturable;
 *   - returns 0 on success, a negative TN_E* code otherwise; tn_last_error_string() describes the last
 *     failure on the calling thread;
 *   - all floating-point tensors are float32, row-major, densely packed; "rows" R = rays, S = samples
 *     per ray, N = points.
 */
#ifndef TN_B200_H
#define TN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TN_OK 0
#define TN_EINVAL (-1)   /* bad shape / unsupported size / null pointer */
#define TN_EALIGN (-2)   /* pointer not aligned as required */
#define TN_ECUDA (-3)    /* a CUDA launch failed; see tn_last_error_string() */

#define TN_MAX_LEVELS 32

/* library identification */
int tn_version(void);                     /* 100 * major + minor */
const char* tn_last_error_string(void);   /* thread-local, never NULL */
const char* tn_build_arch(void);          /* "sm_100a" */

/* ------------------------------------------------------------------------------------------------
 * Multiresolution hash grid.      replaces: field_components/encodings.py:401-461
 *   (HashEncoding.hash_fn + pytorch_fwd) and its autograd backward.
 * x[N,3] in [0,1]; table[L*T, F] with T = 1<<log2_T; scales_host[L] = HashEncoding.scalings (host
 * floats, L <= TN_MAX_LEVELS); out[N, L*F] level-major.  F in {1,2,4,8}.
 * table_dtype: 0 = float32, 1 = float16 (derived cache of the fp32 parameter; out stays float32).
 * idx_out: NULL, or int32[N,L,8] receiving the eight table rows per (point, level) in the reference's
 *   corner order hashed_0..hashed_7 (encodings.py:431-438) -- the bit-exactness test hook.
 * jac_out: NULL, or float32[L, F*3, N] (planar) receiving d out[n, l*F+f] / d x[n, c] at [l, f*3+c, n] for the
 *   FINE levels (scale >= 400, TN_JAC_ENC; coarser levels are left untouched) (exclusive with idx_out): handed
 *   back to tn_hash_encode_bwd it saves the backward its second gather of those levels' corner rows.
 * samples_per_ray: 0, or S when the points are the [R,S] samples of R rays in ray-major order (a HINT: results
 *   do not depend on it).  With S % 8 == 0, S <= 64 and N % 4S == 0 a CTA then owns 4 consecutive rays (one 2x2
 *   pixel patch of data/pixel_samplers.py:389-438) and a warp 8 consecutive samples of each, whose points share
 *   grid cells: fewer L1 wavefronts per gather and fewer table REDs in the backward.
 * ------------------------------------------------------------------------------------------------ */
int tn_hash_encode_fwd(const float* x, const void* table, int table_dtype, const float* scales_host,
                       int64_t N, int L, int F, int log2_T, int samples_per_ray, float* out, int32_t* idx_out,
                       float* jac_out, void* stream);

/* dy[N, L*F].  dtable[L*T, F] float32 is ACCUMULATED into (caller zero-fills for a fresh gradient).
 * dx: NULL or float32[N,3] (overwritten) = dL/dx through the interpolation offsets; the fine levels' share is
 * computed from `jac` (the forward's jac_out) when given, everything else by gathering the corner rows of `table`
 * again. */
int tn_hash_encode_bwd(const float* x, const void* table, int table_dtype, const float* scales_host,
                       const float* dy, int64_t N, int L, int F, int log2_T, int samples_per_ray, float* dtable,
                       float* dx, const float* jac, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Sample positions -> normalised grid coordinates.
 *   replaces: cameras/rays.py:49-58 (Frustums.get_positions), field_components/spatial_distortions.py:66-69
 *   (SceneContraction, L-inf) and fields/nerfacto_field.py:207-215 == fields/density_fields.py:96-103
 *   ((p+2)/4, selector, zeroing).
 * origins/directions [R,3]; ebins[R,S+1] euclidean bin edges (sample s spans ebins[s]..ebins[s+1]).
 * x_out[R*S,3]; selector_out[R*S] float (1.0 inside, 0.0 outside).
 * ------------------------------------------------------------------------------------------------ */
int tn_sample_positions_fwd(const float* origins, const float* directions, const float* ebins, int64_t R,
                            int S, float* x_out, float* selector_out, void* stream);
/* dx[R*S,3] -> d_origins[R,3], d_directions[R,3]: overwritten, or added onto when accumulate != 0 (the
 * gradients of a ray bundle's other consumers are then chained through one buffer instead of being summed by
 * autograd).  Bins carry no gradient (model_components/ray_samplers.py:360). */
int tn_sample_positions_bwd(const float* origins, const float* directions, const float* ebins,
                            const float* dx, int64_t R, int S, int accumulate, float* d_origins,
                            float* d_directions, void* stream);
/* The same normalisation for free-standing points p[N,3] (Field.density_fn, fields/base_field.py:48-69). */
int tn_contract_points_fwd(const float* p, int64_t N, float* x_out, float* selector_out, void* stream);
int tn_contract_points_bwd(const float* p, const float* dx, int64_t N, float* dp, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Fully fused small MLPs.          replaces: field_components/mlp.py:159-178 (MLP.pytorch_fwd)
 * ReLU between layers; out_act: 0 none, 1 sigmoid, 2 trunc_exp (field_components/activations.py:28-42).
 * n_layers in {2,3}; weights w[i] are nn.Linear layout [out_i, in_i] row-major, biases b[i][out_i].
 * Supported shapes: in_dim <= 64, width in {16, 64}, out_dim <= 16.
 * x[N,in_dim] -> y[N,out_dim].  Hidden activations never leave the SM.
 * ------------------------------------------------------------------------------------------------ */
int tn_mlp_fwd(const float* x, int64_t N, int in_dim, int width, int out_dim, int n_layers,
               const float* const* w_host_ptrs, const float* const* b_host_ptrs, int out_act, float* y,
               void* stream);
/* dy[N,out_dim] = gradient w.r.t. the ACTIVATED output.  The forward activations are recomputed from x
 * on chip (nothing was saved).  dx: NULL or [N,in_dim] (overwritten).  dw[i]/db[i] (nn.Linear layout)
 * are ACCUMULATED into with atomics: the caller zero-fills them for a fresh gradient. */
int tn_mlp_bwd(const float* x, const float* dy, int64_t N, int in_dim, int width, int out_dim, int n_layers,
               const float* const* w_host_ptrs, const float* const* b_host_ptrs, int out_act, float* dx,
               float* const* dw_host_ptrs, float* const* db_host_ptrs, void* stream);

/* The same MLPs on the tcgen05 tensor cores: 128-point tiles (UMMA M=128), accumulators in TMEM.  Operands
 * are sums of bf16 terms, products a few MMAs accumulated in fp32: three terms / six MMAs in the forward
 * (fp32-faithful pre-activations, so ReLU masks agree with the fp32 reference), two terms / three MMAs in the
 * backward (~2e-5 relative).  Arguments as tn_mlp_fwd, plus:
 * x_stride: floats between consecutive rows of x (and of dx); 0 = in_dim.  A stride that is a multiple of 4
 *   with 16-byte aligned x lets a tile read whole 16-byte runs (a 63-wide input padded to 64: columns
 *   [in_dim, x_stride) are ignored on read and written as zeros in dx when x_stride <= 64);
 * row_mul / out_scale: every output row is multiplied by out_scale * row_mul[p] (row_mul NULL = 1) after the
 *   activation -- with out_act = 2 and a 1-row last layer this is average_init_density * trunc_exp(z) * selector
 *   (fields/nerfacto_field.py:227-228) as the MLP's epilogue;
 * relu_mask_out: NULL, or uint32[N, n_layers-1, max(1,width/32)] receiving bit j of hidden layer l = (z_lj > 0). */
int tn_mlp_tc_fwd(const float* x, int64_t N, int in_dim, int x_stride, int width, int out_dim, int n_layers,
                  const float* const* w_host_ptrs, const float* const* b_host_ptrs, int out_act,
                  const float* row_mul, float out_scale, float* y, uint32_t* relu_mask_out, void* stream);
/* Backward on the tensor cores as well (arguments as tn_mlp_bwd): activations recomputed per tile, dH = dZ.W
 * and dW^T += A^T.dZ as tcgen05 MMAs, dW/db accumulators resident in TMEM across the tiles of a persistent CTA
 * and flushed once with atomics.  relu_mask: the forward's masks (NULL: gate on the recomputed activations). */
int tn_mlp_tc_bwd(const float* x, const float* dy, const uint32_t* relu_mask, int64_t N, int in_dim, int x_stride,
                  int width, int out_dim, int n_layers, const float* const* w_host_ptrs,
                  const float* const* b_host_ptrs, int out_act, const float* row_mul, float out_scale, float* dx,
                  float* const* dw_host_ptrs, float* const* db_host_ptrs, void* stream);

/* Real spherical-harmonics basis, 4 levels (16 components).
 *   replaces: utils/math.py:29-95 via field_components/encodings.py:792-795.  d[N,3] -> out[N,16]. */
int tn_sh4(const float* d, int64_t N, float* out, void* stream);
/* The per-ray inputs of a NerfactoField's colour head in one launch.
 *   replaces: fields/base_field.py:136-142 (get_normalized_directions) + encodings.py:792-795 (SH of (d+1)/2) +
 *   field_components/embedding.py:48-55 (appearance embedding of the ray's camera; fields/nerfacto_field.py:284-290).
 * directions[R,3] -> sh_out[R,16]; emb_out (may be NULL) [R,emb_dim] = embedding[camera_indices[r]]. */
int tn_ray_features(const float* directions, const float* embedding, const int64_t* camera_indices, int64_t R,
                    int emb_dim, float* sh_out, float* emb_out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Samplers.
 * ------------------------------------------------------------------------------------------------ */
/* replaces: model_components/ray_samplers.py:78-128 with the UniformLinDispPiecewise spacing (:244-245).
 * unit_bins[S+1] = torch.linspace(0,1,S+1) (device copy of the host-computed values, so the bin edges
 * are the reference's to the bit); nears/fars [R]; jitter NULL (eval), [R] (the reference's
 * torch.rand((R,1)) draw, single_jitter; jitter_per_sample = 0) or [R,S+1] (jitter_per_sample = 1).
 * sbins_out/ebins_out [R,S+1]: normalised ("spacing") and euclidean bin edges. */
int tn_piecewise_bins(const float* unit_bins, const float* nears, const float* fars, const float* jitter,
                      int jitter_per_sample, int64_t R, int S, float* sbins_out, float* ebins_out, void* stream);
/* replaces: model_components/ray_samplers.py:301-372 (PDFSampler, include_original=False).
 * weights[R,S_old], sbins_old[R,S_old+1].  anneal_dev: NULL (weights are used as given) or a device float a: the
 * histogram is weights ** a, ProposalNetworkSampler's annealing (:602), read at run time so that a captured CUDA
 * graph follows the BEFORE_TRAIN_ITERATION schedule (models/nerfacto.py:271-281).
 * u_base[S_new+1]: eval (jitter == NULL): linspace(0, 1-1/nb, nb) + 1/(2 nb), nb = S_new+1 (:328-329);
 *                  train (jitter given: [R], or [R,nb] with jitter_per_sample = 1): linspace(0, 1-1/nb, nb),
 *                  the kernel adds jitter/nb (:319-325).
 * Outputs sbins_new/ebins_new [R,S_new+1].  S_old <= 1024. */
int tn_pdf_sample(const float* weights, const float* sbins_old, const float* nears, const float* fars,
                  const float* u_base, const float* jitter, int jitter_per_sample, const float* anneal_dev,
                  int64_t R, int S_old, int S_new, float histogram_padding, float eps, float* sbins_new,
                  float* ebins_new, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Volume rendering.
 * ------------------------------------------------------------------------------------------------ */
/* replaces: cameras/rays.py:128-150 (RaySamples.get_weights).  sigma, deltas [R,S] -> weights [R,S]. */
int tn_weights_fwd(const float* sigma, const float* deltas, int64_t R, int S, float* weights, void* stream);
/* dw[R,S] -> dsigma[R,S] (overwritten). */
int tn_weights_bwd(const float* sigma, const float* deltas, const float* dw, int64_t R, int S,
                   float* dsigma, void* stream);

/* replaces: model_components/renderers.py:118-133 (+:292-307 RGBT), :509 (accumulation), :547-557 (median
 * depth), :558-572 (expected depth before the batch-global clip of :574).
 * weights[R,S], colour[R,S,C] (C <= 4), starts/ends [R,S] with t_stride floats between rows (0 = S; the
 * two pointers may be bins and bins+1 of one [R,S+1] bin-edge array with t_stride = S+1).
 * bg_mode: 0 none ("random": no blending), 1 last_sample, 2 constant bg_host[C].
 * eval_mode != 0 applies nan_to_num to colour and clamps the composite to [0,1] (renderers.py:238-245).
 * Any output pointer may be NULL.  steps_minmax_out: float[2] updated with atomic min/max of
 * (starts+ends)/2 over the launch (caller initialises to +inf/-inf) for the :574 clip. */
int tn_render_fwd(const float* weights, const float* colour, const float* starts, const float* ends,
                  int64_t R, int S, int t_stride, int C, int bg_mode, const float* bg_host, int eval_mode,
                  float* rgb_out, float* acc_out, float* depth_median_out, float* depth_expected_out,
                  float* steps_minmax_out, void* stream);
/* Gradients of rgb_out / acc_out / unclipped expected depth -> dweights[R,S], dcolour[R,S,C]
 * (overwritten).  d_rgb/d_acc/d_depth may each be NULL (= zero). */
int tn_render_bwd(const float* weights, const float* colour, const float* starts, const float* ends,
                  const float* d_rgb, const float* d_acc, const float* d_depth, int64_t R, int S, int t_stride,
                  int C, int bg_mode, const float* bg_host, float* dweights, float* dcolour, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Glue fusions between the field kernels, and the per-ray losses.
 * ------------------------------------------------------------------------------------------------ */
/* replaces: fields/nerfacto_field.py:221-228 (split, trunc_exp, selector) + :335-344 (head input concatenation).
 * h[R*S, h_width] = density-MLP output (column 0 = raw density, columns 1..geo_dim = geometry features),
 * sel[R*S], sh[R,16] (per-ray SH basis), emb_ray[R,emb_dim] (per-ray appearance embedding, NULL if emb_dim = 0).
 * density_out[R*S] = density_scale * exp(h0) * sel;  x_out[R*S, x_stride] = [sh | geo | emb | zero padding]
 * (x_stride = 0 means 16+geo_dim+emb_dim; a stride of 64 gives the tensor-core MLP 16-byte aligned rows). */
int tn_field_split_fwd(const float* h, const float* sel, const float* sh, const float* emb_ray, int64_t R, int S,
                       int h_width, int geo_dim, int emb_dim, int x_stride, float density_scale,
                       float* density_out, float* x_out, void* stream);
/* d_density[R*S] and dx[R*S, in] (either may be NULL = zero) -> dh_out[R*S, h_width] (overwritten) and
 * demb_ray_out[R, emb_dim] (overwritten; sum over the samples of each ray; may be NULL). */
int tn_field_split_bwd(const float* h, const float* sel, const float* d_density, const float* dx, int64_t R, int S,
                       int h_width, int geo_dim, int emb_dim, int x_stride, float density_scale, float* dh_out,
                       float* demb_ray_out, void* stream);
/* The colour head of a NerfactoField with its input assembly, one tensor-core kernel each way: replaces
 * fields/nerfacto_field.py:221-228 (split, trunc_exp, selector) + :335-348 (concatenation, mlp_head) and their autograd.
 * h[R*S,16] = density-MLP output (column 0 raw density, 1..15 geometry features), sel[R*S], sh[R,16] (per-ray SH
 * basis), emb_ray[R,32] (per-ray appearance embedding); head weights 63 -> 64 -> 64 -> out_dim (nn.Linear layout,
 * host tables of device pointers as tn_mlp_tc_fwd).  The 63-wide head input [sh | geo | emb] is assembled in
 * registers per 128-point tile and never stored.
 * forward: density_out[R*S] = density_scale * exp(h0) * sel; y[R*S,out_dim]; relu_mask_out uint32[R*S,2,2] or NULL. */
int tn_field_head_fwd(const float* h, const float* sel, const float* sh, const float* emb_ray, int64_t R, int S,
                      int out_dim, float density_scale, const float* const* w_host_ptrs,
                      const float* const* b_host_ptrs, int out_act, float* density_out, float* y,
                      uint32_t* relu_mask_out, void* stream);
/* backward: dy[R*S,out_dim], d_density[R*S] (may be NULL) -> dh_out[R*S,16] (overwritten) =
 * [d_density*scale*sel*exp(clamp(h0)) | dX[:,16:31]]; dz1_ray[R,64] (ACCUMULATED: zero it first) = per-ray sums of
 * the first-layer pre-activation gradients -- the appearance-embedding gradient is dz1_ray . W1[:, 31:63]; dW/db
 * accumulated as in tn_mlp_tc_bwd.  Needs S >= 19 (a 128-point tile touches at most 8 rays). */
int tn_field_head_bwd(const float* dy, const uint32_t* relu_mask, const float* h, const float* sel, const float* sh,
                      const float* emb_ray, const float* d_density, int64_t R, int S, int out_dim,
                      float density_scale, const float* const* w_host_ptrs, const float* const* b_host_ptrs,
                      int out_act, float* dh_out, float* dz1_ray, float* const* dw_host_ptrs,
                      float* const* db_host_ptrs, void* stream);
/* replaces: fields/density_fields.py:116-117.  density[i] = scale * trunc_exp(raw[i*raw_stride]) * sel[i] and its
 * backward (draw_out[i*draw_stride] is overwritten): raw may be a column of a row-major matrix. */
int tn_density_act_fwd(const float* raw, int raw_stride, const float* sel, int64_t N, float scale,
                       float* density_out, void* stream);
int tn_density_act_bwd(const float* raw, int raw_stride, const float* sel, const float* d_density, int64_t N,
                       float scale, float* draw_out, int draw_stride, void* stream);
/* replaces: model_components/losses.py:139-150 (lossfun_distortion) per ray.  weights[R,S], sbins[R,S+1] ->
 * loss_ray_out[R] (un-averaged) and dweights_out[R,S] = d loss_ray / d weights (may be NULL). */
int tn_distortion_loss(const float* weights, const float* sbins, int64_t R, int S, float* loss_ray_out,
                       float* dweights_out, void* stream);
/* replaces: model_components/losses.py:57-103 (outer + lossfun_outer) per ray for ONE proposal level.
 * fine histogram (w_fine[R,S_fine], sbins_fine[R,S_fine+1]) against the proposal histogram (w_prop, sbins_prop).
 * loss_ray_out[R] = sum_i loss_i (un-averaged); dw_prop_out[R,S_prop] = d loss_ray / d w_prop (may be NULL). */
int tn_interlevel_loss(const float* w_fine, const float* sbins_fine, const float* w_prop, const float* sbins_prop,
                       int64_t R, int S_fine, int S_prop, float* loss_ray_out, float* dw_prop_out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Fused proposal density field: ray samples -> density in one kernel, and its whole backward in one kernel.
 *   replaces: fields/density_fields.py:95-118 (HashMLPDensityField.get_density) end to end, i.e.
 *   cameras/rays.py:49-58 + spatial_distortions.py:66-69 + encodings.py:401-461 (L <= 8 levels, F = 2) +
 *   mlp.py:159-178 (Linear(2L,hidden=16) ReLU Linear(16,1)) + activations.py:28-42 (trunc_exp) + selector.
 * origins/directions [R,3], ebins[R,S+1], table[L*T,2] fp32, w1[16,2L] b1[16] w2[1,16] b2[1] (nn.Linear layout).
 * density_out[R*S] = density_scale * exp(raw) * selector.
 * ------------------------------------------------------------------------------------------------ */
int tn_prop_density_fwd(const float* origins, const float* directions, const float* ebins, const float* table,
                        const float* scales_host, int64_t R, int S, int L, int log2_T, int hidden, const float* w1,
                        const float* b1, const float* w2, const float* b2, float density_scale,
                        float* density_out, void* stream);
/* d_density[R*S] -> dtable (ACCUMULATED), dw1/db1/dw2/db2 (ACCUMULATED, atomics), and, when non-NULL,
 * d_origins/d_directions [R,3] (ACCUMULATED with atomics: zero them first). */
int tn_prop_density_bwd(const float* origins, const float* directions, const float* ebins, const float* table,
                        const float* scales_host, int64_t R, int S, int L, int log2_T, int hidden, const float* w1,
                        const float* b1, const float* w2, const float* b2, float density_scale,
                        const float* d_density, float* dtable, float* dw1, float* db1, float* dw2, float* db2,
                        float* d_origins, float* d_directions, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Model glue as single launches.
 * ------------------------------------------------------------------------------------------------ */
/* replaces: cameras/camera_optimizers.py:132-176 (CameraOptimizer.forward + apply_to_raybundle, mode SO3xR3 or
 * shared_SO3xR3) with cameras/lie_groups.py:24-59.  pose[num_cams|1, 6] = (translation, so(3) log-rotation);
 * frozen: NULL or uint8[num_cams] (1 = non-trainable camera, identity correction); camera_indices int64[R].
 * origins_out = o + t, directions_out = R(w) d. */
int tn_camera_opt_fwd(const float* pose, const uint8_t* frozen, const int64_t* camera_indices, const float* origins,
                      const float* directions, int64_t R, int shared_pose, float* origins_out,
                      float* directions_out, void* stream);
/* gradients w.r.t. the outputs (either may be NULL) -> dpose[num_cams|1, 6], ACCUMULATED with atomics. */
int tn_camera_opt_bwd(const float* pose, const uint8_t* frozen, const int64_t* camera_indices,
                      const float* directions, const float* d_origins_out, const float* d_directions_out, int64_t R,
                      int shared_pose, float* dpose, void* stream);
/* replaces: models/thermal_nerfacto.py:286-354 pixel terms (MSE on RGB / thermal rays, model_components/losses.py:
 * 603-620 tv_pixel_loss, :623-651 cross_channel_loss, utils/rgbt_utils.py:6-33) for patch-ordered batches (groups
 * of four rays = one 2x2 patch of one camera).  rgb[R,3], thermal[R] (NULL in rgb_only mode), image[R,3],
 * is_thermal[R].  losses_out[4] = {rgb MSE, thermal MSE, tv_pixel, cross_channel} (un-multiplied; may be NULL).
 * d_rgb_out[R,3] / d_thermal_out[R] (may be NULL) = sum_k upstream[k] * d losses[k] / d input; upstream: device
 * float[4]. */
int tn_pixel_losses(const float* rgb, const float* thermal, const float* image, const float* is_thermal, int64_t R,
                    const float* upstream, float* losses_out, float* d_rgb_out, float* d_thermal_out, void* stream);
/* replaces: models/thermal_nerfacto.py:328-344 (cross-field density L1 with its stop-gradient pattern).
 * d, d2, dt, d2t [N].  partial_out (may be NULL when only gradients are wanted) [n_partial + 1]: per-CTA partial sums
 * of value_mult * (mean|d2-dt| + mean|d-d2t|); with `ticket` (a zero-initialised uint32 the caller owns; re-armed by
 * the kernel) the CTA that finishes last also writes their sum, in index order, to partial_out[n_partial].
 * Gradients (all four or none), times upstream_dev[0] when given: g_dt = -thermal_grad_mult*sgn(d2-dt)/N,
 * g_d2t = -thermal_grad_mult*sgn(d-d2t)/N, g_d2 = rgb_grad_mult*sgn(d2-dt)/N, g_d = rgb_grad_mult*sgn(d-d2t)/N. */
int tn_density_l1(const float* d, const float* d2, const float* dt, const float* d2t, int64_t N, float value_mult,
                  float thermal_grad_mult, float rgb_grad_mult, const float* upstream_dev, float* partial_out,
                  int n_partial, uint32_t* ticket, float* g_d, float* g_d2, float* g_dt, float* g_d2t, void* stream);
/* replaces: cameras/camera_optimizers.py:188-194 (get_loss_dict) and :200-204 (get_metrics_dict).  pose[num_cameras,6].
 * out3 = {(mean|t_i| * trans_penalty + mean|w_i| * rot_penalty) * penalty_scale, |T|_F, |W|_F}. */
int tn_camera_reg_fwd(const float* pose, int num_cameras, float trans_penalty, float rot_penalty,
                      float penalty_scale, float* out3, void* stream);
/* dpose_out[num_cameras,6] (overwritten; atomically added onto when accumulate != 0) = upstream_dev[0] * d out3[0] /
 * d pose (zero rows where a norm is zero, as torch's norm backward). */
int tn_camera_reg_bwd(const float* pose, const float* upstream_dev, int num_cameras, float trans_penalty,
                      float rot_penalty, float penalty_scale, int accumulate, float* dpose_out, void* stream);
/* Gradient of the appearance embedding (field_components/embedding.py:48-55 under autograd) from the colour head's
 * per-ray first-layer gradient dz1_ray[R,width] (tn_field_head_bwd):
 *   dweight[camera_indices[r], j] += sum_k dz1_ray[r,k] * w0[k, col0 + j],  j < emb_dim,
 * w0[width,in_dim] = the head's first nn.Linear weight, col0 = the embedding's first input column.  clear != 0: the
 * rows of dz1_ray are zeroed after use (a buffer that tn_field_head_bwd accumulates into stays clean between steps). */
int tn_embed_bwd(float* dz1_ray, const float* w0, const int64_t* camera_indices, int64_t R, int width, int in_dim,
                 int col0, int emb_dim, int clear, float* dweight, void* stream);
/* replaces: the loss-dictionary arithmetic of models/thermal_nerfacto.py:284-388 + engine/trainer.py:479
 * (`functools.reduce(torch.add, loss_dict.values())`).  term_host_ptrs[k]: device pointer of scalar term k, which
 * belongs to dictionary entry slot_host[k] in [0, n_slots);
 * out[1+j] = sum_{k: slot k = j} scale[k] * term_k (the dictionary values), out[0] = their sum (the total loss). */
#define TN_MAX_LOSS_TERMS 16
int tn_loss_sum(const float* const* term_host_ptrs, const float* scale_host, const int* slot_host, int n_terms,
                int n_slots, float* out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * The step before the path, on the device: patch pixel sampling, collation, ray generation.
 * ------------------------------------------------------------------------------------------------ */
/* replaces: data/pixel_samplers.py:417-438 (PatchPixelSampler.sample_method, no mask).  u[num_patches,3] = the
 * reference's torch.rand draw (device); indices_out int64[num_patches*patch^2, 3] = (image, row, column). */
int tn_patch_pixel_indices(const float* u, int64_t num_patches, int patch_size, int num_images, int image_height,
                           int image_width, int64_t* indices_out, void* stream);
/* replaces: data/pixel_samplers.py:239-256 (collate_image_dataset_batch, without the CPU index round trip).
 * images[num_images,H,W,channels] float32 (pixel_dtype 0) or uint8 (1, divided by 255); indices int64[R,3] (column
 * 0 is REWRITTEN with image_idx[c] when image_idx is non-NULL); image_out[R,channels];
 * is_thermal_out[R] = is_thermal_per_image[c] (either may be NULL). */
int tn_gather_pixels(const void* images, int pixel_dtype, int64_t num_images, int image_height, int image_width,
                     int channels, int64_t* indices, const int64_t* image_idx, const float* is_thermal_per_image,
                     int64_t R, float* image_out, float* is_thermal_out, void* stream);
/* replaces: model_components/ray_generators.py:40-55 + cameras/cameras.py:504-905 for undistorted PERSPECTIVE
 * cameras.  ray_indices int64[R,3] = (camera, row, column); camera_to_worlds[num_cameras,3,4]; intrinsics
 * [num_cameras,4] = (fx, fy, cx, cy).  Outputs: origins[R,3], directions[R,3] (unit), pixel_area[R],
 * directions_norm[R] (NULL ok), camera_indices int64[R] (NULL ok). */
int tn_generate_rays(const int64_t* ray_indices, const float* camera_to_worlds, const float* intrinsics,
                     int64_t num_cameras, int64_t R, float* origins_out, float* directions_out,
                     float* pixel_area_out, float* directions_norm_out, int64_t* camera_indices_out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Optimiser step over the flat buffers.
 * ------------------------------------------------------------------------------------------------ */
#define TN_ADAM_MAX_GROUPS 16
#define TN_ADAM_HYPER 8
/* replaces: engine/optimizers.py:172-180 (Optimizers.optimizer_step_all: one torch.optim.Adam.step per group),
 * :150-163 (optimizer_scaler_step_all: GradScaler.step = skip on inf/nan) and engine/schedulers.py:109-142
 * (ExponentialDecayScheduler: per-group LambdaLR), for the groups of configs/method_configs.py:274-301.
 * params/grads/exp_avg/exp_avg_sq: flat fp32 device buffers of n elements (16-byte aligned).  Group i covers
 * elements [group_begin_host[i], group_end_host[i]) (ordered, disjoint; elements of no group are left alone) and
 * group_hyper_host[i*8 ..] = {lr_init, lr_final (<=0: lr_init), lr_pre_warmup, eps, weight_decay, warmup_steps,
 * max_steps (0: constant lr), ramp (0 linear, 1 cosine)}.  Step numbers are read from the device when step_dev
 * is non-NULL (so the call can live in a CUDA graph), else step_host serves for both: step_dev[0] = i, the 1-based
 * number of this training iteration -- the learning rate used is lr_init * lr_lambda(i-1) (the schedulers step
 * once per iteration for every group, optimizers.py:182-192), evaluated on the device in double; with
 * per_group_steps != 0, step_dev[1+k] = t_k, the 1-based number of group k's own Adam step (torch's state['step']
 * after the increment: it lags i for groups that sat out iterations without a gradient).  active_mask bit k = 0:
 * group k is left untouched this call (its gradients are None in the reference, :165-170).  Dense update exactly
 * as torch.optim.Adam (amsgrad off): m = lerp(m,g,1-b1); v = b2 v + (1-b2) g^2; p -= lr/(1-b1^t) * m /
 * (sqrt(v)/sqrt(1-b2^t) + eps).  inv_scale_dev: NULL or device float multiplying the gradients (GradScaler);
 * found_inf_dev: NULL or device float, non-zero = skip the update; zero_grads: clear grads after reading them
 * (inactive groups' and padding elements included). */
int tn_adam_step(float* params, float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                 const int64_t* group_begin_host, const int64_t* group_end_host, const float* group_hyper_host,
                 int n_groups, double beta1, double beta2, const int32_t* step_dev, int step_host,
                 int per_group_steps, uint32_t active_mask, const float* inv_scale_dev,
                 const float* found_inf_dev, int zero_grads, void* stream);
/* counters_dev[0] += 1 (iteration) and counters_dev[1+k] += 1 for every group k with active_mask bit k set:
 * the device-side counters tn_adam_step reads (per_group_steps layout), advanced once per iteration. */
int tn_step_counters_tick(int32_t* counters_dev, int n_groups, uint32_t active_mask, void* stream);
/* replaces: torch.cuda.amp.GradScaler.unscale_'s inf check (used by engine/optimizers.py:150-163):
 * found_inf_dev[0] = 1 if any grads[i] * inv_scale is not finite, else 0. */
int tn_grad_unscale_check(const float* grads, int64_t n, const float* inv_scale_dev, float* found_inf_dev,
                          void* stream);
/* counter_dev[0] += value (the device-side step counter read by tn_adam_step inside a captured graph). */
int tn_counter_add(int32_t* counter_dev, int value, void* stream);

/* ------------------------------------------------------------------------------------------------
 * One launch per sampling level (csrc/tn_level.cu): the composite + resample kernels.  Same arithmetic as the
 * single-purpose entry points above, so the results are theirs.
 * ------------------------------------------------------------------------------------------------ */
/* Proposal level.  replaces: cameras/rays.py:128-150 (get_weights) + renderers.py:547-557 (median depth of the
 * level, the prop_depth_i outputs of models/nerfacto.py:346-351) + ray_samplers.py:301-372, :602 (annealed PDF
 * resampling).  sigma[R,S], ebins/sbins [R,S+1] of this level; u_base / jitter / anneal_dev / histogram_padding /
 * eps as tn_pdf_sample.  weights_out[R,S]; depth_median_out[R] or NULL; sbins_new/ebins_new [R,S_new+1], or both
 * NULL for the weights (and depth) alone. */
int tn_level_resample(const float* sigma, const float* ebins, const float* sbins, const float* nears,
                      const float* fars, const float* u_base, const float* jitter, int jitter_per_sample,
                      const float* anneal_dev, int64_t R, int S, int S_new, float histogram_padding, float eps,
                      float* weights_out, float* depth_median_out, float* sbins_new, float* ebins_new,
                      void* stream);
/* Final level.  replaces: get_weights + every renderer of models/nerfacto.py:316-320 (arguments as
 * tn_render_fwd with starts/ends = ebins[:, :-1] / ebins[:, 1:]) + model_components/losses.py:139-158
 * (distortion) + losses.py:57-135 (interlevel loss against n_prop <= 2 proposal histograms prop_w[q][R,S_q],
 * prop_sbins[q][R,S_q+1]).  loss_acc: NULL (no losses), or float[2] ACCUMULATED into: [0] += sum over rays of
 * lossfun_distortion, [1] += sum over rays, levels and fine samples of lossfun_outer (the caller zero-fills and
 * divides by R resp. R*S).  dw_distortion_out[R,S] / prop_dw[q][R,S_q]: NULL, or the gradients of those two sums
 * w.r.t. the final / proposal weights, kept for tn_ray_heads_bwd.
 * scratch5 (may be NULL): five 4-byte words {+inf, -inf, 0, 0, 0u} owned by the caller and never used by two launches at
 * a time.  When given, the launch-wide atomics go THERE, and the CTA that finishes last writes the final values --
 * steps_minmax_out = (min, max), loss_acc = (distortion sum * loss_scale_distortion, interlevel sum *
 * loss_scale_interlevel), depth_expected_clipped_out[R] = clamp(depth_expected_out, min, max) (renderers.py:574) --
 * and puts the scratch back into its initial state: no initialisation, scaling or clamp launches around the call. */
int tn_ray_heads_fwd(const float* sigma, const float* colour, const float* ebins, const float* sbins, int64_t R,
                     int S, int C, int bg_mode, const float* bg_host, int eval_mode, int n_prop,
                     const float* const* prop_w_host_ptrs, const float* const* prop_sbins_host_ptrs,
                     const int* prop_S_host, float* const* prop_dw_host_ptrs, float* weights_out, float* rgb_out,
                     float* acc_out, float* depth_median_out, float* depth_expected_out, float* steps_minmax_out,
                     float* loss_acc, float* dw_distortion_out, float* scratch5, float loss_scale_distortion,
                     float loss_scale_interlevel, float* depth_expected_clipped_out, void* stream);
/* The ray-level backward of a whole branch in one launch.  Final level: d_rgb[R,C], d_acc[R], d_depth_expected[R]
 * (each may be NULL) and g_distortion_dev (device scalar dL/d(mean distortion), NULL = 0) times dw_distortion ->
 * get_weights backward -> dsigma_out[R,S], dcolour_out[R,S,C] (NULL: not wanted).  Proposal level q:
 * g_interlevel_dev (device scalar dL/d(mean interlevel loss)) / (R*S) * prop_dw[q] -> get_weights backward with that
 * level's prop_sigma[q][R,S_q], prop_ebins[q][R,S_q+1] -> prop_dsigma[q][R,S_q]. */
int tn_ray_heads_bwd(const float* sigma, const float* colour, const float* ebins, const float* weights,
                     const float* dw_distortion, const float* d_rgb, const float* d_acc,
                     const float* d_depth_expected, const float* g_distortion_dev, const float* g_interlevel_dev,
                     int64_t R, int S, int C, int bg_mode, const float* bg_host, int n_prop,
                     const float* const* prop_sigma_host_ptrs, const float* const* prop_ebins_host_ptrs,
                     const float* const* prop_dw_host_ptrs, const int* prop_S_host, float* dsigma_out,
                     float* dcolour_out, float* const* prop_dsigma_host_ptrs, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Gradient exchange over NVLink peer memory (replaces, for the main fields' 134 MB, the bucketed all-reduce of
 * DistributedDataParallel, pipelines/base_pipeline.py:280-283): the bulk moves by copy-engine pulls between
 * symmetric buffers (host side: parallel.PeerExchange), this is the only kernel of it.
 * dst[n] = (dst[n] + sum_k src[k * src_stride + n]) * scale for n_src staged copies; n and src_stride multiples of 4,
 * 16-byte aligned pointers.  max_ctas: grid cap (0 = 2 per SM), so that the reduction can stay out of the way of
 * compute kernels running beside it.
 * ------------------------------------------------------------------------------------------------ */
int tn_shard_mean(float* dst, const float* src, int n_src, int64_t src_stride, int64_t n, float scale, int max_ctas,
                  void* stream);

/* ------------------------------------------------------------------------------------------------
 * Measurement aid (bench.py, SURVEY 8d "L2 peak must be measured"): the access pattern of the hash-grid kernels
 * and nothing else.  Every thread of `ctas` x 256 issues `iters` (multiple of 8) operations at pseudo-random rows
 * of table[2^log2_rows, 2] (float32, 16-byte aligned; REDs modify it):
 *   mode 0 8-byte gathers, a row per lane | 1 lanes 2i,2i+1 read an aligned pair of rows | 6 four lanes read one
 *   32-byte sector | 2 8-byte vector REDs, a row per lane | 3 lane pairs add to an aligned pair of rows |
 *   4 16-byte vector REDs | 5 4-byte REDs | 7, 8 gathers / 9, 10 REDs by lane pairs at rows (r, r^3) = same sector,
 *   other 16-byte half / (r, r^7) = same 128-byte line, other sector.   sink: one float, never written in practice.
 * ------------------------------------------------------------------------------------------------ */
int tn_l2_probe(int mode, float* table, int log2_rows, int iters, int ctas, float* sink, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TN_B200_H */
