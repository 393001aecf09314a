/*
 * gbnerf.h — C ABI of libgbnerf.so, the B200 (sm_100a) implementation of GB-NeRF's DS_NeRF
 * volumetric-rendering hot path.
 *
 * This is the drop-in boundary: everything the reference computes with ATen ops inside
 * run.py:render_rays / DS_NeRF/run_nerf_helpers.py is one of the entry points below.  The reference has no
 * FFI of its own for this path (it is pure PyTorch); the binding a maintainer adds is the ctypes stub shown
 * in INTEGRATION.md, and `gb-nerf_b200/_lib.py` is exactly that stub.
 *
 * Conventions
 *   - plain C: raw DEVICE pointers, int64 sizes, scalar flags, a cudaStream_t passed as void*.
 *   - the caller owns every buffer (inputs, outputs, workspaces); the library only borrows them for the
 *     duration of the enqueued work.  Nothing is allocated, freed or synchronised inside.
 *   - all work is enqueued on `stream`; no default-stream use, no hidden syncs.
 *   - every function returns 0 on success, else a GBN_E* code; gbn_last_error_string() describes the last
 *     failure on the calling thread.  Nothing throws, nothing calls exit().
 *   - fp32 tensors are dense row-major unless a stride argument says otherwise.  "ray_stride" is the row
 *     pitch, in floats, of the reference's packed ray batch (run.py:1726-1736: [o(3) d(3) near far (depth)
 *     viewdir(3)] -> 11 or 12), so views into that batch can be passed without a copy.
 *   - there is NO CPU fallback: without a CUDA device every compute entry point returns GBN_ECUDA.
 */
#ifndef GBNERF_H_
#define GBNERF_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GBN_OK 0
#define GBN_EINVAL 1       /* bad argument (null pointer, size, alignment, unsupported shape) */
#define GBN_ECUDA 2        /* a CUDA runtime call / kernel launch failed */
#define GBN_EUNSUPPORTED 3 /* valid request this build cannot serve */

#define GBN_PRECISION_BF16 0 /* bf16 operands, fp32 accumulate (tcgen05 kind::f16)  */
#define GBN_PRECISION_TF32 1 /* tf32 operands, fp32 accumulate (tcgen05 kind::tf32) */
#define GBN_PACK_BWD_BF16 2  /* prepack only: transposed bf16 weights for gbn_mlp_backward_data */

/* Network geometry fixed by the reference (run_nerf_helpers.py:75-104 with D=8, W=256, skips=[4],
 * multires=10, multires_views=4, use_viewdirs=True). */
#define GBN_PTS_CH 63
#define GBN_DIR_CH 27
#define GBN_EMB_CH 90
#define GBN_WIDTH 256
#define GBN_NUM_LINEAR 12 /* pts_linears.0-7, feature_linear, alpha_linear, views_linears.0, rgb_linear */
#define GBN_PARAM_COUNT 595844

int gbn_version(void);
const char* gbn_last_error_string(void);
/* Number of CUDA kernels this library has launched in this process so far (every launch site counts itself). */
unsigned long long gbn_kernel_launches(void);

/* ---- ray setup of render(): run.py:1700-1736 with get_rays / ndc_rays (run_nerf_helpers.py:251-262, 285-302) ------
 * Writes the packed ray batch [R, 8 (+1 if depths) (+3 if use_viewdirs)] = o, d, near, far, (depth), (unit viewdir).
 * Either a camera pose: c2w DEVICE [3,>=4] row-major with pitch c2w_ld, rays of the patch rows [patch_i, +patch_h) x
 * columns [patch_j, +patch_w) of the H x W frame (R = patch_h*patch_w; full frame: 0,0,H,W), optional c2w_static
 * (run.py:1713: origins/directions from it, view directions from c2w); or a ray batch rays_o / rays_d [R,3] (pitches
 * o_ld / d_ld floats) with c2w = NULL.  ndc != 0 applies ndc_rays(H, W, focal, 1., o, d) after the view directions
 * are taken, as render() does. */
int gbn_pack_rays(const float* c2w, int c2w_ld, const float* c2w_static, int c2w_static_ld, const float* rays_o,
                  int64_t o_ld, const float* rays_d, int64_t d_ld, const float* depths, int H, int W, double focal,
                  int patch_i, int patch_j, int patch_h, int patch_w, int use_viewdirs, int ndc, float near, float far,
                  int64_t R, float* rays_out, void* stream);

/* ---- stratified depths: run.py:2291-2315 ---------------------------------------------------------------
 * near/far: [R] with pitch ray_stride (floats).  t_rand: [R,S] uniform [0,1) or NULL (perturb == 0).
 * z out: [R,S].  lindisp != 0 samples linearly in inverse depth. */
int gbn_zvals_stratified(const float* near, const float* far, int64_t ray_stride, int64_t R, int S,
                         int lindisp, const float* t_rand, float* z, void* stream);

/* ---- standalone positional encoding: Embedder.embed (run_nerf_helpers.py:23-53) + run_network's
 * concatenation (run.py:1640-1647).  pts = o + d*z per sample, dirs broadcast along the ray.
 * out: [R*S, 90] fp32 = [pts(3) sin/cos x10 | dir(3) sin/cos x4].  Measurement / test entry point — the
 * MLP kernel fuses this stage and never materialises the tensor. */
int gbn_encode_points(const float* rays_o, const float* rays_d, const float* viewdirs, int64_t ray_stride,
                      const float* z, int64_t R, int S, float* out, void* stream);

/* ---- alpha compositing: raw2outputs (run_nerf_helpers.py:352-406) ---------------------------------------
 * raw [R,S,4], z [R,S], rays_d [R] x3 with pitch ray_stride, noise [R,S] (already scaled by raw_noise_std) or
 * NULL.  Outputs rgb [R,3], disp [R], acc [R], depth [R], weights [R,S], alpha [R,S] or NULL. */
int gbn_composite_forward(const float* raw, const float* z, const float* rays_d, int64_t ray_stride,
                          const float* noise, int64_t R, int S, int white_bkgd, float* rgb, float* disp,
                          float* acc, float* depth, float* weights, float* alpha, void* stream);

/* Backward of the above.  g_weights [R,S] may be NULL.  The forward is recomputed
 * from raw/z, nothing else is read.  g_raw out: [R,S,4]. */
int gbn_composite_backward(const float* raw, const float* z, const float* rays_d, int64_t ray_stride,
                           const float* noise, int64_t R, int S, int white_bkgd, int detach_weights,
                           const float* g_rgb, const float* g_disp, const float* g_acc, const float* g_depth,
                           const float* g_weights, float* g_raw, void* stream);

/* ---- inverse-CDF sampling: sample_pdf (run_nerf_helpers.py:306-349) --------------------------------------
 * bins [R,B], weights [R,B-1], u [R,N] or NULL (deterministic linspace(0,1,N)).  samples out [R,N]. */
int gbn_sample_pdf(const float* bins, const float* weights, const float* u, int64_t R, int B, int N,
                   float* samples, void* stream);

/* The search on its own, for the bit-exact index test: inds[r,n] = #{j : cdf[r,j] <= u[r,n]}
 * == torch.searchsorted(cdf, u, right=True) (run_nerf_helpers.py:333).  inds out: int64 [R,N]. */
int gbn_searchsorted_right(const float* cdf, const float* u, int64_t R, int B, int N, int64_t* inds,
                           void* stream);

/* Fused hierarchical step of render_rays (run.py:2343-2348, 2370): z_mid, sample_pdf on weights[:,1:-1],
 * sort-merge with the coarse depths, std of the new samples.
 * z_vals [R,S] ascending, weights [R,S], u [R,N] or NULL.  Outputs: z_samples [R,N] (may be NULL),
 * z_merged [R,S+N] ascending, z_std [R] (may be NULL). */
int gbn_sample_pdf_merge(const float* z_vals, const float* weights, const float* u, int64_t R, int S, int N,
                         float* z_samples, float* z_merged, float* z_std, void* stream);

/* The same call with the two hooks SURVEY 8(b) asks for, so that the north-star tolerance "bin indices bit-exact given
 * the same CDF and uniforms" can be asserted on the production kernel itself: cdf_in (NULL, or [R, S-1] fp32: used
 * instead of the cdf built from `weights`, which may then be NULL) and inds_out (NULL, or [R, N] int32 receiving
 * searchsorted(cdf, u, right=True) of run_nerf_helpers.py:333 for every sample, in the order of `u`). */
int gbn_sample_pdf_merge_ex(const float* z_vals, const float* weights, const float* u, const float* cdf_in, int64_t R,
                            int S, int N, float* z_samples, float* z_merged, float* z_std, int* inds_out, void* stream);

/* ---- the 8x256 NeRF MLP: NeRF.forward (run_nerf_helpers.py:106-129) --------------------------------------
 * Weights are re-laid-out once per optimiser step into the kernel's shared-memory image
 * (UMMA K-major, 128-byte swizzle, K padded to 64) — `params` is a HOST array of 24 DEVICE pointers in the
 * order weight,bias of: pts_linears.0 … pts_linears.7, feature_linear, alpha_linear, views_linears.0,
 * rgb_linear (nn.Linear layout: weight [out,in] row-major fp32). */
size_t gbn_mlp_packed_bytes(int precision);
int gbn_mlp_prepack_weights(const void* const* params, void* packed, int precision, void* stream);

/* Fused point generation + positional encoding + MLP.  Row p = r*S + s evaluates the point
 * o_r + d_r * z[r,s] with view direction viewdirs_r.  If `pts` is non-NULL it is a dense [R*S,3] tensor
 * used instead of o + d*z (network_query_fn's general form, run.py:2059-2062).  raw out: [R*S,4] fp32 =
 * (r,g,b,sigma_raw).  workspace: gbn_mlp_workspace_bytes(R) bytes. */
size_t gbn_mlp_workspace_bytes(int64_t R);
int gbn_mlp_forward(const void* packed, int precision, const float* rays_o, const float* rays_d,
                    const float* viewdirs, int64_t ray_stride, const float* z, const float* pts, int64_t R,
                    int S, float* raw, void* workspace, void* stash, void* stream);

/* Training stash (bf16 only).  When `stash` is non-NULL the forward also writes, per 128-point tile, the bf16
 * activations the backward needs (8 hidden layers, feature, view-branch hidden, point and direction encoding) as
 * 16 KB blocks of [128 points x 64 channels], laid out [64-point half][8-channel chunk][64 points][16 B]: opaque to
 * the caller, consumed by gbn_mlp_backward_data / gbn_mlp_backward_weights.  gbn_mlp_stash_bytes(P) bytes,
 * 128-byte aligned. */
size_t gbn_mlp_stash_bytes(int64_t P);

/* Diagnostic: when buf is non-NULL (device memory, >= 16 KiB, zeroed by the caller), CTA 0 of every following MLP
 * launch records clock64() stamps of its producer / MMA / encoder / epilogue roles for its `tile`-th tile
 * (layout: tools/mlp_trace.py).  NULL switches tracing off.  Not part of the reference-facing surface. */
int gbn_mlp_set_trace(void* buf, int tile);

/* Post-mortem of the MLP kernels' barrier watchdog.  Every mbarrier wait in nerf_mlp_ts_kernel is bounded (~4 s);
 * the first wait that expires writes a record into zero-copy HOST memory, so it can be read after the CUDA context
 * has died: out[0] = wait code (role << 24 | job/step), out[1] = CTA, out[2] = thread, out[3] = 1 if a record exists,
 * out[8 + 2i .. 9 + 2i] = raw 64-bit state of the i-th barrier counted back from the end of the kernel's barrier block,
 * out[128 + w] = code of the last wait warp w (of any aborting CTA, out[160 + w]) left through the abort path.  Copies up to `words` (<= 256)
 * 32-bit words and returns the number copied (0 before the first MLP launch).  Makes no CUDA call.
 * Not part of the reference-facing surface. */
int gbn_watchdog_report(unsigned int* out, int words);

/* Diagnostic, host only (no CUDA call): the job and epilogue-step tables of the bf16 TMEM-operand kernels, as the
 * kernels read them from constant memory (forward: bwd = 0, dgrad: bwd = 1).  jobs: 16-byte TsJob records, steps:
 * 12-byte TsStep records (csrc/mlp_ts_layout.h); meta[0..9] = jobs, steps, a_ready completions per tile [4],
 * issue-order signals per tile, acc1_empty completions per tile, weight-ring stages, split hand-over in use.
 * tests/test_ts_protocol_cpu.py replays the barrier protocol of these tables under random latencies.
 * bwd = 2: the tables of the two-tiles-in-flight kernel (csrc/mlp_t2.cuh; inference and the stash-writing training
 * forward): 16-byte T2Job and 8-byte T2Step records (a step names the first job of its MMA group and the first H-stash
 * block of its output); meta[0..7] = jobs, steps, weight-ring stages, scheduling mode, slab offsets of alpha_linear (2) and rgb_linear,
 * kernel enabled (tests/test_t2_protocol_cpu.py). */
int gbn_debug_ts_plan(int bwd, void* jobs, int max_jobs, void* steps, int max_steps, int* meta);

/* Which bf16 kernel family this process uses (env GBNERF_MLP): 0 = operands in shared memory ("ss"), 1 = activations
 * in tensor memory ("ts", default).  The packed weight images differ. */
int gbn_mlp_variant(void);

/* ---- optimizer step fused with the weight re-pack: run.py:1529 `optimizer.step()` on the Adam of run.py:2065 --------
 * One network per call.  params / grads / exp_avg / exp_avg_sq: HOST arrays of 24 DEVICE pointers (fp32, nn.Linear
 * layout, the order of gbn_mlp_prepack_weights).  Updates exp_avg, exp_avg_sq and params in place with
 * torch.optim.Adam's arithmetic (no weight decay, no amsgrad); `step` is the 1-based count of this step (bias
 * corrections 1 - beta^step are taken on the host in double, as torch does).  packed_fwd (GBN_PRECISION_BF16 image)
 * and packed_bwd (GBN_PACK_BWD_BF16 image), each NULL or previously filled by gbn_mlp_prepack_weights, receive the
 * new values at their positions in the same launch, so no re-pack pass follows an optimizer step. */
int gbn_adam_step_repack(void* const* params, const void* const* grads, void* const* exp_avg, void* const* exp_avg_sq,
                         double lr, double beta1, double beta2, double eps, int64_t step, void* packed_fwd,
                         void* packed_bwd, void* stream);

/* The same step for a captured CUDA graph (replayed once per training step): nothing step-dependent is a kernel
 * argument.  gbn_adam_tick (one thread) adds 1 to step_state[0] (device double, the 1-based step count) and writes
 * scalars[0] = lr[0] / (1 - beta1^step), scalars[1] = sqrt(1 - beta2^step) (device floats; lr is a device float the
 * host refreshes before a replay - run.py:1540-1544 decays it every step).  gbn_adam_step_repack_dev is
 * gbn_adam_step_repack reading those two scalars from device memory. */
int gbn_adam_tick(double* step_state, const float* lr, double beta1, double beta2, float* scalars, void* stream);
int gbn_adam_step_repack_dev(void* const* params, const void* const* grads, void* const* exp_avg, void* const* exp_avg_sq,
                             const float* scalars, double beta1, double beta2, double eps, void* packed_fwd,
                             void* packed_bwd, void* stream);

/* ---- depth -> normal map: depth2normal_geo, run.py:2458-2474 (called at run.py:1440-1443) ---------------------------
 * points [B,3,H,W] (xyz per pixel, depth2xyz_torch's output moved to channel-first as the reference does) ->
 * normals [B,3,H,W]: per pixel the least-squares plane n.p = 1 through its zero-padded k x k window,
 * n = (A^T A)^-1 A^T 1 (k odd, <= 31; the reference uses 31).  minv: NULL, or [B,6,H,W] to keep (A^T A)^-1 for the
 * backward pass.  A window whose A^T A is singular yields inf/nan (torch.linalg.inv raises there).
 * gbn_normals_backward: g_points [B,3,H,W] = d loss / d points given g_normals. */
int gbn_normals_forward(const float* points, int B, int H, int W, int k, float* normals, float* minv, void* stream);
int gbn_normals_backward(const float* points, const float* normals, const float* minv, const float* g_normals, int B, int H,
                         int W, int k, float* g_points, void* stream);

/* ---- NeRF_TCNN: DS_NeRF/run_nerf_helpers_tcnn.py:13-117 (hash-grid model built from tiny-cuda-nn modules) ------------
 * Parameters are the flat fp32 vectors the tiny-cuda-nn torch bindings expose: encoder.params
 * (gbn_tcnn_grid_params() floats: 16 levels x 2 features, level sizes min(round_up(res^3, 8), 2^19)), sigma_net.params
 * (64x32 + 16x64, row-major [out,in] per layer) and color_net.params (64x32 + 64x64 + 16x64).  gbn_tcnn_prepack
 * rounds them to fp16 into `table` (gbn_tcnn_table_bytes()); re-run after each optimizer step.
 * gbn_tcnn_forward evaluates NeRF_TCNN.forward (lines 90-117) for the points o_r + d_r * z[r,s] with direction
 * viewdirs_r (rays mode), or for explicit rows inputs [R*S, 6] = (point, direction) when `inputs` is non-NULL:
 * raw [R*S, 4] fp32 = (r, g, b, sigma) before activation, values rounded to fp16 as the reference's modules emit.
 * enc_stash: NULL, or [R*S, 32] fp16 to keep the hash-grid encodings for gbn_tcnn_backward. */
size_t gbn_tcnn_table_bytes(void);
size_t gbn_tcnn_grid_params(void);
int gbn_tcnn_prepack(const float* grid_params, const float* sigma_params, const float* color_params, void* table, void* stream);
int gbn_tcnn_forward(const void* table, const float* rays_o, const float* rays_d, const float* viewdirs, int64_t ray_stride,
                     const float* z_vals, const float* inputs, int64_t R, int S, float* raw, void* enc_stash, void* stream);

/* Backward of the same call wrt the parameters (inputs carry no gradient, run.py:2346).  enc_stash: the [R*S,32] fp16
 * hash-grid encodings gbn_tcnn_forward wrote when given a non-NULL enc_stash (64 B/point).  g_raw [R*S,4].
 * Accumulates (atomic adds) into g_grid (gbn_tcnn_grid_params() floats), g_sigma_params (3072) and g_color_params
 * (7168), all fp32 in the parameters' own flat layout; g_enc [R*S,32] fp32 is workspace (d loss / d encoding).
 * The MLP part runs in fp16 with fp32 accumulation; loss_scale (a power of two; tiny-cuda-nn's bindings use 128)
 * multiplies g_raw on the way in and is divided out before anything is written. */
int gbn_tcnn_backward(const void* table, const float* rays_o, const float* rays_d, const float* viewdirs, int64_t ray_stride,
                      const float* z_vals, const float* inputs, int64_t R, int S, const void* enc_stash, const float* g_raw,
                      float loss_scale, float* g_enc, float* g_grid, float* g_sigma_params, float* g_color_params,
                      void* stream);

/* Same network on pre-embedded rows (NeRF.forward's own signature): emb [P,90] fp32 -> raw [P,4]. */
int gbn_mlp_forward_embedded(const void* packed, int precision, const float* emb, int64_t P, float* raw,
                             void* workspace, void* stash, void* stream);

/* ---- MLP backward (autograd of NeRF.forward wrt its parameters; inputs carry no gradient, run.py:2346) --------
 * Step 1, data gradients: g_raw [P,4] -> the pre-activation gradient of every layer, written to `stash_g`
 * (same size and tile layout as the forward stash).  packed_bwd = gbn_mlp_prepack_weights(.., GBN_PACK_BWD_BF16). */
int gbn_mlp_backward_data(const void* packed_bwd, const float* g_raw, int64_t P, const void* stash_h,
                          void* stash_g, void* workspace, void* stream);

/* Step 2, parameter gradients: ACCUMULATES (atomic adds) into `grads`, a HOST array of 24 DEVICE pointers to fp32
 * buffers in nn.Linear layout, same order as gbn_mlp_prepack_weights' params.  g_raw [P,4] as in step 1;
 * viewdirs [R,3] (pitch ray_stride) are needed for the direction columns of views_linears.0; P = R*S.
 * workspace: gbn_mlp_wgrad_workspace_bytes(R). */
size_t gbn_mlp_wgrad_workspace_bytes(int64_t R);
int gbn_mlp_backward_weights(const void* stash_h, const void* stash_g, const float* g_raw, const float* viewdirs,
                             int64_t ray_stride, int64_t R, int S, void* const* grads, void* workspace,
                             void* stream);

/* ---- loss seed: img2mse terms of the training step (run.py:1483,1502,1513-1515) --------------------------
 * loss = mean((rgb-t)^2) + mean((rgb0-t)^2) + depth_lambda*mean((disp-td)^2), means over R_global*3 / R_global.
 * Writes the gradients wrt rgb, rgb0, disp and atomically accumulates the scalar loss into *loss. */
int gbn_loss_seed(const float* rgb, const float* rgb0, const float* disp, const float* target_rgb,
                  const float* target_disp, int64_t R, int64_t R_global, float depth_lambda, float* g_rgb,
                  float* g_rgb0, float* g_disp, float* loss, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GBNERF_H_ */
