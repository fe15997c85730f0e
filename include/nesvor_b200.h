/*
 * nesvor_b200.h -- C ABI of libnesvor_b200.so (sm_100a), the drop-in boundary for NeSVoR's
 * reconstruction hot path.
 *
 * Conventions (all entry points):
 *   - every pointer is DEVICE memory unless the parameter name starts with "h_";
 *   - inputs are borrowed, outputs are caller-allocated; entry points never allocate and never
 *     synchronise; work is enqueued on `stream` (a cudaStream_t passed as void*; NULL = legacy
 *     default stream, which is what the reference's extensions always use);
 *   - optional pointers may be NULL (the reference passes empty tensors for absent masks,
 *     nesvor/slice_acquisition/slice_acq.py:36-39);
 *   - return value: 0 on success, otherwise a cudaError_t (launch/config error) or a negative
 *     NSV_E* code for invalid arguments.  nsv_last_error_string() describes the last failure on
 *     the calling thread.
 *   - bool masks are 1 byte per element (torch.bool).
 *
 * Each block cites the reference interface it replaces (paths relative to /root/reference).
 */
#ifndef NESVOR_B200_H_
#define NESVOR_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NSV_OK 0
#define NSV_EINVAL (-1)      /* bad argument (NULL where required, non-positive size, ...) */
#define NSV_EUNSUPPORTED (-2) /* configuration outside what the kernels are instantiated for */

#define NSV_MAX_LEVELS 32

int nsv_version(void);                     /* ABI version, currently 1 */
const char* nsv_last_error_string(void);   /* thread-local, never NULL */
const char* nsv_build_arch(void);          /* "sm_100a" */

/* ------------------------------------------------------------------------------------------
 * Rigid-pose converters.  Replaces nesvor.transform_convert_cuda.{axisangle2mat_forward,
 * axisangle2mat_backward, mat2axisangle_forward, mat2axisangle_backward}
 * (nesvor/transform/transform_convert_cuda.cpp:27-69; kernels transform_convert_cuda_kernel.cu:15-440).
 * axisangle: [n,6] = (rx,ry,rz,tx,ty,tz); mat: [n,3,4] row-major [R|t].  Outputs fully written.
 * ------------------------------------------------------------------------------------------ */
int nsv_axisangle2mat_fwd_f32(const float* axisangle, float* mat, int n, void* stream);
int nsv_axisangle2mat_bwd_f32(const float* grad_mat, const float* axisangle, float* grad_axisangle, int n, void* stream);
int nsv_mat2axisangle_fwd_f32(const float* mat, float* axisangle, int n, void* stream);
int nsv_mat2axisangle_bwd_f32(const float* mat, const float* grad_axisangle, float* grad_mat, int n, void* stream);
/* transReg of the training loop and its gradient in one launch (replaces NeSVoR.trans_loss, nesvor/nesvor/models.py:357-363,
 * composed there from RigidTransform.inv / compose / axisangle under autograd): err = axisangle(T_init^-1 o T) per slice,
 * loss = mean(err_R^2) + 1e-3 mean(err_T^2).  ACCUMULATES: grad_axisangle[n,6] += weight * dloss/daxisangle, *loss += loss. */
int nsv_trans_reg_f32(const float* axisangle /* [n,6] */, const float* axisangle_init /* [n,6] */, float* grad_axisangle,
                      float* loss, int n, float weight, void* stream);
int nsv_axisangle2mat_fwd_f64(const double* axisangle, double* mat, int n, void* stream);
int nsv_axisangle2mat_bwd_f64(const double* grad_mat, const double* axisangle, double* grad_axisangle, int n, void* stream);
int nsv_mat2axisangle_fwd_f64(const double* mat, double* axisangle, int n, void* stream);
int nsv_mat2axisangle_bwd_f64(const double* mat, const double* grad_axisangle, double* grad_mat, int n, void* stream);

/* ------------------------------------------------------------------------------------------
 * Slice acquisition (PSF-weighted gather A, scatter A^T and their backward passes).  Replaces
 * nesvor.slice_acq_cuda.{forward, backward, adjoint_forward, adjoint_backward}
 * (nesvor/slice_acquisition/slice_acq_cuda.cpp:61-160; kernels slice_acq_cuda_kernel.cu:18-950).
 *   transforms [n,3,4] (voxel units), vol [D,H,W], psf [d_p,h_p,w_p], slices [n,h,w].
 * Unlike the reference's host wrappers, outputs are NOT zero-filled here when the op only writes
 * part of them: the caller passes zero-initialised `slices`, `slices_weight`, `grad_vol`,
 * `grad_transforms`, `vol`, `vol_weight`, `grad_slices` (torch.zeros), exactly as
 * slice_acq_cuda_kernel.cu:965-966,1005-1006,1040-1041,1109-1110 allocate them.
 * nsv_slice_acq_adjoint_forward runs the equalize pass itself when `equalize` != 0;
 * nsv_slice_acq_adjoint_backward first equalizes `grad_vol` IN PLACE when `equalize` != 0
 * (slice_acq_cuda_kernel.cu:1095-1107), as the reference does.
 * ------------------------------------------------------------------------------------------ */
int nsv_slice_acq_forward_f32(const float* transforms, const float* vol, const uint8_t* vol_mask,
                              const uint8_t* slices_mask, const float* psf, float* slices, float* slices_weight,
                              int D, int H, int W, int d_p, int h_p, int w_p, int n, int h, int w,
                              float res_slice, int interp_psf, void* stream);
int nsv_slice_acq_backward_f32(const float* transforms, const float* vol, const uint8_t* vol_mask, const float* psf,
                               const float* grad_slices, const uint8_t* slices_mask, float* grad_vol,
                               float* grad_transforms, int D, int H, int W, int d_p, int h_p, int w_p, int n, int h,
                               int w, float res_slice, int interp_psf, void* stream);
int nsv_slice_acq_adjoint_forward_f32(const float* transforms, const float* psf, const float* slices,
                                      const uint8_t* slices_mask, const uint8_t* vol_mask, float* vol,
                                      float* vol_weight, int D, int H, int W, int d_p, int h_p, int w_p, int n, int h,
                                      int w, float res_slice, int interp_psf, int equalize, void* stream);
int nsv_slice_acq_adjoint_backward_f32(const float* transforms, float* grad_vol, const float* vol_weight,
                                       const uint8_t* vol_mask, const float* psf, const float* slices,
                                       const uint8_t* slices_mask, const float* vol, float* grad_slices,
                                       float* grad_transforms, int D, int H, int W, int d_p, int h_p, int w_p, int n,
                                       int h, int w, float res_slice, int interp_psf, int equalize, void* stream);
/* Flavour switch of the fp32 slice-acquisition family.  0 (default; or env NSV_SLICE_ACQ_EXACT unset): the fast product
 * kernels (csrc/slice_acq_fast.cu: FMA contraction, inside / outside pixel classification, per-slice rotated tap table,
 * axis-matched volume layouts, neighbour-merged reductions) -- equal to the reference within fp32 round-off.
 * 1: the bit-exact flavour (csrc/slice_acq.cu, -fmad=false) whose gather passes reproduce the reference's C arithmetic
 * (slice_acq_cuda_kernel.cu:18-171, :696-950 as compiled for the CPU) bit for bit -- verification mode.
 * nsv_set_slice_acq_tuning: bit mask of the fast flavour's optimisations (1 layouts, 2 row warps, 4 neighbour merge,
 * 8 zero-pixel skip in A^T, 16 pixel classification); profiling / test hook, default all on. */
void nsv_set_slice_acq_exact(int exact);
int nsv_get_slice_acq_exact(void);
void nsv_set_slice_acq_tuning(unsigned bits);
unsigned nsv_get_slice_acq_tuning(void);
int nsv_equalize_f32(float* vol, const float* vol_weight, int is_grad, int64_t DHW, void* stream);

int nsv_slice_acq_forward_f64(const double* transforms, const double* vol, const uint8_t* vol_mask,
                              const uint8_t* slices_mask, const double* psf, double* slices, double* slices_weight,
                              int D, int H, int W, int d_p, int h_p, int w_p, int n, int h, int w,
                              double res_slice, int interp_psf, void* stream);
int nsv_slice_acq_backward_f64(const double* transforms, const double* vol, const uint8_t* vol_mask, const double* psf,
                               const double* grad_slices, const uint8_t* slices_mask, double* grad_vol,
                               double* grad_transforms, int D, int H, int W, int d_p, int h_p, int w_p, int n, int h,
                               int w, double res_slice, int interp_psf, void* stream);
int nsv_slice_acq_adjoint_forward_f64(const double* transforms, const double* psf, const double* slices,
                                      const uint8_t* slices_mask, const uint8_t* vol_mask, double* vol,
                                      double* vol_weight, int D, int H, int W, int d_p, int h_p, int w_p, int n, int h,
                                      int w, double res_slice, int interp_psf, int equalize, void* stream);
int nsv_slice_acq_adjoint_backward_f64(const double* transforms, double* grad_vol, const double* vol_weight,
                                       const uint8_t* vol_mask, const double* psf, const double* slices,
                                       const uint8_t* slices_mask, const double* vol, double* grad_slices,
                                       double* grad_transforms, int D, int H, int W, int d_p, int h_p, int w_p, int n,
                                       int h, int w, double res_slice, int interp_psf, int equalize, void* stream);
int nsv_equalize_f64(double* vol, const double* vol_weight, int is_grad, int64_t DHW, void* stream);

/* ------------------------------------------------------------------------------------------
 * Multiresolution hash-grid encoding.  Replaces tcnn.Encoding(HashGrid) as used through
 * build_encoding (nesvor/nesvor/models.py:22-25, built :102-111, called :146); tiny-cuda-nn is an
 * external dependency of the reference (README.md:88), semantics per SURVEY.md App. A.
 * ------------------------------------------------------------------------------------------ */
typedef struct nsv_grid_meta {
  int32_t n_levels;
  int32_t n_features;                  /* F: 1, 2, 4 or 8 */
  float scale[NSV_MAX_LEVELS];         /* base * s^l - 1 (fp32) */
  uint32_t res[NSV_MAX_LEVELS];        /* ceil(scale)+1 */
  uint32_t size[NSV_MAX_LEVELS];       /* entries in the level, T_l */
  uint32_t offset[NSV_MAX_LEVELS + 1]; /* first entry of the level in the flat table */
  uint32_t hashed[NSV_MAX_LEVELS];     /* 1: XOR-prime hash, 0: dense */
} nsv_grid_meta;

/* host helper: fills `meta` from the tcnn-style hyper-parameters; returns total entry count */
int64_t nsv_grid_meta_init(nsv_grid_meta* h_meta, int n_levels, int n_features, int log2_hashmap_size,
                           int base_resolution, float per_level_scale);

/* x [N,3] fp32 in [0,1] (unclamped); table: flat [sum T_l, F] (fp32 or fp16 copy);
 * out [N, L*F] level-major (fp32 or fp16).  `dy_dx` (optional, [N, L*F, 3] fp32) receives
 * d out / d x for nsv_hashgrid_bwd_input. */
int nsv_hashgrid_fwd_f32(const float* x, const float* table, const nsv_grid_meta* h_meta, float* out, int64_t N, void* stream);
int nsv_hashgrid_fwd_f16(const float* x, const void* table_f16, const nsv_grid_meta* h_meta, void* out_f16, int64_t N, void* stream);
/* grad_table (fp32, flat, same layout as table) += scatter of grad_out; caller zero-fills */
int nsv_hashgrid_bwd_params_f32(const float* x, const float* grad_out, const nsv_grid_meta* h_meta, float* grad_table, int64_t N, void* stream);
int nsv_hashgrid_bwd_params_f16(const float* x, const void* grad_out_f16, const nsv_grid_meta* h_meta, float* grad_table, float grad_scale, int64_t N, void* stream);
/* grad_x [N,3] = sum_l d out_l / d x * grad_out_l (recomputed from the table, nothing stored) */
int nsv_hashgrid_bwd_input_f32(const float* x, const float* table, const float* grad_out, const nsv_grid_meta* h_meta, float* grad_x, int64_t N, void* stream);
int nsv_hashgrid_bwd_input_f16(const float* x, const void* table_f16, const void* grad_out_f16, const nsv_grid_meta* h_meta, float grad_scale, float* grad_x, int64_t N, void* stream);

/* ------------------------------------------------------------------------------------------
 * Fully fused small MLP (fp16 operands, fp32 accumulate, ReLU hidden, linear output, no biases).
 * Replaces tcnn.Network(CutlassMLP) as used through build_network's fp16 branch
 * (nesvor/nesvor/models.py:28-41; instances :113-121, :238-258).
 * weights: fp16, layers concatenated, layer i row-major [out_i, in_i]; n_in / n_out padded to 16,
 * width in {16, 32, 64, 128}; n_hidden >= 1.
 * ------------------------------------------------------------------------------------------ */
int nsv_mlp_fwd_f16(const void* x_f16, const void* weights_f16, void* out_f16, void* hidden_f16 /* [n_hidden, N, width] or NULL */,
                    int64_t N, int n_in, int n_out, int width, int n_hidden, void* stream);
int nsv_mlp_bwd_f16(const void* x_f16, const void* weights_f16, const void* hidden_f16, const void* grad_out_f16,
                    void* grad_x_f16 /* or NULL */, float* grad_weights /* fp32, same layout, caller zero-fills */,
                    int64_t N, int n_in, int n_out, int width, int n_hidden, void* stream);

/* ------------------------------------------------------------------------------------------
 * Kernel A: one fused NeSVoR training iteration (forward + loss + backward) over a batch of B
 * slice pixels x S PSF samples, and the forward-only renderer.
 * Replaces the op sequence NeSVoR.forward -> net_forward -> INR.forward -> losses -> autograd
 * backward (nesvor/nesvor/models.py:260-384, called from train.py:183-190) and
 * INR.sample_batch + INR.forward(...).mean(-1) (models.py:142-174, sample.py:17-53).
 * ------------------------------------------------------------------------------------------ */
typedef struct nsv_inr_config {
  nsv_grid_meta grid;
  int32_t width;            /* hidden width W: 32 or 64 */
  int32_t depth;            /* hidden layers per MLP (args.depth), 1..4 */
  int32_t n_features_z;     /* 15 */
  int32_t n_features_slice; /* 16 (0 disables the embedding) */
  int32_t n_levels_bias;    /* 0 disables b_net */
  int32_t pixel_variance;   /* sigma_net on */
  int32_t slice_variance;
  int32_t slice_scale;
  int32_t pose_grad;        /* back-propagate into axisangle */
  int32_t image_reg;        /* 0 none, 1 TV, 2 edge, 3 L2 */
  float delta;              /* args.delta * v_mean */
  float w_image;            /* loss weights (train.py:167-173); MSE and logVar have weight 1 */
  float w_bias;
  float bbox_lo[3], bbox_hi[3];
  float grad_scale;         /* power of two applied to fp16 backward operands (loss scale) */
} nsv_inr_config;

typedef struct nsv_inr_params {   /* device pointers */
  const void* table_f16;          /* [sum T_l, F] fp16 copy of the master table */
  const void* mlp_f16;            /* packed fp16 weights, see nsv_inr_mlp_layout() */
  const float* axisangle;         /* [n_slices, 6] */
  const float* psf_sigma;         /* [n_slices, 3] */
  const float* slice_embedding;   /* [n_slices, n_features_slice] or NULL */
  const float* logit_coef;        /* [n_slices] or NULL */
  const float* log_var_slice;     /* [n_slices] or NULL */
  int32_t n_slices;
} nsv_inr_params;

typedef struct nsv_inr_grads {    /* device pointers, all fp32, caller zero-fills before the call */
  float* table;                   /* same layout as the table */
  float* mlp;                     /* same layout as mlp_f16 (element for element) */
  float* axisangle;               /* [n_slices, 6] or NULL */
  float* slice_embedding;
  float* slice_scale_c;           /* [n_slices]: receives dL/dlogit_coef (softmax chain rule applied by the finalize kernel) */
  float* log_var_slice;
  float* losses;                  /* [8]: [0] MSE, [1] logVar, [2] biasReg, [3] imageReg (final values);
                                   * [4] INPUT when n_levels_bias > 0: mean(log_bias) over the whole batch, see nsv_inr_bias_mean;
                                   * [5] transReg (written by nsv_trans_reg_f32 when the caller points it there); [6] = [0] + [1],
                                   * the sum the reference logs as "MSE+logVar" (models.py:317-319); [7] unused */
} nsv_inr_grads;

/* number of fp16 elements of the packed MLP buffer and per-net offsets (host helper) */
int64_t nsv_inr_mlp_layout(const nsv_inr_config* h_cfg, int64_t* h_offsets /* [3]: density, sigma, bias */);

/* which implementation of kernel A nsv_inr_train_step uses: 0 = auto (tcgen05/TMEM when instantiated for the
 * configuration -- the warp-specialised kernel for the sigma_net heads, the all-phases kernel otherwise -- else mma.sync),
 * 1 = mma.sync fragments, 2 = tcgen05/TMEM, all warps run every phase, 3 = tcgen05/TMEM warp-specialised (memory
 * warps + chain warps); 2 and 3 return NSV_EUNSUPPORTED when not instantiated for the configuration */
int nsv_set_fused_impl(int impl);
/* tuning / test hooks of kernel A's gather and scatter loops (no reference counterpart; results are identical up to
 * float-atomic ordering): `agg_max_entries` = largest dense level whose gradient is pre-reduced inside a warp before
 * touching global memory (0 = never, < 0 = default: every dense level, or $NSV_AGG_MAX); `fast_path` = 0 forces the generic
 * per-level loops, 1 the chunked branch-free loops, < 0 = default (1 or $NSV_FAST_PATH). */
/* Verification hook: the PSF-sample generator every fused kernel uses when no noise tensor is passed (replaces
 * torch.randn(B, S, 3), nesvor/nesvor/models.py:269 and :161): sample index idx = offset + i -> Philox4x32-10 with counter
 * (idx_lo, idx_hi, 0, 0) and key (seed_lo, seed_hi) -> Box-Muller.  normals [n,3] f32 and / or raw [n,4] u32 (either may be
 * NULL). */
int nsv_debug_normal3(uint64_t seed, uint64_t offset, int64_t n, float* normals, uint32_t* raw, void* stream);
/* 128-sample groups per CTA of the tcgen05 training kernel: 2 (512 threads at 128 registers) or 3 (768 threads at 80
 * registers, one shared set of TMEM weight-gradient accumulators; only for density-only heads with n_samples <= 128 --
 * other configurations keep 2); -1 back to the default (env NSV_TC_GROUPS).  Results are identical up to the order of
 * floating-point accumulation. */
int nsv_set_fused_tc_groups(int groups);
/* Tile order of the tcgen05 training kernel: 0 (default; env NSV_TILE_ORDER) tiles strided over the CTAs, 1 one contiguous
 * run of tiles per CTA -- with a spatially ordered batch (Dataset locality_batch_size) a CTA then walks neighbouring pixels
 * of one slice; -1 back to the default.  Results are identical either way (the losses are means over the batch). */
int nsv_set_fused_tile_order(int contiguous);
/* Number of leading dense levels of the fp16 table that the tcgen05 training kernel stages into shared memory with one
 * bulk copy per CTA (TMA engine, cp.async.bulk) and gathers from there: -1 as many as fit beside the operand tiles
 * (default; env NSV_SMEM_LEVELS), 0 none, -2 back to the default.  Profiling / test hook. */
int nsv_set_fused_smem_levels(int levels);
int nsv_set_fused_tuning(int64_t agg_max_entries, int fast_path);
/* profiling hook: 16 int64 device counters that the tcgen05 kernel A (config-2 instantiation) fills with per-phase
 * warp cycles (gather, barriers, MMA wait, epilogues, losses, scatter, pixel barrier, -); NULL switches it off */
int nsv_set_fused_timers(void* device_counters);

int nsv_inr_train_step(const nsv_inr_config* h_cfg, const nsv_inr_params* h_params, const nsv_inr_grads* h_grads,
                       const float* xyz /* [B,3] */, const float* v /* [B] */, const int64_t* slice_idx /* [B] */,
                       const float* noise /* [B,S,3] or NULL -> in-kernel Philox(seed, offset) */,
                       uint64_t seed, uint64_t offset, float* v_out /* [B] or NULL */, int64_t B, int S, void* stream);

/* Bias-field head (b_net, nesvor/nesvor/models.py:247-258,344-347; n_levels_bias > 0).  biasReg = mean(log_bias)^2
 * (models.py:323) couples every sample of the batch, so its mean must exist before kernel A back-propagates: this
 * forward-only pre-pass evaluates b_net on exactly the samples nsv_inr_train_step will draw (same xyz / slice_idx / noise
 * or Philox(seed, offset)) and ADDS mean(log_bias) to *out_mean (device float, caller zero-fills; normally
 * grads.losses + 4).  Data-parallel callers average it over ranks between the two calls.  nsv_inr_train_step then reads
 * grads.losses[4], applies w_bias * d(mean^2) in its backward pass and reports biasReg in losses[2].
 * Instantiated for 1 <= n_levels_bias <= 4, F = 2, width 64, depth 1, n_features_slice 16 (else NSV_EUNSUPPORTED). */
int nsv_inr_bias_mean(const nsv_inr_config* h_cfg, const nsv_inr_params* h_params, const float* xyz /* [B,3] */,
                      const int64_t* slice_idx /* [B] */, const float* noise /* [B,S,3] or NULL */, uint64_t seed, uint64_t offset,
                      float* out_mean, int64_t B, int S, void* stream);

int nsv_inr_render(const nsv_inr_config* h_cfg, const nsv_inr_params* h_params,
                   const float* xyz /* [M,3] */, const float* mat /* [M,3,4] per-point or [1,3,4] or NULL */, int mat_per_point,
                   const float* psf_sigma /* [M,3] per-point or [3] */, int sigma_per_point,
                   const float* noise /* [M,S,3] or NULL */, uint64_t seed, uint64_t offset,
                   float* out /* [M] mean density */, int64_t M, int S, void* stream);

/* ------------------------------------------------------------------------------------------
 * Optimiser: fused AdamW over a flat fp32 segment, optionally refreshing an fp16 copy.
 * Replaces torch.optim.AdamW as configured in nesvor/nesvor/train.py:134-152 (+ GradScaler
 * unscale, train.py:162-164,195).
 * ------------------------------------------------------------------------------------------ */
int nsv_adamw_step(float* param, float* grad, float* exp_avg, float* exp_avg_sq, void* param_f16 /* or NULL */,
                   int64_t n, float lr, float beta1, float beta2, float eps, float weight_decay, int step,
                   float grad_unscale, int zero_grad /* clear grad for the next iteration */, void* stream);

/* Data-parallel optimiser step (the reference's DDP-less loop has no counterpart; this replaces
 * all-reduce + nsv_adamw_step): reduce-scatter + AdamW + all-gather in one kernel over NVLink peer memory.  `peer_grads`
 * and `peer_param_f16` are HOST arrays of `world` device pointers to every rank's gradient / fp16 parameter buffer
 * (symmetric allocations, peer access enabled); rank `rank` updates the shard given by nsv_adamw_shard_bounds of
 * param / exp_avg / exp_avg_sq (full-size local buffers) and stores the refreshed fp16 values into every rank's copy.
 * Elements [f32_lo, n) -- the per-slice parameters (slice embedding, slice scale / variance, poses), which the training
 * kernel reads in fp32 -- are additionally mirrored in fp32 into every rank's `peer_param_f32[r]` (element e at index
 * e - f32_lo), because only the owner's fp32 master is current.
 * The caller synchronises the ranks before (gradients complete) and after (copies written) and clears its gradient. */
int nsv_adamw_shard_bounds(int64_t n, int world, int rank, int64_t* lo, int64_t* hi);
int nsv_adamw_step_dp(float* param, const void* const* peer_grads, float* exp_avg, float* exp_avg_sq,
                      void* const* peer_param_f16, int world, int rank, int64_t n, float lr, float beta1, float beta2,
                      float eps, float weight_decay, int step, float grad_unscale,
                      int64_t f32_lo /* multiple of 4 */, void* const* peer_param_f32 /* or NULL */, void* stream);
/* The same step with the NVSwitch doing the sum and the replication (NVLS): `mc_grad`, `mc_param_f16`, `mc_param_f32` are the
 * MULTICAST addresses of the three peer buffers (one address bound to the copy on every rank, e.g. the `multicast_ptr` of a
 * torch symmetric-memory allocation plus the buffer's offset).  Each owner reads its shard's gradient sum with
 * multimem.ld_reduce (one shard of inbound link traffic per rank and step instead of world - 1) and writes the refreshed
 * parameters with multimem.st (one shard outbound instead of world - 1).  The switch's summation order is not the rank order of
 * nsv_adamw_step_dp, so the two differ by fp32 round-off; replicas still agree bit for bit with each other (one owner per
 * element).  The unicast peer pointers are still needed (ragged tail of the last shard).  Same barriers around it. */
int nsv_adamw_step_dp_mc(float* param, const void* const* peer_grads, float* exp_avg, float* exp_avg_sq,
                         void* const* peer_param_f16, int world, int rank, int64_t n, float lr, float beta1, float beta2,
                         float eps, float weight_decay, int step, float grad_unscale, int64_t f32_lo,
                         void* const* peer_param_f32, const void* mc_grad, void* mc_param_f16, void* mc_param_f32 /* or NULL */,
                         void* stream);
/* The same kernel with the rendezvous of the ranks INSIDE it (no host-launched barrier around it): `peer_flags[r]` = rank r's
 * flag block (64 x uint64 in peer-accessible memory, zero-initialised once and made visible to all ranks before the first
 * call), `epoch` = a positive number that is the same on every rank for a given step and strictly increases from step to
 * step.  Prologue: every rank announces "my gradient is complete" (it is: the producing kernel precedes this launch on
 * `stream`) to every rank and waits for all announcements before pulling gradients; epilogue: the rank's last block announces
 * "all my reads and my writes into your copies are done" and waits for the same from every rank, so that when the kernel
 * completes the local fp16 / fp32 copies are whole and the local gradient may be cleared.  A wait gives up after a few
 * seconds (a dead rank must not hang the others' GPUs) and records the epoch in flag word 33.
 * sync_mode: 1 = only the rendezvous BEFORE (the caller still synchronises the ranks after the kernel), 2 = only the one AFTER,
 * 3 = both. */
int nsv_adamw_step_dp_sync(float* param, const void* const* peer_grads, float* exp_avg, float* exp_avg_sq,
                           void* const* peer_param_f16, int world, int rank, int64_t n, float lr, float beta1, float beta2,
                           float eps, float weight_decay, int step, float grad_unscale, int64_t f32_lo,
                           void* const* peer_param_f32, void* const* peer_flags, uint64_t epoch, int sync_mode, void* stream);


/* tcgen05 / TMEM bring-up check used by tests/test_gpu_umma.py: runs every tensor-core operand
 * configuration kernel A uses on fixed 128x64 / 64x64 / 128x16 fp16 inputs (no reference counterpart). */
int nsv_umma_selftest(const void* A_f16, const void* W_f16, const void* G_f16, float* out, void* stream);
/* Hardware-behaviour probe behind the 3-group training kernel: `n_issuers` (1..4) threads of different warps each issue
 * `reps` accumulating M64 N64 K128 products A^T A into ONE TMEM accumulator, unordered; out [64,64] f32 must be
 * n_issuers * reps * A^T A if accumulating tcgen05.mma instructions of different issuers compose. */
int nsv_umma_shared_accumulator_test(const void* A_f16, float* out, int n_issuers, int reps, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NESVOR_B200_H_ */
