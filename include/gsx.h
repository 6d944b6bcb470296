/* libgsx -- C ABI of the B200-native generate hot path (StyleGAN-v1 synthesis + segmentation decoder).
 *
 * The reference (author-hidden-name/GAN-segmentation) is pure Python over MXNet and defines no FFI;
 * its boundary for this path is the Python surface of Generator / Decoder / ImageGenerator / SegSolver.
 * Each entry point below states the reference interface it replaces (file:line relative to the
 * reference tree).  INTEGRATION.md shows the ctypes binding a maintainer would add.
 *
 * Conventions: every function returns 0 on success, <0 on error (gsx_last_error() returns a
 * thread-local message).  Handles are opaque, one per GPU (the current CUDA device at create time),
 * not thread-safe per handle.  Device pointers and workspaces are caller-owned (torch tensors on the
 * Python side); forward calls only enqueue work on the caller's stream and never synchronise.
 * There is no CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef GSX_H_
#define GSX_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gsx_synth gsx_synth;
typedef struct gsx_dec gsx_dec;
typedef void* gsx_stream;                 /* cudaStream_t */

/* Generator hyper-parameters: the keys of ImageGenerator._get_config (image_generator.py:46-74)
 * that Generator.__init__ reads (networks_stylegan.py:81-91). */
typedef struct {
  int max_res_log2;
  int base_scale_y, base_scale_x;
  int fmap_base;
  float fmap_decay;
  int fmap_max;
  int latent_size;
  int channels;
} gsx_synth_cfg;

/* Decoder hyper-parameters: cfg['in_channels'], cfg['features'] (incl. the trailing num_classes),
 * cfg['use_bn'] of SegSolver.get_config (seg_solver.py:83-132); base = spatial size of level 0. */
typedef struct {
  int num_levels;
  int in_channels[16];
  int features[17];
  int use_bn;
  int base_y, base_x;
} gsx_dec_cfg;

typedef struct {
  int TH, TW, NB, CBK, N_tile, stages, phase_grid;     /* 0 (phase_grid: -1) keeps the planner's choice */
  int epi_groups, acc_bufs, max_mtiles;
  int hstack;                                         /* must be <= 0: the tap-stacked variant is not built */
  int s2d;                                            /* -1 keeps the planner's choice, 0/1 force (thin 3x3 layers) */
} gsx_plan_override;

const char* gsx_last_error(void);
int gsx_abi_version(void);
/* Number of kernels this library has launched in the calling process (all handles); bench.py reports it. */
uint64_t gsx_launch_count(void);
/* Process-wide build-time options, read when a handle is finalized (gsx_*_finalize).
 *   "fold_apply" (default 1): fold the AdaIN normalise+modulate step of the channel-thin generator blocks into the convs
 *                 that consume it (per-sample modulated weights) instead of a separate pass over the tensor; 0 keeps the
 *                 reference's operator order (networks_stylegan.py:56-73) everywhere -- for A/B tests.
 *   "fold_deconv_maxc" (default 16): the transposed conv + blur + noise/bias/lrelu/statistics of a block run as ONE kernel
 *                 when the block has at most this many channels (0: never).
 *   "pdl" (default 3): programmatic dependent launch -- 0 off, 1 small passes + conv chains (round 1), 2 always with the trigger
 *                 at kernel start, 3 always with the big kernels triggering when their CTAs finish, 4 = policy of 1 with late triggers.
 *   "varn" (default 1): stacked-phase up-convs issue MMAs only over the phase blocks a shift feeds (1: layers with >= 32
 *                 output channels, 2: all, 0: off);  "epi_groups" (default 2; 4 = experiment): epilogue warps per TMEM lane quarter
 *                 of the kernels without the generator epilogue.
 * Read at every call:
 *   "dec_branches" (default 1): gsx_dec_forward runs the cvt block of every level and the 1x1 shortcuts on side streams
 *                 beside the main chain conv_a -> conv_b (joined by events; capturable); "dec_branch_kpx": only for levels
 *                 with at most that many thousand pixels in the batch (default: all).
 *   "wgrad_kxm" (default 1): weight gradients of layers with <= 16 input channels issue one MMA per (ky, channel block) with the
 *                 kx taps along M (6 instead of 9 per 16 pixels).
 *   "wgrad_m64" (default 2): weight gradients of layers with <= 64 input channels use M = 64 MMAs (0: M = 128 everywhere).
 *   "inline_finalize" (default 1): blocks that do not fold their AdaIN compute its coefficients in the apply pass itself
 *                 (no separate finalize launch; same summation order, bit-identical).
 *   "defer_rgb" (default 1): the fused generate calls run the ToRGB / image pass beside the decoder. */
int gsx_set_option(const char* name, int value);

/* ---- generator: replaces Generator(config) / load_parameters / __call__
 *      (networks_stylegan.py:76-197, image_generator.py:20-22, :99) ---- */
int gsx_synth_create(const gsx_synth_cfg* cfg, gsx_synth** out);
void gsx_synth_destroy(gsx_synth* h);
/* Parameter by reference name (structural 'net4.block2.0.weight' or legacy '16_conv_2_weight'),
 * float32 host data in the reference's shape.  Unknown names are ignored (ignore_extra=True,
 * image_generator.py:22) and reported with return value 1. */
int gsx_synth_set_param(gsx_synth* h, const char* name, const float* data, const int64_t* shape, int ndim);
/* Folds wscale/lr_mult, packs bf16 tensor-core operands, uploads.  Fails if a parameter is missing. */
int gsx_synth_finalize(gsx_synth* h);
/* CUDA-graph replays: with the counter enabled, the Philox latents / noise of gsx_synth_forward take their first global
 * sample index from a device-resident counter (initialised to `start`, advanced by n at every forward) instead of the
 * `first_sample` argument, so that a captured forward generates fresh samples on every replay.  (The reference draws
 * fresh noise per call, networks_stylegan.py:300.) */
int gsx_synth_device_counter(gsx_synth* h, int enable, uint64_t start);

int gsx_synth_workspace_bytes(const gsx_synth* h, int n, size_t* bytes);
int gsx_synth_num_layers(const gsx_synth* h);                       /* 2*(max_res_log2-1) */
int gsx_synth_feature_shape(const gsx_synth* h, int level, int* c, int* hgt, int* wid);
/* z_dev [n,latent] fp32 or NULL (Philox(seed, first_sample+i)).  psi_host: NULL -> truncation_psi
 * parameter, else num_layers floats.  noise_dev: NULL -> Philox, else num_layers device pointers to
 * [n,1,h,w] fp32 (the planes AddNoise samples at networks_stylegan.py:300).  Outputs (each nullable):
 * img_f32_dev [n,3,H,W], img_u8_dev [n,H,W,3] (image_generator.py:76-84), feats_f32_dev[level] [n,C,h,w].
 * The blocked 16-bit features stay in the workspace for gsx_dec_forward. */
int gsx_synth_forward(gsx_synth* h, int n, const float* z_dev, const float* psi_host,
                      const float* const* noise_dev, uint64_t seed, uint64_t first_sample, float* img_f32_dev,
                      uint8_t* img_u8_dev, float* const* feats_f32_dev, void* ws, size_t ws_bytes,
                      gsx_stream stream);
/* Copies the noise plane / latents the last forward used (for parity dumps). */
int gsx_synth_export_noise(gsx_synth* h, int n, int layer, float* out_dev, const void* ws, gsx_stream stream);
int gsx_synth_export_latents(gsx_synth* h, int n, float* out_dev, const void* ws, gsx_stream stream);

/* ---- decoder: replaces Decoder(cfg)(*features) + argmax of SegSolver.predict
 *      (networks_seg.py:49-113, seg_solver.py:307-329) ---- */
int gsx_dec_create(const gsx_dec_cfg* cfg, gsx_dec** out);
void gsx_dec_destroy(gsx_dec* h);
int gsx_dec_set_param(gsx_dec* h, const char* name, const float* data, const int64_t* shape, int ndim);
int gsx_dec_finalize(gsx_dec* h);                                   /* folds BatchNorm (inference) */
int gsx_dec_workspace_bytes(const gsx_dec* h, int n, size_t* bytes);
/* Features either as fp32 NCHW device arrays (feats_f32_dev[level], the reference's predict() input)
 * or, when feats_f32_dev is NULL, taken from the workspace of the last gsx_synth_forward(synth, n, ...).
 * mask_dev [n,H,W] uint8 class ids (first maximum wins); logits_dev [n,classes,H,W] fp32 or NULL. */
int gsx_dec_forward(gsx_dec* h, int n, const float* const* feats_f32_dev, const gsx_synth* synth,
                    const void* synth_ws, float* logits_dev, uint8_t* mask_dev, void* ws, size_t ws_bytes,
                    gsx_stream stream);

/* ---- whole generate step on DEVICE buffers (main.py:97-99 per batch, nothing leaves HBM): z_dev [n,latent] (or NULL: Philox
 *      latents of (seed, first_sample)) -> img_u8_dev [n,H,W,3] + mask_dev [n,H,W].  Same results as gsx_synth_forward followed by
 *      gsx_dec_forward(synth = s); fused so that the image pass (ToRGB) runs beside the decoder, which does not need it.  Both
 *      outputs are complete in `stream` order when the call returns. ---- */
int gsx_generate_dev(gsx_synth* s, gsx_dec* d, int n, const float* z_dev, const float* psi_host, uint64_t seed,
                     uint64_t first_sample, uint8_t* img_u8_dev, uint8_t* mask_dev, void* synth_ws, size_t synth_ws_bytes,
                     void* dec_ws, size_t dec_ws_bytes, gsx_stream stream);

/* ---- whole generate step with HOST buffers (main.py:97-99 per batch): z_host (or NULL) in,
 *      uint8 image + uint8 mask out; H2D/D2H copies are enqueued inside the call.
 *      copy_stream == NULL: everything on `stream`, stage_dev >= n*(latent*4 + H*W*4) (+3 KiB) bytes.
 *      copy_stream != NULL: the device-to-host copies go to copy_stream and overlap the next call's kernels;
 *      consecutive calls alternate slot 0/1 (two staging slots: stage_dev twice as large, two sets of host
 *      buffers); outputs of a call are complete once copy_stream has drained (or the same slot is reused). ---- */
int gsx_generate_host(gsx_synth* s, gsx_dec* d, int n, const float* z_host, const float* psi_host, uint64_t seed,
                      uint64_t first_sample, uint8_t* img_u8_host, uint8_t* mask_host, void* synth_ws,
                      size_t synth_ws_bytes, void* dec_ws, size_t dec_ws_bytes, void* stage_dev,
                      size_t stage_bytes, gsx_stream stream, gsx_stream copy_stream, int slot);

/* ---- decoder-training building blocks (seg_solver.py:351-465): loss, weight gradient, BatchNorm / upsample kernels,
 *      optimizer step (the conv forward / data gradient go through gsx_op_conv).
 * gsx_softmax_ce: SoftmaxCELoss(axis=1) with sample_weight = (label > -1) (seg_solver.py:404-407).
 *   logits [n,classes,h,w] fp32, labels [n,h,w] int32 (-1 = ignore) -> loss_dev[n] (mean over all h*w pixels) and,
 *   if dlogits_dev != NULL, grad_scale * d(sum_n loss_n)/dlogits.  grad_scale = h*w keeps the gradient O(1) through
 *   the 16-bit backward pass (1/(h*w) is an fp16 sub-normal at 1024^2); fold 1/grad_scale into gsx_adam_step's
 *   rescale_grad.  scratch_dev: n*256 floats.
 * gsx_adam_step: MXNet Adam on one flat fp32 bucket after the single gradient all-reduce (seg_solver.py:56,421):
 *   lr_t = lr*sqrt(1-beta2^t)/(1-beta1^t), g' = g*rescale_grad + wd*w. ---- */
/* Weight / bias gradient of a stride-1 'same' k x k conv (k = 1 or 3), fp32 NCHW in (converted to the blocked 16-bit
 * layout internally), dw [cout][cin][k][k] and db [cout] (may be NULL) out; deterministic.  Tuning / test hook of the
 * first, CUDA-core version of the decoder's weight gradient (seg_solver.py:411-412 err.backward()). */
int gsx_op_conv_wgrad(int k, int n, int h, int w, int cin, int cout, const float* x_dev, const float* dy_dev,
                      float* dw_dev, float* db_dev, gsx_stream stream);
/* The same weight gradient on tcgen05 tensor cores (csrc/wgrad.cu): a split-K GEMM over the pixels with the blocked
 * activation layout read as MN-major operands, one partial per CTA, fixed-order reduction.  cin % 8 == 0, cout <= 56. */
int gsx_op_conv_wgrad_tc(int k, int n, int h, int w, int cin, int cout, const float* x_dev, const float* dy_dev,
                         float* dw_dev, gsx_stream stream);
/* Train-mode BatchNorm (batch statistics, eps 1e-5) + LeakyReLU(0.2) (+ Dropout(0.5) mask) forward / backward and the
 * nearest-x2 upsample / its adjoint, fp32 NCHW (networks_seg.py:14-29, 70-78, 87).  stats_dev [3][C] = mean, biased
 * variance, rstd; dparam_dev [2][C] = dbeta, dgamma.  Deterministic (fixed-order double partial sums). */
/* The gsx_op_* hooks keep their scratch device memory between calls; this returns it. */
void gsx_op_release_cache(void);
int gsx_op_upsample2(const float* x_dev, float* y_dev, int n, int c, int h, int w, gsx_stream stream);
int gsx_op_sumpool2(const float* dy_dev, float* dx_dev, int n, int c, int h, int w, gsx_stream stream);
int gsx_op_bn_lrelu_fwd(const float* z_dev, const float* gamma_dev, const float* beta_dev, const float* drop_dev, float* y_dev,
                        float* stats_dev, int n, int c, int hw, gsx_stream stream);
int gsx_op_bn_lrelu_bwd(const float* dy_dev, const float* z_dev, const float* stats_dev, const float* gamma_dev,
                        const float* beta_dev, const float* drop_dev, float* dz_dev, float* dparam_dev, int n, int c, int hw,
                        gsx_stream stream);
int gsx_softmax_ce(const float* logits_dev, const int* labels_dev, int n, int num_classes, int h, int w,
                   float* loss_dev, float* dlogits_dev, float grad_scale, float* scratch_dev, size_t scratch_floats,
                   gsx_stream stream);
int gsx_adam_step(float* w_dev, const float* g_dev, float* m_dev, float* v_dev, size_t count, int t, float lr, float beta1,
                  float beta2, float eps, float wd, float rescale_grad, gsx_stream stream);

/* ---- one decoder-training iteration behind one call (seg_solver.py:386-421; csrc/train_step.cu).
 * gsx_train_create: decoder config (use_bn must be 1), the per-rank batch n, Dropout(0.5) after the cvt blocks on/off
 *   (cfg['use_dropout'], networks_seg.py:77-78).  All parameters live in ONE flat fp32 device array: the learnable ones
 *   first (conv weight/bias, BatchNorm gamma/beta, in the order of gsx_train_param_info), then the BatchNorm moving
 *   statistics; gsx_train_param_count returns the number of entries and the two sizes, gsx_train_param_info the reference
 *   (structural) name, offset and element count of entry `index` -- what SegSolver.save/load map to checkpoint_last.params.
 * gsx_train_step: train-mode forward (batch statistics; moving statistics updated in params_dev), SoftmaxCE with weight
 *   (label > -1), backward.  feats_f32_dev[level] [n,C,h,w] fp32 (the reference's features; the generator is frozen, :393),
 *   labels_dev [n,H,W] int32 (-1 = ignore).  Writes loss_dev[n], optionally the argmax of the logits (train metric), and
 *   *grad_scale_out * d(sum_n loss_n)/d(param) into grads_dev[learnable count]: the loss gradient is scaled by H*W for the
 *   16-bit backward pass; the caller all-reduces grads_dev (the step's only collective) and calls gsx_adam_step with
 *   rescale_grad = 1 / (global_batch * grad_scale).  Dropout masks are Philox bits of (dropout_seed, level, sample, element),
 *   recomputed in the backward pass; gsx_train_dropout_mask exports one ({0,1} floats [n,C,h,w]) for parity tests.
 *   Only enqueues work on `stream`; the workspace (gsx_train_workspace_bytes) is caller-owned. ---- */
typedef struct gsx_train gsx_train;
int gsx_train_create(const gsx_dec_cfg* cfg, int n, int use_dropout, gsx_train** out);
void gsx_train_destroy(gsx_train* h);
int gsx_train_param_count(const gsx_train* h, size_t* learnable, size_t* total);
int gsx_train_param_info(const gsx_train* h, int index, const char** name, size_t* offset, size_t* count);
int gsx_train_workspace_bytes(const gsx_train* h, size_t* bytes);
int gsx_train_dropout_mask(const gsx_train* h, int level, uint64_t seed, float* out_dev, gsx_stream stream);
/* CUDA-graph replays of a captured step: with a seed buffer set (one uint64 in device memory; NULL switches it off) the
 * kernels read the dropout seed from there instead of the dropout_seed argument baked into the captured launches. */
int gsx_train_set_seed_buffer(gsx_train* h, const uint64_t* seed_dev);
int gsx_train_step(gsx_train* h, const float* params_dev, float* grads_dev, const float* const* feats_f32_dev,
                   const gsx_synth* synth, const void* synth_ws, const int* labels_dev, uint64_t dropout_seed, float* loss_dev,
                   uint8_t* pred_mask_dev, float* grad_scale_out, void* ws, size_t ws_bytes, gsx_stream stream);

/* ---- per-launch timing of the forward passes (bench.py's per-layer roofline table): enable, run a
 *      forward, dump "label\tms\talgorithmic_bytes\talgorithmic_flops\texecuted_flops\tkernel\n" lines (executed = what the
 *      tensor cores are asked to do after the phase / space-to-depth decompositions).  Off by default. ---- */
int gsx_profile_enable(int on);
int gsx_profile_dump(char* buf, size_t cap);

/* ---- single-operator hooks (tests / tuning; fp32 NCHW device tensors in and out, temporaries
 *      are allocated inside, so not for the hot path) ---- */
int gsx_op_conv(int mode, int n, int h, int w, int cin0, int cin1, int cout, const float* x0_dev,
                const float* x1_dev, const float* w_host, const float* bias_dev, const float* nscale_dev,
                const float* noise_dev, int flags, const float* addsrc_dev, float* out_dev, float* stats_dev,
                uint8_t* mask_dev, float* logits_dev, int num_classes, const gsx_plan_override* ov,
                int* plan_out /* 16 ints */, int repeat, float* ms_out, gsx_stream stream);
/* Planner only (no GPU needed): the tiling gsx_op_conv / the forward passes would use. */
int gsx_plan_query(int mode, int h, int w, int cin0, int cin1, int cout, int num_classes,
                   const gsx_plan_override* ov, int* plan_out /* 16 ints */);
int gsx_op_pass1(int n, int c, int h, int w, const float* x_dev, int blur, int in_broadcast,
                 const float* nscale_dev, const float* bias_dev, const float* noise_dev, float* out_dev,
                 float* stats_dev, gsx_stream stream);
int gsx_op_apply(int n, int c, int h, int w, const float* x_dev, const float* stats_dev, const float* styles_dev,
                 const float* wrgb_dev, const float* brgb_dev, int nc, float* out_dev, float* img_f32_dev,
                 uint8_t* img_u8_dev, gsx_stream stream);
int gsx_op_fill_normal(float* out_dev, size_t per_sample, int n, uint64_t seed, uint64_t first_sample,
                       int stream_id, gsx_stream stream);

#ifdef __cplusplus
}
#endif
#endif /* GSX_H_ */
