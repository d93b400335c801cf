/*
 * eo_b200.h -- C ABI of libeo_b200.so, the B200 (sm_100a) implementation of
 * EO_Diffusion's reverse-sampling hot path.
 *
 * The reference (furio1999/EO_Diffusion) has no FFI: its "operator API" for this path is
 * three Python classes.  Each entry point below replaces the body of one reference
 * method; the Python host in eo_diffusion_b200/ keeps the reference signatures and calls
 * these through ctypes (see INTEGRATION.md for the stub a reference maintainer would add).
 *
 * Conventions
 *   - plain C types only; device pointers are raw CUDA device addresses; `stream` is a
 *     cudaStream_t passed as void* (0 = legacy default stream);
 *   - every call only ENQUEUES work on `stream` (no device-wide synchronisation), except
 *     eo_unet_finalize which may synchronise `stream` once while packing weights;
 *   - return value 0 = success, negative = error; eo_last_error() returns a thread-local
 *     message for the last failing call on this thread.  The library never aborts;
 *   - image tensors at the boundary are the reference's: NCHW, contiguous, fp32;
 *     timesteps are int64 (reference: diffusion/model.py:56);
 *   - there is no CPU fallback: every entry point requires a CUDA device of compute
 *     capability 10.x and fails with EO_ERR_DEVICE otherwise.
 */
#ifndef EO_B200_H
#define EO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define EO_API __attribute__((visibility("default")))
#else
#define EO_API
#endif

#define EO_OK 0
#define EO_ERR_ARG (-1)     /* bad argument / unsupported configuration            */
#define EO_ERR_CUDA (-2)    /* a CUDA runtime or driver call failed                */
#define EO_ERR_STATE (-3)   /* call order violated (e.g. forward before finalize)  */
#define EO_ERR_DEVICE (-4)  /* no sm_100 device                                    */
#define EO_ERR_KEY (-5)     /* unknown state-dict key / shape mismatch             */

/* arithmetic mode of the UNet internals; the sampler state is always fp32 */
#define EO_MODE_FP32 0 /* fp32 activations + fp32 FFMA kernels (parity mode, rel-L2 <= 1e-4) */
#define EO_MODE_BF16 1 /* bf16 activations, tcgen05 tensor-core kernels, fp32 accumulate     */

EO_API const char* eo_last_error(void);
EO_API int eo_version(void);
/* 0 if the current CUDA device is usable (sm_100), EO_ERR_DEVICE otherwise */
EO_API int eo_device_check(void);

/* ------------------------------------------------------------------------------------
 * UNet: replaces UNetModel.forward (backbones/unet_openai.py:746-780) and everything it
 * calls (ResBlock :365-385, AttentionBlock :427-433, QKVAttentionLegacy :465-481,
 * QKVAttention :497-515, Downsample :269-271, Upsample :229-242, GroupNorm32 :11-13,
 * timestep_embedding :81-99, time_embed MLP :597-602).
 * The configuration mirrors UNetModel.__init__ (unet_openai.py:553-575).
 * ---------------------------------------------------------------------------------- */
typedef struct eo_unet eo_unet;

typedef struct eo_unet_cfg {
  int32_t in_channels;      /* channels of x after the optional concat with cond */
  int32_t model_channels;
  int32_t out_channels;
  int32_t num_res_blocks;
  int32_t n_attention_resolutions;
  int32_t attention_resolutions[8];
  int32_t n_channel_mult;
  int32_t channel_mult[8];
  int32_t time_emb_factor;
  int32_t num_classes;      /* 0 = not class-conditional */
  int32_t num_heads;
  int32_t num_head_channels; /* -1 = use num_heads */
  int32_t num_heads_upsample; /* -1 = num_heads */
  int32_t use_new_attention_order;
  /* options of the reference constructor that this path does not implement (no configuration of the
     reference sets them otherwise); they must hold the values below or eo_unet_create fails with EO_ERR_ARG */
  int32_t dims;                 /* 2 */
  int32_t conv_resample;        /* 1 */
  /* the switches of the reference's UNet / UNetBig / UNetSmall factories (unet_openai.py:783-922) */
  int32_t use_scale_shift_norm; /* 0 | 1: FiLM conditioning, unet_openai.py:377-381 */
  int32_t resblock_updown;      /* 0 | 1: up/down-sampling ResBlocks, unet_openai.py:366-371 */
} eo_unet_cfg;

EO_API int eo_unet_create(const eo_unet_cfg* cfg, eo_unet** out);
EO_API void eo_unet_destroy(eo_unet* u);

/* Number of state-dict entries the engine consumes, and their names/shapes, in the
 * reference's state_dict() order (dead `nout.*` / `conv_out.*` entries, unet_openai.py:744,
 * are not consumed).  `shape` receives up to 4 extents; returns ndim or negative. */
EO_API int eo_unet_num_weights(const eo_unet* u);
EO_API const char* eo_unet_weight_name(const eo_unet* u, int index);
EO_API int eo_unet_weight_shape(const eo_unet* u, int index, int64_t shape[4]);

/* Hand one fp32, contiguous, device-resident parameter to the engine under its reference
 * state-dict key (e.g. "input_blocks.1.0.in_layers.2.weight").  The engine copies/repacks
 * it into kernel layout at finalize (the engine keeps private copies of everything it reads later; `dev_ptr` is
 * not dereferenced after eo_unet_finalize returns); the caller keeps ownership of `dev_ptr`. */
EO_API int eo_unet_set_weight(eo_unet* u, const char* key, const float* dev_ptr,
                       const int64_t* shape, int ndim);

/* Pack weights for `mode`, size the activation workspace for batches up to `max_batch`
 * of H x W images and build the launch plan.  May be called again after weights, mode or
 * geometry change. */
EO_API int eo_unet_finalize(eo_unet* u, int mode, int max_batch, int H, int W, void* stream);

/* eps = UNet(cat(x, cond), timesteps[, y]).
 *   x         [B, Cx, H, W] fp32 NCHW
 *   cond      [B, Cc, H, W] fp32 NCHW or NULL (Cx + Cc == in_channels; unet_openai.py:754-756)
 *   timesteps [B] int64, device
 *   y         [B] int64, device, or NULL (must be non-NULL iff num_classes > 0, :758-760)
 *   eps_out   [B, out_channels, H, W] fp32 NCHW
 */
EO_API int eo_unet_forward(eo_unet* u, const float* x, int Cx, const float* cond, int Cc,
                    const int64_t* timesteps, const int64_t* y, float* eps_out, int B,
                    void* stream);

/* Hoist the timestep-embedding path out of the sampling loop (timestep_embedding -> time_embed MLP -> every
 * ResBlock's emb_layers projection; unet_openai.py:763, :374-376 -- work that depends on the timestep VALUE only):
 * computes the rows of the timestep values 0 .. n_timesteps-1 once, with the kernels of the per-step path (rows are
 * bit-identical to it).  While a table is installed, eo_unet_forward with y == NULL gathers row timesteps[b] (one launch
 * instead of four); a timestep outside [0, n_timesteps) then yields NaN eps, so a sampler installs the table of ITS
 * schedule (EODiffusion: n_timesteps = self.timesteps) and clears it afterwards.  Class-conditional forwards
 * (y != NULL) keep the per-step path: label_emb(y) enters before the projections (:764-766).  Building synchronises
 * `stream` once; a later call with n_timesteps <= the built size only switches the table back on.
 * eo_unet_clear_time_tables switches it off (forwards take the per-step path again); the memory stays with the handle
 * until the next eo_unet_finalize / eo_unet_destroy, because row t depends on t alone and stays valid. */
EO_API int eo_unet_build_time_tables(eo_unet* u, int n_timesteps, void* stream);
EO_API int eo_unet_clear_time_tables(eo_unet* u);

/* Profiling variant of eo_unet_forward used by bench.py's roofline leg: brackets every op of
 * the launch plan with CUDA events on `stream`, waits for the last one and writes the device
 * time of op i (milliseconds) to ms_per_op[i], i < eo_unet_num_ops().  eo_unet_op_info names
 * op i (reference module path), its kernel family, and its ALGORITHMIC FLOPs / HBM bytes
 * per image (DESIGN.md states the formulas). */
EO_API int eo_unet_forward_timed(eo_unet* u, const float* x, int Cx, const float* cond, int Cc,
                          const int64_t* timesteps, const int64_t* y, float* eps_out, int B,
                          void* stream, float* ms_per_op);
EO_API int eo_unet_num_ops(const eo_unet* u);
EO_API int eo_unet_op_info(const eo_unet* u, int index, const char** name, const char** kernel,
                    double* flops_per_image, double* bytes_per_image);

/* FLOPs op `index` EXECUTES per image (zero-padded qkv / head / stem rows and head-dimension padding count, the
 * sub-pixel form of Upsample counts 4/9): the figure to hold against ncu's tensor-pipe utilisation, next to the
 * algorithmic one of eo_unet_op_info.  Negative on a bad index. */
EO_API double eo_unet_op_executed_flops(const eo_unet* u, int index);

/* bytes of device memory held by the engine (packed weights + workspace) */
EO_API int64_t eo_unet_device_bytes(const eo_unet* u);
/* number of kernel launches one eo_unet_forward enqueues (for bench.py's gpu_launches) */
EO_API int eo_unet_launches_per_forward(const eo_unet* u);

/* debugging/validation aid used by the parity tests: copy an internal activation
 * (name = reference module path, e.g. "input_blocks.3") of the LAST forward into `out`
 * as fp32 NCHW.  Returns element count, or negative if unknown. */
EO_API int64_t eo_unet_read_activation(eo_unet* u, const char* name, float* out_dev, int64_t capacity,
                                int B, void* stream);

/* ------------------------------------------------------------------------------------
 * DDPM sampler arithmetic: replaces the elementwise tails of EODiffusion.sampling
 * (diffusion/model.py:58-60), _forward_diffusion (:94-98), _reverse_diffusion_with_clip
 * (:133-150) and _reverse_diffusion (:110-122).  Bit-exact with the reference's fp32 op
 * sequence (no FMA contraction).
 *
 * `table` is a device array [T][EO_DDPM_NCOEF] of per-timestep scalars that the host
 * computes with the reference's own torch op sequence; kernels gather row timesteps[b],
 * like the reference's `.gather(-1, t)`.
 * ---------------------------------------------------------------------------------- */
#define EO_DDPM_NCOEF 12
#define EO_COEF_SQRT_ACP 0          /* sqrt_alphas_cumprod[t]                       */
#define EO_COEF_SQRT_1M_ACP 1       /* sqrt_one_minus_alphas_cumprod[t]             */
#define EO_COEF_SQRT_RECIP_ACP 2    /* sqrt(1/acp[t])                                */
#define EO_COEF_SQRT_RECIPM1_ACP 3  /* sqrt(1/acp[t] - 1)                            */
#define EO_COEF_MEAN_X0 4           /* beta*sqrt(acp[t-1])/(1-acp[t])      (t>0)     */
#define EO_COEF_MEAN_XT 5           /* (1-acp[t-1])*sqrt(alpha)/(1-acp[t]) (t>0)     */
#define EO_COEF_STD 6               /* sqrt(beta*(1-acp[t-1])/(1-acp[t]))  (t>0)     */
#define EO_COEF_MEAN_X0_T0 7        /* beta/(1-acp[t])                               */
#define EO_COEF_RECIP_SQRT_ALPHA 8  /* 1/sqrt(alpha[t])                              */
#define EO_COEF_EPS_NOCLIP 9        /* (1-alpha[t])/sqrt_one_minus_acp[t]            */

/* x_out = mask*(sa[t]*gt + sb[t]*noise) + (1-mask)*x_t     (model.py:59-60)
 *   x_t, gt, noise, x_out [B,C,H,W]; mask [B,1,H,W].
 * In this and the two step calls below x_out may be the same buffer as x_t (the kernels read and write element i
 * only, and do not declare the two pointers __restrict__); no other pair of arguments may overlap.  timesteps[b]
 * must lie in [0, T) of `table` (not checked on the device; the reference's gather would raise). */
EO_API int eo_ddpm_sum_mix(const float* x_t, const float* gt, const float* mask, const float* noise,
                    const int64_t* timesteps, const float* table, float* x_out, int B, int C,
                    int HW, void* stream);

/* x_out = posterior_mean(x_t, eps) + std*noise.
 *   clip != 0: _reverse_diffusion_with_clip, else _reverse_diffusion.
 *   all_t_positive: the host's evaluation of the reference's `t.min() > 0` (model.py:140). */
EO_API int eo_ddpm_step(const float* x_t, const float* eps, const float* noise, const int64_t* timesteps,
                 const float* table, float* x_out, int B, int C, int HW, int clip,
                 int all_t_positive, void* stream);

/* Fused: the step of timestep t followed by the 'sum' mix of the NEXT iteration
 * (timesteps_next, noise_next), saving one round trip of x through HBM. */
EO_API int eo_ddpm_step_mix(const float* x_t, const float* eps, const float* noise,
                     const int64_t* timesteps, const float* gt, const float* mask,
                     const float* noise_next, const int64_t* timesteps_next, const float* table,
                     float* x_out, int B, int C, int HW, int clip, int all_t_positive,
                     void* stream);

/* The whole DDPM trajectory in one call: the loop of EODiffusion.sampling (diffusion/model.py:46-75) without its
 * PNG side effect and with the reference's random draws handed in as a tape, so a C caller needs no per-step round trip.
 *   x             [B, Cx, H, W]: in x_T, out x_0 (un-normalised, like the reference's return value)
 *   noise_tape    [T][B, Cx, H, W]: tape[k] = the k-th torch.randn_like draw (k = 0 drives iteration i = T-1; in 'sum'
 *                 mode the same draw forward-diffuses gt, model.py:57-59)
 *   gt, mask      'sum' conditioning ([B, Cx, H, W], [B, 1, H, W]) or both NULL
 *   cond, Cc      concat conditioning ([B, Cc, H, W]) or NULL, 0
 *   y             [B] int64 class labels or NULL
 *   timestep_rows [T][B] int64, row i = the value i repeated (what the reference builds with torch.tensor([i] * n))
 *   table         [T][EO_DDPM_NCOEF], see above
 *   eps_scratch   [B, out_channels, H, W] work buffer for the UNet output
 * Same kernels in the same order as the per-step entry points: bit-identical to driving them from the host.
 * B, Cx, H, W must match the finalized geometry (batch <= max_batch); uses the timestep tables of 0 .. T-1
 * (eo_unet_build_time_tables) for its own loop when y == NULL and leaves the handle's table setting as it found it.
 * Threading: an eo_unet handle carries per-forward state (staging buffers, CUDA graphs); calls on ONE handle
 * must be serialised by the caller -- use one handle per host thread / device. */
EO_API int eo_sample_ddpm(eo_unet* u, float* x, const float* noise_tape, const float* gt, const float* mask,
                   const float* cond, int Cc, const int64_t* y, const int64_t* timestep_rows,
                   const float* table, float* eps_scratch, int T, int B, int Cx, int H, int W, int clip,
                   void* stream);

/* ------------------------------------------------------------------------------------
 * DDIM step: replaces DDIMSampler.p_sample_ddim after the UNet call (diffusion/ddim.py:
 * 187-207).  Scalars are the fp32 values the reference materialises with torch.full:
 *   sqrt_a_t = sqrt(a_t), sqrt_1m_a_t, sqrt_a_prev = sqrt(a_prev),
 *   dir_coef = sqrt(1 - a_prev - sigma_t^2), sigma_t, temperature.
 * noise may be NULL when sigma_t == 0.  Writes x_prev and pred_x0 (either may alias x).
 * ---------------------------------------------------------------------------------- */
EO_API int eo_ddim_step(const float* x, const float* e_t, const float* noise, float* x_prev,
                 float* pred_x0, float sqrt_a_t, float sqrt_1m_a_t, float sqrt_a_prev,
                 float dir_coef, float sigma_t, float temperature, int64_t n_elems,
                 void* stream);

/* The whole DDIM trajectory in one call: the loop of DDIMSampler.ddim_sampling (diffusion/ddim.py:114-164) over
 * p_sample_ddim (:166-207), without classifier-free guidance and callbacks.
 *   x             [B, Cx, H, W]: in x_T, out the final sample
 *   noise_tape    [S][B, Cx, H, W]: tape[k] = the draw of ddim.py:203 at the k-th iteration (index = S-1-k), or NULL
 *                 when every sigma is 0 (eta = 0)
 *   cond, Cc, y   as in eo_unet_forward
 *   timestep_rows [S][B] int64 device, row `index` = ddim_timesteps[index] repeated
 *   scalars       HOST array [S][6] of the fp32 values of eo_ddim_step for each index:
 *                 sqrt_a_t, sqrt_1m_a_t, sqrt_a_prev, dir_coef, sigma_t, temperature
 *   eps_scratch, pred_x0  [B, Cx, H, W] work buffers (pred_x0 holds the last step's x0 prediction on return) */
EO_API int eo_sample_ddim(eo_unet* u, float* x, const float* noise_tape, const float* cond, int Cc, const int64_t* y,
                   const int64_t* timestep_rows, const float* scalars, float* eps_scratch, float* pred_x0,
                   int S, int B, int Cx, int H, int W, void* stream);

/* Classifier-free guidance combine e = e_u + s*(e_c - e_u) (ddim.py:180-181). */
EO_API int eo_cfg_combine(const float* e_uncond, const float* e_cond, float scale, float* e_out,
                   int64_t n_elems, void* stream);

/* ------------------------------------------------------------------------------------
 * Post-processing of finished samples and image-quality metrics (reference inference.py:128-150;
 * SURVEY.md 8f row 2).  Elementwise passes are bit-exact with the reference's fp32 torch ops.
 * ---------------------------------------------------------------------------------- */
/* out = clip(x, 0, 1) (mode 0, inference.py:128), (x + 1) / 2 (mode 1, :128/:136), or torchvision
 * adjust_brightness(x, factor) = clip(factor * x, 0, 1) for float images (mode 2, :141-142, :149).
 * out may alias x. */
EO_API int eo_post_map(const float* x, float* out, int64_t n, int mode, float factor, void* stream);
/* out[b,c,p] = image[b,c,p] * clip(mask[b,0,p] + 0.7, 0, 1)   (inference.py:135) */
EO_API int eo_post_dim_masked(const float* image, const float* mask, float* out, int B, int C, int HW, void* stream);
/* torchvision.utils.save_image's device-side half (reference inference.py:143-150 and, inside the sampling loop,
 * diffusion/model.py:62-66): make_grid(x [B,C,H,W], nrow, padding, pad_value) followed by
 * `mul(255).add_(0.5).clamp_(0, 255).to(uint8)`, written as the HWC uint8 array PIL.Image.fromarray takes, so only
 * 3 bytes per grid pixel cross PCIe.  Grid geometry as torchvision: one image comes back without a border, a
 * single-channel batch is repeated to three channels.  pre = 1 applies (x + 1) / 2 first (model.py:63).
 * grid_hw_or_null (HOST, 3 ints) receives {grid height, grid width, grid channels}; with x = out = NULL the call
 * only answers that query.  Bit-exact with torchvision. */
EO_API int eo_post_grid_u8(const float* x, unsigned char* out, int B, int C, int H, int W, int nrow, int padding,
                    float pad_value, int pre, int* grid_hw_or_null, void* stream);
/* out3 = {mean, min, max} of x (the values the reference's host branches read: image.min() :128,
 * gt.mean() / cond.mean() / samples.mean() :141-149).  workspace4: 4 doubles of device scratch. */
EO_API int eo_post_stats(const float* x, int64_t n, double* workspace4, float* out3, void* stream);
/* torchmetrics peak_signal_noise_ratio(preds, target, data_range) with the default reduction over the
 * whole batch: out1[0] = 10 log10(data_range^2 / mean((preds - target)^2))   (inference.py:138) */
EO_API int eo_psnr(const float* preds, const float* target, int64_t n, float data_range, double* workspace4,
            float* out1, void* stream);
/* torchmetrics structural_similarity_index_measure(preds, target, data_range) with its defaults (Gaussian
 * 11 x 11 window, sigma 1.5, k1 0.01, k2 0.03, mean over the batch).  preds, target [B,C,H,W];
 * gauss11: HOST array of the 11 normalised fp32 window weights (the host computes them with torchmetrics'
 * op sequence); workspaceB: B doubles of device scratch; per_image_or_null [B], mean_out [1] on device. */
EO_API int eo_ssim(const float* preds, const float* target, int B, int C, int H, int W, float data_range,
            const float* gauss11, double* workspaceB, float* per_image_or_null, float* mean_out, void* stream);

/* ------------------------------------------------------------------------------------
 * Kernel self-tests (used by tests/ on the GPU box): run one tensor-core implicit-GEMM
 * convolution / one attention call on caller-provided buffers, outside any UNet.
 * ---------------------------------------------------------------------------------- */
/* y[B,H,W,Cout] (bf16 NHWC) = conv_kxk(x[B,H,W,Cin] bf16 NHWC, w[Cout,Cin,k,k] fp32) + bias
 * [+ residual bf16 NHWC]; k in {1,3}, stride 1, pad k/2. */
EO_API int eo_test_conv_tc(const void* x_bf16, const float* w, const float* bias, const void* residual,
                    void* y_bf16, int B, int H, int W, int Cin, int Cout, int k, void* stream);
/* qkv [B,T,3*heads*ch] in the LEGACY channel order (head, {q,k,v}, ch) bf16 ->
 * out [B,T,heads*ch] bf16 */
EO_API int eo_test_attention_tc(const void* qkv_bf16, void* out_bf16, int B, int T, int heads, int ch,
                         void* stream);

/* Development aid (tools/conv_trace.py): while dev_buf != NULL every tensor-core convolution launch
 * writes, for its first n_ctas CTAs, eight int64 phase stamps (globaltimer at entry, clock64 at setup
 * done / first operands landed / accumulator complete / epilogue done / exit, SM id).  NULL turns it off. */
EO_API int eo_debug_conv_trace(void* dev_buf, int n_ctas);

#ifdef __cplusplus
}
#endif
#endif /* EO_B200_H */
