/*
 * cdc_b200.h — C ABI of the B200-native CDC denoiser engine (libcdc_b200.so).
 *
 * The reference (buggyyang/CDC_compression) is pure Python/PyTorch and has no FFI of its own;
 * this ABI is what a binding for its decoder hot path binds instead of the PyTorch modules.
 * Every entry point below names the reference interface it replaces (paths relative to the
 * upstream tree).  Plain pointers and sizes only — no torch types.  All device pointers are
 * owned by the caller (PyTorch allocates inputs, outputs and the workspace); the engine owns
 * only the repacked weights.  The compute entry points (cdc_unet_forward, cdc_set_context,
 * cdc_set_schedule, cdc_context_decode, cdc_ddim_step, cdc_sample_loop, cdc_sample_loop_noise) enqueue their work on the
 * caller's stream and never synchronise it (cdc_set_schedule stages the table in an engine-owned
 * pinned buffer and waits only for its OWN previous copy, if that is still in flight; growing the
 * table re-allocates it); the introspection entry points marked "(synchronises)" do.  Every entry
 * point leaves the calling thread's current CUDA device as it found it.  Every function returns 0
 * on success and a negative cdc_status otherwise; cdc_last_error() gives the message.  An engine is
 * not re-entrant across host threads.
 */
#ifndef CDC_B200_H_
#define CDC_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CDC_ABI_VERSION 1
#define CDC_MAX_LEVELS 8

typedef enum cdc_status {
  CDC_OK = 0,
  CDC_ERR_INVALID = -1,      /* bad argument / shape */
  CDC_ERR_UNSUPPORTED = -2,  /* configuration outside the kernel family (fails loudly, no fallback) */
  CDC_ERR_MISSING = -3,      /* weight missing / wrong shape at finalize */
  CDC_ERR_CUDA = -4,         /* CUDA runtime error */
  CDC_ERR_STATE = -5         /* call order (e.g. forward before finalize) */
} cdc_status;

typedef enum cdc_variant { CDC_VARIANT_EPS = 0, CDC_VARIANT_X = 1 } cdc_variant;

/* eps: clip_noise in {"none","full","half"} (epsilonparam/modules/denoising_diffusion.py:140-143);
 * x:   clip_denoised bool -> NONE / FULL (xparam/modules/denoising_diffusion.py:163-164). */
typedef enum cdc_clip { CDC_CLIP_NONE = 0, CDC_CLIP_FULL = 1, CDC_CLIP_HALF = 2 } cdc_clip;

/* x-variant pred_mode (xparam/modules/denoising_diffusion.py:157-165). eps variant: NOISE only. */
typedef enum cdc_pred { CDC_PRED_NOISE = 0, CDC_PRED_X = 1, CDC_PRED_V = 2 } cdc_pred;

/* Mirrors the Unet constructor, epsilonparam/modules/unet.py:18-27 (xparam/modules/unet.py:19-29). */
typedef struct cdc_config {
  int32_t abi_version;                      /* CDC_ABI_VERSION */
  int32_t variant;                          /* cdc_variant */
  int32_t dim;                              /* base width (64 in both demos) */
  int32_t channels;                         /* image channels (3) */
  int32_t context_channels;                 /* 3 (eps demo) / 64 (x demo) */
  int32_t n_levels;                         /* len(dim_mults) */
  int32_t dim_mults[CDC_MAX_LEVELS];
  int32_t n_context;                        /* len(context_dim_mults) */
  int32_t context_dim_mults[CDC_MAX_LEVELS];
} cdc_config;

typedef struct cdc_engine cdc_engine;

/* Per-step scalars of the DDIM update; one row per schedule index i.
 * Built by the caller from GaussianDiffusion.set_sample_schedule's tables
 * (epsilonparam/modules/denoising_diffusion.py:81-97; xparam/...:89-108) so linspace rounding
 * stays with PyTorch. */
typedef struct cdc_step_coef {
  float sqrt_recip_acp;     /* sqrt(1/acp_t)            */
  float sqrt_recipm1_acp;   /* sqrt(1/acp_t - 1)        */
  float sqrt_acp_prev;      /* sqrt(acp_prev)           */
  float dir_coef;           /* sqrt(clamp?(1 - acp_prev - (eta*sigma)^2)) — computed by the caller   */
  float noise_coef;         /* eta * sigma_t (0 when eta == 0)                                      */
  float unet_time;          /* value fed to the U-Net: i/S (eps) or index[i]/T (x)                   */
  float sqrt_acp;           /* sqrt(acp_t)     (pred_mode "v" only) */
  float sqrt_1m_acp;        /* sqrt(1 - acp_t) (pred_mode "v" only) */
} cdc_step_coef;

/* ---- lifecycle: replaces Unet.__init__ + load_state_dict (unet.py:18-93, Appendix B keys) ---- */
int cdc_engine_create(const cdc_config* cfg, int device, cdc_engine** out);
void cdc_engine_destroy(cdc_engine* e);
const char* cdc_last_error(const cdc_engine* e);   /* e may be NULL: last create() error */
int cdc_abi_version(void);

/* Register one state_dict entry of the Unet (key WITHOUT the 'denoise_fn.' prefix), fp32,
 * host memory, contiguous, reference layout (Conv2d OIHW, ConvTranspose2d IOHW, Linear [out,in]). */
int cdc_engine_set_weight(cdc_engine* e, const char* key, const float* host_ptr,
                          const int64_t* shape, int ndim);
/* Validate the key set against the config, repack to the kernel layouts (fp16, tap-major,
 * K-major 64-channel chunks) and upload.  May be called again after weights change. */
int cdc_engine_finalize(cdc_engine* e);

/* Bytes of caller-provided device workspace for a B x channels x H x W problem (H, W % 32 == 0). */
int64_t cdc_engine_workspace_bytes(cdc_engine* e, int B, int H, int W);

/* ---- Unet.forward(x, time, context)  (epsilonparam/modules/unet.py:120-124; xparam :131-135) ----
 * x:    [B, channels, H, W] fp32 NCHW (device)      time: [B] fp32 (device; the [B,1] column)
 * ctx:  n_context fp32 NCHW device tensors, level l at (H>>l, W>>l) with the reference channel plan
 * out:  [B, channels, H, W] fp32 NCHW (device) */
int cdc_unet_forward(cdc_engine* e, const float* x, const float* time, const float* const* ctx,
                     int n_ctx, float* out, int B, int H, int W, void* workspace,
                     int64_t workspace_bytes, void* stream);

/* ---- context conversion, once per decode: the step-invariant half of Unet.encode's torch.cat
 * (unet.py:98).  After this call cdc_ddim_step / cdc_sample_loop reuse the converted context. */
int cdc_set_context(cdc_engine* e, const float* const* ctx, int n_ctx, int B, int H, int W,
                    void* workspace, int64_t workspace_bytes, void* stream);

/* ---- context_fn.decode on the engine: BigCompressor.decode
 * (epsilonparam/modules/compress_modules.py:74-82, layers :144-156) / ResnetCompressor.decode
 * (xparam/modules/compress_modules.py:68-74, layers :142-151), vbr=False.  Register the decoder's state_dict entries with
 * cdc_engine_set_weight under their GaussianDiffusion keys ('context_fn.dec.<i>.0.block1.block.0.weight', ...,
 * 'context_fn.dec.<i>.<1|2>.conv.weight') before cdc_engine_finalize.  q_latent: [B, C_latent, H/16, W/16] fp32 NCHW
 * (device) — the dequantised latent Compressor.encode returns.  Runs [ResnetBlock (no time embedding) -> Upsample] per
 * stage and leaves the four maps in the workspace in the layout the U-Net plan consumes: equivalent to cdc_set_context
 * with context_fn.decode(q_latent), without the NCHW fp32 round trip. */
int cdc_engine_has_context_decoder(cdc_engine* e);   /* 1 / 0 after finalize */
int cdc_context_decode(cdc_engine* e, const float* q_latent, int B, int H, int W, void* workspace,
                       int64_t workspace_bytes, void* stream);
/* Copy context map `level` (as stored by cdc_set_context / cdc_context_decode) to `out` as fp32 NCHW (device). */
int cdc_engine_read_context(cdc_engine* e, int level, float* out, int B, int H, int W, void* workspace,
                            int64_t workspace_bytes, void* stream);

/* ---- GaussianDiffusion.set_sample_schedule tables (see cdc_step_coef) ---- */
int cdc_set_schedule(cdc_engine* e, const cdc_step_coef* host_coefs, int S, void* stream);

/* ---- one GaussianDiffusion.ddim() step at schedule index i, in place on x
 * (epsilonparam/modules/denoising_diffusion.py:137-152; xparam/...:152-174).
 * z: optional [B,channels,H,W] fp32 standard-normal tensor (the reference's randn_like), may be
 * NULL when noise_coef == 0.  Requires cdc_set_context + cdc_set_schedule. */
int cdc_ddim_step(cdc_engine* e, float* x_inout, int i, const float* z, int pred_mode, int clip_mode,
                  int B, int H, int W, void* workspace, int64_t workspace_bytes, void* stream);

/* ---- GaussianDiffusion.p_sample_loop for eta == 0 (denoising_diffusion.py:166-192; x :179-205):
 * runs schedule indices i = i_first, i_first-1, ..., i_last (inclusive) in place on x, replaying
 * one captured CUDA graph per step.  i_first = S-1, i_last = 0 is the full loop. */
int cdc_sample_loop(cdc_engine* e, float* x_inout, int i_first, int i_last, int pred_mode,
                    int clip_mode, int B, int H, int W, void* workspace, int64_t workspace_bytes,
                    void* stream);

/* ---- the same loop for eta != 0 (the reference adds eta * sigma_t * randn_like(x) every step, denoising_diffusion.py:150;
 * xparam :171): z holds the standard-normal tensors of the (i_first - i_last + 1) steps of this call back to back, step i
 * at z + (i_first - i) * B*channels*H*W (the caller draws them with PyTorch's generator, in loop order, so the random
 * stream is the reference's).  One graph replay per step; the graph is re-captured when z changes. */
int cdc_sample_loop_noise(cdc_engine* e, float* x_inout, int i_first, int i_last, const float* z,
                          int pred_mode, int clip_mode, int B, int H, int W, void* workspace,
                          int64_t workspace_bytes, void* stream);

/* ---- introspection (tests, bench) ---- */
/* number of kernel launches one U-Net forward / one DDIM step enqueues for this shape */
int cdc_engine_launches_per_forward(cdc_engine* e, int B, int H, int W);
int cdc_engine_launches_per_step(cdc_engine* e, int B, int H, int W);
/* algorithmic FLOPs (2*MAC of every conv / transposed conv + the two attention einsums) of one
 * forward for this shape — SURVEY.md §8(d)'s figure, computed from the layer plan. */
double cdc_engine_flops_per_forward(cdc_engine* e, int B, int H, int W);
/* Debug taps: copy the fp16 NHWC activation produced by plan op `op_index` of the last forward to
 * host fp32 NHWC (synchronises). Returns element count or <0. */
int64_t cdc_engine_debug_read(cdc_engine* e, int op_index, float* host_out, int64_t capacity,
                              int* C, int* H, int* W);
int cdc_engine_num_ops(cdc_engine* e, int B, int H, int W);
const char* cdc_engine_op_name(cdc_engine* e, int op_index);
/* Time every op of the last-run plan in isolation: `iters` back-to-back launches between two CUDA events on
 * `stream`; writes average ms and the op's algorithmic FLOPs.  Returns the number of ops (synchronises). */
int cdc_engine_profile_ops(cdc_engine* e, int iters, float* ms_out, double* flops_out, int capacity, void* stream);
/* Debug: 1 = never reuse workspace buffers, so every op output survives until debug_read. */
int cdc_engine_set_debug(cdc_engine* e, int no_reuse);
/* Select the conv mainloop: 0 = mma.sync (HMMA) baseline kernels, 1 = tcgen05/TMA kernels. */
int cdc_engine_set_mainloop(cdc_engine* e, int kind);
/* How many ops of the plan for this shape run on the tcgen05/TMA kernel under the current mainloop. */
int cdc_engine_tc_ops(cdc_engine* e, int B, int H, int W);

#ifdef __cplusplus
}
#endif
#endif /* CDC_B200_H_ */
