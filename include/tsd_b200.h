/*
 * tsd_b200 - C ABI of the B200-native Tiny-Stable-Diffusion denoising path.
 *
 * The reference (lrmantovani10/Stable-Diffusion.mojo) has no FFI/plugin interface: its
 * boundary is the Mojo struct API  X(...).forward(Matrix) -> Matrix.  Every entry point below
 * names the reference method it replaces (file:line into the reference tree).  A Mojo shim
 * binds these with sys.ffi.DLHandle / external_call; this repo's tests
 * bind them with Python ctypes.
 *
 * Conventions
 *  - plain C: pointers + sizes, no C++/torch types.  Host entry points take host float32
 *    buffers in the reference's layouts: Matrix = 3-D row-major (dim0, dim1, dim2)
 *    (helpers/utils.mojo:521-524, index = z*dim1*dim2 + y*dim2 + x, :805-811); images are
 *    (C, H, W), sequences (1, T, C); conv kernels OIHW (Matrix_Array, :1718); linear weights
 *    [out][in] (:1943).  `_dev` variants take device pointers in the same layouts.
 *  - every function returns int32 status: 0 = OK.  Nothing throws.  The message of the last
 *    failure is tsd_last_error(ctx).  The reference's convention "print + return the null
 *    Matrix(0,0,0)" (e.g. helpers/utils.mojo:1551-1552, 1848-1853, 1956-1957) maps to a
 *    non-zero status; the shim prints tsd_last_error and returns Matrix(0,0,0).
 *  - there is NO CPU fallback: without an sm_100 device tsd_init fails (TSD_ERR_NO_DEVICE).
 *  - a context owns one CUDA stream; calls on one context are serialised by an internal mutex,
 *    distinct contexts may be used from different threads.
 */
#ifndef TSD_B200_H
#define TSD_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(TSD_BUILD)
#pragma GCC visibility push(default)
#endif

#define TSD_OK 0
#define TSD_ERR_INVALID 1   /* bad shape / argument */
#define TSD_ERR_CUDA 2      /* CUDA runtime or driver error */
#define TSD_ERR_NO_DEVICE 3 /* no sm_100 (B200) device */
#define TSD_ERR_OOM 4
#define TSD_ERR_STATE 5     /* e.g. forward before load_weights */

typedef struct tsd_ctx tsd_ctx;             /* device + stream + workspace */
typedef struct tsd_diffusion tsd_diffusion; /* Diffusion (diffusion.mojo:294-318) */
typedef struct tsd_decoder tsd_decoder;     /* VAE Decoder (vae.mojo:162-250) */
typedef struct tsd_clip tsd_clip;           /* CLIP text encoder (clip.mojo:56-109) */
typedef struct tsd_encoder tsd_encoder;     /* VAE Encoder (vae.mojo:70-159) */
typedef struct tsd_tokenizer tsd_tokenizer; /* Tokenizer (helpers/utils.mojo:228-292) */
typedef struct tsd_safetensors tsd_safetensors; /* checkpoint file (the reference's weight-loading TODO, README.md:44,55) */

/* ---- context ------------------------------------------------------------------------ */
int32_t tsd_init(int32_t device, tsd_ctx** out);
int32_t tsd_shutdown(tsd_ctx* ctx);
const char* tsd_last_error(const tsd_ctx* ctx); /* ctx may be NULL: last tsd_init failure */
int32_t tsd_synchronize(tsd_ctx* ctx);
/* Semantic switches of SURVEY.md section 0 and engine knobs:
 *   "softmax_axis"     0 = query axis, as reference Softmax(dim=2) (utils.mojo:435-445) [default]
 *                      1 = key axis (standard attention)
 *   "layernorm_mode"   0 = one mean/std over the whole (C,T) tensor, as reference
 *                          LayerNorm = GroupNorm(1,C) (utils.mojo:2052-2061) [default]
 *                      1 = per token
 *   "norm_eps_mode"    0 = (x - mean) / (std + eps), as reference GroupNorm.forward (utils.mojo:1868-1870) [default]
 *                      1 = (x - mean) / sqrt(var + eps) (what real checkpoints were trained with)
 *   "fused_attention"  1 = fused tcgen05 attention kernel [default], 0 = GEMM+softmax+GEMM
 *   "cuda_graph"       1 = replay model forwards from a captured CUDA graph [default], 0 = eager
 * Execution-plan knobs (never change semantics; results agree to TF32 rounding level):
 *   "autotune"         1 = time tile candidates on first use of a GEMM signature [default], 0 = cost model
 *                          (TSD_TUNE_CACHE=<file> in the environment persists the choices)
 *   "gemm_cg"          0 = plan decides [default], 1 = single CTAs, 2 = CTA pairs (tcgen05 cta_group::2)
 *   "conv_halo"        3x3 convolutions: 0 = implicit GEMM [default], 1 = autotuner may pick the
 *                          halo-in-shared-memory kernel, 2 = whenever eligible
 *   "producer_stats"   1 = GEMM epilogues leave GroupNorm/LayerNorm partial sums for the consumer [default]
 *   "ln_fold"          1 = global-statistics LayerNorm folded into the consuming GEMM epilogue [default]
 *   "fuse_skip"        1 = a ResBlock's 1x1 skip convolution is a second K segment of its conv2 GEMM [default],
 *                          0 = a GEMM of its own whose result conv2 adds as a residual
 *   "fuse_ffn_out"     1 = the two linears that end an attention block (geglu2, conv_out: diffusion.mojo:141-146) run as one
 *                          GEMM over K = 4C + C with weights merged at load [default], 0 = two GEMMs
 *   "gemm_kmerge"      1 = one TMA request per operand and K step of 64 [default]; "gemm_deep_b" 1 = separate deep weight
 *                          ring for narrow tiles (0 [default], measured slower)
 *   "conv_stride_tma"  1 = stride-2 3x3 convolutions are implicit GEMMs through a tensor map with element strides
 *                          [default], 0 = im2col kernel + GEMM
 *   "defer_reduce"     1 = where a split-K GEMM feeds a norm directly, the norm kernel sums the partial tiles
 *                          (no reduce kernel) [default]
 *   "virtual_concat"   1 = the channel concats of the UNet are read (and written out) by the GroupNorm kernel of the
 *                          ResBlock they feed instead of a concat kernel of their own [default]; 0 = concat kernel
 *   "norm_cluster"     1 = GroupNorm by one thread-block cluster per (image, group), slab in shared memory, statistics
 *                          exchanged through distributed shared memory [default]; 0 = grid-barrier kernels only
 *   "norm_v2"          0/1 = which grid-barrier fused norm kernel serves the shapes the cluster kernel does not take
 *   "attn_v2"          1 = attention kernel with three S/P buffers in tensor memory [default], 0 = round-1 kernel
 *   "splitk_cluster"   1 = split-K partials reduced inside the GEMM through distributed shared memory (cluster over
 *                          the splits); 0 [default] (measured slower than partial tiles + consumer-side reduction)
 *   "tune_defer_penalty_us"  autotuner: microseconds charged to a split-K candidate whose consumer sums the partials
 *   "splitk_fixup"     1 = split-K partials reduced in-kernel by the last CTA of each tile, 0 = reduce kernel [default]
 *   "pdl"              1 = programmatic dependent launch between the kernels of a graph [default]
 *   "force_bn" / "force_splits" / "force_stages" / "gemm_debug" / "halo_min_w" / "halo_min_h" /
 *   "tune_verbose" / "tune_flush" / "bench_stats_groups": tests and lab tooling only.
 * Every option can also be preset as TSD_OPT_<name>=<int> in the environment of tsd_init().
 */
int32_t tsd_set_option(tsd_ctx* ctx, const char* name, int32_t value);
int32_t tsd_get_option(tsd_ctx* ctx, const char* name, int32_t* value);
/* number of kernels launched through ctx since creation (bench "gpu_launches") */
int64_t tsd_launch_count(const tsd_ctx* ctx);
/* CUDA-event stopwatch on the context's stream: start records, stop records + waits and returns
 * the device milliseconds in between (bench.py times the `_dev` entry points with it). */
int32_t tsd_timer_start(tsd_ctx* ctx);
int32_t tsd_timer_stop(tsd_ctx* ctx, double* ms);

/* ---- op level (host buffers, reference layouts) ---------------------------------------- */
/* Conv2D.forward, helpers/utils.mojo:1738-1811.  x (cin,h,w); weight OIHW (cout,cin,k,k);
 * bias (cout) or NULL; out (cout, ho, wo), ho = floor((h+2*pad-k)/stride)+1.  n images batched. */
int32_t tsd_conv2d(tsd_ctx* ctx, const float* x, int32_t n, int32_t cin, int32_t h, int32_t w,
                   const float* weight, const float* bias, int32_t cout, int32_t k, int32_t pad,
                   int32_t stride, float* out);
/* Matrix.pad((pad,pad_hi),(pad,pad_hi)) (helpers/utils.mojo:1383-1413) followed by Conv2D.forward with
 * padding (0,0): `pad` zero rows/columns above/left, `pad_hi` below/right.  Encoder.two_stride_pad
 * (vae.mojo:115-116) is pad = 0, pad_hi = 1.  ho = floor((h+pad+pad_hi-k)/stride)+1. */
int32_t tsd_conv2d_pad(tsd_ctx* ctx, const float* x, int32_t n, int32_t cin, int32_t h, int32_t w,
                       const float* weight, const float* bias, int32_t cout, int32_t k, int32_t pad,
                       int32_t pad_hi, int32_t stride, float* out);
/* Linear.forward, helpers/utils.mojo:1954-1976 (bias added to every column).
 * x (b,t,in_f); weight (out_f,in_f); bias (out_f) or NULL; out (b,t,out_f). */
int32_t tsd_linear(tsd_ctx* ctx, const float* x, int32_t b, int32_t t, int32_t in_f,
                   const float* weight, const float* bias, int32_t out_f, float* out);
/* Matrix.matmul, helpers/utils.mojo:1549-1569: out[c] = a[c] (m,k) x b[c] (k,n). */
int32_t tsd_matmul(tsd_ctx* ctx, const float* a, const float* b, int32_t c, int32_t m, int32_t k,
                   int32_t n, float* out);
/* GroupNorm.forward, helpers/utils.mojo:1845-1885: (x-mean)/(std+eps)*gamma, biased std,
 * scalar gamma = 1, no beta.  x,out (c,h,w) x n images.  gamma/beta: optional per-channel
 * vectors (superset for real checkpoints), NULL = reference behaviour. */
int32_t tsd_groupnorm(tsd_ctx* ctx, const float* x, int32_t n, int32_t c, int32_t h, int32_t w,
                      int32_t groups, float eps, const float* gamma, const float* beta, float* out);
/* LayerNorm.forward, helpers/utils.mojo:2052-2061, on a (c,t,1) Matrix; eps = 1e-5. */
int32_t tsd_layernorm(tsd_ctx* ctx, const float* x, int32_t c, int32_t t, float* out);
/* SiLU.forward :1892-1902 / Gelu.forward :1908-1919 (tanh form); n elements. */
int32_t tsd_silu(tsd_ctx* ctx, const float* x, int64_t n, float* out);
int32_t tsd_gelu(tsd_ctx* ctx, const float* x, int64_t n, float* out);
/* Upsample.forward, helpers/utils.mojo:1989-2010 under contract Q8: nearest x2 spatial. */
int32_t tsd_upsample2x(tsd_ctx* ctx, const float* x, int32_t c, int32_t h, int32_t w, float* out);
/* Softmax, helpers/utils.mojo:411-448. x,out (c,r,cc).  dim follows the reference numbering:
 * dim=2 normalises every column over the rows (Q3), dim=1 every row over the columns. */
int32_t tsd_softmax(tsd_ctx* ctx, const float* x, int32_t c, int32_t r, int32_t cc, int32_t dim,
                    float* out);
/* Self_Attention.forward, helpers/attention.mojo:26-65.  x,out (1,t,c); w_in (3c,c), b_in (3c)
 * or NULL; w_out (c,c), b_out (c) or NULL.  Head split by raw reshape (Q4); softmax axis per
 * the "softmax_axis" option. */
int32_t tsd_self_attention(tsd_ctx* ctx, const float* x, int32_t t, int32_t c, int32_t n_heads,
                           const float* w_in, const float* b_in, const float* w_out,
                           const float* b_out, float* out);
/* Cross_Attention.forward, helpers/attention.mojo:96-118.  x,out (1,t,c); context (1,tk,dc). */
int32_t tsd_cross_attention(tsd_ctx* ctx, const float* x, int32_t t, int32_t c, const float* context,
                            int32_t tk, int32_t dc, int32_t n_heads, const float* wq,
                            const float* bq, const float* wk, const float* bk, const float* wv,
                            const float* bv, const float* wo, const float* bo, float* out);
/* Attention core only (config 5 sweep): q (h,tq,d), k,v (h,tk,d) -> out (tq, h*d) merged. */
int32_t tsd_attention_core(tsd_ctx* ctx, const float* q, const float* k, const float* v, int32_t h,
                           int32_t tq, int32_t tk, int32_t d, float* out);
/* Same on DEVICE pointers, asynchronous on the context's stream (tsd_synchronize to wait). */
int32_t tsd_attention_core_dev(tsd_ctx* ctx, const float* q, const float* k, const float* v, int32_t h,
                               int32_t tq, int32_t tk, int32_t d, float* out);
/* DDPMSampler.step, sampler.mojo:75-109 (+ CFG combine pipeline.mojo:117-119 when eps_uncond
 * is non-NULL).  Scalars are the schedule values the host sampler computes (see
 * tsd_b200.DDPMSampler): x0 = (x - sqrt_1mab*eps)/sqrt_ab ; out = c0*x0 + c1*x + sigma*noise. */
int32_t tsd_sampler_step(tsd_ctx* ctx, const float* latents, const float* eps_cond,
                         const float* eps_uncond, float cfg_scale, const float* noise, float sqrt_ab,
                         float sqrt_1mab, float c0, float c1, float sigma, int64_t n, float* out);

/* DDPMSampler.add_noise, sampler.mojo:111-124 (img2img start, pipeline.mojo:78):
 * out = sqrt_ab * x + sqrt_1mab * noise, with sqrt_ab = sqrt(alphas_cumprod[t]) from the host sampler. */
int32_t tsd_sampler_add_noise(tsd_ctx* ctx, const float* x, const float* noise, float sqrt_ab,
                              float sqrt_1mab, int64_t n, float* out);

/* same, device pointers, asynchronous on the context's stream (out may alias latents) */
int32_t tsd_sampler_step_dev(tsd_ctx* ctx, const float* latents, const float* eps_cond,
                             const float* eps_uncond, float cfg_scale, const float* noise,
                             float sqrt_ab, float sqrt_1mab, float c0, float c1, float sigma, int64_t n,
                             float* out);

/* ---- Diffusion (time embedding MLP + UNet + output layer), diffusion.mojo:294-318 --------- */
typedef struct tsd_diffusion_config {
  int32_t latent_h, latent_w; /* 64 x 64 for 512 x 512 images (pipeline.mojo:60) */
  int32_t max_batch;          /* latents evaluated per forward (CFG = 2) */
  int32_t context_len;        /* 77 */
  int32_t context_dim;        /* 768 */
  int32_t mojo_alias_time;    /* 0 [default]; 1 reproduces SiLU^k(t_emb) aliasing */
  int32_t norm_affine;        /* 0 [default]: GroupNorm / LayerNorm own no tensors, as the reference's structs
                                 (helpers/utils.mojo:1817-1819, 2052-2061); 1: every norm carries a per-channel weight and
                                 bias (".weight" / ".bias" under the norm's field name, in struct order) - what a real
                                 checkpoint (README.md:44,55) needs, together with the options softmax_axis = 1,
                                 layernorm_mode = 1 and norm_eps_mode = 1 */
} tsd_diffusion_config;
int32_t tsd_diffusion_create(tsd_ctx* ctx, const tsd_diffusion_config* cfg, tsd_diffusion** out);
int32_t tsd_diffusion_destroy(tsd_diffusion* m);
/* number of float32 values tsd_diffusion_load_weights expects */
int64_t tsd_diffusion_num_params(const tsd_diffusion* m);
/* Flat fp32 blob: every parameter tensor in struct-declaration order, depth first, each layer
 * (weight, bias), in the reference layouts (diffusion.mojo:295-297,
 * 151-173, 25-30, 76-85).  The exact tensor list: tsd_diffusion_param_name(). */
int32_t tsd_diffusion_load_weights(tsd_diffusion* m, const float* blob, int64_t n_floats);
/* Deterministic synthetic weights generated on the device with the reference's init ranges
 * (conv U(+-1/sqrt(fan_in)), utils.mojo:1722-1724; linear U(+-1/sqrt(in))). */
int32_t tsd_diffusion_init_random(tsd_diffusion* m, uint64_t seed);
int32_t tsd_diffusion_param_count(const tsd_diffusion* m);
const char* tsd_diffusion_param_name(const tsd_diffusion* m, int32_t i, int64_t* offset,
                                     int64_t* numel);
/* copies parameter tensor i (reference layout) back to the host - used to feed the oracle */
int32_t tsd_diffusion_get_param(const tsd_diffusion* m, int32_t i, float* out);
/* Diffusion.forward, diffusion.mojo:309-318.  x,out (n,4,H,W); context (n_ctx,77,768) with
 * n_ctx = 1 (shared) or n; time (n_time,320) with n_time = 1 or n. */
int32_t tsd_diffusion_forward(tsd_diffusion* m, const float* x, const float* context, int32_t n_ctx,
                              const float* time, int32_t n_time, int32_t n, float* out);
/* One iteration of the reference denoising loop (pipeline.mojo:107-121) through host buffers in ONE call:
 * Diffusion.forward (diffusion.mojo:309-318; on [cond; uncond] when cfg != 0), the CFG combine (pipeline.mojo:117-119)
 * and DDPMSampler.step (sampler.mojo:75-109) with the schedule scalars of tsd_sampler_step.  latents / noise /
 * latents_out (n,4,H,W); time (320); context rows as tsd_generate_latents: (1|n) without cfg, (2|2n) with cfg, cond
 * rows first.  The K/V projections of `context` are reused while its bytes do not change (content hash), NULL reuses
 * the previous context outright.  Transfers per call: latents, time, noise in; latents out. */
int32_t tsd_diffusion_step(tsd_diffusion* m, const float* latents, const float* context, int32_t n_ctx,
                           const float* time, const float* noise, int32_t cfg, float cfg_scale, float sqrt_ab,
                           float sqrt_1mab, float c0, float c1, float sigma, int32_t n, float* latents_out);
int32_t tsd_diffusion_forward_dev(tsd_diffusion* m, const float* x, const float* context,
                                  int32_t n_ctx, const float* time, int32_t n_time, int32_t n,
                                  float* out);
/* per-family device time of the last profiled forward: families 0 gemm/conv, 1 attention,
 * 2 norm, 3 other.  ms[f], flops[f], launches[f] arrays of 4. */
int32_t tsd_diffusion_profile(tsd_diffusion* m, const float* x_dev, const float* context_dev,
                              int32_t n_ctx, const float* time_dev, int32_t n_time, int32_t n,
                              float* out_dev, double* ms, double* flops, int64_t* launches);

/* ---- VAE decoder, vae.mojo:221-250 (+ rescale/clamp pipeline.mojo:127 when rescale != 0) --- */
int32_t tsd_decoder_create(tsd_ctx* ctx, int32_t latent_h, int32_t latent_w, int32_t max_batch,
                           tsd_decoder** out);
/* Model flags of the *_create_ex entry points.  TSD_MODEL_NORM_AFFINE: every GroupNorm / LayerNorm of the model owns
 * a per-channel weight and bias (what real checkpoints carry; the reference's norms have none, helpers/utils.mojo:1833,
 * 1871-1873).  The extra parameters follow their block's tensors in the blob: Res_Block "<l>.groupnorm1", "<l>.groupnorm2"
 * after its convolutions, Attention_Block "<l>.groupnorm" after out_proj, the output norm ("l24" decoder, "l16" encoder)
 * before the last convolutions; CLIP "player<k>.layer1", "player<k>.layer3" after each layer's linears and "layernorm"
 * last.  tsd_*_param_name lists them.  flags = 0 is exactly tsd_*_create. */
#define TSD_MODEL_NORM_AFFINE 1u
int32_t tsd_decoder_create_ex(tsd_ctx* ctx, int32_t latent_h, int32_t latent_w, int32_t max_batch, uint32_t flags,
                              tsd_decoder** out);
int32_t tsd_decoder_destroy(tsd_decoder* m);
int64_t tsd_decoder_num_params(const tsd_decoder* m);
int32_t tsd_decoder_load_weights(tsd_decoder* m, const float* blob, int64_t n_floats);
int32_t tsd_decoder_init_random(tsd_decoder* m, uint64_t seed);
int32_t tsd_decoder_param_count(const tsd_decoder* m);
const char* tsd_decoder_param_name(const tsd_decoder* m, int32_t i, int64_t* offset, int64_t* numel);
int32_t tsd_decoder_get_param(const tsd_decoder* m, int32_t i, float* out);
/* z (n,4,H,W) -> img (n,3,8H,8W) */
int32_t tsd_decoder_forward(tsd_decoder* m, const float* z, int32_t n, int32_t rescale, float* img);
int32_t tsd_decoder_forward_dev(tsd_decoder* m, const float* z, int32_t n, int32_t rescale,
                                float* img);

/* ---- VAE encoder / img2img entry ----------------------------------------
 * Replaces Encoder.__init__ / Encoder.forward (vae.mojo:91-159): conv 3->128, Res_Blocks, three
 * two_stride_pad + stride-2 convs (one zero row below / column right, vae.mojo:115-116), Attention_Block,
 * GroupNorm(32)+SiLU, conv 512->8, conv 1x1 8->8 and metrics_evals (mean / log-variance chunks, clamp
 * (-30,20), latent = (mean + noise * sqrt(exp(logvar))) * 0.18215; vae.mojo:118-129).
 * img (n,3,8H,8W): values in (-1,1), or in (0,255) when rescale != 0 (the pipeline's
 * rescale((0,255),(-1,1)), pipeline.mojo:71).  noise, z: (n,4,H,W).  Parameters in struct order l1..l19. */
int32_t tsd_encoder_create(tsd_ctx* ctx, int32_t latent_h, int32_t latent_w, int32_t max_batch,
                           tsd_encoder** out);
int32_t tsd_encoder_create_ex(tsd_ctx* ctx, int32_t latent_h, int32_t latent_w, int32_t max_batch, uint32_t flags,
                              tsd_encoder** out);
int32_t tsd_encoder_destroy(tsd_encoder* m);
int64_t tsd_encoder_num_params(const tsd_encoder* m);
int32_t tsd_encoder_load_weights(tsd_encoder* m, const float* blob, int64_t n_floats);
int32_t tsd_encoder_init_random(tsd_encoder* m, uint64_t seed);
int32_t tsd_encoder_param_count(const tsd_encoder* m);
const char* tsd_encoder_param_name(const tsd_encoder* m, int32_t i, int64_t* offset, int64_t* numel);
int32_t tsd_encoder_get_param(const tsd_encoder* m, int32_t i, float* out);
int32_t tsd_encoder_forward(tsd_encoder* m, const float* img, const float* noise, int32_t n, int32_t rescale,
                            float* z);
int32_t tsd_encoder_forward_dev(tsd_encoder* m, const float* img, const float* noise, int32_t n,
                                int32_t rescale, float* z);

/* ---- CLIP text encoder ------------------------------------------------
 * Replaces CLIP.__init__ / CLIP.forward (clip.mojo:70-109): ClipEmbedding (token table + position),
 * 12 ClipPlayer layers (LayerNorm, causal 12-head Self_Attention, quick-GELU MLP; clip.mojo:36-53) and
 * the final LayerNorm.  n_vocab / n_layers <= 0 select the reference's 49408 / 12.  Parameters in
 * struct-declaration order: embedding.token_embedding.weight [n_vocab][768], embedding.position_embedding
 * [77][768], then per layer in_proj / out_proj / layer4 / layer5 (weight [out][in], bias).
 * tokens: up to 77 int32 ids, zero-padded to 77 as clip.mojo:90-92 does; context: [77][768] fp32.
 * "softmax_axis" / "layernorm_mode" apply as in the UNet; the causal mask is the standard triu(1). */
int32_t tsd_clip_create(tsd_ctx* ctx, int32_t n_vocab, int32_t n_layers, tsd_clip** out);
int32_t tsd_clip_create_ex(tsd_ctx* ctx, int32_t n_vocab, int32_t n_layers, uint32_t flags, tsd_clip** out);
int32_t tsd_clip_destroy(tsd_clip* m);
int64_t tsd_clip_num_params(const tsd_clip* m);
int32_t tsd_clip_load_weights(tsd_clip* m, const float* blob, int64_t n_floats);
int32_t tsd_clip_init_random(tsd_clip* m, uint64_t seed);
int32_t tsd_clip_param_count(const tsd_clip* m);
const char* tsd_clip_param_name(const tsd_clip* m, int32_t i, int64_t* offset, int64_t* numel);
int32_t tsd_clip_get_param(const tsd_clip* m, int32_t i, float* out);
int32_t tsd_clip_forward(tsd_clip* m, const int32_t* tokens, int32_t n_tokens, float* context);
int32_t tsd_clip_forward_dev(tsd_clip* m, const int32_t* tokens, int32_t n_tokens, float* context);

/* ---- whole denoising loop on the device, pipeline.mojo:86-122 ------------------------------- */
typedef struct tsd_loop_params {
  int32_t steps;           /* inference steps */
  int32_t cfg;             /* 0 / 1 */
  float cfg_scale;         /* pipeline.mojo:17 */
  const int32_t* timesteps;/* [steps]           (DDPMSampler.set_inference_timesteps) */
  const float* time_emb;   /* [steps][320]      (get_time_embedding, utils.mojo:353-370) */
  const float* coef;       /* [steps][5]: sqrt_ab, sqrt_1mab, c0, c1, sigma */
  const float* noise;      /* [steps][n][4*H*W] host, or NULL for no noise */
} tsd_loop_params;
/* latents (n,4,H,W) in/out; context (n_ctx,77,768): cond rows first, then (if cfg) uncond rows */
int32_t tsd_generate_latents(tsd_diffusion* m, const tsd_loop_params* lp, const float* latents_in,
                             const float* context, int32_t n_ctx, int32_t n, float* latents_out);

/* ---- host-side data formats: no device, no tsd_ctx ------------------------
 * Tokenizer(vocab_size, buf) (helpers/utils.mojo:236-249) over the tokenizer_clip.bin layout written by
 * tokenizer_creation.py:44-48: uint32 max_token_length, then per token float32 score, uint32 length,
 * bytes.  A truncated file is an error (the reference prints and carries on with zeros).  Token strings
 * are C strings as in the reference (cut at the first NUL byte). */
int32_t tsd_tokenizer_load(const char* path, int32_t vocab_size, tsd_tokenizer** out);
int32_t tsd_tokenizer_from_memory(const void* buf, int64_t size, int32_t vocab_size, tsd_tokenizer** out);
int32_t tsd_tokenizer_destroy(tsd_tokenizer* t);
int32_t tsd_tokenizer_vocab_size(const tsd_tokenizer* t);
int32_t tsd_tokenizer_max_token_length(const tsd_tokenizer* t);
const uint8_t* tsd_tokenizer_token(const tsd_tokenizer* t, int32_t id, int32_t* len, float* score);
/* Tokenizer.find (utils.mojo:276-292) including wrap (:197-206); -1 when absent. */
int32_t tsd_tokenizer_find(const tsd_tokenizer* t, const uint8_t* s, int32_t len);
/* bpe_encode (utils.mojo:294-327): one token per byte, then greedy merges of the adjacent pair whose
 * concatenation has the highest score (first wins on ties).  concat_mode 0 = str_concat as written
 * (utils.mojo:214-224: each output position receives the first byte of its source string), 1 = plain
 * concatenation (the intent).  A byte without a token: TSD_ERR_INVALID with the ids collected so far in
 * ids / n_out (the reference prints "Not a good prompt token" and returns that prefix).  n_out > cap:
 * TSD_ERR_OOM, nothing written.  The caller applies prompt.replace(" ", "</w>") (pipeline.mojo:39-40). */
int32_t tsd_tokenizer_encode(const tsd_tokenizer* t, const uint8_t* text, int32_t len, int32_t concat_mode,
                             int32_t* ids, int32_t cap, int32_t* n_out);
/* 8-bit PNG of the (c,h,w) planar float image pipeline.generate returns (values 0..255 after
 * rescale(clamp), pipeline.mojo:127-128; c = 1, 3 or 4): round half up, clamp, filter 0, stored deflate
 * blocks.  tsd_png_encode with out == NULL returns the size needed. */
int32_t tsd_png_encode(const float* img, int32_t c, int32_t h, int32_t w, uint8_t* out, int64_t cap,
                       int64_t* size);
int32_t tsd_png_write(const char* path, const float* img, int32_t c, int32_t h, int32_t w);

/* ---- checkpoint files ------------------------------------------------------
 * The reference has no loader (README.md:44,55 list it as future work); its weights are random.  These entry
 * points read the safetensors container (uint64 LE header length, JSON header, raw tensor bytes) so that the flat
 * blobs tsd_*_load_weights expect can be assembled from a checkpoint: F32 / F16 / BF16 / F64 tensors are delivered
 * as fp32 (exact for F16 / BF16).  Malformed headers, offsets outside the file or sizes that do not match
 * shape x dtype are TSD_ERR_INVALID. */
int32_t tsd_safetensors_open(const char* path, tsd_safetensors** out); /* memory-maps the file */
int32_t tsd_safetensors_from_memory(const void* buf, int64_t size, tsd_safetensors** out); /* copies buf */
int32_t tsd_safetensors_close(tsd_safetensors* st);
int32_t tsd_safetensors_count(const tsd_safetensors* st);
const char* tsd_safetensors_name(const tsd_safetensors* st, int32_t i);
int32_t tsd_safetensors_find(const tsd_safetensors* st, const char* name); /* index or -1 */
int32_t tsd_safetensors_info(const tsd_safetensors* st, int32_t i, char dtype[8], int32_t* rank, int64_t shape[8],
                             int64_t* numel);
int32_t tsd_safetensors_read_f32(const tsd_safetensors* st, int32_t i, float* out, int64_t cap);

/* ---- multi-GPU: batch sharding over the GPUs of one box ---------------------------------------
 * The reference has no distributed code; its only batching hints are pipeline.mojo:12 (a batch of images) and :96-105
 * (the two CFG halves).  The path shards by independent images: one process per GPU, rank r owns images r, r+R, ...,
 * every rank holds a full weight replica, the context is broadcast ONCE per prompt and there is no per-step collective.
 * NCCL (NVLink 5 / NVSwitch) is loaded at run time by tsd_dist_init when nranks > 1.
 * Bootstrap: rank 0 calls tsd_dist_unique_id and hands the 128 bytes to the other ranks by whatever means the host
 * program has (argv, a socket, MPI, torch.distributed), or every rank passes id = NULL and a rendezvous file path
 * (or TSD_DIST_RENDEZVOUS) on a file system they share: rank 0 writes the id there, the others wait for it. */
#define TSD_DIST_ID_BYTES 128
typedef struct tsd_dist tsd_dist;
int32_t tsd_dist_unique_id(uint8_t id[TSD_DIST_ID_BYTES]);
int32_t tsd_dist_init(tsd_ctx* ctx, int32_t nranks, int32_t rank, const uint8_t* id, const char* rendezvous_path,
                      tsd_dist** out);
int32_t tsd_dist_shutdown(tsd_dist* d);
int32_t tsd_dist_rank(const tsd_dist* d);
int32_t tsd_dist_size(const tsd_dist* d);
/* in place on a HOST buffer: root's n_floats values (the (n_ctx,77,768) context of pipeline.mojo:41-53) reach every rank */
int32_t tsd_dist_broadcast_context(tsd_dist* d, float* context, int64_t n_floats, int32_t root);
/* every rank contributes n_floats host values; `all` (root only) receives nranks x n_floats in rank order */
int32_t tsd_dist_gather(tsd_dist* d, const float* local, int64_t n_floats, float* all, int32_t root);
/* tsd_dist_broadcast_context(context) followed by tsd_generate_latents on this rank's n latents */
int32_t tsd_dist_generate(tsd_dist* d, tsd_diffusion* m, const tsd_loop_params* lp, const float* latents_in,
                          float* context, int32_t n_ctx, int32_t n, int32_t root, float* latents_out);

/* ---- tuning probes (synthetic device-resident operands, CUDA-event ms per launch) ------------ */
int32_t tsd_bench_gemm(tsd_ctx* ctx, int32_t m, int32_t n, int32_t k, int32_t batch, int32_t geglu,
                       int32_t force_bn, int32_t force_splits, int32_t iters, double* ms_out);
int32_t tsd_bench_conv(tsd_ctx* ctx, int32_t n, int32_t h, int32_t w, int32_t cin, int32_t cout,
                       int32_t k, int32_t stride, int32_t force_bn, int32_t force_splits,
                       int32_t iters, double* ms_out);
/* attention core (helpers/attention.mojo:46-62) on synthetic N(0,1)-like q,k,v: h heads, tq x tk, head dim d, the
 * context's softmax_axis / fused_attention options; ms per op (both launches of the fused kernel). */
int32_t tsd_bench_attention(tsd_ctx* ctx, int32_t h, int32_t tq, int32_t tk, int32_t d, int32_t iters,
                            double* ms_out);
/* GroupNorm(G, C) + SiLU on (n, C, H, W) device data: mode 0 stand-alone, 1 normalise-only over the partial statistics a
 * producing 3x3 convolution left (convolution untimed), 2 convolution + norm as a pair */
int32_t tsd_bench_norm(tsd_ctx* ctx, int32_t n, int32_t h, int32_t w, int32_t c, int32_t g, int32_t mode,
                       int32_t iters, double* ms_out);

#if defined(TSD_BUILD)
#pragma GCC visibility pop
#endif

#ifdef __cplusplus
}
#endif
#endif /* TSD_B200_H */
