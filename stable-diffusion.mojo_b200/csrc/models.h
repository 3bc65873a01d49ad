// Device-side model graphs composed from the op layer (runtime.h):
//   Diffusion  = Time_Embedding + UNet + UNet_Output_Layer   (reference diffusion.mojo:5-318)
//   Decoder    = VAE decoder                                  (reference vae.mojo:5-67, 162-250)
// plus the denoising loop (reference pipeline.mojo:86-122, sampler.mojo:75-109).
//
// Parameters are enumerated in the reference's struct-declaration order (depth first, each
// layer as weight then bias) and accepted in the reference's layouts (conv OIHW, linear
// [out][in]); on the device conv kernels are re-laid as [O][kh*kw][I] and every GEMM weight is
// rounded to TF32 once at load time (the tensor core would otherwise truncate it on every read).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "runtime.h"

struct tsd_ctx;

namespace tsd {

enum ParamKind { P_CONV_W = 0, P_LIN_W = 1, P_VEC = 2 };

struct Param {
  std::string name;
  int kind = P_VEC;
  int O = 0, I = 0, KK = 1;  // conv: (O, I, k*k) ; linear: (O, I) ; vector: (O)
  long long numel = 0;
  long long offset = 0;      // float offset in the reference-order blob
  float init_scale = 0.f;    // synthetic init: U(-s, s)
  float init_const = 0.f;    // synthetic init when init_scale == 0: this value (norm weights: 1)
  float* dev = nullptr;
};

struct ParamStore {
  Ctx* c = nullptr;
  std::vector<Param> params;
  long long total = 0;
  float* block = nullptr;  // one cudaMalloc for all parameters
  bool loaded = false;

  int add(const std::string& name, int kind, int O, int I, int KK, float init_scale);
  // conv (weight, bias) -> index of the weight; bias is index + 1
  int add_conv(const std::string& name, int cin, int cout, int k);
  // linear (weight[, bias]) -> index of the weight
  int add_linear(const std::string& name, int in_f, int out_f, bool bias);
  // per-channel norm (weight = 1, bias = 0 at synthetic init) -> index of the weight; bias is index + 1
  int add_norm(const std::string& name, int channels);
  const float* gamma(int i) const { return i >= 0 ? params[i].dev : nullptr; }
  const float* beta(int i) const { return i >= 0 ? params[i + 1].dev : nullptr; }
  int allocate();
  int load(const float* blob, long long n_floats);
  int init_random(uint64_t seed);
  int get(int i, float* host_out);
  const float* w(int i) const { return params[i].dev; }
  // row sums of a [O][I] weight matrix (over the values the tensor core sees), computed on first use after
  // every (re)load; for LayerNorm folding (GemmArgs::ln_fold)
  const float* rowsum(int i);
  std::vector<float*> rowsum_dev;
  std::vector<int> rowsum_gen;
  // [O][KK*I + I2] = conv weight rows followed by the rows of a 1x1 convolution with the same O, and the sum of
  // the two biases: the operands of the fused ResBlock tail (conv2 + skip convolution in one GEMM).  Built on first
  // use after every (re)load, keyed by the conv weight's index.
  int fused_skip(int conv_w, int skip_w, const float** w, const float** b);
  // [conv_out . geglu2 | conv_out] and conv_out . b2 + b_out: the two linears that end an attention block as ONE GEMM
  int fused_ffn_out(int w2, int wc, const float** w, const float** b);
  std::vector<float*> ffn_dev;
  std::vector<int> ffn_gen;
  std::vector<float*> fskip_dev;
  std::vector<int> fskip_gen;
  int gen = 0;  // bumped by load / init_random
  void free_all();
};

struct Act {
  float* p = nullptr;
  int N = 0, H = 0, W = 0, C = 0;
  NormHint ns;  // statistics left by the producer for the consumer's first norm (ns.stats == nullptr: none)
  long long numel() const { return (long long)N * H * W * C; }
  long long pixels() const { return (long long)N * H * W; }
};

struct ResBlockW {
  int cin = 0, cout = 0;
  int conv1 = -1, lin_t = -1, conv2 = -1, skip = -1;  // param indices of the weights
  int gn1 = -1, gn2 = -1;  // per-channel norm weights (norm_affine models only)
  int groups = 32;
};
struct AttnBlockW {
  int heads = 8, C = 0;
  int conv_in = -1, in_proj = -1, out_proj = -1, q = -1, k = -1, v = -1, o = -1, geglu1 = -1, geglu2 = -1,
      conv_out = -1;
  int gn = -1, ln1 = -1, ln2 = -1, ln3 = -1;  // per-channel norm weights (norm_affine models only)
};

struct GraphSlot {
  cudaGraphExec_t exec = nullptr;
  int n = 0, n_ctx = 0, n_time = 0;
  int epoch = -1;
  int weights_gen = -1;  // ParamStore::gen the graph was captured with: derived weight buffers (fused skip weights,
                         // LayerNorm-fold row sums) are only refreshed by an eager pass, so a reload re-captures
  const void* arena_base = nullptr;
  long long nodes = 0;  // kernels per replay (bench gpu_launches accounting)
};

struct Diffusion {
  tsd_ctx* h = nullptr;
  Ctx* c = nullptr;
  tsd_diffusion_config cfg{};
  ParamStore ps;
  int te1 = -1, te2 = -1, conv_in = -1, final_conv = -1, down1 = -1, down2 = -1, final_gn = -1;
  ResBlockW res[9];
  AttnBlockW attn[9];
  // persistent device buffers (fixed addresses -> CUDA-graph friendly)
  float* x_in = nullptr;      // [max_batch][4][H][W]  reference layout staging
  float* out_nchw = nullptr;  // [max_batch][4][H][W]
  float* ctx_in = nullptr;    // [max_batch][77][768]
  float* time_in = nullptr;   // [max_batch][320]
  float* kctx[9] = {};        // [n_ctx][77][C] per attention block
  float* vctx[9] = {};
  float* tbias[9] = {};       // [max_batch][Cout] : Linear(SiLU(t_emb)) + linear bias + conv1 bias
  float* temb = nullptr;      // [max_batch][1280]
  float* x_nhwc = nullptr;    // [max_batch][H][W][4]
  float* eps_nhwc = nullptr;  // [max_batch][H][W][4]
  GraphSlot graph;
  struct LoopCache {
    GraphSlot slot;
    int cfg = 0, has_noise = 0;
    int steps = 0;  // the loop-lifetime buffers are carved behind tables whose size depends on the step count
    float scale = 0.f;
    const void* top = nullptr;
  } loop_cache;
  std::vector<size_t> ws_cache;
  int ws_epoch = -1;
  bool ctx_ready = false;
  int ctx_n = 0;
  // host-buffer entry points: the K/V projections of a context are kept while the caller keeps passing the same
  // bytes (the reference rebuilds everything per call, pipeline.mojo:83-122; its loop passes one context for all steps)
  uint64_t ctx_hash = 0;
  int ctx_hash_n = 0, ctx_hash_gen = -1;
  float* noise_in = nullptr;  // [max_batch][4][H][W] staging of tsd_diffusion_step
  float* lat_out = nullptr;

  int create();
  void destroy();
  size_t workspace_bytes(int n) const;
  int prepare_context(int n_ctx);                 // ctx_in -> kctx/vctx
  int prepare_time(int n_time, const float* time_dev /*[n_time][320]*/, float* const* tb_out,
                   int rows_stride);              // time embedding MLP + 9 ResBlock time linears
  int unet(int n, int n_ctx, int n_time);         // x_nhwc -> eps_nhwc (uses kctx/vctx/tbias)
  int core(int n, int n_ctx, int n_time);         // time_in, x_in -> out_nchw: time MLP + layout changes + unet
  int forward_dev(const float* x, const float* context, int n_ctx, const float* time, int n_time, int n,
                  float* out, bool host_ptrs);
  int run_unet_graph(int n, int n_ctx, int n_time);
  // context rows (host) -> ctx_in + kctx/vctx unless the same bytes are already projected
  int upload_context_cached(const float* context_host, int n_ctx);
  // one iteration of the reference loop (pipeline.mojo:107-121): Diffusion.forward (x2 with CFG), combine, DDPMSampler.step
  int step_host(const float* latents, const float* context, int n_ctx, const float* time, const float* noise, int cfg,
                float cfg_scale, const float coef[5], int n, float* latents_out);
};

struct Decoder {
  tsd_ctx* h = nullptr;
  Ctx* c = nullptr;
  int latent_h = 0, latent_w = 0, max_batch = 1;
  ParamStore ps;
  int l1 = -1, l2 = -1, l10 = -1, l15 = -1, l20 = -1, l26 = -1, attn_in = -1, attn_out = -1;
  int norm_affine = 0;           // 1: every GroupNorm owns a per-channel weight and bias (real checkpoints)
  int attn_gn = -1, out_gn = -1; // norm_affine only: Attention_Block's GroupNorm, l24
  ResBlockW res[14];  // l3, l5..l8, l11..l13, l16..l18, l21..l23
  float* z_in = nullptr;     // [max_batch][4][h][w]
  float* img_out = nullptr;  // [max_batch][3][8h][8w]
  float* ping = nullptr;
  float* pong = nullptr;
  size_t pp_elems = 0;
  GraphSlot graph;

  int create();
  void destroy();
  size_t workspace_bytes(int n) const;
  int decode(int n, int rescale);  // z_in -> img_out
  int forward(const float* z, int n, int rescale, float* img, bool host_ptrs);
};

// VAE Encoder (vae.mojo:70-159): image (n,3,8h,8w) + noise (n,4,h,w) -> latent (n,4,h,w)
struct Encoder {
  tsd_ctx* h = nullptr;
  Ctx* c = nullptr;
  int latent_h = 0, latent_w = 0, max_batch = 1;
  ParamStore ps;
  int l1 = -1, l4 = -1, l7 = -1, l10 = -1, l18 = -1, l19 = -1, attn_in = -1, attn_out = -1;
  int norm_affine = 0;           // 1: every GroupNorm owns a per-channel weight and bias (real checkpoints)
  int attn_gn = -1, out_gn = -1; // norm_affine only: Attention_Block's GroupNorm, l16
  ResBlockW res[10];  // l2 l3 l5 l6 l8 l9 l11 l12 l13 l15
  float* img_in = nullptr;    // [max_batch][3][8h][8w]
  float* noise_in = nullptr;  // [max_batch][4][h][w]
  float* z_out = nullptr;     // [max_batch][4][h][w]
  float* ping = nullptr;
  float* pong = nullptr;
  size_t pp_elems = 0;
  GraphSlot graph;

  int create();
  void destroy();
  size_t workspace_bytes(int n) const;
  int encode(int n, int rescale);  // img_in, noise_in -> z_out
  int forward(const float* img, const float* noise, int n, int rescale, float* z, bool host_ptrs);
};

// CLIP text encoder (clip.mojo:5-109) - models_clip.cu
struct Clip {
  static constexpr int kMaxLayers = 24;
  tsd_ctx* h = nullptr;
  Ctx* c = nullptr;
  int n_vocab = 49408, n_embed = 768, n_tokens = 77, n_heads = 12, n_layers = 12;  // clip.mojo:71-83
  ParamStore ps;
  int tok = -1, pos = -1;
  int norm_affine = 0;  // 1: every LayerNorm owns a per-channel weight and bias (real checkpoints)
  int final_ln = -1;
  struct Layer {
    int in_proj = -1, out_proj = -1, fc1 = -1, fc2 = -1;
    int ln1 = -1, ln2 = -1;  // norm_affine only
  } layer[kMaxLayers];
  int* tokens_dev = nullptr;
  float* out_dev = nullptr;  // [n_tokens][n_embed]

  int create();
  void destroy();
  size_t workspace_bytes() const;
  int encode();  // tokens_dev -> out_dev
  int forward(const int32_t* tokens, int n, float* out, bool host_ptrs);
};

int generate_latents(Diffusion& m, const tsd_loop_params& lp, const float* latents_in, const float* context,
                     int n_ctx, int n, float* latents_out);
// next: in = (G, eps) of the norm that will consume `out`; out = next->stats when the last conv produced them
int res_block(Ctx* c, const ParamStore& ps, const ResBlockW& w, const Act& x, const float* tbias,
              int tbias_stride, float eps, float* out, NormHint* next = nullptr);

}  // namespace tsd

struct tsd_diffusion {
  tsd::Diffusion m;
};
struct tsd_decoder {
  tsd::Decoder m;
};
struct tsd_clip {
  tsd::Clip m;
};
struct tsd_encoder {
  tsd::Encoder m;
};
