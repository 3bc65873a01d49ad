// CLIP text encoder on the device (reference clip.mojo:5-109): the stage that produces the 77 x 768
// context of the denoising loop.  SURVEY.md section 8 row (f1): built from the same op layer as the
// UNet (tcgen05 GEMM, norm kernels, attention) once the per-step path met its parity bar.
//
// Semantics (extension of the SURVEY section 0 contract to CLIP):
//   * ClipEmbedding: token_table[token] + position (clip.mojo:17-20)
//   * ClipPlayer (clip.mojo:36-53): LayerNorm -> Self_Attention(12 heads, causal) -> + residue ->
//     LayerNorm -> Linear(768 -> 3072) -> x * sigmoid(1.702 x) -> Linear(3072 -> 768) -> + residue
//   * LayerNorm = GroupNorm(1, C) over the whole (C, T) tensor (Q5, default) or per token
//     ("layernorm_mode"); softmax over the query axis (Q3, default) or the key axis ("softmax_axis");
//     heads split by raw reshape (Q4) - the same deterministic deviations as in the UNet.
//   * Class G (accidents, replaced by the intent): triu(1) as written mirrors the row
//     index (utils.mojo:1588-1592) - the standard causal mask (key index > query index -> -inf) is
//     used; the quick-GELU is computed on aliased buffers in the reference (clip.mojo:49-50, Q2) -
//     x * sigmoid(1.702 x) is used.
#include "models.h"

#include <cstdio>
#include <string>

#include "c_api_internal.h"
#include "elementwise.cuh"

namespace tsd {

#define TRY(expr)            \
  do {                       \
    int rc__ = (expr);       \
    if (rc__) return rc__;   \
  } while (0)
#define LAUNCH(c, expr, what)                          \
  do {                                                 \
    if (!(c)->dry_run) {                               \
      int rc__ = (c)->check((expr), what);             \
      if (rc__) return rc__;                           \
      (c)->launches++;                                 \
    }                                                  \
  } while (0)

int Clip::create() {
  ps.c = c;
  if (n_vocab <= 0 || n_embed <= 0 || n_embed % (4 * n_heads) || n_tokens <= 0 || n_layers <= 0 || n_layers > kMaxLayers)
    return c->fail(TSD_ERR_INVALID, "clip: bad configuration");
  // parameter order = struct declaration order (clip.mojo:5-15, 23-34, 56-87); LayerNorm owns no learnable tensor
  // (GroupNorm's gamma is a scalar 1 and beta is never added, helpers/utils.mojo:1833, 1871-1873)
  tok = ps.add("embedding.token_embedding.weight", P_VEC, n_vocab * n_embed, 1, 1, 1.0f);
  pos = ps.add("embedding.position_embedding", P_VEC, n_tokens * n_embed, 1, 1, 0.0f);  // zero-initialised (clip.mojo:13-14)
  char nm[64];
  for (int l = 0; l < n_layers; ++l) {
    snprintf(nm, sizeof nm, "player%d", l + 1);
    const std::string b(nm);
    layer[l].in_proj = ps.add_linear(b + ".layer2.in_proj", n_embed, 3 * n_embed, true);
    layer[l].out_proj = ps.add_linear(b + ".layer2.out_proj", n_embed, n_embed, true);
    layer[l].fc1 = ps.add_linear(b + ".layer4", n_embed, 4 * n_embed, true);
    layer[l].fc2 = ps.add_linear(b + ".layer5", 4 * n_embed, n_embed, true);
    if (norm_affine) {  // real checkpoints: layer_norm1 / layer_norm2 weight + bias
      layer[l].ln1 = ps.add_norm(b + ".layer1", n_embed);
      layer[l].ln2 = ps.add_norm(b + ".layer3", n_embed);
    }
  }
  if (norm_affine) final_ln = ps.add_norm("layernorm", n_embed);
  if (cudaMalloc(&tokens_dev, sizeof(int) * n_tokens) != cudaSuccess ||
      cudaMalloc(&out_dev, sizeof(float) * (size_t)n_tokens * n_embed) != cudaSuccess)
    return c->fail(TSD_ERR_OOM, "clip: buffer allocation failed");
  return TSD_OK;
}

void Clip::destroy() {
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  if (tokens_dev) cudaFree(tokens_dev);
  if (out_dev) cudaFree(out_dev);
  tokens_dev = nullptr;
  out_dev = nullptr;
  ps.free_all();
}

static float* cw(Ctx* c, long long n) { return c->arena.alloc_n<float>((size_t)(n > 0 ? n : 1)); }

static int clip_linear(Ctx* c, const float* x, int M, int K, const float* w, const float* bias, int N, float* out,
                       const float* residual, int round_out, int split_n = 0, long long split_stride = 0) {
  GemmArgs g;
  g.A = x; g.M = M; g.K = K; g.lda = K;
  g.B = w; g.N = N; g.ldb = K;
  g.D = out; g.ldd = split_n > 0 ? split_n : N;
  g.bias = bias;
  g.residual = residual; g.ldr = g.ldd;
  g.split_n = split_n; g.split_stride = split_stride;
  g.round_tf32 = round_out;
  g.b_static = 1;
  return op_gemm(c, g);
}

// LayerNorm.forward (helpers/utils.mojo:2052-2061) on a (T, C) sequence, both modes
static int clip_layer_norm(Ctx* c, const float* x, float* y, int T, int C, const float* gamma, const float* beta, int round_tf32 = 1) {
  if (c->layernorm_mode == 0) return op_group_norm(c, x, y, 1, T, 1, C, 1, 1e-5f, gamma, beta, 1.0f, 0, 0, round_tf32);
  return op_group_norm(c, x, y, T, 1, 1, C, 1, 1e-5f, gamma, beta, 1.0f, 0, 0, round_tf32);
}

// CLIP.forward, clip.mojo:88-109: tokens_dev -> out_dev
int Clip::encode() {
  const int T = n_tokens, C = n_embed;
  const size_t mark = c->arena.mark();
  float* x = cw(c, (long long)T * C);
  float* v = cw(c, (long long)T * C);
  float* qkv = cw(c, 3LL * T * C);
  float* o = cw(c, (long long)T * C);
  float* x1 = cw(c, (long long)T * C);
  float* hbuf = cw(c, 4LL * T * C);
  if (!x || !v || !qkv || !o || !x1 || !hbuf) return c->fail(TSD_ERR_OOM, "workspace exhausted (clip)");
  LAUNCH(c, launch_clip_embed(tokens_dev, ps.w(tok), ps.w(pos), x, T, C, n_vocab, c->stream), "clip_embed");
  for (int l = 0; l < n_layers; ++l) {
    const Layer& w = layer[l];
    TRY(clip_layer_norm(c, x, v, T, C, ps.gamma(w.ln1), ps.beta(w.ln1)));
    // Self_Attention.forward with causal_mask (helpers/attention.mojo:26-65): in_proj, chunk, raw head split
    TRY(clip_linear(c, v, T, C, ps.w(w.in_proj), ps.w(w.in_proj + 1), 3 * C, qkv, nullptr, 1, C, (long long)T * C));
    AttnArgs at;
    at.Q = qkv; at.K = qkv + (long long)T * C; at.V = qkv + 2LL * T * C;
    at.batch = 1; at.heads = n_heads; at.Tq = T; at.Tk = T; at.d = C / n_heads; at.O = o;
    at.softmax_axis = c->softmax_axis;
    at.causal = 1;
    TRY(op_attention(c, at));
    TRY(clip_linear(c, o, T, C, ps.w(w.out_proj), ps.w(w.out_proj + 1), C, x1, x, 0));
    TRY(clip_layer_norm(c, x1, v, T, C, ps.gamma(w.ln2), ps.beta(w.ln2)));
    TRY(clip_linear(c, v, T, C, ps.w(w.fc1), ps.w(w.fc1 + 1), 4 * C, hbuf, nullptr, 0));
    LAUNCH(c, launch_unary(hbuf, hbuf, 4LL * T * C, UNARY_QUICKGELU, 1.0f, c->stream), "quick_gelu");
    TRY(clip_linear(c, hbuf, T, 4 * C, ps.w(w.fc2), ps.w(w.fc2 + 1), C, x, x1, 0));
  }
  // final LayerNorm; un-rounded (this is the model output, not a GEMM operand)
  TRY(clip_layer_norm(c, x, out_dev, T, C, ps.gamma(final_ln), ps.beta(final_ln), 0));
  c->arena.release_to(mark);
  return TSD_OK;
}

size_t Clip::workspace_bytes() const {
  Clip* self = const_cast<Clip*>(this);
  Ctx* cc = self->c;
  const bool was = cc->dry_run;
  cc->dry_run = true;
  cc->arena.set_virtual(true);
  int rc = self->encode();
  size_t hw = cc->arena.high_water();
  cc->arena.set_virtual(false);
  cc->dry_run = was;
  return rc ? 0 : hw + (64u << 20);
}

int Clip::forward(const int32_t* tokens, int n, float* out, bool host_ptrs) {
  if (!ps.loaded) return c->fail(TSD_ERR_STATE, "clip: forward before load_weights / init_random");
  // n = 0 is the empty prompt (the default backup_prompt, pipeline.mojo:15): 77 zero ids after the padding
  if (!out || n < 0 || n > n_tokens || (!tokens && n)) return c->fail(TSD_ERR_INVALID, "clip: bad token buffer");
  cudaSetDevice(c->device);
  // reshaped_tokens = zeros(77); first n entries = tokens (clip.mojo:90-92)
  std::vector<int32_t> padded(n_tokens, 0);
  if (host_ptrs) {
    for (int i = 0; i < n; ++i) {
      if (tokens[i] < 0 || tokens[i] >= n_vocab) return c->fail(TSD_ERR_INVALID, "clip: token id outside the vocabulary");
      padded[i] = tokens[i];
    }
  }
  const size_t need = workspace_bytes();
  if (need == 0) return TSD_ERR_OOM;
  if (need > c->arena.capacity()) {
    cudaStreamSynchronize(c->stream);
    if (c->arena.reserve(need) != TSD_OK) return c->fail(TSD_ERR_OOM, "clip: workspace allocation failed");
  }
  c->arena.reset();
  if (host_ptrs) {
    TRY(c->check(cudaMemcpyAsync(tokens_dev, padded.data(), sizeof(int32_t) * n_tokens, cudaMemcpyHostToDevice, c->stream),
                 "copy tokens"));
    TRY(c->check(cudaStreamSynchronize(c->stream), "clip token upload"));  // `padded` is a stack-lifetime buffer
  } else {
    TRY(c->check(cudaMemsetAsync(tokens_dev, 0, sizeof(int32_t) * n_tokens, c->stream), "clear tokens"));
    if (n) TRY(c->check(cudaMemcpyAsync(tokens_dev, tokens, sizeof(int32_t) * n, cudaMemcpyDeviceToDevice, c->stream), "copy tokens"));
  }
  TRY(encode());
  TRY(c->check(cudaMemcpyAsync(out, out_dev, sizeof(float) * (size_t)n_tokens * n_embed,
                               host_ptrs ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, c->stream), "copy context"));
  if (host_ptrs) TRY(c->check(cudaStreamSynchronize(c->stream), "clip forward sync"));
  return TSD_OK;
}

}  // namespace tsd
