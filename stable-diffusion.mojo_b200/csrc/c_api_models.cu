// extern "C" surface, model level: Diffusion, Decoder, the denoising loop (include/tsd_b200.h).
#include <cstring>
#include <mutex>
#include <new>

#include "../../include/tsd_b200.h"
#include "c_api_internal.h"
#include "elementwise.cuh"
#include "models.h"

using namespace tsd;

namespace {
struct Guard {
  std::unique_lock<std::mutex> lk;
  explicit Guard(tsd_ctx* h) : lk(h->mu) { cudaSetDevice(h->c->device); }
};
}  // namespace

extern "C" {

// ---- Diffusion -------------------------------------------------------------------------------
int32_t tsd_diffusion_create(tsd_ctx* h, const tsd_diffusion_config* cfg, tsd_diffusion** out) {
  if (!h || !cfg || !out) return TSD_ERR_INVALID;
  *out = nullptr;
  Guard g(h);
  tsd_diffusion* d = new (std::nothrow) tsd_diffusion();
  if (!d) return h->c->fail(TSD_ERR_OOM, "diffusion: host allocation failed");
  d->m.h = h;
  d->m.c = h->c;
  d->m.cfg = *cfg;
  int rc = d->m.create();
  if (rc) {
    d->m.destroy();
    delete d;
    return rc;
  }
  *out = d;
  return TSD_OK;
}
int32_t tsd_diffusion_destroy(tsd_diffusion* d) {
  if (!d) return TSD_ERR_INVALID;
  {
    Guard g(d->m.h);
    d->m.destroy();
  }
  delete d;
  return TSD_OK;
}
int64_t tsd_diffusion_num_params(const tsd_diffusion* d) { return d ? d->m.ps.total : 0; }
int32_t tsd_diffusion_load_weights(tsd_diffusion* d, const float* blob, int64_t n_floats) {
  if (!d || !blob) return TSD_ERR_INVALID;
  Guard g(d->m.h);
  return d->m.ps.load(blob, n_floats);
}
int32_t tsd_diffusion_init_random(tsd_diffusion* d, uint64_t seed) {
  if (!d) return TSD_ERR_INVALID;
  Guard g(d->m.h);
  return d->m.ps.init_random(seed);
}
int32_t tsd_diffusion_param_count(const tsd_diffusion* d) { return d ? (int32_t)d->m.ps.params.size() : 0; }
const char* tsd_diffusion_param_name(const tsd_diffusion* d, int32_t i, int64_t* offset, int64_t* numel) {
  if (!d || i < 0 || i >= (int32_t)d->m.ps.params.size()) return nullptr;
  const Param& p = d->m.ps.params[i];
  if (offset) *offset = p.offset;
  if (numel) *numel = p.numel;
  return p.name.c_str();
}
int32_t tsd_diffusion_get_param(const tsd_diffusion* d, int32_t i, float* out) {
  if (!d || !out) return TSD_ERR_INVALID;
  tsd_diffusion* dd = const_cast<tsd_diffusion*>(d);
  Guard g(dd->m.h);
  return dd->m.ps.get(i, out);
}
int32_t tsd_diffusion_forward(tsd_diffusion* d, const float* x, const float* context, int32_t n_ctx,
                              const float* time, int32_t n_time, int32_t n, float* out) {
  if (!d) return TSD_ERR_INVALID;
  Guard g(d->m.h);
  int rc = d->m.forward_dev(x, context, n_ctx, time, n_time, n, out, true);
  if (rc) {
    cudaStreamSynchronize(d->m.c->stream);
    cudaGetLastError();
  }
  return rc;
}
int32_t tsd_diffusion_step(tsd_diffusion* d, const float* latents, const float* context, int32_t n_ctx,
                           const float* time, const float* noise, int32_t cfg, float cfg_scale, float sqrt_ab,
                           float sqrt_1mab, float c0, float c1, float sigma, int32_t n, float* latents_out) {
  if (!d) return TSD_ERR_INVALID;
  Guard g(d->m.h);
  const float coef[5] = {sqrt_ab, sqrt_1mab, c0, c1, sigma};
  int rc = d->m.step_host(latents, context, n_ctx, time, noise, cfg, cfg_scale, coef, n, latents_out);
  if (rc) {
    cudaStreamSynchronize(d->m.c->stream);
    cudaGetLastError();
  }
  return rc;
}
int32_t tsd_diffusion_forward_dev(tsd_diffusion* d, const float* x, const float* context, int32_t n_ctx,
                                  const float* time, int32_t n_time, int32_t n, float* out) {
  if (!d) return TSD_ERR_INVALID;
  Guard g(d->m.h);
  return d->m.forward_dev(x, context, n_ctx, time, n_time, n, out, false);
}

int32_t tsd_diffusion_profile(tsd_diffusion* d, const float* x_dev, const float* context_dev, int32_t n_ctx,
                              const float* time_dev, int32_t n_time, int32_t n, float* out_dev, double* ms,
                              double* flops, int64_t* launches) {
  if (!d || !ms || !flops || !launches) return TSD_ERR_INVALID;
  Guard g(d->m.h);
  Ctx* c = d->m.c;
  KernelTimer timer;
  // one untimed eager pass (sizes the workspace, sets kernel attributes), then block the stream for
  // 20 ms so the whole timed pass is enqueued before the GPU starts it: the per-launch events then
  // measure device time, not host launch gaps
  int rc0 = d->m.forward_dev(x_dev, context_dev, n_ctx, time_dev, n_time, n, out_dev, false);
  if (rc0) return rc0;
  cudaStreamSynchronize(c->stream);
  launch_spin(20 * 1000 * 1000LL, c->stream);
  c->timer = &timer;
  int rc = d->m.forward_dev(x_dev, context_dev, n_ctx, time_dev, n_time, n, out_dev, false);
  int rc2 = c->check(cudaStreamSynchronize(c->stream), "profile sync");
  c->timer = nullptr;
  for (int f = 0; f < FAM_COUNT; ++f) {
    ms[f] = 0;
    flops[f] = 0;
    launches[f] = 0;
  }
  if (!rc && !rc2) {
    for (auto& r : timer.recs) {
      float t = 0;
      cudaEventElapsedTime(&t, r.a, r.b);
      ms[r.family] += t;
      flops[r.family] += r.flops;
      launches[r.family] += 1;
    }
  }
  for (auto e : timer.pool) cudaEventDestroy(e);
  return rc ? rc : rc2;
}

// ---- Decoder ---------------------------------------------------------------------------------
int32_t tsd_decoder_create(tsd_ctx* h, int32_t latent_h, int32_t latent_w, int32_t max_batch, tsd_decoder** out) {
  return tsd_decoder_create_ex(h, latent_h, latent_w, max_batch, 0, out);
}
int32_t tsd_decoder_create_ex(tsd_ctx* h, int32_t latent_h, int32_t latent_w, int32_t max_batch, uint32_t flags,
                              tsd_decoder** out) {
  if (!h || !out) return TSD_ERR_INVALID;
  *out = nullptr;
  Guard g(h);
  if (flags & ~(uint32_t)TSD_MODEL_NORM_AFFINE) return h->c->fail(TSD_ERR_INVALID, "decoder: unknown flag");
  tsd_decoder* d = new (std::nothrow) tsd_decoder();
  if (!d) return h->c->fail(TSD_ERR_OOM, "decoder: host allocation failed");
  d->m.h = h;
  d->m.c = h->c;
  d->m.latent_h = latent_h;
  d->m.latent_w = latent_w;
  d->m.max_batch = max_batch;
  d->m.norm_affine = (flags & TSD_MODEL_NORM_AFFINE) ? 1 : 0;
  int rc = d->m.create();
  if (rc) {
    d->m.destroy();
    delete d;
    return rc;
  }
  *out = d;
  return TSD_OK;
}
int32_t tsd_decoder_destroy(tsd_decoder* d) {
  if (!d) return TSD_ERR_INVALID;
  {
    Guard g(d->m.h);
    d->m.destroy();
  }
  delete d;
  return TSD_OK;
}
int64_t tsd_decoder_num_params(const tsd_decoder* d) { return d ? d->m.ps.total : 0; }
int32_t tsd_decoder_load_weights(tsd_decoder* d, const float* blob, int64_t n_floats) {
  if (!d || !blob) return TSD_ERR_INVALID;
  Guard g(d->m.h);
  return d->m.ps.load(blob, n_floats);
}
int32_t tsd_decoder_init_random(tsd_decoder* d, uint64_t seed) {
  if (!d) return TSD_ERR_INVALID;
  Guard g(d->m.h);
  return d->m.ps.init_random(seed);
}
int32_t tsd_decoder_param_count(const tsd_decoder* d) { return d ? (int32_t)d->m.ps.params.size() : 0; }
const char* tsd_decoder_param_name(const tsd_decoder* d, int32_t i, int64_t* offset, int64_t* numel) {
  if (!d || i < 0 || i >= (int32_t)d->m.ps.params.size()) return nullptr;
  const Param& p = d->m.ps.params[i];
  if (offset) *offset = p.offset;
  if (numel) *numel = p.numel;
  return p.name.c_str();
}
int32_t tsd_decoder_get_param(const tsd_decoder* d, int32_t i, float* out) {
  if (!d || !out) return TSD_ERR_INVALID;
  tsd_decoder* dd = const_cast<tsd_decoder*>(d);
  Guard g(dd->m.h);
  return dd->m.ps.get(i, out);
}
int32_t tsd_decoder_forward(tsd_decoder* d, const float* z, int32_t n, int32_t rescale, float* img) {
  if (!d) return TSD_ERR_INVALID;
  Guard g(d->m.h);
  int rc = d->m.forward(z, n, rescale, img, true);
  if (rc) {
    cudaStreamSynchronize(d->m.c->stream);
    cudaGetLastError();
  }
  return rc;
}
int32_t tsd_decoder_forward_dev(tsd_decoder* d, const float* z, int32_t n, int32_t rescale, float* img) {
  if (!d) return TSD_ERR_INVALID;
  Guard g(d->m.h);
  return d->m.forward(z, n, rescale, img, false);
}

// ---- VAE Encoder (vae.mojo:70-159) ---------------------------------------------------------------------------------
int32_t tsd_encoder_create(tsd_ctx* h, int32_t latent_h, int32_t latent_w, int32_t max_batch, tsd_encoder** out) {
  return tsd_encoder_create_ex(h, latent_h, latent_w, max_batch, 0, out);
}
int32_t tsd_encoder_create_ex(tsd_ctx* h, int32_t latent_h, int32_t latent_w, int32_t max_batch, uint32_t flags,
                              tsd_encoder** out) {
  if (!h || !out) return TSD_ERR_INVALID;
  *out = nullptr;
  Guard g(h);
  if (flags & ~(uint32_t)TSD_MODEL_NORM_AFFINE) return h->c->fail(TSD_ERR_INVALID, "encoder: unknown flag");
  tsd_encoder* d = new (std::nothrow) tsd_encoder();
  if (!d) return h->c->fail(TSD_ERR_OOM, "encoder: host allocation failed");
  d->m.h = h;
  d->m.c = h->c;
  d->m.latent_h = latent_h;
  d->m.latent_w = latent_w;
  d->m.max_batch = max_batch;
  d->m.norm_affine = (flags & TSD_MODEL_NORM_AFFINE) ? 1 : 0;
  int rc = d->m.create();
  if (rc) {
    d->m.destroy();
    delete d;
    return rc;
  }
  *out = d;
  return TSD_OK;
}
int32_t tsd_encoder_destroy(tsd_encoder* d) {
  if (!d) return TSD_ERR_INVALID;
  {
    Guard g(d->m.h);
    d->m.destroy();
  }
  delete d;
  return TSD_OK;
}
int64_t tsd_encoder_num_params(const tsd_encoder* d) { return d ? d->m.ps.total : 0; }
int32_t tsd_encoder_load_weights(tsd_encoder* d, const float* blob, int64_t n_floats) {
  if (!d || !blob) return TSD_ERR_INVALID;
  Guard g(d->m.h);
  return d->m.ps.load(blob, n_floats);
}
int32_t tsd_encoder_init_random(tsd_encoder* d, uint64_t seed) {
  if (!d) return TSD_ERR_INVALID;
  Guard g(d->m.h);
  return d->m.ps.init_random(seed);
}
int32_t tsd_encoder_param_count(const tsd_encoder* d) { return d ? (int32_t)d->m.ps.params.size() : 0; }
const char* tsd_encoder_param_name(const tsd_encoder* d, int32_t i, int64_t* offset, int64_t* numel) {
  if (!d || i < 0 || i >= (int32_t)d->m.ps.params.size()) return nullptr;
  const Param& p = d->m.ps.params[i];
  if (offset) *offset = p.offset;
  if (numel) *numel = p.numel;
  return p.name.c_str();
}
int32_t tsd_encoder_get_param(const tsd_encoder* d, int32_t i, float* out) {
  if (!d || !out) return TSD_ERR_INVALID;
  tsd_encoder* dd = const_cast<tsd_encoder*>(d);
  Guard g(dd->m.h);
  return dd->m.ps.get(i, out);
}
int32_t tsd_encoder_forward(tsd_encoder* d, const float* img, const float* noise, int32_t n, int32_t rescale,
                            float* z) {
  if (!d) return TSD_ERR_INVALID;
  Guard g(d->m.h);
  int rc = d->m.forward(img, noise, n, rescale, z, true);
  if (rc) {
    cudaStreamSynchronize(d->m.c->stream);
    cudaGetLastError();
  }
  return rc;
}
int32_t tsd_encoder_forward_dev(tsd_encoder* d, const float* img, const float* noise, int32_t n,
                                int32_t rescale, float* z) {
  if (!d) return TSD_ERR_INVALID;
  Guard g(d->m.h);
  return d->m.forward(img, noise, n, rescale, z, false);
}

// ---- CLIP text encoder (clip.mojo:56-109) ------------------------------------------------------
int32_t tsd_clip_create(tsd_ctx* h, int32_t n_vocab, int32_t n_layers, tsd_clip** out) {
  return tsd_clip_create_ex(h, n_vocab, n_layers, 0, out);
}
int32_t tsd_clip_create_ex(tsd_ctx* h, int32_t n_vocab, int32_t n_layers, uint32_t flags, tsd_clip** out) {
  if (!h || !out) return TSD_ERR_INVALID;
  *out = nullptr;
  Guard g(h);
  if (flags & ~(uint32_t)TSD_MODEL_NORM_AFFINE) return h->c->fail(TSD_ERR_INVALID, "clip: unknown flag");
  tsd_clip* d = new (std::nothrow) tsd_clip();
  if (!d) return TSD_ERR_OOM;
  d->m.h = h;
  d->m.c = h->c;
  if (n_vocab > 0) d->m.n_vocab = n_vocab;
  if (n_layers > 0) d->m.n_layers = n_layers;
  d->m.norm_affine = (flags & TSD_MODEL_NORM_AFFINE) ? 1 : 0;
  int rc = d->m.create();
  if (rc) {
    d->m.destroy();
    delete d;
    return rc;
  }
  *out = d;
  return TSD_OK;
}
int32_t tsd_clip_destroy(tsd_clip* d) {
  if (!d) return TSD_ERR_INVALID;
  {
    Guard g(d->m.h);
    d->m.destroy();
  }
  delete d;
  return TSD_OK;
}
int64_t tsd_clip_num_params(const tsd_clip* d) { return d ? d->m.ps.total : 0; }
int32_t tsd_clip_load_weights(tsd_clip* d, const float* blob, int64_t n_floats) {
  if (!d || !blob) return TSD_ERR_INVALID;
  Guard g(d->m.h);
  return d->m.ps.load(blob, n_floats);
}
int32_t tsd_clip_init_random(tsd_clip* d, uint64_t seed) {
  if (!d) return TSD_ERR_INVALID;
  Guard g(d->m.h);
  return d->m.ps.init_random(seed);
}
int32_t tsd_clip_param_count(const tsd_clip* d) { return d ? (int32_t)d->m.ps.params.size() : 0; }
const char* tsd_clip_param_name(const tsd_clip* d, int32_t i, int64_t* offset, int64_t* numel) {
  if (!d || i < 0 || i >= (int32_t)d->m.ps.params.size()) return nullptr;
  const Param& p = d->m.ps.params[i];
  if (offset) *offset = p.offset;
  if (numel) *numel = p.numel;
  return p.name.c_str();
}
int32_t tsd_clip_get_param(const tsd_clip* d, int32_t i, float* out) {
  if (!d || !out) return TSD_ERR_INVALID;
  tsd_clip* dd = const_cast<tsd_clip*>(d);
  Guard g(dd->m.h);
  return dd->m.ps.get(i, out);
}
int32_t tsd_clip_forward(tsd_clip* d, const int32_t* tokens, int32_t n_tokens, float* context) {
  if (!d) return TSD_ERR_INVALID;
  Guard g(d->m.h);
  int rc = d->m.forward(tokens, n_tokens, context, true);
  if (rc) {
    cudaStreamSynchronize(d->m.c->stream);
    cudaGetLastError();
  }
  return rc;
}
int32_t tsd_clip_forward_dev(tsd_clip* d, const int32_t* tokens, int32_t n_tokens, float* context) {
  if (!d) return TSD_ERR_INVALID;
  Guard g(d->m.h);
  return d->m.forward(tokens, n_tokens, context, false);
}

// ---- loop -----------------------------------------------------------------------------------
int32_t tsd_generate_latents(tsd_diffusion* d, const tsd_loop_params* lp, const float* latents_in,
                             const float* context, int32_t n_ctx, int32_t n, float* latents_out) {
  if (!d || !lp) return TSD_ERR_INVALID;
  Guard g(d->m.h);
  int rc = generate_latents(d->m, *lp, latents_in, context, n_ctx, n, latents_out);
  if (rc) {
    cudaStreamSynchronize(d->m.c->stream);
    cudaGetLastError();
  }
  return rc;
}

}  // extern "C"
