// extern "C" surface (include/tsd_b200.h): context + op-level entry points.
// Model-level entry points (Diffusion / Decoder / loop) live in c_api_models.cu.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>

#include "../../include/tsd_b200.h"
#include "c_api_internal.h"
#include "pdl.cuh"
#include "attention_tcgen05.cuh"
#include "elementwise.cuh"
#include "runtime.h"

using namespace tsd;

static std::string g_init_error;

namespace tsd {

// Scoped helper for the host-buffer entry points: lock, reset the arena, stage buffers.
HostCall::HostCall(tsd_ctx* h, size_t reserve_bytes) : h_(h), c(h->c), lock_(h->mu) {
  cudaSetDevice(c->device);
  c->arena.reset();
  if (reserve_bytes > c->arena.capacity()) {
    cudaStreamSynchronize(c->stream);
    if (c->arena.reserve(reserve_bytes + (64u << 20)) != TSD_OK)
      rc = c->fail(TSD_ERR_OOM, "workspace allocation failed");
  }
}
float* HostCall::dev(size_t n) {
  if (rc) return nullptr;
  float* p = c->arena.alloc_n<float>(n ? n : 1);
  if (!p) rc = c->fail(TSD_ERR_OOM, "workspace exhausted");
  return p;
}
float* HostCall::upload(const float* host, size_t n) {
  float* p = dev(n);
  if (p && host) {
    int e = c->check(cudaMemcpyAsync(p, host, n * sizeof(float), cudaMemcpyHostToDevice, c->stream),
                     "H2D copy");
    if (e) rc = e;
  }
  return p;
}
void HostCall::download(float* host, const float* d, size_t n) {
  if (rc) return;
  int e = c->check(cudaMemcpyAsync(host, d, n * sizeof(float), cudaMemcpyDeviceToHost, c->stream),
                   "D2H copy");
  if (e) rc = e;
}
int HostCall::finish() {
  if (rc) {
    cudaStreamSynchronize(c->stream);
    cudaGetLastError();
    return rc;
  }
  return c->check(cudaStreamSynchronize(c->stream), "stream synchronize");
}
void HostCall::run(int code) {
  if (!rc && code) rc = code;
}
void HostCall::cu(cudaError_t e, const char* what) {
  if (!rc) {
    int r = c->check(e, what);
    if (r) rc = r;
    else c->launches++;
  }
}

}  // namespace tsd

extern "C" {

int32_t tsd_init(int32_t device, tsd_ctx** out) {
  if (!out) return TSD_ERR_INVALID;
  *out = nullptr;
  Ctx* c = nullptr;
  std::string err;
  int rc = ctx_create(device, &c, &err);
  if (rc) {
    g_init_error = err;
    return rc;
  }
  tsd_ctx* h = new tsd_ctx();
  h->c = c;
  *out = h;
  // tuning / A-B switches from the environment: TSD_OPT_<option name>=<int>  (same names as tsd_set_option)
  static const char* kEnvOpts[] = {"producer_stats", "norm_v2", "norm_cluster", "gemm_kmerge", "gemm_deep_b", "fuse_ffn_out", "tune_defer_penalty_us", "ln_fold", "fuse_skip", "conv_stride_tma", "defer_reduce", "virtual_concat", "gn_partial", "gn_partial_max_groups", "splitk_fixup", "pdl", "autotune", "conv_halo", "halo_min_w", "halo_min_h", "tune_verbose", "tune_flush", "gemm_cg", "force_bn", "force_splits", "fused_attention", "attn_v2", "attn_poly", "splitk_cluster", "splitk_cluster_max"};
  for (const char* name : kEnvOpts) {
    const std::string key = std::string("TSD_OPT_") + name;
    if (const char* v = getenv(key.c_str())) tsd_set_option(h, name, atoi(v));
  }
  return TSD_OK;
}

int32_t tsd_shutdown(tsd_ctx* h) {
  if (!h) return TSD_ERR_INVALID;
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  ctx_destroy(h->c);
  delete h;
  return TSD_OK;
}

const char* tsd_last_error(const tsd_ctx* h) {
  if (!h) return g_init_error.c_str();
  return h->c->last_error.c_str();
}

int32_t tsd_synchronize(tsd_ctx* h) {
  if (!h) return TSD_ERR_INVALID;
  std::lock_guard<std::mutex> g(h->mu);
  cudaSetDevice(h->c->device);
  return h->c->check(cudaStreamSynchronize(h->c->stream), "stream synchronize");
}

static int* option_slot(tsd_ctx* h, const char* name) {
  if (!strcmp(name, "softmax_axis")) return &h->c->softmax_axis;
  if (!strcmp(name, "layernorm_mode")) return &h->c->layernorm_mode;
  if (!strcmp(name, "norm_eps_mode")) return &h->c->norm_eps_mode;
  if (!strcmp(name, "fused_attention")) return &h->c->fused_attention;
  if (!strcmp(name, "attn_v2")) return &h->c->attn_v2;
  if (!strcmp(name, "attn_poly")) return &h->c->attn_poly;
  if (!strcmp(name, "cuda_graph")) return &h->use_graph;
  if (!strcmp(name, "force_bn")) return &h->c->force_bn;
  if (!strcmp(name, "force_splits")) return &h->c->force_splits;
  if (!strcmp(name, "force_stages")) return &h->c->force_stages;
  if (!strcmp(name, "gemm_debug")) return &h->c->gemm_debug;
  if (!strcmp(name, "gemm_cg")) return &h->c->gemm_cg;
  if (!strcmp(name, "producer_stats")) return &h->c->producer_stats;
  if (!strcmp(name, "norm_v2")) return &h->c->norm_v2;
  if (!strcmp(name, "norm_cluster")) return &h->c->norm_cluster;
  if (!strcmp(name, "gemm_kmerge")) return &h->c->gemm_kmerge;
  if (!strcmp(name, "gemm_deep_b")) return &h->c->gemm_deep_b;
  if (!strcmp(name, "fuse_ffn_out")) return &h->c->fuse_ffn_out;
  if (!strcmp(name, "tune_defer_penalty_us")) return &h->c->tune_defer_penalty_us;
  if (!strcmp(name, "ln_fold")) return &h->c->ln_fold;
  if (!strcmp(name, "fuse_skip")) return &h->c->fuse_skip;
  if (!strcmp(name, "conv_stride_tma")) return &h->c->conv_stride_tma;
  if (!strcmp(name, "defer_reduce")) return &h->c->defer_reduce;
  if (!strcmp(name, "virtual_concat")) return &h->c->virtual_concat;
  if (!strcmp(name, "gn_partial")) return &h->c->gn_partial;
  if (!strcmp(name, "gn_partial_max_groups")) return &h->c->gn_partial_max_groups;
  if (!strcmp(name, "splitk_fixup")) return &h->c->splitk_fixup;
  if (!strcmp(name, "splitk_cluster")) return &h->c->splitk_cluster;
  if (!strcmp(name, "splitk_cluster_max")) return &h->c->splitk_cluster_max;
  if (!strcmp(name, "pdl")) return &tsd::pdl_enabled();
  if (!strcmp(name, "autotune")) return &h->c->autotune;
  if (!strcmp(name, "conv_halo")) return &h->c->conv_halo;
  if (!strcmp(name, "halo_min_w")) return &h->c->halo_min_w;
  if (!strcmp(name, "halo_min_h")) return &h->c->halo_min_h;
  if (!strcmp(name, "tune_verbose")) return &h->c->tune_verbose;
  if (!strcmp(name, "tune_flush")) return &h->c->tune_flush;
  if (!strcmp(name, "bench_stats_groups")) return &h->c->bench_stats_groups;
  return nullptr;
}
int32_t tsd_set_option(tsd_ctx* h, const char* name, int32_t value) {
  if (!h || !name) return TSD_ERR_INVALID;
  std::lock_guard<std::mutex> g(h->mu);
  int* s = option_slot(h, name);
  if (!s) return h->c->fail(TSD_ERR_INVALID, std::string("unknown option ") + name);
  *s = value;
  h->option_epoch++;
  return TSD_OK;
}
int32_t tsd_get_option(tsd_ctx* h, const char* name, int32_t* value) {
  if (!h || !name || !value) return TSD_ERR_INVALID;
  int* s = option_slot(h, name);
  if (!s) return h->c->fail(TSD_ERR_INVALID, std::string("unknown option ") + name);
  *value = *s;
  return TSD_OK;
}
int64_t tsd_launch_count(const tsd_ctx* h) { return h ? h->c->launches : 0; }

// CUDA-event stopwatch on the context's own stream (the stream every kernel of this library is
// launched on; an event recorded on another stream would not see them).
int32_t tsd_timer_start(tsd_ctx* h) {
  if (!h) return TSD_ERR_INVALID;
  std::lock_guard<std::mutex> g(h->mu);
  cudaSetDevice(h->c->device);
  if (!h->ev0) {
    int rc = h->c->check(cudaEventCreate(&h->ev0), "event create");
    if (rc) return rc;
    rc = h->c->check(cudaEventCreate(&h->ev1), "event create");
    if (rc) return rc;
  }
  return h->c->check(cudaEventRecord(h->ev0, h->c->stream), "event record");
}
int32_t tsd_timer_stop(tsd_ctx* h, double* ms) {
  if (!h || !ms || !h->ev0) return TSD_ERR_INVALID;
  std::lock_guard<std::mutex> g(h->mu);
  cudaSetDevice(h->c->device);
  int rc = h->c->check(cudaEventRecord(h->ev1, h->c->stream), "event record");
  if (rc) return rc;
  rc = h->c->check(cudaEventSynchronize(h->ev1), "event synchronize");
  if (rc) return rc;
  float t = 0.f;
  rc = h->c->check(cudaEventElapsedTime(&t, h->ev0, h->ev1), "event elapsed");
  *ms = t;
  return rc;
}

// ---------------------------------------------------------------------------------------
// op level
// ---------------------------------------------------------------------------------------
int32_t tsd_conv2d(tsd_ctx* h, const float* x, int32_t n, int32_t cin, int32_t H, int32_t W,
                   const float* weight, const float* bias, int32_t cout, int32_t k, int32_t pad,
                   int32_t stride, float* out) {
  return tsd_conv2d_pad(h, x, n, cin, H, W, weight, bias, cout, k, pad, pad, stride, out);
}

int32_t tsd_conv2d_pad(tsd_ctx* h, const float* x, int32_t n, int32_t cin, int32_t H, int32_t W,
                       const float* weight, const float* bias, int32_t cout, int32_t k, int32_t pad,
                       int32_t pad_hi, int32_t stride, float* out) {
  if (!h || !x || !weight || !out) return TSD_ERR_INVALID;
  if (n <= 0 || cin <= 0 || cout <= 0 || H <= 0 || W <= 0 || k <= 0 || stride <= 0 || pad < 0 || pad_hi < 0)
    return h->c->fail(TSD_ERR_INVALID, "conv2d: non-positive dimension");
  const int Ho = conv_out_dim(H, k, pad, stride, pad_hi), Wo = conv_out_dim(W, k, pad, stride, pad_hi);
  if (Ho <= 0 || Wo <= 0) return h->c->fail(TSD_ERR_INVALID, "conv2d: kernel larger than padded input");
  const size_t nx = (size_t)n * cin * H * W, nw = (size_t)cout * cin * k * k,
               no = (size_t)n * cout * Ho * Wo;
  const size_t col = stride > 1 ? (size_t)n * Ho * Wo * 9 * cin : 0;
  HostCall hc(h, (2 * nx + 2 * nw + 2 * no + col + cout) * 4 + 8 * (size_t)n * Ho * Wo * cout * 4);
  float* x_nchw = hc.upload(x, nx);
  float* w_oihw = hc.upload(weight, nw);
  float* b = bias ? hc.upload(bias, cout) : nullptr;
  float* x_nhwc = hc.dev(nx);
  float* w_ohwi = hc.dev(nw);
  float* o_nhwc = hc.dev(no);
  float* o_nchw = hc.dev(no);
  if (!hc.rc) {
    Ctx* c = hc.c;
    hc.cu(launch_nchw_to_nhwc(x_nchw, x_nhwc, n, cin, H * W, c->stream), "nchw_to_nhwc");
    hc.cu(launch_oihw_to_ohwi(w_oihw, w_ohwi, cout, cin, k * k, c->stream), "oihw_to_ohwi");
    ConvArgs a;
    a.x = x_nhwc; a.N = n; a.H = H; a.W = W; a.Cin = cin; a.Cout = cout;
    a.k = k; a.pad = pad; a.pad_hi = pad_hi; a.stride = stride; a.w = w_ohwi; a.bias = b; a.out = o_nhwc;
    a.force_bn = c->force_bn; a.force_splits = c->force_splits;
    if (!hc.rc) hc.run(op_conv2d(c, a));
    hc.cu(launch_nhwc_to_nchw(o_nhwc, o_nchw, n, cout, Ho * Wo, c->stream), "nhwc_to_nchw");
    hc.download(out, o_nchw, no);
  }
  return hc.finish();
}

int32_t tsd_linear(tsd_ctx* h, const float* x, int32_t b, int32_t t, int32_t in_f,
                   const float* weight, const float* bias, int32_t out_f, float* out) {
  if (!h || !x || !weight || !out) return TSD_ERR_INVALID;
  if (b <= 0 || t <= 0 || in_f <= 0 || out_f <= 0)
    return h->c->fail(TSD_ERR_INVALID, "linear: non-positive dimension");
  const size_t M = (size_t)b * t;
  HostCall hc(h, (M * in_f + (size_t)out_f * in_f + M * out_f + out_f) * 4 + (16u << 20));
  float* xd = hc.upload(x, M * in_f);
  float* wd = hc.upload(weight, (size_t)out_f * in_f);
  float* bd = bias ? hc.upload(bias, out_f) : nullptr;
  float* od = hc.dev(M * out_f);
  if (!hc.rc) {
    Ctx* c = hc.c;
    if (in_f % 4 || out_f % 4 || M <= 4) {
      hc.cu(launch_gemv(xd, (int)M, in_f, wd, bd, nullptr, od, out_f, 0, 0, c->stream), "gemv");
    } else {
      GemmArgs g;
      g.A = xd; g.M = (int)M; g.K = in_f; g.lda = in_f;
      g.B = wd; g.N = out_f; g.ldb = in_f;
      g.D = od; g.ldd = out_f; g.bias = bd;
      g.force_bn = c->force_bn; g.force_splits = c->force_splits;
      hc.run(op_gemm(c, g));
    }
    hc.download(out, od, M * out_f);
  }
  return hc.finish();
}

int32_t tsd_matmul(tsd_ctx* h, const float* a, const float* b, int32_t cc, int32_t m, int32_t k,
                   int32_t n, float* out) {
  if (!h || !a || !b || !out) return TSD_ERR_INVALID;
  if (cc <= 0 || m <= 0 || k <= 0 || n <= 0)
    return h->c->fail(TSD_ERR_INVALID, "matmul: non-positive dimension");
  const int kp = (k + 3) / 4 * 4, np = (n + 3) / 4 * 4;
  const size_t na = (size_t)cc * m * k, nb = (size_t)cc * k * n;
  HostCall hc(h, (2 * na + nb + (size_t)cc * n * kp + 2 * (size_t)cc * m * np + (size_t)cc * m * kp) * 4 +
                     (16u << 20));
  float* ad = hc.upload(a, na);
  float* bd = hc.upload(b, nb);
  float* bt = hc.dev((size_t)cc * n * kp);
  float* od = hc.dev((size_t)cc * m * np);
  float* oc = hc.dev((size_t)cc * m * n);
  float* ap = ad;
  if (!hc.rc) {
    Ctx* c = hc.c;
    hc.cu(cudaMemsetAsync(bt, 0, (size_t)cc * n * kp * 4, c->stream), "memset");
    hc.cu(launch_transpose_ld(bd, bt, cc, k, n, kp, c->stream), "transpose");
    if (kp != k) {  // re-pitch A rows to a multiple of 4 floats (TMA stride rule)
      ap = hc.dev((size_t)cc * m * kp);
      if (!hc.rc) {
        hc.cu(cudaMemsetAsync(ap, 0, (size_t)cc * m * kp * 4, c->stream), "memset");
        hc.cu(cudaMemcpy2DAsync(ap, (size_t)kp * 4, ad, (size_t)k * 4, (size_t)k * 4, (size_t)cc * m,
                                cudaMemcpyDeviceToDevice, c->stream),
              "repitch");
      }
    }
    GemmArgs g;
    g.A = ap; g.M = m; g.K = kp; g.lda = kp; g.a_bs = (long long)m * kp;
    g.B = bt; g.N = n; g.ldb = kp; g.b_bs = (long long)n * kp;
    g.batch = cc;
    g.D = od; g.ldd = np; g.d_bs = (long long)m * np;
    g.force_bn = c->force_bn;
    if (!hc.rc) hc.run(op_gemm(c, g));
    if (np != n) {
      hc.cu(cudaMemcpy2DAsync(oc, (size_t)n * 4, od, (size_t)np * 4, (size_t)n * 4, (size_t)cc * m,
                              cudaMemcpyDeviceToDevice, c->stream),
            "repitch");
      hc.download(out, oc, (size_t)cc * m * n);
    } else {
      hc.download(out, od, (size_t)cc * m * n);
    }
  }
  return hc.finish();
}

int32_t tsd_groupnorm(tsd_ctx* h, const float* x, int32_t n, int32_t cch, int32_t H, int32_t W,
                      int32_t groups, float eps, const float* gamma, const float* beta, float* out) {
  if (!h || !x || !out) return TSD_ERR_INVALID;
  if (n <= 0 || cch <= 0 || H <= 0 || W <= 0 || groups <= 0)
    return h->c->fail(TSD_ERR_INVALID, "groupnorm: non-positive dimension");
  if (cch % groups)
    return h->c->fail(TSD_ERR_INVALID,
                      "groupnorm: number of channels does not evenly divide the number of groups");
  const size_t nx = (size_t)n * cch * H * W;
  HostCall hc(h, 4 * nx * 4 + (16u << 20));
  float* x_nchw = hc.upload(x, nx);
  float* gd = gamma ? hc.upload(gamma, cch) : nullptr;
  float* bd = beta ? hc.upload(beta, cch) : nullptr;
  float* x_nhwc = hc.dev(nx);
  float* y_nhwc = hc.dev(nx);
  float* y_nchw = hc.dev(nx);
  if (!hc.rc) {
    Ctx* c = hc.c;
    hc.cu(launch_nchw_to_nhwc(x_nchw, x_nhwc, n, cch, H * W, c->stream), "nchw_to_nhwc");
    if (!hc.rc) hc.run(op_group_norm(c, x_nhwc, y_nhwc, n, H, W, cch, groups, eps, gd, bd, 1.0f, 0, 0, 0));
    hc.cu(launch_nhwc_to_nchw(y_nhwc, y_nchw, n, cch, H * W, c->stream), "nhwc_to_nchw");
    hc.download(out, y_nchw, nx);
  }
  return hc.finish();
}

int32_t tsd_layernorm(tsd_ctx* h, const float* x, int32_t cch, int32_t t, float* out) {
  // LayerNorm(n) = GroupNorm(1, n) on a (C,T,1) Matrix (helpers/utils.mojo:2052-2061)
  if (!h) return TSD_ERR_INVALID;
  if (h->c->layernorm_mode == 0) return tsd_groupnorm(h, x, 1, cch, t, 1, 1, 1e-5f, nullptr, nullptr, out);
  // per-token switch: every token is its own "image" of one pixel
  if (!x || !out || cch <= 0 || t <= 0) return h->c->fail(TSD_ERR_INVALID, "layernorm: bad arguments");
  const size_t nx = (size_t)cch * t;
  HostCall hc(h, 4 * nx * 4 + (16u << 20));
  float* x_ct = hc.upload(x, nx);
  float* x_tc = hc.dev(nx);
  float* y_tc = hc.dev(nx);
  float* y_ct = hc.dev(nx);
  if (!hc.rc) {
    Ctx* c = hc.c;
    hc.cu(launch_nchw_to_nhwc(x_ct, x_tc, 1, cch, t, c->stream), "transpose");
    if (!hc.rc) hc.run(op_group_norm(c, x_tc, y_tc, t, 1, 1, cch, 1, 1e-5f, nullptr, nullptr, 1.0f, 0, 0, 0));
    hc.cu(launch_nhwc_to_nchw(y_tc, y_ct, 1, cch, t, c->stream), "transpose");
    hc.download(out, y_ct, nx);
  }
  return hc.finish();
}

static int32_t unary_host(tsd_ctx* h, const float* x, int64_t n, float* out, int op) {
  if (!h || !x || !out || n <= 0) return TSD_ERR_INVALID;
  HostCall hc(h, 2 * (size_t)n * 4 + (1u << 20));
  float* xd = hc.upload(x, n);
  float* yd = hc.dev(n);
  if (!hc.rc) {
    hc.cu(launch_unary(xd, yd, n, op, 1.0f, hc.c->stream), "unary");
    hc.download(out, yd, n);
  }
  return hc.finish();
}
int32_t tsd_silu(tsd_ctx* h, const float* x, int64_t n, float* out) {
  return unary_host(h, x, n, out, UNARY_SILU);
}
int32_t tsd_gelu(tsd_ctx* h, const float* x, int64_t n, float* out) {
  return unary_host(h, x, n, out, UNARY_GELU);
}

int32_t tsd_upsample2x(tsd_ctx* h, const float* x, int32_t cch, int32_t H, int32_t W, float* out) {
  if (!h || !x || !out || cch <= 0 || H <= 0 || W <= 0) return TSD_ERR_INVALID;
  // channel-major input: treat every (c) plane as a 1-channel image, padded to 4 "channels" is
  // not needed - use NHWC with C = 4-aligned only; otherwise transpose path
  const size_t nx = (size_t)cch * H * W;
  HostCall hc(h, 12 * nx * 4 + (1u << 20));
  float* x_chw = hc.upload(x, nx);
  const int cp = (cch + 3) / 4 * 4;
  float* x_hwc = hc.dev((size_t)cp * H * W);
  float* y_hwc = hc.dev((size_t)cp * H * W * 4);
  float* y_chw = hc.dev(nx * 4);
  if (!hc.rc) {
    Ctx* c = hc.c;
    if (cp == cch) {
      hc.cu(launch_nchw_to_nhwc(x_chw, x_hwc, 1, cch, H * W, c->stream), "nchw_to_nhwc");
      hc.cu(launch_upsample2x(x_hwc, y_hwc, 1, H, W, cch, c->stream), "upsample2x");
      hc.cu(launch_nhwc_to_nchw(y_hwc, y_chw, 1, cch, 4 * H * W, c->stream), "nhwc_to_nchw");
    } else {
      // planes as a batch of single-channel... fall back: C planes of (H,W,1) -> use stats-free
      // norm_apply path is overkill; do it per plane with the transpose-free kernel below
      hc.cu(launch_upsample2x_planar(x_chw, y_chw, cch, H, W, c->stream), "upsample2x_planar");
    }
    hc.download(out, y_chw, nx * 4);
  }
  return hc.finish();
}

int32_t tsd_softmax(tsd_ctx* h, const float* x, int32_t cc, int32_t r, int32_t cols, int32_t dim,
                    float* out) {
  if (!h || !x || !out || cc <= 0 || r <= 0 || cols <= 0) return TSD_ERR_INVALID;
  if (dim != 1 && dim != 2)
    return h->c->fail(TSD_ERR_INVALID, "softmax: invalid dimension (only dim=1 and dim=2 are device ops)");
  const size_t nx = (size_t)cc * r * cols;
  HostCall hc(h, (nx + 2 * (size_t)cc * cols) * 4 + (1u << 20));
  float* xd = hc.upload(x, nx);
  float* st = hc.dev(2 * (size_t)cc * cols);
  if (!hc.rc) {
    // reference numbering (helpers/utils.mojo:423-445): dim=2 -> per column over rows; dim=1 -> per row
    hc.cu(launch_softmax(xd, cc, r, cols, cols, dim == 2 ? 0 : 1, 1.0f, st, hc.c->stream), "softmax");
    hc.c->launches++;
    hc.download(out, xd, nx);
  }
  return hc.finish();
}

int32_t tsd_attention_core(tsd_ctx* h, const float* q, const float* k, const float* v, int32_t heads,
                           int32_t tq, int32_t tk, int32_t d, float* out) {
  if (!h || !q || !k || !v || !out || heads <= 0 || tq <= 0 || tk <= 0 || d <= 0) return TSD_ERR_INVALID;
  if (d % 4) return h->c->fail(TSD_ERR_INVALID, "attention: head dim must be a multiple of 4");
  const size_t nq = (size_t)heads * tq * d, nk = (size_t)heads * tk * d;
  size_t need = (2 * nq + 2 * nk) * 4 + (32u << 20);
  if (!h->c->fused_attention || !attention_fused_supported(d, 0)) need += ((size_t)heads * tq * (tk + 4) + (size_t)heads * d * (tk + 4)) * 4;
  HostCall hc(h, need);
  float* qd = hc.upload(q, nq);
  float* kd = hc.upload(k, nk);
  float* vd = hc.upload(v, nk);
  float* od = hc.dev(nq);
  if (!hc.rc) {
    AttnArgs a;
    a.Q = qd; a.K = kd; a.V = vd; a.heads = heads; a.Tq = tq; a.Tk = tk; a.d = d; a.O = od;
    a.softmax_axis = hc.c->softmax_axis;
    hc.run(op_attention(hc.c, a));
    hc.download(out, od, nq);
  }
  return hc.finish();
}

int32_t tsd_attention_core_dev(tsd_ctx* h, const float* q, const float* k, const float* v, int32_t heads,
                               int32_t tq, int32_t tk, int32_t d, float* out) {
  if (!h || !q || !k || !v || !out || heads <= 0 || tq <= 0 || tk <= 0 || d <= 0) return TSD_ERR_INVALID;
  if (d % 4) return h->c->fail(TSD_ERR_INVALID, "attention: head dim must be a multiple of 4");
  size_t need = (size_t)heads * ((size_t)tq + tk) * 64 + (32u << 20);  // softmax statistics
  if (!h->c->fused_attention || !attention_fused_supported(d, 0)) need += ((size_t)heads * tq * (tk + 4) + (size_t)heads * d * (tk + 4)) * 4;
  HostCall hc(h, need);
  if (!hc.rc) {
    AttnArgs a;
    a.Q = q; a.K = k; a.V = v; a.heads = heads; a.Tq = tq; a.Tk = tk; a.d = d; a.O = out;
    a.softmax_axis = hc.c->softmax_axis;
    hc.run(op_attention(hc.c, a));
  }
  return hc.rc;  // asynchronous: no synchronize
}

int32_t tsd_self_attention(tsd_ctx* h, const float* x, int32_t t, int32_t cch, int32_t n_heads,
                           const float* w_in, const float* b_in, const float* w_out,
                           const float* b_out, float* out) {
  if (!h || !x || !w_in || !w_out || !out) return TSD_ERR_INVALID;
  if (t <= 0 || cch <= 0 || n_heads <= 0 || cch % n_heads || (cch / n_heads) % 4 || cch % 16)
    return h->c->fail(TSD_ERR_INVALID, "self_attention: unsupported shape");
  const size_t nx = (size_t)t * cch;
  size_t need = (6 * nx + 4 * (size_t)cch * cch + 4 * cch) * 4 + (32u << 20);
  if (!h->c->fused_attention || !attention_fused_supported(cch / n_heads, 0)) need += ((size_t)n_heads * t * (t + 4) * 2) * 4;
  HostCall hc(h, need);
  float* xd = hc.upload(x, nx);
  float* wi = hc.upload(w_in, (size_t)3 * cch * cch);
  float* bi = b_in ? hc.upload(b_in, 3 * cch) : nullptr;
  float* wo = hc.upload(w_out, (size_t)cch * cch);
  float* bo = b_out ? hc.upload(b_out, cch) : nullptr;
  float* qkv = hc.dev(3 * nx);
  float* o = hc.dev(nx);
  float* y = hc.dev(nx);
  if (!hc.rc) {
    Ctx* c = hc.c;
    GemmArgs g;  // in_proj + chunk(2,3): thirds of the last axis go to separate [T][C] buffers
    g.A = xd; g.M = t; g.K = cch; g.lda = cch;
    g.B = wi; g.N = 3 * cch; g.ldb = cch;
    g.D = qkv; g.ldd = cch; g.bias = bi; g.split_n = cch; g.split_stride = (long long)nx;
    hc.run(op_gemm(c, g));
    AttnArgs a;
    a.Q = qkv; a.K = qkv + nx; a.V = qkv + 2 * nx;
    a.heads = n_heads; a.Tq = t; a.Tk = t; a.d = cch / n_heads; a.O = o;
    a.softmax_axis = c->softmax_axis;
    if (!hc.rc) hc.run(op_attention(c, a));
    GemmArgs p;
    p.A = o; p.M = t; p.K = cch; p.lda = cch;
    p.B = wo; p.N = cch; p.ldb = cch;
    p.D = y; p.ldd = cch; p.bias = bo;
    if (!hc.rc) hc.run(op_gemm(c, p));
    hc.download(out, y, nx);
  }
  return hc.finish();
}

int32_t tsd_cross_attention(tsd_ctx* h, const float* x, int32_t t, int32_t cch, const float* context,
                            int32_t tk, int32_t dc, int32_t n_heads, const float* wq,
                            const float* bq, const float* wk, const float* bk, const float* wv,
                            const float* bv, const float* wo, const float* bo, float* out) {
  if (!h || !x || !context || !wq || !wk || !wv || !wo || !out) return TSD_ERR_INVALID;
  if (t <= 0 || tk <= 0 || cch <= 0 || dc <= 0 || n_heads <= 0 || cch % n_heads ||
      (cch / n_heads) % 4 || cch % 16 || dc % 4)
    return h->c->fail(TSD_ERR_INVALID, "cross_attention: unsupported shape");
  const size_t nx = (size_t)t * cch, nc = (size_t)tk * dc, nkv = (size_t)tk * cch;
  size_t need = (5 * nx + nc + 2 * nkv + 2 * (size_t)cch * cch + 2 * (size_t)cch * dc + 4 * cch) * 4 +
                (32u << 20);
  if (!h->c->fused_attention || !attention_fused_supported(cch / n_heads, 0)) need += ((size_t)n_heads * t * (tk + 4) * 2) * 4;
  HostCall hc(h, need);
  float* xd = hc.upload(x, nx);
  float* cd = hc.upload(context, nc);
  float* wqd = hc.upload(wq, (size_t)cch * cch);
  float* wkd = hc.upload(wk, (size_t)cch * dc);
  float* wvd = hc.upload(wv, (size_t)cch * dc);
  float* wod = hc.upload(wo, (size_t)cch * cch);
  float* bqd = bq ? hc.upload(bq, cch) : nullptr;
  float* bkd = bk ? hc.upload(bk, cch) : nullptr;
  float* bvd = bv ? hc.upload(bv, cch) : nullptr;
  float* bod = bo ? hc.upload(bo, cch) : nullptr;
  float* q = hc.dev(nx);
  float* k = hc.dev(nkv);
  float* v = hc.dev(nkv);
  float* o = hc.dev(nx);
  float* y = hc.dev(nx);
  if (!hc.rc) {
    Ctx* c = hc.c;
    GemmArgs g;
    g.A = xd; g.M = t; g.K = cch; g.lda = cch; g.B = wqd; g.N = cch; g.ldb = cch; g.D = q; g.ldd = cch;
    g.bias = bqd;
    hc.run(op_gemm(c, g));
    GemmArgs gk;
    gk.A = cd; gk.M = tk; gk.K = dc; gk.lda = dc; gk.B = wkd; gk.N = cch; gk.ldb = dc; gk.D = k;
    gk.ldd = cch; gk.bias = bkd;
    if (!hc.rc) hc.run(op_gemm(c, gk));
    gk.B = wvd; gk.D = v; gk.bias = bvd;
    if (!hc.rc) hc.run(op_gemm(c, gk));
    AttnArgs a;
    a.Q = q; a.K = k; a.V = v; a.heads = n_heads; a.Tq = t; a.Tk = tk; a.d = cch / n_heads; a.O = o;
    a.softmax_axis = c->softmax_axis;
    if (!hc.rc) hc.run(op_attention(c, a));
    GemmArgs p;
    p.A = o; p.M = t; p.K = cch; p.lda = cch; p.B = wod; p.N = cch; p.ldb = cch; p.D = y; p.ldd = cch;
    p.bias = bod;
    if (!hc.rc) hc.run(op_gemm(c, p));
    hc.download(out, y, nx);
  }
  return hc.finish();
}

int32_t tsd_sampler_step(tsd_ctx* h, const float* latents, const float* eps_cond,
                         const float* eps_uncond, float cfg_scale, const float* noise, float sqrt_ab,
                         float sqrt_1mab, float c0, float c1, float sigma, int64_t n, float* out) {
  if (!h || !latents || !eps_cond || !out || n <= 0) return TSD_ERR_INVALID;
  HostCall hc(h, 5 * (size_t)n * 4 + (1u << 20));
  float* xd = hc.upload(latents, n);
  float* ec = hc.upload(eps_cond, n);
  float* eu = eps_uncond ? hc.upload(eps_uncond, n) : nullptr;
  float* nz = noise ? hc.upload(noise, n) : nullptr;
  float* od = hc.dev(n);
  if (!hc.rc) {
    hc.cu(launch_ddpm_step(xd, ec, eu, cfg_scale, nz, sqrt_ab, sqrt_1mab, c0, c1, sigma, od, n,
                           hc.c->stream),
          "ddpm_step");
    hc.download(out, od, n);
  }
  return hc.finish();
}

// DDPMSampler.add_noise, sampler.mojo:111-124: out = x * sqrt(ab_t) + noise * sqrt(1 - ab_t).
// Runs on the step kernel with (sqrt_ab, sqrt_1mab, c0, c1, sigma) = (1, 0, sqrt(ab_t), 0, sqrt(1-ab_t)).
int32_t tsd_sampler_add_noise(tsd_ctx* h, const float* x, const float* noise, float sqrt_ab,
                              float sqrt_1mab, int64_t n, float* out) {
  if (!h || !x || !noise || !out || n <= 0) return TSD_ERR_INVALID;
  HostCall hc(h, 3 * (size_t)n * 4 + (1u << 20));
  float* xd = hc.upload(x, n);
  float* nz = hc.upload(noise, n);
  float* od = hc.dev(n);
  if (!hc.rc) {
    hc.cu(launch_ddpm_step(xd, xd, nullptr, 0.0f, nz, 1.0f, 0.0f, sqrt_ab, 0.0f, sqrt_1mab, od, n, hc.c->stream),
          "add_noise");
    hc.download(out, od, n);
  }
  return hc.finish();
}

int32_t tsd_sampler_step_dev(tsd_ctx* h, const float* latents, const float* eps_cond,
                             const float* eps_uncond, float cfg_scale, const float* noise, float sqrt_ab,
                             float sqrt_1mab, float c0, float c1, float sigma, int64_t n, float* out) {
  if (!h || !latents || !eps_cond || !out || n <= 0) return TSD_ERR_INVALID;
  std::lock_guard<std::mutex> g(h->mu);
  Ctx* c = h->c;
  cudaSetDevice(c->device);
  int rc = c->check(launch_ddpm_step(latents, eps_cond, eps_uncond, cfg_scale, noise, sqrt_ab, sqrt_1mab,
                                     c0, c1, sigma, out, n, c->stream),
                    "ddpm_step");
  if (!rc) c->launches++;
  return rc;
}

// ---------------------------------------------------------------------------------------
// tuning probes: device-resident synthetic operands, CUDA-event timing (ms per launch)
// ---------------------------------------------------------------------------------------
static void fill_uniform(tsd::Ctx* c, float* p, size_t n, uint64_t seed, float scale) {
  launch_fill_uniform(p, (long long)n, seed, -scale, scale, c->stream);
}

int32_t tsd_bench_gemm(tsd_ctx* h, int32_t m, int32_t n, int32_t k, int32_t batch, int32_t geglu,
                       int32_t force_bn, int32_t force_splits, int32_t iters, double* ms_out) {
  if (!h || !ms_out || m <= 0 || n <= 0 || k <= 0 || batch <= 0 || iters <= 0) return TSD_ERR_INVALID;
  const size_t na = (size_t)batch * m * k, nb = (size_t)batch * n * k, nd = (size_t)batch * m * n;
  HostCall hc(h, (na + nb + nd + n) * 4 + (size_t)32 * m * n * 4 + (64u << 20));
  float* A = hc.dev(na);
  float* B = hc.dev(nb);
  float* D = hc.dev(nd);
  float* bias = hc.dev(n);
  if (hc.rc) return hc.finish();
  Ctx* c = hc.c;
  fill_uniform(c, A, na, 1, 1.0f);
  fill_uniform(c, B, nb, 2, 0.05f);
  fill_uniform(c, bias, n, 3, 0.05f);
  GemmArgs g;
  g.A = A; g.M = m; g.K = k; g.lda = k; g.a_bs = (long long)m * k;
  g.B = B; g.N = n; g.ldb = k; g.b_bs = (long long)n * k; g.batch = batch;
  g.D = D; g.ldd = geglu ? n / 2 : n; g.d_bs = (long long)m * g.ldd; g.bias = bias; g.geglu = geglu;
  g.force_bn = force_bn; g.force_splits = force_splits;
  NormHint bench_nh;
  if (c->bench_stats_groups > 0) {
    bench_nh.G = c->bench_stats_groups; bench_nh.eps = 1e-5f; bench_nh.imgs = 1;
    bench_nh.scratch_elems = norm_scratch_elems(1, m, n, bench_nh.G);
    bench_nh.scratch = reinterpret_cast<float2*>(hc.dev(2 * bench_nh.scratch_elems));
    g.nh = &bench_nh;
  }
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int i = 0; i < 3 && !hc.rc; ++i) hc.run(op_gemm(c, g));
  launch_spin((long long)iters * 12000, c->stream);  // queue the launches behind a spin: events see device time
  cudaEventRecord(e0, c->stream);
  for (int i = 0; i < iters && !hc.rc; ++i) hc.run(op_gemm(c, g));
  cudaEventRecord(e1, c->stream);
  int rc = hc.finish();
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *ms_out = ms / iters;
  return rc;
}

int32_t tsd_bench_attention(tsd_ctx* h, int32_t heads, int32_t tq, int32_t tk, int32_t d, int32_t iters,
                            double* ms_out) {
  if (!h || !ms_out || heads <= 0 || tq <= 0 || tk <= 0 || d <= 0 || d % 4 || iters <= 0) return TSD_ERR_INVALID;
  const size_t nq = (size_t)heads * tq * d, nk = (size_t)heads * tk * d;
  size_t need = (2 * nq + 2 * nk) * 4 + (size_t)heads * ((size_t)tq + tk) * 64 + (32u << 20);
  if (!h->c->fused_attention || !attention_fused_supported(d, 0)) need += ((size_t)heads * tq * (tk + 4) + (size_t)heads * d * (tk + 4)) * 4;
  HostCall hc(h, need);
  float* q = hc.dev(nq);
  float* k = hc.dev(nk);
  float* v = hc.dev(nk);
  float* o = hc.dev(nq);
  if (hc.rc) return hc.finish();
  Ctx* c = hc.c;
  fill_uniform(c, q, nq, 1, 1.7f);  // variance ~1
  fill_uniform(c, k, nk, 2, 1.7f);
  fill_uniform(c, v, nk, 3, 1.7f);
  AttnArgs a;
  a.Q = q; a.K = k; a.V = v; a.heads = heads; a.Tq = tq; a.Tk = tk; a.d = d; a.O = o;
  a.softmax_axis = c->softmax_axis;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int i = 0; i < 3 && !hc.rc; ++i) hc.run(op_attention(c, a));
  launch_spin((long long)iters * 25000, c->stream);  // queue the launches behind a spin: events see device time
  cudaEventRecord(e0, c->stream);
  for (int i = 0; i < iters && !hc.rc; ++i) hc.run(op_attention(c, a));
  cudaEventRecord(e1, c->stream);
  int rc = hc.finish();
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *ms_out = ms / iters;
  return rc;
}

// Lab: GroupNorm(+SiLU) consumer timed alone.  mode 0: stand-alone (statistics + normalise in one fused launch);
// mode 1: normalise-only pass over producer-side partial statistics left by a 3x3 convolution (run once, untimed);
// mode 2: the producer convolution + its consumer norm, timed as a pair.
int32_t tsd_bench_norm(tsd_ctx* h, int32_t n, int32_t H, int32_t W, int32_t C, int32_t G, int32_t mode, int32_t iters,
                       double* ms_out) {
  if (!h || !ms_out || n <= 0 || H <= 0 || W <= 0 || C <= 0 || G <= 0 || iters <= 0) return TSD_ERR_INVALID;
  const size_t nx = (size_t)n * H * W * C, nw = (size_t)C * 9 * C;
  HostCall hc(h, (3 * nx + nw + C) * 4 + (size_t)34 * nx * 4 + (256u << 20));
  float* x = hc.dev(nx);
  float* w = hc.dev(nw);
  float* y = hc.dev(nx);
  float* o = hc.dev(nx);
  float* bias = hc.dev(C);
  if (hc.rc) return hc.finish();
  Ctx* c = hc.c;
  fill_uniform(c, x, nx, 1, 1.0f);
  fill_uniform(c, w, nw, 2, 0.02f);
  fill_uniform(c, bias, C, 3, 0.05f);
  NormHint nh;
  nh.G = G; nh.eps = 1e-5f; nh.imgs = n;
  nh.scratch_elems = norm_scratch_elems(n, (long long)H * W, C, G);
  nh.scratch = reinterpret_cast<float2*>(hc.dev(2 * nh.scratch_elems));
  ConvArgs a;
  a.x = x; a.N = n; a.H = H; a.W = W; a.Cin = C; a.Cout = C; a.k = 3; a.pad = 1; a.stride = 1;
  a.w = w; a.bias = bias; a.out = y; a.nh = mode >= 1 ? &nh : nullptr;
  auto consumer = [&]() { return op_group_norm(c, y, o, n, H, W, C, G, 1e-5f, nullptr, nullptr, 1.0f, 1, 0, 1, mode >= 1 ? nh.ready() : nullptr, nullptr); };
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int i = 0; i < 3 && !hc.rc; ++i) {
    hc.run(op_conv2d(c, a));
    hc.run(consumer());
  }
  launch_spin((long long)iters * 20000, c->stream);
  cudaEventRecord(e0, c->stream);
  for (int i = 0; i < iters && !hc.rc; ++i) {
    if (mode == 2) hc.run(op_conv2d(c, a));
    hc.run(consumer());
  }
  cudaEventRecord(e1, c->stream);
  int rc = hc.finish();
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *ms_out = ms / iters;
  return rc;
}

int32_t tsd_bench_conv(tsd_ctx* h, int32_t n, int32_t H, int32_t W, int32_t cin, int32_t cout,
                       int32_t k, int32_t stride, int32_t force_bn, int32_t force_splits,
                       int32_t iters, double* ms_out) {
  if (!h || !ms_out || n <= 0 || H <= 0 || W <= 0 || cin <= 0 || cout <= 0 || iters <= 0) return TSD_ERR_INVALID;
  const int pad = k / 2;
  const int Ho = conv_out_dim(H, k, pad, stride), Wo = conv_out_dim(W, k, pad, stride);
  const size_t nx = (size_t)n * H * W * cin, nw = (size_t)cout * k * k * cin, no = (size_t)n * Ho * Wo * cout;
  HostCall hc(h, (nx + nw + no + cout) * 4 + (size_t)34 * no * 4 + (size_t)n * Ho * Wo * 9 * cin * 4 + (64u << 20));
  float* x = hc.dev(nx);
  float* w = hc.dev(nw);
  float* o = hc.dev(no);
  float* bias = hc.dev(cout);
  if (hc.rc) return hc.finish();
  Ctx* c = hc.c;
  fill_uniform(c, x, nx, 1, 1.0f);
  fill_uniform(c, w, nw, 2, 0.02f);
  fill_uniform(c, bias, cout, 3, 0.05f);
  ConvArgs a;
  a.x = x; a.N = n; a.H = H; a.W = W; a.Cin = cin; a.Cout = cout; a.k = k; a.pad = pad; a.stride = stride;
  a.w = w; a.bias = bias; a.out = o; a.force_bn = force_bn; a.force_splits = force_splits;
  NormHint bench_nh;
  if (c->bench_stats_groups > 0) {
    bench_nh.G = c->bench_stats_groups; bench_nh.eps = 1e-5f;
    bench_nh.scratch_elems = norm_scratch_elems(n, (long long)Ho * Wo, cout, bench_nh.G);
    bench_nh.scratch = reinterpret_cast<float2*>(hc.dev(2 * bench_nh.scratch_elems));
    a.nh = &bench_nh;
  }
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int i = 0; i < 3 && !hc.rc; ++i) hc.run(op_conv2d(c, a));
  launch_spin((long long)iters * 12000, c->stream);
  cudaEventRecord(e0, c->stream);
  for (int i = 0; i < iters && !hc.rc; ++i) hc.run(op_conv2d(c, a));
  cudaEventRecord(e1, c->stream);
  int rc = hc.finish();
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *ms_out = ms / iters;
  return rc;
}

}  // extern "C"
