// Model graphs (see models.h).  Reference structure being followed:
//   Time_Embedding.forward        diffusion.mojo:17-21
//   Unet_Residual_Block.forward   diffusion.mojo:54-72
//   Unet_Attention_Block.forward  diffusion.mojo:112-147
//   UNet.forward                  diffusion.mojo:228-273   (Q8 upsample, Q9 dead skips, Q10 [x;x])
//   UNet_Output_Layer.forward     diffusion.mojo:287-291
//   Decoder.forward               vae.mojo:221-250, Res_Block :57-67, Attention_Block :17-27
#include "models.h"

#include <cmath>
#include <cstdio>
#include <cstring>

#include "c_api_internal.h"
#include "elementwise.cuh"

namespace tsd {

#define TRY(expr)            \
  do {                       \
    int rc__ = (expr);       \
    if (rc__) return rc__;   \
  } while (0)

// a raw launch in a model graph: skipped in the planning pass, counted otherwise
#define LAUNCH(c, expr, what)                          \
  do {                                                 \
    if (!(c)->dry_run) {                               \
      int rc__ = (c)->check((expr), what);             \
      if (rc__) return rc__;                           \
      (c)->launches++;                                 \
    }                                                  \
  } while (0)

namespace {

// ------------------------------------------------------------------------------------------
// synthetic parameters: counter-based splitmix64, defined on the REFERENCE-layout element index
// so that oracle/synth.py regenerates bit-identical tensors on the CPU.
//   value(j) = float(int(z >> 40) - 2^23) * (scale * 2^-23),  z = splitmix64(param_seed + j)
// ------------------------------------------------------------------------------------------
__host__ __device__ inline uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  uint64_t z = x;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

__device__ __forceinline__ float rna_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;\n" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// dst is in device layout: conv [O][KK][I], everything else as the reference
__global__ void synth_param_kernel(float* __restrict__ dst, long long n, int kind, int I, int KK,
                                   uint64_t pseed, float step, int round) {
  for (long long d = (long long)blockIdx.x * blockDim.x + threadIdx.x; d < n;
       d += (long long)gridDim.x * blockDim.x) {
    long long j = d;
    if (kind == P_CONV_W && KK > 1) {
      int ci = (int)(d % I);
      long long t = d / I;
      int tap = (int)(t % KK);
      long long o = t / KK;
      j = (o * I + ci) * KK + tap;
    }
    uint64_t z = splitmix64(pseed + (uint64_t)j);
    float v = (float)((int)(z >> 40) - 8388608) * step;
    dst[d] = round ? rna_tf32(v) : v;
  }
}

__global__ void round_tf32_kernel(float* __restrict__ p, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    p[i] = rna_tf32(p[i]);
}

__global__ void ohwi_to_oihw_kernel(const float* __restrict__ src, float* __restrict__ dst, int O,
                                    int I, int KK) {
  const long long total = (long long)O * I * KK;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int tap = (int)(i % KK);
    long long t = i / KK;
    int ci = (int)(t % I);
    long long o = t / I;
    dst[i] = src[(o * KK + tap) * I + ci];
  }
}

inline int blocks_for(long long n) {
  long long b = (n + 255) / 256;
  if (b > 148 * 8) b = 148 * 8;
  if (b < 1) b = 1;
  return (int)b;
}

// ---- denoising-loop glue (one CUDA graph serves every step: the step index lives on the device)
struct StepGather {
  const float* src[9];  // tb_all[r] : [steps][cout]
  float* dst[9];        // tbias[r]  : [cout]
  int cout[9];
};
__global__ void step_prologue_kernel(const int* __restrict__ step, StepGather g,
                                     const float* __restrict__ latents, float* __restrict__ x,
                                     long long n_lat, int copies) {
  const int s = *step;
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (int r = 0; r < 9; ++r)
    for (long long i = tid; i < g.cout[r]; i += stride) g.dst[r][i] = g.src[r][(long long)s * g.cout[r] + i];
  for (long long i = tid; i < n_lat * copies; i += stride) x[i] = latents[i % n_lat];
}
// sampler.mojo:75-109 (+ CFG combine pipeline.mojo:117-119); coef[s] = sqrt_ab, sqrt_1mab, c0, c1, sigma
__global__ void step_epilogue_kernel(const int* __restrict__ step, const float* __restrict__ coef,
                                     const float* __restrict__ eps, int cfg, float cfg_scale,
                                     const float* __restrict__ noise, float* __restrict__ latents,
                                     long long n_lat) {
  const int s = *step;
  const float sqrt_ab = coef[5 * s], sqrt_1mab = coef[5 * s + 1], c0 = coef[5 * s + 2],
              c1 = coef[5 * s + 3], sigma = coef[5 * s + 4];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_lat;
       i += (long long)gridDim.x * blockDim.x) {
    float e = eps[i];
    if (cfg) {
      float u = eps[n_lat + i];
      e = (e - u) * cfg_scale + u;
    }
    float xi = latents[i];
    float x0 = (xi - e * sqrt_1mab) / sqrt_ab;
    float o = x0 * c0 + xi * c1;
    if (noise != nullptr && sigma != 0.0f) o += noise[(long long)s * n_lat + i] * sigma;
    latents[i] = o;
  }
}
__global__ void step_advance_kernel(int* step) { *step += 1; }

}  // namespace

// ------------------------------------------------------------------------------------------
// ParamStore
// ------------------------------------------------------------------------------------------
int ParamStore::add(const std::string& name, int kind, int O, int I, int KK, float init_scale) {
  Param p;
  p.name = name;
  p.kind = kind;
  p.O = O;
  p.I = I;
  p.KK = KK;
  p.numel = (long long)O * (kind == P_VEC ? 1 : I) * (kind == P_CONV_W ? KK : 1);
  p.offset = total;
  p.init_scale = init_scale;
  total += p.numel;
  params.push_back(p);
  return (int)params.size() - 1;
}
int ParamStore::add_conv(const std::string& name, int cin, int cout, int k) {
  // Conv2D init: U(+-1/sqrt(fan_in)) weights, zero bias (helpers/utils.mojo:1717, 1722-1724)
  const float s = 1.0f / sqrtf((float)(cin * k * k));
  int i = add(name + ".weight", P_CONV_W, cout, cin, k * k, s);
  add(name + ".bias", P_VEC, cout, 1, 1, 0.0f);
  return i;
}
int ParamStore::add_linear(const std::string& name, int in_f, int out_f, bool bias) {
  // synthetic Linear init U(+-1/sqrt(in)) for weight and bias (the reference's
  // in^-1/4 overflows its own un-stabilised exp)
  const float s = 1.0f / sqrtf((float)in_f);
  int i = add(name + ".weight", P_LIN_W, out_f, in_f, 1, s);
  if (bias) add(name + ".bias", P_VEC, out_f, 1, 1, s);
  return i;
}
int ParamStore::add_norm(const std::string& name, int channels) {
  int i = add(name + ".weight", P_VEC, channels, 1, 1, 0.0f);
  params[i].init_const = 1.0f;
  add(name + ".bias", P_VEC, channels, 1, 1, 0.0f);
  return i;
}
int ParamStore::allocate() {
  if (block) return TSD_OK;
  // every tensor starts 256 B aligned (TMA base alignment is 16 B)
  long long off = 0;
  std::vector<long long> offs;
  for (auto& p : params) {
    offs.push_back(off);
    off += (p.numel + 63) / 64 * 64;
  }
  if (cudaMalloc(&block, (size_t)off * sizeof(float)) != cudaSuccess) {
    cudaGetLastError();
    return c->fail(TSD_ERR_OOM, "parameter allocation failed");
  }
  for (size_t i = 0; i < params.size(); ++i) params[i].dev = block + offs[i];
  return TSD_OK;
}
void ParamStore::free_all() {
  for (float* r : rowsum_dev)
    if (r) cudaFree(r);
  rowsum_dev.clear();
  rowsum_gen.clear();
  for (float* r : ffn_dev)
    if (r) cudaFree(r);
  ffn_dev.clear();
  ffn_gen.clear();
  for (float* r : fskip_dev)
    if (r) cudaFree(r);
  fskip_dev.clear();
  fskip_gen.clear();
  if (block) cudaFree(block);
  block = nullptr;
}
int ParamStore::load(const float* blob, long long n_floats) {
  if (n_floats != total) {
    char buf[128];
    snprintf(buf, sizeof buf, "load_weights: expected %lld floats, got %lld", total, n_floats);
    return c->fail(TSD_ERR_INVALID, buf);
  }
  TRY(allocate());
  // stage through a bounded device buffer, tensor by tensor
  long long max_numel = 0;
  for (auto& p : params) max_numel = p.numel > max_numel ? p.numel : max_numel;
  float* stage = nullptr;
  if (cudaMalloc(&stage, (size_t)max_numel * sizeof(float)) != cudaSuccess) {
    cudaGetLastError();
    return c->fail(TSD_ERR_OOM, "load_weights: staging allocation failed");
  }
  int rc = TSD_OK;
  for (auto& p : params) {
    if (p.kind == P_CONV_W && p.KK > 1) {
      rc = c->check(cudaMemcpyAsync(stage, blob + p.offset, p.numel * sizeof(float), cudaMemcpyHostToDevice,
                                    c->stream), "load_weights H2D");
      if (rc) break;
      rc = c->check(launch_oihw_to_ohwi(stage, p.dev, p.O, p.I, p.KK, c->stream), "oihw_to_ohwi");
      if (rc) break;
      // the staging buffer is reused by the next tensor: stream order makes that safe
    } else {
      rc = c->check(cudaMemcpyAsync(p.dev, blob + p.offset, p.numel * sizeof(float), cudaMemcpyHostToDevice,
                                    c->stream), "load_weights H2D");
      if (rc) break;
    }
    if (p.kind != P_VEC) {
      round_tf32_kernel<<<blocks_for(p.numel), 256, 0, c->stream>>>(p.dev, p.numel);
      rc = c->check(cudaGetLastError(), "round_tf32");
      if (rc) break;
    }
  }
  int rc2 = c->check(cudaStreamSynchronize(c->stream), "load_weights sync");
  cudaFree(stage);
  if (rc) return rc;
  if (rc2) return rc2;
  loaded = true;
  ++gen;
  return TSD_OK;
}
int ParamStore::init_random(uint64_t seed) {
  TRY(allocate());
  for (size_t i = 0; i < params.size(); ++i) {
    Param& p = params[i];
    if (p.init_scale == 0.0f) {
      if (p.init_const != 0.0f) TRY(c->check(launch_fill_uniform(p.dev, p.numel, 1, p.init_const, p.init_const, c->stream), "fill"));
      else TRY(c->check(cudaMemsetAsync(p.dev, 0, p.numel * sizeof(float), c->stream), "memset"));
      continue;
    }
    const uint64_t pseed = splitmix64(seed ^ splitmix64((uint64_t)i + 1));
    const float step = p.init_scale * (1.0f / 8388608.0f);
    synth_param_kernel<<<blocks_for(p.numel), 256, 0, c->stream>>>(p.dev, p.numel, p.kind, p.I, p.KK, pseed,
                                                                 step, p.kind != P_VEC);
    TRY(c->check(cudaGetLastError(), "synth_param"));
  }
  TRY(c->check(cudaStreamSynchronize(c->stream), "init_random sync"));
  loaded = true;
  ++gen;
  return TSD_OK;
}
int ParamStore::get(int i, float* host_out) {
  if (i < 0 || i >= (int)params.size()) return c->fail(TSD_ERR_INVALID, "get_param: index out of range");
  if (!loaded) return c->fail(TSD_ERR_STATE, "get_param: no weights loaded");
  Param& p = params[i];
  const float* src = p.dev;
  float* tmp = nullptr;
  if (p.kind == P_CONV_W && p.KK > 1) {
    if (cudaMalloc(&tmp, p.numel * sizeof(float)) != cudaSuccess) {
      cudaGetLastError();
      return c->fail(TSD_ERR_OOM, "get_param: allocation failed");
    }
    ohwi_to_oihw_kernel<<<blocks_for(p.numel), 256, 0, c->stream>>>(p.dev, tmp, p.O, p.I, p.KK);
    src = tmp;
  }
  int rc = c->check(cudaMemcpyAsync(host_out, src, p.numel * sizeof(float), cudaMemcpyDeviceToHost, c->stream),
                    "get_param D2H");
  int rc2 = c->check(cudaStreamSynchronize(c->stream), "get_param sync");
  if (tmp) cudaFree(tmp);
  return rc ? rc : rc2;
}

__global__ void weight_rowsum_kernel(const float* __restrict__ w, float* __restrict__ out, int O, int I) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= O) return;
  float s = 0.f;
  for (int k = lane; k < I; k += 32) s += w[(long long)row * I + k];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) out[row] = s;
}

__global__ void fuse_skip_weights_kernel(const float* __restrict__ w1, const float* __restrict__ b1,
                                         const float* __restrict__ w2, const float* __restrict__ b2,
                                         float* __restrict__ out, int O, int K1, int K2) {
  const long long K = (long long)K1 + K2, total = (long long)O * K;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total + O;
       i += (long long)gridDim.x * blockDim.x) {
    if (i < total) {
      const long long o = i / K, k = i - o * K;
      out[i] = k < K1 ? w1[o * K1 + k] : w2[o * K2 + (k - K1)];
    } else {
      const long long o = i - total;
      out[i] = b1[o] + b2[o];
    }
  }
}

// out[o][k] = sum_j wc[o][j] * w2[j][k] for k < K2 (the product of the two weight matrices, fp32), wc[o][k - K2] behind it;
// out[O * (K2 + C) + o] = sum_j wc[o][j] * b2[j] + bc[o].  One thread per output element, 16 x 16 tiles through shared memory.
__global__ void fuse_ffn_weights_kernel(const float* __restrict__ wc, const float* __restrict__ bc,
                                        const float* __restrict__ w2, const float* __restrict__ b2,
                                        float* __restrict__ out, int C, int K2) {
  __shared__ float ta[16][17], tb[16][17];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int o = blockIdx.y * 16 + ty, k = blockIdx.x * 16 + tx;
  const int K = K2 + C;
  // column K2 + C (one extra tile column) carries the bias product
  float acc = 0.f;
  for (int j0 = 0; j0 < C; j0 += 16) {
    ta[ty][tx] = (o < C && j0 + tx < C) ? wc[(long long)o * C + j0 + tx] : 0.f;
    float bv = 0.f;
    if (j0 + ty < C) {
      if (k < K2) bv = w2[(long long)(j0 + ty) * K2 + k];
      else if (k == K) bv = b2[j0 + ty];
    }
    tb[ty][tx] = bv;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 16; ++j) acc = fmaf(ta[ty][j], tb[j][tx], acc);
    __syncthreads();
  }
  if (o >= C) return;
  if (k < K2) out[(long long)o * K + k] = rna_tf32(acc);
  else if (k < K) out[(long long)o * K + k] = wc[(long long)o * C + (k - K2)];  // already TF32-rounded at load
  else if (k == K) out[(long long)C * K + o] = acc + bc[o];
}

int ParamStore::fused_ffn_out(int w2, int wc, const float** w, const float** b) {
  if (ffn_dev.size() != params.size()) {
    ffn_dev.assign(params.size(), nullptr);
    ffn_gen.assign(params.size(), -1);
  }
  const Param& p2 = params[w2];
  const Param& pc = params[wc];
  const int C = pc.O, K2 = p2.I * p2.KK;
  if (p2.O != C || pc.I * pc.KK != C) return c->fail(TSD_ERR_INVALID, "fused ffn out: shapes do not chain");
  const size_t elems = (size_t)C * (K2 + C);
  if (!ffn_dev[wc]) {
    if (cudaMalloc(&ffn_dev[wc], sizeof(float) * (elems + C)) != cudaSuccess) {
      cudaGetLastError();
      return c->fail(TSD_ERR_OOM, "fused ffn out weights: allocation failed");
    }
  }
  if (ffn_gen[wc] != gen && !c->dry_run) {
    const dim3 grid((K2 + C) / 16 + 1, (C + 15) / 16), block(16, 16);
    fuse_ffn_weights_kernel<<<grid, block, 0, c->stream>>>(pc.dev, params[wc + 1].dev, p2.dev, params[w2 + 1].dev, ffn_dev[wc], C, K2);
    TRY(c->check(cudaGetLastError(), "fuse_ffn_weights launch"));
    ffn_gen[wc] = gen;
  }
  *w = ffn_dev[wc];
  *b = ffn_dev[wc] + elems;
  return TSD_OK;
}

int ParamStore::fused_skip(int conv_w, int skip_w, const float** w, const float** b) {
  if (fskip_dev.size() != params.size()) {
    fskip_dev.assign(params.size(), nullptr);
    fskip_gen.assign(params.size(), -1);
  }
  const Param& p1 = params[conv_w];
  const Param& p2 = params[skip_w];
  const int K1 = p1.I * p1.KK, K2 = p2.I * p2.KK;
  const size_t elems = (size_t)p1.O * (K1 + K2);
  if (p2.O != p1.O) return c->fail(TSD_ERR_INVALID, "fused skip: output channels differ");
  if (!fskip_dev[conv_w]) {
    if (cudaMalloc(&fskip_dev[conv_w], sizeof(float) * (elems + p1.O)) != cudaSuccess) {
      cudaGetLastError();
      return c->fail(TSD_ERR_OOM, "fused skip weights: allocation failed");
    }
  }
  if (fskip_gen[conv_w] != gen && !c->dry_run) {
    fuse_skip_weights_kernel<<<1184, 256, 0, c->stream>>>(p1.dev, params[conv_w + 1].dev, p2.dev, params[skip_w + 1].dev,
                                                        fskip_dev[conv_w], p1.O, K1, K2);
    TRY(c->check(cudaGetLastError(), "fuse_skip_weights launch"));
    fskip_gen[conv_w] = gen;
  }
  *w = fskip_dev[conv_w];
  *b = fskip_dev[conv_w] + elems;
  return TSD_OK;
}

const float* ParamStore::rowsum(int i) {
  if (rowsum_dev.size() != params.size()) {
    rowsum_dev.assign(params.size(), nullptr);
    rowsum_gen.assign(params.size(), -1);
  }
  const Param& p = params[i];
  if (!rowsum_dev[i]) {
    if (cudaMalloc(&rowsum_dev[i], sizeof(float) * (size_t)p.O) != cudaSuccess) {
      c->fail(TSD_ERR_OOM, "rowsum allocation failed");
      return nullptr;
    }
  }
  if (rowsum_gen[i] != gen && !c->dry_run) {
    weight_rowsum_kernel<<<(p.O + 7) / 8, 256, 0, c->stream>>>(p.dev, rowsum_dev[i], p.O, p.I * p.KK);
    if (c->check(cudaGetLastError(), "weight_rowsum launch")) return nullptr;
    rowsum_gen[i] = gen;
  }
  return rowsum_dev[i];
}

// ------------------------------------------------------------------------------------------
// blocks
// ------------------------------------------------------------------------------------------
static float* walloc(Ctx* c, long long n) { return c->arena.alloc_n<float>((size_t)(n > 0 ? n : 1)); }
#define WALLOC(var, n)                                                          \
  float* var = walloc(c, (n));                                                  \
  if (!var) return c->fail(TSD_ERR_OOM, "workspace exhausted (" #var ")")

static int conv(Ctx* c, const ParamStore& ps, int wi, const float* x, int N, int H, int W, int cin, int cout,
                int k, int pad, int stride, const float* bias_override, int bias_img_stride,
                const float* residual, float* out, int round_out, NormHint* nh = nullptr, int pad_hi = -1,
                const float* x2 = nullptr, int cin2 = 0, const float* w_override = nullptr) {
  ConvArgs a;
  a.nh = nh;
  a.pad_hi = pad_hi;
  a.x2 = x2;
  a.Cin2 = cin2;
  a.x = x; a.N = N; a.H = H; a.W = W; a.Cin = cin; a.Cout = cout; a.k = k; a.pad = pad; a.stride = stride;
  a.w = w_override ? w_override : ps.w(wi);
  a.bias = bias_override ? bias_override : ps.w(wi + 1);
  a.bias_img_stride = bias_override ? bias_img_stride : 0;
  a.residual = residual;
  a.out = out;
  a.round_tf32 = round_out;
  a.force_bn = c->force_bn; a.force_splits = c->force_splits;
  return op_conv2d(c, a);
}

static int linear(Ctx* c, const float* x, long long M, int K, const float* w, const float* bias, int N,
                  float* out, long long ldd, const float* residual, int round_out, int geglu = 0,
                  int split_n = 0, long long split_stride = 0, NormHint* nh = nullptr,
                  const NormStatsReq* ln_fold = nullptr, const float* wsum = nullptr) {
  GemmArgs g;
  g.nh = nh;
  g.ln_fold = ln_fold;
  g.wsum = wsum;
  g.b_static = 1;  // every Linear of the models multiplies by a parameter matrix
  g.A = x; g.M = (int)M; g.K = K; g.lda = K;
  g.B = w; g.N = N; g.ldb = K;
  g.D = out; g.ldd = ldd;
  g.bias = bias;
  g.residual = residual; g.ldr = ldd;
  g.geglu = geglu;
  g.split_n = split_n; g.split_stride = split_stride;
  g.round_tf32 = round_out;
  g.force_bn = c->force_bn; g.force_splits = c->force_splits;
  return op_gemm(c, g);
}

// Unet_Residual_Block.forward (diffusion.mojo:54-72) / VAE Res_Block.forward (vae.mojo:57-67).
// tbias (optional) = Linear(SiLU(t)) + its bias + conv1 bias, one row per image.
int res_block(Ctx* c, const ParamStore& ps, const ResBlockW& w, const Act& x, const float* tbias,
              int tbias_stride, float eps, float* out, NormHint* next) {
  const int N = x.N, H = x.H, W = x.W;
  const long long px = x.pixels();
  const size_t mark = c->arena.mark();
  WALLOC(h1, px * w.cin);
  const NormStatsReq* xs = (x.ns.G == w.groups && x.ns.eps == eps && x.C == w.cin) ? x.ns.ready() : nullptr;
  const bool xmatch = x.ns.G == w.groups && x.ns.eps == eps && x.C == w.cin;
  TRY(op_group_norm(c, x.p, h1, N, H, W, w.cin, w.groups, eps, ps.gamma(w.gn1), ps.beta(w.gn1), 1.0f, 1, 0, 1, xs,
                    xmatch ? x.ns.deferred() : nullptr));
  WALLOC(h2, px * w.cout);
  NormHint mid;
  mid.G = w.groups;
  mid.eps = eps;
  mid.imgs = N;
  mid.allow_defer = true;  // h2 has one reader, the GroupNorm right below
  mid.scratch_elems = norm_scratch_elems(N, (long long)H * W, w.cout, w.groups);
  mid.scratch = c->arena.alloc_n<float2>(mid.scratch_elems);
  if (!mid.scratch) return c->fail(TSD_ERR_OOM, "workspace exhausted (norm statistics)");
  TRY(conv(c, ps, w.conv1, h1, N, H, W, w.cin, w.cout, 3, 1, 1, tbias, tbias_stride, nullptr, h2, 0, &mid));
  WALLOC(h3, px * w.cout);
  TRY(op_group_norm(c, h2, h3, N, H, W, w.cout, w.groups, eps, ps.gamma(w.gn2), ps.beta(w.gn2), 1.0f, 1, 0, 1, mid.ready(), mid.deferred()));
  const float* r = x.p;
  if (w.cin != w.cout && c->fuse_skip && w.cin % 64 == 0 && w.cout % 64 == 0) {
    // res_conv_layer (1x1, diffusion.mojo:66-70 / vae.mojo:64-66) rides in the K loop of conv2: one GEMM over
    // K = 9*cout + cin with the concatenated weights and the summed biases, no residual operand
    const float *wcat = nullptr, *bcat = nullptr;
    TRY(const_cast<ParamStore&>(ps).fused_skip(w.conv2, w.skip, &wcat, &bcat));
    TRY(conv(c, ps, w.conv2, h3, N, H, W, w.cout, w.cout, 3, 1, 1, bcat, 0, nullptr, out, 0, next, -1, x.p, w.cin, wcat));
    c->arena.release_to(mark);
    return TSD_OK;
  }
  if (w.cin != w.cout) {
    WALLOC(rr, px * w.cout);
    TRY(conv(c, ps, w.skip, x.p, N, H, W, w.cin, w.cout, 1, 0, 1, nullptr, 0, nullptr, rr, 0));
    r = rr;
  }
  TRY(conv(c, ps, w.conv2, h3, N, H, W, w.cout, w.cout, 3, 1, 1, nullptr, 0, r, out, 0, next));
  c->arena.release_to(mark);
  return TSD_OK;
}

// LayerNorm.forward = GroupNorm(1, C) over the whole (C,T) tensor of one image (Q5), or per token
static int layer_norm(Ctx* c, const float* x, float* y, int N, long long T, int C, const NormStatsReq* pre = nullptr,
                      const float* gamma = nullptr, const float* beta = nullptr) {
  if (c->layernorm_mode == 0)
    return op_group_norm(c, x, y, N, (int)T, 1, C, 1, 1e-5f, gamma, beta, 1.0f, 0, 0, 1, pre);
  return op_group_norm(c, x, y, (int)(N * T), 1, 1, C, 1, 1e-5f, gamma, beta, 1.0f, 0, 0, 1);
}

// Unet_Attention_Block.forward, diffusion.mojo:112-147
static int attn_block(Ctx* c, const ParamStore& ps, const AttnBlockW& w, const Act& x, const float* kctx,
                      const float* vctx, int n_ctx, int ctx_len, float* out, NormHint* next = nullptr) {
  const int N = x.N, C = w.C, d = C / w.heads;
  const long long T = (long long)x.H * x.W, M = N * T;
  const size_t mark = c->arena.mark();
  WALLOC(a, M * C);
  const NormStatsReq* xs = (x.ns.G == 32 && x.ns.eps == 1e-6f) ? x.ns.ready() : nullptr;
  TRY(op_group_norm(c, x.p, a, N, x.H, x.W, C, 32, 1e-6f, ps.gamma(w.gn), ps.beta(w.gn), 1.0f, 0, 0, 1, xs,
                    (x.ns.G == 32 && x.ns.eps == 1e-6f) ? x.ns.deferred() : nullptr));
  // LayerNorm with global statistics (Q5) = one group per image: the producing GEMMs fold the sums
  NormHint ln;
  ln.G = c->layernorm_mode == 0 ? 1 : 0;
  ln.eps = 1e-5f;
  ln.imgs = N;
  ln.scratch_elems = norm_scratch_elems(N, T, C, 1);
  ln.scratch = c->arena.alloc_n<float2>(ln.scratch_elems);
  if (!ln.scratch) return c->fail(TSD_ERR_OOM, "workspace exhausted (norm statistics)");
  WALLOC(u, M * C);
  TRY(linear(c, a, M, C, ps.w(w.conv_in), ps.w(w.conv_in + 1), C, u, C, nullptr, 0, 0, 0, 0, &ln));
  WALLOC(v, M * C);
  // LayerNorm folded into the consuming GEMM's epilogue when its statistics came with the producer
  // (a LayerNorm with per-channel weight / bias - norm_affine models - is not a scalar scale and shift: it runs as
  // its own pass)
  auto ln_then_linear = [&](const float* src, int lni, int wi, const float* bias, int Nout, float* dst, long long ldd,
                            int geglu, int split_n, long long split_stride) -> int {
    if (c->ln_fold && ln.ready() && T % 128 == 0 && lni < 0) {
      const float* ws = const_cast<ParamStore&>(ps).rowsum(wi);
      if (!ws) return c->fail(TSD_ERR_OOM, "rowsum");
      const NormStatsReq req = ln.req;  // the producer of `src` filled it; the next producer will overwrite ln
      return linear(c, src, M, C, ps.w(wi), bias, Nout, dst, ldd, nullptr, 1, geglu, split_n, split_stride, nullptr, &req, ws);
    }
    TRY(layer_norm(c, src, v, N, T, C, ln.ready(), ps.gamma(lni), ps.beta(lni)));
    return linear(c, v, M, C, ps.w(wi), bias, Nout, dst, ldd, nullptr, 1, geglu, split_n, split_stride);
  };
  WALLOC(qkv, 3 * M * C);
  TRY(ln_then_linear(u, w.ln1, w.in_proj, nullptr, 3 * C, qkv, C, 0, C, M * C));
  WALLOC(o, M * C);
  {
    AttnArgs at;
    at.Q = qkv; at.K = qkv + M * C; at.V = qkv + 2 * M * C;
    at.batch = N; at.heads = w.heads; at.Tq = (int)T; at.Tk = (int)T; at.d = d; at.O = o;
    at.softmax_axis = c->softmax_axis;
    TRY(op_attention(c, at));
  }
  WALLOC(u2, M * C);
  TRY(linear(c, o, M, C, ps.w(w.out_proj), ps.w(w.out_proj + 1), C, u2, C, u, 0, 0, 0, 0, &ln));
  float* q = qkv;
  TRY(ln_then_linear(u2, w.ln2, w.q, nullptr, C, q, C, 0, 0, 0));
  {
    AttnArgs at;
    at.Q = q; at.K = kctx; at.V = vctx;
    at.batch = N; at.heads = w.heads; at.Tq = (int)T; at.Tk = ctx_len; at.d = d; at.O = o;
    at.softmax_axis = c->softmax_axis;
    at.kv_batch_stride = n_ctx == 1 ? 0 : (long long)ctx_len * C;
    TRY(op_attention(c, at));
  }
  WALLOC(u3, M * C);
  TRY(linear(c, o, M, C, ps.w(w.o), ps.w(w.o + 1), C, u3, C, u2, 0, 0, 0, 0, &ln));
  WALLOC(g, M * 4 * C);
  // GEGLU: Linear(C -> 8C), chunk(2,2), out * gelu(gate)  (diffusion.mojo:138-141)
  TRY(ln_then_linear(u3, w.ln3, w.geglu1, ps.w(w.geglu1 + 1), 8 * C, g, 4 * C, 1, 0, 0));
  if (next) next->imgs = N;
  if (c->fuse_ffn_out && C % 64 == 0) {
    // geglu2 (4C -> C, + u3) and conv_out (1x1, C -> C, + x) have nothing between them (diffusion.mojo:141-146):
    //   out = (g W2^T + b2 + u3) Wc^T + bc + x = [g | u3] [Wc W2 | Wc]^T + (Wc b2 + bc) + x
    // is ONE GEMM over K = 4C + C (second K segment = u3) with weights merged once per weight load - same FLOPs, one
    // launch and one activation round trip fewer per attention block
    const float *wm = nullptr, *bm = nullptr;
    TRY(const_cast<ParamStore&>(ps).fused_ffn_out(w.geglu2, w.conv_out, &wm, &bm));
    TRY(conv(c, ps, w.conv_out, g, N, x.H, x.W, 4 * C, C, 1, 0, 1, bm, 0, x.p, out, 0, next, -1, u3, C, wm));
    c->arena.release_to(mark);
    return TSD_OK;
  }
  float* u4 = u;  // u is dead after u2
  TRY(linear(c, g, M, 4 * C, ps.w(w.geglu2), ps.w(w.geglu2 + 1), C, u4, C, u3, 1));
  TRY(linear(c, u4, M, C, ps.w(w.conv_out), ps.w(w.conv_out + 1), C, out, C, x.p, 0, 0, 0, 0, next));
  c->arena.release_to(mark);
  return TSD_OK;
}

// ------------------------------------------------------------------------------------------
// Diffusion
// ------------------------------------------------------------------------------------------
static const int kResIn[9] = {320, 320, 640, 2560, 1920, 1280, 960, 640, 640};
static const int kResOut[9] = {320, 640, 1280, 1280, 1280, 640, 640, 320, 320};
static const int kResLayer[9] = {2, 5, 8, 10, 12, 15, 17, 20, 22};
static const int kAttnC[9] = {320, 640, 1280, 1280, 1280, 640, 640, 320, 320};
static const int kAttnLayer[9] = {3, 6, 9, 11, 13, 16, 18, 21, 23};

int Diffusion::create() {
  ps.c = c;
  if (cfg.latent_h <= 0 || cfg.latent_w <= 0 || cfg.latent_h % 4 || cfg.latent_w % 4)
    return c->fail(TSD_ERR_INVALID,
                   "diffusion: latent height/width must be positive multiples of 4 (two stride-2 convs and two "
                   "x2 upsamples must round-trip, SURVEY Q8)");
  if (cfg.max_batch <= 0) return c->fail(TSD_ERR_INVALID, "diffusion: max_batch must be positive");
  if (cfg.context_len <= 0 || cfg.context_dim <= 0 || cfg.context_dim % 4)
    return c->fail(TSD_ERR_INVALID, "diffusion: bad context shape");
  // parameter order = struct declaration order (diffusion.mojo:295-297, 151-173, 25-30, 76-85)
  te1 = ps.add_linear("time_embed.layer1", 320, 1280, true);
  te2 = ps.add_linear("time_embed.layer2", 1280, 1280, true);
  int ri = 0, ai = 0;
  char nm[64];
  for (int layer = 1; layer <= 23; ++layer) {
    snprintf(nm, sizeof nm, "unet.layer%d", layer);
    std::string base(nm);
    if (layer == 1) {
      conv_in = ps.add_conv(base, 4, 320, 3);
    } else if (layer == 4) {
      down1 = ps.add_conv(base, 320, 320, 3);
    } else if (layer == 7) {
      down2 = ps.add_conv(base, 640, 640, 3);
    } else if (layer == 14 || layer == 19) {
      // Upsample owns no parameters (helpers/utils.mojo:1979-1987)
    } else if (ri < 9 && kResLayer[ri] == layer) {
      ResBlockW& w = res[ri];
      w.cin = kResIn[ri];
      w.cout = kResOut[ri];
      w.groups = 32;
      // struct order of Unet_Residual_Block (diffusion.mojo:25-30): layer1 GroupNorm, layer2 conv, layer3 linear,
      // layer4 GroupNorm, layer5 conv, layer6 conv
      if (cfg.norm_affine) w.gn1 = ps.add_norm(base + ".layer1", w.cin);
      w.conv1 = ps.add_conv(base + ".layer2", w.cin, w.cout, 3);
      w.lin_t = ps.add_linear(base + ".layer3", 1280, w.cout, true);
      if (cfg.norm_affine) w.gn2 = ps.add_norm(base + ".layer4", w.cout);
      w.conv2 = ps.add_conv(base + ".layer5", w.cout, w.cout, 3);
      // layer6 (1x1 skip conv) is allocated by the reference for every block but only used
      // when in != out (diffusion.mojo:42, 70-72): unused tensors are not part of the blob
      if (w.cin != w.cout) w.skip = ps.add_conv(base + ".layer6", w.cin, w.cout, 1);
      ++ri;
    } else if (ai < 9 && kAttnLayer[ai] == layer) {
      AttnBlockW& w = attn[ai];
      w.C = kAttnC[ai];
      w.heads = 8;
      const int C = w.C;
      // struct order of Unet_Attention_Block (diffusion.mojo:76-85): layer1 GroupNorm, layer2 conv, layer3 LayerNorm,
      // layer4 self-attention, layer5 LayerNorm, layer6 cross-attention, layer7 LayerNorm, layer8, layer9, layer10
      if (cfg.norm_affine) w.gn = ps.add_norm(base + ".layer1", C);
      w.conv_in = ps.add_conv(base + ".layer2", C, C, 1);
      if (cfg.norm_affine) w.ln1 = ps.add_norm(base + ".layer3", C);
      w.in_proj = ps.add_linear(base + ".layer4.in_proj", C, 3 * C, false);
      w.out_proj = ps.add_linear(base + ".layer4.out_proj", C, C, true);
      if (cfg.norm_affine) w.ln2 = ps.add_norm(base + ".layer5", C);
      w.q = ps.add_linear(base + ".layer6.q_proj", C, C, false);
      w.k = ps.add_linear(base + ".layer6.k_proj", cfg.context_dim, C, false);
      w.v = ps.add_linear(base + ".layer6.v_proj", cfg.context_dim, C, false);
      w.o = ps.add_linear(base + ".layer6.out_proj", C, C, true);
      if (cfg.norm_affine) w.ln3 = ps.add_norm(base + ".layer7", C);
      w.geglu1 = ps.add_linear(base + ".layer8", C, 8 * C, true);
      w.geglu2 = ps.add_linear(base + ".layer9", 4 * C, C, true);
      w.conv_out = ps.add_conv(base + ".layer10", C, C, 1);
      ++ai;
    }
  }
  if (cfg.norm_affine) final_gn = ps.add_norm("final.layer1", 320);
  final_conv = ps.add_conv("final.layer2", 320, 4, 3);

  const size_t B = cfg.max_batch, HW = (size_t)cfg.latent_h * cfg.latent_w;
  auto dmalloc = [&](float** p, size_t n) {
    return cudaMalloc(p, n * sizeof(float)) == cudaSuccess ? TSD_OK : c->fail(TSD_ERR_OOM, "diffusion: buffer allocation failed");
  };
  TRY(dmalloc(&x_in, B * 4 * HW));
  TRY(dmalloc(&out_nchw, B * 4 * HW));
  TRY(dmalloc(&x_nhwc, B * 4 * HW));
  TRY(dmalloc(&eps_nhwc, B * 4 * HW));
  TRY(dmalloc(&ctx_in, B * cfg.context_len * cfg.context_dim));
  TRY(dmalloc(&time_in, B * 320));
  TRY(dmalloc(&temb, B * 1280 * 2));
  TRY(dmalloc(&noise_in, B * 4 * HW));
  TRY(dmalloc(&lat_out, B * 4 * HW));
  for (int i = 0; i < 9; ++i) {
    TRY(dmalloc(&kctx[i], B * cfg.context_len * attn[i].C));
    TRY(dmalloc(&vctx[i], B * cfg.context_len * attn[i].C));
    TRY(dmalloc(&tbias[i], B * res[i].cout));
  }
  return TSD_OK;
}

void Diffusion::destroy() {
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  if (graph.exec) cudaGraphExecDestroy(graph.exec);
  float* bufs[] = {x_in, out_nchw, x_nhwc, eps_nhwc, ctx_in, time_in, temb, noise_in, lat_out};
  for (float* b : bufs)
    if (b) cudaFree(b);
  for (int i = 0; i < 9; ++i) {
    if (kctx[i]) cudaFree(kctx[i]);
    if (vctx[i]) cudaFree(vctx[i]);
    if (tbias[i]) cudaFree(tbias[i]);
  }
  ps.free_all();
}

// Cross_Attention k_proj / v_proj of the context (attention.mojo:102-103) depend only on the
// context: hoisted out of the step loop.
int Diffusion::prepare_context(int n_ctx) {
  const long long M = (long long)n_ctx * cfg.context_len;
  for (int i = 0; i < 9; ++i) {
    const AttnBlockW& w = attn[i];
    TRY(linear(c, ctx_in, M, cfg.context_dim, ps.w(w.k), nullptr, w.C, kctx[i], w.C, nullptr, 1));
    TRY(linear(c, ctx_in, M, cfg.context_dim, ps.w(w.v), nullptr, w.C, vctx[i], w.C, nullptr, 1));
  }
  ctx_ready = true;
  ctx_n = n_ctx;
  return TSD_OK;
}

// Time_Embedding.forward (diffusion.mojo:17-21) and the per-block Linear(SiLU(t)) (:61-62) for
// `rows` time vectors at once.  tb_out[r] receives [rows][cout_r] including conv1's bias.
int Diffusion::prepare_time(int rows, const float* time_dev, float* const* tb_out, int) {
  float* t1 = temb;
  float* t2 = temb + (size_t)cfg.max_batch * 1280;
  const size_t mark = c->arena.mark();
  if (rows > cfg.max_batch) {  // loop API: all steps at once
    t1 = walloc(c, (long long)rows * 1280);
    t2 = walloc(c, (long long)rows * 1280);
    if (!t1 || !t2) return c->fail(TSD_ERR_OOM, "workspace exhausted (time embedding)");
  }
  LAUNCH(c, launch_gemv(time_dev, rows, 320, ps.w(te1), ps.w(te1 + 1), nullptr, t1, 1280, 0, 1, c->stream), "gemv");
  LAUNCH(c, launch_gemv(t1, rows, 1280, ps.w(te2), ps.w(te2 + 1), nullptr, t2, 1280, 0, 0, c->stream), "gemv");
  if (!cfg.mojo_alias_time) {
    // the nine block linears read the same SiLU(t_emb): one launch over the concatenated output rows
    GemvMulti gm{};
    gm.nseg = 9;
    int row0 = 0;
    for (int r = 0; r < 9; ++r) {
      gm.seg[r] = GemvSeg{ps.w(res[r].lin_t), ps.w(res[r].lin_t + 1), ps.w(res[r].conv1 + 1), tb_out[r], res[r].cout, row0};
      row0 += res[r].cout;
    }
    gm.total = row0;
    LAUNCH(c, launch_gemv_multi(t2, rows, 1280, gm, 1, c->stream), "gemv_multi");
    c->arena.release_to(mark);
    return TSD_OK;
  }
  float* cur = t2;
  float* other = t1;  // t1 (SiLU(layer1)) is dead once layer2 has run
  for (int r = 0; r < 9; ++r) {
    int silu_in = 1;
    const float* src = cur;
    if (cfg.mojo_alias_time) {
      // SiLU().forward(time) overwrites the shared embedding, block k sees SiLU^k(t)
      LAUNCH(c, launch_unary(cur, other, (long long)rows * 1280, UNARY_SILU, 1.0f, c->stream), "silu");
      float* t = cur; cur = other; other = t;
      src = cur;
      silu_in = 0;
    }
    LAUNCH(c, launch_gemv(src, rows, 1280, ps.w(res[r].lin_t), ps.w(res[r].lin_t + 1),
                          ps.w(res[r].conv1 + 1), tb_out[r], res[r].cout, silu_in, 0, c->stream), "gemv");
  }
  c->arena.release_to(mark);
  return TSD_OK;
}

// UNet.forward + UNet_Output_Layer.forward: x_nhwc [n][H][W][4] -> eps_nhwc [n][H][W][4]
int Diffusion::unet(int n, int n_ctx, int n_time) {
  const int H = cfg.latent_h, W = cfg.latent_w, L = cfg.context_len;
  const int tstride = n_time == 1 ? 0 : 1;
  auto act = [&](int C, int h, int w) {
    Act a;
    a.N = n; a.H = h; a.W = w; a.C = C;
    a.p = walloc(c, a.numel());
    return a;
  };
#define NEED(a) if (!(a).p) return c->fail(TSD_ERR_OOM, "workspace exhausted (unet activation)")
  auto hint_scratch = [&](Act& a) -> bool {  // partial-statistics buffer living as long as the activation
    a.ns.imgs = a.N;
    a.ns.scratch_elems = norm_scratch_elems(a.N, (long long)a.H * a.W, a.C, a.ns.G);
    a.ns.scratch = c->arena.alloc_n<float2>(a.ns.scratch_elems);
    return a.ns.scratch != nullptr;
  };
  // The first reader of these activations is the GroupNorm of the next block: a split-K producer may leave its
  // partial tiles for that norm kernel (NormHint::allow_defer).  Only where split-K happens (<= 32x32 latents).
  auto defer_ws = [&](Act& a) -> bool {
    if (!c->defer_reduce || a.pixels() * a.C > 1024ll * 1280) return true;
    a.ns.imgs = a.N;
    a.ns.defer_ws_elems = (size_t)8 * a.pixels() * a.C;  // up to 8 splits
    a.ns.defer_ws = c->arena.alloc_n<float>(a.ns.defer_ws_elems);
    a.ns.allow_defer = a.ns.defer_ws != nullptr;
    return a.ns.allow_defer;
  };
  auto RES = [&](int i, const Act& in, Act& out_) -> int {
    out_ = act(res[i].cout, in.H, in.W);
    if (!out_.p) return c->fail(TSD_ERR_OOM, "workspace exhausted (unet activation)");
    Act v = in;
    v.C = res[i].cin;
    // every ResBlock of the UNet feeds an attention block: GroupNorm(32, eps 1e-6) (diffusion.mojo:104)
    out_.ns.G = 32;
    out_.ns.eps = 1e-6f;
    if (!hint_scratch(out_)) return c->fail(TSD_ERR_OOM, "workspace exhausted (norm statistics)");
    if (!defer_ws(out_)) return c->fail(TSD_ERR_OOM, "workspace exhausted (split-K partial tiles)");
    return res_block(c, ps, res[i], v, tbias[i], tstride * res[i].cout, 1e-5f, out_.p, &out_.ns);
  };
  auto ATT = [&](int i, const Act& in, Act& out_) -> int {
    out_ = act(attn[i].C, in.H, in.W);
    if (!out_.p) return c->fail(TSD_ERR_OOM, "workspace exhausted (unet activation)");
    NormHint* nh = nullptr;
    if (i == 8) {  // a23 feeds UNet_Output_Layer's GroupNorm(320 groups) (diffusion.mojo:280)
      out_.ns.G = 320;
      out_.ns.eps = 1e-5f;
      if (!hint_scratch(out_)) return c->fail(TSD_ERR_OOM, "workspace exhausted (norm statistics)");
      nh = &out_.ns;
    } else if (i == 4 || i == 6) {
      // a13 / a18 are upsampled (nearest x2) and then normalised by a ResBlock's GroupNorm(32): replicating every
      // pixel four times leaves mean and variance unchanged, so the statistics of the source serve the upsampled tensor
      out_.ns.G = 32;
      out_.ns.eps = 1e-5f;
      if (!hint_scratch(out_)) return c->fail(TSD_ERR_OOM, "workspace exhausted (norm statistics)");
      nh = &out_.ns;
    }
    return attn_block(c, ps, attn[i], in, kctx[i], vctx[i], n_ctx, L, out_.p, nh);
  };
  auto CAT = [&](const Act& a, const Act& b, Act& out_) -> int {
    out_ = act(a.C + b.C, a.H, a.W);
    if (!out_.p) return c->fail(TSD_ERR_OOM, "workspace exhausted (unet activation)");
    if (c->virtual_concat && a.C % 4 == 0 && b.C % 4 == 0 &&
        ((c->norm_cluster && norm_cluster_supported(a.N, (long long)a.H * a.W, a.C + b.C, 32)) ||
         norm_fused2_supported(a.N, (long long)a.H * a.W, a.C + b.C, 32, c->sm_count))) {
      // every concat of the UNet feeds a ResBlock whose first GroupNorm(32, eps 1e-5) is its first reader: that
      // kernel reads the two tensors and writes the concatenation (for the block's convolutions) as a by-product
      out_.ns.G = 32;
      out_.ns.eps = 1e-5f;
      out_.ns.imgs = a.N;
      NormHint::Deferred& d = out_.ns.def;
      d.ws = a.p;
      d.x2 = b.p;
      d.c_a = a.C;
      d.splits = 1;
      d.raw = out_.p;
      d.rows = (int)a.pixels();
      d.C = a.C + b.C;
      return TSD_OK;
    }
    LAUNCH(c, launch_concat_channels(a.p, a.C, b.p, b.C, out_.p, a.pixels(), c->stream), "concat");
    return TSD_OK;
  };
  auto UP = [&](const Act& a, Act& out_) -> int {
    out_ = act(a.C, a.H * 2, a.W * 2);
    if (!out_.p) return c->fail(TSD_ERR_OOM, "workspace exhausted (unet activation)");
    LAUNCH(c, launch_upsample2x(a.p, out_.p, a.N, a.H, a.W, a.C, c->stream), "upsample2x");
    out_.ns = a.ns;  // statistics are invariant under nearest-neighbour replication
    return TSD_OK;
  };

  // encoders (diffusion.mojo:236-250)
  Act s1 = act(320, H, W);
  NEED(s1);
  TRY(conv(c, ps, conv_in, x_nhwc, n, H, W, 4, 320, 3, 1, 1, nullptr, 0, nullptr, s1.p, 0));
  Act r2, a3, r5, a6, r8, a9;
  TRY(RES(0, s1, r2));
  TRY(ATT(0, r2, a3));  // skip2: dead input of layer20 (Q9)
  Act d4 = act(320, H / 2, W / 2);
  NEED(d4);
  d4.ns.G = 32;  // consumed by ResBlock 1's first GroupNorm
  d4.ns.eps = 1e-5f;
  if (!hint_scratch(d4)) return c->fail(TSD_ERR_OOM, "workspace exhausted (norm statistics)");
  if (!defer_ws(d4)) return c->fail(TSD_ERR_OOM, "workspace exhausted (split-K partial tiles)");
  TRY(conv(c, ps, down1, a3.p, n, H, W, 320, 320, 3, 1, 2, nullptr, 0, nullptr, d4.p, 0, &d4.ns));
  TRY(RES(1, d4, r5));
  TRY(ATT(1, r5, a6));  // skip4: dead input of layer15 (Q9)
  Act d7 = act(640, H / 4, W / 4);
  NEED(d7);
  d7.ns.G = 32;
  d7.ns.eps = 1e-5f;
  if (!hint_scratch(d7)) return c->fail(TSD_ERR_OOM, "workspace exhausted (norm statistics)");
  if (!defer_ws(d7)) return c->fail(TSD_ERR_OOM, "workspace exhausted (split-K partial tiles)");
  TRY(conv(c, ps, down2, a6.p, n, H / 2, W / 2, 640, 640, 3, 1, 2, nullptr, 0, nullptr, d7.p, 0, &d7.ns));
  TRY(RES(2, d7, r8));
  TRY(ATT(2, r8, a9));
  // decoders (diffusion.mojo:252-272)
  Act c10, r10, a11, c12, r12, a13, u14, r15, a16, c17, r17, a18, u19, r20, a21, c22, r22, a23;
  TRY(CAT(a9, a9, c10));  // out.concat(skip6) with skip6 == out (Q10)
  TRY(RES(3, c10, r10));
  TRY(ATT(3, r10, a11));
  TRY(CAT(a11, d7, c12));
  TRY(RES(4, c12, r12));
  TRY(ATT(4, r12, a13));
  TRY(UP(a13, u14));       // layer14 ; concat(skip4) adds channels layer15 never reads (Q9)
  TRY(RES(5, u14, r15));
  TRY(ATT(5, r15, a16));
  TRY(CAT(a16, d4, c17));
  TRY(RES(6, c17, r17));
  TRY(ATT(6, r17, a18));
  TRY(UP(a18, u19));       // layer19 ; concat(skip2) dead (Q9)
  TRY(RES(7, u19, r20));
  TRY(ATT(7, r20, a21));
  TRY(CAT(a21, s1, c22));
  TRY(RES(8, c22, r22));
  TRY(ATT(8, r22, a23));
  // UNet_Output_Layer: GroupNorm(320 groups) -> SiLU -> conv 320->4 (diffusion.mojo:280, 287-291)
  Act f = act(320, H, W);
  NEED(f);
  TRY(op_group_norm(c, a23.p, f.p, n, H, W, 320, 320, 1e-5f, ps.gamma(final_gn), ps.beta(final_gn), 1.0f, 1, 0, 1,
                    (a23.ns.G == 320) ? a23.ns.ready() : nullptr));
  TRY(conv(c, ps, final_conv, f.p, n, H, W, 320, 4, 3, 1, 1, nullptr, 0, nullptr, eps_nhwc, 0));
#undef NEED
  return TSD_OK;
}

size_t Diffusion::workspace_bytes(int n) const {
  // planning pass over the same composition: exact high-water mark of the bump allocator
  Diffusion* self = const_cast<Diffusion*>(this);
  Ctx* cc = self->c;
  if (n < (int)self->ws_cache.size() && self->ws_cache[n] && self->ws_epoch == h->option_epoch) return self->ws_cache[n];
  if (self->ws_epoch != h->option_epoch) self->ws_cache.clear();
  self->ws_epoch = h->option_epoch;
  const bool was = cc->dry_run;
  cc->dry_run = true;
  cc->arena.set_virtual(true);
  int rc = self->unet(n, 1, 1);
  size_t hw = cc->arena.high_water();
  cc->arena.set_virtual(false);
  cc->dry_run = was;
  if (rc) return 0;
  if ((int)self->ws_cache.size() <= n) self->ws_cache.resize(n + 1, 0);
  self->ws_cache[n] = hw + (64u << 20);
  return self->ws_cache[n];
}

static long long g_launches_before_capture = 0;
static int capture_begin(Ctx* c) {
  g_launches_before_capture = c->launches;
  return c->check(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal), "graph capture begin");
}
static int capture_end(Ctx* c, GraphSlot* slot) {
  cudaGraph_t g = nullptr;
  int rc = c->check(cudaStreamEndCapture(c->stream, &g), "graph capture end");
  if (rc) return rc;
  if (slot->exec) {
    cudaGraphExecDestroy(slot->exec);
    slot->exec = nullptr;
  }
  rc = c->check(cudaGraphInstantiate(&slot->exec, g, 0), "graph instantiate");
  cudaGraphDestroy(g);
  // kernels recorded (not run) during capture: remember the count, credit it per replay
  slot->nodes = c->launches - g_launches_before_capture;
  c->launches = g_launches_before_capture;
  return rc;
}

// Everything of a forward that works on the handle's own fixed buffers: time-embedding MLP + the nine block linears
// (time_in -> tbias), NCHW -> NHWC of the latents (x_in -> x_nhwc), the UNet, NHWC -> NCHW of the result (eps_nhwc ->
// out_nchw).  One captured graph replays all of it: the six small launches around the UNet no longer pay a stream-launch
// gap each.
int Diffusion::core(int n, int n_ctx, int n_time) {
  const int HW = cfg.latent_h * cfg.latent_w;
  TRY(prepare_time(n_time, time_in, tbias, 0));
  LAUNCH(c, launch_nchw_to_nhwc(x_in, x_nhwc, n, 4, HW, c->stream), "nchw_to_nhwc");
  TRY(unet(n, n_ctx, n_time));
  LAUNCH(c, launch_nhwc_to_nchw(eps_nhwc, out_nchw, n, 4, HW, c->stream), "nhwc_to_nchw");
  return TSD_OK;
}

int Diffusion::run_unet_graph(int n, int n_ctx, int n_time) {
  c->arena.reset();
  if (!h->use_graph || c->timer) return core(n, n_ctx, n_time);
  GraphSlot& g = graph;
  const bool valid = g.exec && g.n == n && g.n_ctx == n_ctx && g.n_time == n_time && g.epoch == h->option_epoch &&
                     g.arena_base == c->arena.base() && g.weights_gen == ps.gen;
  if (!valid) {
    // one eager pass first: validates shapes, initialises per-kernel attributes outside capture
    TRY(core(n, n_ctx, n_time));
    TRY(c->check(cudaStreamSynchronize(c->stream), "unet eager pass"));
    c->arena.reset();
    TRY(capture_begin(c));
    int rc = core(n, n_ctx, n_time);
    int rc2 = capture_end(c, &g);
    if (rc) return rc;
    if (rc2) return rc2;
    g.n = n; g.n_ctx = n_ctx; g.n_time = n_time; g.epoch = h->option_epoch; g.arena_base = c->arena.base();
    g.weights_gen = ps.gen;
    return TSD_OK;  // the eager pass already produced this call's result
  }
  TRY(c->check(cudaGraphLaunch(g.exec, c->stream), "graph launch"));
  c->launches += g.nodes;
  return TSD_OK;
}

int Diffusion::forward_dev(const float* x, const float* context, int n_ctx, const float* time, int n_time, int n,
                           float* out, bool host_ptrs) {
  if (!ps.loaded) return c->fail(TSD_ERR_STATE, "diffusion: forward before load_weights / init_random");
  if (n <= 0 || n > cfg.max_batch) return c->fail(TSD_ERR_INVALID, "diffusion: batch exceeds max_batch");
  if (!x || !out || !time) return c->fail(TSD_ERR_INVALID, "diffusion: null buffer");
  if (!(n_time == 1 || n_time == n)) return c->fail(TSD_ERR_INVALID, "diffusion: n_time must be 1 or n");
  if (context && !(n_ctx == 1 || n_ctx == n)) return c->fail(TSD_ERR_INVALID, "diffusion: n_ctx must be 1 or n");
  if (!context && !ctx_ready) return c->fail(TSD_ERR_STATE, "diffusion: no context set");
  cudaSetDevice(c->device);
  const size_t need = workspace_bytes(n);
  if (need == 0) return TSD_ERR_OOM;
  if (need > c->arena.capacity()) {
    cudaStreamSynchronize(c->stream);
    if (c->arena.reserve(need) != TSD_OK) return c->fail(TSD_ERR_OOM, "diffusion: workspace allocation failed");
  }
  const cudaMemcpyKind kin = host_ptrs ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
  const cudaMemcpyKind kout = host_ptrs ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
  const size_t HW = (size_t)cfg.latent_h * cfg.latent_w;
  TRY(c->check(cudaMemcpyAsync(x_in, x, n * 4 * HW * sizeof(float), kin, c->stream), "copy x"));
  TRY(c->check(cudaMemcpyAsync(time_in, time, (size_t)n_time * 320 * sizeof(float), kin, c->stream), "copy time"));
  c->arena.reset();
  if (context && host_ptrs) {
    TRY(upload_context_cached(context, n_ctx));
  } else if (context) {
    TRY(c->check(cudaMemcpyAsync(ctx_in, context, (size_t)n_ctx * cfg.context_len * cfg.context_dim * sizeof(float),
                                 kin, c->stream), "copy context"));
    TRY(prepare_context(n_ctx));
    ctx_hash_gen = -1;  // device-side context: the host-content cache no longer describes kctx/vctx
  }
  TRY(run_unet_graph(n, ctx_n, n_time));
  TRY(c->check(cudaMemcpyAsync(out, out_nchw, n * 4 * HW * sizeof(float), kout, c->stream), "copy out"));
  if (host_ptrs) TRY(c->check(cudaStreamSynchronize(c->stream), "diffusion forward sync"));
  return TSD_OK;
}

// 64-bit content hash of a host buffer (four independent multiply-xorshift lanes over 8-byte words: ~10 GB/s, i.e.
// ~25 us for a 77x768 context against a 2.5 ms step)
static uint64_t hash_bytes(const void* p, size_t nbytes) {
  const uint64_t* w = static_cast<const uint64_t*>(p);
  const size_t nw = nbytes / 8;
  uint64_t h[4] = {0x9E3779B97F4A7C15ull, 0xBF58476D1CE4E5B9ull, 0x94D049BB133111EBull, 0xD6E8FEB86659FD93ull};
  size_t i = 0;
  for (; i + 4 <= nw; i += 4)
    for (int k = 0; k < 4; ++k) {
      uint64_t v = (h[k] ^ w[i + k]) * 0xFF51AFD7ED558CCDull;
      h[k] = v ^ (v >> 29);
    }
  for (; i < nw; ++i) {
    uint64_t v = (h[0] ^ w[i]) * 0xFF51AFD7ED558CCDull;
    h[0] = v ^ (v >> 29);
  }
  uint64_t tail = 0;
  memcpy(&tail, static_cast<const uint8_t*>(p) + nw * 8, nbytes - nw * 8);
  uint64_t r = (h[0] ^ tail) * 0xC4CEB9FE1A85EC53ull;
  for (int k = 1; k < 4; ++k) r = (r ^ (r >> 31) ^ h[k]) * 0xC4CEB9FE1A85EC53ull;
  return r ^ (r >> 33) ^ nbytes;
}

int Diffusion::upload_context_cached(const float* context_host, int n_ctx) {
  const size_t bytes = (size_t)n_ctx * cfg.context_len * cfg.context_dim * sizeof(float);
  const uint64_t hsh = hash_bytes(context_host, bytes);
  if (ctx_ready && ctx_hash_gen == ps.gen && ctx_hash_n == n_ctx && ctx_n == n_ctx && ctx_hash == hsh) return TSD_OK;
  TRY(c->check(cudaMemcpyAsync(ctx_in, context_host, bytes, cudaMemcpyHostToDevice, c->stream), "copy context"));
  TRY(prepare_context(n_ctx));
  ctx_hash = hsh;
  ctx_hash_n = n_ctx;
  ctx_hash_gen = ps.gen;
  return TSD_OK;
}

int Diffusion::step_host(const float* latents, const float* context, int n_ctx, const float* time, const float* noise,
                         int use_cfg, float cfg_scale, const float coef[5], int n, float* latents_out) {
  if (!ps.loaded) return c->fail(TSD_ERR_STATE, "diffusion: step before load_weights / init_random");
  const int nb = use_cfg ? 2 * n : n;
  if (n <= 0 || nb > cfg.max_batch) return c->fail(TSD_ERR_INVALID, "diffusion step: batch (x2 with cfg) exceeds max_batch");
  if (!latents || !time || !latents_out) return c->fail(TSD_ERR_INVALID, "diffusion step: null buffer");
  const int groups = use_cfg ? 2 : 1;
  if (context && !(n_ctx == groups || n_ctx == nb))
    return c->fail(TSD_ERR_INVALID, "diffusion step: n_ctx must be (1|n) without cfg, (2|2n) with cfg (cond rows, then uncond rows)");
  if (!context && !ctx_ready) return c->fail(TSD_ERR_STATE, "diffusion step: no context set");
  cudaSetDevice(c->device);
  const size_t need = workspace_bytes(nb);
  if (need == 0) return TSD_ERR_OOM;
  if (need > c->arena.capacity()) {
    cudaStreamSynchronize(c->stream);
    if (c->arena.reserve(need) != TSD_OK) return c->fail(TSD_ERR_OOM, "diffusion: workspace allocation failed");
  }
  cudaStream_t s = c->stream;
  const size_t HW = (size_t)cfg.latent_h * cfg.latent_w, n_lat = (size_t)n * 4 * HW;
  TRY(c->check(cudaMemcpyAsync(x_in, latents, n_lat * 4, cudaMemcpyHostToDevice, s), "copy latents"));
  if (use_cfg) TRY(c->check(cudaMemcpyAsync(x_in + n_lat, x_in, n_lat * 4, cudaMemcpyDeviceToDevice, s), "copy latents"));
  TRY(c->check(cudaMemcpyAsync(time_in, time, 320 * sizeof(float), cudaMemcpyHostToDevice, s), "copy time"));
  if (noise) TRY(c->check(cudaMemcpyAsync(noise_in, noise, n_lat * 4, cudaMemcpyHostToDevice, s), "copy noise"));
  c->arena.reset();
  if (context) {
    if (n_ctx == 1 || n_ctx == nb) {
      TRY(upload_context_cached(context, n_ctx));
    } else {
      // (cond, uncond) shared by all images: one row per UNet batch entry, cond rows first
      const size_t row = (size_t)cfg.context_len * cfg.context_dim;
      const uint64_t hsh = hash_bytes(context, 2 * row * sizeof(float)) ^ 0x5bd1e995u;
      if (!(ctx_ready && ctx_hash_gen == ps.gen && ctx_hash_n == -nb && ctx_n == nb && ctx_hash == hsh)) {
        for (int i = 0; i < nb; ++i)
          TRY(c->check(cudaMemcpyAsync(ctx_in + (size_t)i * row, context + (size_t)(i / n) * row, row * sizeof(float),
                                       cudaMemcpyHostToDevice, s), "copy context"));
        TRY(prepare_context(nb));
        ctx_hash = hsh;
        ctx_hash_n = -nb;
        ctx_hash_gen = ps.gen;
      }
    }
  }
  TRY(run_unet_graph(nb, ctx_n, 1));
  LAUNCH(c, launch_ddpm_step(x_in, out_nchw, use_cfg ? out_nchw + n_lat : nullptr, cfg_scale, noise ? noise_in : nullptr,
                             coef[0], coef[1], coef[2], coef[3], coef[4], lat_out, (long long)n_lat, s), "ddpm_step");
  TRY(c->check(cudaMemcpyAsync(latents_out, lat_out, n_lat * 4, cudaMemcpyDeviceToHost, s), "copy latents out"));
  return c->check(cudaStreamSynchronize(s), "diffusion step sync");
}

// ------------------------------------------------------------------------------------------
// Decoder
// ------------------------------------------------------------------------------------------
int Decoder::create() {
  ps.c = c;
  if (latent_h <= 0 || latent_w <= 0 || max_batch <= 0) return c->fail(TSD_ERR_INVALID, "decoder: bad shape");
  // parameter order = Decoder struct order l1..l26 (vae.mojo:163-188)
  int ri = 0;
  auto add_res = [&](const char* name, int cin, int cout) {
    ResBlockW& w = res[ri++];
    w.cin = cin; w.cout = cout; w.groups = 16;  // Res_Block GroupNorm(16, .) (vae.mojo:42-43, Q17)
    std::string b(name);
    w.conv1 = ps.add_conv(b + ".conv1", cin, cout, 3);
    w.conv2 = ps.add_conv(b + ".conv2", cout, cout, 3);
    if (cin != cout) w.skip = ps.add_conv(b + ".res_conv_layer", cin, cout, 1);
    if (norm_affine) {
      w.gn1 = ps.add_norm(b + ".groupnorm1", cin);
      w.gn2 = ps.add_norm(b + ".groupnorm2", cout);
    }
  };
  l1 = ps.add_conv("l1", 4, 4, 1);
  l2 = ps.add_conv("l2", 4, 512, 3);
  add_res("l3", 512, 512);
  attn_in = ps.add_linear("l4.attention.in_proj", 512, 1536, true);
  attn_out = ps.add_linear("l4.attention.out_proj", 512, 512, true);
  if (norm_affine) attn_gn = ps.add_norm("l4.groupnorm", 512);
  add_res("l5", 512, 512);
  add_res("l6", 512, 512);
  add_res("l7", 512, 512);
  add_res("l8", 512, 512);
  l10 = ps.add_conv("l10", 512, 512, 3);
  add_res("l11", 512, 512);
  add_res("l12", 512, 512);
  add_res("l13", 512, 512);
  l15 = ps.add_conv("l15", 512, 512, 3);
  add_res("l16", 512, 256);
  add_res("l17", 256, 256);
  add_res("l18", 256, 256);
  l20 = ps.add_conv("l20", 256, 256, 3);
  add_res("l21", 256, 128);
  add_res("l22", 128, 128);
  add_res("l23", 128, 128);
  if (norm_affine) out_gn = ps.add_norm("l24", 128);
  l26 = ps.add_conv("l26", 128, 3, 3);

  const size_t B = max_batch, hw = (size_t)latent_h * latent_w;
  pp_elems = B * hw * 16384;  // largest activation: 256 ch at 8h x 8w
  auto dmalloc = [&](float** p, size_t n) {
    return cudaMalloc(p, n * sizeof(float)) == cudaSuccess ? TSD_OK : c->fail(TSD_ERR_OOM, "decoder: buffer allocation failed");
  };
  TRY(dmalloc(&z_in, B * 4 * hw));
  TRY(dmalloc(&img_out, B * 3 * 64 * hw));
  TRY(dmalloc(&ping, pp_elems));
  TRY(dmalloc(&pong, pp_elems));
  return TSD_OK;
}
void Decoder::destroy() {
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  if (graph.exec) cudaGraphExecDestroy(graph.exec);
  float* bufs[] = {z_in, img_out, ping, pong};
  for (float* b : bufs)
    if (b) cudaFree(b);
  ps.free_all();
}

// Attention_Block.forward, vae.mojo:17-27: GroupNorm(32) -> 1-head self-attention -> + residue
static int vae_attn_block(Ctx* c, const ParamStore& ps, int attn_in, int attn_out, int attn_gn, const float* x, int n,
                          int H, int W, float* out, const NormHint* xns = nullptr, NormHint* next = nullptr) {
  const int C = 512;
  const long long T = (long long)H * W, M = n * T;
  const size_t mark = c->arena.mark();
  WALLOC(a, M * C);
  TRY(op_group_norm(c, x, a, n, H, W, C, 32, 1e-5f, ps.gamma(attn_gn), ps.beta(attn_gn), 1.0f, 0, 0, 1,
                    (xns && xns->G == 32 && xns->eps == 1e-5f) ? xns->ready() : nullptr));
  WALLOC(qkv, 3 * M * C);
  TRY(linear(c, a, M, C, ps.w(attn_in), ps.w(attn_in + 1), 3 * C, qkv, C, nullptr, 1, 0, C, M * C));
  WALLOC(o, M * C);
  AttnArgs at;
  at.Q = qkv; at.K = qkv + M * C; at.V = qkv + 2 * M * C;
  at.batch = n; at.heads = 1; at.Tq = (int)T; at.Tk = (int)T; at.d = C; at.O = o;
  at.softmax_axis = c->softmax_axis;
  TRY(op_attention(c, at));
  TRY(linear(c, o, M, C, ps.w(attn_out), ps.w(attn_out + 1), C, out, C, x, 0, 0, 0, 0, next));
  c->arena.release_to(mark);
  return TSD_OK;
}

// Decoder.forward, vae.mojo:221-250; z_in (n,4,h,w) -> img_out (n,3,8h,8w)
int Decoder::decode(int n, int rescale) {
  int H = latent_h, W = latent_w;
  float* cur = ping;
  float* nxt = pong;
  auto swap = [&]() { float* t = cur; cur = nxt; nxt = t; };
  c->arena.reset();
  // x / 0.18215 (vae.mojo:222) then to NHWC
  {
    const size_t mark = c->arena.mark();
    WALLOC(zs, (long long)n * 4 * H * W);
    LAUNCH(c, launch_unary(z_in, zs, (long long)n * 4 * H * W, UNARY_SCALE, 1.0f / 0.18215f, c->stream), "scale");
    LAUNCH(c, launch_nchw_to_nhwc(zs, cur, n, 4, H * W, c->stream), "nchw_to_nhwc");
    c->arena.release_to(mark);
  }
  TRY(conv(c, ps, l1, cur, n, H, W, 4, 4, 1, 0, 1, nullptr, 0, nullptr, nxt, 0));
  swap();
  TRY(conv(c, ps, l2, cur, n, H, W, 4, 512, 3, 1, 1, nullptr, 0, nullptr, nxt, 0));
  swap();
  int ri = 0;
  // Producer-side norm statistics along the chain (as in the UNet): every GEMM whose output is normalised next leaves
  // the partial sums for that norm (`groups` of the consumer; 0 = the consumer is not a norm).  cur_ns describes `cur`.
  NormHint cur_ns;
  auto make_hint = [&](NormHint& h, int groups, int C, int hh, int ww) -> bool {
    h = NormHint();
    if (groups <= 0) return true;
    h.G = groups;
    h.eps = 1e-5f;
    h.imgs = n;
    h.scratch_elems = norm_scratch_elems(n, (long long)hh * ww, C, groups);
    h.scratch = c->arena.alloc_n<float2>(h.scratch_elems);  // lives until the next decode resets the arena
    return h.scratch != nullptr;
  };
  auto RES = [&](int next_groups) -> int {
    Act x;
    x.p = cur; x.N = n; x.H = H; x.W = W; x.C = res[ri].cin;
    x.ns = cur_ns;
    NormHint out_ns;
    if (!make_hint(out_ns, next_groups, res[ri].cout, H, W)) return c->fail(TSD_ERR_OOM, "workspace exhausted (norm statistics)");
    int rc = res_block(c, ps, res[ri], x, nullptr, 0, 1e-5f, nxt, next_groups > 0 ? &out_ns : nullptr);
    cur_ns = out_ns;
    ++ri;
    swap();
    return rc;
  };
  auto UPCONV = [&](int wi, int cin, int cout, int next_groups) -> int {
    NormHint out_ns;
    if (!make_hint(out_ns, next_groups, cout, 2 * H, 2 * W)) return c->fail(TSD_ERR_OOM, "workspace exhausted (norm statistics)");
    const size_t mark = c->arena.mark();
    WALLOC(up, (long long)n * 4 * H * W * cin);
    LAUNCH(c, launch_upsample2x(cur, up, n, H, W, cin, c->stream), "upsample2x");
    H *= 2;
    W *= 2;
    TRY(conv(c, ps, wi, up, n, H, W, cin, cout, 3, 1, 1, nullptr, 0, nullptr, nxt, 0, next_groups > 0 ? &out_ns : nullptr));
    c->arena.release_to(mark);
    cur_ns = out_ns;
    swap();
    return TSD_OK;
  };
  TRY(RES(32));  // l3 -> Attention_Block's GroupNorm(32)
  {
    NormHint out_ns;
    if (!make_hint(out_ns, 16, 512, H, W)) return c->fail(TSD_ERR_OOM, "workspace exhausted (norm statistics)");
    TRY(vae_attn_block(c, ps, attn_in, attn_out, attn_gn, cur, n, H, W, nxt, &cur_ns, &out_ns));  // l4
    cur_ns = out_ns;
  }
  swap();
  for (int i = 0; i < 4; ++i) TRY(RES(i < 3 ? 16 : 0));   // l5..l8 (l8 feeds the upsample)
  TRY(UPCONV(l10, 512, 512, 16));                         // l9, l10
  for (int i = 0; i < 3; ++i) TRY(RES(i < 2 ? 16 : 0));   // l11..l13
  TRY(UPCONV(l15, 512, 512, 16));                         // l14, l15
  for (int i = 0; i < 3; ++i) TRY(RES(i < 2 ? 16 : 0));   // l16..l18
  TRY(UPCONV(l20, 256, 256, 16));                         // l19, l20
  for (int i = 0; i < 3; ++i) TRY(RES(i < 2 ? 16 : 32));  // l21..l23 (l23 -> l24 GroupNorm(32))
  {
    // l24 GroupNorm(32,128), l25 SiLU, l26 conv 128->3, then rescale/clamp (pipeline.mojo:127)
    const size_t mark = c->arena.mark();
    WALLOC(f, (long long)n * H * W * 128);
    TRY(op_group_norm(c, cur, f, n, H, W, 128, 32, 1e-5f, ps.gamma(out_gn), ps.beta(out_gn), 1.0f, 1, 0, 1,
                      (cur_ns.G == 32 && cur_ns.eps == 1e-5f) ? cur_ns.ready() : nullptr));
    TRY(conv(c, ps, l26, f, n, H, W, 128, 3, 3, 1, 1, nullptr, 0, nullptr, nxt, 0));
    c->arena.release_to(mark);
    LAUNCH(c, launch_rescale_to_nchw(nxt, img_out, n, 3, H * W, rescale, c->stream), "rescale_to_nchw");
  }
  return TSD_OK;
}

size_t Decoder::workspace_bytes(int n) const {
  Decoder* self = const_cast<Decoder*>(this);
  Ctx* cc = self->c;
  const bool was = cc->dry_run;
  cc->dry_run = true;
  cc->arena.set_virtual(true);
  int rc = self->decode(n, 0);
  size_t hw = cc->arena.high_water();
  cc->arena.set_virtual(false);
  cc->dry_run = was;
  return rc ? 0 : hw + (64u << 20);
}

int Decoder::forward(const float* z, int n, int rescale, float* img, bool host_ptrs) {
  if (!ps.loaded) return c->fail(TSD_ERR_STATE, "decoder: forward before load_weights / init_random");
  if (n <= 0 || n > max_batch) return c->fail(TSD_ERR_INVALID, "decoder: batch exceeds max_batch");
  if (!z || !img) return c->fail(TSD_ERR_INVALID, "decoder: null buffer");
  cudaSetDevice(c->device);
  const size_t need = workspace_bytes(n);
  if (need == 0) return TSD_ERR_OOM;
  if (need > c->arena.capacity()) {
    cudaStreamSynchronize(c->stream);
    if (c->arena.reserve(need) != TSD_OK) return c->fail(TSD_ERR_OOM, "decoder: workspace allocation failed");
  }
  const size_t hw = (size_t)latent_h * latent_w;
  TRY(c->check(cudaMemcpyAsync(z_in, z, n * 4 * hw * sizeof(float),
                               host_ptrs ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, c->stream), "copy z"));
  GraphSlot& g = graph;
  const bool use = h->use_graph && !c->timer;
  const bool valid = g.exec && g.n == n && g.n_ctx == rescale && g.epoch == h->option_epoch &&
                     g.arena_base == c->arena.base() && g.weights_gen == ps.gen;
  if (!use) {
    TRY(decode(n, rescale));
  } else if (!valid) {
    TRY(decode(n, rescale));
    TRY(c->check(cudaStreamSynchronize(c->stream), "decoder eager pass"));
    TRY(capture_begin(c));
    int rc = decode(n, rescale);
    int rc2 = capture_end(c, &g);
    if (rc) return rc;
    if (rc2) return rc2;
    g.n = n; g.n_ctx = rescale; g.epoch = h->option_epoch; g.arena_base = c->arena.base();
    g.weights_gen = ps.gen;
  } else {
    TRY(c->check(cudaGraphLaunch(g.exec, c->stream), "graph launch"));
    c->launches += g.nodes;
  }
  TRY(c->check(cudaMemcpyAsync(img, img_out, n * 3 * 64 * hw * sizeof(float),
                               host_ptrs ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, c->stream), "copy img"));
  if (host_ptrs) TRY(c->check(cudaStreamSynchronize(c->stream), "decoder forward sync"));
  return TSD_OK;
}

// ---------------------------------------------------------------------------------------
// VAE Encoder (vae.mojo:70-159)
// ---------------------------------------------------------------------------------------
int Encoder::create() {
  ps.c = c;
  if (latent_h <= 0 || latent_w <= 0 || max_batch <= 0) return c->fail(TSD_ERR_INVALID, "encoder: bad shape");
  // parameter order = Encoder struct order l1..l19 (vae.mojo:71-89, 94-112)
  int ri = 0;
  auto add_res = [&](const char* name, int cin, int cout) {
    ResBlockW& w = res[ri++];
    w.cin = cin; w.cout = cout; w.groups = 16;  // Res_Block GroupNorm(16, .) (vae.mojo:42-43, Q17)
    std::string b(name);
    w.conv1 = ps.add_conv(b + ".conv1", cin, cout, 3);
    w.conv2 = ps.add_conv(b + ".conv2", cout, cout, 3);
    if (cin != cout) w.skip = ps.add_conv(b + ".res_conv_layer", cin, cout, 1);
    if (norm_affine) {
      w.gn1 = ps.add_norm(b + ".groupnorm1", cin);
      w.gn2 = ps.add_norm(b + ".groupnorm2", cout);
    }
  };
  l1 = ps.add_conv("l1", 3, 128, 3);
  add_res("l2", 128, 128);
  add_res("l3", 128, 128);
  l4 = ps.add_conv("l4", 128, 128, 3);
  add_res("l5", 128, 256);
  add_res("l6", 256, 256);
  l7 = ps.add_conv("l7", 256, 256, 3);
  add_res("l8", 256, 512);
  add_res("l9", 512, 512);
  l10 = ps.add_conv("l10", 512, 512, 3);
  add_res("l11", 512, 512);
  add_res("l12", 512, 512);
  add_res("l13", 512, 512);
  attn_in = ps.add_linear("l14.attention.in_proj", 512, 1536, true);
  attn_out = ps.add_linear("l14.attention.out_proj", 512, 512, true);
  if (norm_affine) attn_gn = ps.add_norm("l14.groupnorm", 512);
  add_res("l15", 512, 512);
  if (norm_affine) out_gn = ps.add_norm("l16", 512);
  l18 = ps.add_conv("l18", 512, 8, 3);
  l19 = ps.add_conv("l19", 8, 8, 1);

  const size_t B = max_batch, hw = (size_t)latent_h * latent_w;
  pp_elems = B * hw * 64 * 128;  // largest activation: 128 ch at 8h x 8w
  auto dmalloc = [&](float** p, size_t n) {
    return cudaMalloc(p, n * sizeof(float)) == cudaSuccess ? TSD_OK : c->fail(TSD_ERR_OOM, "encoder: buffer allocation failed");
  };
  TRY(dmalloc(&img_in, B * 3 * 64 * hw));
  TRY(dmalloc(&noise_in, B * 4 * hw));
  TRY(dmalloc(&z_out, B * 4 * hw));
  TRY(dmalloc(&ping, pp_elems));
  TRY(dmalloc(&pong, pp_elems));
  return TSD_OK;
}
void Encoder::destroy() {
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  if (graph.exec) cudaGraphExecDestroy(graph.exec);
  float* bufs[] = {img_in, noise_in, z_out, ping, pong};
  for (float* b : bufs)
    if (b) cudaFree(b);
  ps.free_all();
}

// Encoder.forward, vae.mojo:131-159; rescale = the pipeline's rescale((0,255),(-1,1)) (pipeline.mojo:71)
int Encoder::encode(int n, int rescale) {
  int H = latent_h * 8, W = latent_w * 8;
  float* cur = ping;
  float* nxt = pong;
  auto swap = [&]() { float* t = cur; cur = nxt; nxt = t; };
  c->arena.reset();
  LAUNCH(c, launch_rescale_to_nhwc(img_in, cur, n, 3, H * W, rescale, c->stream), "rescale_to_nhwc");
  TRY(conv(c, ps, l1, cur, n, H, W, 3, 128, 3, 1, 1, nullptr, 0, nullptr, nxt, 0));
  swap();
  int ri = 0;
  auto RES = [&]() -> int {
    Act x;
    x.p = cur; x.N = n; x.H = H; x.W = W; x.C = res[ri].cin;
    int rc = res_block(c, ps, res[ri], x, nullptr, 0, 1e-5f, nxt);
    ++ri;
    swap();
    return rc;
  };
  // two_stride_pad (vae.mojo:115-116: one zero row below, one zero column right) + 3x3 stride-2 conv, no padding
  auto DOWN = [&](int wi, int ch) -> int {
    TRY(conv(c, ps, wi, cur, n, H, W, ch, ch, 3, 0, 2, nullptr, 0, nullptr, nxt, 0, nullptr, 1));
    H = conv_out_dim(H, 3, 0, 2, 1);
    W = conv_out_dim(W, 3, 0, 2, 1);
    swap();
    return TSD_OK;
  };
  TRY(RES());  // l2
  TRY(RES());  // l3
  TRY(DOWN(l4, 128));
  TRY(RES());  // l5
  TRY(RES());  // l6
  TRY(DOWN(l7, 256));
  TRY(RES());  // l8
  TRY(RES());  // l9
  TRY(DOWN(l10, 512));
  for (int i = 0; i < 3; ++i) TRY(RES());  // l11..l13
  TRY(vae_attn_block(c, ps, attn_in, attn_out, attn_gn, cur, n, H, W, nxt));  // l14
  swap();
  TRY(RES());  // l15
  {
    // l16 GroupNorm(32,512), l17 SiLU, l18 conv 512->8, l19 conv 1x1 8->8, metrics_evals (vae.mojo:118-129)
    const size_t mark = c->arena.mark();
    WALLOC(f, (long long)n * H * W * 512);
    TRY(op_group_norm(c, cur, f, n, H, W, 512, 32, 1e-5f, ps.gamma(out_gn), ps.beta(out_gn), 1.0f, 1, 0, 1));
    WALLOC(m8, (long long)n * H * W * 8);
    TRY(conv(c, ps, l18, f, n, H, W, 512, 8, 3, 1, 1, nullptr, 0, nullptr, m8, 0));
    TRY(conv(c, ps, l19, m8, n, H, W, 8, 8, 1, 0, 1, nullptr, 0, nullptr, nxt, 0));
    c->arena.release_to(mark);
    LAUNCH(c, launch_latent_from_moments(nxt, noise_in, z_out, n, H * W, c->stream), "latent_from_moments");
  }
  return TSD_OK;
}

size_t Encoder::workspace_bytes(int n) const {
  Encoder* self = const_cast<Encoder*>(this);
  Ctx* cc = self->c;
  const bool was = cc->dry_run;
  cc->dry_run = true;
  cc->arena.set_virtual(true);
  int rc = self->encode(n, 0);
  size_t hw = cc->arena.high_water();
  cc->arena.set_virtual(false);
  cc->dry_run = was;
  return rc ? 0 : hw + (64u << 20);
}
int Encoder::forward(const float* img, const float* noise, int n, int rescale, float* z, bool host_ptrs) {
  if (!ps.loaded) return c->fail(TSD_ERR_STATE, "encoder: forward before load_weights / init_random");
  if (n <= 0 || n > max_batch) return c->fail(TSD_ERR_INVALID, "encoder: batch exceeds max_batch");
  if (!img || !noise || !z) return c->fail(TSD_ERR_INVALID, "encoder: null buffer");
  cudaSetDevice(c->device);
  const size_t need = workspace_bytes(n);
  if (need == 0) return TSD_ERR_OOM;
  if (need > c->arena.capacity()) {
    cudaStreamSynchronize(c->stream);
    if (c->arena.reserve(need) != TSD_OK) return c->fail(TSD_ERR_OOM, "encoder: workspace allocation failed");
  }
  const size_t hw = (size_t)latent_h * latent_w;
  const cudaMemcpyKind in_kind = host_ptrs ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
  TRY(c->check(cudaMemcpyAsync(img_in, img, n * 3 * 64 * hw * sizeof(float), in_kind, c->stream), "copy image"));
  TRY(c->check(cudaMemcpyAsync(noise_in, noise, n * 4 * hw * sizeof(float), in_kind, c->stream), "copy noise"));
  GraphSlot& g = graph;
  const bool use = h->use_graph && !c->timer;
  const bool valid = g.exec && g.n == n && g.n_ctx == rescale && g.epoch == h->option_epoch &&
                     g.arena_base == c->arena.base() && g.weights_gen == ps.gen;
  if (!use) {
    TRY(encode(n, rescale));
  } else if (!valid) {
    TRY(encode(n, rescale));
    TRY(c->check(cudaStreamSynchronize(c->stream), "encoder eager pass"));
    TRY(capture_begin(c));
    int rc = encode(n, rescale);
    int rc2 = capture_end(c, &g);
    if (rc) return rc;
    if (rc2) return rc2;
    g.n = n; g.n_ctx = rescale; g.epoch = h->option_epoch; g.arena_base = c->arena.base();
    g.weights_gen = ps.gen;
  } else {
    TRY(c->check(cudaGraphLaunch(g.exec, c->stream), "graph launch"));
    c->launches += g.nodes;
  }
  TRY(c->check(cudaMemcpyAsync(z, z_out, n * 4 * hw * sizeof(float),
                               host_ptrs ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, c->stream), "copy latent"));
  if (host_ptrs) TRY(c->check(cudaStreamSynchronize(c->stream), "encoder forward sync"));
  return TSD_OK;
}

// ------------------------------------------------------------------------------------------
// denoising loop on the device (pipeline.mojo:86-122)
// ------------------------------------------------------------------------------------------
int generate_latents(Diffusion& m, const tsd_loop_params& lp, const float* latents_in, const float* context,
                     int n_ctx, int n, float* latents_out) {
  Ctx* c = m.c;
  if (!m.ps.loaded) return c->fail(TSD_ERR_STATE, "generate: no weights loaded");
  if (lp.steps <= 0 || !lp.timesteps || !lp.time_emb || !lp.coef || !latents_in || !context || !latents_out)
    return c->fail(TSD_ERR_INVALID, "generate: bad loop parameters");
  const int nb = lp.cfg ? 2 * n : n;
  if (n <= 0 || nb > m.cfg.max_batch) return c->fail(TSD_ERR_INVALID, "generate: batch (x2 with cfg) exceeds max_batch");
  const int groups = lp.cfg ? 2 : 1;
  if (!(n_ctx == groups || n_ctx == groups * n))
    return c->fail(TSD_ERR_INVALID, "generate: n_ctx must be (1|n) without cfg, (2|2n) with cfg (cond rows, then uncond rows)");
  cudaSetDevice(c->device);
  const size_t HW = (size_t)m.cfg.latent_h * m.cfg.latent_w, n_lat = (size_t)n * 4 * HW;
  const size_t ctx_row = (size_t)m.cfg.context_len * m.cfg.context_dim;
  const int steps = lp.steps;

  // workspace: UNet arena + loop-lifetime buffers carved from the top of the same arena
  size_t need = m.workspace_bytes(nb);
  if (need == 0) return TSD_ERR_OOM;
  size_t extra = 0;
  auto carve = [&](size_t floats) { size_t o = extra; extra += (floats * 4 + 1023) / 1024 * 1024; return o; };
  const size_t o_lat = carve(n_lat), o_lat_nchw = carve(n_lat), o_noise = carve(lp.noise ? steps * n_lat : 1),
               o_noise_nchw = carve(lp.noise ? steps * n_lat : 1), o_temb = carve((size_t)steps * 320),
               o_coef = carve((size_t)steps * 5), o_step = carve(256), o_t12 = carve((size_t)steps * 1280 * 2);
  size_t o_tb[9];
  for (int r = 0; r < 9; ++r) o_tb[r] = carve((size_t)steps * m.res[r].cout);
  const size_t total = need + extra + (1u << 20);
  if (total > c->arena.capacity()) {
    cudaStreamSynchronize(c->stream);
    if (c->arena.reserve(total) != TSD_OK) return c->fail(TSD_ERR_OOM, "generate: workspace allocation failed");
  }
  uint8_t* top = (uint8_t*)c->arena.base() + ((need + 1023) / 1024 * 1024);
  auto P = [&](size_t off) { return reinterpret_cast<float*>(top + off); };
  float* lat = P(o_lat);
  float* lat_nchw = P(o_lat_nchw);
  float* noise = lp.noise ? P(o_noise) : nullptr;
  float* temb_all = P(o_temb);
  float* coef = P(o_coef);
  int* step = reinterpret_cast<int*>(P(o_step));
  float* tb_all[9];
  for (int r = 0; r < 9; ++r) tb_all[r] = P(o_tb[r]);

  cudaStream_t s = c->stream;
  // ---- uploads (inside the caller's timed region: this is the end-to-end path) ----
  TRY(c->check(cudaMemcpyAsync(lat_nchw, latents_in, n_lat * 4, cudaMemcpyHostToDevice, s), "H2D latents"));
  LAUNCH(c, launch_nchw_to_nhwc(lat_nchw, lat, n, 4, (int)HW, s), "nchw_to_nhwc");
  if (lp.noise) {
    TRY(c->check(cudaMemcpyAsync(P(o_noise_nchw), lp.noise, (size_t)steps * n_lat * 4, cudaMemcpyHostToDevice, s),
                 "H2D noise"));
    LAUNCH(c, launch_nchw_to_nhwc(P(o_noise_nchw), noise, steps * n, 4, (int)HW, s), "nchw_to_nhwc");
  }
  TRY(c->check(cudaMemcpyAsync(temb_all, lp.time_emb, (size_t)steps * 320 * 4, cudaMemcpyHostToDevice, s), "H2D time"));
  TRY(c->check(cudaMemcpyAsync(coef, lp.coef, (size_t)steps * 5 * 4, cudaMemcpyHostToDevice, s), "H2D coef"));
  TRY(c->check(cudaMemsetAsync(step, 0, sizeof(int), s), "memset"));
  // contexts: one row per UNet batch entry unless a single shared row
  int n_ctx_eff = 1;
  if (n_ctx == 1) {
    TRY(c->check(cudaMemcpyAsync(m.ctx_in, context, ctx_row * 4, cudaMemcpyHostToDevice, s), "H2D context"));
  } else {
    n_ctx_eff = nb;
    for (int i = 0; i < nb; ++i) {
      const int grp = i / n, img = i % n;
      const size_t src_row = (n_ctx == groups) ? (size_t)grp : (size_t)grp * n + img;
      TRY(c->check(cudaMemcpyAsync(m.ctx_in + (size_t)i * ctx_row, context + src_row * ctx_row, ctx_row * 4,
                                   cudaMemcpyHostToDevice, s), "H2D context"));
    }
  }
  c->arena.reset();
  m.ctx_hash_gen = -1;  // kctx / vctx no longer belong to the context the host-buffer entry points cached
  TRY(m.prepare_context(n_ctx_eff));
  // time embedding MLP and the 9 block projections for every step at once (M = steps)
  {
    // prepare_time allocates from the arena when rows > max_batch; those land in the UNet region,
    // which is free until the first step
    TRY(m.prepare_time(steps, temb_all, tb_all, 0));
  }

  StepGather sg;
  for (int r = 0; r < 9; ++r) {
    sg.src[r] = tb_all[r];
    sg.dst[r] = m.tbias[r];
    sg.cout[r] = m.res[r].cout;
  }
  auto one_step = [&]() -> int {
    step_prologue_kernel<<<64, 256, 0, s>>>(step, sg, lat, m.x_nhwc, (long long)n_lat, lp.cfg ? 2 : 1);
    TRY(c->check(cudaGetLastError(), "step_prologue"));
    c->arena.reset();
    TRY(m.unet(nb, n_ctx_eff, 1));
    step_epilogue_kernel<<<64, 256, 0, s>>>(step, coef, m.eps_nhwc, lp.cfg, lp.cfg_scale, noise, lat,
                                            (long long)n_lat);
    TRY(c->check(cudaGetLastError(), "step_epilogue"));
    step_advance_kernel<<<1, 1, 0, s>>>(step);
    TRY(c->check(cudaGetLastError(), "step_advance"));
    c->launches += 3;
    return TSD_OK;
  };

  const bool use_graph = m.h->use_graph && !c->timer;
  if (!use_graph) {
    for (int i = 0; i < steps; ++i) TRY(one_step());
  } else {
    // graph of one step, valid for (nb, ctx rows, cfg, noise?, arena base); replayed `steps` times
    Diffusion::LoopCache& cache = m.loop_cache;
    GraphSlot& g = cache.slot;
    // (the captured step reads step / coef / time-bias tables / noise at addresses carved behind the UNet region:
    // those offsets depend on the step count, so `steps` is part of the key; so is the weight generation)
    const bool valid = g.exec && g.n == nb && g.n_ctx == n_ctx_eff && cache.cfg == lp.cfg &&
                       cache.has_noise == (lp.noise != nullptr) && cache.scale == lp.cfg_scale && cache.steps == steps &&
                       cache.top == (const void*)top && g.epoch == m.h->option_epoch && g.arena_base == c->arena.base() &&
                       g.weights_gen == m.ps.gen;
    int first = 0;
    if (!valid) {
      TRY(one_step());  // eager step 0 (also the warm-up that sets kernel attributes)
      TRY(c->check(cudaStreamSynchronize(s), "generate eager step"));
      first = 1;
      if (steps > 1) {
        TRY(capture_begin(c));
        int rc = one_step();
        int rc2 = capture_end(c, &g);
        if (rc) return rc;
        if (rc2) return rc2;
        cache.cfg = lp.cfg; cache.has_noise = lp.noise != nullptr; cache.scale = lp.cfg_scale;
        cache.top = top;
        cache.steps = steps;
        g.n = nb; g.n_ctx = n_ctx_eff; g.epoch = m.h->option_epoch; g.arena_base = c->arena.base();
        g.weights_gen = m.ps.gen;
      }
    }
    for (int i = first; i < steps; ++i) {
      TRY(c->check(cudaGraphLaunch(g.exec, s), "graph launch"));
      c->launches += g.nodes;
    }
  }
  LAUNCH(c, launch_nhwc_to_nchw(lat, lat_nchw, n, 4, (int)HW, s), "nhwc_to_nchw");
  TRY(c->check(cudaMemcpyAsync(latents_out, lat_nchw, n_lat * 4, cudaMemcpyDeviceToHost, s), "D2H latents"));
  TRY(c->check(cudaStreamSynchronize(s), "generate sync"));
  return TSD_OK;
}

}  // namespace tsd
