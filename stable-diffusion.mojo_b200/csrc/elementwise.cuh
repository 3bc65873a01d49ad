// K5..K9: HBM-bound kernels around the tensor-core GEMMs: GroupNorm/LayerNorm statistics
// and apply(+SiLU)(+nearest 2x upsample), layout changes (reference CHW <-> device NHWC),
// channel concat, stride-2 im2col, direct convolutions for degenerate channel counts,
// M=1 GEMV (time embedding), softmax for the unfused attention path, the DDPM sampler step.
// All activations are fp32, NHWC (= token-major [pixels][channels]).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

#include "norm_stats.cuh"

namespace tsd {

// ---- layout -----------------------------------------------------------------------------
cudaError_t launch_nchw_to_nhwc(const float* src, float* dst, int N, int C, int HW, cudaStream_t s);
cudaError_t launch_nhwc_to_nchw(const float* src, float* dst, int N, int C, int HW, cudaStream_t s);
// per batch entry: src [R][Cc] -> dst [Cc][ld_out] (ld_out >= R; padding columns untouched)
cudaError_t launch_transpose_ld(const float* src, float* dst, int B, int R, int Cc, int ld_out,
                                cudaStream_t s);
// conv weights: reference OIHW (Matrix_Array [out][in][k][k], helpers/utils.mojo:1718) ->
// K-major [O][kh*kw][I] rows for the implicit-GEMM B operand
cudaError_t launch_oihw_to_ohwi(const float* src, float* dst, int O, int I, int KK, cudaStream_t s);
cudaError_t launch_concat_channels(const float* a, int Ca, const float* b, int Cb, float* out,
                                   long long pixels, cudaStream_t s);
// copies the first `Cout` channels of a [pixels][Cin] tensor (Q9: truncated concat)
cudaError_t launch_upsample2x(const float* x, float* y, int N, int H, int W, int C, cudaStream_t s);
// channel-major (C,H,W) -> (C,2H,2W), any C
cudaError_t launch_upsample2x_planar(const float* x, float* y, int C, int H, int W, cudaStream_t s);
cudaError_t launch_rescale_to_nhwc(const float* src, float* dst, int N, int C, int HW, int rescale,
                                   cudaStream_t s);  // NCHW (0..255 when rescale) -> NHWC (-1..1)
cudaError_t launch_latent_from_moments(const float* m, const float* noise, float* out, int N, int HW,
                                       cudaStream_t s);
cudaError_t launch_im2col3x3(const float* x, float* col, int N, int H, int W, int C, int stride,
                             int pad_lo, int Ho, int Wo, cudaStream_t s);

// ---- normalisation ------------------------------------------------------------------------
// stats[n][g] = (mean, 1/(std+eps)) with the reference's biased std and eps added to std
// (helpers/utils.mojo:1360-1380, 1868-1870).  `scratch` holds group_stats_scratch_bytes() bytes;
// `ticket` is a zero-initialised device counter owned by the context (self-resetting).
// Deterministic: no floating-point atomics, fixed reduction order.
size_t group_stats_scratch_bytes(int N, long long pixels, int C, int G);
cudaError_t launch_group_stats(const float* x, int N, long long pixels, int C, int G, float eps,
                               void* scratch, unsigned int* ticket, float2* stats, cudaStream_t s);
// y = (x - mean) * inv [* gamma[c] + beta[c]] ; optional SiLU ; optional TF32 rounding ;
// optional nearest 2x upsample on write (x is [N,H,W,C], y is [N,2H,2W,C]).
// fused norm v2 (one launch, activation read once; optionally also the split-K reduction of its producer)
constexpr int kNormBarrierCounters = 512;  // 128 B apart: root/generation line + sub-group counters
struct NormFused2Src {
  const float* x = nullptr;       // activation [N][pixels][C], or split-K partials [splits][N*pixels][ldx]
  int splits = 1;
  long long split_stride = 0;     // floats between splits
  int ldx = 0;                    // row stride of the partials (floats)
  const float* bias = nullptr;    // split-K source: per-image bias rows
  int bias_img_stride = 0;
  const float* residual = nullptr;  // split-K source: [N*pixels][C]
  float* raw = nullptr;             // split-K source: un-normalised result [N*pixels][C] (nullptr: not needed)
  const float* x2 = nullptr;        // two-source (channel concat [x | x2]): x is [rows][c_a], x2 [rows][C - c_a];
  int c_a = 0;                      //   raw (optional) receives the concatenated tensor
};
bool norm_fused2_supported(int N, long long pixels, int C, int G, int sm_count);
size_t norm_fused2_scratch_bytes(int N, long long pixels, int C, int G, int sm_count);
cudaError_t launch_norm_fused2(const NormFused2Src& src, float* y, int N, long long pixels, int C, int G, float eps,
                               const float* gamma, const float* beta, float gamma_scalar, int silu, int round_tf32,
                               void* scratch, unsigned int* barrier_words, int sm_count, cudaStream_t s);
// cluster GroupNorm: one thread-block cluster per (image, group), slab in shared memory, no grid barrier, no scratch
bool norm_cluster_supported(int N, long long pixels, int C, int G);
cudaError_t launch_norm_cluster(const NormFused2Src& src, float* y, int N, long long pixels, int C, int G, float eps,
                                const float* gamma, const float* beta, float gamma_scalar, int silu, int round_tf32,
                                cudaStream_t s);
bool norm_apply_partial_supported(int C, int G);
cudaError_t launch_norm_apply_partial(const float* x, float* y, const NormStatsReq& req, int N, long long pixels,
                                      const float* gamma, const float* beta, float gamma_scalar, int silu,
                                      int round_tf32, cudaStream_t s);
cudaError_t launch_norm_apply(const float* x, const float2* stats, const float* gamma,
                              const float* beta, float gamma_scalar, float* y, int N, int H, int W,
                              int C, int G, int silu, int upsample2x, int round_tf32,
                              cudaStream_t s);

// One-launch GroupNorm/LayerNorm (+SiLU, +TF32 rounding): statistics, grid barrier, normalise.
// Needs C % 4 == 0 and C <= 3072; grid <= sm_count blocks (all resident).  `barrier_words` = two
// zero-initialised device words owned by the context.
bool norm_fused_supported(int C);
size_t norm_fused_scratch_bytes(int N, long long pixels, int C, int G);
cudaError_t launch_norm_fused(const float* x, float* y, int N, long long pixels, int C, int G, float eps,
                              const float* gamma, const float* beta, float gamma_scalar, int silu, int round_tf32,
                              void* scratch, unsigned int* barrier_words, int sm_count, cudaStream_t s);

// deterministic U(lo,hi) fill (synthetic weights / probes): counter-based splitmix64 hash
cudaError_t launch_fill_uniform(float* p, long long n, uint64_t seed, float lo, float hi, cudaStream_t s);

// ---- simple elementwise --------------------------------------------------------------------
enum UnaryOp { UNARY_SILU = 0, UNARY_GELU = 1, UNARY_SCALE = 2, UNARY_COPY = 3, UNARY_QUICKGELU = 4 };
cudaError_t launch_unary(const float* x, float* y, long long n, int op, float scalar, cudaStream_t s);
cudaError_t launch_add(const float* a, const float* b, float* y, long long n, cudaStream_t s);
// y[p][c] = x[p][c] + v[c]
cudaError_t launch_add_channel_vec(const float* x, const float* v, float* y, long long pixels, int C,
                                   cudaStream_t s);

// ---- direct convolution (CUDA cores) for shapes the tensor-core path does not take ---------
// x [N,H,W,Cin] ; w [Cout][k*k][Cin] ; out [N,Ho,Wo,Cout]
cudaError_t launch_conv_direct(const float* x, const float* w, const float* bias, float* out, int N,
                               int H, int W, int Cin, int Cout, int k, int pad, int stride, int Ho,
                               int Wo, cudaStream_t s);

// ---- GEMV: y[r][n] = sum_k act(x[r][k]) * Wt[n][k] + bias[n] + bias2[n] ---------------------
struct GemvSeg {
  const float* W;      // [N][K]
  const float* bias;   // [N] or nullptr
  const float* bias2;  // [N] or nullptr
  float* y;            // [rows][N]
  int N, row0;
};
struct GemvMulti {
  static constexpr int kMax = 12;
  GemvSeg seg[kMax];
  int nseg, total;
};
cudaError_t launch_gemv_multi(const float* x, int rows, int K, const GemvMulti& m, int silu_in, cudaStream_t s);
cudaError_t launch_gemv(const float* x, int rows, int K, const float* Wt, const float* bias,
                        const float* bias2, float* y, int N, int silu_in, int silu_out,
                        cudaStream_t s);

// ---- softmax over a [B][R][Cc] score tensor (row stride ld) -------------------------------
// axis 0: normalise each column over rows (reference Softmax(dim=2) behaviour, Q3)
// axis 1: normalise each row over columns (standard attention)
cudaError_t launch_softmax(float* S, int B, int R, int Cc, int ld, int axis, float scale,
                           float* col_scratch, cudaStream_t s, int causal = 0);

// keeps the stream busy for `ns` nanoseconds (profiling aid)
cudaError_t launch_clip_embed(const int* tokens, const float* table, const float* pos, float* out, int T, int d,
                              int n_vocab, cudaStream_t s);
cudaError_t launch_spin(long long ns, cudaStream_t s);

// ---- DDPM step (+ optional CFG combine), sampler.mojo:75-109, pipeline.mojo:117-119 --------
// eps = cfg ? u + cfg_scale * (c - u) : c ; x0 = (x - sqrt(1-ab_t) eps)/sqrt(ab_t)
// out = c0 * x0 + c1 * x + sigma * noise
cudaError_t launch_ddpm_step(const float* x, const float* eps_c, const float* eps_u, float cfg_scale,
                             const float* noise, float sqrt_ab, float sqrt_1mab, float c0, float c1,
                             float sigma, float* out, long long n, cudaStream_t s);

// img = clamp((v + 1) * 127.5, 0, 255), NHWC -> NCHW  (pipeline.mojo:127, utils.mojo:577-597)
cudaError_t launch_rescale_to_nchw(const float* src, float* dst, int N, int C, int HW, int rescale,
                                   cudaStream_t s);

}  // namespace tsd
