// Device runtime under the C ABI: context (device, stream, workspace arena, TMA descriptor
// encoder), and the op layer the model graphs are composed from.  All pointers handled here
// are device pointers; activations are fp32 NHWC.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/tsd_b200.h"  // status codes
#include "norm_stats.cuh"

namespace tsd {

struct Status {
  int code = 0;  // 0 = OK (C ABI status)
  std::string msg;
  bool ok() const { return code == 0; }
};


// Bump allocator over one cudaMalloc'd block; reset() at the top of every forward so the same
// call sequence yields the same addresses (CUDA-graph friendly, no allocation in the hot loop).
class Arena {
 public:
  ~Arena();
  int reserve(size_t bytes);
  void reset() { off_ = 0; }
  void* alloc(size_t bytes);  // 1024 B aligned; nullptr on exhaustion
  template <class T>
  T* alloc_n(size_t n) { return static_cast<T*>(alloc(n * sizeof(T))); }
  size_t capacity() const { return cap_; }
  size_t used() const { return off_; }
  size_t high_water() const { return high_; }
  size_t mark() const { return off_; }
  void release_to(size_t m) { off_ = m; }
  // planning mode: allocations succeed against an unbounded fake address range and only the
  // high-water mark is recorded; nothing may dereference the returned pointers
  void set_virtual(bool v) { virtual_ = v; off_ = 0; high_ = 0; }
  bool is_virtual() const { return virtual_; }
  const void* base() const { return base_; }
  int generation() const { return gen_; }

 private:
  uint8_t* base_ = nullptr;
  size_t cap_ = 0, off_ = 0, high_ = 0;
  bool virtual_ = false;
  int gen_ = 0;
};

using PFN_encodeTiled = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                     const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                     const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct KernelTimer;  // optional per-launch CUDA-event timing (bench roofline leg)
struct TuneCache;    // per-context GEMM tile autotuning results (runtime.cu)
constexpr int kTileTickets = 1 << 16;
constexpr size_t kFlushBytes = 192u << 20;  // > L2: written before every autotune timing run

// Request to the producer of an activation: "the next consumer normalises this tensor with G groups
// and this eps".  If the producing kernel can fold partial statistics into its epilogue
// (norm_stats.cuh) it fills `req` (req.partial != nullptr) and the consumer runs a single
// normalise pass; otherwise the consumer computes the statistics itself.  `scratch` is owned by
// the caller and must stay valid until the consumer has been enqueued.
struct NormHint {
  int G = 0;
  float eps = 0.f;
  int imgs = 1;  // plain GEMM outputs: rows = imgs * rows_per_image
  float2* scratch = nullptr;
  size_t scratch_elems = 0;
  NormStatsReq req;
  const NormStatsReq* ready() const { return req.partial ? &req : nullptr; }
  // Deferred split-K reduction: the consumer promises that the very next reader of the producer's output is its
  // norm.  A split-K producer then leaves its partial tiles in the arena (not released) and launches no reduce
  // kernel; the norm kernel sums them, adds bias / residual, writes the raw output and normalises in one launch.
  bool allow_defer = false;
  // Partial-tile buffer that outlives the producer's own arena scope (consumer in another block); nullptr: the
  // producer allocates from the arena and leaves it to the caller's mark.  Capacity in floats.
  float* defer_ws = nullptr;
  size_t defer_ws_elems = 0;
  struct Deferred {
    const float* ws = nullptr;  // [splits][rows][ld]
    int splits = 0, ld = 0;
    long long split_stride = 0;
    const float* bias = nullptr;
    int bias_img_stride = 0;
    const float* residual = nullptr;
    float* raw = nullptr;       // [rows][C] un-normalised output
    int rows = 0, C = 0;
    // virtual channel concat (splits == 1): ws = first tensor [rows][c_a], x2 = second [rows][C - c_a]; the norm
    // kernel writes the concatenated tensor to `raw` as a by-product
    const float* x2 = nullptr;
    int c_a = 0;
  } def;
  const Deferred* deferred() const { return def.ws ? &def : nullptr; }
};
// upper bound of the partial entries a producer may write for an [imgs][rows_per_img][C] activation
inline size_t norm_scratch_elems(int imgs, long long rows_per_img, int C, int G) {
  const long long slabs = (rows_per_img + 63) / 64 + 64;  // conv tiles are pixel boxes: allow ragged edges
  return (size_t)imgs * (size_t)slabs * (size_t)(G + C / 8 + 2);
}


struct Ctx {
  int device = -1;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  PFN_encodeTiled encode = nullptr;
  Arena arena;
  std::string last_error;
  int sm_count = 148;
  int softmax_axis = 0;     // 0 = query axis (reference Softmax(dim=2), Q3) ; 1 = key axis
  int fused_attention = 1;  // 1 = tcgen05 fused kernel ; 0 = GEMM + softmax + GEMM
  int attn_poly = 0;        // lab switch, compiled out (FMA-pipe polynomial for every fourth exponential: measured no gain)
  int attn_v2 = 1;          // fused attention: 1 = software-pipelined softmax role (attn2_kernel) where the shape allows, 0 = attn_kernel
  int layernorm_mode = 0;   // 0 = global statistics (reference, Q5) ; 1 = per token
  int norm_eps_mode = 0;    // 0 = (x - mean) / (std + eps) (reference, Q6) ; 1 = (x - mean) / sqrt(var + eps)
  int force_bn = 0, force_splits = 0;  // test/tuning overrides for the GEMM tile heuristic
  int force_stages = 0;     // tuning: operand ring depth override
  int gemm_cg = 0;          // tuning: 0 = model decides, 1 = single CTAs, 2 = CTA pairs (cta_group::2)
  int gemm_debug = 0;       // lab only: bit0 skip MMAs, bit1 skip A loads, bit2 skip B loads (results invalid)
  unsigned int* ticket = nullptr;  // zero-initialised device counters for last-block reductions
  int bench_stats_groups = 0;      // lab: tsd_bench_conv / tsd_bench_gemm request norm statistics with this many groups
  unsigned int* tile_tickets = nullptr;  // zero-initialised per-tile arrival counters of the split-K fix-up
  int splitk_fixup = 0;            // 1: split-K partials reduced by the last CTA of each tile (measured slower than the separate reduce kernel: serial tail)
  int splitk_cluster = 0;          // 1: split-K reduced inside the GEMM by the thread-block cluster of a tile's K splits (cluster size <= splitk_cluster_max); measured: the in-kernel reduction costs about what the reduce kernel does, so it stays off
  int splitk_cluster_max = 8;      // largest cluster (CTA pair x splits) the in-kernel reduction is used with; 16 needs the non-portable opt-in
  unsigned int* norm_bar = nullptr;  // zero-initialised barrier words of the fused norm (elementwise.cuh)
  TuneCache* tune = nullptr;
  void* flush_buf = nullptr;
  int autotune = 1;                // time tile candidates on first use of a GEMM signature (0: cost model only)
  int conv_halo = 0;               // 3x3 convolutions: 0 never (default: measured on par with the implicit GEMM), 1 autotuner may choose, 2 whenever eligible
  int halo_min_w = 16, halo_min_h = 18;  // lab: smallest image the halo TMA box is used on
  int tune_verbose = 0;
  int tune_flush = 0;              // 1: flush L2 before every autotune timing run (weights AND activations cold)
  int gn_partial = 1;              // GroupNorm consumes producer-side partial statistics (0: always the stand-alone fused norm)
  int gn_partial_max_groups = 64;  // ... only up to this many groups (every block folds all groups of its image)
  int ln_fold = 1;                 // fold global-statistics LayerNorm into the consuming GEMM epilogue (0: separate pass)
  int defer_reduce = 1;            // split-K partials summed by the consuming norm kernel where the call site allows it
  int virtual_concat = 1;          // 1: channel concats feeding a ResBlock are read (and written out) by its first GroupNorm kernel (+0.5 % with the cluster norm)
  int conv_stride_tma = 1;         // stride-2 3x3 convolutions as implicit GEMMs through a strided tensor map (0: im2col + GEMM)
  int fuse_ffn_out = 1;            // the two linears that end an attention block (geglu2, conv_out) as one GEMM with merged weights
  int fuse_skip = 1;               // ResBlock 1x1 skip convolution as a second K segment of conv2 (0: GEMM of its own + residual add)
  int norm_cluster = 1;            // 1: GroupNorm with one thread-block cluster per (image, group) where the slab fits shared memory
  int tune_defer_penalty_us = 3;   // autotuner: cost charged to a split-K candidate whose GroupNorm consumer is the cluster kernel (it sums the partials)
  int gemm_kmerge = 1;             // 1: one TMA request per operand and K step of 64 where the shapes allow (A/B switch)
  int gemm_deep_b = 0;             // 1: narrow GEMM tiles stream their weights through a separate, deeper ring (A/B switch)
  int norm_v2 = 0;                 // 0: previous fused norm kernel (A/B switch)
  int producer_stats = 1;          // 0: never fold norm statistics into GEMM epilogues (A/B switch)
  KernelTimer* timer = nullptr;
  long long launches = 0;   // kernels launched through this context (bench "gpu_launches")
  bool dry_run = false;     // planning pass: ops allocate workspace but launch nothing

  int fail(int code, const std::string& msg);
  int check(cudaError_t e, const char* what);
};

int ctx_create(int device, Ctx** out, std::string* err);
void ctx_destroy(Ctx* ctx);

// ---- per-launch timing (used by bench.py to attribute device time to kernel families) ----
struct KernelTimer {
  struct Rec {
    int family;
    double flops;
    cudaEvent_t a, b;
  };
  std::vector<Rec> recs;
  std::vector<cudaEvent_t> pool;
  size_t next = 0;
  cudaEvent_t get();
};
enum KernelFamily { FAM_GEMM = 0, FAM_ATTN = 1, FAM_NORM = 2, FAM_OTHER = 3, FAM_COUNT = 4 };
struct TimedScope {
  Ctx* c;
  int idx = -1;
  TimedScope(Ctx* ctx, int family, double flops);
  ~TimedScope();
};

// ---------------------------------------------------------------------------------------
// op layer
// ---------------------------------------------------------------------------------------
struct GemmArgs {
  // A: [batch][M][K] row-major, row stride lda (elements), batch stride a_bs
  const float* A = nullptr;
  int M = 0, K = 0;
  long long lda = 0, a_bs = 0;
  // B: [batch][N][K] row-major (K-major "weight" layout), row stride ldb
  const float* B = nullptr;
  int N = 0;
  long long ldb = 0, b_bs = 0;
  int batch = 1;
  // D: [batch][M][ldd]
  float* D = nullptr;
  long long ldd = 0, d_bs = 0;
  const float* bias = nullptr;      // [N]
  const float* row_bias = nullptr;  // [M]
  const float* residual = nullptr;  // [batch][M][ldr]
  long long ldr = 0, r_bs = 0;
  float alpha = 1.0f;
  int geglu = 0;                    // N counts both halves; D gets N/2 columns
  int split_n = 0;                  // >0: column n goes to D + (n/split_n)*split_stride + n%split_n
  long long split_stride = 0;
  int round_tf32 = 0;
  int force_bn = 0, force_splits = 0;  // tuning / tests
  NormHint* nh = nullptr;              // optional: statistics of D for the next norm
  int b_static = 0;                    // B is a weight matrix no kernel writes during the forward pass
  // LayerNorm (global statistics) of A folded into the epilogue: statistics of A left by ITS producer,
  // and the row sums of B; the GEMM then reads the un-normalised A
  const NormStatsReq* ln_fold = nullptr;
  const float* wsum = nullptr;
};
int op_gemm(Ctx* c, const GemmArgs& a);

struct ConvArgs {
  const float* x = nullptr;  // [N][H][W][Cin]
  int N = 1, H = 0, W = 0, Cin = 0, Cout = 0;
  int k = 3, pad = 1, stride = 1;
  int pad_hi = -1;  // bottom/right padding; -1 = same as pad (top/left).  Encoder: pad 0, pad_hi 1 (vae.mojo:115-116)
  const float* w = nullptr;     // [Cout][k*k*Cin]  (O,(kh,kw),I)
  const float* bias = nullptr;  // [Cout] (image i uses bias + i * bias_img_stride)
  int bias_img_stride = 0;
  const float* residual = nullptr;  // [N][Ho][Wo][Cout]
  float* out = nullptr;             // [N][Ho][Wo][Cout]
  int round_tf32 = 0;
  int force_bn = 0, force_splits = 0;
  NormHint* nh = nullptr;  // optional: statistics of `out` for the next norm
  // optional fused 1x1 convolution of a second tensor accumulated into the same output (3x3 / stride 1 / pad 1
  // only): x2 [N][H][W][Cin2]; `w` then is [Cout][k*k*Cin + Cin2] and `bias` the sum of both biases
  const float* x2 = nullptr;
  int Cin2 = 0;
};
int op_conv2d(Ctx* c, const ConvArgs& a);
inline int conv_out_dim(int in, int k, int pad, int stride, int pad_hi = -1) {
  return (in + pad + (pad_hi < 0 ? pad : pad_hi) - k) / stride + 1;
}

// GroupNorm(+SiLU)(+2x nearest upsample). stats scratch comes from the arena.
int op_group_norm(Ctx* c, const float* x, float* y, int N, int H, int W, int C, int G, float eps,
                  const float* gamma, const float* beta, float gamma_scalar, int silu, int upsample,
                  int round_tf32, const NormStatsReq* pre = nullptr, const NormHint::Deferred* def = nullptr);

struct AttnArgs {
  // per (batch b, head h): Q [Tq][d], K [Tk][d], V [Tk][d], contiguous blocks
  // (reference raw-reshape head split, attention.mojo:29-44: head h = flat block h of the
  // [T][C] projection).  Batch stride = heads * T * d.
  const float* Q = nullptr;
  const float* K = nullptr;
  const float* V = nullptr;
  int batch = 1, heads = 1, Tq = 0, Tk = 0, d = 0;
  float* O = nullptr;  // merged [batch][Tq][heads*d]: O[t][h*d + j]  (attention.mojo:61-62)
  int softmax_axis = 0;
  int causal = 0;
  long long kv_batch_stride = -1;  // elements between batch entries of K/V; -1 = heads*Tk*d, 0 = shared
};
int op_attention(Ctx* c, const AttnArgs& a);
int op_attention_unfused(Ctx* c, const AttnArgs& a);

}  // namespace tsd
