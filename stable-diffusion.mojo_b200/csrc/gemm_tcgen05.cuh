// K1/K2: TF32 GEMM / implicit-GEMM convolution on tcgen05 tensor cores.
//
//   D[b][m][n] = epilogue( sum_{tap,k} A[b][pixel(m) + shift(tap)][k] * B[b][n][tap*Cin + k] )
//
// * A is a 4-D fp32 tensor (K innermost, then W, H, batch) - an NHWC activation, or a plain
//   [M,K] matrix viewed as W=M,H=1.  One CTA owns a 128-row tile = a (bh x bw) pixel box; the
//   TMA producer issues one box load per (tap, 32-channel chunk) with the tap's (dy,dx) shift
//   added to the box coordinates; out-of-bounds elements are zero-filled by TMA, which *is*
//   the convolution's zero padding (reference Conv2D.forward pads first, helpers/utils.mojo:1749).
// * B is a K-major [N][taps*Cin] weight matrix (reference OIHW re-laid as O,(kh,kw),I).
// * Both operands land in shared memory in the canonical K-major SWIZZLE_128B layout
//   (rows of 32 fp32 = 128 B) and feed tcgen05.mma.kind::tf32 (M=128, N=BN, K=8 per instruction)
//   issued by one thread; the accumulator lives in TMEM and is read back by 4 epilogue warps.
// * Epilogue options: column bias, row bias, residual add, alpha scale, GEGLU
//   (out * gelu(gate), diffusion.mojo:138-141), column-block routing (QKV split), TF32 rounding,
//   or raw split-K partials (reduced by splitk_reduce_kernel, which then applies bias/residual).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>

#include "norm_stats.cuh"

namespace tsd {

constexpr int GEMM_BM = 128;       // rows per CTA tile (TMEM lanes)
constexpr int GEMM_BK = 64;        // preferred K step (fp32 elements) = two 128 B swizzle atoms per operand row; 32 when smem-bound
// Role pairs (TMA producer + MMA issuer threads) per CTA: 2 spreads the K steps over two thread pairs
// with separate accumulators.  Measured on B200: no gain (the K loop is bound by shared-memory
// bandwidth, not by instruction issue) and the 12-warp CTA caps registers at 168 (epilogue spills),
// so 1 is used; the two-pair path stays for wider tiles / other precisions.
constexpr int GEMM_ROLE_PAIRS = 1;
// With one role pair, warp 10 is a second TMA producer: warp 0 issues the A (activation) boxes and arms the barrier,
// warp 10 issues the B (weight) boxes.  A lone thread sustains one TMA request per 180-420 clk (tools/lab/l2bench.cu),
// and a K step of 64 is four requests: the under-filled small-M GEMMs of the UNet were bound by that chain.
constexpr int GEMM_B_PRODUCER = GEMM_ROLE_PAIRS == 1 ? 1 : 0;
constexpr int GEMM_THREADS = 320 + 64 * (GEMM_ROLE_PAIRS - 1) + 32 * GEMM_B_PRODUCER;  // warps 0/10 TMA producers, 1/11 MMA issuers (1: TMEM alloc), 2..9 epilogue
constexpr int GEMM_MAX_STAGES = 6;
constexpr int GEMM_MAX_B_STAGES = 24;
constexpr int GEMM_DSM_TILE_OFFSET = 48 * 1024;  // cluster split-K (fixup 3): the raw partial tile sits behind the epilogue staging area  // deep weight ring (GemmKParams::b_stages)

struct GemmKParams {
  // output-pixel tiling (plain GEMM: H = 1, W = M, bw = 128, bh = 1)
  int H, W, bw, bh, tiles_w, tiles_h;
  int m_per_batch;  // img_n * H * W : rows of D per batch entry
  // K loop
  int taps, cin, chunks_per_tap, total_iters, splits, iters_per_split;
  int cstride, coff;       // strided 3x3 convolution: input coordinate = out * cstride + tap offset (-1..1) + coff
  int cin2;                // > 0: second K segment - cin2 channels of a second NHWC operand (tmA2) after the taps
  int a_box_bytes;
  // N tiling
  int BN, n_valid, geglu, n_half;
  int num_stages, tmem_cols;
  int bk;                  // K step of this launch: 64 or 32
  int halo;                // 1: conv3x3_halo_kernel (total_iters / iters_per_split count 32-channel chunks; num_stages = B ring depth)
  int acc_stride;          // TMEM columns between the two issuers' accumulator tiles
  // output
  float* D;
  long long d_batch_stride;
  int ldd;
  int split_n;             // column routing: col n -> (n / split_n) * split_stride + n % split_n
  long long split_stride;
  const float* bias;       // [n] or nullptr
  int bias_img_stride;     // conv: bias row of image i starts at bias + i * bias_img_stride
  const float* row_bias;   // [m] or nullptr
  const float* residual;   // [b][m][ldr] or nullptr
  long long r_batch_stride;
  int ldr;
  float alpha;
  int round_tf32;
  float* partial;          // split-K workspace [split][b][m][n_pad] or nullptr
  int n_pad;
  int fixup;               // split-K: the last CTA of every output tile reduces the partials and runs the epilogue
  unsigned int* tile_tickets;  // [grid.x * grid.y] zero-initialised, self-resetting arrival counters (fixup)
  int b_stages;            // > 0: the B (weight) operand has its own ring of this depth behind the A ring (num_stages deep), with its
                           //      own barriers: narrow tiles stream many small weight boxes from HBM, and a ring sized for the
                           //      32 KiB A boxes keeps too few of them in flight to cover the DRAM latency
  int kmerge;              // 1: a K step of 64 is ONE TMA request per operand (tensor maps carry the 32-channel chunk index as
                           //    an extra outermost dimension, box extent 2): half the requests of the single-thread producers
  int debug;               // lab only (Ctx::gemm_debug)
  int cg;                  // 1, or 2 = CTA pairs over consecutive M tiles (cluster 2x1x1, grid.x even)
  NormStatsReq ns;         // producer-side GroupNorm statistics of D (ns.partial == nullptr: off)
  // LayerNorm (global statistics, one group per image) of the A operand folded into the epilogue:
  //   W.((x - mu) r) + b  =  r (W.x) + (b - r mu rowsum(W));  ln.partial == nullptr: off
  NormStatsReq ln;
  const float* wsum;       // [N] row sums of the (TF32-rounded) weight matrix
  int ln_rows_per_img;
  int b_static;            // B holds weights (never written by a predecessor kernel): may be loaded before pdl_wait
  int imgs;                // images (rows of tiles past the last image are phantom: loaded as zeros, never stored)
};

struct SplitKReduceParams {
  const float* partial;
  int splits;
  long long split_stride;  // elements between consecutive splits
  int m, n_pad, n_valid;
  float* D;
  int ldd;
  const float* bias;
  int bias_img_stride, rows_per_img;
  const float* residual;
  int ldr;
  int round_tf32;
  NormStatsReq ns;  // optional statistics of D (slab variant below when ns.partial != nullptr)
  int slab_rows;    // rows per block of the slab variant
};

cudaError_t launch_gemm_tf32(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmKParams& p,
                             dim3 grid, size_t smem_bytes, cudaStream_t stream, const CUtensorMap* tmA2 = nullptr);
cudaError_t launch_splitk_reduce(const SplitKReduceParams& p, cudaStream_t stream);
int halo_pick_sb(int BN, int cg);
size_t halo_smem_bytes(int BN, int cg, int sb);
cudaError_t launch_conv_halo(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmKParams& p, dim3 grid,
                             size_t smem_bytes, cudaStream_t stream);
size_t gemm_smem_bytes(int BN, int num_stages, int cg, int bk, int b_stages = 0, int dsm_tile = 0);
void gemm_pick_ring(int BN, int cg, int* bk, int* stages);

}  // namespace tsd
