// K5..K9 kernel bodies - see elementwise.cuh.
#include "elementwise.cuh"

#include "pdl.cuh"

#include <cstdio>
#include <cstdlib>

#include <cfloat>

namespace tsd {

namespace {

constexpr int kSMs = 148;

inline int grid_for(long long work_items, int threads, int max_waves = 8) {
  long long b = (work_items + threads - 1) / threads;
  long long cap = (long long)kSMs * max_waves;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

__device__ __forceinline__ float silu_f(float x) { return __fdividef(x, 1.0f + __expf(-x)); }
__device__ __forceinline__ float gelu_f(float x) {
  const float k = 0.7978845608028654f;
  return 0.5f * x * (1.0f + tanhf(k * (x + 0.044715f * x * x * x)));
}
__device__ __forceinline__ float rna_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;\n" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// ------------------------------------------------------------------------------------------
// layout: per image [R][Cc] -> [Cc][R] transpose through a padded smem tile
// ------------------------------------------------------------------------------------------
__global__ void transpose_kernel(const float* __restrict__ src, float* __restrict__ dst, int R,
                                 int Cc, int rescale) {
  pdl_wait();
  pdl_launch_dependents();
  __shared__ float tile[32][33];
  const long long img = (long long)blockIdx.z * R * Cc;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int r = r0 + i, c = c0 + threadIdx.x;
    if (r < R && c < Cc) tile[i][threadIdx.x] = src[img + (long long)r * Cc + c];
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int c = c0 + i, r = r0 + threadIdx.x;
    if (r < R && c < Cc) {
      float v = tile[threadIdx.x][i];
      if (rescale == 1) v = fminf(fmaxf((v + 1.0f) * 127.5f, 0.0f), 255.0f);
      else if (rescale == 2) v = v * 2.0f / 255.0f - 1.0f;  // rescale((0,255),(-1,1)), pipeline.mojo:71
      dst[img + (long long)c * R + r] = v;
    }
  }
}

cudaError_t launch_transpose(const float* src, float* dst, int N, int R, int Cc, int rescale,
                             cudaStream_t s) {
  dim3 grid((Cc + 31) / 32, (R + 31) / 32, N), block(32, 8);
  { cudaError_t e_ = launch_pdl(transpose_kernel, dim3(grid), dim3(block), 0, s, src, dst, R, Cc, rescale); if (e_ != cudaSuccess) return e_; }
  return cudaGetLastError();
}

__global__ void transpose_ld_kernel(const float* __restrict__ src, float* __restrict__ dst, int R,
                                    int Cc, int ld_out) {
  __shared__ float tile[32][33];
  const float* sb = src + (long long)blockIdx.z * R * Cc;
  float* db = dst + (long long)blockIdx.z * Cc * ld_out;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int r = r0 + i, c = c0 + threadIdx.x;
    if (r < R && c < Cc) tile[i][threadIdx.x] = sb[(long long)r * Cc + c];
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int c = c0 + i, r = r0 + threadIdx.x;
    if (r < R && c < Cc) db[(long long)c * ld_out + r] = tile[threadIdx.x][i];
  }
}

__global__ void oihw_to_ohwi_kernel(const float* __restrict__ src, float* __restrict__ dst, int O,
                                    int I, int KK) {
  const long long total = (long long)O * I * KK;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int ci = (int)(i % I);
    long long t = i / I;
    int tap = (int)(t % KK);
    long long o = t / KK;
    dst[i] = src[(o * I + ci) * KK + tap];
  }
}

__global__ void concat_kernel(const float* __restrict__ a, int Ca, const float* __restrict__ b,
                              int Cb, float* __restrict__ out, long long pixels) {
  pdl_wait();
  pdl_launch_dependents();
  const int C = Ca + Cb;
  const long long total = pixels * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    long long p = i / C;
    int c = (int)(i - p * C);
    out[i] = c < Ca ? a[p * Ca + c] : b[p * Cb + (c - Ca)];
  }
}

__global__ void upsample2x_kernel(const float4* __restrict__ x, float4* __restrict__ y, int N, int H,
                                  int W, int C4) {
  pdl_wait();
  pdl_launch_dependents();
  const long long total = (long long)N * 4 * H * W * C4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % C4);
    long long t = i / C4;
    int wo = (int)(t % (2 * W));
    t /= (2 * W);
    int ho = (int)(t % (2 * H));
    int n = (int)(t / (2 * H));
    y[i] = x[(((long long)n * H + (ho >> 1)) * W + (wo >> 1)) * C4 + c];
  }
}

__global__ void upsample2x_planar_kernel(const float* __restrict__ x, float* __restrict__ y, int C,
                                         int H, int W) {
  const long long total = (long long)C * 4 * H * W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int wo = (int)(i % (2 * W));
    long long t = i / (2 * W);
    int ho = (int)(t % (2 * H));
    long long c = t / (2 * H);
    y[i] = x[(c * H + (ho >> 1)) * W + (wo >> 1)];
  }
}

__global__ void im2col3x3_kernel(const float4* __restrict__ x, float4* __restrict__ col, int N, int H,
                                 int W, int C4, int stride, int pad_lo, int Ho, int Wo) {
  pdl_wait();
  pdl_launch_dependents();
  const long long total = (long long)N * Ho * Wo * 9 * C4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % C4);
    long long t = i / C4;
    int tap = (int)(t % 9);
    t /= 9;
    int wo = (int)(t % Wo);
    t /= Wo;
    int ho = (int)(t % Ho);
    int n = (int)(t / Ho);
    int hi = ho * stride + tap / 3 - pad_lo, wi = wo * stride + tap % 3 - pad_lo;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (hi >= 0 && hi < H && wi >= 0 && wi < W)
      v = x[(((long long)n * H + hi) * W + wi) * C4 + c];
    col[i] = v;
  }
}

// ------------------------------------------------------------------------------------------
// GroupNorm / LayerNorm statistics
// ------------------------------------------------------------------------------------------
// Fast path (C % 4 == 0, C/4 <= 768): deterministic two-level reduction, no float atomics.
// grid (nb, N): block b of image n owns a slab of pixels and all channels.  A thread owns up to
// three fixed channel quads (coalesced float4 loads along C) and walks the slab's pixels; the
// per-channel sums are combined through shared memory in a fixed order, folded to per-group
// (sum, sum of squares) in double and written to partial[n][b][g].  The last block to finish
// (atomic ticket) reduces the partials in block order and writes (mean, 1/(std+eps)), so the
// result is bit-reproducible run to run (CUDA-graph replay == eager).
constexpr int GS_THREADS = 256;
constexpr int GS_MAXQ = 3;
__global__ void __launch_bounds__(GS_THREADS)
group_stats_vec_kernel(const float4* __restrict__ x, long long pixels, int C4, int G, int cpg, int slab,
                       double2* __restrict__ partial, unsigned int* __restrict__ ticket, double count, float eps,
                       float2* __restrict__ stats) {
  extern __shared__ float sm[];  // [ppl][2][C]
  __shared__ bool last_block;
  const int C = C4 * 4;
  const int n = blockIdx.y, nb = gridDim.x;
  const int TU = C4 < GS_THREADS ? C4 : GS_THREADS;  // threads across the channel-quad axis
  const int ppl = C4 < GS_THREADS ? GS_THREADS / C4 : 1;
  const int u = threadIdx.x % TU, pl = threadIdx.x / TU;
  float s[GS_MAXQ][4], q[GS_MAXQ][4];
#pragma unroll
  for (int i = 0; i < GS_MAXQ; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) s[i][j] = q[i][j] = 0.f;
  long long p0 = (long long)blockIdx.x * slab, p1 = p0 + slab;
  if (p1 > pixels) p1 = pixels;
  if (pl < ppl) {
    const float4* base = x + (long long)n * pixels * C4;
    for (long long p = p0 + pl; p < p1; p += ppl) {
#pragma unroll
      for (int i = 0; i < GS_MAXQ; ++i) {
        const int qd = u + i * TU;
        if (qd < C4) {
          const float4 v = base[p * C4 + qd];
          s[i][0] += v.x; q[i][0] += v.x * v.x;
          s[i][1] += v.y; q[i][1] += v.y * v.y;
          s[i][2] += v.z; q[i][2] += v.z * v.z;
          s[i][3] += v.w; q[i][3] += v.w * v.w;
        }
      }
    }
#pragma unroll
    for (int i = 0; i < GS_MAXQ; ++i) {
      const int qd = u + i * TU;
      if (qd < C4) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          sm[(pl * 2 + 0) * C + qd * 4 + j] = s[i][j];
          sm[(pl * 2 + 1) * C + qd * 4 + j] = q[i][j];
        }
      }
    }
  }
  __syncthreads();
  // per-channel totals over the pixel lanes (fixed order), kept in plane 0 of the staging buffer
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float a = sm[c], b = sm[C + c];
    for (int l = 1; l < ppl; ++l) {
      a += sm[(l * 2 + 0) * C + c];
      b += sm[(l * 2 + 1) * C + c];
    }
    sm[c] = a;
    sm[C + c] = b;
  }
  __syncthreads();
  // one warp per group: lanes stride over the group's channels, then a fixed xor-shuffle tree
  {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    for (int g = warp; g < G; g += nwarps) {
      double ds = 0.0, dq = 0.0;
      for (int c = g * cpg + lane; c < (g + 1) * cpg; c += 32) {
        ds += (double)sm[c];
        dq += (double)sm[C + c];
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        ds += __shfl_xor_sync(0xffffffffu, ds, o);
        dq += __shfl_xor_sync(0xffffffffu, dq, o);
      }
      if (lane == 0) partial[((long long)n * nb + blockIdx.x) * G + g] = make_double2(ds, dq);
    }
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int total = gridDim.x * gridDim.y;
    const unsigned int t = atomicAdd(ticket, 1u);
    last_block = (t == total - 1);
    if (last_block) *ticket = 0u;  // self-resetting: launches on one stream are serialised
  }
  __syncthreads();
  if (!last_block) return;
  __threadfence();
  // 8 lanes per (image, group) entry: lane j sums blocks j, j+8, ... in order, then a fixed
  // xor-shuffle tree combines the 8 lanes: parallel loads, reproducible summation order
  const int NG = gridDim.y * G;
  for (int e0 = 0; e0 < NG; e0 += GS_THREADS / 8) {
    const int i = e0 + (threadIdx.x >> 3), j = threadIdx.x & 7;
    double ds = 0.0, dq = 0.0;
    if (i < NG) {
      const int nn = i / G, g = i - nn * G;
      for (int b = j; b < nb; b += 8) {
        const double2 v = partial[((long long)nn * nb + b) * G + g];
        ds += v.x;
        dq += v.y;
      }
    }
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
      ds += __shfl_xor_sync(0xffffffffu, ds, o);
      dq += __shfl_xor_sync(0xffffffffu, dq, o);
    }
    if (i < NG && j == 0) {
      const double mean = ds / count;
      double var = dq / count - mean * mean;
      if (var < 0.0) var = 0.0;
      // reference: (x - mean) / (std + eps), biased std  (helpers/utils.mojo:1380, 1868-1870)
      stats[i] = make_float2((float)mean, (float)norm_rstd(var, (double)eps));
    }
  }
}

// ------------------------------------------------------------------------------------------
// Fused GroupNorm / LayerNorm: statistics + normalise (+SiLU, +TF32 rounding) in ONE launch.
// The grid is at most one block per SM (all blocks co-resident), so a generation-counter grid
// barrier separates the two phases:
//   phase 1  every work item (image n, slab of pixels) -> per-group (sum, sum^2) partials in double
//            (same deterministic thread/channel ownership as group_stats_vec_kernel)
//   barrier  last arriver resets the arrival counter and bumps the generation (CUDA-graph safe)
//   phase 2  every block reduces the partials of its image in a fixed order (redundantly, from L2)
//            and normalises its own slabs, which it has just read (L2 / L1 hits).
// The slab partition depends on the image shape only, never on the batch.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(GS_THREADS)
norm_fused_kernel(const float4* __restrict__ x, float4* __restrict__ y, long long pixels, int C4, int G, int cpg,
                  int slabs, int slab, int items, double2* __restrict__ partial, unsigned int* __restrict__ bar,
                  double count, float eps, const float* __restrict__ gamma, const float* __restrict__ beta,
                  float gamma_scalar, int silu, int round) {
  pdl_wait();
  extern __shared__ float sm[];  // [ppl][2][C] staging, then [G] float2 statistics
  const int C = C4 * 4;
  const int TU = C4 < GS_THREADS ? C4 : GS_THREADS;
  const int ppl = C4 < GS_THREADS ? GS_THREADS / C4 : 1;
  const int u = threadIdx.x % TU, pl = threadIdx.x / TU;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;

  // ---- phase 1 ----
  for (int item = blockIdx.x; item < items; item += gridDim.x) {
    const int n = item / slabs, sl = item - n * slabs;
    float s[GS_MAXQ][4], q[GS_MAXQ][4];
#pragma unroll
    for (int i = 0; i < GS_MAXQ; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) s[i][j] = q[i][j] = 0.f;
    long long p0 = (long long)sl * slab, p1 = p0 + slab;
    if (p1 > pixels) p1 = pixels;
    if (pl < ppl) {
      const float4* base = x + (long long)n * pixels * C4;
#pragma unroll 4
      for (long long p = p0 + pl; p < p1; p += ppl) {
#pragma unroll
        for (int i = 0; i < GS_MAXQ; ++i) {
          const int qd = u + i * TU;
          if (qd < C4) {
            const float4 v = base[p * C4 + qd];
            s[i][0] += v.x; q[i][0] += v.x * v.x;
            s[i][1] += v.y; q[i][1] += v.y * v.y;
            s[i][2] += v.z; q[i][2] += v.z * v.z;
            s[i][3] += v.w; q[i][3] += v.w * v.w;
          }
        }
      }
#pragma unroll
      for (int i = 0; i < GS_MAXQ; ++i) {
        const int qd = u + i * TU;
        if (qd < C4) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            sm[(pl * 2 + 0) * C + qd * 4 + j] = s[i][j];
            sm[(pl * 2 + 1) * C + qd * 4 + j] = q[i][j];
          }
        }
      }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      float a = sm[c], b = sm[C + c];
      for (int l = 1; l < ppl; ++l) {
        a += sm[(l * 2 + 0) * C + c];
        b += sm[(l * 2 + 1) * C + c];
      }
      sm[c] = a;
      sm[C + c] = b;
    }
    __syncthreads();
    for (int g = warp; g < G; g += nwarps) {
      double ds = 0.0, dq = 0.0;
      for (int c = g * cpg + lane; c < (g + 1) * cpg; c += 32) {
        ds += (double)sm[c];
        dq += (double)sm[C + c];
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        ds += __shfl_xor_sync(0xffffffffu, ds, o);
        dq += __shfl_xor_sync(0xffffffffu, dq, o);
      }
      if (lane == 0) partial[(long long)item * G + g] = make_double2(ds, dq);
    }
    __syncthreads();
  }

  // ---- grid barrier (all blocks are resident: gridDim.x <= number of SMs) ----
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    volatile unsigned int* gen = bar + 1;
    const unsigned int g0 = *gen;
    if (atomicAdd(bar, 1u) == gridDim.x - 1) {
      *bar = 0u;
      __threadfence();
      atomicAdd(bar + 1, 1u);
    } else {
      unsigned int spins = 0;
      while (*gen == g0) {
        __nanosleep(20);
        if (++spins > (1u << 24)) {  // bounded: lost co-residency traps instead of hanging the GPU
          printf("tsd: norm_fused grid barrier timed out (block %d)\n", blockIdx.x);
          __trap();
        }
      }
    }
    __threadfence();
  }
  __syncthreads();

  // ---- phase 2 ----
  float2* st = reinterpret_cast<float2*>(sm);  // [G]
  int cur_n = -1;
  for (int item = blockIdx.x; item < items; item += gridDim.x) {
    const int n = item / slabs, sl = item - n * slabs;
    if (n != cur_n) {
      __syncthreads();
      // 8 lanes per group: lane j sums slabs j, j+8, ... in order; fixed xor-shuffle tree
      for (int e0 = 0; e0 < G; e0 += GS_THREADS / 8) {
        const int g = e0 + (threadIdx.x >> 3), j = threadIdx.x & 7;
        double ds = 0.0, dq = 0.0;
        if (g < G) {
#pragma unroll 4
          for (int b = j; b < slabs; b += 8) {
            const double2 v = __ldcg(&partial[((long long)n * slabs + b) * G + g]);
            ds += v.x;
            dq += v.y;
          }
        }
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
          ds += __shfl_xor_sync(0xffffffffu, ds, o);
          dq += __shfl_xor_sync(0xffffffffu, dq, o);
        }
        if (g < G && j == 0) {
          const double mean = ds / count;
          double var = dq / count - mean * mean;
          if (var < 0.0) var = 0.0;
          // reference: (x - mean) / (std + eps), biased std  (helpers/utils.mojo:1380, 1868-1870)
          st[g] = make_float2((float)mean, (float)norm_rstd(var, (double)eps));
        }
      }
      __syncthreads();
      cur_n = n;
    }
    if (pl < ppl) {
      float mu[GS_MAXQ][4], sc[GS_MAXQ][4], sh[GS_MAXQ][4];
#pragma unroll
      for (int i = 0; i < GS_MAXQ; ++i) {
        const int qd = u + i * TU;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          mu[i][j] = 0.f; sc[i][j] = 0.f; sh[i][j] = 0.f;
          if (qd < C4) {
            const int c = qd * 4 + j;
            const float2 t = st[c / cpg];
            mu[i][j] = t.x;
            sc[i][j] = t.y * gamma_scalar * (gamma ? gamma[c] : 1.0f);
            sh[i][j] = beta ? beta[c] : 0.0f;
          }
        }
      }
      long long p0 = (long long)sl * slab, p1 = p0 + slab;
      if (p1 > pixels) p1 = pixels;
      const float4* base = x + (long long)n * pixels * C4;
      float4* obase = y + (long long)n * pixels * C4;
#pragma unroll 2
      for (long long p = p0 + pl; p < p1; p += ppl) {
#pragma unroll
        for (int i = 0; i < GS_MAXQ; ++i) {
          const int qd = u + i * TU;
          if (qd < C4) {
            const float4 v = base[p * C4 + qd];
            float o[4] = {(v.x - mu[i][0]) * sc[i][0] + sh[i][0], (v.y - mu[i][1]) * sc[i][1] + sh[i][1],
                          (v.z - mu[i][2]) * sc[i][2] + sh[i][2], (v.w - mu[i][3]) * sc[i][3] + sh[i][3]};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if (silu) o[j] = silu_f(o[j]);
              if (round) o[j] = rna_tf32(o[j]);
            }
            obase[p * C4 + qd] = make_float4(o[0], o[1], o[2], o[3]);
          }
        }
      }
    }
  }
}


// ------------------------------------------------------------------------------------------
// Fused GroupNorm / LayerNorm v2: statistics + normalise (+SiLU, +TF32 rounding) in ONE launch with
// the activation read from L2 ONCE.  One block per (image, slab of pixels); all blocks are
// co-resident (grid <= 2 per SM), so a two-level barrier separates the phases:
//   phase 1  load the slab into registers, per-group (sum, sum^2) -> part[item][g]
//   level 1  blocks arrive on the counter of their sub-group (<= 16 consecutive slabs of one image);
//            the last arriver folds the sub-group's partials -> part2[sub][g] and arrives on the root
//   level 2  the last sub-group bumps the generation word every block is polling
//   phase 2  fold part2 of the image (<= 19 entries per group), normalise the registers, store.
// Contended atomics stay <= 16-19 per address; every fold has a fixed order (bit-reproducible).
// The source is either a plain activation or the split-K partials of a GEMM (+bias, +residual): then
// this kernel IS the split-K reduction and `raw` (optional) receives the un-normalised sum.
// ------------------------------------------------------------------------------------------
constexpr int NF_THREADS = 256, NF_SUB = 16;
struct NormFused2Params {
  const float4* x;      // plain source, or the first split's partial
  int splits;           // 1 = plain
  long long split_stride4;  // float4 elements between splits
  int ldx4;             // row stride of the source in float4 (n_pad / 4 for split-K partials, C4 otherwise)
  const float* bias;    // split-K source only: [C] per image (bias + img * bias_img_stride) or nullptr
  int bias_img_stride;
  const float4* residual;  // split-K source only
  float4* raw;          // split-K / two-source: un-normalised result (nullptr: not needed)
  const float4* x2;     // two-source (channel concat): channels [c4a*4, C) come from x2 [rows][C4 - c4a]
  int c4a;
  float4* y;
  int pixels, C4, G, cpg;
  int slabs_per_img, slab, subs_per_img;
  float2* part;         // [items][G]
  float2* part2;        // [N * subs_per_img][G]
  unsigned int* bar;    // [0] root count, [1] generation, [32 * (1 + k)] sub-group counters
  float inv_count, eps;
  const float* gamma;
  const float* beta;
  float gamma_scalar;
  int silu, round;
  int trace;  // lab: block 0 prints its phase timestamps
};

template <int NQ, int MAXIT, bool CACHE>
__global__ void __launch_bounds__(NF_THREADS, 2)
norm_fused2_kernel(const NormFused2Params p) {
  pdl_wait();
  extern __shared__ float nf_sm[];  // [ppl][2][C] staging; later [G] float2 statistics
  __shared__ unsigned int flag_s;
  __shared__ unsigned int gen_s;
  const int C4 = p.C4, C = C4 * 4;
  const int TU = C4 < NF_THREADS ? C4 : NF_THREADS;
  const int ppl = C4 < NF_THREADS ? NF_THREADS / C4 : 1;
  const int u = threadIdx.x % TU, pl = threadIdx.x / TU;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int item = blockIdx.x;
  const int n = item / p.slabs_per_img, sl = item - n * p.slabs_per_img;
  const int p0 = sl * p.slab;
  int p1 = p0 + p.slab;
  if (p1 > p.pixels) p1 = p.pixels;
  long long t0 = 0, t1 = 0, t2 = 0, t3 = 0, t4 = 0;
  if (threadIdx.x == 0) {
    t0 = clock64();
    gen_s = *reinterpret_cast<volatile unsigned int*>(p.bar + 1);  // before this block arrives
  }

  auto load = [&](int px, int qd) -> float4 {
    const long long row = (long long)n * p.pixels + px;
    if (p.x2 != nullptr) {  // channel concat of two tensors, written out as a by-product
      const float4 a = qd < p.c4a ? p.x[row * p.c4a + qd] : p.x2[row * (C4 - p.c4a) + (qd - p.c4a)];
      if (p.raw) p.raw[row * C4 + qd] = a;
      return a;
    }
    float4 a = p.x[row * p.ldx4 + qd];
    if (p.splits > 1) {
      // every split's tile is requested before the first add (a warp issues in order: load -> add -> load -> add
      // costs one L2 round trip per split); the sum stays in split order
      const float4* src = p.x + row * p.ldx4 + qd;
      for (int s0 = 1; s0 < p.splits; s0 += 4) {
        float4 t[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          t[k] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (s0 + k < p.splits) t[k] = src[(long long)(s0 + k) * p.split_stride4];
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          a.x += t[k].x; a.y += t[k].y; a.z += t[k].z; a.w += t[k].w;
        }
      }
      if (p.bias) {
        const float4 b = *reinterpret_cast<const float4*>(p.bias + (long long)n * p.bias_img_stride + qd * 4);
        a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
      }
      if (p.residual) {
        const float4 t = p.residual[row * C4 + qd];
        a.x += t.x; a.y += t.y; a.z += t.z; a.w += t.w;
      }
      if (p.raw) p.raw[row * C4 + qd] = a;
    }
    return a;
  };

  // ---- phase 1 ----
  float4 v[CACHE ? MAXIT : 1][NQ];
  float s[NQ][4], q[NQ][4];
#pragma unroll
  for (int i = 0; i < NQ; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) s[i][j] = q[i][j] = 0.f;
  if (pl < ppl) {
    if (CACHE) {
      const int iters = (p1 - p0 - pl + ppl - 1) / ppl;  // pixels this thread owns (uniform per pixel lane)
#pragma unroll
      for (int it = 0; it < MAXIT; ++it) {
        if (it >= iters) break;  // a branch, not predication: unused iterations cost no issue slots
        const int px = p0 + pl + it * ppl;
#pragma unroll
        for (int i = 0; i < NQ; ++i) {
          const int qd = u + i * TU;
          v[it][i] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (qd < C4) v[it][i] = load(px, qd);
        }
      }
#pragma unroll
      for (int it = 0; it < MAXIT; ++it) {
        if (it >= iters) break;
#pragma unroll
        for (int i = 0; i < NQ; ++i) {
          const float4 t = v[it][i];
          s[i][0] += t.x; q[i][0] = fmaf(t.x, t.x, q[i][0]);
          s[i][1] += t.y; q[i][1] = fmaf(t.y, t.y, q[i][1]);
          s[i][2] += t.z; q[i][2] = fmaf(t.z, t.z, q[i][2]);
          s[i][3] += t.w; q[i][3] = fmaf(t.w, t.w, q[i][3]);
        }
      }
    } else {
#pragma unroll 4
      for (int px = p0 + pl; px < p1; px += ppl) {
#pragma unroll
        for (int i = 0; i < NQ; ++i) {
          const int qd = u + i * TU;
          if (qd < C4) {
            const float4 t = load(px, qd);
            s[i][0] += t.x; q[i][0] = fmaf(t.x, t.x, q[i][0]);
            s[i][1] += t.y; q[i][1] = fmaf(t.y, t.y, q[i][1]);
            s[i][2] += t.z; q[i][2] = fmaf(t.z, t.z, q[i][2]);
            s[i][3] += t.w; q[i][3] = fmaf(t.w, t.w, q[i][3]);
          }
        }
      }
    }
#pragma unroll
    for (int i = 0; i < NQ; ++i) {
      const int qd = u + i * TU;
      if (qd < C4) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          nf_sm[(pl * 2 + 0) * C + qd * 4 + j] = s[i][j];
          nf_sm[(pl * 2 + 1) * C + qd * 4 + j] = q[i][j];
        }
      }
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += NF_THREADS) {  // fold the pixel lanes in a fixed order
    float a = nf_sm[c], b = nf_sm[C + c];
    for (int l = 1; l < ppl; ++l) {
      a += nf_sm[(l * 2 + 0) * C + c];
      b += nf_sm[(l * 2 + 1) * C + c];
    }
    nf_sm[c] = a;
    nf_sm[C + c] = b;
  }
  __syncthreads();
  for (int g = warp; g < p.G; g += NF_THREADS / 32) {  // one warp per group
    float a = 0.f, b = 0.f;
    for (int c = g * p.cpg + lane; c < (g + 1) * p.cpg; c += 32) {
      a += nf_sm[c];
      b += nf_sm[C + c];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o);
      b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    if (lane == 0) p.part[(long long)item * p.G + g] = make_float2(a, b);
  }

  if (threadIdx.x == 0) t1 = clock64();
  // ---- level 1: sub-group of <= NF_SUB consecutive slabs of this image ----
  const int sub = sl / NF_SUB;
  const int sub_first = sub * NF_SUB;
  int members = p.slabs_per_img - sub_first;
  if (members > NF_SUB) members = NF_SUB;
  const int sub_global = n * p.subs_per_img + sub;
  unsigned int* sub_ctr = p.bar + 32 * (1 + sub_global);
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    flag_s = (atomicAdd(sub_ctr, 1u) == (unsigned int)members - 1) ? 1u : 0u;
  }
  __syncthreads();
  if (flag_s) {
    __threadfence();
    const float2* src = p.part + ((long long)n * p.slabs_per_img + sub_first) * p.G;
    for (int g0 = warp; g0 < p.G; g0 += 4 * (NF_THREADS / 32)) {  // 4 groups per pass: their L2 loads overlap
      float2 t[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int g = g0 + k * (NF_THREADS / 32);
        t[k] = make_float2(0.f, 0.f);
        if (g < p.G && lane < members) t[k] = __ldcg(src + (long long)lane * p.G + g);
      }
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) {  // NF_SUB = 16 lanes
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          t[k].x += __shfl_xor_sync(0xffffffffu, t[k].x, o);
          t[k].y += __shfl_xor_sync(0xffffffffu, t[k].y, o);
        }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int g = g0 + k * (NF_THREADS / 32);
        if (lane == 0 && g < p.G) p.part2[(long long)sub_global * p.G + g] = t[k];
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      *sub_ctr = 0u;  // self-resetting: launches on one stream are serialised
      __threadfence();
      const unsigned int total_subs = (gridDim.x / p.slabs_per_img) * p.subs_per_img;
      if (atomicAdd(p.bar, 1u) == total_subs - 1) {
        *p.bar = 0u;
        __threadfence();
        atomicAdd(p.bar + 1, 1u);  // release every block
      }
    }
  }
  // ---- level 2: wait for the generation to move ----
  if (threadIdx.x == 0) {
    t2 = clock64();
    volatile unsigned int* gen = p.bar + 1;
    unsigned int spins = 0;
    while (*gen == gen_s) {
      __nanosleep(32);
      if (++spins > (1u << 24)) {
        printf("tsd: norm_fused2 grid barrier timed out (block %d)\n", blockIdx.x);
        __trap();
      }
    }
    __threadfence();
    t3 = clock64();
  }
  __syncthreads();

  // ---- phase 2: statistics of this image, normalise ----
  float2* st = reinterpret_cast<float2*>(nf_sm);
  {
    const float2* src = p.part2 + (long long)n * p.subs_per_img * p.G;
    for (int g0 = warp; g0 < p.G; g0 += 4 * (NF_THREADS / 32)) {  // 4 groups per pass: their L2 loads overlap
      float2 t[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int g = g0 + k * (NF_THREADS / 32);
        t[k] = make_float2(0.f, 0.f);
        if (g < p.G && lane < p.subs_per_img) t[k] = __ldcg(src + (long long)lane * p.G + g);  // subs_per_img <= 32
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {  // fp32 tree (fp64 issue is slow on this part); the subtraction below is fp64
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          t[k].x += __shfl_xor_sync(0xffffffffu, t[k].x, o);
          t[k].y += __shfl_xor_sync(0xffffffffu, t[k].y, o);
        }
      }
      // every lane holds the four totals: lane k finishes group k
      float2 mine = t[0];
      if (lane == 1) mine = t[1];
      if (lane == 2) mine = t[2];
      if (lane == 3) mine = t[3];
      const int g = g0 + lane * (NF_THREADS / 32);
      if (lane < 4 && g < p.G) {
        const double mean = (double)mine.x * (double)p.inv_count;
        double var = (double)mine.y * (double)p.inv_count - mean * mean;
        if (var < 0.0) var = 0.0;
        // reference: (x - mean) / (std + eps), biased std  (helpers/utils.mojo:1380, 1868-1870)
        st[g] = make_float2((float)mean, norm_rstd((float)var, p.eps));
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) t4 = clock64();
  if (pl >= ppl) return;
  float mu[NQ][4], sc[NQ][4], sh[NQ][4];
#pragma unroll
  for (int i = 0; i < NQ; ++i) {
    const int qd = u + i * TU;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      mu[i][j] = 0.f; sc[i][j] = 0.f; sh[i][j] = 0.f;
      if (qd < C4) {
        const int c = qd * 4 + j;
        const float2 t = st[c / p.cpg];
        mu[i][j] = t.x;
        sc[i][j] = t.y * p.gamma_scalar * (p.gamma ? p.gamma[c] : 1.0f);
        sh[i][j] = p.beta ? p.beta[c] : 0.0f;
      }
    }
  }
  float4* obase = p.y + (long long)n * p.pixels * C4;
  auto emit = [&](int px, int i, int qd, float4 t) {
    float o[4] = {(t.x - mu[i][0]) * sc[i][0] + sh[i][0], (t.y - mu[i][1]) * sc[i][1] + sh[i][1],
                  (t.z - mu[i][2]) * sc[i][2] + sh[i][2], (t.w - mu[i][3]) * sc[i][3] + sh[i][3]};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (p.silu) o[j] = silu_f(o[j]);
      if (p.round) o[j] = rna_tf32(o[j]);
    }
    obase[(long long)px * C4 + qd] = make_float4(o[0], o[1], o[2], o[3]);
  };
  if (CACHE) {
    const int iters = (p1 - p0 - pl + ppl - 1) / ppl;
#pragma unroll
    for (int it = 0; it < MAXIT; ++it) {
      if (it >= iters) break;
      const int px = p0 + pl + it * ppl;
#pragma unroll
      for (int i = 0; i < NQ; ++i) {
        const int qd = u + i * TU;
        if (qd < C4) emit(px, i, qd, v[it][i]);
      }
    }
  } else {
    const bool from_raw = p.splits > 1 || p.x2 != nullptr;  // re-read mode of a split-K / two-source input reads back
    const float4* src = from_raw ? p.raw : p.x;             // the raw tensor this kernel wrote in phase 1
    const int ld = from_raw ? C4 : p.ldx4;
#pragma unroll 2
    for (int px = p0 + pl; px < p1; px += ppl) {
#pragma unroll
      for (int i = 0; i < NQ; ++i) {
        const int qd = u + i * TU;
        if (qd < C4) emit(px, i, qd, src[((long long)n * p.pixels + px) * ld + qd]);
      }
    }
  }
  if (p.trace && threadIdx.x == 0 && blockIdx.x == 0)
    printf("norm2 trace: phase1 %lld arrive..spin %lld spin %lld fold %lld apply %lld (clk) grid %d\n", t1 - t0, t2 - t1, t3 - t2,
           t4 - t3, clock64() - t4, gridDim.x);
}


// ------------------------------------------------------------------------------------------
// Cluster GroupNorm: ONE thread-block cluster per (image, group).  The CTAs of a cluster split the pixels of the
// group; each keeps its slab of the group (cpg channels x its pixels) in shared memory, reduces (sum, sum^2) over the
// block, exchanges the partials through distributed shared memory (one cluster barrier - a few hundred cycles where
// the grid barrier of norm_fused2 costs ~9K) and normalises out of shared memory.  Groups never talk to each other,
// so there is no grid-wide synchronisation and no co-residency requirement beyond what a cluster launch guarantees.
// The source is a plain activation, the split-K partials of a GEMM (+bias, +residual; `raw` receives the reduced
// tensor: this kernel IS the reduction) or a channel concat of two tensors (`raw` receives the concatenation).
// V = floats per access (4 when cpg % 4 == 0, else 2); fixed summation orders (bit-reproducible).
// ------------------------------------------------------------------------------------------
constexpr int NC_THREADS = 256;
struct NormClusterParams {
  const float* x;
  int splits;
  long long split_stride;  // floats between splits
  int ldx;                 // row stride of the source (floats)
  const float* bias;
  int bias_img_stride;
  const float* residual;
  float* raw;
  const float* x2;
  int c_a;
  float* y;
  int pixels, C, G, cpg;
  int cs, ppc;             // CTAs per cluster, pixels per CTA
  int hc, ppl;             // vectors per pixel of the group (cpg / V), pixel lanes per pass (NC_THREADS / hc)
  float inv_count, eps;
  const float* gamma;
  const float* beta;
  float gamma_scalar;
  int silu, round;
};

template <int V> struct NcVec;
template <> struct NcVec<2> { using T = float2; };
template <> struct NcVec<4> { using T = float4; };

__device__ __forceinline__ uint32_t nc_cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r));
  return r;
}
__device__ __forceinline__ float2 nc_ld_remote_f2(const void* local, uint32_t cta) {
  const uint32_t laddr = (uint32_t)__cvta_generic_to_shared(local);
  uint32_t raddr;
  float2 v;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(raddr) : "r"(laddr), "r"(cta));
  asm volatile("ld.shared::cluster.v2.f32 {%0, %1}, [%2];\n" : "=f"(v.x), "=f"(v.y) : "r"(raddr) : "memory");
  return v;
}

template <int V>
__global__ void __launch_bounds__(NC_THREADS)
norm_cluster_kernel(const NormClusterParams p) {
  using VT = typename NcVec<V>::T;
  extern __shared__ __align__(16) float nc_sm[];  // the slab: [ppc][hc] vectors, thread-private entries
  __shared__ float2 warp_part[NC_THREADS / 32];
  __shared__ float2 cta_part;   // read by the other CTAs of the cluster
  __shared__ float2 stat_s;     // (mean, 1 / (std + eps))
  const uint32_t rank = nc_cluster_rank();
  const int cluster = blockIdx.x / p.cs;
  const int n = cluster / p.G, g = cluster - n * p.G;
  const int u = threadIdx.x % p.hc, pl = threadIdx.x / p.hc;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int p0 = (int)rank * p.ppc;
  int p1 = p0 + p.ppc;
  if (p1 > p.pixels) p1 = p.pixels;
  const int c0 = g * p.cpg + u * V;  // first channel of this thread's vector
  const bool active = pl < p.ppl;
  VT* slab = reinterpret_cast<VT*>(nc_sm);

  // per-channel affine terms do not depend on the producer: fetch them before the dependency wait
  float ga[V], be[V];
#pragma unroll
  for (int j = 0; j < V; ++j) {
    ga[j] = p.gamma_scalar * ((p.gamma && active) ? p.gamma[c0 + j] : 1.0f);
    be[j] = (p.beta && active) ? p.beta[c0 + j] : 0.0f;
  }
  pdl_wait();

  float bias_v[V];
#pragma unroll
  for (int j = 0; j < V; ++j) bias_v[j] = (p.bias && active) ? p.bias[(long long)n * p.bias_img_stride + c0 + j] : 0.0f;

  auto ldv = [&](const float* q) -> VT { return *reinterpret_cast<const VT*>(q); };
  auto add = [&](VT& a, const VT& b) {
    a.x += b.x; a.y += b.y;
    if constexpr (V == 4) { a.z += b.z; a.w += b.w; }
  };
  auto load = [&](int px) -> VT {
    const long long row = (long long)n * p.pixels + px;
    VT a;
    if (p.x2 != nullptr) {
      a = c0 < p.c_a ? ldv(p.x + row * p.c_a + c0) : ldv(p.x2 + row * (p.C - p.c_a) + (c0 - p.c_a));
    } else if (p.splits > 1) {
      const float* src = p.x + row * p.ldx + c0;
      a = ldv(src);
      for (int s0 = 1; s0 < p.splits; s0 += 4) {  // every split requested before the first add; sum in split order
        VT t[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if constexpr (V == 4) t[k] = make_float4(0.f, 0.f, 0.f, 0.f); else t[k] = make_float2(0.f, 0.f);
          if (s0 + k < p.splits) t[k] = ldv(src + (long long)(s0 + k) * p.split_stride);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) add(a, t[k]);
      }
    } else {
      a = ldv(p.x + row * p.ldx + c0);
    }
    if (p.splits > 1 || p.bias != nullptr || p.residual != nullptr) {
      a.x += bias_v[0]; a.y += bias_v[1];
      if constexpr (V == 4) { a.z += bias_v[2]; a.w += bias_v[3]; }
      if (p.residual) add(a, ldv(p.residual + row * p.C + c0));
    }
    if (p.raw) *reinterpret_cast<VT*>(p.raw + row * p.C + c0) = a;
    return a;
  };

  // ---- phase 1: slab -> shared memory, (sum, sum^2) ----
  float s = 0.f, q = 0.f;
  if (active) {
    int it = 0;
    int px = p0 + pl;
    for (; px + p.ppl < p1; px += 2 * p.ppl, it += 2) {  // two pixels per trip: both rows' loads are in flight together
      const VT a = load(px), b = load(px + p.ppl);
      slab[(it * p.ppl + pl) * p.hc + u] = a;
      slab[((it + 1) * p.ppl + pl) * p.hc + u] = b;
      s += a.x; q = fmaf(a.x, a.x, q); s += a.y; q = fmaf(a.y, a.y, q);
      if constexpr (V == 4) { s += a.z; q = fmaf(a.z, a.z, q); s += a.w; q = fmaf(a.w, a.w, q); }
      s += b.x; q = fmaf(b.x, b.x, q); s += b.y; q = fmaf(b.y, b.y, q);
      if constexpr (V == 4) { s += b.z; q = fmaf(b.z, b.z, q); s += b.w; q = fmaf(b.w, b.w, q); }
    }
    if (px < p1) {
      const VT a = load(px);
      slab[(it * p.ppl + pl) * p.hc + u] = a;
      s += a.x; q = fmaf(a.x, a.x, q); s += a.y; q = fmaf(a.y, a.y, q);
      if constexpr (V == 4) { s += a.z; q = fmaf(a.z, a.z, q); s += a.w; q = fmaf(a.w, a.w, q); }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    q += __shfl_xor_sync(0xffffffffu, q, o);
  }
  if (lane == 0) warp_part[warp] = make_float2(s, q);
  __syncthreads();
  if (threadIdx.x == 0) {
    float2 t = warp_part[0];
#pragma unroll
    for (int w = 1; w < NC_THREADS / 32; ++w) { t.x += warp_part[w].x; t.y += warp_part[w].y; }
    cta_part = t;
  }
  // ---- exchange inside the cluster ----
  asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
  if (threadIdx.x == 0) {
    float2 t = nc_ld_remote_f2(&cta_part, 0);
    for (int r = 1; r < p.cs; ++r) {
      const float2 o = nc_ld_remote_f2(&cta_part, (uint32_t)r);
      t.x += o.x; t.y += o.y;
    }
    const double mean = (double)t.x * (double)p.inv_count;
    double var = (double)t.y * (double)p.inv_count - mean * mean;
    if (var < 0.0) var = 0.0;
    // reference: (x - mean) / (std + eps), biased std  (helpers/utils.mojo:1380, 1868-1870)
    stat_s = make_float2((float)mean, norm_rstd((float)var, p.eps));
  }
  __syncthreads();
  // the second barrier keeps every CTA (and its shared memory) alive until all remote reads are done
  asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
  pdl_launch_dependents();

  // ---- phase 2: normalise out of shared memory ----
  if (active) {
    const float mu = stat_s.x, rstd = stat_s.y;
    float sc[V];
#pragma unroll
    for (int j = 0; j < V; ++j) sc[j] = rstd * ga[j];
    float* obase = p.y + (long long)n * p.pixels * p.C + c0;
    int it = 0;
#pragma unroll 2
    for (int px = p0 + pl; px < p1; px += p.ppl, ++it) {
      const VT a = slab[(it * p.ppl + pl) * p.hc + u];
      float o[V];
      o[0] = (a.x - mu) * sc[0] + be[0];
      o[1] = (a.y - mu) * sc[1] + be[1];
      if constexpr (V == 4) { o[2] = (a.z - mu) * sc[2] + be[2]; o[3] = (a.w - mu) * sc[3] + be[3]; }
#pragma unroll
      for (int j = 0; j < V; ++j) {
        if (p.silu) o[j] = silu_f(o[j]);
        if (p.round) o[j] = rna_tf32(o[j]);
      }
      VT out;
      out.x = o[0]; out.y = o[1];
      if constexpr (V == 4) { out.z = o[2]; out.w = o[3]; }
      *reinterpret_cast<VT*>(obase + (long long)px * p.C) = out;
    }
  }
  asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}

// General path: one block per (n, g).
__global__ void group_stats_general_kernel(const float* __restrict__ x, long long pixels, int C,
                                           int G, int cpg, double* __restrict__ accum) {
  const int g = blockIdx.x, n = blockIdx.y;
  const float* base = x + (long long)n * pixels * C + (long long)g * cpg;
  double s = 0.0, q = 0.0;
  const long long total = pixels * cpg;
  for (long long i = threadIdx.x; i < total; i += blockDim.x) {
    long long p = i / cpg;
    int c = (int)(i - p * cpg);
    float v = base[p * C + c];
    s += v;
    q += (double)v * v;
  }
  __shared__ double rs[32], rq[32];
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffff, s, o);
    q += __shfl_xor_sync(0xffffffff, q, o);
  }
  if ((threadIdx.x & 31) == 0) {
    rs[threadIdx.x >> 5] = s;
    rq[threadIdx.x >> 5] = q;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ts = 0, tq = 0;
    for (int i = 0; i < (blockDim.x + 31) / 32; ++i) {
      ts += rs[i];
      tq += rq[i];
    }
    accum[((long long)n * G + g) * 2] = ts;
    accum[((long long)n * G + g) * 2 + 1] = tq;
  }
}

__global__ void group_stats_finalize_kernel(const double* __restrict__ accum, int NG, double count,
                                            float eps, float2* __restrict__ stats) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= NG) return;
  double mean = accum[2 * i] / count;
  double var = accum[2 * i + 1] / count - mean * mean;
  if (var < 0.0) var = 0.0;
  // reference: (x - mean) / (std + eps), biased std  (helpers/utils.mojo:1380, 1868-1870)
  double inv = norm_rstd(var, (double)eps);
  stats[i] = make_float2((float)mean, (float)inv);
}

template <int VEC>
__global__ void norm_apply_kernel(const float* __restrict__ x, const float2* __restrict__ stats,
                                  const float* __restrict__ gamma, const float* __restrict__ beta,
                                  float gamma_scalar, float* __restrict__ y, int N, int H, int W,
                                  int C, int G, int cpg, int silu, int up, int round) {
  pdl_wait();
  pdl_launch_dependents();
  const int CV = C / VEC;
  const int Ho = up ? 2 * H : H, Wo = up ? 2 * W : W;
  const long long total = (long long)N * Ho * Wo * CV;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int cv = (int)(i % CV);
    long long t = i / CV;
    int wo = (int)(t % Wo);
    t /= Wo;
    int ho = (int)(t % Ho);
    int n = (int)(t / Ho);
    int hi = up ? (ho >> 1) : ho, wi = up ? (wo >> 1) : wo;
    const float* src = x + ((((long long)n * H + hi) * W + wi) * C + (long long)cv * VEC);
    float v[VEC];
    if (VEC == 4) {
      float4 t4 = *reinterpret_cast<const float4*>(src);
      v[0] = t4.x; v[1] = t4.y; v[2] = t4.z; v[3] = t4.w;
    } else {
      v[0] = src[0];
    }
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      int c = cv * VEC + j;
      float2 st = stats[n * G + c / cpg];
      float o = (v[j] - st.x) * st.y * gamma_scalar;
      if (gamma) o *= gamma[c];
      if (beta) o += beta[c];
      if (silu) o = silu_f(o);
      if (round) o = rna_tf32(o);
      v[j] = o;
    }
    float* dst = y + i * VEC;
    if (VEC == 4)
      *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
    else
      dst[0] = v[0];
  }
}

// Normalise pass for producer-side partial statistics (norm_stats.cuh): grid (slabs, N).  Every
// block first folds the partials of its image to (mean, 1/(std+eps)) per group (a few hundred
// float2 from L2), then a thread owns up to three fixed channel quads (coalesced float4 along C)
// and walks the pixels of its slab, `ppl` pixels in flight per block - no index divisions.
// NQ = channel quads per thread (1 for C <= 1024 per segment).  One instantiation per NQ keeps the code a launch touches
// small: these kernels run a few microseconds between GEMMs whose own code evicts them from the instruction caches, so
// every launch starts cold and pays for each 128 B line of straight-line code it walks through.
template <int NQ>
__global__ void __launch_bounds__(GS_THREADS)
norm_apply_partial_kernel(const float4* __restrict__ x, float4* __restrict__ y, const NormStatsReq req,
                          const float* __restrict__ gamma, const float* __restrict__ beta, float gamma_scalar,
                          int pixels, int C4, int slab, int silu, int round, int trace, int c4_seg) {
  long long t0 = 0, t1 = 0, t2 = 0;
  if (trace) t0 = clock64();
  pdl_wait();
  pdl_launch_dependents();
  if (trace) t1 = clock64();
  __shared__ float2 st[512];
  const int n = blockIdx.y;
  // blockIdx.z: channel segment of c4_seg quads (many-group norms: a block then folds only the groups of its segment)
  const int c4_0 = blockIdx.z * c4_seg;
  const int C4s = (C4 - c4_0) < c4_seg ? (C4 - c4_0) : c4_seg;  // quads of this segment
  const int TU = C4s < GS_THREADS ? C4s : GS_THREADS;
  const int ppl = C4s < GS_THREADS ? GS_THREADS / C4s : 1;
  const int u = threadIdx.x % TU, pl = threadIdx.x / TU;
  int p0 = blockIdx.x * slab, p1 = p0 + slab;
  if (p1 > pixels) p1 = pixels;
  const float4* base = x + (long long)n * pixels * C4 + c4_0;
  float4* obase = y + (long long)n * pixels * C4 + c4_0;
  const int g_begin = norm_fastdiv(c4_0 * 4, req.cpg_magic);
  const int g_end = norm_fastdiv((c4_0 + C4s) * 4 - 1, req.cpg_magic) + 1;
  // The first NP pixels of this thread are requested BEFORE the statistics are folded: the fold is a latency chain
  // (L2 round trip, shuffles, barrier: 6 000+ cycles measured) that needs no activation data, and the activation loads
  // need no statistics - the two now overlap instead of running back to back.
  constexpr int NP = 3;
  float4 pre[NP][NQ];
  const bool active = pl < ppl;
  auto prefetch = [&]() {
#pragma unroll
    for (int k = 0; k < NP; ++k) {
      const int p = p0 + pl + k * ppl;
#pragma unroll
      for (int i = 0; i < NQ; ++i) {
        const int qd = u + i * TU;
        pre[k][i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (active && p < p1 && qd < C4s) pre[k][i] = base[(long long)p * C4 + qd];
      }
    }
  };
  long long trf[3] = {0, 0, 0};
  // the (tiny) statistics table is requested first, the activation prefetch queues behind it
  norm_stats_fold(req, n, threadIdx.x, GS_THREADS, st, trace ? trf : nullptr, prefetch, g_begin, g_end);
  const long long t1b = trace ? clock64() : 0;
  __syncthreads();
  if (trace) t2 = clock64();
  if (trace && threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0)
    printf("  fold detail: prefetch issue + index math %lld, partial loads %lld, shuffles %lld, finalise %lld, barrier %lld (clk)\n",
           trf[0] - t1, trf[1] - trf[0], trf[2] - trf[1], t1b - trf[2], t2 - t1b);
  if (!active) return;
  const unsigned cpg_magic = req.cpg_magic;  // c / cpg = umulhi(c, magic), exact for c, cpg < 2^16
  float mu[NQ][4], sc[NQ][4], sh[NQ][4];
#pragma unroll
  for (int i = 0; i < NQ; ++i) {
    const int qd = u + i * TU;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      mu[i][j] = 0.f; sc[i][j] = 0.f; sh[i][j] = 0.f;
      if (qd < C4s) {
        const int c = (c4_0 + qd) * 4 + j;
        const float2 t = st[norm_fastdiv(c, cpg_magic) - g_begin];
        mu[i][j] = t.x;
        sc[i][j] = t.y * gamma_scalar * (gamma ? gamma[c] : 1.0f);
        sh[i][j] = beta ? beta[c] : 0.0f;
      }
    }
  }
  auto emit = [&](const float4 v, int i, long long idx) {
    float o[4] = {(v.x - mu[i][0]) * sc[i][0] + sh[i][0], (v.y - mu[i][1]) * sc[i][1] + sh[i][1],
                  (v.z - mu[i][2]) * sc[i][2] + sh[i][2], (v.w - mu[i][3]) * sc[i][3] + sh[i][3]};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (silu) o[j] = silu_f(o[j]);
      if (round) o[j] = __uint_as_float((__float_as_uint(o[j]) + 0x1000u) & 0xffffe000u);  // round to nearest (ties away) at TF32 precision: the bits cvt.rna.tf32 gives for finite values, two instructions
    }
    obase[idx] = make_float4(o[0], o[1], o[2], o[3]);
  };
#pragma unroll
  for (int k = 0; k < NP; ++k) {
    const int p = p0 + pl + k * ppl;
#pragma unroll
    for (int i = 0; i < NQ; ++i) {
      const int qd = u + i * TU;
      if (p < p1 && qd < C4s) emit(pre[k][i], i, (long long)p * C4 + qd);
    }
  }
#pragma unroll 4
  for (int p = p0 + pl + NP * ppl; p < p1; p += ppl) {
#pragma unroll
    for (int i = 0; i < NQ; ++i) {
      const int qd = u + i * TU;
      if (qd < C4s) emit(base[(long long)p * C4 + qd], i, (long long)p * C4 + qd);
    }
  }
  if (trace && threadIdx.x == 0 && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1) && blockIdx.y == 0)
    printf("norm_apply_partial trace block %d/%d: pdl_wait %lld fold %lld normalise %lld (clk)\n", blockIdx.x, gridDim.x,
           t1 - t0, t2 - t1, clock64() - t2);
}

// ------------------------------------------------------------------------------------------
// elementwise
// ------------------------------------------------------------------------------------------
__global__ void fill_uniform_kernel(float* __restrict__ p, long long n, uint64_t seed, float lo,
                                    float hi) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    uint64_t z = seed * 0x9E3779B97F4A7C15ull + (uint64_t)i + 0x632BE59BD9B4E019ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    float u = (float)(z >> 40) * (1.0f / 16777216.0f);  // [0,1)
    p[i] = lo + (hi - lo) * u;
  }
}

__global__ void unary_kernel(const float* __restrict__ x, float* __restrict__ y, long long n, int op,
                             float scalar) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    float v = x[i];
    if (op == UNARY_SILU) v = v / (1.0f + expf(-v));
    else if (op == UNARY_GELU) v = gelu_f(v);
    else if (op == UNARY_SCALE) v = v * scalar;
    else if (op == UNARY_QUICKGELU) v = v / (1.0f + expf(-1.702f * v));  // x * sigmoid(1.702 x), clip.mojo:49-50
    y[i] = v;
  }
}
__global__ void add_kernel(const float* __restrict__ a, const float* __restrict__ b,
                           float* __restrict__ y, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    y[i] = a[i] + b[i];
}
__global__ void add_channel_vec_kernel(const float* __restrict__ x, const float* __restrict__ v,
                                       float* __restrict__ y, long long total, int C) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x)
    y[i] = x[i] + v[i % C];
}

// ------------------------------------------------------------------------------------------
// direct convolution: thread per (pixel, cout), cout fastest
// ------------------------------------------------------------------------------------------
__global__ void conv_direct_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                   const float* __restrict__ bias, float* __restrict__ out, int N,
                                   int H, int W, int Cin, int Cout, int k, int pad, int stride,
                                   int Ho, int Wo) {
  const long long total = (long long)N * Ho * Wo * Cout;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int co = (int)(i % Cout);
    long long t = i / Cout;
    int wo = (int)(t % Wo);
    t /= Wo;
    int ho = (int)(t % Ho);
    int n = (int)(t / Ho);
    float acc = bias ? bias[co] : 0.f;
    const float* wrow = w + (long long)co * k * k * Cin;
    for (int ky = 0; ky < k; ++ky) {
      int hi = ho * stride + ky - pad;
      if (hi < 0 || hi >= H) continue;
      for (int kx = 0; kx < k; ++kx) {
        int wi = wo * stride + kx - pad;
        if (wi < 0 || wi >= W) continue;
        const float* xp = x + (((long long)n * H + hi) * W + wi) * Cin;
        const float* wp = wrow + (ky * k + kx) * Cin;
        for (int ci = 0; ci < Cin; ++ci) acc = fmaf(xp[ci], wp[ci], acc);
      }
    }
    out[i] = acc;
  }
}

// Small-K direct convolution (k*k*Cin <= 64: the 4-channel input convs, diffusion.mojo:177,
// vae.mojo:195): a block stages the zero-padded input patches of 32 output pixels in shared
// memory; thread t owns output channel t (and t + 256, ...), keeps its k*k*Cin weights in
// registers and walks the pixels: patch reads are shared-memory broadcasts, output writes are
// coalesced along Cout.
constexpr int CS_PIX = 16, CS_MAXK = 64;
template <int KPAD>  // k*k*Cin rounded up to a multiple of 4 (36 for the 4-channel 3x3 input convs), <= CS_MAXK
__global__ void __launch_bounds__(512)
conv_smallk_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                   float* __restrict__ out, int N, int H, int W, int Cin, int Cout, int k, int pad, int stride,
                   int Ho, int Wo) {
  pdl_wait();
  pdl_launch_dependents();
  __shared__ __align__(16) float patch[CS_PIX][KPAD];
  const int K = k * k * Cin;
  const long long total_px = (long long)N * Ho * Wo;
  const long long px0 = (long long)blockIdx.x * CS_PIX;
  for (int i = threadIdx.x; i < CS_PIX * KPAD; i += blockDim.x) {
    const int pp = i / KPAD, kk = i - pp * KPAD;
    const long long px = px0 + pp;
    float v = 0.f;
    if (px < total_px && kk < K) {
      const int ci = kk % Cin, tap = kk / Cin;
      const int wo = (int)(px % Wo);
      long long t = px / Wo;
      const int ho = (int)(t % Ho), n = (int)(t / Ho);
      const int hi = ho * stride + tap / k - pad, wi = wo * stride + tap % k - pad;
      if (hi >= 0 && hi < H && wi >= 0 && wi < W) v = x[(((long long)n * H + hi) * W + wi) * Cin + ci];
    }
    patch[pp][kk] = v;
  }
  __syncthreads();
  for (int co = threadIdx.x; co < Cout; co += blockDim.x) {
    float wr[KPAD];
#pragma unroll
    for (int kk = 0; kk < KPAD; ++kk) wr[kk] = kk < K ? w[(long long)co * K + kk] : 0.f;
    const float b = bias ? bias[co] : 0.f;
#pragma unroll 1
    for (int pp = 0; pp < CS_PIX; pp += 4) {  // four pixels in flight: independent accumulation chains
      float acc[4] = {b, b, b, b};
#pragma unroll
      for (int kk = 0; kk < KPAD; kk += 4) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 pv = *reinterpret_cast<const float4*>(&patch[pp + j][kk]);  // shared-memory broadcast
          acc[j] = fmaf(pv.x, wr[kk], acc[j]);
          acc[j] = fmaf(pv.y, wr[kk + 1], acc[j]);
          acc[j] = fmaf(pv.z, wr[kk + 2], acc[j]);
          acc[j] = fmaf(pv.w, wr[kk + 3], acc[j]);
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (px0 + pp + j < total_px) out[(px0 + pp + j) * Cout + co] = acc[j];
    }
  }
}

// ------------------------------------------------------------------------------------------
// GEMV: one warp per output element (row r, column n)
// ------------------------------------------------------------------------------------------
__global__ void gemv_kernel(const float* __restrict__ x, int K, const float* __restrict__ Wt,
                            const float* __restrict__ bias, const float* __restrict__ bias2,
                            float* __restrict__ y, int N, int silu_in, int silu_out) {
  pdl_wait();
  pdl_launch_dependents();
  const int warps = blockDim.x >> 5;
  const int n = blockIdx.x * warps + (threadIdx.x >> 5);
  const int r = blockIdx.y;
  if (n >= N) return;
  const int lane = threadIdx.x & 31;
  const float* xr = x + (long long)r * K;
  const float* wr = Wt + (long long)n * K;
  float acc = 0.f;
  if ((K & 3) == 0) {
    for (int k = lane * 4; k < K; k += 128) {
      float4 a = *reinterpret_cast<const float4*>(xr + k);
      float4 b = *reinterpret_cast<const float4*>(wr + k);
      if (silu_in) {
        a.x = a.x / (1.0f + expf(-a.x));
        a.y = a.y / (1.0f + expf(-a.y));
        a.z = a.z / (1.0f + expf(-a.z));
        a.w = a.w / (1.0f + expf(-a.w));
      }
      acc = fmaf(a.x, b.x, acc);
      acc = fmaf(a.y, b.y, acc);
      acc = fmaf(a.z, b.z, acc);
      acc = fmaf(a.w, b.w, acc);
    }
  } else {
    for (int k = lane; k < K; k += 32) {
      float a = xr[k];
      if (silu_in) a = a / (1.0f + expf(-a));
      acc = fmaf(a, wr[k], acc);
    }
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffff, acc, o);
  if (lane == 0) {
    if (bias) acc += bias[n];
    if (bias2) acc += bias2[n];
    if (silu_out) acc = acc / (1.0f + expf(-acc));
    y[(long long)r * N + n] = acc;
  }
}

// Several GEMVs over the same input vector(s) in one launch: output row n of the concatenation belongs to the
// segment whose [row0, row0 + N) holds it.  Used for the nine per-ResBlock Linear(SiLU(t)) (diffusion.mojo:61-62),
// which depend only on the time embedding.
__global__ void gemv_multi_kernel(const float* __restrict__ x, int K, const GemvMulti m, int silu_in) {
  pdl_wait();
  pdl_launch_dependents();
  const int warps = blockDim.x >> 5;
  const int ng = blockIdx.x * warps + (threadIdx.x >> 5);
  const int r = blockIdx.y;
  if (ng >= m.total) return;
  int s = 0;
#pragma unroll 1
  while (s + 1 < m.nseg && ng >= m.seg[s + 1].row0) ++s;
  const GemvSeg sg = m.seg[s];
  const int n = ng - sg.row0;
  const int lane = threadIdx.x & 31;
  const float* xr = x + (long long)r * K;
  const float* wr = sg.W + (long long)n * K;
  float acc = 0.f;
  for (int k = lane * 4; k < K; k += 128) {
    float4 a = *reinterpret_cast<const float4*>(xr + k);
    const float4 b = *reinterpret_cast<const float4*>(wr + k);
    if (silu_in) {
      a.x = a.x / (1.0f + expf(-a.x));
      a.y = a.y / (1.0f + expf(-a.y));
      a.z = a.z / (1.0f + expf(-a.z));
      a.w = a.w / (1.0f + expf(-a.w));
    }
    acc = fmaf(a.x, b.x, acc);
    acc = fmaf(a.y, b.y, acc);
    acc = fmaf(a.z, b.z, acc);
    acc = fmaf(a.w, b.w, acc);
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffff, acc, o);
  if (lane == 0) {
    if (sg.bias) acc += sg.bias[n];
    if (sg.bias2) acc += sg.bias2[n];
    sg.y[(long long)r * sg.N + n] = acc;
  }
}

// ------------------------------------------------------------------------------------------
// softmax (unfused attention path)
// ------------------------------------------------------------------------------------------
// axis 0: per column j, over rows.  block (32, 8): 32 columns, 8 row lanes.
// `causal`: entry (row i = query, column j = key) is masked (= -inf before the softmax) when j > i
// (Self_Attention.forward with causal_mask, helpers/attention.mojo:48-56, with the standard triu(1)).
__global__ void softmax_colstats_kernel(const float* __restrict__ S, int R, int Cc, int ld,
                                        float scale, float2* __restrict__ st, int causal) {
  __shared__ float sm_m[8][33], sm_l[8][33];
  const int b = blockIdx.y;
  const int j = blockIdx.x * 32 + threadIdx.x;
  const float* base = S + (long long)b * R * ld;
  float m = -FLT_MAX, l = 0.f;
  if (j < Cc) {
    for (int i = threadIdx.y; i < R; i += 8) {
      if (causal && j > i) continue;
      float v = base[(long long)i * ld + j] * scale;
      if (v > m) {
        l = l * expf(m - v) + 1.0f;
        m = v;
      } else {
        l += expf(v - m);
      }
    }
  }
  sm_m[threadIdx.y][threadIdx.x] = m;
  sm_l[threadIdx.y][threadIdx.x] = l;
  __syncthreads();
  if (threadIdx.y == 0 && j < Cc) {
    float M = sm_m[0][threadIdx.x];
    for (int t = 1; t < 8; ++t) M = fmaxf(M, sm_m[t][threadIdx.x]);
    float L = 0.f;
    for (int t = 0; t < 8; ++t) L += sm_l[t][threadIdx.x] * expf(sm_m[t][threadIdx.x] - M);
    st[(long long)b * Cc + j] = make_float2(M, L > 0.f ? 1.0f / L : 0.f);  // a fully masked column (j >= R) yields zeros
  }
}
__global__ void softmax_colapply_kernel(float* __restrict__ S, int R, int Cc, int ld, float scale,
                                        const float2* __restrict__ st, long long total, int causal) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int j = (int)(i % Cc);
    long long t = i / Cc;  // b * R + r
    long long b = t / R;
    float2 s = st[b * Cc + j];
    float* p = S + t * ld + j;
    const int r = (int)(t - b * R);
    *p = (causal && j > r) ? 0.f : expf(*p * scale - s.x) * s.y;
  }
}
// axis 1: block per row
__global__ void softmax_row_kernel(float* __restrict__ S, int Cc_all, int ld, float scale, int R, int causal) {
  float* row = S + (long long)blockIdx.x * ld;
  const int r = blockIdx.x % R;
  const int Cc = (causal && r + 1 < Cc_all) ? r + 1 : Cc_all;  // causal: keys 0..r only
  if (causal)
    for (int j = Cc + threadIdx.x; j < Cc_all; j += blockDim.x) row[j] = 0.f;
  __shared__ float red[32];
  float m = -FLT_MAX;
  for (int j = threadIdx.x; j < Cc; j += blockDim.x) m = fmaxf(m, row[j] * scale);
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffff, m, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  m = red[0];
  for (int i = 1; i < (blockDim.x >> 5); ++i) m = fmaxf(m, red[i]);
  __syncthreads();
  float l = 0.f;
  for (int j = threadIdx.x; j < Cc; j += blockDim.x) l += expf(row[j] * scale - m);
  for (int o = 16; o > 0; o >>= 1) l += __shfl_xor_sync(0xffffffff, l, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = l;
  __syncthreads();
  l = 0.f;
  for (int i = 0; i < (blockDim.x >> 5); ++i) l += red[i];
  const float inv = 1.0f / l;
  for (int j = threadIdx.x; j < Cc; j += blockDim.x) row[j] = expf(row[j] * scale - m) * inv;
}

// ClipEmbedding.forward (clip.mojo:17-20; Embedding.forward, helpers/utils.mojo:2032-2046):
// out[t][:] = token_table[tokens[t]][:] + position[t][:]
__global__ void clip_embed_kernel(const int* __restrict__ tokens, const float4* __restrict__ table,
                                  const float4* __restrict__ pos, float4* __restrict__ out, int T, int d4, int n_vocab) {
  const long long total = (long long)T * d4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int t = (int)(i / d4), c = (int)(i - (long long)t * d4);
    int id = tokens[t];
    id = id < 0 ? 0 : (id >= n_vocab ? n_vocab - 1 : id);  // device-side token ids (tsd_clip_forward_dev) are not host-validated
    const float4 a = table[(long long)id * d4 + c], b = pos[i];
    out[i] = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
  }
}

// Holds the stream busy for `ns` nanoseconds so that the host can enqueue a whole step behind it
// (profiling passes: per-launch CUDA events then measure kernel time, not host launch gaps).
__global__ void spin_kernel(long long ns) {
  long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (;;) {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    if (t - t0 >= ns) break;
    __nanosleep(1000);
  }
}

__global__ void ddpm_step_kernel(const float* __restrict__ x, const float* __restrict__ ec,
                                 const float* __restrict__ eu, float cfg_scale,
                                 const float* __restrict__ noise, float sqrt_ab, float sqrt_1mab,
                                 float c0, float c1, float sigma, float* __restrict__ out,
                                 long long n) {
  pdl_wait();
  pdl_launch_dependents();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    float e = ec[i];
    if (eu) {
      float u = eu[i];
      e = (e - u) * cfg_scale + u;  // pipeline.mojo:117-119
    }
    float xi = x[i];
    float x0 = (xi - e * sqrt_1mab) / sqrt_ab;  // sampler.mojo:88-90
    float o = x0 * c0 + xi * c1;                // sampler.mojo:91-99
    if (noise) o += noise[i] * sigma;           // sampler.mojo:101-108
    out[i] = o;
  }
}

}  // namespace

// ------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------
cudaError_t launch_nchw_to_nhwc(const float* src, float* dst, int N, int C, int HW, cudaStream_t s) {
  return launch_transpose(src, dst, N, C, HW, 0, s);
}
cudaError_t launch_nhwc_to_nchw(const float* src, float* dst, int N, int C, int HW, cudaStream_t s) {
  return launch_transpose(src, dst, N, HW, C, 0, s);
}
cudaError_t launch_rescale_to_nchw(const float* src, float* dst, int N, int C, int HW, int rescale,
                                   cudaStream_t s) {
  return launch_transpose(src, dst, N, HW, C, rescale, s);
}
cudaError_t launch_rescale_to_nhwc(const float* src, float* dst, int N, int C, int HW, int rescale,
                                   cudaStream_t s) {
  return launch_transpose(src, dst, N, C, HW, rescale ? 2 : 0, s);
}

// Encoder.metrics_evals, vae.mojo:118-129: moments NHWC [N][HW][8] (mean = channels 0..3,
// log-variance = 4..7), noise / out NCHW [N][4][HW].
__global__ void latent_from_moments_kernel(const float* __restrict__ m, const float* __restrict__ noise,
                                           float* __restrict__ out, int N, int HW) {
  pdl_wait();
  pdl_launch_dependents();
  const long long total = (long long)N * 4 * HW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int px = (int)(i % HW);
    const long long t = i / HW;
    const int ch = (int)(t % 4);
    const long long n = t / 4;
    const float* mp = m + (n * HW + px) * 8;
    const float mean = mp[ch];
    const float lv = fminf(fmaxf(mp[4 + ch], -30.0f), 20.0f);
    const float sd = sqrtf(expf(lv));
    out[i] = (mean + noise[i] * sd) * 0.18215f;
  }
}
cudaError_t launch_latent_from_moments(const float* m, const float* noise, float* out, int N, int HW,
                                       cudaStream_t s) {
  const long long total = (long long)N * 4 * HW;
  return launch_pdl(latent_from_moments_kernel, dim3(grid_for(total, 256)), dim3(256), 0, s, m, noise, out, N, HW);
}
cudaError_t launch_transpose_ld(const float* src, float* dst, int B, int R, int Cc, int ld_out,
                                cudaStream_t s) {
  dim3 grid((Cc + 31) / 32, (R + 31) / 32, B), block(32, 8);
  transpose_ld_kernel<<<grid, block, 0, s>>>(src, dst, R, Cc, ld_out);
  return cudaGetLastError();
}
cudaError_t launch_oihw_to_ohwi(const float* src, float* dst, int O, int I, int KK, cudaStream_t s) {
  long long total = (long long)O * I * KK;
  oihw_to_ohwi_kernel<<<grid_for(total, 256), 256, 0, s>>>(src, dst, O, I, KK);
  return cudaGetLastError();
}
cudaError_t launch_concat_channels(const float* a, int Ca, const float* b, int Cb, float* out,
                                   long long pixels, cudaStream_t s) {
  long long total = pixels * (Ca + Cb);
  { cudaError_t e_ = launch_pdl(concat_kernel, dim3(grid_for(total, 256)), dim3(256), 0, s, a, Ca, b, Cb, out, pixels); if (e_ != cudaSuccess) return e_; }
  return cudaGetLastError();
}
cudaError_t launch_upsample2x(const float* x, float* y, int N, int H, int W, int C, cudaStream_t s) {
  if (C % 4) return cudaErrorInvalidValue;
  long long total = (long long)N * 4 * H * W * (C / 4);
  { cudaError_t e_ = launch_pdl(upsample2x_kernel, dim3(grid_for(total, 256)), dim3(256), 0, s, reinterpret_cast<const float4*>(x),
                                                         reinterpret_cast<float4*>(y), N, H, W, C / 4); if (e_ != cudaSuccess) return e_; }
  return cudaGetLastError();
}
cudaError_t launch_upsample2x_planar(const float* x, float* y, int C, int H, int W, cudaStream_t s) {
  long long total = (long long)C * 4 * H * W;
  upsample2x_planar_kernel<<<grid_for(total, 256), 256, 0, s>>>(x, y, C, H, W);
  return cudaGetLastError();
}
cudaError_t launch_im2col3x3(const float* x, float* col, int N, int H, int W, int C, int stride,
                             int pad_lo, int Ho, int Wo, cudaStream_t s) {
  if (C % 4) return cudaErrorInvalidValue;
  long long total = (long long)N * Ho * Wo * 9 * (C / 4);
  { cudaError_t e_ = launch_pdl(im2col3x3_kernel, dim3(grid_for(total, 256)), dim3(256), 0, s, reinterpret_cast<const float4*>(x),
                                                        reinterpret_cast<float4*>(col), N, H, W,
                                                        C / 4, stride, pad_lo, Ho, Wo); if (e_ != cudaSuccess) return e_; }
  return cudaGetLastError();
}

int group_stats_blocks(int N, long long pixels, int C) {
  // blocks per image: a function of the image shape only (never of the batch), so that one
  // image's statistics are bit-identical whatever batch it is evaluated in
  (void)N;
  const int C4 = C / 4;
  const int ppl = C4 < GS_THREADS ? GS_THREADS / C4 : 1;
  long long nb = pixels / (8LL * ppl);
  if (nb > 64) nb = 64;
  if (nb < 1) nb = 1;
  return (int)nb;
}

size_t group_stats_scratch_bytes(int N, long long pixels, int C, int G) {
  if (C % 4 == 0 && C / 4 <= GS_THREADS * GS_MAXQ)
    return sizeof(double2) * (size_t)N * group_stats_blocks(N, pixels, C) * G;
  return sizeof(double) * 2 * (size_t)N * G;
}

cudaError_t launch_group_stats(const float* x, int N, long long pixels, int C, int G, float eps,
                               void* scratch, unsigned int* ticket, float2* stats, cudaStream_t s) {
  if (G <= 0 || C % G) return cudaErrorInvalidValue;
  const int cpg = C / G;
  const int NG = N * G;
  if (C % 4 == 0 && C / 4 <= GS_THREADS * GS_MAXQ) {
    const int C4 = C / 4;
    const int ppl = C4 < GS_THREADS ? GS_THREADS / C4 : 1;
    const int nb = group_stats_blocks(N, pixels, C);
    const int slab = (int)((pixels + nb - 1) / nb);
    const size_t smem = (size_t)ppl * 2 * C * sizeof(float);
    {
      cudaError_t e = optin_dyn_smem(reinterpret_cast<const void*>(group_stats_vec_kernel), 64 * 1024, nullptr);
      if (e != cudaSuccess) return e;
    }
    dim3 grid((unsigned)nb, N);
    group_stats_vec_kernel<<<grid, GS_THREADS, smem, s>>>(reinterpret_cast<const float4*>(x), pixels, C4, G, cpg, slab,
                                                          reinterpret_cast<double2*>(scratch), ticket,
                                                          (double)pixels * cpg, eps, stats);
    return cudaGetLastError();
  }
  dim3 grid(G, N);
  double* accum = reinterpret_cast<double*>(scratch);
  group_stats_general_kernel<<<grid, 256, 0, s>>>(x, pixels, C, G, cpg, accum);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  group_stats_finalize_kernel<<<(NG + 127) / 128, 128, 0, s>>>(accum, NG, (double)pixels * cpg, eps, stats);
  return cudaGetLastError();
}

bool norm_fused_supported(int C) { return C % 4 == 0 && C / 4 <= GS_THREADS * GS_MAXQ; }

static int norm_fused_slabs(long long pixels, int C) {
  const int C4 = C / 4;
  const int ppl = C4 < GS_THREADS ? GS_THREADS / C4 : 1;
  long long s = pixels / (4LL * ppl);
  if (s > 128) s = 128;
  if (s < 1) s = 1;
  return (int)s;
}

size_t norm_fused_scratch_bytes(int N, long long pixels, int C, int G) {
  return sizeof(double2) * (size_t)N * norm_fused_slabs(pixels, C) * G;
}

cudaError_t launch_norm_fused(const float* x, float* y, int N, long long pixels, int C, int G, float eps,
                              const float* gamma, const float* beta, float gamma_scalar, int silu, int round_tf32,
                              void* scratch, unsigned int* barrier_words, int sm_count, cudaStream_t s) {
  if (G <= 0 || C % G || !norm_fused_supported(C)) return cudaErrorInvalidValue;
  const int C4 = C / 4, cpg = C / G;
  const int ppl = C4 < GS_THREADS ? GS_THREADS / C4 : 1;
  const int slabs = norm_fused_slabs(pixels, C);
  const int slab = (int)((pixels + slabs - 1) / slabs);
  const int items = N * slabs;
  int grid = items < sm_count ? items : sm_count;
  size_t smem = (size_t)ppl * 2 * C * sizeof(float);
  if (smem < (size_t)G * sizeof(float2)) smem = (size_t)G * sizeof(float2);
  {
    cudaError_t e = optin_dyn_smem(reinterpret_cast<const void*>(norm_fused_kernel), 64 * 1024, nullptr);
    if (e != cudaSuccess) return e;
  }
  {
    cudaError_t e = grid_coresident(reinterpret_cast<const void*>(norm_fused_kernel), GS_THREADS, smem, grid);
    if (e != cudaSuccess) return e;
  }
  { cudaError_t e_ = launch_pdl(norm_fused_kernel, dim3(grid), dim3(GS_THREADS), smem, s, reinterpret_cast<const float4*>(x), reinterpret_cast<float4*>(y),
                                                   pixels, C4, G, cpg, slabs, slab, items,
                                                   reinterpret_cast<double2*>(scratch), barrier_words,
                                                   (double)pixels * cpg, eps, gamma, beta, gamma_scalar, silu,
                                                   round_tf32); if (e_ != cudaSuccess) return e_; }
  return cudaGetLastError();
}


// ---- fused norm v2 ----
struct NormFused2Plan {
  int ppl, slab, slabs_per_img, subs_per_img, nq, iters;
  bool cache;
};
static NormFused2Plan norm_fused2_plan(int N, long long pixels, int C, int sm_count) {
  NormFused2Plan pl{};
  const int C4 = C / 4;
  pl.ppl = C4 < NF_THREADS ? NF_THREADS / C4 : 1;
  pl.nq = (C4 + NF_THREADS - 1) / NF_THREADS;
  int max_blocks = 2 * sm_count;
  int spi = max_blocks / N;
  if (spi < 1) spi = 1;
  const long long groups = (pixels + pl.ppl - 1) / pl.ppl;  // pixel groups in flight per block pass
  if (spi > groups) spi = (int)groups;
  long long slab = (pixels + spi - 1) / spi;
  slab = (slab + pl.ppl - 1) / pl.ppl * pl.ppl;
  pl.slab = (int)slab;
  pl.slabs_per_img = (int)((pixels + slab - 1) / slab);
  pl.subs_per_img = (pl.slabs_per_img + NF_SUB - 1) / NF_SUB;
  pl.iters = (int)(slab / pl.ppl);
  const int maxit = pl.nq == 1 ? 16 : (pl.nq == 2 ? 8 : 5);
  pl.cache = pl.iters <= maxit;
  return pl;
}
// the kernel instantiation and dynamic shared memory a plan launches with
static const void* norm_fused2_func(const NormFused2Plan& pl);
static size_t norm_fused2_smem(const NormFused2Plan& pl, int C, int G) {
  size_t smem = (size_t)pl.ppl * 2 * C * sizeof(float);
  if (smem < (size_t)G * sizeof(float2)) smem = (size_t)G * sizeof(float2);
  return smem;
}
bool norm_fused2_supported(int N, long long pixels, int C, int G, int sm_count) {
  if (C % 4 || C / 4 > 3 * NF_THREADS || G <= 0 || C % G || pixels >= (1 << 30)) return false;
  if (N > 2 * sm_count) return false;
  const NormFused2Plan pl = norm_fused2_plan(N, pixels, C, sm_count);
  if (!(pl.subs_per_img <= 32 && (long long)N * pl.subs_per_img + 1 <= kNormBarrierCounters)) return false;
  // the grid barrier needs every block resident at once: ask the runtime instead of assuming two blocks per SM
  const void* fn = norm_fused2_func(pl);
  if (optin_dyn_smem(fn, 64 * 1024, nullptr) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  if (grid_coresident(fn, NF_THREADS, norm_fused2_smem(pl, C, G), (long long)N * pl.slabs_per_img) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return true;
}
size_t norm_fused2_scratch_bytes(int N, long long pixels, int C, int G, int sm_count) {
  const NormFused2Plan pl = norm_fused2_plan(N, pixels, C, sm_count);
  return sizeof(float2) * (size_t)N * G * ((size_t)pl.slabs_per_img + pl.subs_per_img) + 256;
}
cudaError_t launch_norm_fused2(const NormFused2Src& src, float* y, int N, long long pixels, int C, int G, float eps,
                               const float* gamma, const float* beta, float gamma_scalar, int silu, int round_tf32,
                               void* scratch, unsigned int* barrier_words, int sm_count, cudaStream_t s) {
  if (!norm_fused2_supported(N, pixels, C, G, sm_count)) return cudaErrorInvalidValue;
  const NormFused2Plan pl = norm_fused2_plan(N, pixels, C, sm_count);
  NormFused2Params p{};
  p.x = reinterpret_cast<const float4*>(src.x);
  p.splits = src.splits > 1 ? src.splits : 1;
  p.split_stride4 = src.split_stride / 4;
  p.ldx4 = src.splits > 1 ? src.ldx / 4 : C / 4;
  p.bias = src.bias;
  p.bias_img_stride = src.bias_img_stride;
  p.residual = reinterpret_cast<const float4*>(src.residual);
  p.raw = reinterpret_cast<float4*>(src.raw);
  p.x2 = reinterpret_cast<const float4*>(src.x2);
  p.c4a = src.c_a / 4;
  if (src.x2 != nullptr && (src.splits > 1 || src.c_a <= 0 || src.c_a >= C || src.c_a % 4)) return cudaErrorInvalidValue;
  p.y = reinterpret_cast<float4*>(y);
  p.pixels = (int)pixels;
  p.C4 = C / 4;
  p.G = G;
  p.cpg = C / G;
  p.slabs_per_img = pl.slabs_per_img;
  p.slab = pl.slab;
  p.subs_per_img = pl.subs_per_img;
  p.part = reinterpret_cast<float2*>(scratch);
  p.part2 = p.part + (size_t)N * pl.slabs_per_img * G;
  p.bar = barrier_words;
  p.inv_count = (float)(1.0 / ((double)pixels * (C / G)));
  p.eps = eps;
  p.gamma = gamma;
  p.beta = beta;
  p.gamma_scalar = gamma_scalar;
  p.silu = silu;
  p.round = round_tf32;
  {
    static int tr = -1;
    if (tr < 0) { const char* v = getenv("TSD_NORM_TRACE"); tr = v ? atoi(v) : 0; }
    p.trace = tr;
  }
  bool cache = pl.cache;
  if ((p.splits > 1 || p.x2 != nullptr) && !cache && p.raw == nullptr) return cudaErrorInvalidValue;  // re-read mode needs the raw tensor
  const int grid = N * pl.slabs_per_img;
  const size_t smem = norm_fused2_smem(pl, C, G);  // opt-in limit and co-residency were established by norm_fused2_supported
#define NF2_LAUNCH(NQ, MAXIT, CACHE)                                                                        \
  do {                                                                                                      \
    { cudaError_t e_ = launch_pdl(norm_fused2_kernel<NQ, MAXIT, CACHE>, dim3(grid), dim3(NF_THREADS), smem, s, p); if (e_ != cudaSuccess) return e_; } \
  } while (0)
  if (pl.nq == 1) { if (cache) NF2_LAUNCH(1, 16, true); else NF2_LAUNCH(1, 16, false); }
  else if (pl.nq == 2) { if (cache) NF2_LAUNCH(2, 8, true); else NF2_LAUNCH(2, 8, false); }
  else { if (cache) NF2_LAUNCH(3, 5, true); else NF2_LAUNCH(3, 5, false); }
#undef NF2_LAUNCH
  return cudaGetLastError();
}

static const void* norm_fused2_func(const NormFused2Plan& pl) {
  if (pl.nq == 1) return pl.cache ? reinterpret_cast<const void*>(norm_fused2_kernel<1, 16, true>) : reinterpret_cast<const void*>(norm_fused2_kernel<1, 16, false>);
  if (pl.nq == 2) return pl.cache ? reinterpret_cast<const void*>(norm_fused2_kernel<2, 8, true>) : reinterpret_cast<const void*>(norm_fused2_kernel<2, 8, false>);
  return pl.cache ? reinterpret_cast<const void*>(norm_fused2_kernel<3, 5, true>) : reinterpret_cast<const void*>(norm_fused2_kernel<3, 5, false>);
}


// ---- cluster GroupNorm ----
struct NormClusterPlan {
  int v, hc, ppl, cs, ppc;
  size_t smem;
  bool ok;
};
static int norm_cluster_target_ctas() {  // lab: TSD_NORM_CLUSTER_CTAS overrides the CTA count a plan aims for
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("TSD_NORM_CLUSTER_CTAS");
    v = e ? atoi(e) : 256;
    if (v < 1) v = 256;
  }
  return v;
}
static NormClusterPlan norm_cluster_plan(int N, long long pixels, int C, int G) {
  NormClusterPlan pl{};
  if (G <= 0 || C % G || pixels <= 0 || pixels >= (1 << 30)) return pl;
  const int cpg = C / G;
  if (cpg % 2 || C % 4) return pl;
  pl.v = cpg % 4 == 0 ? 4 : 2;
  pl.hc = cpg / pl.v;
  if (pl.hc > NC_THREADS) return pl;
  pl.ppl = NC_THREADS / pl.hc;
  const long long clusters = (long long)N * G;
  if (clusters > (1 << 20)) return pl;
  // smallest cluster that fills the device (>= 256 CTAs) and whose slab fits shared memory; <= 8 (portable size)
  const size_t limit = 200 * 1024;
  for (int cs = 1; cs <= 8; cs *= 2) {
    const long long ppc = (pixels + cs - 1) / cs;
    const size_t smem = (size_t)((ppc + pl.ppl - 1) / pl.ppl * pl.ppl) * cpg * sizeof(float);
    if (smem > limit) continue;
    pl.cs = cs;
    pl.ppc = (int)ppc;
    pl.smem = smem;
    pl.ok = true;
    if (clusters * cs >= norm_cluster_target_ctas() || ppc <= 2 * pl.ppl) break;
  }
  if (pl.ok && clusters * pl.cs < 64) pl.ok = false;  // too few CTAs to be worth it (LayerNorm of one image)
  return pl;
}
bool norm_cluster_supported(int N, long long pixels, int C, int G) {
  const NormClusterPlan pl = norm_cluster_plan(N, pixels, C, G);
  if (!pl.ok) return false;
  const void* fn = pl.v == 4 ? reinterpret_cast<const void*>(norm_cluster_kernel<4>) : reinterpret_cast<const void*>(norm_cluster_kernel<2>);
  if (optin_dyn_smem(fn, 200 * 1024, nullptr) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return true;
}
cudaError_t launch_norm_cluster(const NormFused2Src& src, float* y, int N, long long pixels, int C, int G, float eps,
                                const float* gamma, const float* beta, float gamma_scalar, int silu, int round_tf32,
                                cudaStream_t s) {
  if (!norm_cluster_supported(N, pixels, C, G)) return cudaErrorInvalidValue;
  const NormClusterPlan pl = norm_cluster_plan(N, pixels, C, G);
  NormClusterParams p{};
  p.x = src.x;
  p.splits = src.splits > 1 ? src.splits : 1;
  p.split_stride = src.split_stride;
  p.ldx = src.splits > 1 ? src.ldx : C;
  p.bias = src.bias;
  p.bias_img_stride = src.bias_img_stride;
  p.residual = src.residual;
  p.raw = src.raw;
  p.x2 = src.x2;
  p.c_a = src.c_a;
  if (src.x2 != nullptr && (src.splits > 1 || src.c_a <= 0 || src.c_a >= C || src.c_a % pl.v || (C - src.c_a) % pl.v)) return cudaErrorInvalidValue;
  if (p.splits > 1 && (src.ldx % pl.v || src.split_stride % pl.v)) return cudaErrorInvalidValue;
  p.y = y;
  p.pixels = (int)pixels;
  p.C = C;
  p.G = G;
  p.cpg = C / G;
  p.cs = pl.cs;
  p.ppc = pl.ppc;
  p.hc = pl.hc;
  p.ppl = pl.ppl;
  p.inv_count = (float)(1.0 / ((double)pixels * (C / G)));
  p.eps = eps;
  p.gamma = gamma;
  p.beta = beta;
  p.gamma_scalar = gamma_scalar;
  p.silu = silu;
  p.round = round_tf32;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)((long long)N * G * pl.cs));
  cfg.blockDim = dim3(NC_THREADS);
  cfg.dynamicSmemBytes = pl.smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  attr[1].id = cudaLaunchAttributeClusterDimension;
  attr[1].val.clusterDim.x = (unsigned)pl.cs;
  attr[1].val.clusterDim.y = 1;
  attr[1].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  cudaError_t e = pl.v == 4 ? cudaLaunchKernelEx(&cfg, norm_cluster_kernel<4>, p) : cudaLaunchKernelEx(&cfg, norm_cluster_kernel<2>, p);
  if (e != cudaSuccess) return e;
  return cudaGetLastError();
}

static int norm_trace_env() {
  static int tr = -1;
  if (tr < 0) { const char* v = getenv("TSD_NORM_TRACE"); tr = v ? atoi(v) : 0; }
  return tr;
}

// channel segments of a many-group norm: each block folds at most ~64 groups (0: not expressible)
static int norm_apply_partial_segments(int C, int G) {
  if (G <= 64) return 1;
  const int C4 = C / 4, cpg = C / G;
  for (int nseg = (G + 63) / 64; nseg <= C4; ++nseg) {
    if (C4 % nseg) continue;
    const int ch = (C4 / nseg) * 4;
    if (ch % cpg == 0 && ch / cpg <= 512) return nseg;
  }
  return 0;
}
bool norm_apply_partial_supported(int C, int G) {
  return C % 4 == 0 && C / 4 <= GS_MAXQ * GS_THREADS && G > 0 && C % G == 0 && C < 65536 && norm_apply_partial_segments(C, G) > 0 &&
         (G <= 512 || norm_apply_partial_segments(C, G) > 1);
}

cudaError_t launch_norm_apply_partial(const float* x, float* y, const NormStatsReq& req, int N, long long pixels_ll,
                                      const float* gamma, const float* beta, float gamma_scalar, int silu,
                                      int round_tf32, cudaStream_t s) {
  if (!norm_apply_partial_supported(req.C, req.G) || pixels_ll >= (1 << 30)) return cudaErrorInvalidValue;
  const int C4 = req.C / 4, pixels = (int)pixels_ll;
  const int nseg = norm_apply_partial_segments(req.C, req.G);
  const int c4_seg = C4 / nseg;
  const int ppl = c4_seg < GS_THREADS ? GS_THREADS / c4_seg : 1;
  // blocks over the whole batch (each folds the partial statistics of its image first, a fixed cost);
  // a slab is a multiple of the pixels in flight
  static int target = -1;
  if (target < 0) {
    const char* v = getenv("TSD_NORM_BLOCKS");  // lab override
    target = v ? atoi(v) : 2 * 148;  // two blocks per SM (measured best: 6.7 us for 64 x 64 x 320 against 9.6 at four)
    if (target < 1) target = 2 * 148;
  }
  // small tensors (the UNet's <= 5 MB activations) are latency chains: two blocks per SM; the decoder's tensors of tens
  // to thousands of MB need loads in flight to fill HBM: up to eight blocks per SM
  const long long bytes = 4ll * N * pixels_ll * req.C;
  int tgt = target;
  if (bytes > (256ll << 20)) tgt = 8 * target / 2;
  else if (bytes > (32ll << 20)) tgt = 4 * target / 2;
  int slabs = (tgt + N * nseg - 1) / (N * nseg);
  int slab = (pixels + slabs - 1) / slabs;
  slab = (slab + ppl - 1) / ppl * ppl;
  if (slab < ppl) slab = ppl;
  slabs = (pixels + slab - 1) / slab;
  const int nq = (c4_seg + GS_THREADS - 1) / GS_THREADS;
#define TSD_NAP(NQ)                                                                                                        \
  launch_pdl(norm_apply_partial_kernel<NQ>, dim3(slabs, N, nseg), dim3(GS_THREADS), 0, s, reinterpret_cast<const float4*>(x), \
             reinterpret_cast<float4*>(y), req, gamma, beta, gamma_scalar, pixels, C4, slab, silu, round_tf32,             \
             norm_trace_env(), c4_seg)
  const cudaError_t e_ = nq <= 1 ? TSD_NAP(1) : (nq == 2 ? TSD_NAP(2) : TSD_NAP(3));
#undef TSD_NAP
  if (e_ != cudaSuccess) return e_;
  return cudaGetLastError();
}

cudaError_t launch_norm_apply(const float* x, const float2* stats, const float* gamma,
                              const float* beta, float gamma_scalar, float* y, int N, int H, int W,
                              int C, int G, int silu, int upsample2x, int round_tf32,
                              cudaStream_t s) {
  if (G <= 0 || C % G) return cudaErrorInvalidValue;
  const int cpg = C / G;
  const long long outpix = (long long)N * H * W * (upsample2x ? 4 : 1);
  if (C % 4 == 0) {
    long long total = outpix * (C / 4);
    { cudaError_t e_ = launch_pdl(norm_apply_kernel<4>, dim3(grid_for(total, 256)), dim3(256), 0, s, x, stats, gamma, beta, gamma_scalar, y,
                                                              N, H, W, C, G, cpg, silu, upsample2x,
                                                              round_tf32); if (e_ != cudaSuccess) return e_; }
  } else {
    long long total = outpix * C;
    { cudaError_t e_ = launch_pdl(norm_apply_kernel<1>, dim3(grid_for(total, 256)), dim3(256), 0, s, x, stats, gamma, beta, gamma_scalar, y,
                                                              N, H, W, C, G, cpg, silu, upsample2x,
                                                              round_tf32); if (e_ != cudaSuccess) return e_; }
  }
  return cudaGetLastError();
}

cudaError_t launch_fill_uniform(float* p, long long n, uint64_t seed, float lo, float hi, cudaStream_t s) {
  fill_uniform_kernel<<<grid_for(n, 256), 256, 0, s>>>(p, n, seed, lo, hi);
  return cudaGetLastError();
}
cudaError_t launch_unary(const float* x, float* y, long long n, int op, float scalar, cudaStream_t s) {
  unary_kernel<<<grid_for(n, 256), 256, 0, s>>>(x, y, n, op, scalar);
  return cudaGetLastError();
}
cudaError_t launch_add(const float* a, const float* b, float* y, long long n, cudaStream_t s) {
  add_kernel<<<grid_for(n, 256), 256, 0, s>>>(a, b, y, n);
  return cudaGetLastError();
}
cudaError_t launch_add_channel_vec(const float* x, const float* v, float* y, long long pixels, int C,
                                   cudaStream_t s) {
  add_channel_vec_kernel<<<grid_for(pixels * C, 256), 256, 0, s>>>(x, v, y, pixels * C, C);
  return cudaGetLastError();
}

cudaError_t launch_conv_direct(const float* x, const float* w, const float* bias, float* out, int N,
                               int H, int W, int Cin, int Cout, int k, int pad, int stride, int Ho,
                               int Wo, cudaStream_t s) {
  if (k * k * Cin <= CS_MAXK && Cout >= 32) {
    const long long total_px = (long long)N * Ho * Wo;
    const int K = k * k * Cin;
    int threads = (Cout + 31) / 32 * 32;
    if (threads > 512) threads = 512;
    const dim3 grid((unsigned)((total_px + CS_PIX - 1) / CS_PIX));
    if (K <= 36)
      return launch_pdl(conv_smallk_kernel<36>, grid, dim3(threads), 0, s, x, w, bias, out, N, H, W, Cin, Cout, k, pad, stride, Ho, Wo);
    return launch_pdl(conv_smallk_kernel<CS_MAXK>, grid, dim3(threads), 0, s, x, w, bias, out, N, H, W, Cin, Cout, k, pad, stride, Ho, Wo);
  }
  long long total = (long long)N * Ho * Wo * Cout;
  conv_direct_kernel<<<grid_for(total, 128, 16), 128, 0, s>>>(x, w, bias, out, N, H, W, Cin, Cout, k,
                                                              pad, stride, Ho, Wo);
  return cudaGetLastError();
}

cudaError_t launch_gemv_multi(const float* x, int rows, int K, const GemvMulti& m, int silu_in, cudaStream_t s) {
  if (K % 4 || m.nseg <= 0 || m.nseg > GemvMulti::kMax) return cudaErrorInvalidValue;
  const int warps = 8;
  dim3 grid((m.total + warps - 1) / warps, rows);
  return launch_pdl(gemv_multi_kernel, grid, dim3(warps * 32), 0, s, x, K, m, silu_in);
}
cudaError_t launch_gemv(const float* x, int rows, int K, const float* Wt, const float* bias,
                        const float* bias2, float* y, int N, int silu_in, int silu_out,
                        cudaStream_t s) {
  const int warps = 8;
  dim3 grid((N + warps - 1) / warps, rows);
  return launch_pdl(gemv_kernel, grid, dim3(warps * 32), 0, s, x, K, Wt, bias, bias2, y, N, silu_in, silu_out);
}

cudaError_t launch_softmax(float* S, int B, int R, int Cc, int ld, int axis, float scale,
                           float* col_scratch, cudaStream_t s, int causal) {
  if (axis == 0) {
    dim3 grid((Cc + 31) / 32, B), block(32, 8);
    float2* st = reinterpret_cast<float2*>(col_scratch);
    softmax_colstats_kernel<<<grid, block, 0, s>>>(S, R, Cc, ld, scale, st, causal);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    long long total = (long long)B * R * Cc;
    softmax_colapply_kernel<<<grid_for(total, 256), 256, 0, s>>>(S, R, Cc, ld, scale, st, total, causal);
  } else {
    softmax_row_kernel<<<B * R, 256, 0, s>>>(S, Cc, ld, scale, R, causal);
  }
  return cudaGetLastError();
}

cudaError_t launch_clip_embed(const int* tokens, const float* table, const float* pos, float* out, int T, int d,
                              int n_vocab, cudaStream_t s) {
  if (d % 4) return cudaErrorInvalidValue;
  clip_embed_kernel<<<grid_for((long long)T * (d / 4), 256), 256, 0, s>>>(tokens, reinterpret_cast<const float4*>(table),
                                                                        reinterpret_cast<const float4*>(pos),
                                                                        reinterpret_cast<float4*>(out), T, d / 4, n_vocab);
  return cudaGetLastError();
}

cudaError_t launch_spin(long long ns, cudaStream_t s) {
  spin_kernel<<<1, 1, 0, s>>>(ns);
  return cudaGetLastError();
}

cudaError_t launch_ddpm_step(const float* x, const float* eps_c, const float* eps_u, float cfg_scale,
                             const float* noise, float sqrt_ab, float sqrt_1mab, float c0, float c1,
                             float sigma, float* out, long long n, cudaStream_t s) {
  { cudaError_t e_ = launch_pdl(ddpm_step_kernel, dim3(grid_for(n, 256)), dim3(256), 0, s, x, eps_c, eps_u, cfg_scale, noise, sqrt_ab,
                                                    sqrt_1mab, c0, c1, sigma, out, n); if (e_ != cudaSuccess) return e_; }
  return cudaGetLastError();
}

}  // namespace tsd
