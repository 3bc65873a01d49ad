// K1/K2 kernel bodies - see gemm_tcgen05.cuh for the contract.
#include "gemm_tcgen05.cuh"

#include <cstdio>

#include "ptx_sm100.cuh"

namespace tsd {

namespace {

__device__ __forceinline__ float gelu_tanh(float x) {
  // reference Gelu.forward, helpers/utils.mojo:1908-1919 (tanh form)
  const float k = 0.7978845608028654f;  // sqrt(2/pi)
  float u = k * (x + 0.044715f * x * x * x);
  return 0.5f * x * (1.0f + tanhf(u));
}

__device__ __forceinline__ float round_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;\n" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

}  // namespace

__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tf32_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const GemmKParams p) {
  extern __shared__ uint8_t smem_dyn[];
  __shared__ __align__(8) uint64_t full_bar[GEMM_MAX_STAGES];
  __shared__ __align__(8) uint64_t empty_bar[GEMM_MAX_STAGES];
  __shared__ __align__(8) uint64_t accum_bar;
  __shared__ uint32_t tmem_slot;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // 1024 B aligned operand ring (SWIZZLE_128B atoms are 8 rows x 128 B)
  const uint32_t smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  const uint32_t a_bytes = GEMM_BM * GEMM_BK * 4;      // 16 KiB
  const uint32_t b_bytes = (uint32_t)p.BN * GEMM_BK * 4;
  const uint32_t stage_bytes = a_bytes + b_bytes;

  // tile coordinates
  const int nt = blockIdx.x;
  int mt = blockIdx.y;
  const int batch = blockIdx.z / p.splits;
  const int split = blockIdx.z - batch * p.splits;
  const int tw = mt % p.tiles_w;
  mt /= p.tiles_w;
  const int th = mt % p.tiles_h;
  const int img = mt / p.tiles_h;
  const int w0 = tw * p.bw, h0 = th * p.bh;

  const int it_begin = split * p.iters_per_split;
  int it_end = it_begin + p.iters_per_split;
  if (it_end > p.total_iters) it_end = p.total_iters;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.num_stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&accum_bar, 1);
    fence_barrier_init();
    prefetch_tensormap(&tmA);
    prefetch_tensormap(&tmB);
  }
  if (warp == 1) {
    tmem_alloc(&tmem_slot, (uint32_t)p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_d = tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      const int half_rows = p.BN >> 1;
      const uint32_t tx = (uint32_t)p.a_box_bytes + b_bytes;
      int stage = 0;
      uint32_t phase = 0;
      for (int it = it_begin; it < it_end; ++it) {
        const int tap = it / p.chunks_per_tap;
        const int kc = (it - tap * p.chunks_per_tap) * GEMM_BK;
        int dy = 0, dx = 0;
        if (p.taps == 9) {
          dy = tap / 3 - 1;
          dx = tap - (tap / 3) * 3 - 1;
        }
        mbar_wait(&empty_bar[stage], phase ^ 1u);
        uint8_t* sa = smem_dyn + (smem_base - smem_u32(smem_dyn)) + stage * stage_bytes;
        uint8_t* sb = sa + a_bytes;
        mbar_arrive_expect_tx(&full_bar[stage], tx);
        tma_load_4d(sa, &tmA, &full_bar[stage], kc, w0 + dx, h0 + dy, img + batch);
        const int kb = tap * p.cin + kc;
        if (!p.geglu) {
          tma_load_3d(sb, &tmB, &full_bar[stage], kb, nt * p.BN, batch);
        } else {
          tma_load_3d(sb, &tmB, &full_bar[stage], kb, nt * half_rows, batch);
          tma_load_3d(sb + half_rows * GEMM_BK * 4, &tmB, &full_bar[stage], kb,
                      p.n_half + nt * half_rows, batch);
        }
        if (++stage == p.num_stages) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      const uint32_t idesc = umma_idesc(UMMA_FMT_TF32, GEMM_BM, (uint32_t)p.BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      for (int it = it_begin; it < it_end; ++it) {
        const int tap = it / p.chunks_per_tap;
        const int kc = (it - tap * p.chunks_per_tap) * GEMM_BK;
        int nk = (p.cin - kc + 7) >> 3;  // K=8 MMAs with real data in this chunk
        if (nk > GEMM_BK / 8) nk = GEMM_BK / 8;
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after_sync();
        const uint32_t sa = smem_base + stage * stage_bytes;
        const uint32_t sb = sa + a_bytes;
        for (int kk = 0; kk < nk; ++kk) {
          const uint64_t adesc = umma_smem_desc(sa + kk * 32, 16, 1024, UMMA_SWIZZLE_128B);
          const uint64_t bdesc = umma_smem_desc(sb + kk * 32, 16, 1024, UMMA_SWIZZLE_128B);
          umma_tf32(tmem_d, adesc, bdesc, idesc, (it > it_begin || kk > 0) ? 1u : 0u);
        }
        umma_commit(&empty_bar[stage]);  // smem slot reusable once these MMAs retire
        if (++stage == p.num_stages) {
          stage = 0;
          phase ^= 1u;
        }
      }
      umma_commit(&accum_bar);  // accumulator complete
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int q = warp & 3;  // TMEM lane quadrant this warp may read
    const int r = q * 32 + lane;
    const int lh = r / p.bw, lw = r - lh * p.bw;
    const int h = h0 + lh, w = w0 + lw;
    const bool row_ok = (lh < p.bh) && (h < p.H) && (w < p.W);
    const long long gm = ((long long)img * p.H + h) * p.W + w;

    mbar_wait(&accum_bar, 0);
    tc_fence_after_sync();

    const uint32_t trow = tmem_d + ((uint32_t)(q * 32) << 16);
    if (p.partial != nullptr) {
      // raw split-K partials
      float* dst = p.partial +
                   (((long long)split * gridDim.z / p.splits + batch) * p.m_per_batch + gm) * p.n_pad +
                   (long long)nt * p.BN;
      for (int c = 0; c < p.BN; c += 16) {
        uint32_t v[16];
        tmem_ld16(trow + c, v);
        tmem_ld_wait();
        if (row_ok && nt * p.BN + c < p.n_pad) {
          float4* o = reinterpret_cast<float4*>(dst + c);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            o[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                               __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
        }
      }
    } else {
      const int out_cols = p.geglu ? (p.BN >> 1) : p.BN;
      const int n0 = nt * out_cols;
      const float rb = (p.row_bias != nullptr && row_ok) ? p.row_bias[gm] : 0.0f;
      const float* cbias = p.bias ? p.bias + (long long)img * p.bias_img_stride : nullptr;
      float* drow = p.D + (long long)batch * p.d_batch_stride + gm * p.ldd;
      const float* rrow =
          p.residual ? p.residual + (long long)batch * p.r_batch_stride + gm * p.ldr : nullptr;
      for (int c = 0; c < out_cols; c += 16) {
        uint32_t v[16];
        float f[16];
        tmem_ld16(trow + c, v);
        if (p.geglu) {
          uint32_t g[16];
          tmem_ld16(trow + out_cols + c, g);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int n = n0 + c + j;
            float a = __uint_as_float(v[j]), b = __uint_as_float(g[j]);
            if (cbias != nullptr && n < p.n_valid) {
              a += cbias[n];
              b += cbias[p.n_half + n];
            }
            f[j] = a * gelu_tanh(b);
          }
        } else {
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int n = n0 + c + j;
            float a = __uint_as_float(v[j]) * p.alpha + rb;
            if (cbias != nullptr && n < p.n_valid) a += cbias[n];
            f[j] = a;
          }
        }
        const int n = n0 + c;
        if (!row_ok || n >= p.n_valid) continue;
        const long long col = (long long)(n / p.split_n) * p.split_stride + (n % p.split_n);
        if (n + 16 <= p.n_valid) {
          if (rrow != nullptr) {
            const float4* rr = reinterpret_cast<const float4*>(rrow + n);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float4 t = rr[j];
              f[4 * j] += t.x;
              f[4 * j + 1] += t.y;
              f[4 * j + 2] += t.z;
              f[4 * j + 3] += t.w;
            }
          }
          if (p.round_tf32) {
#pragma unroll
            for (int j = 0; j < 16; ++j) f[j] = round_tf32(f[j]);
          }
          float4* o = reinterpret_cast<float4*>(drow + col);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            o[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
        } else {
          for (int j = 0; j < 16 && n + j < p.n_valid; ++j) {
            float a = f[j];
            if (rrow != nullptr) a += rrow[n + j];
            if (p.round_tf32) a = round_tf32(a);
            drow[col + j] = a;
          }
        }
      }
    }
    tc_fence_before_sync();
  }

  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_d, (uint32_t)p.tmem_cols);
  }
}

// Sums split-K partials and applies the (bias, residual, rounding) epilogue.
__global__ void splitk_reduce_kernel(const SplitKReduceParams p) {
  const int n4 = p.n_pad >> 2;
  const long long total = (long long)p.m * n4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int row = (int)(i / n4);
    const int n = (int)(i - (long long)row * n4) * 4;
    if (n >= p.n_valid) continue;
    const float* src = p.partial + (long long)row * p.n_pad + n;
    float4 acc = *reinterpret_cast<const float4*>(src);
    for (int s = 1; s < p.splits; ++s) {
      float4 t = *reinterpret_cast<const float4*>(src + s * p.split_stride);
      acc.x += t.x;
      acc.y += t.y;
      acc.z += t.z;
      acc.w += t.w;
    }
    float f[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (n + j < p.n_valid) {
        float a = f[j];
        if (p.bias) a += p.bias[(long long)(row / p.rows_per_img) * p.bias_img_stride + n + j];
        if (p.residual) a += p.residual[(long long)row * p.ldr + n + j];
        if (p.round_tf32) a = round_tf32(a);
        f[j] = a;
      }
    }
    float* dst = p.D + (long long)row * p.ldd + n;
    if (n + 4 <= p.n_valid && (p.ldd & 3) == 0) {
      *reinterpret_cast<float4*>(dst) = make_float4(f[0], f[1], f[2], f[3]);
    } else {
      for (int j = 0; j < 4 && n + j < p.n_valid; ++j) dst[j] = f[j];
    }
  }
}

size_t gemm_smem_bytes(int BN, int num_stages) {
  return (size_t)num_stages * (GEMM_BM * GEMM_BK * 4 + (size_t)BN * GEMM_BK * 4) + 1024;
}

int gemm_pick_stages(int BN) {
  const size_t budget = 200 * 1024;
  int s = (int)((budget - 1024) / (GEMM_BM * GEMM_BK * 4 + (size_t)BN * GEMM_BK * 4));
  if (s > GEMM_MAX_STAGES) s = GEMM_MAX_STAGES;
  if (s < 2) s = 2;
  return s;
}

cudaError_t launch_gemm_tf32(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmKParams& p,
                             dim3 grid, size_t smem_bytes, cudaStream_t stream) {
  // static + dynamic shared memory must fit the 227 KiB opt-in limit together
  static int max_dyn = -1;
  if (max_dyn < 0) {
    cudaFuncAttributes fa;
    cudaError_t e = cudaFuncGetAttributes(&fa, gemm_tf32_kernel);
    if (e != cudaSuccess) return e;
    int dev = 0, optin = 0;
    cudaGetDevice(&dev);
    e = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (e != cudaSuccess) return e;
    const int lim = optin - (int)fa.sharedSizeBytes;
    e = cudaFuncSetAttribute(gemm_tf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, lim);
    if (e != cudaSuccess) return e;
    max_dyn = lim;
  }
  if ((long long)smem_bytes > max_dyn) return cudaErrorInvalidConfiguration;
  gemm_tf32_kernel<<<grid, GEMM_THREADS, smem_bytes, stream>>>(tmA, tmB, p);
  return cudaGetLastError();
}

cudaError_t launch_splitk_reduce(const SplitKReduceParams& p, cudaStream_t stream) {
  const long long total = (long long)p.m * (p.n_pad >> 2);
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  splitk_reduce_kernel<<<blocks, 256, 0, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace tsd
