// K3/K4: fused attention core on tcgen05 tensor cores - see attention_tcgen05.cuh.
//
// Reference semantics (helpers/attention.mojo:46-62, 105-115): per head S = Q K^T / sqrt(d),
// Softmax(dim=2) (which normalises every COLUMN of S over the query axis, utils.mojo:435-445,
// SURVEY Q3; "softmax_axis = 1" selects the standard key-axis form), O = P V, merged heads.
//
// Both softmax axes are served by the same two launches of one kernel:
//   STATS  rows of X against all rows of Y: per X-row r the log2-domain max m_r of
//          c * <x_r, y_j> and l_r = sum_j exp2(c * <x_r, y_j> - m_r)   (c = log2(e)/sqrt(d));
//          query axis: X = K, Y = Q (one statistic per key column); key axis: X = Q, Y = K.
//          The Y loop can be split over grid.z; partial (m, l) pairs are merged by the consumer.
//   APPLY  per 128-query tile: S = Q K_t^T in TMEM, P = exp2(c * S - mu) with mu = m + log2(l)
//          per key column (query axis) or per query row (key axis), written back over S in TMEM
//          (tcgen05.st) and fed straight to the second MMA as the TMEM A operand:
//          O += P V_t, V staged by TMA in its natural [key][d] layout and read MN-major.
//          S / P never touch shared or global memory; no running-max correction of O is needed
//          because mu is known before the pass starts.
// Warp roles: warp 0 TMA producer, warp 1 MMA issuer (+ TMEM owner), warps 2..9 softmax /
// epilogue (two threads per S row, each a column half; TMEM lane quadrant = warp & 3).
#include "attention_tcgen05.cuh"

#include <cmath>
#include <cstdlib>

#include "pdl.cuh"
#include "ptx_sm100.cuh"

namespace tsd {

int make_tmap_f32(Ctx* c, CUtensorMap* tm, const float* base, int rank, const uint64_t* dims,
                  const uint64_t* strides_elems, const uint32_t* box, int swizzle_atom_32b,
                  const uint32_t* elem_strides);

namespace {

constexpr int ATT_BM = 128;
constexpr int ATT_SM_THREADS = 256;               // 8 softmax warps
constexpr int ATT_THREADS = 64 + ATT_SM_THREADS;  // + TMA producer warp + MMA issuer warp
constexpr int ATT_BOX_ROW_BYTES = 128;  // 32 fp32 = one SWIZZLE_128B row
enum { ATT_STATS = 0, ATT_APPLY = 1 };

struct AttnKParams {
  int mode;
  int Tx, Ty;      // rows of the stationary operand (owned by CTAs) / of the streamed operand
  int d, nbox;     // head dim, number of 32-float boxes covering it
  int dpad;        // N of the P.V MMA: d rounded up to 16
  int BN;          // streamed rows per tile (multiple of 16, <= 128)
  int sY, sV;      // ring depths (1 or 2)
  int heads;
  int x_shared, y_shared, v_shared;  // tensor-map batch coordinate = head only (operand shared by the batch)
  int tiles_per_split, total_tiles;
  int tmem_cols;
  float c;                // softmax scale * log2(e)
  float2* part_out;       // STATS: [split][bh][Tx]
  const float2* part_in;  // APPLY: [mu_splits][bh][Tmu]
  int mu_splits, mu_per_row, Tmu;
  long long part_stride;  // elements between splits = bh_count * T
  float* O;               // APPLY: merged output [b][Tx][heads*d]
  int ldo;
  int round_out;
  int debug;  // TSD_ATTN_DEBUG: 1 = P := 1, 2 = P := S (bring-up probes)
};

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;\n" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;\n" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t rna_tf32_bits(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;\n" : "=r"(r) : "f"(x));
  return r;
}

// merge `splits` partial (max, sum) pairs of one row/column into mu = max + log2(sum)
__device__ __forceinline__ float merge_partials(const float2* __restrict__ part, long long stride, int splits,
                                                long long idx) {
  float2 v = part[idx];
  float M = v.x, L = v.y;
  for (int s = 1; s < splits; ++s) {
    v = part[idx + (long long)s * stride];
    const float Mn = fmaxf(M, v.x);
    L = L * ex2(M - Mn) + v.y * ex2(v.x - Mn);
    M = Mn;
  }
  return M + lg2(L);
}

__global__ void __launch_bounds__(ATT_THREADS, 1)
attn_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmY,
            const __grid_constant__ CUtensorMap tmV, const AttnKParams p) {
  extern __shared__ uint8_t smem_dyn[];
  __shared__ __align__(8) uint64_t x_full, y_full[2], y_empty[2], v_full[2], v_empty[2];
  __shared__ __align__(8) uint64_t s_full[2], s_free[2], p_full[2], o_full;
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(16) float mu_s[2][ATT_BM];

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const uint32_t dyn0 = smem_u32(smem_dyn);
  const uint32_t xs = (dyn0 + 1023u) & ~1023u;
  const uint32_t x_bytes = (uint32_t)p.nbox * ATT_BM * ATT_BOX_ROW_BYTES;
  const uint32_t ybox = (uint32_t)p.BN * ATT_BOX_ROW_BYTES;
  const uint32_t ystage = (uint32_t)p.nbox * ybox;
  const uint32_t ys = xs + x_bytes;
  const uint32_t vs = ys + (uint32_t)p.sY * ystage;

  const int xt = blockIdx.x, bh = blockIdx.y, split = blockIdx.z;
  const int h = bh % p.heads;
  const int x0 = xt * ATT_BM;
  const int it0 = split * p.tiles_per_split;
  int n_it = p.total_tiles - it0;
  if (n_it > p.tiles_per_split) n_it = p.tiles_per_split;
  const bool apply = p.mode == ATT_APPLY;

  if (threadIdx.x == 0) {
    mbar_init(&x_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&y_full[s], 1);
      mbar_init(&y_empty[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 1);
      mbar_init(&s_full[s], 1);
      mbar_init(&s_free[s], ATT_SM_THREADS);
      mbar_init(&p_full[s], ATT_SM_THREADS);
    }
    mbar_init(&o_full, 1);
    fence_barrier_init();
    prefetch_tensormap(&tmX);
    prefetch_tensormap(&tmY);
    if (apply) prefetch_tensormap(&tmV);
  }
  if (warp == 1) {
    tmem_alloc(&tmem_slot, (uint32_t)p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = tmem_slot;
  pdl_wait();               // Q/K/V and the statistics come from predecessor kernels
  pdl_launch_dependents();  // resources are held: the next kernel may start its prologue
  const uint32_t tmem_o = tmem_base + 256u;  // S0: cols [0,128), S1: [128,256), O: [256, 256+dpad)

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      mbar_arrive_expect_tx(&x_full, x_bytes);
      for (int bx = 0; bx < p.nbox; ++bx)
        tma_load_3d_s(xs + bx * (ATT_BM * ATT_BOX_ROW_BYTES), &tmX, &x_full, bx * 32, x0,
                    p.x_shared ? h : bh);
      for (int it = 0; it < n_it; ++it) {
        const int yrow = (it0 + it) * p.BN;
        {
          const int st = it % p.sY;
          const uint32_t ph = (uint32_t)(it / p.sY) & 1u;
          mbar_wait(&y_empty[st], ph ^ 1u);
          mbar_arrive_expect_tx(&y_full[st], ystage);
          for (int bx = 0; bx < p.nbox; ++bx)
            tma_load_3d_s(ys + st * ystage + bx * ybox, &tmY, &y_full[st], bx * 32, yrow,
                        p.y_shared ? h : bh);
        }
        if (apply) {
          const int st = it % p.sV;
          const uint32_t ph = (uint32_t)(it / p.sV) & 1u;
          mbar_wait(&v_empty[st], ph ^ 1u);
          mbar_arrive_expect_tx(&v_full[st], ystage);
          for (int bx = 0; bx < p.nbox; ++bx)
            tma_load_3d_s(vs + st * ystage + bx * ybox, &tmV, &v_full[st], bx * 32, yrow,
                        p.v_shared ? h : bh);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      const uint32_t idesc_s = umma_idesc(UMMA_FMT_TF32, ATT_BM, (uint32_t)p.BN, 0, 0);
      const uint32_t idesc_o = umma_idesc(UMMA_FMT_TF32, ATT_BM, (uint32_t)p.dpad, 0, 1);  // B (= V) MN-major
      const int ksteps = p.d >> 3;
      const int pv_steps = p.BN >> 3;
      auto issue_pv = [&](int j) {
        const int bj = j & 1;
        const int st = j % p.sV;
        mbar_wait(&p_full[bj], (uint32_t)(j >> 1) & 1u);
        mbar_wait(&v_full[st], (uint32_t)(j / p.sV) & 1u);
        tc_fence_after_sync();
        const uint32_t vb = vs + st * ystage;
        for (int ks = 0; ks < pv_steps; ++ks) {
          // V tile: [BN keys][32 floats] boxes read MN-major.  For 32-bit MN-major operands the only
          // legal layout is SWIZZLE_128B with 32 B atoms: 4-key groups of 512 B (SBO), N atoms
          // (boxes of 32 floats) ybox apart (LBO); one K = 8 step spans two groups.
          const uint64_t bdesc = umma_smem_desc(vb + ks * 1024, ybox, 512, UMMA_SWIZZLE_128B_BASE32B);
          umma_tf32_ts(tmem_o, tmem_base + (uint32_t)(bj * 128 + ks * 8), bdesc, idesc_o,
                       (j > 0 || ks > 0) ? 1u : 0u);
        }
        umma_commit(&v_empty[st]);
      };
      mbar_wait(&x_full, 0);
      for (int it = 0; it < n_it; ++it) {
        const int b = it & 1;
        const int st = it % p.sY;
        mbar_wait(&y_full[st], (uint32_t)(it / p.sY) & 1u);
        if (!apply && it >= 2) mbar_wait(&s_free[b], (uint32_t)((it >> 1) - 1) & 1u);
        tc_fence_after_sync();
        const uint32_t yb = ys + st * ystage;
        for (int ks = 0; ks < ksteps; ++ks) {
          const uint32_t off = (uint32_t)(ks & 3) * 32u;
          const uint64_t adesc =
              umma_smem_desc(xs + (ks >> 2) * (ATT_BM * ATT_BOX_ROW_BYTES) + off, 16, 1024, UMMA_SWIZZLE_128B);
          const uint64_t bdesc = umma_smem_desc(yb + (ks >> 2) * ybox + off, 16, 1024, UMMA_SWIZZLE_128B);
          umma_tf32(tmem_base + (uint32_t)(b * 128), adesc, bdesc, idesc_s, ks > 0 ? 1u : 0u);
        }
        umma_commit(&y_empty[st]);
        umma_commit(&s_full[b]);
        if (apply && it >= 1) issue_pv(it - 1);
      }
      if (apply) {
        issue_pv(n_it - 1);
        umma_commit(&o_full);
      }
    }
  } else {
    // ===================== softmax / epilogue (warps 2..9) =====================
    // Two warps per TMEM lane quadrant (a warp may only touch lanes 32*(warp%4)..+31): both own the
    // same 32 S rows, each a contiguous half of the tile's columns, so that every SM scheduler has
    // two softmax warps to interleave (one alone is issue/latency bound).
    const int sw = warp - 2;        // 0..7
    const int q = warp & 3;         // lane quadrant
    const int half = sw >> 2;       // column half
    const int row = q * 32 + lane;  // S row == TMEM lane
    const int st = sw * 32 + lane;  // 0..255 among the softmax threads
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    const bool row_ok = x0 + row < p.Tx;
    const float c = p.c;
    const int h0 = ((p.BN >> 4) + 1) >> 1;  // 16-column chunks owned by half 0
    const int cb = half ? h0 * 16 : 0, ce = half ? p.BN : h0 * 16;

    if (!apply) {
      float m = -INFINITY, l = 0.f;
      for (int it = 0; it < n_it; ++it) {
        const int b = it & 1;
        const int ycol0 = (it0 + it) * p.BN;
        int nvalid = p.Ty - ycol0;
        if (nvalid > p.BN) nvalid = p.BN;
        mbar_wait(&s_full[b], (uint32_t)(it >> 1) & 1u);
        tc_fence_after_sync();
        for (int c0 = cb; c0 < ce; c0 += 32) {
          uint32_t v[32];
          if (c0 + 32 <= ce) {
            tmem_ld32(lane_base + (uint32_t)(b * 128 + c0), v);
          } else {
            uint32_t t16[16];
            tmem_ld16(lane_base + (uint32_t)(b * 128 + c0), t16);
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = t16[j];
#pragma unroll
            for (int j = 16; j < 32; ++j) v[j] = 0xff800000u;  // -inf
          }
          tmem_ld_wait();
          if (c0 >= nvalid) continue;  // whole chunk past the last valid column
          if (c0 + 32 > nvalid) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (c0 + j >= nvalid) v[j] = 0xff800000u;
          }
          float cm = __uint_as_float(v[0]);
#pragma unroll
          for (int j = 1; j < 32; ++j) cm = fmaxf(cm, __uint_as_float(v[j]));
          cm *= c;
          if (cm > m) {
            l *= ex2(m - cm);
            m = cm;
          }
          float a0 = 0.f, a1 = 0.f;
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            a0 += ex2(fmaf(__uint_as_float(v[j]), c, -m));
            a1 += ex2(fmaf(__uint_as_float(v[j + 1]), c, -m));
          }
          l += a0 + a1;
        }
        tc_fence_before_sync();
        mbar_arrive(&s_free[b]);  // this thread's reads of S[b] are complete
      }
      // merge the two column halves of each row (half 1 -> shared -> half 0), then store
      float2* comb = reinterpret_cast<float2*>(&mu_s[0][0]);  // 128 float2 = the whole mu_s array
      if (half == 1) comb[row] = make_float2(m, l);
      named_bar_sync(1, ATT_SM_THREADS);
      if (half == 0 && row_ok) {
        const float2 o = comb[row];
        const float M = fmaxf(m, o.x);
        const float L = l * ex2(m - M) + o.y * ex2(o.x - M);
        p.part_out[(long long)split * p.part_stride + (long long)bh * p.Tx + x0 + row] = make_float2(M, L);
      }
    } else {
      // P is written with a 2^-11 relative boost so that the tensor core's truncation of the fp32
      // bit pattern to TF32 acts as round-to-nearest (unbiased), at no instruction cost
      const float rnd = 7.0436e-4f;  // log2(1 + 2^-11)
      float mu_row = 0.f;
      if (p.mu_per_row && row_ok)
        mu_row = merge_partials(p.part_in, p.part_stride, p.mu_splits, (long long)bh * p.Tmu + x0 + row) - rnd;
      float2 pre = make_float2(0.f, 1.f);
      const bool fast_mu = !p.mu_per_row && p.mu_splits == 1;
      if (fast_mu && st < p.BN && it0 * p.BN + st < p.Ty) pre = p.part_in[(long long)bh * p.Tmu + it0 * p.BN + st];
      for (int it = 0; it < n_it; ++it) {
        const int b = it & 1;
        const int ycol0 = (it0 + it) * p.BN;
        int nvalid = p.Ty - ycol0;
        if (nvalid > p.BN) nvalid = p.BN;
        if (!p.mu_per_row) {
          if (st < p.BN) {
            const int col = ycol0 + st;
            float mu = INFINITY;  // masked key column: exp2(-inf) = 0
            if (col < p.Ty)
              mu = (fast_mu ? pre.x + lg2(pre.y)
                            : merge_partials(p.part_in, p.part_stride, p.mu_splits, (long long)bh * p.Tmu + col)) - rnd;
            mu_s[b][st] = mu;
          }
          named_bar_sync(1, ATT_SM_THREADS);
          if (fast_mu && it + 1 < n_it && st < p.BN && ycol0 + p.BN + st < p.Ty)
            pre = p.part_in[(long long)bh * p.Tmu + ycol0 + p.BN + st];
        }
        mbar_wait(&s_full[b], (uint32_t)(it >> 1) & 1u);
        tc_fence_after_sync();
        for (int c0 = cb; c0 < ce; c0 += 32) {
          const uint32_t taddr = lane_base + (uint32_t)(b * 128 + c0);
          const bool full = c0 + 32 <= ce;
          uint32_t v[32];
          if (full) {
            tmem_ld32(taddr, v);
          } else {
            uint32_t t16[16];
            tmem_ld16(taddr, t16);
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = t16[j];
#pragma unroll
            for (int j = 16; j < 32; ++j) v[j] = 0;
          }
          tmem_ld_wait();
          if (p.debug) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float e = p.debug == 1 ? 1.0f : __uint_as_float(v[j]);
              v[j] = (c0 + j < nvalid) ? rna_tf32_bits(e) : 0u;
            }
          } else if (!p.mu_per_row) {
            const float4* mup = reinterpret_cast<const float4*>(&mu_s[b][c0]);
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              if (!full && j4 >= 4) break;
              const float4 m4 = mup[j4];
              v[4 * j4 + 0] = __float_as_uint(ex2(fmaf(__uint_as_float(v[4 * j4 + 0]), c, -m4.x)));
              v[4 * j4 + 1] = __float_as_uint(ex2(fmaf(__uint_as_float(v[4 * j4 + 1]), c, -m4.y)));
              v[4 * j4 + 2] = __float_as_uint(ex2(fmaf(__uint_as_float(v[4 * j4 + 2]), c, -m4.z)));
              v[4 * j4 + 3] = __float_as_uint(ex2(fmaf(__uint_as_float(v[4 * j4 + 3]), c, -m4.w)));
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float e = ex2(fmaf(__uint_as_float(v[j]), c, -mu_row));
              v[j] = (c0 + j < nvalid) ? __float_as_uint(e) : 0u;
            }
          }
          if (full) {
            tmem_st32(taddr, v);
          } else {
            uint32_t t16[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) t16[j] = v[j];
            tmem_st16(taddr, t16);
          }
        }
        tmem_st_wait();
        tc_fence_before_sync();
        mbar_arrive(&p_full[b]);
      }
      // ---- epilogue: O (TMEM) -> merged [b][t][h*d + j]  (attention.mojo:61-62) ----
      mbar_wait(&o_full, 0);
      tc_fence_after_sync();
      const int bidx = bh / p.heads;
      float* orow = p.O + ((long long)bidx * p.Tx + x0 + row) * p.ldo + (long long)h * p.d;
      const int e0 = (((p.dpad >> 4) + 1) >> 1) * 16;
      const int ob = half ? e0 : 0, oe = half ? p.dpad : e0;
      for (int c0 = ob; c0 < oe; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(lane_base + 256u + (uint32_t)c0, v);
        tmem_ld_wait();
        if (row_ok) {
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            if (c0 + 4 * j4 < p.d) {
              float4 o;
              if (p.round_out) {
                o = make_float4(__uint_as_float(rna_tf32_bits(__uint_as_float(v[4 * j4]))),
                                __uint_as_float(rna_tf32_bits(__uint_as_float(v[4 * j4 + 1]))),
                                __uint_as_float(rna_tf32_bits(__uint_as_float(v[4 * j4 + 2]))),
                                __uint_as_float(rna_tf32_bits(__uint_as_float(v[4 * j4 + 3]))));
              } else {
                o = make_float4(__uint_as_float(v[4 * j4]), __uint_as_float(v[4 * j4 + 1]),
                                __uint_as_float(v[4 * j4 + 2]), __uint_as_float(v[4 * j4 + 3]));
              }
              *reinterpret_cast<float4*>(orow + c0 + 4 * j4) = o;
            }
          }
        }
      }
    }
    tc_fence_before_sync();
  }

  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}


// ------------------------------------------------------------------------------------------
// v2 softmax role: the same protocol as attn_kernel, with the tensor-memory reads software-pipelined.
//
// Measured on the v1 kernel (profiles/r01_ncu_attn.txt, T = 4096, d = 40): 2240 (statistics) / 2917 (apply) cycles per
// 128 x 128 tile against 1024 cycles of MUFU.EX2 issue.  A tile is 64 KiB of fp32 scores that every softmax thread
// must pull out of tensor memory (tcgen05.ld, ~64 B/clk per SM) before it can exponentiate them, and in v1 all eight
// softmax warps did the two phases in lock step: load, wait, compute, (store, wait,) barrier.  Here each warp keeps
// the load of chunk g+1 in flight while it exponentiates chunk g (two register buffers, loop unrolled by two), runs
// straight across tile boundaries, and the per-tile all-warp barrier is gone: mu (query axis) is staged in shared
// memory for every key of the pass once, before the loop.
// Preconditions (attention_fused picks v1 otherwise): BN in {64, 128}, Ty a multiple of BN, and for the query-axis
// apply pass Ty * 4 bytes of extra shared memory.
// ------------------------------------------------------------------------------------------
// 2^x for x <= ~0 on the FMA / ALU pipes (no MUFU): Cody-Waite split x = n + f, f in [-0.5, 0.5], degree-3 minimax
// polynomial for 2^f (max relative error 7.5e-5, below the 2^-11 rounding P takes on its way into the tensor core),
// exponent added to the bit pattern.  MUFU.EX2 retires one warp instruction per ~10.4 cycles per scheduler on this part
// (measured, tools/lab/ubench.cu), which bounds both softmax passes; every fourth exponential taken here shortens the
// MUFU queue by a quarter while the FMA pipe (otherwise one FFMA per element) stays below it.
__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -125.0f);                    // keeps the result a normal number (values this small do not matter)
  const float y = x + 12582912.0f;          // 1.5 * 2^23: the integer nearest to x lands in the low mantissa bits
  const float n = y - 12582912.0f;
  const float f = x - n;
  float p = fmaf(f, 0.05517157f, 0.24261111f);
  p = fmaf(p, f, 0.69326103f);
  p = fmaf(p, f, 0.99992806f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(y) << 23));
}

__device__ __forceinline__ void reg_fence32(uint32_t (&v)[32]) {
  // zero instructions: pins the uses of v[] behind the preceding tcgen05.wait::ld in the compiler's schedule
  asm volatile("" : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
                    "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]),
                    "+r"(v[16]), "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]),
                    "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31]));
}

template <int MODE, bool PER_ROW, int NCH, bool TRACE, bool POLY>
__global__ void __launch_bounds__(ATT_THREADS, 1)
attn2_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmY,
             const __grid_constant__ CUtensorMap tmV, const AttnKParams p) {
  extern __shared__ uint8_t smem_dyn[];
  __shared__ __align__(8) uint64_t x_full, y_full[2], y_empty[2], v_full[2], v_empty[2];
  __shared__ __align__(8) uint64_t s_full[3], s_free[3], p_full[3], o_full;  // three S / P buffers in tensor memory
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(16) float2 comb[ATT_BM];
  constexpr bool apply = MODE == ATT_APPLY;
  constexpr int BN = NCH * 64;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const uint32_t dyn0 = smem_u32(smem_dyn);
  const uint32_t xs = (dyn0 + 1023u) & ~1023u;
  const uint32_t x_bytes = (uint32_t)p.nbox * ATT_BM * ATT_BOX_ROW_BYTES;
  const uint32_t ybox = (uint32_t)BN * ATT_BOX_ROW_BYTES;
  const uint32_t ystage = (uint32_t)p.nbox * ybox;
  const uint32_t ys = xs + x_bytes;
  const uint32_t vs = ys + (uint32_t)p.sY * ystage;
  float* mu_all = reinterpret_cast<float*>(smem_dyn + (vs - dyn0) + (size_t)(apply ? p.sV : 0) * ystage);

  const int xt = blockIdx.x, bh = blockIdx.y, split = blockIdx.z;
  const int h = bh % p.heads;
  const int x0 = xt * ATT_BM;
  const int it0 = split * p.tiles_per_split;
  int n_it = p.total_tiles - it0;
  if (n_it > p.tiles_per_split) n_it = p.tiles_per_split;

  if (threadIdx.x == 0) {
    mbar_init(&x_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&y_full[s], 1);
      mbar_init(&y_empty[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 1);
    }
    for (int s = 0; s < 3; ++s) {
      mbar_init(&s_full[s], 1);
      mbar_init(&s_free[s], ATT_SM_THREADS);
      mbar_init(&p_full[s], ATT_SM_THREADS);
    }
    mbar_init(&o_full, 1);
    fence_barrier_init();
    prefetch_tensormap(&tmX);
    prefetch_tensormap(&tmY);
    if (apply) prefetch_tensormap(&tmV);
  }
  if (warp == 1) {
    tmem_alloc(&tmem_slot, (uint32_t)p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = tmem_slot;
  // Statistics pass: Q / K come from the predecessor kernel - everybody waits.  Apply pass: its stream predecessor is
  // the statistics pass of the same attention op, which itself waited for every earlier kernel BEFORE it allowed this
  // launch (pdl_wait precedes pdl_launch_dependents in every thread of it), so Q / K / V are complete and visible the
  // moment this kernel runs: the producer and the MMA issuer start at once, only the softmax warps - readers of the
  // statistics, writers of O - wait.
  if (!apply || warp >= 2) pdl_wait();
  pdl_launch_dependents();  // resources are held: the next kernel may start its prologue
  // S / P buffer b of a tile: columns [b * BN, (b + 1) * BN), b = tile % 3; O: [3 * BN, 3 * BN + dpad)
  const uint32_t tmem_o = tmem_base + (uint32_t)(3 * BN);

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      mbar_arrive_expect_tx(&x_full, x_bytes);
      for (int bx = 0; bx < p.nbox; ++bx)
        tma_load_3d_s(xs + bx * (ATT_BM * ATT_BOX_ROW_BYTES), &tmX, &x_full, bx * 32, x0, p.x_shared ? h : bh);
      // K tiles run one ahead of V tiles: S = Q K^T of tile it+1 is issued while P V of tile it-2 is still waiting
      // for its softmax, so the K load must never queue behind a V slot that frees late
      auto load_y = [&](int it) {
        const int st = it % p.sY;
        const uint32_t ph = (uint32_t)(it / p.sY) & 1u;
        mbar_wait(&y_empty[st], ph ^ 1u);
        mbar_arrive_expect_tx(&y_full[st], ystage);
        for (int bx = 0; bx < p.nbox; ++bx)
          tma_load_3d_s(ys + st * ystage + bx * ybox, &tmY, &y_full[st], bx * 32, (it0 + it) * BN, p.y_shared ? h : bh);
      };
      if (n_it > 0) load_y(0);
      for (int it = 0; it < n_it; ++it) {
        if (it + 1 < n_it) load_y(it + 1);
        if (apply) {
          const int st = it % p.sV;
          const uint32_t ph = (uint32_t)(it / p.sV) & 1u;
          mbar_wait(&v_empty[st], ph ^ 1u);
          mbar_arrive_expect_tx(&v_full[st], ystage);
          for (int bx = 0; bx < p.nbox; ++bx)
            tma_load_3d_s(vs + st * ystage + bx * ybox, &tmV, &v_full[st], bx * 32, (it0 + it) * BN, p.v_shared ? h : bh);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // A lone thread pays 4-6 cycles per dependent instruction and ~90 per mbarrier try_wait: with descriptors rebuilt
    // per MMA this role took ~100 cycles per instruction (21 per 128 x 128 tile at d = 40) and bounded the apply pass
    // (trace: the softmax warps waited 625 of 2146 cycles per tile for their scores).  Descriptors are therefore built
    // once and advanced by adds of compile-time constants; the K loops are unrolled.
    if (elect_one()) {
      const uint32_t idesc_s = umma_idesc(UMMA_FMT_TF32, ATT_BM, (uint32_t)BN, 0, 0);
      const uint32_t idesc_o = umma_idesc(UMMA_FMT_TF32, ATT_BM, (uint32_t)p.dpad, 0, 1);  // B (= V) MN-major
      const int ksteps = p.d >> 3;
      constexpr int pv_steps = BN >> 3;
      // descriptor start-address fields are in 16-byte units
      const uint64_t xdesc0 = umma_smem_desc(xs, 16, 1024, UMMA_SWIZZLE_128B);
      const uint64_t ydesc0 = umma_smem_desc(ys, 16, 1024, UMMA_SWIZZLE_128B);
      // V tile: [BN keys][32 floats] boxes read MN-major.  For 32-bit MN-major operands the only legal layout is
      // SWIZZLE_128B with 32 B atoms: 4-key groups of 512 B (SBO), N atoms (boxes of 32 floats) ybox apart (LBO); one
      // K = 8 step spans two groups (1024 B).
      const uint64_t vdesc0 = umma_smem_desc(vs, ybox, 512, UMMA_SWIZZLE_128B_BASE32B);
      const uint32_t stage_step = ystage >> 4, ybox_step = ybox >> 4;
      constexpr uint32_t xbox_step = (ATT_BM * ATT_BOX_ROW_BYTES) >> 4;
      auto issue_pv = [&](int j, int bj, int stv, uint32_t par_p, uint32_t par_v) {
        mbar_wait(&p_full[bj], par_p);
        mbar_wait(&v_full[stv], par_v);
        tc_fence_after_sync();
        const uint64_t vd = vdesc0 + (uint64_t)((uint32_t)stv * stage_step);
        const uint32_t pa = tmem_base + (uint32_t)(bj * BN);
        umma_tf32_ts(tmem_o, pa, vd, idesc_o, j > 0 ? 1u : 0u);
#pragma unroll
        for (int ks = 1; ks < pv_steps; ++ks) umma_tf32_ts(tmem_o, pa + (uint32_t)(ks * 8), vd + (uint64_t)(ks * 64), idesc_o, 1u);
        umma_commit(&v_empty[stv]);
      };
      mbar_wait(&x_full, 0);
      // running tile state: S buffer b = it % 3 with its barrier parity, ring slots and parities of K (sY) and V (sV)
      int b = 0, sty = 0;
      uint32_t par_s = 0, par_y = 0;
      int pj = 0, pb = 0, stv = 0;      // P V bookkeeping of tile pj = it - 2
      uint32_t par_p = 0, par_v = 0;
      auto advance_pv = [&]() {
        ++pj;
        if (++pb == 3) { pb = 0; par_p ^= 1u; }
        if (++stv == p.sV) { stv = 0; par_v ^= 1u; }
      };
      for (int it = 0; it < n_it; ++it) {
        // Three S buffers: S of tile `it` is issued as soon as P V of tile it-3 has been (the tensor pipe runs in
        // order), two tiles ahead of the softmax that consumes it - the P V of a tile, which has to wait for that
        // tile's exponentials, is then never on the path to the next tile's scores
        mbar_wait(&y_full[sty], par_y);
        if (!apply && it >= 3) mbar_wait(&s_free[b], par_s ^ 1u);
        tc_fence_after_sync();
        const uint64_t yd = ydesc0 + (uint64_t)((uint32_t)sty * stage_step);
        const uint32_t sd = tmem_base + (uint32_t)(b * BN);
        int kleft = ksteps;
        for (int bx = 0; bx < p.nbox; ++bx, kleft -= 4) {
          const uint64_t xa = xdesc0 + (uint64_t)((uint32_t)bx * xbox_step), ya = yd + (uint64_t)((uint32_t)bx * ybox_step);
          if (kleft >= 4) {
            umma_tf32(sd, xa, ya, idesc_s, bx > 0 ? 1u : 0u);
            umma_tf32(sd, xa + 2, ya + 2, idesc_s, 1u);
            umma_tf32(sd, xa + 4, ya + 4, idesc_s, 1u);
            umma_tf32(sd, xa + 6, ya + 6, idesc_s, 1u);
          } else {
            for (int kk = 0; kk < kleft; ++kk) umma_tf32(sd, xa + (uint64_t)(2 * kk), ya + (uint64_t)(2 * kk), idesc_s, (bx > 0 || kk > 0) ? 1u : 0u);
          }
        }
        umma_commit(&y_empty[sty]);
        umma_commit(&s_full[b]);
        if (++b == 3) { b = 0; par_s ^= 1u; }
        if (++sty == p.sY) { sty = 0; par_y ^= 1u; }
        if (apply && it >= 2) {
          issue_pv(pj, pb, stv, par_p, par_v);
          advance_pv();
        }
      }
      if (apply) {
        while (pj < n_it) {
          issue_pv(pj, pb, stv, par_p, par_v);
          advance_pv();
        }
        umma_commit(&o_full);
      }
    }
  } else {
    // ===================== softmax / epilogue (warps 2..9) =====================
    const int sw = warp - 2;        // 0..7
    const int q = warp & 3;         // lane quadrant
    const int half = sw >> 2;       // column half: this warp owns columns [cb, cb + 32 * NCH) of every tile
    const int row = q * 32 + lane;  // S row == TMEM lane
    const int st = sw * 32 + lane;  // 0..255 among the softmax threads
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    const bool row_ok = x0 + row < p.Tx;
    const float c = p.c;
    const int cb = half * (BN / 2);
    const int total = n_it * NCH;   // 32-column chunks of this warp over the whole pass
    [[maybe_unused]] long long t_wait_s = 0, t_wait_ld = 0, t_wait_st = 0, t_begin = 0;
    const bool tr = TRACE && st == 0;

    // load of chunk g into `buf`; at a tile boundary first wait for the tile's scores
    auto issue = [&](int g, uint32_t (&buf)[32]) {
      const int it = g / NCH, k = g - it * NCH;
      if (k == 0) {
        long long t0 = 0;
        if (tr) t0 = clock64();
        mbar_wait(&s_full[it % 3], (uint32_t)(it / 3) & 1u);
        tc_fence_after_sync();
        if (tr) t_wait_s += clock64() - t0;
      }
      tmem_ld32(lane_base + (uint32_t)((it % 3) * BN + cb + k * 32), buf);
    };
    auto ld_wait = [&](uint32_t (&buf)[32]) {
      long long t0 = 0;
      if (tr) t0 = clock64();
      tmem_ld_wait();
      reg_fence32(buf);
      if (tr) t_wait_ld += clock64() - t0;
    };

    if constexpr (!apply) {
      float m = -INFINITY, l = 0.f;
      auto consume = [&](int g, uint32_t (&v)[32]) {
        const int it = g / NCH, k = g - it * NCH;
        float cm0 = fmaxf(__uint_as_float(v[0]), __uint_as_float(v[1]));
        float cm1 = fmaxf(__uint_as_float(v[2]), __uint_as_float(v[3]));
#pragma unroll
        for (int j = 4; j < 32; j += 4) {
          cm0 = fmaxf(cm0, fmaxf(__uint_as_float(v[j]), __uint_as_float(v[j + 1])));
          cm1 = fmaxf(cm1, fmaxf(__uint_as_float(v[j + 2]), __uint_as_float(v[j + 3])));
        }
        const float cm = fmaxf(cm0, cm1) * c;  // c > 0
        if (cm > m) {
          l *= ex2(m - cm);
          m = cm;
        }
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          a0 += ex2(fmaf(__uint_as_float(v[j]), c, -m));
          a1 += ex2(fmaf(__uint_as_float(v[j + 1]), c, -m));
          a2 += ex2(fmaf(__uint_as_float(v[j + 2]), c, -m));
          a3 += POLY ? ex2_poly(fmaf(__uint_as_float(v[j + 3]), c, -m)) : ex2(fmaf(__uint_as_float(v[j + 3]), c, -m));
        }
        l += (a0 + a1) + (a2 + a3);
        if (k == NCH - 1) {
          tc_fence_before_sync();
          mbar_arrive(&s_free[it % 3]);  // this thread's reads of this tile's S are complete
        }
      };
      uint32_t A[32], B[32];
      if (tr) t_begin = clock64();
      if (total > 0) issue(0, A);
      // reg_fence32 after each issue(): the consumers of the current buffer are pinned BEHIND the next chunk's
      // tcgen05.ld in the instruction stream (left alone, the compiler sinks the load below the exponentials - the
      // arithmetic does not depend on it - and the two never overlap)
      // The next load is issued unconditionally (past the end it re-reads the last chunk into the idle buffer): under
      // an `if` the compiler moves the whole conditional block below the exponentials.
      const int last = total - 1;
      for (int g = 0; g < total; g += 2) {
        ld_wait(A);
        issue(g + 1 < total ? g + 1 : last, B);
        reg_fence32(A);
        consume(g, A);
        if (g + 1 < total) {
          ld_wait(B);
          issue(g + 2 < total ? g + 2 : last, A);
          reg_fence32(B);
          consume(g + 1, B);
        }
      }
      tmem_ld_wait();  // the trailing dummy load
      // merge the two column halves of each row (half 1 -> shared -> half 0), then store
      if (half == 1) comb[row] = make_float2(m, l);
      named_bar_sync(1, ATT_SM_THREADS);
      if (half == 0 && row_ok) {
        const float2 o = comb[row];
        const float M = fmaxf(m, o.x);
        const float L = l * ex2(m - M) + o.y * ex2(o.x - M);
        p.part_out[(long long)split * p.part_stride + (long long)bh * p.Tx + x0 + row] = make_float2(M, L);
      }
    } else {
      // P is written with a 2^-11 relative boost so that the tensor core's truncation of the fp32
      // bit pattern to TF32 acts as round-to-nearest (unbiased), at no instruction cost
      const float rnd = 7.0436e-4f;  // log2(1 + 2^-11)
      float mu_row = 0.f;
      if constexpr (PER_ROW) {
        if (row_ok) mu_row = merge_partials(p.part_in, p.part_stride, p.mu_splits, (long long)bh * p.Tmu + x0 + row) - rnd;
      } else {
        // mu = max + log2(sum) - rnd of every key column of this pass, once
        for (int col = st; col < n_it * BN; col += ATT_SM_THREADS) {
          const int gcol = it0 * BN + col;
          float mu = INFINITY;  // masked key column: exp2(-inf) = 0
          if (gcol < p.Ty) mu = merge_partials(p.part_in, p.part_stride, p.mu_splits, (long long)bh * p.Tmu + gcol) - rnd;
          mu_all[col] = mu;
        }
        named_bar_sync(1, ATT_SM_THREADS);
      }
      auto consume = [&](int g, uint32_t (&v)[32]) {
        const int it = g / NCH, k = g - it * NCH;
        if constexpr (PER_ROW) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float xj = fmaf(__uint_as_float(v[j]), c, -mu_row);
            v[j] = __float_as_uint((POLY && (j & 3) == 3) ? ex2_poly(xj) : ex2(xj));
          }
        } else {
          const float4* mup = reinterpret_cast<const float4*>(mu_all + it * BN + cb + k * 32);
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            const float4 m4 = mup[j4];
            v[4 * j4 + 0] = __float_as_uint(ex2(fmaf(__uint_as_float(v[4 * j4 + 0]), c, -m4.x)));
            v[4 * j4 + 1] = __float_as_uint(ex2(fmaf(__uint_as_float(v[4 * j4 + 1]), c, -m4.y)));
            v[4 * j4 + 2] = __float_as_uint(ex2(fmaf(__uint_as_float(v[4 * j4 + 2]), c, -m4.z)));
            const float x3 = fmaf(__uint_as_float(v[4 * j4 + 3]), c, -m4.w);
            v[4 * j4 + 3] = __float_as_uint(POLY ? ex2_poly(x3) : ex2(x3));
          }
        }
        tmem_st32(lane_base + (uint32_t)((it % 3) * BN + cb + k * 32), v);
        if (k == NCH - 1) {
          long long t0 = 0;
          if (tr) t0 = clock64();
          tmem_st_wait();
          tc_fence_before_sync();
          mbar_arrive(&p_full[it % 3]);
          if (tr) t_wait_st += clock64() - t0;
        }
      };
      uint32_t A[32], B[32];
      if (tr) t_begin = clock64();
      if (total > 0) issue(0, A);
      // reg_fence32 after each issue(): the consumers of the current buffer are pinned BEHIND the next chunk's
      // tcgen05.ld in the instruction stream (left alone, the compiler sinks the load below the exponentials - the
      // arithmetic does not depend on it - and the two never overlap)
      // The next load is issued unconditionally (past the end it re-reads the last chunk into the idle buffer): under
      // an `if` the compiler moves the whole conditional block below the exponentials.
      const int last = total - 1;
      for (int g = 0; g < total; g += 2) {
        ld_wait(A);
        issue(g + 1 < total ? g + 1 : last, B);
        reg_fence32(A);
        consume(g, A);
        if (g + 1 < total) {
          ld_wait(B);
          issue(g + 2 < total ? g + 2 : last, A);
          reg_fence32(B);
          consume(g + 1, B);
        }
      }
      tmem_ld_wait();  // the trailing dummy load
      // ---- epilogue: O (TMEM) -> merged [b][t][h*d + j]  (attention.mojo:61-62) ----
      mbar_wait(&o_full, 0);
      tc_fence_after_sync();
      const int bidx = bh / p.heads;
      float* orow = p.O + ((long long)bidx * p.Tx + x0 + row) * p.ldo + (long long)h * p.d;
      const int e0 = (((p.dpad >> 4) + 1) >> 1) * 16;
      const int ob = half ? e0 : 0, oe = half ? p.dpad : e0;
      for (int c0 = ob; c0 < oe; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(lane_base + (uint32_t)(3 * BN + c0), v);
        tmem_ld_wait();
        if (row_ok) {
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            if (c0 + 4 * j4 < p.d) {
              float4 o;
              if (p.round_out) {
                o = make_float4(__uint_as_float(rna_tf32_bits(__uint_as_float(v[4 * j4]))),
                                __uint_as_float(rna_tf32_bits(__uint_as_float(v[4 * j4 + 1]))),
                                __uint_as_float(rna_tf32_bits(__uint_as_float(v[4 * j4 + 2]))),
                                __uint_as_float(rna_tf32_bits(__uint_as_float(v[4 * j4 + 3]))));
              } else {
                o = make_float4(__uint_as_float(v[4 * j4]), __uint_as_float(v[4 * j4 + 1]),
                                __uint_as_float(v[4 * j4 + 2]), __uint_as_float(v[4 * j4 + 3]));
              }
              *reinterpret_cast<float4*>(orow + c0 + 4 * j4) = o;
            }
          }
        }
      }
    }
    if (TRACE && st == 0 && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1) && blockIdx.y == 0 && blockIdx.z == 0) {
      const long long t_all = clock64() - t_begin;
      printf("attn2 trace mode %d block %d: %d tiles of 128x%d, softmax loop %lld clk (%lld per tile): wait scores %lld, wait tmem ld %lld, wait tmem st + arrive %lld\n",
             MODE, blockIdx.x, n_it, BN, t_all, t_all / (n_it > 0 ? n_it : 1), t_wait_s, t_wait_ld, t_wait_st);
    }
    tc_fence_before_sync();
  }

  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

struct AttnPlan {
  int nbox, dpad, BN, sY, sV;
  size_t smem;
};

// tile of the streamed operand and ring depths that fit the 227 KiB shared-memory budget
AttnPlan plan_for(int d, int Ty, bool with_v, size_t extra = 0) {
  AttnPlan pl{};
  pl.nbox = (d + 31) / 32;
  pl.dpad = (d + 15) / 16 * 16;
  int BN = pl.nbox <= 2 ? 128 : 64;
  if (Ty < BN) BN = (Ty + 15) / 16 * 16;
  pl.BN = BN;
  const size_t xb = (size_t)pl.nbox * ATT_BM * ATT_BOX_ROW_BYTES;
  const size_t yb = (size_t)pl.nbox * BN * ATT_BOX_ROW_BYTES;
  const size_t budget = 212 * 1024 - extra;
  pl.sY = 2;
  pl.sV = with_v ? 2 : 0;
  if (xb + (pl.sY + pl.sV) * yb > budget && with_v) pl.sV = 1;
  if (xb + (pl.sY + pl.sV) * yb > budget) pl.sY = 1;
  pl.smem = xb + (size_t)(pl.sY + pl.sV) * yb + extra + 1024;
  return pl;
}

int tmap3(Ctx* c, CUtensorMap* tm, const float* base, int d, int T, long long nb, int box_rows, int atom32 = 0) {
  uint64_t dims[3] = {(uint64_t)d, (uint64_t)T, (uint64_t)nb};
  uint64_t str[3] = {1, (uint64_t)d, (uint64_t)T * d};
  uint32_t box[3] = {32, (uint32_t)box_rows, 1};
  return make_tmap_f32(c, tm, base, 3, dims, str, box, atom32, nullptr);
}

cudaError_t launch_attn(const CUtensorMap& tmX, const CUtensorMap& tmY, const CUtensorMap& tmV,
                        const AttnKParams& p, dim3 grid, size_t smem, cudaStream_t stream) {
  int max_dyn = 0;
  {
    cudaError_t e = optin_dyn_smem(reinterpret_cast<const void*>(attn_kernel), -1, &max_dyn);
    if (e != cudaSuccess) return e;
  }
  if ((long long)smem > max_dyn) return cudaErrorInvalidConfiguration;
  return launch_pdl(attn_kernel, grid, dim3(ATT_THREADS), smem, stream, tmX, tmY, tmV, p);
}

template <int MODE, bool PER_ROW, int NCH, bool TRACE, bool POLY>
cudaError_t launch_attn2_t(const CUtensorMap& tmX, const CUtensorMap& tmY, const CUtensorMap& tmV, const AttnKParams& p,
                           dim3 grid, size_t smem, cudaStream_t stream) {
  int max_dyn = 0;
  cudaError_t e = optin_dyn_smem(reinterpret_cast<const void*>(attn2_kernel<MODE, PER_ROW, NCH, TRACE, POLY>), -1, &max_dyn);
  if (e != cudaSuccess) return e;
  if ((long long)smem > max_dyn) return cudaErrorInvalidConfiguration;
  return launch_pdl(attn2_kernel<MODE, PER_ROW, NCH, TRACE, POLY>, grid, dim3(ATT_THREADS), smem, stream, tmX, tmY, tmV, p);
}
cudaError_t launch_attn2(const CUtensorMap& tmX, const CUtensorMap& tmY, const CUtensorMap& tmV, const AttnKParams& p,
                         dim3 grid, size_t smem, cudaStream_t stream, bool trace, bool poly) {
  const int nch = p.BN / 64;
  // POLY (every fourth exponential as an FMA-pipe polynomial, ex2_poly) is compiled out: measured on B200 it does not
  // shorten either pass (T = 4096, d = 40: 122.6 us without, 123.2 us with) - the statistics loop is bound by
  // instruction issue (~5 instructions per score), not by the MUFU queue, and the polynomial adds six more.
  (void)poly;
#define TSD_A2(MODE, PR, NCH)                                                                                    \
  (trace ? launch_attn2_t<MODE, PR, NCH, true, false>(tmX, tmY, tmV, p, grid, smem, stream)                        \
         : launch_attn2_t<MODE, PR, NCH, false, false>(tmX, tmY, tmV, p, grid, smem, stream))
  if (p.mode == ATT_STATS) return nch == 2 ? TSD_A2(ATT_STATS, false, 2) : TSD_A2(ATT_STATS, false, 1);
  if (p.mu_per_row) return nch == 2 ? TSD_A2(ATT_APPLY, true, 2) : TSD_A2(ATT_APPLY, true, 1);
  return nch == 2 ? TSD_A2(ATT_APPLY, false, 2) : TSD_A2(ATT_APPLY, false, 1);
#undef TSD_A2
}

// v2 takes whole tiles of 64 or 128 streamed rows; its three S / P buffers and the O tile must fit the 512 columns
bool attn2_ok(const AttnPlan& pl, int Ty) { return (pl.BN == 64 || pl.BN == 128) && Ty % pl.BN == 0 && 3 * pl.BN + pl.dpad <= 512; }
int attn2_tmem_cols(const AttnPlan& pl, bool apply) {
  const int need = 3 * pl.BN + (apply ? pl.dpad : 0);
  int c = 32;
  while (c < need) c <<= 1;
  return c;
}

}  // namespace

bool attention_fused_supported(int d, int causal) { return !causal && d >= 8 && d % 8 == 0 && d <= 160; }

int attention_fused(Ctx* c, const AttnArgs& a) {
  if (!attention_fused_supported(a.d, a.causal)) return c->fail(TSD_ERR_INVALID, "attention(fused): unsupported head dim");
  const long long kv_full = (long long)a.heads * a.Tk * a.d;
  if (a.kv_batch_stride != 0 && a.kv_batch_stride != kv_full)
    return c->fail(TSD_ERR_INVALID, "attention(fused): K/V batch stride must be 0 (shared) or heads*Tk*d");
  const int kv_shared = (a.kv_batch_stride == 0 && a.batch > 1) ? 1 : 0;
  const int BH = a.batch * a.heads;
  const long long kv_nb = kv_shared ? a.heads : BH;
  const float cs = 1.4426950408889634f / sqrtf((float)a.d);
  const bool query_axis = a.softmax_axis == 0;
  const bool attn_v2 = c->attn_v2 != 0;
  static int attn_trace_env = -1;  // TSD_ATTN_TRACE=1: in-kernel cycle counts of the v2 softmax loops (lab)
  if (attn_trace_env < 0) {
    const char* v = getenv("TSD_ATTN_TRACE");
    attn_trace_env = v ? atoi(v) : 0;
  }
  const bool attn_trace = attn_trace_env != 0;

  // ---- pass 1: statistics --------------------------------------------------------------------
  // query axis: one (max, sum) per key column over all queries  -> X = K, Y = Q
  // key axis:   one (max, sum) per query row over all keys       -> X = Q, Y = K
  const int Tx = query_axis ? a.Tk : a.Tq, Ty = query_axis ? a.Tq : a.Tk;
  const AttnPlan sp = plan_for(a.d, Ty, false);
  const int x_tiles = (Tx + ATT_BM - 1) / ATT_BM;
  const int y_tiles = (Ty + sp.BN - 1) / sp.BN;
  int splits = 1;
  {
    // fill the machine when the stationary side is short (cross-attention: 77 keys = 1 tile)
    // (a function of the per-image shape only: results do not depend on the batch size)
    const long long ctas = (long long)x_tiles * a.heads;
    if (ctas < c->sm_count) {
      splits = (int)((c->sm_count + ctas - 1) / ctas);
      if (splits > y_tiles) splits = y_tiles;
      if (splits > 32) splits = 32;
      if (splits < 1) splits = 1;
    }
  }
  int tps = (y_tiles + splits - 1) / splits;
  splits = (y_tiles + tps - 1) / tps;

  const size_t mark = c->arena.mark();
  float2* part = c->arena.alloc_n<float2>((size_t)splits * BH * Tx);
  if (!part) return c->fail(TSD_ERR_OOM, "attention(fused): arena exhausted (softmax statistics)");
  if (c->dry_run) {
    c->arena.release_to(mark);
    return TSD_OK;
  }
  const double core_flops = 4.0 * BH * (double)a.Tq * a.Tk * a.d;
  {
    TimedScope ts(c, FAM_ATTN, core_flops);
    CUtensorMap tmX, tmY;
    const float* X = query_axis ? a.K : a.Q;
    const float* Y = query_axis ? a.Q : a.K;
    const long long x_nb = query_axis ? kv_nb : BH, y_nb = query_axis ? BH : kv_nb;
    int rc = tmap3(c, &tmX, X, a.d, Tx, x_nb, ATT_BM);
    if (rc) return rc;
    rc = tmap3(c, &tmY, Y, a.d, Ty, y_nb, sp.BN);
    if (rc) return rc;
    AttnKParams p{};
    p.mode = ATT_STATS;
    p.Tx = Tx; p.Ty = Ty; p.d = a.d; p.nbox = sp.nbox; p.dpad = sp.dpad; p.BN = sp.BN;
    p.sY = sp.sY; p.sV = 1; p.heads = a.heads;
    p.x_shared = query_axis ? kv_shared : 0;
    p.y_shared = query_axis ? 0 : kv_shared;
    p.tiles_per_split = tps; p.total_tiles = y_tiles;
    p.tmem_cols = 256;
    p.c = cs;
    p.part_out = part;
    p.part_stride = (long long)BH * Tx;
    const bool v2s = attn_v2 && attn2_ok(sp, Ty);
    if (v2s) p.tmem_cols = attn2_tmem_cols(sp, false);
    rc = c->check(v2s ? launch_attn2(tmX, tmY, tmY, p, dim3(x_tiles, BH, splits), sp.smem, c->stream, attn_trace, c->attn_poly != 0)
                      : launch_attn(tmX, tmY, tmY, p, dim3(x_tiles, BH, splits), sp.smem, c->stream),
                  "attn_kernel (stats) launch");
    if (rc) return rc;
    c->launches++;

    // ---- pass 2: P = exp2(c S - mu), O = P V ----------------------------------------------------
    AttnPlan ap = plan_for(a.d, a.Tk, true);
    bool v2a = attn_v2 && attn2_ok(ap, a.Tk);
    if (v2a && query_axis) {
      // mu of every key staged in shared memory once (4 B per key)
      const size_t mu_bytes = ((size_t)a.Tk * 4 + 1023) / 1024 * 1024;
      const AttnPlan ap2 = plan_for(a.d, a.Tk, true, mu_bytes);
      if (mu_bytes <= 32 * 1024 && attn2_ok(ap2, a.Tk)) ap = ap2;
      else v2a = false;
    }
    CUtensorMap tmQ, tmK, tmV;
    rc = tmap3(c, &tmQ, a.Q, a.d, a.Tq, BH, ATT_BM);
    if (rc) return rc;
    rc = tmap3(c, &tmK, a.K, a.d, a.Tk, kv_nb, ap.BN);
    if (rc) return rc;
    rc = tmap3(c, &tmV, a.V, a.d, a.Tk, kv_nb, ap.BN, 1);  // MN-major tf32 operand: 128B swizzle, 32B atoms
    if (rc) return rc;
    AttnKParams q{};
    q.mode = ATT_APPLY;
    q.Tx = a.Tq; q.Ty = a.Tk; q.d = a.d; q.nbox = ap.nbox; q.dpad = ap.dpad; q.BN = ap.BN;
    q.sY = ap.sY; q.sV = ap.sV; q.heads = a.heads;
    q.x_shared = 0; q.y_shared = kv_shared; q.v_shared = kv_shared;
    q.total_tiles = (a.Tk + ap.BN - 1) / ap.BN;
    q.tiles_per_split = q.total_tiles;
    q.tmem_cols = 512;
    q.c = cs;
    q.part_in = part;
    q.mu_splits = splits;
    q.mu_per_row = query_axis ? 0 : 1;
    q.Tmu = Tx;
    q.part_stride = (long long)BH * Tx;
    q.O = a.O;
    q.ldo = a.heads * a.d;
    q.round_out = 1;
    {
      const char* dbg = getenv("TSD_ATTN_DEBUG");
      q.debug = dbg ? atoi(dbg) : 0;
      if (q.debug == 3) {  // return the statistics instead of the output
        size_t nb = sizeof(float2) * (size_t)splits * BH * Tx, cap = sizeof(float) * (size_t)a.batch * a.Tq * a.heads * a.d;
        cudaMemcpyAsync(a.O, part, nb < cap ? nb : cap, cudaMemcpyDeviceToDevice, c->stream);
        c->arena.release_to(mark);
        return TSD_OK;
      }
    }
    if (v2a && !q.debug) q.tmem_cols = attn2_tmem_cols(ap, true);
    rc = c->check(v2a && !q.debug ? launch_attn2(tmQ, tmK, tmV, q, dim3((a.Tq + ATT_BM - 1) / ATT_BM, BH, 1), ap.smem, c->stream, attn_trace, c->attn_poly != 0)
                                  : launch_attn(tmQ, tmK, tmV, q, dim3((a.Tq + ATT_BM - 1) / ATT_BM, BH, 1), ap.smem, c->stream),
                  "attn_kernel (apply) launch");
    if (rc) return rc;
    c->launches++;
  }
  c->arena.release_to(mark);  // stream-ordered reuse
  return TSD_OK;
}

}  // namespace tsd
