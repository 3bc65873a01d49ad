// K1/K2 kernel bodies - see gemm_tcgen05.cuh for the contract.
#include "gemm_tcgen05.cuh"

#include <cstdio>
#include <map>
#include <mutex>

#include "pdl.cuh"
#include "ptx_sm100.cuh"

namespace tsd {

// Lab instrumentation: in-kernel cycle stamps (gemm_debug bit 3 prints them).  Compiled in only with -DTSD_LAB_TRACE
// (csrc/build.sh honours TSD_LAB_TRACE=1): the stamps re-read %tid and the clock in the hot loops.
#ifdef TSD_LAB_TRACE
#define TSD_TRACE(cond, slot) do { if (cond) tr[slot] = clock64(); } while (0)
#else
#define TSD_TRACE(cond, slot) do { } while (0)
#endif

namespace {

__device__ __forceinline__ float gelu_tanh(float x) {
  // reference Gelu.forward, helpers/utils.mojo:1908-1919 (tanh form)
  const float k = 0.7978845608028654f;  // sqrt(2/pi)
  const float u = x * fmaf(x * x, 0.044715f * k, k);  // k (x + 0.044715 x^3)
  float t;
  asm("tanh.approx.f32 %0, %1;\n" : "=f"(t) : "f"(u));  // MUFU.TANH, rel. error 2^-11 (TF32 level)
  const float hx = 0.5f * x;
  return fmaf(hx, t, hx);
}

// Round to nearest (ties away) at TF32 precision.  ptxas expands cvt.rna.tf32.f32 into this add-and-mask plus an
// Inf/NaN guard per element; the unguarded form gives the same bits for every finite input, keeps +-Inf, turns an
// overflowing finite value into Inf as rounding must, and keeps NaN a NaN (except the all-ones payload).
__device__ __forceinline__ float round_tf32(float x) {
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
}

}  // namespace

// ---- raw shared-address forms of the PTX wrappers (addresses precomputed outside the loops) ----
__device__ __forceinline__ bool mbar_try_wait_a(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __noinline__ void mbar_timeout_trap() {
  printf("tsd: gemm mbarrier wait timed out (block %d,%d,%d thread %d)\n", blockIdx.x, blockIdx.y, blockIdx.z,
         threadIdx.x);
  __trap();
}
// bounded wait: a protocol bug surfaces as a trapped kernel, never as a hung GPU
__device__ __forceinline__ void mbar_wait_a(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait_a(bar, parity)) return;
  uint32_t spins = 0;
  while (!mbar_try_wait_a(bar, parity))
    if (++spins > (1u << 26)) mbar_timeout_trap();
}
__device__ __forceinline__ void mbar_expect_tx_a(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
template <int CG>
__device__ __forceinline__ void tma_a_4d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1, int c2, int c3) {
  if constexpr (CG == 1) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6}], [%2];\n" ::"r"(dst),
        "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
  } else {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6}], [%2];\n" ::"r"(dst),
        "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
  }
}
// merged K atoms: coordinate 0 is always 0 (the 32 channels of a chunk), the last coordinate is the chunk index
template <int CG>
__device__ __forceinline__ void tma_a_5d(uint32_t dst, const void* tmap, uint32_t bar, int c1, int c2, int c3, int c4) {
  if constexpr (CG == 1) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];\n" ::"r"(dst),
        "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
  } else {
    asm volatile(
        "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];\n" ::"r"(dst),
        "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
  }
}
template <int CG>
__device__ __forceinline__ void tma_b_4d(uint32_t dst, const void* tmap, uint32_t bar, int c1, int c2, int c3) {
  if constexpr (CG == 1) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6}], [%2];\n" ::"r"(dst),
        "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
  } else {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6}], [%2];\n" ::"r"(dst),
        "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
  }
}
template <int CG>
__device__ __forceinline__ void tma_b_3d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1, int c2) {
  if constexpr (CG == 1) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5}], [%2];\n" ::"r"(dst),
        "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
  } else {
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5}], [%2];\n" ::"r"(dst),
        "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
  }
}
template <int CG>
__device__ __forceinline__ void umma_tf32_cg(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  if constexpr (CG == 1) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
// all MMAs issued so far arrive on `bar` when complete; CG == 2: on the same barrier of both CTAs
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r));
  return r;
}
template <int CG>
__device__ __forceinline__ void umma_commit_cg(uint32_t bar) {
  if constexpr (CG == 1) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(bar)
                 : "memory");
  } else {
    // both CTAs of the pair: cluster ranks 2k and 2k+1 (the cluster may hold several pairs when it also spans K splits)
    const uint16_t mask = (uint16_t)(3u << (cluster_ctarank() & ~1u));
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n" ::"r"(bar),
        "h"(mask)
        : "memory");
  }
}
__device__ __forceinline__ uint32_t cluster_ctaid_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctaid.x;\n" : "=r"(r));
  return r;
}
// float2 at shared-window address `laddr` of the CTA with rank `cta` of this cluster (distributed shared memory)
__device__ __forceinline__ float2 ld_dsmem_f2(uint32_t laddr, uint32_t cta) {
  uint32_t raddr;
  float2 v;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(raddr) : "r"(laddr), "r"(cta));
  asm volatile("ld.shared::cluster.v2.f32 {%0, %1}, [%2];\n" : "=f"(v.x), "=f"(v.y) : "r"(raddr) : "memory");
  return v;
}
// arrive (release, cluster scope) on the mbarrier at shared-window address `laddr` of CTA `cta` of this cluster
__device__ __forceinline__ float4 ld_dsmem_f4(uint32_t laddr, uint32_t cta) {
  uint32_t raddr;
  float4 v;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(raddr) : "r"(laddr), "r"(cta));
  asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];\n" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(raddr) : "memory");
  return v;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t laddr, uint32_t cta) {
  uint32_t raddr;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(raddr) : "r"(laddr), "r"(cta));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];\n" ::"r"(raddr) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait_cluster(bar, parity))
    if (++spins > (1u << 26)) mbar_timeout_trap();
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}

// Store phase of one 32-column chunk: this lane's 8 rows x 4 columns, staged tile -> (scale, bias, residual,
// TF32 rounding) -> global, plus the per-column (sum, sum^2) of what was stored.  Compile-time variants keep the
// row loop free of branches and of re-derived predicates: measured, the branchy runtime form cost ~130 cycles per
// row (a constant-bank reload and a branch per row, nothing overlapping) against ~25 here.
//   FULL  - all 8 rows of every lane of the warp lie inside the image (no row predicate at all)
//   RES / ROUND / STATS - residual add, TF32 rounding of the output, column statistics
template <bool FULL, bool RES, bool ROUND, bool STATS>
__device__ __forceinline__ void store_chunk_rows(const float* __restrict__ st, float sc, float4 b4,
                                                 float* const (&drow)[8], const float* const (&rrow)[8],
                                                 long long col, int n, float4& ss, float4& qq) {
  float4 t[8], rr[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) t[i] = *reinterpret_cast<const float4*>(st + i * (4 * 36));
  if (RES) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      rr[i] = (FULL || drow[i] != nullptr) ? *reinterpret_cast<const float4*>(rrow[i] + n) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float4 u = t[i];
    u.x = fmaf(u.x, sc, b4.x); u.y = fmaf(u.y, sc, b4.y); u.z = fmaf(u.z, sc, b4.z); u.w = fmaf(u.w, sc, b4.w);
    if (RES) {
      u.x += rr[i].x; u.y += rr[i].y; u.z += rr[i].z; u.w += rr[i].w;
    }
    if (ROUND) {
      u.x = round_tf32(u.x); u.y = round_tf32(u.y); u.z = round_tf32(u.z); u.w = round_tf32(u.w);
    }
    if (FULL) {
#ifndef TSD_LAB_NOSTORE
      *reinterpret_cast<float4*>(drow[i] + col) = u;
#else
      if (u.x == 1.2345e-30f) *reinterpret_cast<float4*>(drow[i] + col) = u;  // lab: keeps the math alive, never stores
#endif
      if (STATS) {
        ss.x += u.x; ss.y += u.y; ss.z += u.z; ss.w += u.w;
        qq.x = fmaf(u.x, u.x, qq.x); qq.y = fmaf(u.y, u.y, qq.y); qq.z = fmaf(u.z, u.z, qq.z); qq.w = fmaf(u.w, u.w, qq.w);
      }
    } else if (drow[i] != nullptr) {
      *reinterpret_cast<float4*>(drow[i] + col) = u;
      if (STATS) {
        ss.x += u.x; ss.y += u.y; ss.z += u.z; ss.w += u.w;
        qq.x = fmaf(u.x, u.x, qq.x); qq.y = fmaf(u.y, u.y, qq.y); qq.z = fmaf(u.z, u.z, qq.z); qq.w = fmaf(u.w, u.w, qq.w);
      }
    }
  }
}
// Ragged tiles (some rows outside the image / past M): one generic variant with run-time switches and row predicates.
__device__ __forceinline__ void store_chunk_rows_ragged(const float* __restrict__ st, float sc, float4 b4,
                                                     float* const (&drow)[8], const float* const (&rrow)[8],
                                                     long long col, int n, uint32_t flags, float4& ss, float4& qq) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    if (drow[i] == nullptr) continue;
    float4 u = *reinterpret_cast<const float4*>(st + i * (4 * 36));
    u.x = fmaf(u.x, sc, b4.x); u.y = fmaf(u.y, sc, b4.y); u.z = fmaf(u.z, sc, b4.z); u.w = fmaf(u.w, sc, b4.w);
    if (flags & 2u) {
      const float4 r = *reinterpret_cast<const float4*>(rrow[i] + n);
      u.x += r.x; u.y += r.y; u.z += r.z; u.w += r.w;
    }
    if (flags & 4u) {
      u.x = round_tf32(u.x); u.y = round_tf32(u.y); u.z = round_tf32(u.z); u.w = round_tf32(u.w);
    }
    *reinterpret_cast<float4*>(drow[i] + col) = u;
    ss.x += u.x; ss.y += u.y; ss.z += u.z; ss.w += u.w;
    qq.x = fmaf(u.x, u.x, qq.x); qq.y = fmaf(u.y, u.y, qq.y); qq.z = fmaf(u.z, u.z, qq.z); qq.w = fmaf(u.w, u.w, qq.w);
  }
}

// ===================== epilogue (8 warps: kernel warps 2..9) =====================
// TMEM -> registers (thread = output row) -> per-warp shared staging tile -> coalesced global
// stores (8 lanes x 16 B cover 128 B of one row; a warp instruction writes 4 full rows).
// Two warps per TMEM lane quadrant, interleaved over the 32-column chunks.  The staging tiles
// reuse the operand ring: every MMA has retired once the accumulator barrier completes.
// Shared by the GEMM / implicit-GEMM kernel and the halo convolution kernel.
struct EpilogueCtx {
  uint8_t* ring;        // 1024 B aligned start of the (idle) operand ring: staging tiles + column statistics
  uint32_t tmem_d;      // accumulator tile(s)
  uint32_t accum_bar;   // shared address of the accumulator-complete barrier
  int n_iters;          // > 1: two accumulator tiles (even / odd K steps) to be added
  int nt, img, h0, w0, batch, split;
  long long* tr;        // lab trace slots
  uint32_t red_bar = 0;  // shared address of two mbarriers (count = splits) of the cluster split-K reduction
  uint32_t pair_rank = 0, cg = 1;
};
__device__ __forceinline__ void gemm_epilogue(const GemmKParams& p, const EpilogueCtx& ec) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* const smem_ring = ec.ring;
  const uint32_t tmem_d = ec.tmem_d;
  const uint32_t accum_a = ec.accum_bar;
  const int n_iters = ec.n_iters;
  const int nt = ec.nt, img = ec.img, h0 = ec.h0, w0 = ec.w0, batch = ec.batch, split = ec.split;
  [[maybe_unused]] long long* const tr = ec.tr;
  pdl_wait();                // before the first global access of this role (row bias, residual, D)
  const int ew = warp - 2;   // 0..7
  const int q = warp & 3;    // TMEM lane quadrant this warp may read
  const int half = ew >> 2;  // which of the two warps of the quadrant
  constexpr int ST = 36;     // staging row stride in floats (conflict-free float4 rows)
  float* stg = reinterpret_cast<float*>(smem_ring) + ew * (32 * ST);
  const uint32_t trow = tmem_d + ((uint32_t)(q * 32) << 16);
  const bool partial = p.partial != nullptr;
  const bool geglu = p.geglu && !partial;
  const int out_cols = geglu ? (p.BN >> 1) : p.BN;
  const int n0 = nt * out_cols;
  const int n_valid = p.n_valid;
  const int n_lim = partial ? p.n_pad : n_valid;
  const bool img_ok = img < p.imgs;

  // the row this lane owns while in the thread = row layout (row bias / alpha / GEGLU)
  float rb = 0.0f;
  if (p.row_bias != nullptr && !partial) {
    const int r_own = q * 32 + lane;
    const int lh = r_own / p.bw, lw = r_own - lh * p.bw;
    const int h = h0 + lh, w = w0 + lw;
    if (img_ok && lh < p.bh && h < p.H && w < p.W) rb = p.row_bias[((long long)img * p.H + h) * p.W + w];
  }
  const float* cbias = (p.bias && !partial) ? p.bias + (long long)img * p.bias_img_stride : nullptr;
  const float alpha = partial ? 1.0f : p.alpha;
  const bool plain = (alpha == 1.0f) && (p.row_bias == nullptr || partial);

  // the 8 rows this lane stores (row = q*32 + i*4 + lane/8), 4 consecutive columns at lane%8*4
  const int sub = lane >> 3, c4 = (lane & 7) * 4;
  float* dbase;
  long long ldd;
  if (partial) {
    dbase = p.partial + ((long long)split * (gridDim.z / p.splits) + batch) * p.m_per_batch * p.n_pad;
    ldd = p.n_pad;
  } else {
    dbase = p.D + (long long)batch * p.d_batch_stride;
    ldd = p.ldd;
  }
  const float* rbase = (p.residual && !partial) ? p.residual + (long long)batch * p.r_batch_stride : nullptr;
  float* drow[8];
  const float* rrow[8];
  long long grow[8];  // global output row of each of this lane's 8 rows (-1: outside the image)
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = q * 32 + i * 4 + sub;
    const int lh = r / p.bw, lw = r - lh * p.bw;
    const int h = h0 + lh, w = w0 + lw;
    const bool ok = img_ok && (lh < p.bh) && (h < p.H) && (w < p.W);
    const long long g = ((long long)img * p.H + h) * p.W + w;
    grow[i] = ok ? g : -1;
    drow[i] = ok ? dbase + g * ldd : nullptr;
    rrow[i] = rbase ? rbase + g * p.ldr : nullptr;
  }
  bool rows_all = true;
#pragma unroll
  for (int i = 0; i < 8; ++i) rows_all = rows_all && (drow[i] != nullptr);
  const bool rows_full = __all_sync(0xffffffffu, rows_all) != 0;  // warp-uniform: no row predicates in the store phase
  const bool routed = !partial && p.split_n < (1 << 30);
  const bool vec_ok = partial || (((n_valid | p.ldd | p.split_n) & 3) == 0 && (p.split_stride & 3) == 0 &&
                                  (rbase == nullptr || (p.ldr & 3) == 0) &&
                                  (cbias == nullptr || (((p.bias_img_stride & 3) == 0) &&
                                                        (reinterpret_cast<uintptr_t>(p.bias) & 15) == 0)));
  const bool round_out = p.round_tf32 && !partial;
  // LayerNorm of the A operand folded into this epilogue: scale r and shift -r*mu*rowsum(W)
  float ln_r = 1.0f, ln_rmu = 0.0f;
  const bool fixup = p.fixup != 0 && partial;  // split-K: the last CTA of a tile reduces and runs the real epilogue
  const bool ln_fold = p.ln.partial != nullptr && (!partial || fixup);
  if (ln_fold) {
    __shared__ float2 ln_st[1];
    const int img_ln = (int)(((long long)blockIdx.x * GEMM_BM) / p.ln_rows_per_img);
    norm_stats_fold(p.ln, img_ln < p.ln.imgs ? img_ln : 0, threadIdx.x - 64, 256, ln_st);
    named_bar_sync(3, 256);
    ln_r = ln_st[0].y;
    ln_rmu = ln_st[0].x * ln_st[0].y;
  }
  // Effective column bias of this tile, bias[n] - r*mu*rowsum(W)[n] (LayerNorm fold), staged in shared memory
  // while the K loop is still running: the thread = row GEGLU math reads it as broadcast LDS.128 and the store
  // phase as one LDS.128 per chunk, instead of global loads on the critical path after the accumulator barrier.
  // GEGLU: [0, out_cols) value half, [128, 128 + out_cols) gate half.
  __shared__ __align__(16) float eb[256];
  const bool use_eb = !partial || fixup;
  if (use_eb) {
    const int te = threadIdx.x - 64;
    const float* cb = p.bias ? p.bias + (long long)img * p.bias_img_stride : nullptr;
    float e = 0.0f;
    int nsrc = -1;
    if (geglu) {
      const int hcol = te & 127;
      if (hcol < out_cols && n0 + hcol < n_valid) nsrc = (te >> 7) * p.n_half + n0 + hcol;
    } else if (te < p.BN && n0 + te < n_valid) {
      nsrc = n0 + te;
    }
    if (nsrc >= 0) {
      if (cb != nullptr) e = __ldg(cb + nsrc);
      if (ln_fold) e -= ln_rmu * __ldg(p.wsum + nsrc);
    }
    eb[te] = e;
    named_bar_sync(3, 256);
  }
  // producer-side norm statistics: per-column (sum, sum^2) of the stored values of this tile
  const bool want_stats = p.ns.partial != nullptr && (!partial || fixup);
  const bool stats_pass1 = want_stats && !partial;
  float2* cs = reinterpret_cast<float2*>(smem_ring + 8 * 32 * ST * 4);  // [4][256]

  const bool two_acc = n_iters > 1 && GEMM_ROLE_PAIRS > 1;
  const uint32_t trow1 = trow + (uint32_t)p.acc_stride;  // the odd-K-step issuer's accumulator
  // Every decision of the chunk loop is a bit of one register that the compiler cannot re-derive: left to itself it
  // re-read the kernel parameters from the constant bank and branched on them in every chunk (a chain of ~6 dependent
  // constant loads, ~300 cycles per chunk with only two epilogue warps per scheduler to hide them).
  enum : uint32_t { F_FULL = 1, F_RES = 2, F_ROUND = 4, F_STATS = 8, F_GEGLU = 16, F_TWOACC = 32, F_PLAIN = 64,
                    F_PARTIAL = 128, F_ROUTED = 256, F_VEC = 512, F_B4 = 1024, F_STATS1 = 2048, F_SMEMTILE = 4096 };
  // cluster split-K through distributed shared memory (p.fixup == 3): the raw partial tile of this CTA stays in its own
  // shared memory ([128 rows][BN rounded up to 32, + 4] floats behind the staging area of the idle operand ring); the CTAs of the tile read
  // each other's tiles with ld.shared::cluster after a cluster barrier - no partial tile ever reaches L2 / HBM
  const bool dsm = fixup && p.fixup == 3;
  const int ldt = ((p.BN + 31) & ~31) + 4;  // whole 32-column chunks (the last one may be partly padding) + 4: conflict-free rows
  float* const tile = reinterpret_cast<float*>(smem_ring + GEMM_DSM_TILE_OFFSET);
  uint32_t flags = (rows_full ? F_FULL : 0u) | (rbase != nullptr ? F_RES : 0u) | (round_out ? F_ROUND : 0u) |
                   (want_stats ? F_STATS : 0u) | (geglu ? F_GEGLU : 0u) | (two_acc ? F_TWOACC : 0u) |
                   (plain ? F_PLAIN : 0u) | (partial ? F_PARTIAL : 0u) | (routed ? F_ROUTED : 0u) |
                   (vec_ok ? F_VEC : 0u) | ((!geglu && !partial) ? F_B4 : 0u) | (stats_pass1 ? F_STATS1 : 0u) |
                   (dsm ? F_SMEMTILE : 0u);
  asm volatile("mov.b32 %0, %0;\n" : "+r"(flags));
  const float sc_store = (!geglu && !partial && ln_fold) ? ln_r : 1.0f;
  const long long partial_col0 = (long long)nt * p.BN;

  mbar_wait_a(accum_a, 0);
  tc_fence_after_sync();
  TSD_TRACE(threadIdx.x == 64, 4);

  for (int c = half * 32; c < out_cols; c += 64) {
    uint32_t v[32];
    tmem_ld32(trow + c, v);
    if (flags & F_TWOACC) {
      uint32_t v1[32];
      tmem_ld32(trow1 + c, v1);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(v1[j]));
    }
    if (flags & F_GEGLU) {
      uint32_t g[32];
      tmem_ld32(trow + out_cols + c, g);
      if (flags & F_TWOACC) {
        uint32_t g1[32];
        tmem_ld32(trow1 + out_cols + c, g1);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) g[j] = __float_as_uint(__uint_as_float(g[j]) + __uint_as_float(g1[j]));
      }
      TSD_TRACE(threadIdx.x == 64 && c == 0, 8);
      tmem_ld_wait();
      TSD_TRACE(threadIdx.x == 64 && c == 0, 9);
      const float4* bo4 = reinterpret_cast<const float4*>(eb + c);
      const float4* bg4 = reinterpret_cast<const float4*>(eb + 128 + c);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 bo = bo4[j], bg = bg4[j];
        v[4 * j + 0] = __float_as_uint(fmaf(__uint_as_float(v[4 * j + 0]), ln_r, bo.x) * gelu_tanh(fmaf(__uint_as_float(g[4 * j + 0]), ln_r, bg.x)));
        v[4 * j + 1] = __float_as_uint(fmaf(__uint_as_float(v[4 * j + 1]), ln_r, bo.y) * gelu_tanh(fmaf(__uint_as_float(g[4 * j + 1]), ln_r, bg.y)));
        v[4 * j + 2] = __float_as_uint(fmaf(__uint_as_float(v[4 * j + 2]), ln_r, bo.z) * gelu_tanh(fmaf(__uint_as_float(g[4 * j + 2]), ln_r, bg.z)));
        v[4 * j + 3] = __float_as_uint(fmaf(__uint_as_float(v[4 * j + 3]), ln_r, bo.w) * gelu_tanh(fmaf(__uint_as_float(g[4 * j + 3]), ln_r, bg.w)));
      }
    } else {
      TSD_TRACE(threadIdx.x == 64 && c == 0, 8);
      tmem_ld_wait();
      TSD_TRACE(threadIdx.x == 64 && c == 0, 9);
      if (!(flags & F_PLAIN)) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(fmaf(__uint_as_float(v[j]), alpha, rb));
      }
    }
    TSD_TRACE(threadIdx.x == 64 && c == 0, 10);
    if (flags & F_SMEMTILE) {  // thread = row: 32 consecutive floats of this row, quarter-warps hit distinct banks (ldt % 32 == 4)
      uint4* trow_s = reinterpret_cast<uint4*>(tile + (q * 32 + lane) * ldt + c);
#pragma unroll
      for (int j = 0; j < 8; ++j) trow_s[j] = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      continue;
    }
    uint4* srow = reinterpret_cast<uint4*>(stg + lane * ST);
#pragma unroll
    for (int j = 0; j < 8; ++j) srow[j] = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    __syncwarp();
    TSD_TRACE(threadIdx.x == 64 && c == 0, 11);
    TSD_TRACE(threadIdx.x == 64 && c == 64, 13);
    const int n = n0 + c + c4;  // first of this lane's 4 columns
    float4 ss = make_float4(0.f, 0.f, 0.f, 0.f), qq = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c + c4 < out_cols && n < n_lim) {
      long long col;
      if (flags & F_PARTIAL) col = partial_col0 + c + c4;
      else if (flags & F_ROUTED) col = (long long)(n / p.split_n) * p.split_stride + (n % p.split_n);
      else col = n;
      if (flags & F_VEC) {
        float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (flags & F_B4) b4 = *reinterpret_cast<const float4*>(eb + c + c4);
        const float* st = stg + sub * ST + c4;
#define TSD_STORE_CASE(k)                                                                                      \
  case k:                                                                                                      \
    store_chunk_rows<true, ((k) & 1) != 0, ((k) & 2) != 0, ((k) & 4) != 0>(st, sc_store, b4, drow, rrow, col, n, ss, qq); \
    break;
        if (flags & F_FULL) {
          switch ((flags >> 1) & 7u) {
            TSD_STORE_CASE(0) TSD_STORE_CASE(1) TSD_STORE_CASE(2) TSD_STORE_CASE(3)
            TSD_STORE_CASE(4) TSD_STORE_CASE(5) TSD_STORE_CASE(6) TSD_STORE_CASE(7)
          }
        } else {
          store_chunk_rows_ragged(st, sc_store, b4, drow, rrow, col, n, flags, ss, qq);
        }
#undef TSD_STORE_CASE
      } else {
        // unaligned shapes (rare): scalar stores, output routing resolved once per column
        long long co[4];
        float badd[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int nn = n + j;
          co[j] = nn < n_valid ? (long long)(nn / p.split_n) * p.split_stride + (nn % p.split_n) : -1;
          badd[j] = (nn < n_valid && cbias != nullptr && !(flags & F_GEGLU)) ? cbias[nn] : 0.0f;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if (drow[i] == nullptr) continue;
          const float4 t = *reinterpret_cast<const float4*>(stg + (i * 4 + sub) * ST + c4);
          const float e[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (co[j] < 0) continue;
            float a = e[j] + badd[j];
            if (rbase != nullptr) a += rrow[i][n + j];
            if (flags & F_ROUND) a = round_tf32(a);
            drow[i][co[j]] = a;
          }
        }
      }
    }
    TSD_TRACE(threadIdx.x == 64 && c == 0, 12);
    TSD_TRACE(threadIdx.x == 64 && c == 64, 14);
    if (flags & F_STATS1) {
      // fold the 4 row groups of the warp (lanes with equal lane % 8): fixed xor tree, all lanes take part
#pragma unroll
      for (int o = 8; o <= 16; o <<= 1) {
        ss.x += __shfl_xor_sync(0xffffffffu, ss.x, o); ss.y += __shfl_xor_sync(0xffffffffu, ss.y, o);
        ss.z += __shfl_xor_sync(0xffffffffu, ss.z, o); ss.w += __shfl_xor_sync(0xffffffffu, ss.w, o);
        qq.x += __shfl_xor_sync(0xffffffffu, qq.x, o); qq.y += __shfl_xor_sync(0xffffffffu, qq.y, o);
        qq.z += __shfl_xor_sync(0xffffffffu, qq.z, o); qq.w += __shfl_xor_sync(0xffffffffu, qq.w, o);
      }
      if (sub == 0 && c + c4 < out_cols) {
        float2* d = cs + q * 256 + c + c4;
        d[0] = make_float2(ss.x, qq.x); d[1] = make_float2(ss.y, qq.y);
        d[2] = make_float2(ss.z, qq.z); d[3] = make_float2(ss.w, qq.w);
      }
    }
    __syncwarp();
  }
  bool stats_leader = true;  // this CTA writes the statistics of its tile (cluster split-K: the split-0 CTA only)
  if (fixup) {
    // ---- in-kernel split-K reduction: every CTA has stored its raw partial tile (L2-resident).
    //  fixup 1: the last CTA to arrive for an output tile sums the partials in split order and runs the real epilogue
    //           (serial tail, measured slower than the separate reduce kernel);
    //  fixup 2: the `splits` CTAs of a tile form a thread-block CLUSTER (launch attribute, co-scheduled by hardware):
    //           after a cluster-scope barrier among their epilogue warps each CTA reduces and finishes every
    //           splits-th 32-column chunk - the reduction runs `splits`-wide, no reduce kernel, no ticket, and the tile's
    //           norm statistics still come out of this epilogue (column sums gathered through distributed shared memory).
    __shared__ unsigned int fix_last;
    const int te0 = threadIdx.x - 64;
    int c_first = half * 32, c_stride = 64;
    const int S = p.splits;
    if (p.fixup == 1) {
      named_bar_sync(2, 256);
      if (te0 == 0) {
        __threadfence();
        unsigned int* tk = p.tile_tickets + (blockIdx.y * gridDim.x + blockIdx.x);
        const unsigned int old = atomicAdd(tk, 1u);
        fix_last = (old == (unsigned int)p.splits - 1) ? 1u : 0u;
        if (fix_last) *tk = 0u;  // self-resetting: launches on one stream are serialised
      }
      named_bar_sync(2, 256);
      if (!fix_last) return;
      __threadfence();
    } else {
      if (!dsm) __threadfence();  // this thread's partial stores are performed before the arrival below is observed
      named_bar_sync(2, 256);
      if (te0 < S) mbar_arrive_remote(ec.red_bar, ec.pair_rank + ec.cg * (uint32_t)te0);  // one arrival per CTA of the tile, on each of them
      mbar_wait_cluster(ec.red_bar, 0);
      c_first = (split + S * half) * 32;
      c_stride = 2 * S * 32;
      stats_leader = split == 0;
      if (want_stats) {  // columns this CTA does not reduce stay zero in its table
        for (int i = te0; i < 4 * 256; i += 256) cs[i] = make_float2(0.f, 0.f);
        named_bar_sync(2, 256);
      }
    }
    const float* cb2 = p.bias ? p.bias + (long long)img * p.bias_img_stride : nullptr;
    const float* ws0 = p.partial;
    const long long ws_split = (long long)p.m_per_batch * p.n_pad;  // floats between splits (batch == 1)
    for (int c = c_first; c < out_cols; c += c_stride) {
      const int n = n0 + c + c4;
      float4 ss = make_float4(0.f, 0.f, 0.f, 0.f), qq = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c + c4 < out_cols && n < n_valid) {
        float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (cb2 != nullptr) b4 = __ldg(reinterpret_cast<const float4*>(cb2 + n));
        float sc = 1.0f;
        if (ln_fold) {
          const float4 w4 = __ldg(reinterpret_cast<const float4*>(p.wsum + n));
          b4.x -= ln_rmu * w4.x; b4.y -= ln_rmu * w4.y; b4.z -= ln_rmu * w4.z; b4.w -= ln_rmu * w4.w;
          sc = ln_r;
        }
        // all loads of the chunk first (8 rows x splits), then the sums in split order
        float4 t[8];
        if (dsm) {
          // this lane's 8 rows of the chunk in the tile of split s2: CTA rank pair_rank + cg * s2 of the cluster
          const uint32_t a0 = smem_u32(tile + (q * 32 + sub) * ldt + c + c4);
#pragma unroll
          for (int i = 0; i < 8; ++i) t[i] = ld_dsmem_f4(a0 + (uint32_t)(i * 4 * ldt) * 4u, ec.pair_rank);
          for (int s2 = 1; s2 < S; ++s2) {
            float4 u[8];
            const uint32_t owner = ec.pair_rank + ec.cg * (uint32_t)s2;
#pragma unroll
            for (int i = 0; i < 8; ++i) u[i] = ld_dsmem_f4(a0 + (uint32_t)(i * 4 * ldt) * 4u, owner);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              t[i].x += u[i].x; t[i].y += u[i].y; t[i].z += u[i].z; t[i].w += u[i].w;
            }
          }
        } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          t[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (grow[i] >= 0) t[i] = __ldcg(reinterpret_cast<const float4*>(ws0 + grow[i] * p.n_pad + (long long)nt * p.BN + c + c4));
        }
        for (int s2 = 1; s2 < S; ++s2) {
          float4 u[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            u[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (grow[i] >= 0)
              u[i] = __ldcg(reinterpret_cast<const float4*>(ws0 + grow[i] * p.n_pad + (long long)nt * p.BN + c + c4 + s2 * ws_split));
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            t[i].x += u[i].x; t[i].y += u[i].y; t[i].z += u[i].z; t[i].w += u[i].w;
          }
        }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if (grow[i] < 0) continue;
          float4 v4 = t[i];
          v4.x = fmaf(v4.x, sc, b4.x); v4.y = fmaf(v4.y, sc, b4.y); v4.z = fmaf(v4.z, sc, b4.z); v4.w = fmaf(v4.w, sc, b4.w);
          if (p.residual != nullptr) {
            const float4 rr = *reinterpret_cast<const float4*>(p.residual + grow[i] * p.ldr + n);
            v4.x += rr.x; v4.y += rr.y; v4.z += rr.z; v4.w += rr.w;
          }
          if (p.round_tf32) {
            v4.x = round_tf32(v4.x); v4.y = round_tf32(v4.y); v4.z = round_tf32(v4.z); v4.w = round_tf32(v4.w);
          }
          *reinterpret_cast<float4*>(p.D + grow[i] * p.ldd + n) = v4;
          ss.x += v4.x; ss.y += v4.y; ss.z += v4.z; ss.w += v4.w;
          qq.x = fmaf(v4.x, v4.x, qq.x); qq.y = fmaf(v4.y, v4.y, qq.y); qq.z = fmaf(v4.z, v4.z, qq.z); qq.w = fmaf(v4.w, v4.w, qq.w);
        }
      }
      if (want_stats) {
#pragma unroll
        for (int o = 8; o <= 16; o <<= 1) {
          ss.x += __shfl_xor_sync(0xffffffffu, ss.x, o); ss.y += __shfl_xor_sync(0xffffffffu, ss.y, o);
          ss.z += __shfl_xor_sync(0xffffffffu, ss.z, o); ss.w += __shfl_xor_sync(0xffffffffu, ss.w, o);
          qq.x += __shfl_xor_sync(0xffffffffu, qq.x, o); qq.y += __shfl_xor_sync(0xffffffffu, qq.y, o);
          qq.z += __shfl_xor_sync(0xffffffffu, qq.z, o); qq.w += __shfl_xor_sync(0xffffffffu, qq.w, o);
        }
        if (sub == 0 && c + c4 < out_cols) {
          float2* d = cs + q * 256 + c + c4;
          d[0] = make_float2(ss.x, qq.x); d[1] = make_float2(ss.y, qq.y);
          d[2] = make_float2(ss.z, qq.z); d[3] = make_float2(ss.w, qq.w);
        }
      }
    }
    if (p.fixup >= 2 && want_stats) {
      // the column sums of a chunk live in the CTA that reduced it: the split-0 CTA gathers them through distributed
      // shared memory once every CTA of the tile has finished (second barrier), then writes the tile's statistics
      named_bar_sync(2, 256);
      if (te0 == 0) mbar_arrive_remote(ec.red_bar + 8u, ec.pair_rank);  // on the split-0 CTA of this tile
      if (!stats_leader) return;
      mbar_wait_cluster(ec.red_bar + 8u, 0);
      {
        const int te = te0;
        float2 tot = make_float2(0.f, 0.f);
        if (te < out_cols) {
          const uint32_t owner = ec.pair_rank + ec.cg * (uint32_t)((te >> 5) % S);
          const uint32_t a0 = smem_u32(cs + te);
          const float2 a = ld_dsmem_f2(a0, owner), b = ld_dsmem_f2(a0 + 256 * 8, owner), c2 = ld_dsmem_f2(a0 + 512 * 8, owner),
                       d2 = ld_dsmem_f2(a0 + 768 * 8, owner);
          tot = make_float2((a.x + b.x) + (c2.x + d2.x), (a.y + b.y) + (c2.y + d2.y));
        }
        named_bar_sync(2, 256);  // every remote read of this CTA's own table (owner == self) is done before it is overwritten
        cs[te] = tot;
        cs[256 + te] = make_float2(0.f, 0.f);
        cs[512 + te] = make_float2(0.f, 0.f);
        cs[768 + te] = make_float2(0.f, 0.f);
      }
    }
  }
  if (want_stats) {
    // quadrant sums -> column sums -> sums of the groups overlapping this tile -> partial; the last CTA finalises
    const int te = threadIdx.x - 64;  // 0..255 among the epilogue threads
    named_bar_sync(2, 256);
    {
      float2 a = cs[te], b = cs[256 + te], c2 = cs[512 + te], d2 = cs[768 + te];
      const bool okc = te < out_cols && n0 + te < n_valid;
      cs[te] = okc ? make_float2((a.x + b.x) + (c2.x + d2.x), (a.y + b.y) + (c2.y + d2.y)) : make_float2(0.f, 0.f);
    }
    named_bar_sync(2, 256);
    if (img_ok) {
      const int cpg = p.ns.cpg;
      const int g_lo = n0 / cpg;
      int n_end = n0 + out_cols;
      if (n_end > n_valid) n_end = n_valid;
      const int g_hi = (n_end - 1) / cpg;
      float2* dst = p.ns.partial + (long long)nt * p.ns.lg * p.ns.slabs_total + blockIdx.x;  // [n-tile][group][slab]
      for (int g = g_lo + ew; g <= g_hi; g += 8) {  // one warp per group: lanes stride over its columns in the tile
        int c_lo = g * cpg - n0, c_hi = (g + 1) * cpg - n0;
        if (c_lo < 0) c_lo = 0;
        if (c_hi > n_end - n0) c_hi = n_end - n0;
        float s = 0.f, qv = 0.f;
        for (int cc = c_lo + lane; cc < c_hi; cc += 32) {
          const float2 t = cs[cc];
          s += t.x;
          qv += t.y;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          s += __shfl_xor_sync(0xffffffffu, s, o);
          qv += __shfl_xor_sync(0xffffffffu, qv, o);
        }
        if (lane == 0) dst[(long long)(g - g_lo) * p.ns.slabs_total] = make_float2(s, qv);
      }
    }
  }
  TSD_TRACE(threadIdx.x == 64, 5);
}

// Role loops are single-thread instruction streams: every dependent SASS instruction costs ~4-6
// cycles and an mbarrier try_wait ~90, so (measured, profiles/r01_gemm_lab_notes.md) the first
// version spent ~500 cycles per 32-wide K step on bookkeeping while the MMAs of the step need
// 2*BN cycles.  Hence: K steps of 64 (two 128 B swizzle atoms per operand row, 8 MMAs per
// barrier round trip), incremental index math, shared addresses and descriptors precomputed.
//
// CG == 2: CTA pairs (cluster 2x1x1 over the M tiles).  Each CTA loads its own 128 A rows and
// HALF of the B rows; the leader issues tcgen05.mma.cta_group::2 (M = 256) which reads both
// CTAs' shared memory and writes both CTAs' TMEM - the B operand crosses L2->SM once per pair.
template <int CG>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tf32_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmA2, const GemmKParams p) {
  extern __shared__ uint8_t smem_dyn[];
  __shared__ __align__(8) uint64_t full_bar[GEMM_MAX_STAGES];
  __shared__ __align__(8) uint64_t empty_bar[GEMM_MAX_STAGES];
  __shared__ __align__(8) uint64_t fullb_bar[GEMM_MAX_B_STAGES];   // deep weight ring (p.b_stages > 0)
  __shared__ __align__(8) uint64_t emptyb_bar[GEMM_MAX_B_STAGES];
  __shared__ __align__(8) uint64_t accum_bar;
  __shared__ __align__(8) uint64_t red_bar[2];  // cluster split-K (p.fixup == 2): partial tiles stored / column sums ready
  __shared__ uint32_t tmem_slot;
  __shared__ long long tr[16];  // lab trace (debug bit 3)
  TSD_TRACE(threadIdx.x == 0, 0);

  const int warp = threadIdx.x >> 5;
  const uint32_t rank = (CG == 2) ? cluster_ctaid_x() : 0u;  // position in the CTA pair (the cluster may also span K splits along z)
  const bool cluster_split = p.fixup >= 2;  // the K splits of a tile are the z extent of this CTA's cluster

  // 1024 B aligned operand ring (SWIZZLE_128B atoms are 8 rows x 128 B)
  const uint32_t smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  const int b_rows = p.BN / CG;                        // B rows staged by this CTA
  const uint32_t a_atom = GEMM_BM * 128;               // 128 rows x 32 fp32
  const uint32_t b_atom = (uint32_t)b_rows * 128;
  const int bk = p.bk;            // K step: 64 (two 128 B atoms per operand row) or 32 (one)
  const uint32_t natoms = (uint32_t)bk >> 5;
  const uint32_t stage_bytes = natoms * (a_atom + b_atom);

  // tile coordinates
  const int nt = blockIdx.y;
  int mt = blockIdx.x;  // M tiles along x: a CTA pair is two consecutive M tiles (cluster 2x1x1)
  const int batch = blockIdx.z / p.splits;
  const int split = blockIdx.z - batch * p.splits;
  const int tw = mt % p.tiles_w;
  mt /= p.tiles_w;
  const int th = mt % p.tiles_h;
  const int img = mt / p.tiles_h;  // >= p.imgs for the phantom tile that pads an odd tile count to a pair
  const int w0 = tw * p.bw, h0 = th * p.bh;

  const int it_begin = split * p.iters_per_split;
  int it_end = it_begin + p.iters_per_split;
  if (it_end > p.total_iters) it_end = p.total_iters;
  const int n_iters = it_end - it_begin;
  const int num_stages = p.num_stages;
  const uint32_t full0 = smem_u32(&full_bar[0]), empty0 = smem_u32(&empty_bar[0]), accum_a = smem_u32(&accum_bar);
  const uint32_t fullb0 = smem_u32(&fullb_bar[0]), emptyb0 = smem_u32(&emptyb_bar[0]);
  const int nb_stages = GEMM_B_PRODUCER ? p.b_stages : 0;      // > 0: separate rings
  const uint32_t a_stage = natoms * a_atom, b_stage = natoms * b_atom;
  const uint32_t b_ring = smem_base + (uint32_t)num_stages * a_stage;  // deep mode: the weight ring follows the A ring

  if (threadIdx.x == 0) {
    for (int s = 0; s < num_stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < nb_stages; ++s) {
      mbar_init(&fullb_bar[s], 1);
      mbar_init(&emptyb_bar[s], 1);
    }
    mbar_init(&accum_bar, (n_iters > 1 && GEMM_ROLE_PAIRS > 1) ? 2 : 1);  // one commit per MMA issuer
    if (cluster_split) {
      mbar_init(&red_bar[0], (uint32_t)p.splits);
      mbar_init(&red_bar[1], (uint32_t)p.splits);
    }
    fence_barrier_init();
    prefetch_tensormap(&tmA);
    prefetch_tensormap(&tmB);
    if (p.cin2 > 0) prefetch_tensormap(&tmA2);
  }
  if (warp == 1) {
    if constexpr (CG == 1) {
      tmem_alloc(&tmem_slot, (uint32_t)p.tmem_cols);
      tmem_relinquish();
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_slot)),
                   "r"((uint32_t)p.tmem_cols)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;\n" ::: "memory");
    }
  }
  tc_fence_before_sync();
  if (CG == 2 || cluster_split) cluster_sync_all();  // the peers' barriers must exist before anything arrives on them
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_d = tmem_slot;
  TSD_TRACE(threadIdx.x == 0, 1);
  pdl_launch_dependents();  // resources are held: the next kernel may start its own prologue

  // Role warps: 0 / 10 = TMA producers, 1 / 11 = MMA issuers (1 also owns TMEM), 2..9 = epilogue.
  // Both role pairs run the same loop over every other K step (parity 0 / 1): the per-step barrier
  // round trip and the instruction issue of a single thread (~40 clk per MMA, ~65 per TMA) are what
  // bounds small-N tiles, so they are spread over two threads.  The two issuers accumulate into two
  // separate TMEM tiles (no ordering between them is needed); the epilogue adds them.
  const int role_parity = (GEMM_ROLE_PAIRS > 1 && warp >= 10) ? 1 : 0;
  const int n_par = (n_iters > 1 && GEMM_ROLE_PAIRS > 1) ? 2 : 1;  // issuers / producers with work
  // GEMM_B_PRODUCER: warp 0 = A boxes + barrier arming, warp 10 = B boxes (both walk every K step)
  const bool do_a = !GEMM_B_PRODUCER || warp == 0, do_b = !GEMM_B_PRODUCER || warp == 10;
  if ((warp == 0 || warp == 10) && role_parity < n_par) {
    // ===================== TMA producer (K steps it = parity, parity + 2, ...) =====================
    if (elect_one()) {
      const int cin = p.cin, debug = p.debug;
      const bool geglu = p.geglu != 0;
      const uint32_t a_tx = (debug & 2) ? 0u : natoms * (uint32_t)p.a_box_bytes;
      const uint32_t b_tx = (debug & 4) ? 0u : natoms * b_atom;
      const uint32_t tx = (uint32_t)CG * (a_tx + b_tx);  // CG == 2: the leader's barrier counts both CTAs' bytes
      const int c3 = img + batch;
      // first B row of this CTA's (first) box, and of the second (gate) box for single-CTA GEGLU
      int brow0, brow1 = 0;
      const int half_rows = p.BN >> 1;
      if (geglu) {
        if (CG == 1) {
          brow0 = nt * half_rows;
          brow1 = p.n_half + nt * half_rows;
        } else {
          brow0 = (rank == 0 ? 0 : p.n_half) + nt * half_rows;  // leader: value rows, peer: gate rows
        }
      } else {
        brow0 = nt * p.BN + (int)rank * b_rows;
      }
      const bool two_b = geglu && CG == 1;
      const bool kmerge = p.kmerge != 0;
      const int step = n_par;
      // (tap, channel chunk) of this thread's first iteration; afterwards advanced incrementally
      // Second K segment (p.cin2 > 0): after the taps over the first operand, `cin2` more channels of a second
      // NHWC tensor (tmA2, no spatial shift) against the weight columns that follow - a ResBlock's 1x1 skip
      // convolution accumulated into its second 3x3 convolution (diffusion.mojo:66-72) instead of a GEMM of its own.
      const int it_first = it_begin + role_parity;
      const int taps = p.taps, cin2 = p.cin2;
      const int cstride = p.cstride, coff = p.coff;
      int tap = it_first / p.chunks_per_tap;
      int kc = (it_first - tap * p.chunks_per_tap) * bk;
      int dy = 0, dx = 0;
      int cin_cur = cin;
      if (tap >= taps) {  // this split starts inside the second segment
        kc = (it_first - taps * p.chunks_per_tap) * bk;
        tap = taps;
        cin_cur = cin2;
      } else if (taps == 9) {
        dy = tap / 3 - 1;
        dx = tap - (tap / 3) * 3 - 1;
      }
      int kb = tap * cin + kc;
      auto advance = [&]() {
        for (int s = 0; s < step; ++s) {
          kc += bk;
          kb += bk;
          if (kc >= cin_cur) {
            kc = 0;
            ++tap;
            kb = tap * cin;
            if (++dx > 1) {
              dx = -1;
              ++dy;
            }
            if (tap >= taps) {
              cin_cur = cin2;
              dx = 0;
              dy = 0;
            }
          }
        }
      };
      int stage = role_parity % num_stages;
      uint32_t phase = (uint32_t)(role_parity / num_stages) & 1u;
      const uint32_t fb_mask = (CG == 2) ? 0xFEFFFFFFu : 0xFFFFFFFFu;  // CG == 2: signal the leader's barrier
      auto load_b = [&](uint32_t sb, uint32_t fbs_, int kb_) {
        if (kmerge) {
          tma_b_4d<CG>(sb, &tmB, fbs_, brow0, batch, kb_ >> 5);
        } else if (!two_b) {
          tma_b_3d<CG>(sb, &tmB, fbs_, kb_, brow0, batch);
          if (natoms == 2) tma_b_3d<CG>(sb + b_atom, &tmB, fbs_, kb_ + 32, brow0, batch);
        } else {
          const uint32_t hb = (uint32_t)half_rows * 128;
          tma_b_3d<CG>(sb, &tmB, fbs_, kb_, brow0, batch);
          tma_b_3d<CG>(sb + hb, &tmB, fbs_, kb_, brow1, batch);
          if (natoms == 2) {
            tma_b_3d<CG>(sb + b_atom, &tmB, fbs_, kb_ + 32, brow0, batch);
            tma_b_3d<CG>(sb + b_atom + hb, &tmB, fbs_, kb_ + 32, brow1, batch);
          }
        }
      };
      if (nb_stages > 0) {
        // ---- separate rings: warp 10 streams the weight boxes through their own (deep) ring, warp 0 the A boxes ----
        if (do_b) {
          const bool stat = p.b_static && !(debug & 4);
          if (!stat) pdl_wait();  // B is an activation: it belongs to the predecessor (static weights need no wait at all)
          int t = 0;
          uint32_t ph = 0;
          for (int it = 0; it < n_iters; ++it) {
            const uint32_t fb = fullb0 + 8u * t;
            if (it >= nb_stages) mbar_wait_a(emptyb0 + 8u * t, ph ^ 1u);
            if (CG == 1 || rank == 0) mbar_expect_tx_a(fb, (uint32_t)CG * b_tx);
            if (!(debug & 4)) load_b(b_ring + (uint32_t)t * b_stage, fb & fb_mask, kb);
            advance();
            if (++t == nb_stages) {
              t = 0;
              ph ^= 1u;
            }
          }
        } else {
          pdl_wait();
          for (int it = 0; it < n_iters; ++it) {
            const uint32_t sa = smem_base + (uint32_t)stage * a_stage;
            const uint32_t fb = full0 + 8u * stage;
            const uint32_t fbs = fb & fb_mask;
            if (it >= num_stages) mbar_wait_a(empty0 + 8u * stage, phase ^ 1u);
            if (CG == 1 || rank == 0) mbar_expect_tx_a(fb, (uint32_t)CG * a_tx);
            if (!(debug & 2)) {
              const CUtensorMap* ta = tap >= taps ? &tmA2 : &tmA;
              const int wx = w0 * cstride + dx + coff, hy = h0 * cstride + dy + coff;
              if (kmerge) {
                tma_a_5d<CG>(sa, ta, fbs, wx, hy, c3, kc >> 5);
              } else {
                tma_a_4d<CG>(sa, ta, fbs, kc, wx, hy, c3);
                if (natoms == 2) tma_a_4d<CG>(sa + a_atom, ta, fbs, kc + 32, wx, hy, c3);
              }
            }
            advance();
            if (++stage == num_stages) {
              stage = 0;
              phase ^= 1u;
            }
          }
          TSD_TRACE(true, 2);
        }
      } else {
      // Weights do not depend on the predecessor kernel: the B halves of the first ring pass are
      // requested BEFORE the programmatic-dependency wait, so their HBM latency overlaps its tail.
      // (With a separate B producer the barrier is armed by the A producer in the main loop; bytes that land before
      // the expect_tx only drive the transaction count negative, the phase cannot complete before the arrival.)
      int pre_end = role_parity;  // iterations < pre_end (of this parity) already have their B tile in flight
      if (p.b_static && !(debug & 4)) {
        pre_end = n_iters < num_stages ? n_iters : num_stages;
        int kc2 = kc, tap2 = tap, kb2 = kb, cin_cur2 = cin_cur;
        for (int it = role_parity; it < pre_end && do_b; it += step) {
          const uint32_t fb2 = full0 + 8u * it;
          if (!GEMM_B_PRODUCER && (CG == 1 || rank == 0)) mbar_expect_tx_a(fb2, tx);
          load_b(smem_base + it * stage_bytes + natoms * a_atom, fb2 & fb_mask, kb2);
          for (int s = 0; s < step; ++s) {
            kc2 += bk;
            kb2 += bk;
            if (kc2 >= cin_cur2) {
              kc2 = 0;
              ++tap2;
              kb2 = tap2 * cin;
              if (tap2 >= taps) cin_cur2 = cin2;
            }
          }
        }
      }
      pdl_wait();  // A (and any aliasing of the arena) belongs to the predecessor: no other global access before this
      for (int it = role_parity; it < n_iters; it += step) {
        const uint32_t sa = smem_base + stage * stage_bytes;
        const uint32_t fb = full0 + 8u * stage, eb = empty0 + 8u * stage;
        const uint32_t fbs = fb & fb_mask;
        if (GEMM_B_PRODUCER) {
          if (it >= num_stages || (do_b && it >= pre_end)) mbar_wait_a(eb, phase ^ 1u);  // the first ring pass finds every slot free
          if (do_a && (CG == 1 || rank == 0)) mbar_expect_tx_a(fb, tx);
          if (do_b && it >= pre_end && !(debug & 4)) load_b(sa + natoms * a_atom, fbs, kb);
        } else if (it >= pre_end) {
          mbar_wait_a(eb, phase ^ 1u);
          if (CG == 1 || rank == 0) mbar_expect_tx_a(fb, tx);
          if (!(debug & 4)) load_b(sa + natoms * a_atom, fbs, kb);
        }
        if (do_a && !(debug & 2)) {
          const CUtensorMap* ta = tap >= taps ? &tmA2 : &tmA;
          const int wx = w0 * cstride + dx + coff, hy = h0 * cstride + dy + coff;
          if (kmerge) {
            tma_a_5d<CG>(sa, ta, fbs, wx, hy, c3, kc >> 5);
          } else {
            tma_a_4d<CG>(sa, ta, fbs, kc, wx, hy, c3);
            if (natoms == 2) tma_a_4d<CG>(sa + a_atom, ta, fbs, kc + 32, wx, hy, c3);
          }
        }
        advance();
        stage += step;
        if (stage >= num_stages) {
          stage -= num_stages;
          phase ^= 1u;
        }
      }
      TSD_TRACE(role_parity == 0 && do_a, 2);
      }  // shared ring
    }
  } else if ((warp == 1 || (GEMM_ROLE_PAIRS > 1 && warp == 11)) && role_parity < n_par) {
    // ===================== MMA issuer (leader CTA only when paired) =====================
    if ((CG == 1 || rank == 0) && elect_one()) {
      const uint32_t idesc = umma_idesc(UMMA_FMT_TF32, GEMM_BM * CG, (uint32_t)p.BN, 0, 0);
      const int chunks = p.chunks_per_tap;
      const int step = n_par;
      // K = 8 MMAs with real data in a tap's last chunk (all 8 when cin is a multiple of 64)
      int nk_last = (p.cin - (chunks - 1) * bk + 7) >> 3;
      if (p.debug & 1) nk_last = 0;
      const int nk_full = (p.debug & 1) ? 0 : bk / 8;
      int chunk = (it_begin + role_parity) % chunks;
      const bool deep = nb_stages > 0;  // separate A / B rings
      const uint64_t adesc0 = umma_smem_desc(smem_base, 16, 1024, UMMA_SWIZZLE_128B);
      const uint64_t bdesc0 = umma_smem_desc(deep ? b_ring : smem_base + natoms * a_atom, 16, 1024, UMMA_SWIZZLE_128B);
      const uint32_t a_step = a_atom >> 4, b_step = b_atom >> 4;
      const uint32_t a_stage_step = (deep ? a_stage : stage_bytes) >> 4, b_stage_step = (deep ? b_stage : stage_bytes) >> 4;
      int tb = 0;
      uint32_t phb = 0;
      const uint32_t tmem_acc = tmem_d + (uint32_t)(role_parity * p.acc_stride);  // this issuer's accumulator tile
      int stage = role_parity % num_stages;
      uint32_t phase = (uint32_t)(role_parity / num_stages) & 1u;
      uint32_t acc = 0;
      for (int it = role_parity; it < n_iters; it += step) {
        const int nk = (chunk == chunks - 1) ? nk_last : nk_full;
        chunk += step;
        while (chunk >= chunks) chunk -= chunks;
        mbar_wait_a(full0 + 8u * stage, phase);
        if (deep) mbar_wait_a(fullb0 + 8u * tb, phb);
        tc_fence_after_sync();
        const uint64_t ad = adesc0 + (uint32_t)stage * a_stage_step, bd = bdesc0 + (uint32_t)(deep ? tb : stage) * b_stage_step;
        if (nk == 8) {
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {
            umma_tf32_cg<CG>(tmem_acc, ad + (kk >> 2) * a_step + 2u * (kk & 3), bd + (kk >> 2) * b_step + 2u * (kk & 3),
                             idesc, acc);
            acc = 1u;
          }
        } else if (nk == 4) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            umma_tf32_cg<CG>(tmem_acc, ad + 2u * kk, bd + 2u * kk, idesc, acc);
            acc = 1u;
          }
        } else {
          for (int kk = 0; kk < nk; ++kk) {
            umma_tf32_cg<CG>(tmem_acc, ad + (kk >> 2) * a_step + 2u * (kk & 3), bd + (kk >> 2) * b_step + 2u * (kk & 3),
                             idesc, acc);
            acc = 1u;
          }
        }
        umma_commit_cg<CG>(empty0 + 8u * stage);  // smem slot reusable (in both CTAs) once these MMAs retire
        if (deep) {
          umma_commit_cg<CG>(emptyb0 + 8u * tb);
          if (++tb == nb_stages) {
            tb = 0;
            phb ^= 1u;
          }
        }
        stage += step;
        if (stage >= num_stages) {
          stage -= num_stages;
          phase ^= 1u;
        }
      }
      umma_commit_cg<CG>(accum_a);  // this issuer's accumulator complete
      TSD_TRACE(role_parity == 0, 3);
    }
  } else if (warp >= 2 && warp < 10) {
    // ===================== epilogue (warps 2..9) =====================
    EpilogueCtx ec;
    ec.ring = smem_dyn + (smem_base - smem_u32(smem_dyn));
    ec.tmem_d = tmem_d;
    ec.accum_bar = accum_a;
    ec.n_iters = n_iters;
    ec.nt = nt; ec.img = img; ec.h0 = h0; ec.w0 = w0; ec.batch = batch; ec.split = split;
    ec.tr = tr;
    ec.red_bar = smem_u32(&red_bar[0]);
    ec.pair_rank = rank;
    ec.cg = CG;
    gemm_epilogue(p, ec);
    tc_fence_before_sync();
  }

  __syncthreads();
#ifdef TSD_LAB_TRACE
  if ((p.debug & 8) && threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0)
    printf("gemm trace: setup %lld producer_done %lld mma_issued %lld accum_ready %lld epilogue_done %lld end %lld (clk since entry)\n",
           tr[1] - tr[0], tr[2] - tr[0], tr[3] - tr[0], tr[4] - tr[0], tr[5] - tr[0], clock64() - tr[0]);
  if ((p.debug & 8) && threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0)
    printf("  epilogue chunk 0 of warp 2: loads issued %lld tmem ready %lld math done %lld staged %lld stored %lld; chunk 1: staged %lld stored %lld\n",
           tr[8] - tr[4], tr[9] - tr[4], tr[10] - tr[4], tr[11] - tr[4], tr[12] - tr[4], tr[13] - tr[4], tr[14] - tr[4]);
#endif

  // neither CTA of a pair may release TMEM / exit while the pair is in flight; no CTA of a split-K cluster may exit while
  // the split-0 CTA still reads its column sums through distributed shared memory
  if (CG == 2 || cluster_split) cluster_sync_all();
  if (warp == 1) {
    tc_fence_after_sync();
    if constexpr (CG == 1) {
      tmem_dealloc(tmem_d, (uint32_t)p.tmem_cols);
    } else {
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_d), "r"((uint32_t)p.tmem_cols)
                   : "memory");
    }
  }
}


// ------------------------------------------------------------------------------------------
// 3x3 / stride 1 / pad 1 convolution with the activation halo kept in shared memory.
//
// The implicit GEMM above re-loads the 128-pixel A tile once per tap: 9 x 16 KiB per 32 channels,
// and the L2 -> SM port (~76 B/clk/SM measured) - not the tensor core - bounds every tile with
// N <= 256.  Here a CTA owns a 16 x 8 pixel box and loads its (18 x 16) x 32-channel halo ONCE per
// channel chunk (36 KiB, TMA zero fill = padding); the nine taps are nine shifted VIEWS of that
// tile: with a row pitch of 16 pixels the 8-pixel rows of the box sit a uniform 2 KiB apart, so a
// tap is just another start address with stride-dimension byte offset 2048 (the hardware swizzles on
// absolute address bits, so TMA's pattern and the shifted UMMA view agree).  A traffic drops 4x; B (weights) streams through its own ring,
// three taps (one kernel row) per stage.  K order: channel chunk major, tap minor.
// ------------------------------------------------------------------------------------------
constexpr int HALO_A_BYTES = 18 * 16 * 128;  // 36 KiB per 32-channel chunk
constexpr int HALO_SA = 2, HALO_SB_MAX = 16, HALO_TB = 3;

template <int CG>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
conv3x3_halo_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const GemmKParams p) {
  extern __shared__ uint8_t smem_dyn[];
  __shared__ __align__(8) uint64_t a_full[HALO_SA], a_empty[HALO_SA];
  __shared__ __align__(8) uint64_t b_full[HALO_SB_MAX], b_empty[HALO_SB_MAX];
  __shared__ __align__(8) uint64_t accum_bar;
  __shared__ uint32_t tmem_slot;
  __shared__ long long tr[16];
  TSD_TRACE(threadIdx.x == 0, 0);

  const int warp = threadIdx.x >> 5;
  const uint32_t rank = (CG == 2) ? cluster_ctaid_x() : 0u;  // position in the CTA pair (the cluster may also span K splits along z)
  const uint32_t smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  const int b_rows = p.BN / CG;
  const uint32_t b_atom = (uint32_t)b_rows * 128;
  const uint32_t b_stage = HALO_TB * b_atom;
  const uint32_t b_ring = smem_base + HALO_SA * HALO_A_BYTES;
  const int SB = p.num_stages;

  const int nt = blockIdx.y;
  int mt = blockIdx.x;
  const int batch = 0;
  const int split = blockIdx.z;
  const int tw = mt % p.tiles_w;
  mt /= p.tiles_w;
  const int th = mt % p.tiles_h;
  const int img = mt / p.tiles_h;
  const int w0 = tw * 8, h0 = th * 16;

  // K range of this split, in 32-channel chunks
  const int c_begin = split * p.iters_per_split;
  int c_end = c_begin + p.iters_per_split;
  if (c_end > p.total_iters) c_end = p.total_iters;
  const int n_chunks = c_end - c_begin;
  const uint32_t af0 = smem_u32(&a_full[0]), ae0 = smem_u32(&a_empty[0]);
  const uint32_t bf0 = smem_u32(&b_full[0]), be0 = smem_u32(&b_empty[0]), accum_a = smem_u32(&accum_bar);

  if (threadIdx.x == 0) {
    for (int s = 0; s < HALO_SA; ++s) {
      mbar_init(&a_full[s], 1);
      mbar_init(&a_empty[s], 1);
    }
    for (int s = 0; s < SB; ++s) {
      mbar_init(&b_full[s], 1);
      mbar_init(&b_empty[s], 1);
    }
    mbar_init(&accum_bar, 1);
    fence_barrier_init();
    prefetch_tensormap(&tmA);
    prefetch_tensormap(&tmB);
  }
  if (warp == 1) {
    if constexpr (CG == 1) {
      tmem_alloc(&tmem_slot, (uint32_t)p.tmem_cols);
      tmem_relinquish();
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_slot)),
                   "r"((uint32_t)p.tmem_cols)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;\n" ::: "memory");
    }
  }
  tc_fence_before_sync();
  if constexpr (CG == 2) cluster_sync_all();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_d = tmem_slot;
  TSD_TRACE(threadIdx.x == 0, 1);
  pdl_launch_dependents();

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      const int cin = p.cin;
      const uint32_t a_tx = (uint32_t)CG * HALO_A_BYTES;
      const uint32_t b_tx = (uint32_t)CG * b_stage;
      const uint32_t mask = (CG == 2) ? 0xFEFFFFFFu : 0xFFFFFFFFu;  // pairs: signal the leader's barrier
      const int brow0 = nt * p.BN + (int)rank * b_rows;
      const int nb = n_chunks * 3;  // B stages of this CTA: (chunk, kernel row)
      auto load_b = [&](int j) {
        const int st = j % SB;
        const int chunk = c_begin + j / 3, r = j - (j / 3) * 3;
        const uint32_t fb = bf0 + 8u * st;
        if (CG == 1 || rank == 0) mbar_expect_tx_a(fb, b_tx);
        const uint32_t dst = b_ring + st * b_stage;
#pragma unroll
        for (int t = 0; t < HALO_TB; ++t)
          tma_b_3d<CG>(dst + t * b_atom, &tmB, fb & mask, (r * 3 + t) * cin + chunk * 32, brow0, 0);
      };
      // weights first: they do not depend on the predecessor kernel
      int jb = 0;
      const int pre = nb < SB ? nb : SB;
      for (; jb < pre; ++jb) load_b(jb);
      pdl_wait();
      uint32_t a_phase = 0;
      int sa = 0;
      for (int ci = 0; ci < n_chunks; ++ci) {
        mbar_wait_a(ae0 + 8u * sa, a_phase ^ 1u);
        const uint32_t fa = af0 + 8u * sa;
        if (CG == 1 || rank == 0) mbar_expect_tx_a(fa, a_tx);
        tma_a_4d<CG>(smem_base + sa * HALO_A_BYTES, &tmA, fa & mask, (c_begin + ci) * 32, w0 - 1, h0 - 1, img);
        if (++sa == HALO_SA) {
          sa = 0;
          a_phase ^= 1u;
        }
        // B stages of this chunk (never wait on a slot whose consumer needs an A tile not yet requested)
        const int want = (ci + 1) * 3 < nb ? (ci + 1) * 3 : nb;
        for (; jb < want; ++jb) {
          if (jb >= SB) mbar_wait_a(be0 + 8u * (jb % SB), (uint32_t)(((jb / SB) & 1) ^ 1));
          load_b(jb);
        }
      }
      TSD_TRACE(true, 2);
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only when paired) =====================
    if ((CG == 1 || rank == 0) && elect_one()) {
      const uint32_t idesc = umma_idesc(UMMA_FMT_TF32, GEMM_BM * CG, (uint32_t)p.BN, 0, 0);
      const int nk_full = 4;
      const int nk_last = (p.cin - (p.total_iters - 1) * 32 + 7) >> 3;  // K = 8 MMAs with data in the last chunk
      // measured (tools/lab/halo_probe.py): the 128 B swizzle is a function of the absolute shared-memory
      // address bits, so a tap view needs no base-offset field; debug bit 4 sets it (lab: gives wrong results)
      const bool use_base_off = (p.debug & 16) != 0;
      const uint64_t adesc0 = umma_smem_desc(smem_base, 16, 2048, UMMA_SWIZZLE_128B);
      const uint64_t bdesc0 = umma_smem_desc(b_ring, 16, 1024, UMMA_SWIZZLE_128B);
      uint32_t acc = 0;
      int sa = 0, j = 0;
      uint32_t a_phase = 0;
      for (int ci = 0; ci < n_chunks; ++ci) {
        const int nk = (c_begin + ci == p.total_iters - 1) ? nk_last : nk_full;
        mbar_wait_a(af0 + 8u * sa, a_phase);
        // descriptors advance by plain adds: (bytes >> 4) in the low word, no carry out of the 14-bit field
        const uint64_t ad_chunk = adesc0 + (uint32_t)((sa * HALO_A_BYTES) >> 4) + (use_base_off ? 0u : 0u);
        for (int r = 0; r < 3; ++r, ++j) {
          const int st = j % SB;
          mbar_wait_a(bf0 + 8u * st, (uint32_t)((j / SB) & 1));
          tc_fence_after_sync();
          const uint64_t bd_stage = bdesc0 + (uint32_t)((st * b_stage) >> 4);
#pragma unroll
          for (int t = 0; t < HALO_TB; ++t) {
            uint64_t ad = ad_chunk + (uint32_t)((r * 16 + t) * 8);  // tap view: +((dy+1)*16 + (dx+1)) rows of 128 B
            if (use_base_off) ad |= (uint64_t)(((smem_base + sa * HALO_A_BYTES + (r * 16 + t) * 128) >> 7) & 7u) << 49;
            const uint64_t bd = bd_stage + (uint32_t)((t * b_atom) >> 4);
            if (nk == 4) {
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) {
                umma_tf32_cg<CG>(tmem_d, ad + 2u * kk, bd + 2u * kk, idesc, acc);
                acc = 1u;
              }
            } else {
              for (int kk = 0; kk < nk; ++kk) {
                umma_tf32_cg<CG>(tmem_d, ad + 2u * kk, bd + 2u * kk, idesc, acc);
                acc = 1u;
              }
            }
          }
          umma_commit_cg<CG>(be0 + 8u * st);
        }
        umma_commit_cg<CG>(ae0 + 8u * sa);
        if (++sa == HALO_SA) {
          sa = 0;
          a_phase ^= 1u;
        }
      }
      umma_commit_cg<CG>(accum_a);
      TSD_TRACE(true, 3);
    }
  } else if (warp >= 2 && warp < 10) {
    EpilogueCtx ec;
    ec.ring = smem_dyn + (smem_base - smem_u32(smem_dyn));
    ec.tmem_d = tmem_d;
    ec.accum_bar = accum_a;
    ec.n_iters = 1;  // one accumulator tile
    ec.nt = nt; ec.img = img; ec.h0 = h0; ec.w0 = w0; ec.batch = batch; ec.split = split;
    ec.tr = tr;
    gemm_epilogue(p, ec);
    tc_fence_before_sync();
  }

  __syncthreads();
#ifdef TSD_LAB_TRACE
  if ((p.debug & 8) && threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0)
    printf("halo trace: setup %lld producer_done %lld mma_issued %lld accum_ready %lld epilogue_done %lld end %lld (clk since entry)\n",
           tr[1] - tr[0], tr[2] - tr[0], tr[3] - tr[0], tr[4] - tr[0], tr[5] - tr[0], clock64() - tr[0]);
#endif
  if constexpr (CG == 2) cluster_sync_all();
  if (warp == 1) {
    tc_fence_after_sync();
    if constexpr (CG == 1) {
      tmem_dealloc(tmem_d, (uint32_t)p.tmem_cols);
    } else {
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_d), "r"((uint32_t)p.tmem_cols)
                   : "memory");
    }
  }
}

// Sums split-K partials and applies the (bias, residual, rounding) epilogue.
__global__ void splitk_reduce_kernel(const SplitKReduceParams p) {
  pdl_wait();
  pdl_launch_dependents();
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

static size_t gemm_stage_bytes(int BN, int cg, int bk) { return (size_t)(bk / 32) * ((size_t)GEMM_BM * 128 + (size_t)(BN / cg) * 128); }

size_t gemm_smem_bytes(int BN, int num_stages, int cg, int bk, int b_stages, int dsm_tile) {
  size_t ring = (size_t)num_stages * gemm_stage_bytes(BN, cg, bk);
  if (b_stages > 0)  // separate rings: num_stages A stages + b_stages B stages
    ring = (size_t)(bk / 32) * ((size_t)num_stages * GEMM_BM * 128 + (size_t)b_stages * (BN / cg) * 128);
  const size_t staging = 8 * 32 * 36 * 4 + 4 * 256 * 8;  // epilogue staging tiles + column statistics live in the (then idle) ring
  if (ring < staging) ring = staging;
  const size_t tile = (size_t)GEMM_DSM_TILE_OFFSET + (size_t)GEMM_BM * (((BN + 31) & ~31) + 4) * 4;  // cluster split-K: raw partial tile
  if (dsm_tile && ring < tile) ring = tile;
  return ring + 1024;
}

// Ring depth and K step.  The two role pairs (even / odd K steps) must own disjoint stage sets for
// their completions to stay ordered, so the depth is even; K steps of 64 are preferred (fewer
// barrier round trips) unless they leave fewer than 4 stages in shared memory.
void gemm_pick_ring(int BN, int cg, int* bk, int* stages) {
  const size_t budget = 224 * 1024 - 1024;  // 227 KiB opt-in minus static shared memory and alignment
  for (int k : {64, 32}) {
    int s = (int)(budget / gemm_stage_bytes(BN, cg, k));
    if (s > GEMM_MAX_STAGES) s = GEMM_MAX_STAGES;
    if (GEMM_ROLE_PAIRS > 1) s &= ~1;
    if (s >= (GEMM_ROLE_PAIRS > 1 ? 4 : 2) || k == 32) {
      *bk = k;
      *stages = s < 2 ? 2 : s;
      return;
    }
  }
}

template <int CG>
static cudaError_t launch_gemm_cg(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmA2,
                                  const GemmKParams& p, dim3 grid, size_t smem_bytes, cudaStream_t stream) {
  // static + dynamic shared memory must fit the 227 KiB opt-in limit together
  int max_dyn = 0;
  {
    cudaError_t e = optin_dyn_smem(reinterpret_cast<const void*>(gemm_tf32_kernel<CG>), -1, &max_dyn);
    if (e != cudaSuccess) return e;
  }
  if ((long long)smem_bytes > max_dyn) return cudaErrorInvalidConfiguration;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(GEMM_THREADS, 1, 1);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = p.fixup >= 2 ? p.splits : 1;  // cluster split-K: the K splits of a tile are co-scheduled
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  if (p.fixup >= 2 && CG * p.splits > 8) {  // beyond the portable cluster size: opt in once per (kernel, device)
    static std::mutex mu;
    static std::map<int, bool> done;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(mu);
    if (!done[dev]) {
      cudaError_t e = cudaFuncSetAttribute(gemm_tf32_kernel<CG>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
      if (e != cudaSuccess) return e;
      done[dev] = true;
    }
  }
  return cudaLaunchKernelEx(&cfg, gemm_tf32_kernel<CG>, tmA, tmB, tmA2, p);
}

// halo convolution: B ring depth that fits next to the two halo tiles (0: configuration unsupported)
int halo_pick_sb(int BN, int cg) {
  const size_t budget = 224 * 1024 - 1024;
  const size_t b_stage = (size_t)HALO_TB * (BN / cg) * 128;
  if ((size_t)HALO_SA * HALO_A_BYTES + 2 * b_stage > budget) return 0;
  int sb = (int)((budget - (size_t)HALO_SA * HALO_A_BYTES) / b_stage);
  if (sb > HALO_SB_MAX) sb = HALO_SB_MAX;
  return sb;
}
size_t halo_smem_bytes(int BN, int cg, int sb) {
  size_t ring = (size_t)HALO_SA * HALO_A_BYTES + (size_t)sb * HALO_TB * (BN / cg) * 128;
  const size_t staging = 8 * 32 * 36 * 4 + 4 * 256 * 8;
  if (ring < staging) ring = staging;
  return ring + 1024;
}

template <int CG>
static cudaError_t launch_halo_cg(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmKParams& p, dim3 grid,
                                  size_t smem_bytes, cudaStream_t stream) {
  int max_dyn = 0;
  {
    cudaError_t e = optin_dyn_smem(reinterpret_cast<const void*>(conv3x3_halo_kernel<CG>), -1, &max_dyn);
    if (e != cudaSuccess) return e;
  }
  if ((long long)smem_bytes > max_dyn) return cudaErrorInvalidConfiguration;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(GEMM_THREADS, 1, 1);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  return cudaLaunchKernelEx(&cfg, conv3x3_halo_kernel<CG>, tmA, tmB, p);
}

cudaError_t launch_conv_halo(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmKParams& p, dim3 grid,
                             size_t smem_bytes, cudaStream_t stream) {
  if (p.cg == 2) return launch_halo_cg<2>(tmA, tmB, p, grid, smem_bytes, stream);
  return launch_halo_cg<1>(tmA, tmB, p, grid, smem_bytes, stream);
}

cudaError_t launch_gemm_tf32(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmKParams& p,
                             dim3 grid, size_t smem_bytes, cudaStream_t stream, const CUtensorMap* tmA2) {
  const CUtensorMap& a2 = tmA2 ? *tmA2 : tmA;
  if (p.cg == 2) return launch_gemm_cg<2>(tmA, tmB, a2, p, grid, smem_bytes, stream);
  return launch_gemm_cg<1>(tmA, tmB, a2, p, grid, smem_bytes, stream);
}

// Slab variant with producer-side norm statistics (norm_stats.cuh): a block owns `slab_rows` rows x all
// columns; a thread owns up to three fixed column quads, so the per-column sums of the values it
// stores accumulate in registers with no atomics.
constexpr int RK_THREADS = 256, RK_MAXQ = 3;
__global__ void __launch_bounds__(RK_THREADS)
splitk_reduce_stats_kernel(const SplitKReduceParams p) {
  extern __shared__ float2 rk_sm[];  // [ppl][C] column sums, folded into [C]
  pdl_wait();
  pdl_launch_dependents();
  const int n4 = p.n_valid >> 2, C = p.n_valid;
  const int TU = n4 < RK_THREADS ? n4 : RK_THREADS;
  const int ppl = n4 < RK_THREADS ? RK_THREADS / n4 : 1;
  const int u = threadIdx.x % TU, pl = threadIdx.x / TU;
  const int row0 = blockIdx.x * p.slab_rows;
  float4 ss[RK_MAXQ], qq[RK_MAXQ];
#pragma unroll
  for (int i = 0; i < RK_MAXQ; ++i) ss[i] = qq[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (pl < ppl) {
    for (int r = pl; r < p.slab_rows; r += ppl) {
      const int row = row0 + r;
      if (row >= p.m) break;
      const float* brow = p.bias ? p.bias + (long long)(row / p.rows_per_img) * p.bias_img_stride : nullptr;
#pragma unroll
      for (int i = 0; i < RK_MAXQ; ++i) {
        const int qd = u + i * TU;
        if (qd < n4) {
          const float* src = p.partial + (long long)row * p.n_pad + qd * 4;
          float4 acc = *reinterpret_cast<const float4*>(src);
          for (int sp = 1; sp < p.splits; ++sp) {
            const float4 t = *reinterpret_cast<const float4*>(src + sp * p.split_stride);
            acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
          }
          if (brow) {
            const float4 b = *reinterpret_cast<const float4*>(brow + qd * 4);
            acc.x += b.x; acc.y += b.y; acc.z += b.z; acc.w += b.w;
          }
          if (p.residual) {
            const float4 t = *reinterpret_cast<const float4*>(p.residual + (long long)row * p.ldr + qd * 4);
            acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
          }
          if (p.round_tf32) {
            acc.x = round_tf32(acc.x); acc.y = round_tf32(acc.y); acc.z = round_tf32(acc.z); acc.w = round_tf32(acc.w);
          }
          *reinterpret_cast<float4*>(p.D + (long long)row * p.ldd + qd * 4) = acc;
          ss[i].x += acc.x; ss[i].y += acc.y; ss[i].z += acc.z; ss[i].w += acc.w;
          qq[i].x = fmaf(acc.x, acc.x, qq[i].x); qq[i].y = fmaf(acc.y, acc.y, qq[i].y);
          qq[i].z = fmaf(acc.z, acc.z, qq[i].z); qq[i].w = fmaf(acc.w, acc.w, qq[i].w);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < RK_MAXQ; ++i) {
      const int qd = u + i * TU;
      if (qd < n4) {
        float2* d = rk_sm + (long long)pl * C + qd * 4;
        d[0] = make_float2(ss[i].x, qq[i].x); d[1] = make_float2(ss[i].y, qq[i].y);
        d[2] = make_float2(ss[i].z, qq[i].z); d[3] = make_float2(ss[i].w, qq[i].w);
      }
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += RK_THREADS) {  // fold the row lanes in a fixed order
    float2 a = rk_sm[c];
    for (int l = 1; l < ppl; ++l) {
      const float2 t = rk_sm[(long long)l * C + c];
      a.x += t.x;
      a.y += t.y;
    }
    rk_sm[c] = a;
  }
  __syncthreads();
  {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float2* dst = p.ns.partial + blockIdx.x;  // [group][slab]
    for (int g = warp; g < p.ns.G; g += RK_THREADS / 32) {
      float s = 0.f, qv = 0.f;
      for (int cc = g * p.ns.cpg + lane; cc < (g + 1) * p.ns.cpg; cc += 32) {
        const float2 t = rk_sm[cc];
        s += t.x;
        qv += t.y;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        qv += __shfl_xor_sync(0xffffffffu, qv, o);
      }
      if (lane == 0) dst[(long long)g * p.ns.slabs_total] = make_float2(s, qv);
    }
  }
}

cudaError_t launch_splitk_reduce(const SplitKReduceParams& p, cudaStream_t stream) {
  if (p.ns.partial != nullptr) {
    const int n4 = p.n_valid >> 2;
    const int ppl = n4 < RK_THREADS ? RK_THREADS / n4 : 1;
    const size_t smem = (size_t)ppl * p.n_valid * sizeof(float2);
    {
      cudaError_t e = optin_dyn_smem(reinterpret_cast<const void*>(splitk_reduce_stats_kernel), 96 * 1024, nullptr);
      if (e != cudaSuccess) return e;
    }
    const int blocks = (p.m + p.slab_rows - 1) / p.slab_rows;
    return launch_pdl(splitk_reduce_stats_kernel, dim3(blocks), dim3(RK_THREADS), smem, stream, p);
  }
  const long long total = (long long)p.m * (p.n_pad >> 2);
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  return launch_pdl(splitk_reduce_kernel, dim3(blocks), dim3(256), 0, stream, p);
}

}  // namespace tsd
