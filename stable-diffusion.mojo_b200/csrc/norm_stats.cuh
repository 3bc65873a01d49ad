// Producer-side GroupNorm / LayerNorm statistics.
//
// A kernel that writes an activation [rows][C] (GEMM epilogue, split-K reduction) also writes, per
// tile it owns (slab of rows x n-tile of BN columns), the partial sums (sum x, sum x^2) of the
// values it stored, folded to the GROUPS that overlap its column range:
//   partial[((nt * lg) + (g - first_group(nt))) * slabs_total + slab]      (float2, fp32 sums)
// (slab innermost: the entries of one group are contiguous runs, so the consumer's fold is a coalesced read - with the
// slab-major layout of round 1 every lane of a fold hit a different 128 B line, and the few hundred blocks of a
// normalise pass folding the same table serialised on those lines in L2: 8 000 - 20 000 cycles per block, measured)
// The consumer (norm_apply_partial_kernel) folds the few hundred partials of its image in a fixed
// order at the top of every block - redundantly, in parallel, from L2 - and normalises:
//   (x - mean) / (std + eps), biased std       (reference helpers/utils.mojo:1380, 1868-1870)
// so no statistics pass over the activation, no grid barrier and no atomics exist on this path;
// the result is bit-reproducible.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace tsd {

// 1 / (std + eps) as the reference computes it (helpers/utils.mojo:1868-1870), or - eps < 0, the
// "norm_eps_mode" option used with real checkpoints - the usual 1 / sqrt(var + |eps|)
__host__ __device__ __forceinline__ float norm_rstd(float var, float eps) {
  return eps >= 0.f ? 1.0f / (sqrtf(var) + eps) : 1.0f / sqrtf(var - eps);
}
__host__ __device__ __forceinline__ double norm_rstd(double var, double eps) {
  return eps >= 0.0 ? 1.0 / (sqrt(var) + eps) : 1.0 / sqrt(var - eps);
}

struct NormStatsReq {
  float2* partial = nullptr;     // (nullptr: no statistics available)
  int lg = 0;                    // entries per tile = max groups overlapping BN columns
  int BN = 0, n_tiles = 0;
  int C = 0, G = 0, cpg = 0;
  int slabs_per_img = 0, imgs = 0;
  int slabs_total = 0;           // imgs * slabs_per_img: the stride between (n-tile, group) rows of the table
  // n / BN and n / cpg as umulhi(n, magic), magic = 2^32 / d + 1 (exact for n, d < 2^16): integer division by a
  // run-time value costs ~100 cycles on this part and the consumers' folds are latency chains
  unsigned bn_magic = 0, cpg_magic = 0;  // 0 = divisor 1 (2^32 does not fit): norm_fastdiv returns n
  void set_magics() {
    bn_magic = BN > 1 ? 0xFFFFFFFFu / (unsigned)BN + 1u : 0u;
    cpg_magic = cpg > 1 ? 0xFFFFFFFFu / (unsigned)cpg + 1u : 0u;
  }
  float inv_count = 0.f;         // 1 / elements per (image, group)
  float eps = 0.f;
};

// All `nthreads` (multiple of 32) threads of a block: st[g] = (mean, 1/(std+eps)) of image `img`.
// One warp per group; lanes stride over the (slab, n-tile) entries of the group.
__device__ __forceinline__ int norm_fastdiv(int n, unsigned magic) {
  return magic ? (int)__umulhi((unsigned)n, magic) : n;
}

struct NormFoldNoHook {
  __device__ __forceinline__ void operator()() const {}
};
// `after_issue` runs once, right after the first pass's loads of the partial table have been issued and before their
// values are used: a caller puts its own independent loads there so that they queue BEHIND the (tiny) table reads.
template <class Hook = NormFoldNoHook>
__device__ __forceinline__ void norm_stats_fold(const NormStatsReq& r, int img, int tid, int nthreads, float2* st,
                                                long long* tr = nullptr, Hook after_issue = Hook(), int g_begin = 0,
                                                int g_end = -1) {
  // groups [g_begin, g_end) only (default: all); st is indexed by g - g_begin
  if (g_end < 0) g_end = r.G;
  const int lane = tid & 31, warp = tid >> 5, nwarps = nthreads >> 5;
  const float2* base = r.partial + (long long)img * r.slabs_per_img;
  const unsigned bn_magic = r.bn_magic, cpg_magic = r.cpg_magic;
  bool hooked = false;
  for (int g0 = g_begin + warp; g0 < g_end; g0 += 4 * nwarps) {
    // Four groups per pass, their entries walked in lock step: a warp issues in order and stalls at the first use of a
    // loaded value, so one loop nest per group costs one L2 round trip per group and n-tile (measured: 7 000+ cycles
    // for 32 groups); here up to 16 independent loads are in flight before the first add.
    float fs[4], fq[4];
    int ent[4], row0[4], nt0s[4], emax = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      fs[k] = fq[k] = 0.f;
      ent[k] = 0;
      row0[k] = 0;
      nt0s[k] = 0;
      const int g = g0 + k * nwarps;
      if (g >= g_end) continue;
      const int nt0 = norm_fastdiv(g * r.cpg, bn_magic);              // n-tiles overlapping the group
      const int nt1 = norm_fastdiv((g + 1) * r.cpg - 1, bn_magic);
      ent[k] = (nt1 - nt0 + 1) * r.slabs_per_img;
      row0[k] = nt0 * r.lg + (g - norm_fastdiv(nt0 * r.BN, cpg_magic));  // a later tile of the group starts inside it: its entry is row nt * lg
      nt0s[k] = nt0;
      emax = ent[k] > emax ? ent[k] : emax;
    }
    if (tr) tr[0] = clock64();
    for (int e0 = lane; e0 < emax; e0 += 128) {
      float2 v[4][4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int e = e0 + 32 * j;
          v[k][j] = make_float2(0.f, 0.f);
          if (e < ent[k]) {
            int t = 0, sl = e, row = row0[k];
            if (e >= r.slabs_per_img) {  // groups wider than an n-tile only (LayerNorm: one group over every tile)
              t = e / r.slabs_per_img;
              sl = e - t * r.slabs_per_img;
              row = (nt0s[k] + t) * r.lg;
            }
            v[k][j] = __ldcg(base + (long long)row * r.slabs_total + sl);
          }
        }
      }
      if (!hooked) {
        after_issue();
        hooked = true;
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        fs[k] += (v[k][0].x + v[k][1].x) + (v[k][2].x + v[k][3].x);
        fq[k] += (v[k][0].y + v[k][1].y) + (v[k][2].y + v[k][3].y);
      }
    }
    if (tr) tr[1] = clock64();
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {  // fp32 trees (fp64 issue is slow on this part); the subtraction below is fp64
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        fs[k] += __shfl_xor_sync(0xffffffffu, fs[k], o);
        fq[k] += __shfl_xor_sync(0xffffffffu, fq[k], o);
      }
    }
    // every lane holds the four totals: lane k finishes group k
    float ms = fs[0], mq = fq[0];
    if (lane == 1) { ms = fs[1]; mq = fq[1]; }
    if (lane == 2) { ms = fs[2]; mq = fq[2]; }
    if (lane == 3) { ms = fs[3]; mq = fq[3]; }
    if (tr) tr[2] = clock64();
    const int g = g0 + lane * nwarps;
    if (lane < 4 && g < g_end) {
      const double mean = (double)ms * (double)r.inv_count;
      double var = (double)mq * (double)r.inv_count - mean * mean;
      if (var < 0.0) var = 0.0;
      st[g - g_begin] = make_float2((float)mean, norm_rstd((float)var, r.eps));
    }
  }
  if (!hooked) after_issue();  // threads without a table entry of their own
}

}  // namespace tsd
