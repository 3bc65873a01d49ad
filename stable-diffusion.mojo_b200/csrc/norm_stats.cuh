// Producer-side GroupNorm / LayerNorm statistics.
//
// A kernel that writes an activation [rows][C] (GEMM epilogue, split-K reduction) also writes, per
// tile it owns (slab of rows x n-tile of BN columns), the partial sums (sum x, sum x^2) of the
// values it stored, folded to the GROUPS that overlap its column range:
//   partial[(slab * n_tiles + nt) * lg + (g - first_group(nt))]      (float2, fp32 sums)
// The consumer (norm_apply_partial_kernel) folds the few hundred partials of its image in a fixed
// order at the top of every block - redundantly, in parallel, from L2 - and normalises:
//   (x - mean) / (std + eps), biased std       (reference helpers/utils.mojo:1380, 1868-1870)
// so no statistics pass over the activation, no grid barrier and no atomics exist on this path;
// the result is bit-reproducible.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace tsd {

struct NormStatsReq {
  float2* partial = nullptr;     // (nullptr: no statistics available)
  int lg = 0;                    // entries per tile = max groups overlapping BN columns
  int BN = 0, n_tiles = 0;
  int C = 0, G = 0, cpg = 0;
  int slabs_per_img = 0, imgs = 0;
  float inv_count = 0.f;         // 1 / elements per (image, group)
  float eps = 0.f;
};

// All `nthreads` (multiple of 32) threads of a block: st[g] = (mean, 1/(std+eps)) of image `img`.
// One warp per group; lanes stride over the (slab, n-tile) entries of the group.
__device__ __forceinline__ void norm_stats_fold(const NormStatsReq& r, int img, int tid, int nthreads, float2* st) {
  const int lane = tid & 31, warp = tid >> 5, nwarps = nthreads >> 5;
  const float2* base = r.partial + (long long)img * r.slabs_per_img * r.n_tiles * r.lg;
  for (int g0 = warp; g0 < r.G; g0 += 4 * nwarps) {  // 4 groups per pass so that their L2 loads overlap
    float fs[4], fq[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      fs[k] = fq[k] = 0.f;
      const int g = g0 + k * nwarps;
      if (g >= r.G) continue;
      const int nt0 = (g * r.cpg) / r.BN, nt1 = ((g + 1) * r.cpg - 1) / r.BN;  // n-tiles overlapping the group
      const int span = nt1 - nt0 + 1;
      const int entries = r.slabs_per_img * span;
      for (int k0 = 0; k0 < entries; k0 += 128) {
        float2 v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int e = k0 + j * 32 + lane;
          v[j] = make_float2(0.f, 0.f);
          if (e < entries) {
            const int sl = e / span, nt = nt0 + (e - sl * span);
            v[j] = __ldcg(base + ((long long)sl * r.n_tiles + nt) * r.lg + (g - (nt * r.BN) / r.cpg));
          }
        }
        fs[k] += (v[0].x + v[1].x) + (v[2].x + v[3].x);
        fq[k] += (v[0].y + v[1].y) + (v[2].y + v[3].y);
      }
    }
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
    const int g = g0 + lane * nwarps;
    if (lane < 4 && g < r.G) {
      const double mean = (double)ms * (double)r.inv_count;
      double var = (double)mq * (double)r.inv_count - mean * mean;
      if (var < 0.0) var = 0.0;
      st[g] = make_float2((float)mean, 1.0f / (sqrtf((float)var) + r.eps));
    }
  }
}

}  // namespace tsd
