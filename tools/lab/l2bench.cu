// Lab micro-benchmark (not part of the product): how many bytes per clock the SMs can pull out of L2 with bulk async
// copies, and what sharing does to it.  The GEMM / implicit-GEMM convolution kernels are bound by exactly this number
// (profiles/r02_ncu_decoder_conv256: 43 B/clk/SM delivered = the full-chip L2 cap), so the question is whether operand
// tiles that several CTAs need should be (a) loaded by each CTA on its own, or (b) loaded once and multicast.
//   mode 0  unicast, every CTA streams its OWN region                     (the cap)
//   mode 1  unicast, the CTAs of a group of `csz` stream the SAME region  (what neighbouring tiles do today)
//   mode 2  multicast: each CTA of a cluster loads 1/csz of every chunk and multicasts it to the whole cluster
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l2bench l2bench.cu && ./l2bench
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok)
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok)
                 : "r"(bar), "r"(parity)
                 : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void bulk_load_mc(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;\n" ::"r"(dst),
      "l"(src), "r"(bytes), "r"(bar), "h"(mask)
      : "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r));
  return r;
}

constexpr int MAXSTAGES = 32;

__global__ void __launch_bounds__(128) stream_kernel(const char* src, long long region, int iters, int mode, int csz,
                                                    unsigned long long* out, int STAGES, int CHUNK, int T, int R) {
  extern __shared__ __align__(128) char buf[];  // STAGES x CHUNK
  __shared__ __align__(8) unsigned long long full[MAXSTAGES];
  const uint32_t rank = cluster_rank();
  if (threadIdx.x == 0)
    for (int s = 0; s < STAGES; ++s) mbar_init(smem_u32(&full[s]), 1);
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;\n" ::: "memory");
  // T issuing threads (lane 0 of warps 0..T-1); thread w owns stages w, w+T, ...; every chunk goes out as R requests
  if ((threadIdx.x & 31) == 0 && (int)(threadIdx.x >> 5) < T) {
    const int w = threadIdx.x >> 5;
    const int group = mode == 0 ? blockIdx.x : blockIdx.x / csz;
    const char* base = src + (long long)group * region;
    const uint16_t mask = (uint16_t)((1u << csz) - 1u);
    const uint32_t slice = CHUNK / csz;
    const uint32_t piece = (mode == 2 ? slice : CHUNK) / R;
    const long long t0 = clock64();
    int s = w;
    uint32_t par = 1;
    long long off = (long long)w * CHUNK;
    const int my_iters = iters / T, my_stages = STAGES / T;
    for (int i = 0; i < my_iters + my_stages; ++i) {  // no divisions in this loop: a lone thread pays ~100+ clk for each
      const uint32_t bar = smem_u32(&full[s]);
      if (i >= my_stages) mbar_wait(bar, par);
      if (i < my_iters) {
        mbar_expect_tx(bar, CHUNK);
        const uint32_t dst = smem_u32(buf + s * CHUNK) + (mode == 2 ? rank * slice : 0);
        const char* g = base + off + (mode == 2 ? rank * slice : 0);
        for (int r = 0; r < R; ++r) {
          if (mode == 2)
            bulk_load_mc(dst + r * piece, g + r * piece, piece, bar, mask);
          else
            bulk_load(dst + r * piece, g + r * piece, piece, bar);
        }
        off += (long long)T * CHUNK;
        if (off >= region) off -= region;
      }
      s += T;
      if (s >= STAGES) {
        s = w;
        par ^= 1u;
      }
    }
    const long long dt = clock64() - t0;
    atomicMax(out, (unsigned long long)dt);
  }
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}

int main(int argc, char** argv) {
  const int max_mc = argc > 1 ? atoi(argv[1]) : 2;  // largest multicast cluster to try
  const long long region = 192 * 1024;  // per group; 296 groups = 57 MB: L2-resident
  char* src;
  unsigned long long* out;
  cudaMalloc(&src, 296 * region);
  cudaMemset(src, 1, 296 * region);
  cudaMalloc(&out, 8);
  cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  const long long per_cta = 32ll << 20;  // bytes delivered per CTA
  const char* names[] = {"unicast, own region", "unicast, region shared by the group", "multicast inside the cluster"};
  struct Cfg { int stages, chunk, per_sm, T, R; };
  const Cfg cfgs[] = {{6, 16384, 1, 1, 1}, {6, 32768, 1, 1, 2}, {6, 32768, 1, 1, 1}, {6, 32768, 1, 1, 4}, {6, 16384, 1, 2, 1}, {6, 16384, 1, 3, 1},
                      {12, 8192, 1, 4, 1}, {12, 16384, 1, 4, 1}, {6, 32768, 1, 2, 1}, {6, 32768, 1, 2, 2}};
  for (const Cfg& cf : cfgs)
  for (int mode = 0; mode < 3; ++mode)
    for (int csz : {1, -32, 4, 2}) {
      const int STAGES = cf.stages, CHUNK = cf.chunk;
      const int iters = (int)(per_cta / CHUNK);
      int grid_override = 0;
      if (csz < 0) {  // mode 0 only: fewer SMs pulling (the per-SM port limit)
        grid_override = -csz;
        csz = 1;
        if (mode != 0) continue;
      }
      if (mode == 0 && csz > 1) continue;
      if (mode != 0 && csz == 1) continue;
      if (mode == 2 && csz > max_mc) continue;
      const int grid = (grid_override ? grid_override : 148 / csz * csz) * cf.per_sm;
      cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, STAGES * CHUNK);
      unsigned long long clk = 0;
      for (int rep = 0; rep < 2; ++rep) {
        cudaMemset(out, 0, 8);
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(grid);
        cfg.blockDim = dim3(128);
        cfg.dynamicSmemBytes = STAGES * CHUNK;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = mode == 2 ? csz : 1;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        cudaError_t e = cudaLaunchKernelEx(&cfg, stream_kernel, (const char*)src, region, iters, mode, csz, out, STAGES, CHUNK, cf.T, cf.R);
        if (e == cudaSuccess) e = cudaMemcpy(&clk, out, 8, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) {
          printf("mode %d csz %d: %s\n", mode, csz, cudaGetErrorString(e));
          cudaGetLastError();
          clk = 0;
          break;
        }
      }
      if (!clk) continue;
      const double per_sm = (double)iters * CHUNK / (double)clk;
      const double unique = mode == 0 ? 1.0 : 1.0 / csz;  // fraction of the delivered bytes that are distinct L2 lines
      printf("%2d x %5d B, %d CTA/SM, %d thr, %d req | %-38s csz %2d grid %3d: %7.1f B/clk/CTA delivered, %8.0f B/clk chip delivered, %8.0f B/clk distinct\n",
             STAGES, CHUNK, cf.per_sm, cf.T, cf.R, names[mode], csz, grid, per_sm, per_sm * grid, per_sm * grid * unique);
      fflush(stdout);
    }
  return 0;
}
