// lab: which cluster launch configurations does the driver accept?
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(320, 1) k(int* out) {
  extern __shared__ unsigned char sm[];
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r));
  if (threadIdx.x == 0) { atomicAdd(out, 1 + (int)r * 0); }
}
static void tryl(dim3 grid, dim3 cl, size_t smem, int* d) {
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 64);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = dim3(320); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute a[1];
  a[0].id = cudaLaunchAttributeClusterDimension;
  a[0].val.clusterDim.x = cl.x; a[0].val.clusterDim.y = cl.y; a[0].val.clusterDim.z = cl.z;
  cfg.attrs = a; cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, k, d);
  cudaError_t e2 = cudaDeviceSynchronize();
  int nc = -1;
  cudaOccupancyMaxActiveClusters(&nc, k, &cfg);
  printf("grid (%u,%u,%u) cluster (%u,%u,%u) smem %zu: launch=%s sync=%s maxActiveClusters=%d\n", grid.x, grid.y, grid.z,
         cl.x, cl.y, cl.z, smem, cudaGetErrorName(e), cudaGetErrorName(e2), nc);
  cudaGetLastError();
}
int main() {
  int* d; cudaMalloc(&d, 4); cudaMemset(d, 0, 4);
  for (size_t smem : {(size_t)1024, (size_t)100 * 1024, (size_t)160 * 1024, (size_t)200 * 1024, (size_t)222 * 1024}) {
    tryl(dim3(2, 32, 1), dim3(1, 2, 1), smem, d);
    tryl(dim3(32, 2, 1), dim3(2, 1, 1), smem, d);
    tryl(dim3(2, 32, 2), dim3(1, 2, 1), smem, d);
  }
  return 0;
}
