// Programmatic dependent launch (PDL): a kernel launched with the attribute may start while its
// stream predecessor is still running; its prologue (barrier init, TMEM allocation, descriptor
// prefetch, weight loads) overlaps the predecessor's tail.  Contract used throughout this library:
//   * every PDL-launched kernel calls pdl_wait() before its first access (read OR write - the
//     workspace arena reuses addresses) to memory its predecessors may touch; pdl_wait() returns
//     once all prerequisite grids have completed and flushed, so no other ordering is needed;
//   * pdl_launch_dependents() is a hint placed after the resource-acquiring prologue.
// Works under stream capture (the edge becomes a programmatic graph dependency).
#pragma once
#include <cuda_runtime.h>

#include <utility>

namespace tsd {

__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;\n" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory"); }

// process-wide switch (tsd_set_option "pdl"); default on
inline int& pdl_enabled() {
  static int v = 1;
  return v;
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

}  // namespace tsd
