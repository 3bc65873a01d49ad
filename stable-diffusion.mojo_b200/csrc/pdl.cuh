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

#include <map>
#include <mutex>
#include <tuple>
#include <utility>

namespace tsd {

// cudaFuncSetAttribute applies to the CURRENT device only and tsd_init() may create contexts on several devices in
// one process, so the opt-in dynamic shared-memory limit is tracked per (kernel, device).  `request` < 0 asks for
// everything the device allows next to the kernel's static shared memory; the granted limit is returned in *granted.
inline cudaError_t optin_dyn_smem(const void* func, int request, int* granted) {
  static std::mutex mu;
  static std::map<std::pair<const void*, int>, int> done;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  std::lock_guard<std::mutex> lock(mu);
  auto it = done.find({func, dev});
  if (it == done.end()) {
    int lim = request;
    if (request < 0) {
      cudaFuncAttributes fa;
      e = cudaFuncGetAttributes(&fa, func);
      if (e != cudaSuccess) return e;
      int optin = 0;
      e = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
      if (e != cudaSuccess) return e;
      lim = optin - (int)fa.sharedSizeBytes;
    }
    e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, lim);
    if (e != cudaSuccess) return e;
    it = done.emplace(std::make_pair(func, dev), lim).first;
  }
  if (granted) *granted = it->second;
  return cudaSuccess;
}

__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;\n" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory"); }

// process-wide switch (tsd_set_option "pdl"); default on
inline int& pdl_enabled() {
  static int v = 1;
  return v;
}

// Kernels that synchronise their whole grid through a spin barrier need every block resident at once.  They are
// launched as ordinary (PDL) kernels, so the launcher proves co-residency from the occupancy the runtime reports for
// (kernel, block size, dynamic shared memory) on the current device instead of assuming it; a launch that does not
// fit is refused (cudaErrorCooperativeLaunchTooLarge) and the caller takes its barrier-free path.  The barriers
// themselves are bounded spins that trap, so a second context running such a kernel concurrently on the same device
// (which can break co-residency) surfaces as a failed kernel, never as a hung GPU.  Contexts that share a device
// should not run norm kernels concurrently.
inline cudaError_t grid_coresident(const void* func, int threads, size_t smem, long long grid_blocks) {
  static std::mutex mu;
  static std::map<std::tuple<const void*, int, int, size_t>, long long> cap;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  std::lock_guard<std::mutex> lock(mu);
  const auto key = std::make_tuple(func, dev, threads, smem);
  auto it = cap.find(key);
  if (it == cap.end()) {
    int per_sm = 0, sms = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, func, threads, smem);
    if (e != cudaSuccess) return e;
    e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return e;
    it = cap.emplace(key, (long long)per_sm * sms).first;
  }
  return grid_blocks <= it->second ? cudaSuccess : cudaErrorCooperativeLaunchTooLarge;
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
