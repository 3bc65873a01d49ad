// Shared between the C-ABI translation units.
#pragma once
#include <mutex>

#include "runtime.h"

struct tsd_ctx {
  tsd::Ctx* c = nullptr;
  std::mutex mu;
  int use_graph = 1;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;  // tsd_timer_start / tsd_timer_stop
  int option_epoch = 0;  // bumped by tsd_set_option: captured graphs are stale afterwards
};

namespace tsd {

// Scoped helper for host-buffer entry points: lock the context, reset (and grow) the arena,
// stage host buffers to the device, collect the first error.
struct HostCall {
  tsd_ctx* h_;
  Ctx* c;
  std::unique_lock<std::mutex> lock_;
  int rc = 0;
  HostCall(tsd_ctx* h, size_t reserve_bytes);
  float* dev(size_t n);
  float* upload(const float* host, size_t n);
  void download(float* host, const float* d, size_t n);
  void run(int code);
  void cu(cudaError_t e, const char* what);
  int finish();
};

}  // namespace tsd
