// Multi-GPU entry points of the C ABI (include/tsd_b200.h, "multi-GPU"): batch sharding over the GPUs of one box.
//
// The path shards by independent images - one latent per rank, full weight replica per rank - so the only exchange
// steps are ONE broadcast of the CLIP context per prompt (236 544 B per row, pipeline.mojo:41-53 computes it once) and
// an optional gather of the results; there is NO per-step collective (the reference's only batching hints are
// pipeline.mojo:12 and :96-105).  One process per GPU, one NCCL communicator per process, collectives on the context's
// stream.  NCCL is bound at run time (dlopen of libnccl.so.2): libtsd_b200.so keeps no link-time dependency on it and a
// single-GPU caller never loads it.
#include <dlfcn.h>
#include <nccl.h>
#include <unistd.h>

#include <cstdio>
#include <cstring>
#include <mutex>
#include <new>
#include <string>

#include "../../include/tsd_b200.h"
#include "c_api_internal.h"
#include "models.h"

using namespace tsd;

namespace {

struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  std::string error;
};

NcclApi* nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    // an already loaded NCCL (e.g. the one bundled with torch in the same process) is found by soname
    api.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!api.lib) api.lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!api.lib) {
      api.error = std::string("NCCL not found (dlopen libnccl.so.2): ") + dlerror();
      return;
    }
    bool ok = true;
    auto sym = [&](const char* n) {
      void* p = dlsym(api.lib, n);
      if (!p) {
        ok = false;
        api.error = std::string("NCCL symbol missing: ") + n;
      }
      return p;
    };
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
    api.Broadcast = reinterpret_cast<decltype(api.Broadcast)>(sym("ncclBroadcast"));
    api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
    if (!ok) {
      dlclose(api.lib);
      api.lib = nullptr;
    }
  });
  return &api;
}

static_assert(sizeof(ncclUniqueId) == TSD_DIST_ID_BYTES, "ncclUniqueId is 128 bytes");

}  // namespace

struct tsd_dist {
  tsd_ctx* h = nullptr;
  ncclComm_t comm = nullptr;
  int nranks = 1, rank = 0;
  float* stage = nullptr;  // device staging buffer of the collectives
  size_t stage_floats = 0;
};

namespace {

int nccl_fail(tsd_ctx* h, const char* what, ncclResult_t r) {
  NcclApi* a = nccl_api();
  return h->c->fail(TSD_ERR_CUDA, std::string(what) + ": " + (a->GetErrorString ? a->GetErrorString(r) : "NCCL error"));
}

int ensure_stage(tsd_dist* d, size_t floats) {
  if (floats <= d->stage_floats) return TSD_OK;
  if (d->stage) cudaFree(d->stage);
  d->stage = nullptr;
  d->stage_floats = 0;
  if (cudaMalloc(&d->stage, floats * sizeof(float)) != cudaSuccess) {
    cudaGetLastError();
    return d->h->c->fail(TSD_ERR_OOM, "dist: staging allocation failed");
  }
  d->stage_floats = floats;
  return TSD_OK;
}

// File rendezvous for the 128-byte communicator id: rank 0 writes `<path>.tmp` and renames it, the others poll.
int rendezvous_file(tsd_ctx* h, const char* path, int rank, ncclUniqueId* id) {
  NcclApi* a = nccl_api();
  if (rank == 0) {
    ncclResult_t r = a->GetUniqueId(id);
    if (r != ncclSuccess) return nccl_fail(h, "ncclGetUniqueId", r);
    const std::string tmp = std::string(path) + ".tmp";
    FILE* f = fopen(tmp.c_str(), "wb");
    if (!f || fwrite(id, 1, sizeof *id, f) != sizeof *id) {
      if (f) fclose(f);
      return h->c->fail(TSD_ERR_INVALID, std::string("dist: cannot write the rendezvous file ") + path);
    }
    fclose(f);
    if (rename(tmp.c_str(), path) != 0) return h->c->fail(TSD_ERR_INVALID, "dist: rename of the rendezvous file failed");
    return TSD_OK;
  }
  for (int tries = 0; tries < 6000; ++tries) {  // up to 60 s
    FILE* f = fopen(path, "rb");
    if (f) {
      const size_t n = fread(id, 1, sizeof *id, f);
      fclose(f);
      if (n == sizeof *id) return TSD_OK;
    }
    usleep(10000);
  }
  return h->c->fail(TSD_ERR_STATE, std::string("dist: timed out waiting for the rendezvous file ") + path);
}

}  // namespace

extern "C" {

int32_t tsd_dist_unique_id(uint8_t id[TSD_DIST_ID_BYTES]) {
  if (!id) return TSD_ERR_INVALID;
  NcclApi* a = nccl_api();
  if (!a->lib) return TSD_ERR_STATE;
  ncclUniqueId u;
  if (a->GetUniqueId(&u) != ncclSuccess) return TSD_ERR_CUDA;
  memcpy(id, &u, sizeof u);
  return TSD_OK;
}

int32_t tsd_dist_init(tsd_ctx* h, int32_t nranks, int32_t rank, const uint8_t* id, const char* rendezvous_path,
                      tsd_dist** out) {
  if (!h || !out) return TSD_ERR_INVALID;
  *out = nullptr;
  std::lock_guard<std::mutex> g(h->mu);
  if (nranks <= 0 || rank < 0 || rank >= nranks) return h->c->fail(TSD_ERR_INVALID, "dist: bad rank / nranks");
  tsd_dist* d = new (std::nothrow) tsd_dist();
  if (!d) return TSD_ERR_OOM;
  d->h = h;
  d->nranks = nranks;
  d->rank = rank;
  if (nranks > 1) {
    NcclApi* a = nccl_api();
    if (!a->lib) {
      delete d;
      return h->c->fail(TSD_ERR_STATE, a->error);
    }
    ncclUniqueId u;
    if (id) {
      memcpy(&u, id, sizeof u);
    } else {
      const char* path = rendezvous_path ? rendezvous_path : getenv("TSD_DIST_RENDEZVOUS");
      if (!path) {
        delete d;
        return h->c->fail(TSD_ERR_INVALID, "dist: pass the 128-byte id of tsd_dist_unique_id or a rendezvous file path (TSD_DIST_RENDEZVOUS)");
      }
      int rc = rendezvous_file(h, path, rank, &u);
      if (rc) {
        delete d;
        return rc;
      }
    }
    cudaSetDevice(h->c->device);
    ncclResult_t r = a->CommInitRank(&d->comm, nranks, u, rank);
    if (r != ncclSuccess) {
      delete d;
      return nccl_fail(h, "ncclCommInitRank", r);
    }
  }
  *out = d;
  return TSD_OK;
}

int32_t tsd_dist_shutdown(tsd_dist* d) {
  if (!d) return TSD_ERR_INVALID;
  {
    std::lock_guard<std::mutex> g(d->h->mu);
    cudaSetDevice(d->h->c->device);
    cudaStreamSynchronize(d->h->c->stream);
    if (d->comm) nccl_api()->CommDestroy(d->comm);
    if (d->stage) cudaFree(d->stage);
  }
  delete d;
  return TSD_OK;
}

int32_t tsd_dist_rank(const tsd_dist* d) { return d ? d->rank : -1; }
int32_t tsd_dist_size(const tsd_dist* d) { return d ? d->nranks : 0; }

// In place on a host buffer: root's `n_floats` values reach every rank (the (n_ctx,77,768) context, once per prompt).
int32_t tsd_dist_broadcast_context(tsd_dist* d, float* context, int64_t n_floats, int32_t root) {
  if (!d || !context || n_floats <= 0 || root < 0 || root >= d->nranks) return TSD_ERR_INVALID;
  if (d->nranks == 1) return TSD_OK;
  tsd_ctx* h = d->h;
  std::lock_guard<std::mutex> g(h->mu);
  Ctx* c = h->c;
  cudaSetDevice(c->device);
  int rc = ensure_stage(d, (size_t)n_floats);
  if (rc) return rc;
  if (d->rank == root) {
    rc = c->check(cudaMemcpyAsync(d->stage, context, n_floats * sizeof(float), cudaMemcpyHostToDevice, c->stream), "dist H2D");
    if (rc) return rc;
  }
  ncclResult_t r = nccl_api()->Broadcast(d->stage, d->stage, (size_t)n_floats, ncclFloat, root, d->comm, c->stream);
  if (r != ncclSuccess) return nccl_fail(h, "ncclBroadcast", r);
  if (d->rank != root) {
    rc = c->check(cudaMemcpyAsync(context, d->stage, n_floats * sizeof(float), cudaMemcpyDeviceToHost, c->stream), "dist D2H");
    if (rc) return rc;
  }
  return c->check(cudaStreamSynchronize(c->stream), "dist broadcast sync");
}

// Every rank contributes `n_floats` host values; `all` (root only, may be NULL elsewhere) receives nranks x n_floats in
// rank order (decoded images or latents at the end of a job).
int32_t tsd_dist_gather(tsd_dist* d, const float* local, int64_t n_floats, float* all, int32_t root) {
  if (!d || !local || n_floats <= 0 || root < 0 || root >= d->nranks) return TSD_ERR_INVALID;
  if (d->rank == root && !all) return TSD_ERR_INVALID;
  tsd_ctx* h = d->h;
  if (d->nranks == 1) {
    memcpy(all, local, n_floats * sizeof(float));
    return TSD_OK;
  }
  std::lock_guard<std::mutex> g(h->mu);
  Ctx* c = h->c;
  cudaSetDevice(c->device);
  int rc = ensure_stage(d, (size_t)n_floats * (d->nranks + 1));
  if (rc) return rc;
  float* send = d->stage + (size_t)n_floats * d->nranks;
  rc = c->check(cudaMemcpyAsync(send, local, n_floats * sizeof(float), cudaMemcpyHostToDevice, c->stream), "dist H2D");
  if (rc) return rc;
  ncclResult_t r = nccl_api()->AllGather(send, d->stage, (size_t)n_floats, ncclFloat, d->comm, c->stream);
  if (r != ncclSuccess) return nccl_fail(h, "ncclAllGather", r);
  if (d->rank == root) {
    rc = c->check(cudaMemcpyAsync(all, d->stage, (size_t)n_floats * d->nranks * sizeof(float), cudaMemcpyDeviceToHost, c->stream), "dist D2H");
    if (rc) return rc;
  }
  return c->check(cudaStreamSynchronize(c->stream), "dist gather sync");
}

// The denoising loop of this rank's shard: context broadcast from `root` (its K/V projections are then hoisted out of
// the step loop as in tsd_generate_latents), then `n` local latents run `lp->steps` steps.  No per-step collective.
int32_t tsd_dist_generate(tsd_dist* d, tsd_diffusion* m, const tsd_loop_params* lp, const float* latents_in,
                          float* context, int32_t n_ctx, int32_t n, int32_t root, float* latents_out) {
  if (!d || !m || !lp || !context) return TSD_ERR_INVALID;
  const int64_t n_floats = (int64_t)n_ctx * m->m.cfg.context_len * m->m.cfg.context_dim;
  int32_t rc = tsd_dist_broadcast_context(d, context, n_floats, root);
  if (rc) return rc;
  return tsd_generate_latents(m, lp, latents_in, context, n_ctx, n, latents_out);
}

}  // extern "C"
