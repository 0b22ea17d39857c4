// Multi-GPU combine of a point-range-sharded MSM behind the C ABI.
//
// sum_i s_i P_i = sum_g sum_{i in shard g} s_i P_i: every rank runs the complete single-GPU
// pipeline on its shard, the Jacobian partials (144 B for BLS12-381 G1) are exchanged with ONE
// ncclAllGather on the context's stream and summed on the device by every rank.  It is an
// all-gather + local adds rather than an all-reduce because elliptic-curve addition is not an NCCL
// reduction operator.  Everything stays stream-ordered: no host round trip between the shard MSM
// and the combined result.
//
// NCCL is bound at run time (dlopen of libnccl.so.2) so that libozl_b200.so loads on a box without
// NCCL and, inside a PyTorch process, shares the libnccl torch already loaded.
#include <dlfcn.h>
#include <nccl.h>

#include "runtime.cuh"

namespace {

struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  std::string error;
};

NcclApi& nccl() {
  static NcclApi api = []() {
    NcclApi a;
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
      a.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (a.lib) break;
    }
    if (!a.lib) {
      const char* e = dlerror();
      a.error = std::string("libnccl.so.2 not loadable: ") + (e ? e : "?");
      return a;
    }
    a.GetUniqueId = (decltype(a.GetUniqueId))dlsym(a.lib, "ncclGetUniqueId");
    a.CommInitRank = (decltype(a.CommInitRank))dlsym(a.lib, "ncclCommInitRank");
    a.CommDestroy = (decltype(a.CommDestroy))dlsym(a.lib, "ncclCommDestroy");
    a.AllGather = (decltype(a.AllGather))dlsym(a.lib, "ncclAllGather");
    a.GetErrorString = (decltype(a.GetErrorString))dlsym(a.lib, "ncclGetErrorString");
    if (!a.GetUniqueId || !a.CommInitRank || !a.CommDestroy || !a.AllGather || !a.GetErrorString) {
      a.error = "libnccl.so.2 lacks a required symbol";
      a.lib = nullptr;
    }
    return a;
  }();
  return api;
}

}  // namespace

struct ozl_comm {
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1;
  ozl_ctx* ctx = nullptr;
  DevBuf gathered;   // world Jacobian points
};

#define NCCL_TRY(ctx, expr)                                                                          \
  do {                                                                                               \
    ncclResult_t _r = (expr);                                                                        \
    if (_r != ncclSuccess) {                                                                         \
      if (ctx) (ctx)->last_error = std::string(#expr) + ": " + nccl().GetErrorString(_r);            \
      return OZL_ERR_NCCL;                                                                           \
    }                                                                                                \
  } while (0)

extern "C" {

int ozl_comm_unique_id(uint8_t* id_out) {
  if (!id_out) return OZL_ERR_ARG;
  if (!nccl().lib) return OZL_ERR_NCCL;
  ncclUniqueId id;
  if (nccl().GetUniqueId(&id) != ncclSuccess) return OZL_ERR_NCCL;
  static_assert(sizeof(id) == OZL_COMM_ID_BYTES, "ncclUniqueId size");
  memcpy(id_out, &id, sizeof(id));
  return OZL_OK;
}

int ozl_comm_create(ozl_ctx* ctx, const uint8_t* id_in, int rank, int world, ozl_comm** out) {
  if (!ctx || !id_in || !out || world < 1 || rank < 0 || rank >= world) return OZL_ERR_ARG;
  if (!nccl().lib) {
    ctx->last_error = nccl().error;
    return OZL_ERR_NCCL;
  }
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  ncclUniqueId id;
  memcpy(&id, id_in, sizeof(id));
  ozl_comm* c = new ozl_comm();
  c->rank = rank;
  c->world = world;
  c->ctx = ctx;
  ncclResult_t r = nccl().CommInitRank(&c->comm, world, id, rank);
  if (r != ncclSuccess) {
    ctx->last_error = std::string("ncclCommInitRank: ") + nccl().GetErrorString(r);
    delete c;
    return OZL_ERR_NCCL;
  }
  *out = c;
  return OZL_OK;
}

int ozl_comm_destroy(ozl_comm* c) {
  if (!c) return OZL_OK;
  if (c->ctx) cudaSetDevice(c->ctx->device);
  if (c->comm) nccl().CommDestroy(c->comm);
  if (c->gathered.p) cudaFree(c->gathered.p);
  delete c;
  return OZL_OK;
}

// d_partial: this rank's Jacobian partial on the device; d_out: the sum over all ranks (same on
// every rank up to the Jacobian representative: ranks add in the same order, so it is identical).
int ozl_comm_allgather_sum_async(ozl_ctx* ctx, ozl_comm* c, int curve, const uint64_t* d_partial, uint64_t* d_out) {
  if (!ctx || !c || c->ctx != ctx || !d_partial || !d_out || !coord_u32(curve)) return OZL_ERR_ARG;
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  const size_t pt_u32 = 3 * (size_t)coord_u32(curve);
  int r;
  if ((r = ensure(ctx, c->gathered, (size_t)c->world * pt_u32 * 4))) return r;
  NCCL_TRY(ctx, nccl().AllGather(d_partial, c->gathered.p, pt_u32, ncclUint32, c->comm, ctx->stream));
  curve_ops_for(curve)->jacobian_sum(ctx->stream, (const uint32_t*)c->gathered.p, (uint32_t)c->world, (uint32_t*)d_out);
  LAUNCH_CHECK(ctx);
  return OZL_OK;
}

int ozl_msm_sharded_device_async(ozl_ctx* ctx, ozl_comm* c, uint32_t handle, const uint64_t* d_scalars, size_t n,
                                 uint64_t* d_out_jacobian) {
  if (!ctx || !c || c->ctx != ctx || !d_out_jacobian) return OZL_ERR_ARG;
  auto it = ctx->bases.find(handle);
  if (it == ctx->bases.end()) return OZL_ERR_HANDLE;
  const int curve = it->second.curve;
  int r = ozl_msm_device_async(ctx, handle, d_scalars, n, d_out_jacobian);
  if (r) return r;
  // in place: the gather reads the partial before the sum kernel (stream order) overwrites it
  return ozl_comm_allgather_sum_async(ctx, c, curve, d_out_jacobian, d_out_jacobian);
}

int ozl_msm_sharded(ozl_ctx* ctx, ozl_comm* c, uint32_t handle, const uint64_t* scalars, size_t n, uint64_t* out_jacobian) {
  if (!ctx || !c || c->ctx != ctx || !out_jacobian || (!scalars && n)) return OZL_ERR_ARG;
  auto it = ctx->bases.find(handle);
  if (it == ctx->bases.end()) return OZL_ERR_HANDLE;
  const Bases& b = it->second;
  const int curve = b.curve;
  if (n > b.n) return OZL_ERR_ARG;
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  const size_t out_bytes = 3 * (size_t)coord_u32(curve) * 4;
  int r;
  if ((r = ensure(ctx, ctx->out, 1024))) return r;
  if (ctx->timing) stages_clear(ctx);
  // shard MSM with the host->device copy of the shard's scalars hidden under the accumulation
  // (batches, runtime.cuh: ozl_rt_msm_host), then the all-gather + sum, all stream-ordered
  if ((r = ozl_rt_msm_host(ctx, b, scalars, n, (uint32_t*)ctx->out.p))) return r;
  if ((r = ozl_comm_allgather_sum_async(ctx, c, curve, (const uint64_t*)ctx->out.p, (uint64_t*)ctx->out.p))) return r;
  uint32_t flags = 0;
  CUDA_TRY(ctx, cudaMemcpyAsync(out_jacobian, ctx->out.p, out_bytes, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, cudaMemcpyAsync(&flags, msm_err_flags(ctx->ws), 4, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  if (flags & 1u) {
    ctx->last_error = "msm_sharded: a scalar has bits above the window plan (scalars must be canonical)";
    return OZL_ERR_ARG;
  }
  return OZL_OK;
}

}  // extern "C"
