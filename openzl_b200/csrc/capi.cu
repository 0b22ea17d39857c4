// C ABI of libozl_b200 (see include/ozl.h): argument checking, bases registry, host<->device
// staging and dispatch to the per-curve / per-field translation units.  No CPU compute path.
#include "runtime.cuh"

namespace {

const OzlCurveOps* curve_ops(int curve) {
  switch (curve) {
    case OZL_BLS12_381_G1: return &ozl_ops_bls12_381_g1;
    case OZL_BLS12_381_G2: return &ozl_ops_bls12_381_g2;
    case OZL_BN254_G1: return &ozl_ops_bn254_g1;
    case OZL_BN254_G2: return &ozl_ops_bn254_g2;
  }
  return nullptr;
}

int msm_dispatch(ozl_ctx* ctx, const Bases& b, const uint32_t* d_scalars, size_t n, uint32_t* d_out) {
  const OzlCurveOps* ops = curve_ops(b.curve);
  if (!ops) return OZL_ERR_ARG;
  return ops->msm(ctx, ctx->ws, ctx->stream, b, d_scalars, n, d_out);
}

int find_bases(ozl_ctx* ctx, uint32_t handle, Bases** out) {
  auto it = ctx->bases.find(handle);
  if (it == ctx->bases.end()) return OZL_ERR_HANDLE;
  *out = &it->second;
  return OZL_OK;
}

int alloc_bases(ozl_ctx* ctx, int curve, size_t n, bool with_inf, Bases* b) {
  const int cu = coord_u32(curve);
  if (!cu) return OZL_ERR_ARG;
  b->curve = curve;
  b->n = n;
  CUDA_TRY(ctx, cudaMalloc((void**)&b->d_pts, std::max<size_t>(n, 1) * 2 * cu * 4));
  if (with_inf) {
    const size_t bytes = (((n + 7) / 8 + 3) & ~(size_t)3) + 4;   // whole 32-bit words (k_mark_zero_points ORs words)
    cudaError_t e = cudaMalloc((void**)&b->d_inf, bytes);
    if (e == cudaSuccess) e = cudaMemsetAsync(b->d_inf, 0, bytes, ctx->stream);
    if (e != cudaSuccess) {
      cudaFree(b->d_pts);
      if (b->d_inf) cudaFree(b->d_inf);
      b->d_pts = nullptr;
      b->d_inf = nullptr;
      ctx->last_error = cudaGetErrorString(e);
      return OZL_ERR_OOM;
    }
  }
  return OZL_OK;
}

void release_bases(Bases& b) {
  if (b.d_pts) cudaFree(b.d_pts);
  if (b.d_inf) cudaFree(b.d_inf);
  b.d_pts = nullptr;
  b.d_inf = nullptr;
}

// An all-zero affine point is not on any of the curves (b != 0): ark never produces it, but a caller that
// encodes GroupAffine::infinity as (0, 0) without an inf_mask means the identity.  Marking such points
// in the bitset at upload time makes every path treat them alike (skipped at digit extraction), with
// and without precomputed copies.
__global__ void k_mark_zero_points(const uint32_t* __restrict__ pts, uint32_t n, int aff_words, uint32_t* __restrict__ inf_words) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint4* p = reinterpret_cast<const uint4*>(pts + (size_t)i * aff_words);
  uint32_t t = 0;
  for (int k = 0; k < aff_words / 4; k++) {
    const uint4 v = p[k];
    t |= v.x | v.y | v.z | v.w;
  }
  if (t == 0) atomicOr(&inf_words[i >> 5], 1u << (i & 31));
}

int finish_upload(ozl_ctx* ctx, Bases& b, uint32_t* handle) {
  if (b.n) {
    k_mark_zero_points<<<(unsigned)((b.n + 255) / 256), 256, 0, ctx->stream>>>(b.d_pts, (uint32_t)b.n, 2 * coord_u32(b.curve), (uint32_t*)b.d_inf);
    ctx->launches++;
  }
  cudaError_t e = cudaStreamSynchronize(ctx->stream);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) {
    ctx->last_error = std::string("bases_upload: ") + cudaGetErrorString(e);
    release_bases(b);
    return OZL_ERR_CUDA;
  }
  *handle = ctx->next_handle++;
  ctx->bases[*handle] = b;
  return OZL_OK;
}

// D2H of the result and of the input-error flags, then the verdict of a synchronous MSM call
int finish_msm(ozl_ctx* ctx, const Bases& b, uint64_t* out_jacobian) {
  const size_t out_bytes = 3 * coord_u32(b.curve) * 4;
  uint32_t flags = 0;
  CUDA_TRY(ctx, cudaMemcpyAsync(out_jacobian, ctx->out.p, out_bytes, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, cudaMemcpyAsync(&flags, msm_err_flags(ctx->ws), 4, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  if (flags & 1u) {
    ctx->last_error = "msm: a scalar has bits above the window plan (scalars must be canonical, i.e. reduced modulo r)";
    return OZL_ERR_ARG;
  }
  return OZL_OK;
}

}  // namespace

// =============================================================================================
extern "C" {

int ozl_version(void) { return 100; }

const char* ozl_strerror(int s) {
  switch (s) {
    case OZL_OK: return "ok";
    case OZL_ERR_ARG: return "invalid argument";
    case OZL_ERR_CUDA: return "CUDA error";
    case OZL_ERR_NO_DEVICE: return "no usable CUDA device";
    case OZL_ERR_OOM: return "out of memory";
    case OZL_ERR_HANDLE: return "unknown bases handle";
    case OZL_ERR_DOMAIN: return "domain larger than the field's two-adicity";
    case OZL_ERR_NCCL: return "NCCL unavailable or an NCCL call failed";
  }
  return "unknown status";
}

const char* ozl_last_error(const ozl_ctx* ctx) { return ctx ? ctx->last_error.c_str() : ""; }

int ozl_ctx_create(int device, ozl_ctx** out) {
  if (!out) return OZL_ERR_ARG;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) return OZL_ERR_NO_DEVICE;
  if (device < 0 || device >= count) return OZL_ERR_ARG;
  if (cudaSetDevice(device) != cudaSuccess) return OZL_ERR_NO_DEVICE;
  ozl_ctx* ctx = new (std::nothrow) ozl_ctx();
  if (!ctx) return OZL_ERR_OOM;
  ctx->device = device;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major != 10) {
    delete ctx;   // the library holds sm_100a code only: anything but a Blackwell data-centre part cannot run it
    return OZL_ERR_NO_DEVICE;
  }
  ctx->sm_count = prop.multiProcessorCount;
  if (cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
    delete ctx;
    return OZL_ERR_CUDA;
  }
  ctx->stream = ctx->own_stream;
  *out = ctx;
  return OZL_OK;
}

void ozl_ctx_destroy(ozl_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  stages_clear(ctx);
  if (ctx->pk_deleter) {
    std::map<uint32_t, ozl_rt::Groth16Pk*> pks;
    pks.swap(ctx->pks);
    for (auto& kv : pks) ctx->pk_deleter(ctx, kv.second);   // frees the pk's device buffers and its five bases handles
  }
  for (cudaEvent_t e : ctx->ev_batch)
    if (e) cudaEventDestroy(e);
  if (ctx->ev_prior) cudaEventDestroy(ctx->ev_prior);
  for (auto& kv : ctx->bases) {
    cudaFree(kv.second.d_pts);
    if (kv.second.d_inf) cudaFree(kv.second.d_inf);
  }
  free_workspace(ctx->ws);
  for (int k = 0; k < 2; k++) {
    if (ctx->pipe_scalars[k].p) cudaFree(ctx->pipe_scalars[k].p);
    if (ctx->pipe_out[k].p) cudaFree(ctx->pipe_out[k].p);
    if (ctx->ev_h2d[k]) cudaEventDestroy(ctx->ev_h2d[k]);
    if (ctx->ev_done[k]) cudaEventDestroy(ctx->ev_done[k]);
  }
  if (ctx->copy_stream) { cudaStreamSynchronize(ctx->copy_stream); cudaStreamDestroy(ctx->copy_stream); }
  DevBuf* bufs[] = {&ctx->scalars, &ctx->out};
  for (DevBuf* b : bufs)
    if (b->p) cudaFree(b->p);
  NttWorkspace& nw = ctx->ntt_ws;
  void* nb[] = {nw.scratch, nw.tw[0], nw.tw[1], nw.glo[0], nw.glo[1], nw.ghi[0], nw.ghi[1], nw.consts[0], nw.consts[1]};
  for (void* p : nb)
    if (p) cudaFree(p);
  cudaStreamDestroy(ctx->own_stream);
  delete ctx;
}

int ozl_ctx_set_stream(ozl_ctx* ctx, void* s) {
  if (!ctx) return OZL_ERR_ARG;
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->stream = (cudaStream_t)s;   // NULL is CUDA's legacy default stream
  return OZL_OK;
}
int ozl_ctx_use_own_stream(ozl_ctx* ctx) {
  if (!ctx) return OZL_ERR_ARG;
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->stream = ctx->own_stream;
  return OZL_OK;
}
void* ozl_ctx_get_stream(ozl_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

int ozl_ctx_synchronize(ozl_ctx* ctx) {
  if (!ctx) return OZL_ERR_ARG;
  if (ctx->copy_stream) CUDA_TRY(ctx, cudaStreamSynchronize(ctx->copy_stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return OZL_OK;
}

int ozl_curve_coord_limbs(int curve) { return coord_u32(curve) / 2; }

// ---- bases ----------------------------------------------------------------------------------
int ozl_msm_bases_upload(ozl_ctx* ctx, int curve, const uint64_t* bases, const uint8_t* inf_mask, size_t n,
                         uint32_t* handle) {
  if (!ctx || !handle || (!bases && n) || n >= 0x7fffffffull) return OZL_ERR_ARG;
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  Bases b;
  int r = alloc_bases(ctx, curve, n, true, &b);
  if (r) return r;
  const size_t bytes = n * 2 * coord_u32(curve) * 4;
  cudaError_t e = cudaMemcpyAsync(b.d_pts, bases, bytes, cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess && inf_mask) e = cudaMemcpyAsync(b.d_inf, inf_mask, (n + 7) / 8, cudaMemcpyHostToDevice, ctx->stream);
  if (e != cudaSuccess) {
    ctx->last_error = std::string("bases_upload: ") + cudaGetErrorString(e);
    cudaStreamSynchronize(ctx->stream);
    release_bases(b);
    return OZL_ERR_CUDA;
  }
  return finish_upload(ctx, b, handle);
}

int ozl_msm_bases_upload_device(ozl_ctx* ctx, int curve, const uint64_t* d_bases, const uint8_t* d_inf_mask, size_t n,
                                uint32_t* handle) {
  if (!ctx || !handle || (!d_bases && n) || n >= 0x7fffffffull) return OZL_ERR_ARG;
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  Bases b;
  int r = alloc_bases(ctx, curve, n, true, &b);
  if (r) return r;
  const size_t bytes = n * 2 * coord_u32(curve) * 4;
  cudaError_t e = cudaMemcpyAsync(b.d_pts, d_bases, bytes, cudaMemcpyDeviceToDevice, ctx->stream);
  if (e == cudaSuccess && d_inf_mask) e = cudaMemcpyAsync(b.d_inf, d_inf_mask, (n + 7) / 8, cudaMemcpyDeviceToDevice, ctx->stream);
  if (e != cudaSuccess) {
    ctx->last_error = std::string("bases_upload_device: ") + cudaGetErrorString(e);
    cudaStreamSynchronize(ctx->stream);
    release_bases(b);
    return OZL_ERR_CUDA;
  }
  return finish_upload(ctx, b, handle);
}

int ozl_msm_bases_generate(ozl_ctx* ctx, int curve, uint64_t start, size_t n, uint32_t* handle) {
  if (!ctx || !handle || start == 0 || n >= 0x7fffffffull) return OZL_ERR_ARG;
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  Bases b;
  int r = alloc_bases(ctx, curve, n, false, &b);
  if (r) return r;
  if (n) {
    curve_ops(curve)->generate(ctx->stream, start, (uint32_t)n, b.d_pts);
    ctx->launches++;
  }
  cudaError_t e = cudaStreamSynchronize(ctx->stream);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) {
    ctx->last_error = std::string("bases_generate: ") + cudaGetErrorString(e);
    cudaFree(b.d_pts);
    return OZL_ERR_CUDA;
  }
  *handle = ctx->next_handle++;
  ctx->bases[*handle] = b;
  return OZL_OK;
}

int ozl_msm_bases_download(ozl_ctx* ctx, uint32_t handle, size_t first, size_t n, uint64_t* out) {
  if (!ctx || !out) return OZL_ERR_ARG;
  Bases* b;
  int r = find_bases(ctx, handle, &b);
  if (r) return r;
  if (first + n > b->n) return OZL_ERR_ARG;
  const size_t stride = 2 * coord_u32(b->curve);
  CUDA_TRY(ctx, cudaMemcpyAsync(out, b->d_pts + first * stride, n * stride * 4, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return OZL_OK;
}

int ozl_msm_bases_free(ozl_ctx* ctx, uint32_t handle) {
  if (!ctx) return OZL_ERR_ARG;
  auto it = ctx->bases.find(handle);
  if (it == ctx->bases.end()) return OZL_ERR_HANDLE;
  cudaStreamSynchronize(ctx->stream);
  cudaFree(it->second.d_pts);
  if (it->second.d_inf) cudaFree(it->second.d_inf);
  ctx->bases.erase(it);
  return OZL_OK;
}

int ozl_msm_bases_precompute(ozl_ctx* ctx, uint32_t handle, int factor) {
  if (!ctx || factor < 1 || factor > 32) return OZL_ERR_ARG;
  Bases* b;
  int r = find_bases(ctx, handle, &b);
  if (r) return r;
  if (b->factor != 1) return OZL_ERR_ARG;  // already precomputed
  if (factor == 1 || b->n == 0) return OZL_OK;
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  const MsmPlan p = make_plan(b->curve, b->n, ctx->forced_c, 0, factor);
  const int Wc = (p.W + factor - 1) / factor;
  const int copies = (p.W + Wc - 1) / Wc;        // a factor above the number of windows means one copy per window
  if ((uint64_t)b->n * copies >= 0x7fffffffull) {
    ctx->last_error = "bases_precompute: n * copies exceeds the 31-bit point index of the sorted entries";
    return OZL_ERR_ARG;
  }
  const size_t stride = (size_t)b->n * 2 * coord_u32(b->curve);
  uint32_t* nd = nullptr;
  CUDA_TRY(ctx, cudaMalloc((void**)&nd, stride * copies * 4));
  CUDA_TRY(ctx, cudaMemcpyAsync(nd, b->d_pts, stride * 4, cudaMemcpyDeviceToDevice, ctx->stream));
  const OzlCurveOps* ops = curve_ops(b->curve);
  for (int q = 1; q < copies; q++) {
    ops->precompute(ctx->stream, nd + (size_t)(q - 1) * stride, nd + (size_t)q * stride, (uint32_t)b->n, p.c * Wc);
    ctx->launches++;
  }
  cudaError_t e = cudaStreamSynchronize(ctx->stream);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) {
    cudaFree(nd);
    ctx->last_error = std::string("bases_precompute: ") + cudaGetErrorString(e);
    return OZL_ERR_CUDA;
  }
  cudaFree(b->d_pts);
  b->d_pts = nd;
  b->factor = copies;
  b->pc = p.c;
  b->pWc = Wc;
  return OZL_OK;
}

// ---- msm ------------------------------------------------------------------------------------
int ozl_msm_device_async(ozl_ctx* ctx, uint32_t handle, const uint64_t* d_scalars, size_t n, uint64_t* d_out) {
  if (!ctx || !d_out || (!d_scalars && n)) return OZL_ERR_ARG;
  Bases* b;
  int r = find_bases(ctx, handle, &b);
  if (r) return r;
  if (n > b->n) return OZL_ERR_ARG;
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  if (ctx->timing) stages_clear(ctx);
  return msm_dispatch(ctx, *b, (const uint32_t*)d_scalars, n, (uint32_t*)d_out);
}

int ozl_msm(ozl_ctx* ctx, uint32_t handle, const uint64_t* scalars, size_t n, uint64_t* out_jacobian) {
  if (!ctx || !out_jacobian || (!scalars && n)) return OZL_ERR_ARG;
  Bases* b;
  int r = find_bases(ctx, handle, &b);
  if (r) return r;
  if (n > b->n) return OZL_ERR_ARG;
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  if ((r = ensure(ctx, ctx->out, 1024))) return r;
  if (ctx->timing) stages_clear(ctx);
  // scalars cross PCIe in point-range batches while earlier batches are already being accumulated
  if ((r = ozl_rt_msm_host(ctx, *b, scalars, n, (uint32_t*)ctx->out.p))) return r;
  return finish_msm(ctx, *b, out_jacobian);
}

int ozl_msm_submit(ozl_ctx* ctx, uint32_t handle, const uint64_t* scalars, size_t n, uint64_t* out_jacobian) {
  if (!ctx || !out_jacobian || (!scalars && n)) return OZL_ERR_ARG;
  Bases* b;
  int r = find_bases(ctx, handle, &b);
  if (r) return r;
  if (n > b->n) return OZL_ERR_ARG;
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  if (!ctx->copy_stream) CUDA_TRY(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
  for (int k = 0; k < 2; k++) {
    if (!ctx->ev_h2d[k]) CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->ev_h2d[k], cudaEventDisableTiming));
    if (!ctx->ev_done[k]) CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->ev_done[k], cudaEventDisableTiming));
  }
  const int k = (int)(ctx->pipe_idx++ & 1u);
  const size_t out_bytes = 3 * coord_u32(b->curve) * 4;
  // the staging buffer may still feed the MSM submitted two calls ago
  CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_done[k], 0));
  if ((r = ensure(ctx, ctx->pipe_scalars[k], std::max<size_t>(n, 1) * 32))) return r;
  if ((r = ensure(ctx, ctx->pipe_out[k], 1024))) return r;
  if (n) CUDA_TRY(ctx, cudaMemcpyAsync(ctx->pipe_scalars[k].p, scalars, n * 32, cudaMemcpyHostToDevice, ctx->copy_stream));
  CUDA_TRY(ctx, cudaEventRecord(ctx->ev_h2d[k], ctx->copy_stream));
  CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_h2d[k], 0));
  if (ctx->timing) stages_clear(ctx);
  if ((r = msm_dispatch(ctx, *b, (const uint32_t*)ctx->pipe_scalars[k].p, n, (uint32_t*)ctx->pipe_out[k].p))) return r;
  CUDA_TRY(ctx, cudaMemcpyAsync(out_jacobian, ctx->pipe_out[k].p, out_bytes, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, cudaEventRecord(ctx->ev_done[k], ctx->stream));
  return OZL_OK;
}

int ozl_msm_set_window_bits(ozl_ctx* ctx, int c) {
  if (!ctx || (c != 0 && (c < 2 || c > 24))) return OZL_ERR_ARG;
  ctx->forced_c = c;
  return OZL_OK;
}

int ozl_msm_set_batch_affine(ozl_ctx* ctx, int levels) {
  if (!ctx || levels < -1 || levels > 8) return OZL_ERR_ARG;
  ctx->batch_levels = levels;
  return OZL_OK;
}

int ozl_msm_get_window_bits(ozl_ctx* ctx, int curve, size_t n) {
  if (!ctx || !coord_u32(curve)) return -1;
  return make_plan(curve, n, ctx->forced_c).c;
}


int ozl_msm_bases_info(ozl_ctx* ctx, uint32_t handle, size_t n, int* c, int* windows, int* bucket_sets, int* factor) {
  if (!ctx) return OZL_ERR_ARG;
  Bases* b;
  int r = find_bases(ctx, handle, &b);
  if (r) return r;
  const MsmPlan p = b->factor > 1 ? make_plan(b->curve, n, b->pc, b->pWc) : make_plan(b->curve, n, ctx->forced_c);
  if (c) *c = p.c;
  if (windows) *windows = p.W;
  if (bucket_sets) *bucket_sets = p.Wc;
  if (factor) *factor = b->factor;
  return OZL_OK;
}

int ozl_jacobian_sum(ozl_ctx* ctx, int curve, const uint64_t* points, size_t k, uint64_t* out_jacobian) {
  if (!ctx || !out_jacobian || (!points && k) || !coord_u32(curve)) return OZL_ERR_ARG;
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  const size_t pt_bytes = 3 * coord_u32(curve) * 4;
  int r;
  if ((r = ensure(ctx, ctx->scalars, std::max<size_t>(k, 1) * pt_bytes))) return r;
  if ((r = ensure(ctx, ctx->out, 1024))) return r;
  if (k) CUDA_TRY(ctx, cudaMemcpyAsync(ctx->scalars.p, points, k * pt_bytes, cudaMemcpyHostToDevice, ctx->stream));
  const uint32_t* in = (const uint32_t*)ctx->scalars.p;
  uint32_t* out = (uint32_t*)ctx->out.p;
  curve_ops(curve)->jacobian_sum(ctx->stream, in, (uint32_t)k, out);
  LAUNCH_CHECK(ctx);
  CUDA_TRY(ctx, cudaMemcpyAsync(out_jacobian, out, pt_bytes, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return OZL_OK;
}

int ozl_jacobian_to_affine(ozl_ctx* ctx, int curve, const uint64_t* jacobian, uint64_t* out_affine, int* is_identity) {
  if (!ctx || !jacobian || !out_affine || !is_identity || !coord_u32(curve)) return OZL_ERR_ARG;
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  const size_t cu = coord_u32(curve);
  int r;
  if ((r = ensure(ctx, ctx->scalars, 3 * cu * 4))) return r;
  if ((r = ensure(ctx, ctx->out, 1024))) return r;
  CUDA_TRY(ctx, cudaMemcpyAsync(ctx->scalars.p, jacobian, 3 * cu * 4, cudaMemcpyHostToDevice, ctx->stream));
  const uint32_t* in = (const uint32_t*)ctx->scalars.p;
  uint32_t* out = (uint32_t*)ctx->out.p;
  int* flag = (int*)((char*)ctx->out.p + 512);
  curve_ops(curve)->jacobian_to_affine(ctx->stream, in, out, flag);
  LAUNCH_CHECK(ctx);
  CUDA_TRY(ctx, cudaMemcpyAsync(out_affine, out, 2 * cu * 4, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, cudaMemcpyAsync(is_identity, flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return OZL_OK;
}

// ---- ntt ------------------------------------------------------------------------------------
int ozl_ntt_device_async(ozl_ctx* ctx, int field, uint64_t* d_data, uint32_t log_n, int inverse, int coset) {
  if (!ctx || !d_data) return OZL_ERR_ARG;
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  if (ctx->timing) stages_clear(ctx);
  int launches = 0;
  int r = OZL_ERR_ARG;
  STAGE(ctx, "ntt");
  if (field == OZL_BN254_FR) r = ozl_ntt_run_bn254_fr(ctx->stream, ctx->ntt_ws, (uint32_t*)d_data, log_n, inverse != 0, coset != 0, &launches);
  else if (field == OZL_BLS12_381_FR) r = ozl_ntt_run_bls12_381_fr(ctx->stream, ctx->ntt_ws, (uint32_t*)d_data, log_n, inverse != 0, coset != 0, &launches);
  ctx->launches += launches;
  if (ctx->timing && !ctx->stages.empty()) ctx->stages.back().launches += launches;
  STAGE_END(ctx);
  if (r == -2) return OZL_ERR_DOMAIN;
  if (r == -4) return OZL_ERR_OOM;
  if (r == -3) {
    ctx->last_error = std::string("ntt launch: ") + cudaGetErrorString(cudaGetLastError());
    return OZL_ERR_CUDA;
  }
  return r;
}

int ozl_ntt(ozl_ctx* ctx, int field, uint64_t* data, uint32_t log_n, int inverse, int coset) {
  if (!ctx || !data) return OZL_ERR_ARG;
  if (field != OZL_BN254_FR && field != OZL_BLS12_381_FR) return OZL_ERR_ARG;
  if (log_n > (uint32_t)(field == OZL_BN254_FR ? Bn254Fr::TWO_ADICITY : Bls12381Fr::TWO_ADICITY)) return OZL_ERR_DOMAIN;
  if (log_n > 30) return OZL_ERR_ARG;
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  const size_t bytes = ((size_t)1 << log_n) * 32;
  int r;
  if ((r = ensure(ctx, ctx->scalars, bytes))) return r;
  CUDA_TRY(ctx, cudaMemcpyAsync(ctx->scalars.p, data, bytes, cudaMemcpyHostToDevice, ctx->stream));
  if ((r = ozl_ntt_device_async(ctx, field, (uint64_t*)ctx->scalars.p, log_n, inverse, coset))) return r;
  CUDA_TRY(ctx, cudaMemcpyAsync(data, ctx->scalars.p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return OZL_OK;
}

// ---- instrumentation ------------------------------------------------------------------------
int ozl_ctx_enable_timing(ozl_ctx* ctx, int on) {
  if (!ctx) return OZL_ERR_ARG;
  ctx->timing = on != 0;
  if (!on) stages_clear(ctx);
  return OZL_OK;
}

int ozl_ctx_get_stage_times(ozl_ctx* ctx, ozl_stage_time* out, int cap) {
  if (!ctx || !out) return -1;
  if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) return -1;
  int k = 0;
  for (auto& s : ctx->stages) {
    if (k >= cap) break;
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, s.e0, s.e1) != cudaSuccess) ms = -1.f;
    memset(&out[k], 0, sizeof(out[k]));
    snprintf(out[k].name, sizeof(out[k].name), "%s", s.name.c_str());
    out[k].ms = ms;
    out[k].launches = s.launches;
    k++;
  }
  return k;
}

int ozl_ctx_get_stage_spans(ozl_ctx* ctx, ozl_stage_span* out, int cap) {
  if (!ctx || !out) return -1;
  if (cudaDeviceSynchronize() != cudaSuccess) return -1;   // stages live on several streams
  int k = 0;
  for (auto& s : ctx->stages) {
    if (k >= cap) break;
    float t0 = 0.f, t1 = 0.f;
    if (cudaEventElapsedTime(&t0, ctx->stages.front().e0, s.e0) != cudaSuccess) t0 = -1.f;
    if (cudaEventElapsedTime(&t1, ctx->stages.front().e0, s.e1) != cudaSuccess) t1 = -1.f;
    memset(&out[k], 0, sizeof(out[k]));
    snprintf(out[k].name, sizeof(out[k].name), "%s", s.name.c_str());
    out[k].start_ms = t0;
    out[k].end_ms = t1;
    out[k].launches = s.launches;
    k++;
  }
  return k;
}

uint64_t ozl_ctx_launch_count(const ozl_ctx* ctx) { return ctx ? ctx->launches : 0; }

int ozl_bench_field_mul(ozl_ctx* ctx, int field_id, int iters, double* mul_per_sec) {
  if (!ctx || !mul_per_sec || iters <= 0) return OZL_ERR_ARG;
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  int r;
  if ((r = ensure(ctx, ctx->out, 1024))) return r;
  cudaEvent_t e0, e1;
  CUDA_TRY(ctx, cudaEventCreate(&e0));
  CUDA_TRY(ctx, cudaEventCreate(&e1));
  const int blocks = ctx->sm_count * 8, threads = 128;
  for (int rep = 0; rep < 2; rep++) {
    if (rep == 1) CUDA_TRY(ctx, cudaEventRecord(e0, ctx->stream));
    if (field_id == 0) ozl_ops_bls12_381_g1.bench_mul(ctx->stream, blocks, threads, (uint32_t*)ctx->out.p, iters);
    else if (field_id == 1) ozl_ops_bn254_g1.bench_mul(ctx->stream, blocks, threads, (uint32_t*)ctx->out.p, iters);
    else if (field_id >= 2 && field_id <= 4) ozl_ops_bls12_381_g1.bench_mul_fp64(ctx->stream, blocks, threads, (uint32_t*)ctx->out.p, iters, field_id - 2);
    else return OZL_ERR_ARG;
    LAUNCH_CHECK(ctx);
  }
  CUDA_TRY(ctx, cudaEventRecord(e1, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  float ms = 0.f;
  CUDA_TRY(ctx, cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *mul_per_sec = (double)blocks * threads * iters * 4.0 / (ms * 1e-3);
  return OZL_OK;
}

}  // extern "C"
