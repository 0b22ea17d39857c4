// Groth16 prover orchestration on the device: the restatement of ark_groth16::create_proof /
// R1CStoQAP::witness_map (ark-groth16 0.3.0) that plugins/arkworks reaches at
// /root/reference/plugins/arkworks/src/groth16.rs:454, with the step order of SURVEY.md section 3.1:
//   a = A z, b = B z, c = C z (instance rows appended to a) -> ifft x3 -> coset_fft x3 ->
//   h = (a*b - c) / Z(g) -> coset_ifft -> MSM(h_query, h) ; MSM(l_query, aux) ; MSM(a_query, z) ;
//   MSM(b_g1_query, z) ; MSM(b_g2_query, z) -> A = alpha + a_acc + r delta, B = beta + b_acc + s delta,
//   C = l_acc + h_acc + s A + r B1 - r s delta  (computed as l + h + s A + r (beta1 + b1_acc)).
// Everything between the H2D copy of z and the D2H copy of the three proof points stays in HBM.
#include "runtime.cuh"
#include "frops.cuh"

namespace ozl_rt {

struct DevCsr {
  uint32_t n_rows = 0;
  uint32_t* row_ptr = nullptr;
  uint32_t* col = nullptr;
  uint32_t* cidx = nullptr;
};

struct Groth16Pk {
  int pairing = 0;
  uint32_t n_constraints = 0, n_instance = 0, n_vars = 0, log_n = 0;
  DevCsr A, B, C;
  uint32_t* coef = nullptr;
  uint32_t h_a = 0, h_b1 = 0, h_b2 = 0, h_h = 0, h_l = 0;
  uint32_t* consts_g1 = nullptr;  // Jacobian alpha1, beta1, delta1
  uint32_t* consts_g2 = nullptr;  // Jacobian beta2, delta2
  uint32_t* tables_g1 = nullptr;  // byte tables (32 Jacobian entries each) of alpha1, beta1, delta1, delta1
  uint32_t* tables_g2 = nullptr;  // byte table of delta2
  uint32_t* zinv = nullptr;       // 1/(g^n - 1)
  // per-proof workspace
  uint32_t *z = nullptr, *zc = nullptr, *a = nullptr, *b = nullptr, *c = nullptr, *hc = nullptr;
  uint32_t *acc = nullptr;        // MSM outputs + assembly scratch
  uint32_t *rs = nullptr;         // scalars for the assembly
  // the five MSMs run on three streams (main: h after the NTTs; s1: a, b1, l; s2: b2)
  cudaStream_t s1 = nullptr, s2 = nullptr, s3 = nullptr;   // s3: the two variable-point scalar multiples of the assembly
  cudaEvent_t ev_z = nullptr, ev_s1 = nullptr, ev_s2 = nullptr, ev_ntt = nullptr, ev_a = nullptr, ev_b1 = nullptr, ev_s3 = nullptr, ev_fixed = nullptr;
  MsmWorkspace ws1, ws2;
  // b_g1 and b_g2 queries have the same points at infinity (checked at creation): the b1 MSM reuses the b2 MSM's
  // digit extraction and bucket sort (same scalars, same plan) instead of repeating them
  bool share_b_sort = false;
};

}  // namespace ozl_rt

namespace {

// Proving keys live in their context (ozl_ctx::pks): one context per thread is the library's
// threading model, so the registry needs no lock, a handle cannot be used with a foreign context, and
// ozl_ctx_destroy releases whatever the caller left behind.
Groth16Pk* find_pk(ozl_ctx* ctx, uint32_t handle) {
  auto it = ctx->pks.find(handle);
  return it == ctx->pks.end() ? nullptr : it->second;
}

void free_csr(DevCsr& m) {
  if (m.row_ptr) cudaFree(m.row_ptr);
  if (m.col) cudaFree(m.col);
  if (m.cidx) cudaFree(m.cidx);
  m = DevCsr();
}

int upload_csr(ozl_ctx* ctx, const ozl_csr* h, DevCsr* d) {
  d->n_rows = h->n_rows;
  const size_t nnz = h->row_ptr[h->n_rows];
  CUDA_TRY(ctx, cudaMalloc((void**)&d->row_ptr, ((size_t)h->n_rows + 1) * 4));
  CUDA_TRY(ctx, cudaMalloc((void**)&d->col, std::max<size_t>(nnz, 1) * 4));
  CUDA_TRY(ctx, cudaMalloc((void**)&d->cidx, std::max<size_t>(nnz, 1) * 4));
  CUDA_TRY(ctx, cudaMemcpyAsync(d->row_ptr, h->row_ptr, ((size_t)h->n_rows + 1) * 4, cudaMemcpyHostToDevice, ctx->stream));
  CUDA_TRY(ctx, cudaMemcpyAsync(d->col, h->col_idx, nnz * 4, cudaMemcpyHostToDevice, ctx->stream));
  CUDA_TRY(ctx, cudaMemcpyAsync(d->cidx, h->coef_idx, nnz * 4, cudaMemcpyHostToDevice, ctx->stream));
  return OZL_OK;
}

const OzlFieldOps* field_ops(int field) {
  if (field == OZL_BN254_FR) return &ozl_fops_bn254_fr;
  if (field == OZL_BLS12_381_FR) return &ozl_fops_bls12_381_fr;
  return nullptr;
}
const OzlCurveOps* g1_ops(int pairing) { return pairing == OZL_PAIRING_BN254 ? &ozl_ops_bn254_g1 : &ozl_ops_bls12_381_g1; }
const OzlCurveOps* g2_ops(int pairing) { return pairing == OZL_PAIRING_BN254 ? &ozl_ops_bn254_g2 : &ozl_ops_bls12_381_g2; }
int g1_curve(int pairing) { return pairing == OZL_PAIRING_BN254 ? OZL_BN254_G1 : OZL_BLS12_381_G1; }
int g2_curve(int pairing) { return pairing == OZL_PAIRING_BN254 ? OZL_BN254_G2 : OZL_BLS12_381_G2; }
int fr_field(int pairing) { return pairing == OZL_PAIRING_BN254 ? OZL_BN254_FR : OZL_BLS12_381_FR; }

void destroy_pk(ozl_ctx* ctx, Groth16Pk* pkp) {
  if (!pkp) return;
  Groth16Pk& pk = *pkp;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  if (pk.s1) cudaStreamSynchronize(pk.s1);
  if (pk.s2) cudaStreamSynchronize(pk.s2);
  free_csr(pk.A); free_csr(pk.B); free_csr(pk.C);
  void* ptrs[] = {pk.tables_g1, pk.tables_g2, pk.coef, pk.consts_g1, pk.consts_g2, pk.zinv, pk.z, pk.zc, pk.a, pk.b, pk.c, pk.hc, pk.acc, pk.rs};
  for (void* p : ptrs) if (p) cudaFree(p);
  for (uint32_t h : {pk.h_a, pk.h_b1, pk.h_b2, pk.h_h, pk.h_l})
    if (h) ozl_msm_bases_free(ctx, h);
  free_workspace(pk.ws1);
  free_workspace(pk.ws2);
  if (pk.s1) cudaStreamDestroy(pk.s1);
  if (pk.s2) cudaStreamDestroy(pk.s2);
  if (pk.s3) cudaStreamDestroy(pk.s3);
  for (cudaEvent_t e : {pk.ev_a, pk.ev_b1, pk.ev_s3, pk.ev_fixed})
    if (e) cudaEventDestroy(e);
  for (cudaEvent_t e : {pk.ev_z, pk.ev_s1, pk.ev_s2, pk.ev_ntt}) if (e) cudaEventDestroy(e);
  delete pkp;
}

}  // namespace

extern "C" {

int ozl_fr_spmv(ozl_ctx* ctx, int field, const ozl_csr* M, const uint64_t* coef_table, uint32_t n_coef,
                const uint64_t* x, uint32_t n_cols, uint64_t* y) {
  if (!ctx || !M || !M->row_ptr || !coef_table || !x || !y) return OZL_ERR_ARG;
  const OzlFieldOps* f = field_ops(field);
  if (!f) return OZL_ERR_ARG;
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  DevCsr d;
  int r = upload_csr(ctx, M, &d);
  uint32_t *dc = nullptr, *dx = nullptr, *dy = nullptr;
  auto cleanup = [&]() { free_csr(d); if (dc) cudaFree(dc); if (dx) cudaFree(dx); if (dy) cudaFree(dy); };
  if (r) { cleanup(); return r; }
  if (cudaMalloc((void**)&dc, std::max<size_t>(n_coef, 1) * 32) != cudaSuccess ||
      cudaMalloc((void**)&dx, std::max<size_t>(n_cols, 1) * 32) != cudaSuccess ||
      cudaMalloc((void**)&dy, std::max<size_t>(M->n_rows, 1) * 32) != cudaSuccess) { cleanup(); return OZL_ERR_OOM; }
  cudaMemcpyAsync(dc, coef_table, (size_t)n_coef * 32, cudaMemcpyHostToDevice, ctx->stream);
  cudaMemcpyAsync(dx, x, (size_t)n_cols * 32, cudaMemcpyHostToDevice, ctx->stream);
  f->spmv(ctx->stream, d.row_ptr, d.col, d.cidx, dc, dx, M->n_rows, dy);
  ctx->launches++;
  cudaMemcpyAsync(y, dy, (size_t)M->n_rows * 32, cudaMemcpyDeviceToHost, ctx->stream);
  cudaError_t e = cudaStreamSynchronize(ctx->stream);
  if (e == cudaSuccess) e = cudaGetLastError();
  cleanup();
  if (e != cudaSuccess) { ctx->last_error = std::string("fr_spmv: ") + cudaGetErrorString(e); return OZL_ERR_CUDA; }
  return OZL_OK;
}

int ozl_fr_poseidon_permute(ozl_ctx* ctx, int field, uint64_t* states, size_t batch, uint32_t width, uint32_t full_rounds,
                            uint32_t partial_rounds, const uint64_t* round_keys, const uint64_t* mds) {
  if (!ctx || !states || !round_keys || !mds) return OZL_ERR_ARG;
  if (width < 2 || width > (uint32_t)ozl::POSEIDON_MAX_WIDTH || (full_rounds & 1) || full_rounds + partial_rounds == 0 ||
      full_rounds + partial_rounds > 4096 || batch >= 0x7fffffffull / width) return OZL_ERR_ARG;
  const OzlFieldOps* f = field_ops(field);
  if (!f) return OZL_ERR_ARG;
  if (!batch) return OZL_OK;
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  const size_t st_bytes = batch * width * 32, rk_bytes = (size_t)(full_rounds + partial_rounds) * width * 32, mds_bytes = (size_t)width * width * 32;
  uint32_t *ds = nullptr, *dk = nullptr, *dm = nullptr;
  auto cleanup = [&]() { if (ds) cudaFree(ds); if (dk) cudaFree(dk); if (dm) cudaFree(dm); };
  if (cudaMalloc((void**)&ds, st_bytes) != cudaSuccess || cudaMalloc((void**)&dk, rk_bytes) != cudaSuccess ||
      cudaMalloc((void**)&dm, mds_bytes) != cudaSuccess) { cleanup(); cudaGetLastError(); return OZL_ERR_OOM; }
  cudaMemcpyAsync(ds, states, st_bytes, cudaMemcpyHostToDevice, ctx->stream);
  cudaMemcpyAsync(dk, round_keys, rk_bytes, cudaMemcpyHostToDevice, ctx->stream);
  cudaMemcpyAsync(dm, mds, mds_bytes, cudaMemcpyHostToDevice, ctx->stream);
  f->poseidon(ctx->stream, ds, (uint32_t)batch, (int)width, (int)full_rounds, (int)partial_rounds, dk, dm);
  ctx->launches++;
  cudaMemcpyAsync(states, ds, st_bytes, cudaMemcpyDeviceToHost, ctx->stream);
  cudaError_t e = cudaStreamSynchronize(ctx->stream);
  if (e == cudaSuccess) e = cudaGetLastError();
  cleanup();
  if (e != cudaSuccess) { ctx->last_error = std::string("fr_poseidon_permute: ") + cudaGetErrorString(e); return OZL_ERR_CUDA; }
  return OZL_OK;
}

int ozl_fixed_base_mul(ozl_ctx* ctx, int curve, const uint64_t* scalars, size_t n, uint64_t* out_affine,
                       uint8_t* identity_flags) {
  if (!ctx || (!scalars && n) || !out_affine || !identity_flags || !coord_u32(curve) || n >= 0x7fffffffull) return OZL_ERR_ARG;
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  const OzlCurveOps* ops = nullptr;
  switch (curve) {
    case OZL_BLS12_381_G1: ops = &ozl_ops_bls12_381_g1; break;
    case OZL_BLS12_381_G2: ops = &ozl_ops_bls12_381_g2; break;
    case OZL_BN254_G1: ops = &ozl_ops_bn254_g1; break;
    case OZL_BN254_G2: ops = &ozl_ops_bn254_g2; break;
  }
  if (!n) return OZL_OK;
  const size_t aff_bytes = 2 * (size_t)coord_u32(curve) * 4;
  uint32_t *ds = nullptr, *dout = nullptr;
  uint8_t* dfl = nullptr;
  auto cleanup = [&]() { if (ds) cudaFree(ds); if (dout) cudaFree(dout); if (dfl) cudaFree(dfl); };
  if (cudaMalloc((void**)&ds, n * 32) != cudaSuccess || cudaMalloc((void**)&dout, n * aff_bytes) != cudaSuccess ||
      cudaMalloc((void**)&dfl, n) != cudaSuccess) { cleanup(); return OZL_ERR_OOM; }
  cudaMemcpyAsync(ds, scalars, n * 32, cudaMemcpyHostToDevice, ctx->stream);
  ops->fixed_base_mul(ctx->stream, ds, (uint32_t)n, dout, dfl);
  ctx->launches++;
  cudaMemcpyAsync(out_affine, dout, n * aff_bytes, cudaMemcpyDeviceToHost, ctx->stream);
  cudaMemcpyAsync(identity_flags, dfl, n, cudaMemcpyDeviceToHost, ctx->stream);
  cudaError_t e = cudaStreamSynchronize(ctx->stream);
  if (e == cudaSuccess) e = cudaGetLastError();
  cleanup();
  if (e != cudaSuccess) { ctx->last_error = std::string("fixed_base_mul: ") + cudaGetErrorString(e); return OZL_ERR_CUDA; }
  return OZL_OK;
}

}  // extern "C"

namespace {
// device-side part of pk_create; any failure leaves *pkp for destroy_pk (the caller keeps its bases handles)
int pk_build(ozl_ctx* ctx, Groth16Pk* pkp, const ozl_csr* A, const ozl_csr* B, const ozl_csr* C, const uint64_t* coef_table,
             uint32_t n_coef, const uint64_t* alpha_g1, const uint64_t* beta_g1, const uint64_t* delta_g1,
             const uint64_t* beta_g2, const uint64_t* delta_g2) {
  Groth16Pk& pk = *pkp;
  const int pairing = pk.pairing;
  const uint32_t n_vars = pk.n_vars, log_n = pk.log_n;
  const size_t n = (size_t)1 << log_n;
  int r;
  if ((r = upload_csr(ctx, A, &pk.A)) || (r = upload_csr(ctx, B, &pk.B)) || (r = upload_csr(ctx, C, &pk.C))) return r;
  const int c1 = coord_u32(g1_curve(pairing)), c2 = coord_u32(g2_curve(pairing));
  CUDA_TRY(ctx, cudaMalloc((void**)&pk.coef, std::max<size_t>(n_coef, 1) * 32));
  CUDA_TRY(ctx, cudaMemcpyAsync(pk.coef, coef_table, (size_t)n_coef * 32, cudaMemcpyHostToDevice, ctx->stream));
  CUDA_TRY(ctx, cudaMalloc((void**)&pk.consts_g1, 3 * 3 * c1 * 4));
  CUDA_TRY(ctx, cudaMalloc((void**)&pk.consts_g2, 2 * 3 * c2 * 4));
  CUDA_TRY(ctx, cudaMalloc((void**)&pk.zinv, 64));
  CUDA_TRY(ctx, cudaMalloc((void**)&pk.z, (size_t)n_vars * 32));
  CUDA_TRY(ctx, cudaMalloc((void**)&pk.zc, (size_t)n_vars * 32));
  CUDA_TRY(ctx, cudaMalloc((void**)&pk.a, n * 32));
  CUDA_TRY(ctx, cudaMalloc((void**)&pk.b, n * 32));
  CUDA_TRY(ctx, cudaMalloc((void**)&pk.c, n * 32));
  CUDA_TRY(ctx, cudaMalloc((void**)&pk.hc, n * 32));
  CUDA_TRY(ctx, cudaMalloc((void**)&pk.acc, 24 * 3 * c2 * 4));
  CUDA_TRY(ctx, cudaMalloc((void**)&pk.rs, 32 * 32));
  // constants: affine (host) -> Jacobian (device)
  if ((r = ensure(ctx, ctx->out, 4096))) return r;
  const uint64_t* g1c[3] = {alpha_g1, beta_g1, delta_g1};
  for (int i = 0; i < 3; i++) {
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->out.p, g1c[i], 2 * c1 * 4, cudaMemcpyHostToDevice, ctx->stream));
    g1_ops(pairing)->affine_to_jacobian(ctx->stream, (const uint32_t*)ctx->out.p, pk.consts_g1 + (size_t)i * 3 * c1);
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  }
  const uint64_t* g2c[2] = {beta_g2, delta_g2};
  for (int i = 0; i < 2; i++) {
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->out.p, g2c[i], 2 * c2 * 4, cudaMemcpyHostToDevice, ctx->stream));
    g2_ops(pairing)->affine_to_jacobian(ctx->stream, (const uint32_t*)ctx->out.p, pk.consts_g2 + (size_t)i * 3 * c2);
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  }
  // byte tables for the fixed points of the proof assembly: alpha1, beta1, delta1, delta1 | delta2
  CUDA_TRY(ctx, cudaMalloc((void**)&pk.tables_g1, (size_t)4 * 32 * 3 * c1 * 4));
  CUDA_TRY(ctx, cudaMalloc((void**)&pk.tables_g2, (size_t)1 * 32 * 3 * c2 * 4));
  {
    const int order[4] = {0, 1, 2, 2};
    for (int i = 0; i < 4; i++)
      g1_ops(pairing)->build_byte_table(ctx->stream, pk.consts_g1 + (size_t)order[i] * 3 * c1, pk.tables_g1 + (size_t)i * 32 * 3 * c1);
    g2_ops(pairing)->build_byte_table(ctx->stream, pk.consts_g2 + (size_t)1 * 3 * c2, pk.tables_g2);
    ctx->launches += 5;
  }
  {
    // s1 carries the longest chain of a proof (a, b1, l MSMs): greatest priority, so that wherever another stream's
    // accumulation hands an SM back (yield_ctas) the chain's kernels go first.  OZL_G16_PRIO=0: all default.
    static const bool kPrio = []() { const char* e = getenv("OZL_G16_PRIO"); return !(e && e[0] == '0'); }();
    static const bool kYield = []() { const char* e = getenv("OZL_G16_YIELD"); return !(e && e[0] == '0'); }();
    int lo = 0, hi = 0;
    CUDA_TRY(ctx, cudaDeviceGetStreamPriorityRange(&lo, &hi));
    CUDA_TRY(ctx, cudaStreamCreateWithPriority(&pk.s1, cudaStreamNonBlocking, kPrio ? hi : lo));
    CUDA_TRY(ctx, cudaStreamCreateWithPriority(&pk.s3, cudaStreamNonBlocking, kPrio ? hi : lo));
    CUDA_TRY(ctx, cudaStreamCreateWithFlags(&pk.s2, cudaStreamNonBlocking));
    pk.ws1.yield_ctas = pk.ws2.yield_ctas = kYield;
  }
  CUDA_TRY(ctx, cudaEventCreateWithFlags(&pk.ev_a, cudaEventDisableTiming));
  CUDA_TRY(ctx, cudaEventCreateWithFlags(&pk.ev_b1, cudaEventDisableTiming));
  CUDA_TRY(ctx, cudaEventCreateWithFlags(&pk.ev_s3, cudaEventDisableTiming));
  CUDA_TRY(ctx, cudaEventCreateWithFlags(&pk.ev_fixed, cudaEventDisableTiming));
  CUDA_TRY(ctx, cudaEventCreateWithFlags(&pk.ev_z, cudaEventDisableTiming));
  CUDA_TRY(ctx, cudaEventCreateWithFlags(&pk.ev_s1, cudaEventDisableTiming));
  CUDA_TRY(ctx, cudaEventCreateWithFlags(&pk.ev_s2, cudaEventDisableTiming));
  CUDA_TRY(ctx, cudaEventCreateWithFlags(&pk.ev_ntt, cudaEventDisableTiming));
  field_ops(fr_field(pairing))->vanishing_inv(ctx->stream, (int)log_n, pk.zinv);
  ctx->launches += 6;
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return OZL_OK;
}
}  // namespace

extern "C" {

int ozl_groth16_pk_create(ozl_ctx* ctx, int pairing, uint32_t n_constraints, uint32_t n_instance, uint32_t n_vars,
                          const ozl_csr* A, const ozl_csr* B, const ozl_csr* C, const uint64_t* coef_table,
                          uint32_t n_coef, uint32_t a_query, uint32_t b_g1_query, uint32_t b_g2_query,
                          uint32_t h_query, uint32_t l_query, const uint64_t* alpha_g1, const uint64_t* beta_g1,
                          const uint64_t* delta_g1, const uint64_t* beta_g2, const uint64_t* delta_g2,
                          uint32_t* pk_handle) {
  if (!ctx || !A || !B || !C || !coef_table || !alpha_g1 || !beta_g1 || !delta_g1 || !beta_g2 || !delta_g2 || !pk_handle)
    return OZL_ERR_ARG;
  if (pairing != OZL_PAIRING_BN254 && pairing != OZL_PAIRING_BLS12_381) return OZL_ERR_ARG;
  if (A->n_rows != n_constraints || B->n_rows != n_constraints || C->n_rows != n_constraints) return OZL_ERR_ARG;
  if (n_instance == 0 || n_instance > n_vars) return OZL_ERR_ARG;
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  const uint64_t need = (uint64_t)n_constraints + n_instance;   // ark: domain over num_constraints + num_inputs
  uint32_t log_n = 0;
  while (((uint64_t)1 << log_n) < need) log_n++;
  const int two_adicity = pairing == OZL_PAIRING_BN254 ? 28 : 32;
  if ((int)log_n > two_adicity || log_n > 30) return OZL_ERR_DOMAIN;
  const size_t n = (size_t)1 << log_n;
  // handle sanity
  Bases *ba, *bb1, *bb2, *bh, *bl;
  for (auto hp : {std::make_pair(a_query, &ba), std::make_pair(b_g1_query, &bb1), std::make_pair(b_g2_query, &bb2),
                  std::make_pair(h_query, &bh), std::make_pair(l_query, &bl)}) {
    auto it = ctx->bases.find(hp.first);
    if (it == ctx->bases.end()) return OZL_ERR_HANDLE;
    *hp.second = &it->second;
  }
  if (ba->curve != g1_curve(pairing) || bb1->curve != g1_curve(pairing) || bh->curve != g1_curve(pairing) ||
      bl->curve != g1_curve(pairing) || bb2->curve != g2_curve(pairing)) return OZL_ERR_ARG;
  if (ba->n < n_vars || bb1->n < n_vars || bb2->n < n_vars || bh->n < n - 1 || bl->n < n_vars - n_instance) return OZL_ERR_ARG;
  Groth16Pk* pkp = new (std::nothrow) Groth16Pk();
  if (!pkp) return OZL_ERR_OOM;
  pkp->pairing = pairing;
  pkp->n_constraints = n_constraints; pkp->n_instance = n_instance; pkp->n_vars = n_vars;
  pkp->log_n = log_n;
  const int rc = pk_build(ctx, pkp, A, B, C, coef_table, n_coef, alpha_g1, beta_g1, delta_g1, beta_g2, delta_g2);
  if (rc) {
    destroy_pk(ctx, pkp);   // the five bases handles are not the pk's yet: the caller keeps them
    return rc;
  }
  pkp->h_a = a_query; pkp->h_b1 = b_g1_query; pkp->h_b2 = b_g2_query; pkp->h_h = h_query; pkp->h_l = l_query;
  {
    // same infinity sets?  (b_i(tau) = 0 kills the point in both groups; compared rather than assumed)
    static const bool kShare = []() { const char* e = getenv("OZL_G16_SHARE_SORT"); return !(e && e[0] == '0'); }();
    bool same = kShare && bb1->n == bb2->n && (bb1->d_inf == nullptr) == (bb2->d_inf == nullptr);
    if (same && bb1->d_inf) {
      const size_t bytes = (bb1->n + 7) / 8;
      std::vector<uint8_t> m1(bytes), m2(bytes);
      same = cudaMemcpy(m1.data(), bb1->d_inf, bytes, cudaMemcpyDeviceToHost) == cudaSuccess &&
             cudaMemcpy(m2.data(), bb2->d_inf, bytes, cudaMemcpyDeviceToHost) == cudaSuccess && m1 == m2;
    }
    pkp->share_b_sort = same;
  }
  ctx->pk_deleter = destroy_pk;
  *pk_handle = ctx->next_pk++;
  ctx->pks[*pk_handle] = pkp;
  return OZL_OK;
}

int ozl_groth16_pk_destroy(ozl_ctx* ctx, uint32_t pk_handle) {
  if (!ctx) return OZL_ERR_ARG;
  auto it = ctx->pks.find(pk_handle);
  if (it == ctx->pks.end()) return OZL_ERR_HANDLE;
  Groth16Pk* pkp = it->second;
  ctx->pks.erase(it);
  destroy_pk(ctx, pkp);
  return OZL_OK;
}

int ozl_groth16_domain_size(ozl_ctx* ctx, uint32_t pk_handle, uint32_t* out) {
  if (!ctx || !out) return OZL_ERR_ARG;
  Groth16Pk* pkp = find_pk(ctx, pk_handle);
  if (!pkp) return OZL_ERR_HANDLE;
  *out = 1u << pkp->log_n;
  return OZL_OK;
}

int ozl_groth16_prove(ozl_ctx* ctx, uint32_t pk_handle, const uint64_t* z, const uint64_t* r_scalar, const uint64_t* s_scalar,
                      uint64_t* proof_a, uint64_t* proof_b, uint64_t* proof_c, uint64_t* h_out) {
  if (!ctx || !z || !r_scalar || !s_scalar || !proof_a || !proof_b || !proof_c) return OZL_ERR_ARG;
  Groth16Pk* pkp = find_pk(ctx, pk_handle);
  if (!pkp) return OZL_ERR_HANDLE;      // unknown, destroyed, or created on another context
  Groth16Pk& pk = *pkp;
  const Bases *q_a, *q_b1, *q_b2, *q_h, *q_l;
  for (auto hp : {std::make_pair(pk.h_a, &q_a), std::make_pair(pk.h_b1, &q_b1), std::make_pair(pk.h_b2, &q_b2),
                  std::make_pair(pk.h_h, &q_h), std::make_pair(pk.h_l, &q_l)}) {
    auto it = ctx->bases.find(hp.first);
    if (it == ctx->bases.end()) return OZL_ERR_HANDLE;
    *hp.second = &it->second;
  }
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  const OzlFieldOps* f = field_ops(fr_field(pk.pairing));
  const OzlCurveOps* o1 = g1_ops(pk.pairing);
  const OzlCurveOps* o2 = g2_ops(pk.pairing);
  const int c1 = coord_u32(g1_curve(pk.pairing)), c2 = coord_u32(g2_curve(pk.pairing));
  const int J1 = 3 * c1, J2 = 3 * c2;  // u32 per Jacobian point
  const size_t n = (size_t)1 << pk.log_n;
  cudaStream_t st = ctx->stream;
  const uint32_t nc = pk.n_constraints, ni = pk.n_instance, m = pk.n_vars;
  int rc;
  if (ctx->timing) stages_clear(ctx);
  // NVTX ranges named after ark-groth16's own start_timer! labels (prover.rs / r1cs_to_qap.rs)
  struct NvtxScope {
    explicit NvtxScope(const char* n) { nvtxRangePushA(n); }
    ~NvtxScope() { nvtxRangePop(); }
  } prover_range("Groth16::Prover");

  STAGE(ctx, "g16_h2d_witness");
  CUDA_TRY(ctx, cudaMemcpyAsync(pk.z, z, (size_t)m * 32, cudaMemcpyHostToDevice, st));
  CUDA_TRY(ctx, cudaMemcpyAsync(pk.rs, r_scalar, 32, cudaMemcpyHostToDevice, st));
  CUDA_TRY(ctx, cudaMemcpyAsync(pk.rs + 8, s_scalar, 32, cudaMemcpyHostToDevice, st));
  STAGE_END(ctx);

  nvtxRangePushA("R1CS to QAP witness map");
  STAGE(ctx, "g16_matvec");
  CUDA_TRY(ctx, cudaMemsetAsync(pk.a, 0, n * 32, st));
  CUDA_TRY(ctx, cudaMemsetAsync(pk.b, 0, n * 32, st));
  CUDA_TRY(ctx, cudaMemsetAsync(pk.c, 0, n * 32, st));
  f->spmv(st, pk.A.row_ptr, pk.A.col, pk.A.cidx, pk.coef, pk.z, nc, pk.a);
  f->spmv(st, pk.B.row_ptr, pk.B.col, pk.B.cidx, pk.coef, pk.z, nc, pk.b);
  f->spmv(st, pk.C.row_ptr, pk.C.col, pk.C.cidx, pk.coef, pk.z, nc, pk.c);
  ctx->launches += 3;
  // ark: a[num_constraints + j] = full_assignment[j] for the num_inputs instance variables
  CUDA_TRY(ctx, cudaMemcpyAsync(pk.a + (size_t)nc * 8, pk.z, (size_t)ni * 32, cudaMemcpyDeviceToDevice, st));
  f->from_mont(st, pk.z, pk.zc, m);
  ctx->launches += 1;
  STAGE_END(ctx);
  CUDA_TRY(ctx, cudaEventRecord(pk.ev_z, st));
  // The witness map's seven transforms are enqueued before the side-stream MSMs (kernels start in enqueue order as
  // resources allow).  They still take ~10.8 ms of wall time for 2 ms of work, interleaved with the accumulation
  // kernels of the other streams; that sharing is what packs the proof into 24.9 ms (see the gate experiment below).
  STAGE(ctx, "g16_ntt");
  {
    int launches = 0;
    NttWorkspace& ws = ctx->ntt_ws;
    uint32_t* vecs[3] = {pk.a, pk.b, pk.c};
    int nrc = 0;
    for (int i = 0; i < 3 && !nrc; i++) nrc = f->ntt(st, ws, vecs[i], pk.log_n, true, false, &launches);
    for (int i = 0; i < 3 && !nrc; i++) nrc = f->ntt(st, ws, vecs[i], pk.log_n, false, true, &launches);
    if (!nrc) {
      f->h_pointwise(st, pk.a, pk.b, pk.c, pk.zinv, (uint32_t)n);
      launches++;
      nrc = f->ntt(st, ws, pk.a, pk.log_n, true, true, &launches);
    }
    if (!nrc) {
      f->from_mont(st, pk.a, pk.hc, (uint32_t)n);
      launches++;
    }
    ctx->launches += launches;
    if (ctx->timing && !ctx->stages.empty()) ctx->stages.back().launches += launches;
    if (nrc) {
      nvtxRangePop();
      nvtxRangePop();
      if (nrc == -2) return OZL_ERR_DOMAIN;
      if (nrc == -4) return OZL_ERR_OOM;
      ctx->last_error = "groth16_prove: ntt launch failed";
      return OZL_ERR_CUDA;
    }
  }
  STAGE_END(ctx);
  nvtxRangePop();   // R1CS to QAP witness map
  {
    // Measured and left OFF: holding the side-stream accumulations back until the transforms are through brings the
    // transform stage from 10.8 to 2.4 ms but the proof from 24.85 to 26.2 ms -- sharing the SMs was the better packing.
    static const bool kGate = []() { const char* e = getenv("OZL_G16_GATE"); return e && e[0] == '1'; }();
    CUDA_TRY(ctx, cudaEventRecord(pk.ev_ntt, st));
    pk.ws1.accumulate_gate = pk.ws2.accumulate_gate = kGate ? pk.ev_ntt : nullptr;
  }
  // the four MSMs over the assignment do not depend on the NTTs: they go to the side streams
  uint32_t* const acc = pk.acc;
  uint32_t* const acc_g2 = pk.acc + 16 * J2;
  CUDA_TRY(ctx, cudaStreamWaitEvent(pk.s1, pk.ev_z, 0));
  CUDA_TRY(ctx, cudaStreamWaitEvent(pk.s2, pk.ev_z, 0));
  // Proof assembly terms are issued where their inputs become available, so only the final sums and
  // the affine conversion remain after the last MSM:
  //   s2: [s]alpha1, [r]beta1, [rs]delta1, [r]delta1, [s]delta2 (fixed tables; need only r, s), then the b2 MSM
  //   s1: a MSM, b1 MSM, then [s]a_acc and [r]b1_acc, then the l MSM
  static const uint32_t one[8] = {1, 0, 0, 0, 0, 0, 0, 0};
  uint32_t* const S = pk.rs;                // 16+ scalars x 8 words: [0]=r [1]=s [2]=1 [3]=rs
  uint32_t* const fs = S + 32;              // 4 scalars for the fixed G1 tables (alpha1, beta1, delta1, delta1)
  uint32_t* const vs = S + 64;              // 2 scalars for the variable points
  uint32_t* const fixed_out = acc + 12 * J1;     // s*alpha1, r*beta1, rs*delta1, r*delta1
  uint32_t* const var_out = acc + 16 * J1;       // s*a_acc, r*b1_acc
  uint32_t* const g2_fixed_out = acc_g2 + J2;    // s*delta2
  auto cp_on = [&](cudaStream_t q, uint32_t* dst, const uint32_t* src, size_t words) {
    return cudaMemcpyAsync(dst, src, words * 4, cudaMemcpyDeviceToDevice, q);
  };
  // A and B are assembled and converted to affine where their last input appears (A on s3 after the a MSM, B on s2
  // after the b2 MSM) instead of after the last MSM: the tail of a proof is then C alone.  OZL_G16_EARLY_AB=0: all at the end.
  static const bool kEarlyAB = []() { const char* e = getenv("OZL_G16_EARLY_AB"); return !(e && e[0] == '0'); }();
  static const bool kS3 = []() { const char* e = getenv("OZL_G16_S3"); return !(e && e[0] == '0'); }();
  const bool early_ab = kEarlyAB && kS3;
  if ((rc = ensure(ctx, ctx->out, 4096))) return rc;
  uint32_t* const outb = (uint32_t*)ctx->out.p;
  int* const flags = (int*)(outb + 512);
  uint32_t* const ones = S + 80;                  // 8 unit scalars
  uint32_t* const A_jac = acc + 8 * J1;
  uint32_t* const C_jac = A_jac + J1;
  uint32_t* const pts2 = acc_g2 + 2 * J2;
  uint32_t* const B_jac = pts2 + 3 * J2;
  {
    nvtxRangePushA("Compute A / Compute B (a, b_g1, b_g2 query MSMs) + l query MSM");
    CUDA_TRY(ctx, cudaMemcpyAsync(S + 16, one, 32, cudaMemcpyHostToDevice, pk.s2));
    for (int i = 0; i < 8; i++) CUDA_TRY(ctx, cp_on(pk.s2, ones + 8 * i, S + 16, 8));
    f->mul_canonical(pk.s2, S + 0, S + 8, S + 24);
    CUDA_TRY(ctx, cp_on(pk.s2, fs + 0, S + 8, 8));
    CUDA_TRY(ctx, cp_on(pk.s2, fs + 8, S + 0, 8));
    CUDA_TRY(ctx, cp_on(pk.s2, fs + 16, S + 24, 8));
    CUDA_TRY(ctx, cp_on(pk.s2, fs + 24, S + 0, 8));
    o1->scalar_mul_table(pk.s2, pk.tables_g1, fs, 4, fixed_out);
    o2->scalar_mul_table(pk.s2, pk.tables_g2, S + 8, 1, g2_fixed_out);
    CUDA_TRY(ctx, cudaEventRecord(pk.ev_fixed, pk.s2));
    pk.ws2.publish_sort = pk.share_b_sort;
    if ((rc = ozl_rt_msm(ctx, pk.ws2, pk.s2, *q_b2, pk.zc, m, acc_g2))) return rc;
    if (early_ab) {   // B = beta2 + b2_acc + [s]delta2
      CUDA_TRY(ctx, cp_on(pk.s2, pts2 + 0 * J2, pk.consts_g2 + 0 * J2, J2));
      CUDA_TRY(ctx, cp_on(pk.s2, pts2 + 1 * J2, acc_g2, J2));
      CUDA_TRY(ctx, cp_on(pk.s2, pts2 + 2 * J2, g2_fixed_out, J2));
      o2->lincomb(pk.s2, pts2, ones, 3, B_jac);
      o2->jacobian_to_affine(pk.s2, B_jac, outb + 4 * c1, flags + 2);
      ctx->launches += 2;
    }
    if ((rc = ozl_rt_msm(ctx, pk.ws1, pk.s1, *q_a, pk.zc, m, acc + 2 * J1))) return rc;
    if (kS3) {
      // [s]a_acc and [r]b1_acc (two warps, 1.25 ms of dependent doublings) leave the chain: each starts on s3 as soon
      // as its MSM is through and runs under the next MSM of s1
      CUDA_TRY(ctx, cudaEventRecord(pk.ev_a, pk.s1));
      CUDA_TRY(ctx, cudaStreamWaitEvent(pk.s3, pk.ev_a, 0));
      CUDA_TRY(ctx, cp_on(pk.s3, vs + 0, pk.rs + 8, 8));
      CUDA_TRY(ctx, cp_on(pk.s3, vs + 8, pk.rs + 0, 8));
      o1->scalar_mul_var(pk.s3, acc + 2 * J1, vs, 1, var_out);
      if (early_ab) {   // A = alpha1 + a_acc + [r]delta1
        uint32_t* ptsA = acc + 28 * J1;
        CUDA_TRY(ctx, cudaStreamWaitEvent(pk.s3, pk.ev_fixed, 0));
        CUDA_TRY(ctx, cp_on(pk.s3, ptsA + 0 * J1, pk.consts_g1 + 0 * J1, J1));
        CUDA_TRY(ctx, cp_on(pk.s3, ptsA + 1 * J1, acc + 2 * J1, J1));
        CUDA_TRY(ctx, cp_on(pk.s3, ptsA + 2 * J1, fixed_out + 3 * J1, J1));
        o1->lincomb(pk.s3, ptsA, ones, 3, A_jac);
        o1->jacobian_to_affine(pk.s3, A_jac, outb, flags);
        ctx->launches += 2;
      }
    }
    pk.ws1.sort_from = pk.share_b_sort ? &pk.ws2 : nullptr;
    rc = ozl_rt_msm(ctx, pk.ws1, pk.s1, *q_b1, pk.zc, m, acc + 3 * J1);
    pk.ws1.sort_from = nullptr;
    if (rc) return rc;
    if (kS3) {
      CUDA_TRY(ctx, cudaEventRecord(pk.ev_b1, pk.s1));
      CUDA_TRY(ctx, cudaStreamWaitEvent(pk.s3, pk.ev_b1, 0));
      o1->scalar_mul_var(pk.s3, acc + 3 * J1, vs + 8, 1, var_out + J1);
      CUDA_TRY(ctx, cudaEventRecord(pk.ev_s3, pk.s3));
    } else {
      CUDA_TRY(ctx, cp_on(pk.s1, vs + 0, pk.rs + 8, 8));
      CUDA_TRY(ctx, cp_on(pk.s1, vs + 8, pk.rs + 0, 8));
      o1->scalar_mul_var(pk.s1, acc + 2 * J1, vs, 2, var_out);   // acc[2] = a_acc, acc[3] = b1_acc (adjacent)
    }
    if ((rc = ozl_rt_msm(ctx, pk.ws1, pk.s1, *q_l, pk.zc + (size_t)ni * 8, m - ni, acc + 1 * J1))) return rc;
    if (kS3) CUDA_TRY(ctx, cudaStreamWaitEvent(pk.s1, pk.ev_s3, 0));   // ev_s1 then covers s3 as well
    CUDA_TRY(ctx, cudaEventRecord(pk.ev_s1, pk.s1));
    CUDA_TRY(ctx, cudaEventRecord(pk.ev_s2, pk.s2));
    nvtxRangePop();
  }

  if (h_out) CUDA_TRY(ctx, cudaMemcpyAsync(h_out, pk.a, n * 32, cudaMemcpyDeviceToHost, st));

  {
    // MSM outputs (Jacobian): acc[0]=h, [1]=l, [2]=a, [3]=b1 in G1 slots; b2 in a G2 slot after them
    nvtxRangePushA("Compute C (h query MSM)");
    const bool ws_yield = ctx->ws.yield_ctas;
    ctx->ws.yield_ctas = pk.ws1.yield_ctas;
    rc = ozl_rt_msm(ctx, ctx->ws, st, *q_h, pk.hc, n - 1, acc + 0 * J1);
    ctx->ws.yield_ctas = ws_yield;
    nvtxRangePop();
    if (rc) return rc;
    CUDA_TRY(ctx, cudaStreamWaitEvent(st, pk.ev_s1, 0));
    CUDA_TRY(ctx, cudaStreamWaitEvent(st, pk.ev_s2, 0));

    STAGE(ctx, "g16_assemble");
    // C = l + h + [s]alpha1 + [r]beta1 + [rs]delta1 + [s]a_acc + [r]b1_acc      (= l + h + sA + rB1 - rs delta1)
    // A = alpha1 + a_acc + [r]delta1 ;  B = beta2 + b2_acc + [s]delta2
    // (the scalar multiples were issued on the side streams above)
    auto cp = [&](uint32_t* dst, const uint32_t* src, size_t words) { return cp_on(st, dst, src, words); };
    uint32_t* pts = acc + 20 * J1;          // up to 8 G1 points
    if (!early_ab) {
      CUDA_TRY(ctx, cp(pts + 0 * J1, pk.consts_g1 + 0 * J1, J1));   // alpha1
      CUDA_TRY(ctx, cp(pts + 1 * J1, acc + 2 * J1, J1));            // a_acc
      CUDA_TRY(ctx, cp(pts + 2 * J1, fixed_out + 3 * J1, J1));      // r delta1
      o1->lincomb(st, pts, ones, 3, A_jac);
    }
    CUDA_TRY(ctx, cp(pts + 0 * J1, acc + 1 * J1, J1));            // l_acc
    CUDA_TRY(ctx, cp(pts + 1 * J1, acc + 0 * J1, J1));            // h_acc
    CUDA_TRY(ctx, cp(pts + 2 * J1, fixed_out, 3 * J1));           // s alpha1, r beta1, rs delta1
    CUDA_TRY(ctx, cp(pts + 5 * J1, var_out, 2 * J1));             // s a_acc, r b1_acc
    o1->lincomb(st, pts, ones, 7, C_jac);
    if (!early_ab) {
      CUDA_TRY(ctx, cp(pts2 + 0 * J2, pk.consts_g2 + 0 * J2, J2));  // beta2
      CUDA_TRY(ctx, cp(pts2 + 1 * J2, acc_g2, J2));                 // b2_acc
      CUDA_TRY(ctx, cp(pts2 + 2 * J2, g2_fixed_out, J2));           // s delta2
      o2->lincomb(st, pts2, ones, 3, B_jac);
    }
    // affine outputs
    if (!early_ab) o1->jacobian_to_affine(st, A_jac, outb, flags);
    o1->jacobian_to_affine(st, C_jac, outb + 2 * c1, flags + 1);
    if (!early_ab) o2->jacobian_to_affine(st, B_jac, outb + 4 * c1, flags + 2);
    ctx->launches += early_ab ? 2 : 6;
    STAGE_END(ctx);
    CUDA_TRY(ctx, cudaMemcpyAsync(proof_a, outb, 2 * c1 * 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(ctx, cudaMemcpyAsync(proof_c, outb + 2 * c1, 2 * c1 * 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(ctx, cudaMemcpyAsync(proof_b, outb + 4 * c1, 2 * c2 * 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(ctx, cudaStreamSynchronize(st));
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { ctx->last_error = std::string("groth16_prove: ") + cudaGetErrorString(e); return OZL_ERR_CUDA; }
  }
  return OZL_OK;

}

}  // extern "C"
