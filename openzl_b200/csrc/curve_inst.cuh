// Defines the OzlCurveOps table for one curve: include after defining OZL_F (coordinate field),
// OZL_C (curve constants), OZL_BASE (base prime field) and OZL_OPS (table symbol).
#include "runtime.cuh"

namespace {
typedef OZL_F F_;
typedef OZL_C C_;
int msm_entry(ozl_ctx* ctx, ozl_rt::MsmWorkspace& ws, cudaStream_t st, const ozl_rt::Bases& b, const uint32_t* d_scalars,
              size_t n, uint32_t* d_out) {
  return ozl_rt::msm_run<F_>(ctx, ws, st, b, d_scalars, n, d_out);
}
int msm_batched_entry(ozl_ctx* ctx, ozl_rt::MsmWorkspace& ws, cudaStream_t st, const ozl_rt::Bases& b, const uint32_t* d_scalars,
                      size_t n, const ozl_rt::MsmBatches& mb, uint32_t* d_out) {
  return ozl_rt::msm_run_batched<F_>(ctx, ws, st, b, d_scalars, n, mb, d_out);
}
void generate_entry(cudaStream_t st, uint64_t start, uint32_t n, uint32_t* d_pts) {
  const uint32_t threads = (n + GEN_RUN - 1) / GEN_RUN;
  k_generate_bases<F_, C_><<<(threads + 127) / 128, 128, 0, st>>>(start, n, d_pts);
}
void jsum_entry(cudaStream_t st, const uint32_t* d_pts, uint32_t k, uint32_t* d_out) {
  k_jacobian_sum<F_><<<1, 32, 0, st>>>(d_pts, k, d_out);
}
void jaff_entry(cudaStream_t st, const uint32_t* d_jac, uint32_t* d_out, int* d_flag) {
  k_jacobian_to_affine<F_><<<1, 32, 0, st>>>(d_jac, d_out, d_flag);
}
void bench_entry(cudaStream_t st, int blocks, int threads, uint32_t* d_out, int iters) {
  k_bench_mul<Fp<OZL_BASE>><<<blocks, threads, 0, st>>>(d_out, iters);
}
void fixed_base_entry(cudaStream_t st, const uint32_t* d_scalars, uint32_t n, uint32_t* d_out, uint8_t* d_flags) {
  k_fixed_base_mul<F_, C_><<<(n + 127) / 128, 128, 0, st>>>(d_scalars, n, d_out, d_flags);
}
void lincomb_entry(cudaStream_t st, const uint32_t* d_pts, const uint32_t* d_scalars, uint32_t k, uint32_t* d_out) {
  k_lincomb<F_><<<1, 32, 0, st>>>(d_pts, d_scalars, k, d_out);
}
void smul_var_entry(cudaStream_t st, const uint32_t* d_pts, const uint32_t* d_scalars, uint32_t k, uint32_t* d_out) {
  if (k) k_scalar_mul_warp<F_, false><<<k, 32, 0, st>>>(d_pts, d_scalars, d_out);
}
void smul_table_entry(cudaStream_t st, const uint32_t* d_tables, const uint32_t* d_scalars, uint32_t k, uint32_t* d_out) {
  if (k) k_scalar_mul_warp<F_, true><<<k, 32, 0, st>>>(d_tables, d_scalars, d_out);
}
void build_table_entry(cudaStream_t st, const uint32_t* d_pt, uint32_t* d_table) {
  k_build_byte_table<F_><<<1, 32, 0, st>>>(d_pt, d_table);
}
void precompute_entry(cudaStream_t st, const uint32_t* d_src, uint32_t* d_dst, uint32_t n, int shift) {
  const uint32_t threads = (n + GEN_RUN - 1) / GEN_RUN;
  k_precompute<F_><<<(threads + 127) / 128, 128, 0, st>>>(d_src, d_dst, n, shift);
}
void a2j_entry(cudaStream_t st, const uint32_t* d_aff, uint32_t* d_out) {
  k_affine_to_jacobian<F_><<<1, 32, 0, st>>>(d_aff, d_out);
}
#ifdef OZL_FP64_BENCH
void bench_fp64_entry(cudaStream_t st, int blocks, int threads, uint32_t* d_out, int iters, int mix) {
  k_bench_mul_fp64<OZL_BASE><<<blocks, threads, 0, st>>>(d_out, iters, mix);
}
#define OZL_FP64_BENCH_PTR bench_fp64_entry
#else
#define OZL_FP64_BENCH_PTR nullptr
#endif
}  // namespace

const OzlCurveOps OZL_OPS = {msm_entry, msm_batched_entry, generate_entry, jsum_entry, jaff_entry, bench_entry, fixed_base_entry, lincomb_entry, smul_var_entry, smul_table_entry, build_table_entry, precompute_entry, a2j_entry, OZL_FP64_BENCH_PTR};
