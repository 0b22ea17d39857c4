#include "runtime.cuh"
#include "ntt.cuh"
#include "frops.cuh"
namespace {
typedef ozl_params::Bls12381Fr P_;
void spmv_entry(cudaStream_t st, const uint32_t* row_ptr, const uint32_t* col, const uint32_t* cidx, const uint32_t* coef,
                const uint32_t* x, uint32_t n_rows, uint32_t* y) {
  if (n_rows) ozl::k_spmv<P_><<<(n_rows + 255) / 256, 256, 0, st>>>(row_ptr, col, cidx, coef, x, n_rows, y);
}
void from_mont_entry(cudaStream_t st, const uint32_t* in, uint32_t* out, uint32_t n) {
  if (n) ozl::k_from_mont<P_><<<(n + 255) / 256, 256, 0, st>>>(in, out, n);
}
void h_pointwise_entry(cudaStream_t st, uint32_t* a, const uint32_t* b, const uint32_t* c, const uint32_t* scale, uint32_t n) {
  if (n) ozl::k_h_pointwise<P_><<<(n + 255) / 256, 256, 0, st>>>(a, b, c, (const ozl::Fp<P_>*)scale, n);
}
void vanishing_inv_entry(cudaStream_t st, int log_n, uint32_t* out) {
  ozl::k_vanishing_inv<P_><<<1, 32, 0, st>>>(log_n, (ozl::Fp<P_>*)out);
}
void mul_canonical_entry(cudaStream_t st, const uint32_t* a, const uint32_t* b, uint32_t* out) {
  ozl::k_mul_canonical<P_><<<1, 32, 0, st>>>(a, b, out);
}
void poseidon_entry(cudaStream_t st, uint32_t* states, uint32_t batch, int width, int full_rounds, int partial_rounds,
                    const uint32_t* round_keys, const uint32_t* mds) {
  if (batch) ozl::k_poseidon_permute<P_><<<(batch + 127) / 128, 128, 0, st>>>(states, batch, width, full_rounds, partial_rounds, round_keys, mds);
}
}  // namespace
int ozl_ntt_run_bls12_381_fr(cudaStream_t st, ozl::NttWorkspace& ws, uint32_t* d, uint32_t log_n, bool inverse, bool coset, int* launches) {
  return ozl::ntt_run<P_>(st, ws, OZL_BLS12_381_FR, d, log_n, inverse, coset, launches);
}
const OzlFieldOps ozl_fops_bls12_381_fr = {ozl_ntt_run_bls12_381_fr, spmv_entry, from_mont_entry, h_pointwise_entry, vanishing_inv_entry, mul_canonical_entry, poseidon_entry};
