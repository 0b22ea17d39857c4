#include "runtime.cuh"
int ozl_ntt_run_bls12_381_fr(cudaStream_t st, ozl::NttWorkspace& ws, uint32_t* d, uint32_t log_n, bool inverse, bool coset, int* launches) {
  return ozl::ntt_run<ozl_params::Bls12381Fr>(st, ws, OZL_BLS12_381_FR, d, log_n, inverse, coset, launches);
}
