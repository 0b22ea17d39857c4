// Per-curve / per-field entry points.  Each curve (and each scalar field) is instantiated in its
// own translation unit so the library builds in parallel; capi.cu dispatches through these tables.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

struct ozl_ctx;
namespace ozl_rt {
struct Bases;
}
namespace ozl {
struct NttWorkspace;
}

struct OzlCurveOps {
  int (*msm)(ozl_ctx* ctx, const ozl_rt::Bases& b, const uint32_t* d_scalars, size_t n, uint32_t* d_out);
  void (*generate)(cudaStream_t st, uint64_t start, uint32_t n, uint32_t* d_pts);
  void (*jacobian_sum)(cudaStream_t st, const uint32_t* d_pts, uint32_t k, uint32_t* d_out);
  void (*jacobian_to_affine)(cudaStream_t st, const uint32_t* d_jac, uint32_t* d_out, int* d_flag);
  void (*bench_mul)(cudaStream_t st, int blocks, int threads, uint32_t* d_out, int iters);  // base-field multiplier
};

extern const OzlCurveOps ozl_ops_bls12_381_g1;
extern const OzlCurveOps ozl_ops_bls12_381_g2;
extern const OzlCurveOps ozl_ops_bn254_g1;
extern const OzlCurveOps ozl_ops_bn254_g2;

int ozl_ntt_run_bn254_fr(cudaStream_t st, ozl::NttWorkspace& ws, uint32_t* d_data, uint32_t log_n, bool inverse, bool coset, int* launches);
int ozl_ntt_run_bls12_381_fr(cudaStream_t st, ozl::NttWorkspace& ws, uint32_t* d_data, uint32_t log_n, bool inverse, bool coset, int* launches);
