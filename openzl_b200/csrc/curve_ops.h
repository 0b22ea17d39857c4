// Per-curve / per-field entry points.  Each curve (and each scalar field) is instantiated in its
// own translation unit so the library builds in parallel; capi.cu dispatches through these tables.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

struct ozl_ctx;
namespace ozl_rt {
struct Bases;
struct MsmWorkspace;
struct MsmBatches;
}
namespace ozl {
struct NttWorkspace;
}

struct OzlCurveOps {
  int (*msm)(ozl_ctx* ctx, ozl_rt::MsmWorkspace& ws, cudaStream_t st, const ozl_rt::Bases& b, const uint32_t* d_scalars,
             size_t n, uint32_t* d_out);
  // same MSM with the scalars arriving in point-range batches (host->device copy overlapped with the accumulation)
  int (*msm_batched)(ozl_ctx* ctx, ozl_rt::MsmWorkspace& ws, cudaStream_t st, const ozl_rt::Bases& b, const uint32_t* d_scalars,
                     size_t n, const ozl_rt::MsmBatches& mb, uint32_t* d_out);
  void (*generate)(cudaStream_t st, uint64_t start, uint32_t n, uint32_t* d_pts);
  void (*jacobian_sum)(cudaStream_t st, const uint32_t* d_pts, uint32_t k, uint32_t* d_out);
  void (*jacobian_to_affine)(cudaStream_t st, const uint32_t* d_jac, uint32_t* d_out, int* d_flag);
  void (*bench_mul)(cudaStream_t st, int blocks, int threads, uint32_t* d_out, int iters);  // base-field multiplier
  // out_affine[j] = [k_j] G (canonical 256-bit k_j); flags[j] = 1 when the result is the identity
  void (*fixed_base_mul)(cudaStream_t st, const uint32_t* d_scalars, uint32_t n, uint32_t* d_out_affine, uint8_t* d_flags);
  // out_jac = sum_{i<k} scalars[i] * pts_jac[i]   (k <= 32; one lane per term)
  void (*lincomb)(cudaStream_t st, const uint32_t* d_pts_jac, const uint32_t* d_scalars, uint32_t k, uint32_t* d_out_jac);
  // out[b] = scalars[b] * P_b for k Jacobian points, one warp each (variable points)
  void (*scalar_mul_var)(cudaStream_t st, const uint32_t* d_pts_jac, const uint32_t* d_scalars, uint32_t k, uint32_t* d_out_jac);
  // same with per-point byte tables (32 Jacobian entries each) built by build_byte_table
  void (*scalar_mul_table)(cudaStream_t st, const uint32_t* d_tables, const uint32_t* d_scalars, uint32_t k, uint32_t* d_out_jac);
  void (*build_byte_table)(cudaStream_t st, const uint32_t* d_pt_jac, uint32_t* d_table);
  // dst_i = 2^shift * src_i for n affine points (base precomputation)
  void (*precompute)(cudaStream_t st, const uint32_t* d_src, uint32_t* d_dst, uint32_t n, int shift);
  // Jacobian <- affine (x||y) on the device, for constants uploaded from the host
  void (*affine_to_jacobian)(cudaStream_t st, const uint32_t* d_aff, uint32_t* d_out_jac);
  // base-field multiplier on the FP64 pipe (mix = 0) or two integer + two FP64 chains per thread (mix = 1); NULL where
  // the field has no 48-bit constants
  void (*bench_mul_fp64)(cudaStream_t st, int blocks, int threads, uint32_t* d_out, int iters, int mix);
};

extern const OzlCurveOps ozl_ops_bls12_381_g1;
extern const OzlCurveOps ozl_ops_bls12_381_g2;
extern const OzlCurveOps ozl_ops_bn254_g1;
extern const OzlCurveOps ozl_ops_bn254_g2;

struct OzlFieldOps {
  int (*ntt)(cudaStream_t st, ozl::NttWorkspace& ws, uint32_t* d_data, uint32_t log_n, bool inverse, bool coset, int* launches);
  void (*spmv)(cudaStream_t st, const uint32_t* row_ptr, const uint32_t* col, const uint32_t* cidx, const uint32_t* coef,
               const uint32_t* x, uint32_t n_rows, uint32_t* y);
  void (*from_mont)(cudaStream_t st, const uint32_t* in, uint32_t* out, uint32_t n);
  void (*h_pointwise)(cudaStream_t st, uint32_t* a, const uint32_t* b, const uint32_t* c, const uint32_t* scale, uint32_t n);
  void (*vanishing_inv)(cudaStream_t st, int log_n, uint32_t* out);
  // out = a * b mod r, all canonical 256-bit integers on the device
  void (*mul_canonical)(cudaStream_t st, const uint32_t* a, const uint32_t* b, uint32_t* out);
  // Poseidon permutation of `batch` states of `width` elements in place (Montgomery), x^5 S-box
  void (*poseidon)(cudaStream_t st, uint32_t* states, uint32_t batch, int width, int full_rounds, int partial_rounds,
                   const uint32_t* round_keys, const uint32_t* mds);
};
extern const OzlFieldOps ozl_fops_bn254_fr;
extern const OzlFieldOps ozl_fops_bls12_381_fr;

int ozl_ntt_run_bn254_fr(cudaStream_t st, ozl::NttWorkspace& ws, uint32_t* d_data, uint32_t log_n, bool inverse, bool coset, int* launches);
int ozl_ntt_run_bls12_381_fr(cudaStream_t st, ozl::NttWorkspace& ws, uint32_t* d_data, uint32_t log_n, bool inverse, bool coset, int* launches);
