/* libozl_b200 -- C ABI of the Blackwell-native proving backend for OpenZL.
 *
 * This header is the drop-in boundary for the ONE hot path this library accelerates: the two
 * compute kernels underneath `Groth16::<E>::prove`
 * (/root/reference/plugins/arkworks/src/groth16.rs:445-457, trait at
 * /root/reference/openzl-crypto/src/constraint.rs:73-79):
 *
 *   ark_ec::msm::VariableBaseMSM::multi_scalar_mul(bases, scalars) -> G::Projective
 *       reached through `pub use ec`   (/root/reference/plugins/arkworks/src/lib.rs:28-29)
 *   ark_poly::EvaluationDomain::{fft,ifft,coset_fft,coset_ifft}_in_place
 *       reached through `pub use poly` (/root/reference/plugins/arkworks/src/lib.rs:70-71)
 *
 * The reference has no FFI of its own (it is pure Rust); these entry points are what a Rust
 * shim binds with `extern "C"`.  Conventions follow what arkworks 0.3.0 keeps
 * in memory so that slices can be passed without conversion:
 *
 *   field elements / point coordinates : little-endian u64 limbs, MONTGOMERY form (aR mod p)
 *   MSM scalars                         : little-endian 4 x u64, CANONICAL (`into_repr()`)
 *   affine point                        : x || y            (G2: x.c0 || x.c1 || y.c0 || y.c1)
 *   projective result                   : Jacobian X || Y || Z (x = X/Z^2, y = Y/Z^3),
 *                                         identity = (0, 1, 0) like GroupProjective::zero()
 *
 * Error behaviour mirrors the plugin's opaque `Error` (groth16.rs:35-45): every call returns
 * 0 on success or a non-zero ozl_status; nothing aborts or throws across the ABI.  All host
 * buffers are caller-owned; device memory behind a bases handle is library-owned.  A context is
 * thread-compatible (use one per thread or lock externally); calls are synchronous on return
 * unless the name ends in `_async`.
 *
 * There is NO CPU fallback: every compute entry point runs hand-written sm_100a kernels and
 * fails with OZL_ERR_CUDA / OZL_ERR_NO_DEVICE when no B200-class GPU is usable.
 */
#ifndef OZL_H
#define OZL_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ozl_ctx ozl_ctx;

typedef enum {
  OZL_OK = 0,
  OZL_ERR_ARG = 1,        /* bad argument (null pointer, unknown curve/field, size out of range) */
  OZL_ERR_CUDA = 2,       /* a CUDA runtime call or kernel failed; see ozl_last_error()        */
  OZL_ERR_NO_DEVICE = 3,  /* no usable CUDA device                                             */
  OZL_ERR_OOM = 4,        /* device or host allocation failed                                  */
  OZL_ERR_HANDLE = 5,     /* unknown or freed bases handle                                     */
  OZL_ERR_DOMAIN = 6      /* log_n exceeds the field's two-adicity (ark: `new` returns None)   */
} ozl_status;

/* Groups behind `Pairing::{G1, G2}` (/root/reference/plugins/arkworks/src/pairing.rs:14-23). */
typedef enum {
  OZL_BLS12_381_G1 = 0,
  OZL_BLS12_381_G2 = 1,
  OZL_BN254_G1 = 2,
  OZL_BN254_G2 = 3
} ozl_curve;

/* Scalar fields (`Pairing::Scalar`, pairing.rs:11). */
typedef enum {
  OZL_BN254_FR = 0,
  OZL_BLS12_381_FR = 1
} ozl_field;

/* ---- context ------------------------------------------------------------------------------ */
int ozl_version(void);
const char* ozl_strerror(int status);
/* Text of the most recent CUDA error seen by this context (empty string if none). */
const char* ozl_last_error(const ozl_ctx* ctx);

/* One context per (device, stream).  Creates its own non-blocking stream. */
int ozl_ctx_create(int device, ozl_ctx** out);
void ozl_ctx_destroy(ozl_ctx* ctx);
/* Run on a caller-supplied cudaStream_t (passed as void*), e.g. torch's current stream, so the
 * caller can bracket calls with its own CUDA events.  NULL restores the context's own stream. */
int ozl_ctx_set_stream(ozl_ctx* ctx, void* cuda_stream);
void* ozl_ctx_get_stream(ozl_ctx* ctx);
int ozl_ctx_synchronize(ozl_ctx* ctx);
/* Bytes of u64 limbs per coordinate / per affine point / per Jacobian point of a curve. */
int ozl_curve_coord_limbs(int curve);

/* ---- MSM: replaces VariableBaseMSM::multi_scalar_mul -------------------------------------- */
/* Upload `n` packed affine bases once (the MSM bases of a Groth16 proving key are constant
 * across proofs: `ProvingContext<E>(pub ProvingKey<E>)`, groth16.rs:127-129).  `inf_mask` is an
 * optional bitset (bit i of byte i/8) marking points at infinity (GroupAffine::infinity). */
int ozl_msm_bases_upload(ozl_ctx* ctx, int curve, const uint64_t* bases, const uint8_t* inf_mask,
                         size_t n, uint32_t* handle);
/* Same, from a device pointer (copied device-to-device). */
int ozl_msm_bases_upload_device(ozl_ctx* ctx, int curve, const uint64_t* d_bases,
                                const uint8_t* d_inf_mask, size_t n, uint32_t* handle);
/* Synthesize bases P_i = [start + i]G (G = the curve's generator) directly in device memory;
 * used by benchmarks and size-independent parity checks (sum s_i P_i = [sum s_i (start+i)]G). */
int ozl_msm_bases_generate(ozl_ctx* ctx, int curve, uint64_t start, size_t n, uint32_t* handle);
/* Copy bases [first, first + n) of a handle back to the host (packed affine, Montgomery). */
int ozl_msm_bases_download(ozl_ctx* ctx, uint32_t handle, size_t first, size_t n, uint64_t* out);
int ozl_msm_bases_free(ozl_ctx* ctx, uint32_t handle);

/* sum_{i<n} scalars[i] * bases[i] over the first n bases of `handle` (n <= uploaded count,
 * like ark's `size = min(bases.len(), scalars.len())`).  Host scalars in, host Jacobian out. */
int ozl_msm(ozl_ctx* ctx, uint32_t handle, const uint64_t* scalars, size_t n,
            uint64_t* out_jacobian);
/* Device scalars in, device Jacobian out; enqueued on the context's stream, not synchronized. */
int ozl_msm_device_async(ozl_ctx* ctx, uint32_t handle, const uint64_t* d_scalars, size_t n,
                         uint64_t* d_out_jacobian);
/* Pippenger window width in bits; 0 = choose from n (default). */
int ozl_msm_set_window_bits(ozl_ctx* ctx, int c);
int ozl_msm_get_window_bits(ozl_ctx* ctx, int curve, size_t n);

/* Sum k Jacobian points (host, X||Y||Z each) -- the combine step after the multi-GPU exchange
 * of per-rank partial sums (EC addition is not an NCCL reduction op). */
int ozl_jacobian_sum(ozl_ctx* ctx, int curve, const uint64_t* points, size_t k,
                     uint64_t* out_jacobian);
/* GroupProjective::into_affine: writes x || y (zeros for the identity) and *is_identity. */
int ozl_jacobian_to_affine(ozl_ctx* ctx, int curve, const uint64_t* jacobian, uint64_t* out_affine,
                           int* is_identity);

/* ---- NTT: replaces Radix2EvaluationDomain::{fft,ifft,coset_fft,coset_ifft}_in_place ------- */
/* data = 2^log_n elements x 4 u64 limbs, Montgomery, natural order in and out, in place.
 * inverse: multiply by size_inv as ark does; coset: generator g = F::multiplicative_generator()
 * (forward: scale by g^i first; inverse: scale by g^-i last). */
int ozl_ntt(ozl_ctx* ctx, int field, uint64_t* data, uint32_t log_n, int inverse, int coset);
int ozl_ntt_device_async(ozl_ctx* ctx, int field, uint64_t* d_data, uint32_t log_n, int inverse,
                         int coset);

/* ---- instrumentation ---------------------------------------------------------------------- */
/* When enabled, every pipeline stage of the next calls is bracketed with CUDA events on the
 * context's stream.  ozl_ctx_get_stage_times copies up to `cap` (name, ms, launches) records of
 * the most recent call and returns the count. */
typedef struct {
  char name[32];
  float ms;
  int launches;
} ozl_stage_time;
int ozl_ctx_enable_timing(ozl_ctx* ctx, int on);
int ozl_ctx_get_stage_times(ozl_ctx* ctx, ozl_stage_time* out, int cap);
/* Total kernel launches issued by this context since creation. */
uint64_t ozl_ctx_launch_count(const ozl_ctx* ctx);

/* Micro-benchmark of the field multiplier that bounds every kernel here: runs `iters` dependent
 * Montgomery multiplications per thread on a full grid and returns multiplications per second.
 * field_id: 0 = BLS12-381 Fq (12 limbs), 1 = BN254 Fq (8 limbs). */
int ozl_bench_field_mul(ozl_ctx* ctx, int field_id, int iters, double* mul_per_sec);

#ifdef __cplusplus
}
#endif
#endif /* OZL_H */
