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
  OZL_ERR_DOMAIN = 6,     /* log_n exceeds the field's two-adicity (ark: `new` returns None)   */
  OZL_ERR_NCCL = 7        /* libnccl.so.2 not loadable, or an NCCL call failed                 */
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

/* One context per (device, stream).  Creates its own non-blocking stream.  OZL_ERR_NO_DEVICE unless
 * `device` is a compute-capability-10.x GPU (the library contains sm_100a code only). */
int ozl_ctx_create(int device, ozl_ctx** out);
void ozl_ctx_destroy(ozl_ctx* ctx);
/* Run on a caller-supplied cudaStream_t (passed as void*), e.g. torch's current stream, so the
 * caller can bracket calls with its own CUDA events.  NULL is CUDA's legacy default stream;
 * ozl_ctx_use_own_stream goes back to the context's private stream. */
int ozl_ctx_set_stream(ozl_ctx* ctx, void* cuda_stream);
int ozl_ctx_use_own_stream(ozl_ctx* ctx);
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
/* Trade HBM for work: store `factor` copies 2^(c*Wc*q) * P_i of the bases (q < factor,
 * Wc = ceil(W / factor) windows per copy) so that only Wc bucket sets are reduced and the final
 * Horner shrinks by the same factor.  One-time cost at key-load time; the window width c is
 * fixed from the handle's size at this point.  Results are unchanged (same group element). */
int ozl_msm_bases_precompute(ozl_ctx* ctx, uint32_t handle, int factor);
/* Copy bases [first, first + n) of a handle back to the host (packed affine, Montgomery). */
int ozl_msm_bases_download(ozl_ctx* ctx, uint32_t handle, size_t first, size_t n, uint64_t* out);
int ozl_msm_bases_free(ozl_ctx* ctx, uint32_t handle);

/* sum_{i<n} scalars[i] * bases[i] over the first n bases of `handle` (n <= uploaded count,
 * like ark's `size = min(bases.len(), scalars.len())`).  Host scalars in, host Jacobian out.
 * `scalars` may be pageable or page-locked: they cross PCIe in point-range batches while earlier
 * batches are already being accumulated, so the transfer is hidden under the computation.
 *
 * Scalars must be CANONICAL (< r, what `into_repr()` yields).  A scalar with bits the window plan does
 * not cover (>= 2^255 for BN254, or a recoding carry out of the top window) makes the synchronous
 * calls (ozl_msm, ozl_msm_sharded) return OZL_ERR_ARG; the *_async / submit forms do not report it.
 * An all-zero affine base (0, 0) is treated as the point at infinity, with or without `inf_mask`.
 *
 * Size limits (32-bit sort entries): n * windows < 2^32 and n * copies < 2^31, where windows =
 * ceil((scalar bits + 1) / c) and copies is the precompute factor in effect -- e.g. 2^28 points take
 * at most 7 copies, and 2^29 points with c = 20 (13 windows) are refused with OZL_ERR_ARG. */
int ozl_msm(ozl_ctx* ctx, uint32_t handle, const uint64_t* scalars, size_t n,
            uint64_t* out_jacobian);
/* Pipelined form of ozl_msm for back-to-back MSMs (a prover streaming proofs): returns as soon as
 * the work is enqueued.  Scalars are copied on a separate copy stream into one of two staging
 * buffers, so the host->device transfer of call i+1 overlaps the kernels of call i.  `scalars`
 * and `out_jacobian` must stay valid (and should be page-locked) until ozl_ctx_synchronize. */
int ozl_msm_submit(ozl_ctx* ctx, uint32_t handle, const uint64_t* scalars, size_t n,
                   uint64_t* out_jacobian);
/* Device scalars in, device Jacobian out; enqueued on the context's stream, not synchronized. */
int ozl_msm_device_async(ozl_ctx* ctx, uint32_t handle, const uint64_t* d_scalars, size_t n,
                         uint64_t* d_out_jacobian);
/* Pippenger window width in bits; 0 = choose from n (default). */
int ozl_msm_set_window_bits(ozl_ctx* ctx, int c);
int ozl_msm_get_window_bits(ozl_ctx* ctx, int curve, size_t n);
/* EXPERIMENTAL: run the first `levels` halving levels of bucket accumulation with batched-affine
 * additions (one shared inversion per thread, msm_batch.cuh) before the XYZZ accumulation.
 * 0 = off, -1 = library default (off; measured no faster than XYZZ on B200).
 * The result is the same group element either way. */
int ozl_msm_set_batch_affine(ozl_ctx* ctx, int levels);
/* The plan an n-scalar MSM on this handle will use: window bits, windows, bucket sets, copies. */
int ozl_msm_bases_info(ozl_ctx* ctx, uint32_t handle, size_t n, int* c, int* windows, int* bucket_sets,
                       int* factor);

/* Sum k Jacobian points (host, X||Y||Z each) -- the combine step after the multi-GPU exchange
 * of per-rank partial sums (EC addition is not an NCCL reduction op). */
int ozl_jacobian_sum(ozl_ctx* ctx, int curve, const uint64_t* points, size_t k,
                     uint64_t* out_jacobian);
/* GroupProjective::into_affine: writes x || y (zeros for the identity) and *is_identity. */
int ozl_jacobian_to_affine(ozl_ctx* ctx, int curve, const uint64_t* jacobian, uint64_t* out_affine,
                           int* is_identity);

/* ---- multi-GPU MSM: one rank per GPU, point-range shards, one all-gather of partials -------- */
/* Sum_i s_i P_i = Sum_g Sum_{i in shard g} s_i P_i.  Every rank uploads ITS point range as a bases
 * handle, runs the full single-GPU pipeline on its scalars, and the Jacobian partials are exchanged
 * with one ncclAllGather on the context's stream and summed on the device (elliptic-curve addition
 * is not an NCCL reduction op, so it is not an all-reduce).  The communicator is created from a
 * 128-byte NCCL unique id that rank 0 makes and the caller distributes by its own means (MPI,
 * torch.distributed, a file ...).  NCCL is bound at run time (dlopen of libnccl.so.2). */
typedef struct ozl_comm ozl_comm;
#define OZL_COMM_ID_BYTES 128
int ozl_comm_unique_id(uint8_t* id_out /* OZL_COMM_ID_BYTES */);
int ozl_comm_create(ozl_ctx* ctx, const uint8_t* id, int rank, int world, ozl_comm** out);
int ozl_comm_destroy(ozl_comm* comm);
/* Host scalars of this rank's shard in, the COMBINED result (identical on every rank) out. */
int ozl_msm_sharded(ozl_ctx* ctx, ozl_comm* comm, uint32_t handle, const uint64_t* scalars, size_t n,
                    uint64_t* out_jacobian);
/* Same with device scalars / device output, enqueued on the context's stream. */
int ozl_msm_sharded_device_async(ozl_ctx* ctx, ozl_comm* comm, uint32_t handle, const uint64_t* d_scalars,
                                 size_t n, uint64_t* d_out_jacobian);
/* The combine step alone: all-gather one device-resident Jacobian partial per rank and sum. */
int ozl_comm_allgather_sum_async(ozl_ctx* ctx, ozl_comm* comm, int curve, const uint64_t* d_partial,
                                 uint64_t* d_out_jacobian);

/* ---- NTT: replaces Radix2EvaluationDomain::{fft,ifft,coset_fft,coset_ifft}_in_place ------- */
/* data = 2^log_n elements x 4 u64 limbs, Montgomery, natural order in and out, in place.
 * inverse: multiply by size_inv as ark does; coset: generator g = F::multiplicative_generator()
 * (forward: scale by g^i first; inverse: scale by g^-i last). */
int ozl_ntt(ozl_ctx* ctx, int field, uint64_t* data, uint32_t log_n, int inverse, int coset);
int ozl_ntt_device_async(ozl_ctx* ctx, int field, uint64_t* d_data, uint32_t log_n, int inverse,
                         int coset);

/* ---- Groth16 prover: replaces ark_groth16::create_proof behind groth16.rs:454 -------------- */
/* Pairing-friendly curve families (`Pairing`, /root/reference/plugins/arkworks/src/pairing.rs:9-38). */
typedef enum {
  OZL_PAIRING_BN254 = 0,
  OZL_PAIRING_BLS12_381 = 1
} ozl_pairing;

/* One R1CS matrix in CSR form.  Coefficients are indices into a shared table of field elements
 * (Montgomery) because gadget-built systems have very few distinct constants. */
typedef struct {
  uint32_t n_rows;
  const uint32_t* row_ptr;  /* n_rows + 1 */
  const uint32_t* col_idx;  /* row_ptr[n_rows] */
  const uint32_t* coef_idx; /* row_ptr[n_rows] */
} ozl_csr;

/* y = M x over the pairing's scalar field; x has n_cols elements, y gets M->n_rows (host buffers,
 * Montgomery).  Used by the setup (transposed matrices) and by tests of the witness map. */
int ozl_fr_spmv(ozl_ctx* ctx, int field, const ozl_csr* M, const uint64_t* coef_table, uint32_t n_coef,
                const uint64_t* x, uint32_t n_cols, uint64_t* y);

/* Poseidon permutation over the pairing's scalar field, `batch` states of `width` elements each, in
 * place (host buffers, Montgomery): the hash of the reference's Groth16 workload circuit
 * (/root/reference/openzl-crypto/src/poseidon/mod.rs:156-283; S-box x^5,
 * /root/reference/plugins/arkworks/src/poseidon/mod.rs:147-159).  round_keys = (full_rounds +
 * partial_rounds) x width elements, round-major; mds = width x width, row-major.  Evaluates hash-chain
 * witnesses on the device; with the reference's width-3 known-answer vector
 * (/root/reference/openzl-tutorials/src/poseidon.rs:388-401) it pins the device multiplier directly.
 * width in 2..12, full_rounds even. */
int ozl_fr_poseidon_permute(ozl_ctx* ctx, int field, uint64_t* states, size_t batch, uint32_t width,
                            uint32_t full_rounds, uint32_t partial_rounds, const uint64_t* round_keys,
                            const uint64_t* mds);

/* out_affine[j] = [scalars[j]] G for the curve's generator G; identity_flags[j] = 1 for the point
 * at infinity.  The fixed-base work of a (known-trapdoor) Groth16 setup. */
int ozl_fixed_base_mul(ozl_ctx* ctx, int curve, const uint64_t* scalars, size_t n, uint64_t* out_affine,
                       uint8_t* identity_flags);

/* Proving key resident on the device: `ProvingKey{vk.alpha_g1, beta_g1, beta_g2, delta_g1, delta_g2,
 * a_query, b_g1_query, b_g2_query, h_query, l_query}` (field names as destructured at
 * groth16.rs:200-205) plus the circuit's A, B, C matrices.  The five query vectors are bases
 * handles uploaded beforehand (sizes: a/b_g1/b_g2 = n_vars, h = domain - 1, l = n_vars - n_instance);
 * the pk takes ownership of them.  n_instance counts the leading constant 1. */
int ozl_groth16_pk_create(ozl_ctx* ctx, int pairing, uint32_t n_constraints, uint32_t n_instance, uint32_t n_vars,
                          const ozl_csr* A, const ozl_csr* B, const ozl_csr* C, const uint64_t* coef_table,
                          uint32_t n_coef, uint32_t a_query, uint32_t b_g1_query, uint32_t b_g2_query,
                          uint32_t h_query, uint32_t l_query, const uint64_t* alpha_g1, const uint64_t* beta_g1,
                          const uint64_t* delta_g1, const uint64_t* beta_g2, const uint64_t* delta_g2,
                          uint32_t* pk_handle);
int ozl_groth16_pk_destroy(ozl_ctx* ctx, uint32_t pk_handle);

/* `create_proof(circuit, pk, r, s)`: z = full assignment (1, instance..., witness...) as n_vars
 * Montgomery elements; r, s = the two blinding scalars (canonical, drawn by the caller's rng like
 * `create_random_proof`).  Outputs affine A (G1), B (G2), C (G1).  h_out, if non-NULL, receives the
 * domain_size quotient coefficients (Montgomery) for inspection. */
int ozl_groth16_prove(ozl_ctx* ctx, uint32_t pk_handle, const uint64_t* z, const uint64_t* r, const uint64_t* s,
                      uint64_t* proof_a, uint64_t* proof_b, uint64_t* proof_c, uint64_t* h_out);
/* Domain size 2^k = next_pow2(n_constraints + n_instance) of a pk (ark's witness_map sizing). */
int ozl_groth16_domain_size(ozl_ctx* ctx, uint32_t pk_handle, uint32_t* out);

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
/* Same records as a timeline: start and end of every stage in milliseconds after the first stage's start (stages
 * of one call run on several streams and overlap; this is the Gantt chart of the call). */
typedef struct {
  char name[32];
  float start_ms;
  float end_ms;
  int launches;
} ozl_stage_span;
int ozl_ctx_get_stage_spans(ozl_ctx* ctx, ozl_stage_span* out, int cap);
/* Total kernel launches issued by this context since creation. */
uint64_t ozl_ctx_launch_count(const ozl_ctx* ctx);

/* Micro-benchmark of the field multiplier that bounds every kernel here: runs `iters` dependent
 * Montgomery multiplications per thread on a full grid and returns multiplications per second.
 * field_id: 0 = BLS12-381 Fq (12 limbs), 1 = BN254 Fq (8 limbs); 2 = BLS12-381 Fq on the FP64 pipe (8 x 48-bit limbs in
 * doubles, fp64mul.cuh), 3 = two integer and two FP64 product chains per thread (the hybrid's ceiling). */
int ozl_bench_field_mul(ozl_ctx* ctx, int field_id, int iters, double* mul_per_sec);

#ifdef __cplusplus
}
#endif
#endif /* OZL_H */
