/* Plain-C replay of the call sequence the Rust shim (plugins/b200/src/groth16.rs) makes for
 * `ProofSystem::compile` + `ProofSystem::prove`, with every host buffer in PAGEABLE memory (malloc),
 * the way a Rust `Vec` is -- no Python, no pinned memory, no torch between the caller and
 * libozl_b200.so:
 *
 *   compile:  ozl_ctx_create -> 5 x (ozl_msm_bases_upload + ozl_msm_bases_precompute)
 *             -> ozl_groth16_pk_create (CSR matrices + alpha/beta/delta)
 *   prove:    ozl_groth16_prove(z, r, s) -> affine A, B, C          (twice: must be identical)
 *   extra:    ozl_msm over a_query with the assignment as pageable scalars (the batched H2D path)
 *
 * Input: a key/witness blob written by tests/test_c_replay.py (layout below); output: the proof limbs and
 * the MSM result, which the test compares with the in-process ctypes path.  Every array of the blob is
 * copied into its own malloc'd buffer before use.
 *
 *   u64 header[8] = {magic, pairing, n_constraints, n_instance, n_vars, n_coef, domain, precompute}
 *   then for M in A, B, C: u32 row_ptr[n_constraints + 1], u32 col_idx[nnz], u32 coef_idx[nnz]
 *   u64 coef_table[n_coef * 4]
 *   for Q in a, b_g1, b_g2, h, l: u64 count, u64 pts[count * 2 * L(Q)], u8 inf[(count + 7) / 8] (padded to 8)
 *   u64 alpha_g1[2 L1], beta_g1[2 L1], delta_g1[2 L1], beta_g2[2 L2], delta_g2[2 L2]
 *   u64 z[n_vars * 4] (Montgomery), u64 zc[n_vars * 4] (canonical), u64 r[4], u64 s[4]
 *
 * Build: gcc -O2 -Iinclude examples/ozl_groth16_replay.c -Lopenzl_b200 -lozl_b200 -Wl,-rpath,$PWD/openzl_b200 -o replay
 * Exit status: 0 ok, 3 no usable GPU, 1 failure. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ozl.h"

static ozl_ctx* ctx = NULL;

#define CHECK(call)                                                                         \
  do {                                                                                      \
    int _s = (call);                                                                        \
    if (_s != OZL_OK) {                                                                     \
      fprintf(stderr, "%s -> %s (%s)\n", #call, ozl_strerror(_s), ctx ? ozl_last_error(ctx) : ""); \
      return _s == OZL_ERR_NO_DEVICE ? 3 : 1;                                               \
    }                                                                                       \
  } while (0)

static void* take(FILE* f, size_t bytes) {   /* next `bytes` of the blob in a fresh pageable buffer */
  void* p = malloc(bytes ? bytes : 1);
  if (!p || fread(p, 1, bytes, f) != bytes) {
    fprintf(stderr, "blob truncated\n");
    exit(1);
  }
  return p;
}

int main(int argc, char** argv) {
  if (argc < 3) {
    fprintf(stderr, "usage: %s key_and_witness.blob proof.out\n", argv[0]);
    return 1;
  }
  CHECK(ozl_ctx_create(0, &ctx));   /* before touching the blob: without a GPU this is the only thing that runs */
  FILE* f = fopen(argv[1], "rb");
  if (!f) { perror(argv[1]); return 1; }
  uint64_t* hd = (uint64_t*)take(f, 64);
  if (hd[0] != 0x4f5a4c5245504c59ull) { fprintf(stderr, "bad magic\n"); return 1; }
  const int pairing = (int)hd[1];
  const uint32_t nc = (uint32_t)hd[2], ni = (uint32_t)hd[3], nv = (uint32_t)hd[4], ncoef = (uint32_t)hd[5];
  const int precompute = (int)hd[7];
  const int g1 = pairing == OZL_PAIRING_BN254 ? OZL_BN254_G1 : OZL_BLS12_381_G1;
  const int g2 = pairing == OZL_PAIRING_BN254 ? OZL_BN254_G2 : OZL_BLS12_381_G2;
  const size_t L1 = (size_t)ozl_curve_coord_limbs(g1), L2 = (size_t)ozl_curve_coord_limbs(g2);

  ozl_csr M[3];
  for (int k = 0; k < 3; k++) {
    uint32_t* rp = (uint32_t*)take(f, ((size_t)nc + 1) * 4);
    const size_t nnz = rp[nc];
    M[k].n_rows = nc;
    M[k].row_ptr = rp;
    M[k].col_idx = (uint32_t*)take(f, nnz * 4);
    M[k].coef_idx = (uint32_t*)take(f, nnz * 4);
  }
  uint64_t* coef = (uint64_t*)take(f, (size_t)ncoef * 32);

  /* compile: the five query vectors go to the device once (ProvingContext is constant across proofs) */
  uint32_t q[5];
  size_t qn[5];
  const int qcurve[5] = {g1, g1, g2, g1, g1};
  for (int k = 0; k < 5; k++) {
    uint64_t* cnt = (uint64_t*)take(f, 8);
    qn[k] = (size_t)*cnt;
    const size_t L = qcurve[k] == g1 ? L1 : L2;
    uint64_t* pts = (uint64_t*)take(f, qn[k] * 2 * L * 8);
    uint8_t* inf = (uint8_t*)take(f, (((qn[k] + 7) / 8) + 7) & ~(size_t)7);
    CHECK(ozl_msm_bases_upload(ctx, qcurve[k], pts, inf, qn[k], &q[k]));
    if (precompute > 1) CHECK(ozl_msm_bases_precompute(ctx, q[k], precompute));
    free(pts); free(inf); free(cnt);        /* the library owns device copies now */
  }
  uint64_t* alpha1 = (uint64_t*)take(f, 2 * L1 * 8);
  uint64_t* beta1 = (uint64_t*)take(f, 2 * L1 * 8);
  uint64_t* delta1 = (uint64_t*)take(f, 2 * L1 * 8);
  uint64_t* beta2 = (uint64_t*)take(f, 2 * L2 * 8);
  uint64_t* delta2 = (uint64_t*)take(f, 2 * L2 * 8);
  uint64_t* z = (uint64_t*)take(f, (size_t)nv * 32);
  uint64_t* zc = (uint64_t*)take(f, (size_t)nv * 32);
  uint64_t* r = (uint64_t*)take(f, 32);
  uint64_t* s = (uint64_t*)take(f, 32);
  fclose(f);

  /* the a_query MSM through ozl_msm with pageable scalars, before the pk takes the handle over */
  uint64_t* msm_out = (uint64_t*)calloc(3 * L1, 8);
  CHECK(ozl_msm(ctx, q[0], zc, nv, msm_out));

  uint32_t pk = 0;
  CHECK(ozl_groth16_pk_create(ctx, pairing, nc, ni, nv, &M[0], &M[1], &M[2], coef, ncoef, q[0], q[1], q[2], q[3], q[4],
                              alpha1, beta1, delta1, beta2, delta2, &pk));
  uint32_t domain = 0;
  CHECK(ozl_groth16_domain_size(ctx, pk, &domain));
  if (domain != (uint32_t)hd[6]) { fprintf(stderr, "domain %u != %llu\n", domain, (unsigned long long)hd[6]); return 1; }

  /* prove, twice */
  const size_t plen = 2 * L1 + 2 * L2 + 2 * L1;
  uint64_t* p1 = (uint64_t*)calloc(plen, 8);
  uint64_t* p2 = (uint64_t*)calloc(plen, 8);
  CHECK(ozl_groth16_prove(ctx, pk, z, r, s, p1, p1 + 2 * L1, p1 + 2 * L1 + 2 * L2, NULL));
  CHECK(ozl_groth16_prove(ctx, pk, z, r, s, p2, p2 + 2 * L1, p2 + 2 * L1 + 2 * L2, NULL));
  if (memcmp(p1, p2, plen * 8) != 0) { fprintf(stderr, "two proofs of the same (z, r, s) differ\n"); return 1; }

  FILE* o = fopen(argv[2], "wb");
  if (!o) { perror(argv[2]); return 1; }
  fwrite(p1, 8, plen, o);
  fwrite(msm_out, 8, 3 * L1, o);
  fclose(o);
  CHECK(ozl_groth16_pk_destroy(ctx, pk));     /* frees the five bases handles it owns */
  if (ozl_msm_bases_free(ctx, q[0]) != OZL_ERR_HANDLE) { fprintf(stderr, "pk did not take the bases over\n"); return 1; }
  ozl_ctx_destroy(ctx);
  printf("replay ok: domain %u, %zu proof limbs\n", domain, plen);
  return 0;
}
