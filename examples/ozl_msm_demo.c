/* Plain-C caller of libozl_b200 (include/ozl.h): what a non-Python host -- the Rust shim's FFI, a C++
 * prover -- does for one MSM and one NTT.  Self-checking without any oracle:
 *
 *   MSM   bases P_i = [i + 1]G (ozl_msm_bases_generate), scalars s_i = 2 for all i
 *         =>  sum_i s_i P_i = [n (n + 1)] G, compared with ozl_fixed_base_mul of that scalar
 *   NTT   forward then inverse transform returns the input (BN254 Fr, Montgomery limbs)
 *
 * Build:  gcc -O2 -Iinclude examples/ozl_msm_demo.c -Lopenzl_b200 -lozl_b200 -Wl,-rpath,$PWD/openzl_b200 -o ozl_msm_demo
 * Exit status: 0 = both checks passed, 3 = no usable GPU (the library has no CPU fallback), 1 = failure. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ozl.h"

#define CHECK(call)                                                                         \
  do {                                                                                      \
    int _s = (call);                                                                        \
    if (_s != OZL_OK) {                                                                     \
      fprintf(stderr, "%s -> %s (%s)\n", #call, ozl_strerror(_s), ctx ? ozl_last_error(ctx) : ""); \
      return _s == OZL_ERR_NO_DEVICE ? 3 : 1;                                               \
    }                                                                                       \
  } while (0)

int main(int argc, char** argv) {
  const size_t n = argc > 1 ? (size_t)strtoull(argv[1], NULL, 10) : 100000;
  ozl_ctx* ctx = NULL;
  CHECK(ozl_ctx_create(0, &ctx));

  /* ---- MSM over BLS12-381 G1 ---- */
  const int curve = OZL_BLS12_381_G1;
  const int L = ozl_curve_coord_limbs(curve);            /* 6 u64 limbs per coordinate */
  uint32_t bases = 0;
  CHECK(ozl_msm_bases_generate(ctx, curve, 1, n, &bases));
  uint64_t* scalars = (uint64_t*)calloc(n, 32);          /* canonical BigInteger256, little-endian limbs */
  for (size_t i = 0; i < n; i++) scalars[4 * i] = 2;
  uint64_t jac[18], aff[12], expect[12];
  int is_inf = 0;
  CHECK(ozl_msm(ctx, bases, scalars, n, jac));
  CHECK(ozl_jacobian_to_affine(ctx, curve, jac, aff, &is_inf));
  uint64_t k[4] = {(uint64_t)n * (uint64_t)(n + 1), 0, 0, 0};   /* n < 2^32 keeps this in one limb */
  uint8_t flag = 0;
  CHECK(ozl_fixed_base_mul(ctx, curve, k, 1, expect, &flag));
  const int msm_ok = !is_inf && !flag && memcmp(aff, expect, sizeof(uint64_t) * 2 * (size_t)L) == 0;
  printf("MSM  n=%zu  sum_i 2*[i+1]G == [n(n+1)]G : %s\n", n, msm_ok ? "ok" : "MISMATCH");
  CHECK(ozl_msm_bases_free(ctx, bases));
  free(scalars);

  /* ---- NTT over BN254 Fr: inverse(forward(x)) == x ---- */
  const uint32_t log_n = 12;
  const size_t m = (size_t)1 << log_n;
  uint64_t* x = (uint64_t*)malloc(m * 32);
  uint64_t* y = (uint64_t*)malloc(m * 32);
  uint64_t state = 0x9E3779B97F4A7C15ull;
  for (size_t i = 0; i < 4 * m; i++) {                   /* any residues below 2^253 are valid field elements */
    state = state * 6364136223846793005ull + 1442695040888963407ull;
    x[i] = (i % 4 == 3) ? (state >> 11) & 0x0FFFFFFFFFFFFFFFull : state;
  }
  memcpy(y, x, m * 32);
  CHECK(ozl_ntt(ctx, OZL_BN254_FR, y, log_n, 0, 0));
  const int changed = memcmp(x, y, m * 32) != 0;
  CHECK(ozl_ntt(ctx, OZL_BN254_FR, y, log_n, 1, 0));
  const int ntt_ok = changed && memcmp(x, y, m * 32) == 0;
  printf("NTT  2^%u  ifft(fft(x)) == x : %s\n", log_n, ntt_ok ? "ok" : "MISMATCH");
  free(x);
  free(y);
  ozl_ctx_destroy(ctx);
  return msm_ok && ntt_ok ? 0 : 1;
}
