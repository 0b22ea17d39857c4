// Scalar-field vector kernels around the NTT for the Groth16 witness map
// (ark_groth16::R1CStoQAP::witness_map, ark-groth16 0.3.0, called behind
// /root/reference/plugins/arkworks/src/groth16.rs:454): sparse matrix-vector products <A_i, z>,
// the pointwise (a*b - c) / Z(g) step on the coset, and Montgomery -> canonical conversion
// (`into_repr`) of the MSM scalars.  All HBM-streaming except the SpMV gather.
#pragma once
#include <cuda_runtime.h>
#include "fp.cuh"

namespace ozl {

// y[row] = sum_k coef[cidx[k]] * x[col[k]]   (CSR; coefficients come from a small table because an
// R1CS built from a few gadgets has very few distinct constants)
template <class P>
__global__ void __launch_bounds__(256)
k_spmv(const uint32_t* __restrict__ row_ptr, const uint32_t* __restrict__ col, const uint32_t* __restrict__ cidx,
       const uint32_t* __restrict__ coef, const uint32_t* __restrict__ x, uint32_t n_rows, uint32_t* __restrict__ y) {
  typedef Fp<P> F;
  const uint32_t row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= n_rows) return;
  F acc = F::zero();
  const uint32_t k1 = row_ptr[row + 1];
  for (uint32_t k = row_ptr[row]; k < k1; k++) {
    const F c = F::load(coef + (size_t)cidx[k] * F::N);
    const F v = F::load(x + (size_t)col[k] * F::N);
    acc = acc + c * v;
  }
  acc.store(y + (size_t)row * F::N);
}

template <class P>
__global__ void __launch_bounds__(256) k_from_mont(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, uint32_t n) {
  typedef Fp<P> F;
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  F::load(in + (size_t)i * F::N).from_mont().store(out + (size_t)i * F::N);
}

// a[i] = (a[i] * b[i] - c[i]) * scale      (scale = 1 / (g^n - 1): the vanishing polynomial is
// constant on the coset gH, ark's divide_by_vanishing_poly_on_coset_in_place)
template <class P>
__global__ void __launch_bounds__(256)
k_h_pointwise(uint32_t* __restrict__ a, const uint32_t* __restrict__ b, const uint32_t* __restrict__ c,
              const Fp<P>* __restrict__ scale, uint32_t n) {
  typedef Fp<P> F;
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const F s = *scale;
  F v = F::load(a + (size_t)i * F::N) * F::load(b + (size_t)i * F::N) - F::load(c + (size_t)i * F::N);
  (v * s).store(a + (size_t)i * F::N);
}

// out = 1 / (g^(2^log_n) - 1), Montgomery form
template <class P>
__global__ void k_vanishing_inv(int log_n, Fp<P>* out) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  typedef Fp<P> F;
  F g = F::from_limbs(P::gen());
  for (int i = 0; i < log_n; i++) g = F::sqr_ni(g);
  *out = (g - F::one()).inverse();
}

// out = a * b mod r on canonical integers (r*s for the proof assembly)
template <class P>
__global__ void k_mul_canonical(const uint32_t* __restrict__ a, const uint32_t* __restrict__ b, uint32_t* __restrict__ out) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  typedef Fp<P> F;
  F x = F::from_limbs(a).to_mont(), y = F::from_limbs(b).to_mont();
  F r = (x * y).from_mont();
#pragma unroll
  for (int i = 0; i < F::N; i++) out[i] = r.v[i];
}

// Poseidon permutation, one state per thread: the workload hash of the Groth16 benchmark circuit
// (/root/reference/openzl-crypto/src/poseidon/mod.rs:156-283: a round adds the round keys, applies the
// S-box x^5 -- to every element in the R_F/2 leading and trailing full rounds, to element 0 in the R_P
// partial rounds between them -- and multiplies by the MDS matrix, row-major, next[i] = sum_j mds[i][j] state[j];
// S-box at /root/reference/plugins/arkworks/src/poseidon/mod.rs:147-159).  Evaluates witnesses for
// batches of hashes on the device and pins the device's Montgomery arithmetic to the reference's
// width-3 known-answer test (tests/test_gpu_poseidon.py).
static constexpr int POSEIDON_MAX_WIDTH = 12;   // the reference tabulates MDS matrices for widths 2..12

template <class P>
__global__ void __launch_bounds__(128)
k_poseidon_permute(uint32_t* __restrict__ states, uint32_t batch, int width, int full_rounds, int partial_rounds,
                   const uint32_t* __restrict__ round_keys, const uint32_t* __restrict__ mds) {
  typedef Fp<P> F;
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= batch) return;
  F st[POSEIDON_MAX_WIDTH], nx[POSEIDON_MAX_WIDTH];
  uint32_t* mine = states + (size_t)i * width * F::N;
  for (int k = 0; k < width; k++) st[k] = F::load(mine + k * F::N);
  const int half = full_rounds / 2;
  for (int rnd = 0; rnd < full_rounds + partial_rounds; rnd++) {
    const bool full = rnd < half || rnd >= half + partial_rounds;
    for (int k = 0; k < width; k++) {
      F v = st[k] + F::load(round_keys + ((size_t)rnd * width + k) * F::N);
      if (full || k == 0) {
        const F v2 = F::mul_ni(v, v);
        v = F::mul_ni(F::mul_ni(v2, v2), v);
      }
      st[k] = v;
    }
    for (int r = 0; r < width; r++) {
      F acc = F::zero();
      for (int k = 0; k < width; k++) acc = acc + F::mul_ni(F::load(mds + ((size_t)r * width + k) * F::N), st[k]);
      nx[r] = acc;
    }
    for (int k = 0; k < width; k++) st[k] = nx[k];
  }
  for (int k = 0; k < width; k++) st[k].store(mine + k * F::N);
}

}  // namespace ozl
