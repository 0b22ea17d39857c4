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

}  // namespace ozl
