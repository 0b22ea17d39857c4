// g++ build of the device arithmetic headers (fp.cuh / ec.cuh) with the PTX carry-chain
// primitives emulated in software (-DOZL_HOST_EMU).  TEST ONLY: lets the CPU test-suite check
// the limb-level algorithms against the big-int oracle on a machine without a GPU.  Never
// linked into libozl_b200.so.
#include <cstring>
#include "../../openzl_b200/csrc/params_gen.cuh"
#include "../../openzl_b200/csrc/ec.cuh"
#include "../../openzl_b200/csrc/fp64mul.cuh"

using namespace ozl;
using namespace ozl_params;

template <class F>
static void fp_op(int op, const uint32_t* a, const uint32_t* b, uint32_t* out) {
  F x = F::from_limbs(a), y = F::from_limbs(b), r;
  switch (op) {
    case 0: r = x * y; break;
    case 1: r = x + y; break;
    case 2: r = x - y; break;
    case 3: r = x.neg(); break;
    case 4: r = x.inverse(); break;
    case 5: r = x.sqr(); break;
    case 6: r = x.dbl(); break;
    case 7: r = x.sqr_sos(); break;
    case 8: r = x.mul_kara(y); break;
    case 9:   // x y - y (x + y) through the single-reduction form (fields with p < R / 4 only)
      if constexpr (F::Params::BITS <= 32 * F::Params::N - 2) r = F::mul_add2_ni(x, y, y.neg(), x + y);
      else r = x * y - y * (x + y);
      break;
    case 10:  // x x + y y: all four operands at their maximum when x = y = p - 1
      if constexpr (F::Params::BITS <= 32 * F::Params::N - 2) r = F::mul_add2_ni(x, x, y, y);
      else r = x * x + y * y;
      break;
    case 11:  // unreduced two-accumulator product followed by the stand-alone CIOS reduction (prime fields)
      if constexpr (F::N == F::Params::N) {
        uint32_t t[2 * F::N];
        F::mul_wide(x.v, y.v, t);
        r = F::redc_cios(t);
      } else {
        r = x.mul_lazy(y);                       // Fq2: lazily reduced Karatsuba (3 products, 2 reductions)
      }
      break;
    case 12:  // Fq2: x y - y (x + y) with two reductions
      if constexpr (F::N != F::Params::N) r = F::mul_sub2_lazy_ni(x, y, y, x + y);
      else r = x * y - y * (x + y);
      break;
    default: r = F::zero();
  }
  memcpy(out, &r, sizeof(F));
}

// raw limb-level entry points: a, b are ARBITRARY N-limb integers (not residues), t an arbitrary 2N-limb integer < p R
template <class F>
static void wide_op(int op, const uint32_t* a, const uint32_t* b, uint32_t* out) {
  if (op == 0) {
    F::mul_wide(a, b, out);
  } else {
    F r = F::redc_cios(a);
    memcpy(out, &r, sizeof(F));
  }
}

template <class F>
static void ec_op(int op, const uint32_t* a, const uint32_t* b, uint32_t k, uint32_t* out) {
  XYZZ<F> acc;
  memcpy(&acc, a, sizeof(acc));
  switch (op) {
    case 0: { Affine<F> p; memcpy(&p, b, sizeof(p)); acc.add_mixed(p); break; }
    case 1: { XYZZ<F> o; memcpy(&o, b, sizeof(o)); acc.add(o); break; }
    case 2: acc = acc.dbl(); break;
    case 3: acc = acc.mul_u32(k); break;
    case 6: { Affine<F> p; memcpy(&p, b, sizeof(p)); acc.add_mixed_cold(p); break; }
    case 4: {  // to affine: out = x||y, returns via k? (identity -> zeros)
      Affine<F> p;
      memset(out, 0, 2 * sizeof(F));
      if (acc.to_affine(p)) memcpy(out, &p, sizeof(p));
      return;
    }
    case 5: {  // to ark jacobian
      F X, Y, Z;
      acc.to_jacobian(X, Y, Z);
      memcpy(out, &X, sizeof(F)); memcpy(out + F::N, &Y, sizeof(F)); memcpy(out + 2 * F::N, &Z, sizeof(F));
      return;
    }
  }
  memcpy(out, &acc, sizeof(acc));
}

extern "C" {
// BLS12-381 Fq product on the emulated FP64 pipe (fp64mul.cuh); aborts if any step is inexact
void emu_mul_fp64(const uint32_t* a, const uint32_t* b, uint32_t* out) {
  typedef Fp<Bls12381Fq> F;
  F r = mul_fp64<Bls12381Fq>(F::from_limbs(a), F::from_limbs(b));
  memcpy(out, &r, sizeof(F));
}
// field ids: 0 Bls12381Fq, 1 Bls12381Fr, 2 Bn254Fq, 3 Bn254Fr, 4 Bls12381Fq2, 5 Bn254Fq2
void emu_fp_op(int field, int op, const uint32_t* a, const uint32_t* b, uint32_t* out) {
  switch (field) {
    case 0: fp_op<Fp<Bls12381Fq>>(op, a, b, out); break;
    case 1: fp_op<Fp<Bls12381Fr>>(op, a, b, out); break;
    case 2: fp_op<Fp<Bn254Fq>>(op, a, b, out); break;
    case 3: fp_op<Fp<Bn254Fr>>(op, a, b, out); break;
    case 4: fp_op<Fp2<Bls12381Fq>>(op, a, b, out); break;
    case 5: fp_op<Fp2<Bn254Fq>>(op, a, b, out); break;
  }
}
// curve ids: 0 Bls12381G1, 1 Bls12381G2, 2 Bn254G1, 3 Bn254G2
void emu_wide_op(int field, int op, const uint32_t* a, const uint32_t* b, uint32_t* out) {
  switch (field) {
    case 0: wide_op<Fp<Bls12381Fq>>(op, a, b, out); break;
    case 1: wide_op<Fp<Bls12381Fr>>(op, a, b, out); break;
    case 2: wide_op<Fp<Bn254Fq>>(op, a, b, out); break;
    case 3: wide_op<Fp<Bn254Fr>>(op, a, b, out); break;
  }
}

void emu_ec_op(int curve, int op, const uint32_t* a, const uint32_t* b, uint32_t k, uint32_t* out) {
  switch (curve) {
    case 0: ec_op<Fp<Bls12381Fq>>(op, a, b, k, out); break;
    case 1: ec_op<Fp2<Bls12381Fq>>(op, a, b, k, out); break;
    case 2: ec_op<Fp<Bn254Fq>>(op, a, b, k, out); break;
    case 3: ec_op<Fp2<Bn254Fq>>(op, a, b, k, out); break;
  }
}
}
