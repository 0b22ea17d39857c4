// Montgomery product on the FP64 pipe: the experiment VERDICT r1 item 5 / DESIGN.md section 6 name.
//
// B200 issues ~0.9 DFMA per clock per SM BESIDE a saturated integer multiplier (tools/pipe_probe.cu), and the
// hot kernel leaves the FP64 pipe idle.  This computes the same value as Fp::operator* -- a b / 2^384 mod p,
// fully reduced, for BLS12-381 Fq -- with double-precision arithmetic only, so that some of an addition's
// products can run there while other warps keep the integer pipe busy.
//
// Representation: eight 48-bit limbs (381 <= 8 x 48 = 384, the same R = 2^384 as the 12 x 32-bit form, so
// Montgomery residues are interchangeable), each an exact integer in a double.  A product of two limbs is
// below 2^96; it is cut at bit 48 with the fused multiply-add:
//     s' = fma_rz(x, y, s)      s = 2^100 + (multiple of 2^48): the sum lies in [2^100, 2^101), where one ulp is
//                               2^48, so rounding toward zero drops exactly (x y mod 2^48) -- and the high parts
//                               of up to 16 products ACCUMULATE exactly in the same chain;
//     hp = s' - s               the high part of this product alone (exact);
//     l  = fma(x, y, -hp)       its low part, in [0, 2^48) (exact);
//     L += l                    up to 16 low parts stay below 2^52 (exact).
// Four FP64 operations per 48 x 48 product and no integer instruction.  Product scanning by column k keeps
// three accumulators live; q_k = (column mod 2^48) * (-p^-1) mod 2^48 is obtained with the same cut, and
// adding q_k p_0 clears the column's low 48 bits, whose carry (below 32) moves on.  Every partial sum stays
// below 2^53: at most 14 + 15 or 16 + 14 terms of < 2^48 plus the carry meet in one column.
// All intermediate values are non-negative, so the cuts are floors.  The g++ build of the tests emulates
// these primitives with 128-bit integers and aborts on any inexact step (tests/test_host_emu.py).
#pragma once
#include "fp.cuh"

namespace ozl {

template <class P>
struct Fp64Consts;

// p in 48-bit limbs and -p^-1 mod 2^48 (checked against the 32-bit constants by the tests).  They live in
// constant memory so that DFMA takes them as constant-bank operands: as 64-bit immediates every use cost
// two IMAD.MOV on the very pipe this multiplier is meant to relieve (first build: 490 of them per product).
#if defined(__CUDACC__)
static __device__ __constant__ const double Bls12381Fq_P48[9] = {
#else
static const double Bls12381Fq_P48[9] = {
#endif
    281474976688811.0, 194974335351294.0, 270634993844222.0, 113459389855408.0, 83034393350847.0,
    73992301405303.0,  253550359455670.0, 28591897852287.0,  281462091612157.0 /* [8] = -1/p mod 2^48 */};

template <>
struct Fp64Consts<ozl_params::Bls12381Fq> {
  static OZL_DEV double p48(int k) { return Bls12381Fq_P48[k]; }
  static OZL_DEV double pinv48() { return Bls12381Fq_P48[8]; }
};

// 12 x 32-bit words -> 8 x 48-bit limbs as doubles
OZL_DEV void fp64_split48(const uint32_t* w, double* A) {
#pragma unroll
  for (int m = 0; m < 4; m++) {
    A[2 * m] = ptx::u48_to_double(w[3 * m], w[3 * m + 1] & 0xffffu);
    A[2 * m + 1] = ptx::u48_to_double((w[3 * m + 1] >> 16) | (w[3 * m + 2] << 16), w[3 * m + 2] >> 16);
  }
}

template <class P>
OZL_DEV Fp<P> mul_fp64(const Fp<P>& a, const Fp<P>& b) {
  static_assert(P::N == 12, "48-bit limbs: 12 words = 8 limbs");
  typedef Fp64Consts<P> C;
  constexpr int K = 8;
  const double C1 = 1267650600228229401496703205376.0;      // 2^100
  const double INV48 = 3.552713678800500929355621337890625e-15;   // 2^-48
  double A[K], B[K], Q[K], T[K];
  fp64_split48(a.v, A);
  fp64_split48(b.v, B);
  double carry = 0.0, hprev = 0.0;
#pragma unroll
  for (int k = 0; k < 2 * K; k++) {
    double s = C1, L = 0.0;
    const int i0 = k < K ? 0 : k - (K - 1), i1 = k < K ? k : K - 1;
#pragma unroll
    for (int i = i0; i <= i1; i++) {            // a_i b_(k-i)
      const double sn = ptx::fma_rz(A[i], B[k - i], s);
      const double hp = ptx::add_x(sn, -s);
      L = ptx::add_x(L, ptx::fma_x(A[i], B[k - i], -hp));
      s = sn;
    }
#pragma unroll
    for (int i = i0; i <= (k < K ? k - 1 : K - 1); i++) {   // q_i p_(k-i), q_i already known
      const double pj = C::p48(k - i);
      const double sn = ptx::fma_rz(Q[i], pj, s);
      const double hp = ptx::add_x(sn, -s);
      L = ptx::add_x(L, ptx::fma_x(Q[i], pj, -hp));
      s = sn;
    }
    const double V = ptx::add_x(ptx::add_x(carry, hprev), L);
    const double t = ptx::add_x(ptx::add_rz(V, C1), -C1);    // V cut at bit 48
    const double r = ptx::add_x(V, -t);                        // V mod 2^48
    if (k < K) {
      const double ph = ptx::fma_rz(r, C::pinv48(), C1);
      const double q = ptx::fma_x(r, C::pinv48(), -ptx::add_x(ph, -C1));   // r * (-1/p) mod 2^48
      Q[k] = q;
      const double sn = ptx::fma_rz(q, C::p48(0), s);
      const double hp = ptx::add_x(sn, -s);
      const double l0 = ptx::fma_x(q, C::p48(0), -hp);
      s = sn;
      carry = ptx::mul_x(ptx::add_x(V, l0), INV48);            // low 48 bits of V + l0 are zero
    } else {
      T[k - K] = r;
      carry = ptx::mul_x(t, INV48);
    }
    hprev = ptx::mul_x(ptx::add_x(s, -C1), INV48);
  }
  // T < 2 p: back to 32-bit words, one conditional subtraction
  Fp<P> out;
#pragma unroll
  for (int m = 0; m < 4; m++) {
    uint32_t l0, h0, l1, h1;
    ptx::double_to_u48(T[2 * m], l0, h0);
    ptx::double_to_u48(T[2 * m + 1], l1, h1);
    out.v[3 * m] = l0;
    out.v[3 * m + 1] = h0 | (l1 << 16);
    out.v[3 * m + 2] = (l1 >> 16) | (h1 << 16);
  }
  Fp<P>::final_sub(out.v);
  return out;
}

template <class P>
OZL_DEV_NOINLINE Fp<P> mul_fp64_ni(Fp<P> a, Fp<P> b) { return mul_fp64<P>(a, b); }

// One integer-pipe product and one FP64-pipe product in ONE body, so that ptxas interleaves the two
// instruction streams and a single warp feeds both pipes (phase mixing across warps alone was measured
// to give no overlap: at 3-4 warps per scheduler each kind of product is latency-bound on its own).
template <class P>
struct FpPairIF {
  Fp<P> i, f;
};
template <class P>
OZL_DEV_NOINLINE FpPairIF<P> mul_int_fp64_pair_ni(Fp<P> a, Fp<P> b, Fp<P> c, Fp<P> d) {
  FpPairIF<P> r;
  r.f = mul_fp64<P>(c, d);
  r.i = a * b;
  return r;
}

}  // namespace ozl
