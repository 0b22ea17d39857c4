// Short-Weierstrass group arithmetic (a = 0) generic over the coordinate field
// (Fp for G1, Fp2 for G2).
//
// Replaces ark_ec::models::short_weierstrass_jacobian::{GroupAffine, GroupProjective}
// (ark-ec 0.3.0) behind Pairing::{G1, G2} (/root/reference/plugins/arkworks/src/pairing.rs:14-23).
// Bucket accumulators use extended Jacobian ("XYZZ") coordinates: x = X/ZZ, y = Y/ZZZ with
// ZZ^3 = ZZZ^2, identity <=> ZZ == 0.  A mixed addition costs 8M + 2S instead of ark's
// madd-2007-bl 7M + 4S, needs no final doubling of Z, and the result is the same group
// element.  The C ABI returns ark's Jacobian (X, Y, Z) layout: (X*ZZ, Y*ZZZ, ZZ).
#pragma once
#include "fp.cuh"
#include "fp64mul.cuh"

namespace ozl {

template <class F>
struct Affine {
  F x, y;
};

template <class F>
struct XYZZ {
  F x, y, zz, zzz;

  static OZL_DEV XYZZ identity() {
    XYZZ r;
    r.x = F::zero(); r.y = F::one(); r.zz = F::zero(); r.zzz = F::zero();
    return r;
  }
  OZL_DEV bool is_identity() const { return zz.is_zero(); }

  static OZL_DEV XYZZ from_affine(const Affine<F>& p) {
    XYZZ r;
    r.x = p.x; r.y = p.y; r.zz = F::one(); r.zzz = F::one();
    return r;
  }

  OZL_DEV XYZZ neg() const { XYZZ r = *this; r.y = y.neg(); return r; }

  // The cold-path formulas below issue their field products in independent PAIRS through
  // F::mul2_ni / F::sqr2_ni (one out-of-line body, two interleaved carry chains): the kernels that
  // use them run few warps per SM and are bound by dependent-issue latency, not by pipe throughput.

  // 2 * (affine p), mdbl-2008-s with a = 0
  static OZL_DEV_NOINLINE XYZZ dbl_affine(const Affine<F>& p) {
    XYZZ r;
    F u = p.y.dbl();
    typename F::Pair a = F::sqr2_ni(u, p.x);              // v = u^2, xx = x^2
    F m = a.b.dbl() + a.b;
    typename F::Pair b = F::mul2_ni(u, a.a, p.x, a.a);    // w = u v, s = x v
    typename F::Pair c = F::mul2_ni(m, m, b.a, p.y);      // m^2, w y
    r.x = c.a - b.b.dbl();
    r.y = F::mul_ni(m, b.b - r.x) - c.b;
    r.zz = a.a;
    r.zzz = b.a;
    return r;
  }

  // dbl-2008-s with a = 0
  OZL_DEV_NOINLINE XYZZ dbl() const {
    if (is_identity()) return *this;
    XYZZ r;
    F u = y.dbl();
    typename F::Pair a = F::sqr2_ni(u, x);                // v = u^2, xx = x^2
    F m = a.b.dbl() + a.b;
    typename F::Pair b = F::mul2_ni(u, a.a, x, a.a);      // w = u v, s = x v
    typename F::Pair c = F::mul2_ni(m, m, b.a, y);        // m^2, w y
    r.x = c.a - b.b.dbl();
    typename F::Pair d = F::mul2_ni(m, b.b - r.x, a.a, zz);
    r.y = d.a - c.b;
    r.zz = d.b;
    r.zzz = F::mul_ni(b.a, zzz);
    return r;
  }

  // this += affine p (madd-2008-s).  p must be a finite point.  FUSED: y3 = r (q - x3) - y p3 as one
  // dual product with a single Montgomery reduction (Fp::mul_add2) -- base fields only.
  template <bool FUSED = false>
  OZL_DEV void add_mixed(const Affine<F>& p) {
    if (is_identity()) {
      *this = from_affine(p);
      return;
    }
    F u2 = p.x * zz;
    F s2 = p.y * zzz;
    F pp = u2 - x;   // P
    F r = s2 - y;    // R
    if (pp.is_zero()) {
      if (r.is_zero()) {
        *this = dbl_affine(p);
      } else {
        *this = identity();
      }
      return;
    }
    F p2 = pp.sqr();
    F p3 = pp * p2;
    F q = x * p2;
    F x3 = r.sqr() - p3 - q.dbl();
    if (FUSED) y = F::mul_add2(r, q - x3, y.neg(), p3);
    else y = r * (q - x3) - y * p3;
    x = x3;
    zz = zz * p2;
    zzz = zzz * p3;
  }

  // add_mixed with its ten field products issued through out-of-line multiplier bodies, operands by value
  // (registers).  MODE 1: ten calls of F::mul_ni; MODE 2: five calls of the paired F::mul2_ni; MODE 3: MODE 1 with
  // the two squarings through the dedicated F::sqr_sos_ni; MODE 4: MODE 3 with the eight products through the
  // Karatsuba body F::mul_kara_ni; MODE 5: MODE 3 with y3 = r (q - x3) - y p3 as ONE fused product pair
  // (F::mul_add2_ni: 3 N^2 instead of 4 N^2 wide multiplies, a single reduction).  Exists because
  // the fully inlined loop body of k_accumulate (~100 KB of SASS) does not fit the instruction cache
  // (ncu: sm__icc_request_hit_rate 83.5 %, stalled_no_instruction 1.26 warps per issue); which variant
  // the hot kernel uses is decided by measurement (OZL_ACC_MODE, see msm.cuh).
  // product on the FP64 pipe where the field has the 48-bit constants (BLS12-381 Fq), the integer body otherwise
  template <class G>
  static OZL_DEV G mul_fp64_or_int(const G& a, const G& b) { return G::mul_ni(a, b); }
  static OZL_DEV Fp<ozl_params::Bls12381Fq> mul_fp64_or_int(const Fp<ozl_params::Bls12381Fq>& a, const Fp<ozl_params::Bls12381Fq>& b) {
    return mul_fp64_ni<ozl_params::Bls12381Fq>(a, b);
  }

  // MODE 8 (quadratic extension only): lazily reduced Fq2 products -- three unreduced base products and two Montgomery
  // reductions per multiplication (5 N^2 instead of 6 N^2 wide multiplies), y3 as six products and two reductions
  // (8 N^2 instead of 12 N^2).  On a prime field MODE 8 is MODE 5.
  static constexpr bool kExt = F::N != F::Params::N;
  template <int MODE>
  static OZL_DEV F mul_m(const F& a, const F& b) {
    if constexpr (MODE == 8 && kExt) return F::mul_lazy_ni(a, b);
    else return MODE == 4 ? F::mul_kara_ni(a, b) : F::mul_ni(a, b);
  }
  template <int MODE>
  static OZL_DEV F sqr_m(const F& a) { return MODE >= 3 ? F::sqr_sos_ni(a) : F::mul_ni(a, a); }

  template <int MODE>
  OZL_DEV void add_mixed_calls(const Affine<F>& p) {
    if (is_identity()) {
      *this = from_affine(p);
      return;
    }
    F u2, s2;
    if (MODE == 2) {
      typename F::Pair a = F::mul2_ni(p.x, zz, p.y, zzz);
      u2 = a.a; s2 = a.b;
    } else {
      u2 = mul_m<MODE>(p.x, zz); s2 = mul_m<MODE>(p.y, zzz);
    }
    F pp = u2 - x;
    F r = s2 - y;
    if (pp.is_zero()) {
      if (r.is_zero()) {
        *this = dbl_affine(p);
      } else {
        *this = identity();
      }
      return;
    }
    if (MODE == 2) {
      typename F::Pair b = F::sqr2_ni(pp, r);               // p2, r^2
      typename F::Pair c = F::mul2_ni(pp, b.a, x, b.a);     // p3, q
      F x3 = b.b - c.a - c.b.dbl();
      typename F::Pair d = F::mul2_ni(r, c.b - x3, y, c.a);
      typename F::Pair e = F::mul2_ni(zz, b.a, zzz, c.a);
      y = d.a - d.b;
      x = x3;
      zz = e.a;
      zzz = e.b;
    } else {
      F p2 = sqr_m<MODE>(pp);
      F p3 = mul_m<MODE>(pp, p2);
      F q = mul_m<MODE>(x, p2);
      F x3 = sqr_m<MODE>(r) - p3 - q.dbl();
      // y3 = r (q - x3) - y p3: MODE >= 5 folds the two products into one Montgomery reduction.
      // MODE 7: the two products nothing in this addition waits for (zz3, zzz3 feed the NEXT addition) run on
      // the FP64 pipe (fp64mul.cuh) while other warps keep the integer multiplier busy.
      if (MODE == 7) {
        y = F::mul_add2_ni(r, q - x3, y.neg(), p3);
        zz = mul_fp64_or_int(zz, p2);
        zzz = mul_fp64_or_int(zzz, p3);
      } else {
        zz = mul_m<MODE>(zz, p2);
        zzz = mul_m<MODE>(zzz, p3);
      }
      if (MODE == 7) {
      } else if (MODE == 8 && kExt) {
        if constexpr (kExt) y = F::mul_sub2_lazy_ni(r, q - x3, y, p3);
      } else if (MODE == 5 || MODE == 8) y = F::mul_add2_ni(r, q - x3, y.neg(), p3);
      else y = mul_m<MODE>(r, q - x3) - mul_m<MODE>(y, p3);
      x = x3;
    }
  }

  // out-of-line copy of add_mixed for cold kernels (keeps their code size small)
  OZL_DEV_NOINLINE void add_mixed_cold(const Affine<F>& p) {
    if (is_identity()) {
      *this = from_affine(p);
      return;
    }
    typename F::Pair a = F::mul2_ni(p.x, zz, p.y, zzz);   // u2, s2
    F pp = a.a - x;
    F r = a.b - y;
    if (pp.is_zero()) {
      if (r.is_zero()) {
        *this = dbl_affine(p);
      } else {
        *this = identity();
      }
      return;
    }
    typename F::Pair b = F::sqr2_ni(pp, r);               // p2, r^2
    typename F::Pair c = F::mul2_ni(pp, b.a, x, b.a);     // p3, q
    F x3 = b.b - c.a - c.b.dbl();
    typename F::Pair e = F::mul2_ni(zz, b.a, zzz, c.a);
    y = F::mul_add2_ni(r, c.b - x3, y.neg(), c.a);        // r (q - x3) - y p3, one reduction
    x = x3;
    zz = e.a;
    zzz = e.b;
  }

  // this += o (add-2008-s)
  OZL_DEV_NOINLINE void add(const XYZZ& o) {
    if (o.is_identity()) return;
    if (is_identity()) {
      *this = o;
      return;
    }
    typename F::Pair u = F::mul2_ni(x, o.zz, o.x, zz);    // u1, u2
    typename F::Pair s = F::mul2_ni(y, o.zzz, o.y, zzz);  // s1, s2
    F pp = u.b - u.a;
    F r = s.b - s.a;
    if (pp.is_zero()) {
      if (r.is_zero()) {
        *this = dbl();
      } else {
        *this = identity();
      }
      return;
    }
    typename F::Pair b = F::sqr2_ni(pp, r);               // p2, r^2
    typename F::Pair c = F::mul2_ni(pp, b.a, u.a, b.a);   // p3, q
    F x3 = b.b - c.a - c.b.dbl();
    typename F::Pair z = F::mul2_ni(zz, o.zz, zzz, o.zzz);
    typename F::Pair e = F::mul2_ni(z.a, b.a, z.b, c.a);
    y = F::mul_add2_ni(r, c.b - x3, s.a.neg(), c.a);      // r (q - x3) - s1 p3, one reduction
    x = x3;
    zz = e.a;
    zzz = e.b;
  }

  // [k] * this for a small unsigned k (left-to-right double-and-add); cold path.
  OZL_DEV_NOINLINE XYZZ mul_u32(uint32_t k) const {
    XYZZ acc = identity();
    for (int bit = 31; bit >= 0; bit--) {
      acc = acc.dbl();
      if ((k >> bit) & 1) acc.add(*this);
    }
    return acc;
  }

  // ark Jacobian (X, Y, Z): x = X/Z^2, y = Y/Z^3 with Z := ZZ
  OZL_DEV void to_jacobian(F& X, F& Y, F& Z) const {
    if (is_identity()) {
      X = F::zero(); Y = F::one(); Z = F::zero();  // ark GroupProjective::zero()
      return;
    }
    X = x * zz;
    Y = y * zzz;
    Z = zz;
  }

  // affine (x, y); returns false for the identity
  OZL_DEV_NOINLINE bool to_affine(Affine<F>& p) const {
    if (is_identity()) return false;
    F zi = zzz.inverse();                      // 1/Z^3
    F zi2 = F::sqr_ni(F::mul_ni(zi, zz));      // (1/Z)^2 = (ZZ/ZZZ)^2
    p.x = F::mul_ni(x, zi2);
    p.y = F::mul_ni(y, zi);
    return true;
  }

  static OZL_DEV XYZZ load(const uint32_t* p) {
    XYZZ r;
    r.x = F::load(p); r.y = F::load(p + F::N); r.zz = F::load(p + 2 * F::N); r.zzz = F::load(p + 3 * F::N);
    return r;
  }
  OZL_DEV void store(uint32_t* p) const {
    x.store(p); y.store(p + F::N); zz.store(p + 2 * F::N); zzz.store(p + 3 * F::N);
  }
};

template <class F>
OZL_DEV Affine<F> load_affine(const uint32_t* p) {
  Affine<F> r;
  r.x = F::load(p);
  r.y = F::load(p + F::N);
  return r;
}

}  // namespace ozl
