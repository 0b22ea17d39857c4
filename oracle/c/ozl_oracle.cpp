// CPU oracle (C++17, g++) for the openzl_b200 hot path.  TEST INFRASTRUCTURE ONLY -- only
// tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// load this library; the product never links it.
//
// It RESTATES (it is not the reference binary: the arithmetic lives in un-vendored crates.io
// dependencies ark-ec / ark-ff / ark-poly ^0.3.0, /root/reference/plugins/arkworks/Cargo.toml:113-146,
// and no Rust toolchain exists in this image):
//   * ark_ff Fp256/Fp384 Montgomery arithmetic on 64-bit limbs            (msm / ntt below)
//   * ark_ec short_weierstrass_jacobian add_assign_mixed (madd-2007-bl),
//     add_assign (add-2007-bl), double_in_place (dbl-2009-l, a = 0), with ark's guards
//   * ark_ec::msm::VariableBaseMSM::multi_scalar_mul: window rule c = ceil(log2 n)*69/100+2,
//     unsigned digits, zero filter, scalar==1 shortcut, 2^c-1 Jacobian buckets, running sum,
//     high-to-low fold; windows run in parallel like rayon's cfg_into_iter!(window_starts)
//   * ark_poly Radix2EvaluationDomain in-order fft / ifft / coset variants (serial)
// Call chain being restated: /root/reference/plugins/arkworks/src/groth16.rs:445-457
// (Groth16::prove -> ark_groth16::create_proof -> the two families above).
// PARITY: field arithmetic is pinned to the reference's Poseidon golden vectors through the
// Python oracle (tests/test_oracle_golden.py, tests/test_oracle_c.py); MSM / NTT results are
// "parity unpinned" by the reference's own tests and pinned by mathematics only.
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include <thread>
#include <vector>
#include <algorithm>

#include "params_gen.h"

typedef unsigned __int128 u128;
using namespace ozl_params;

// ------------------------------------------------------------------------------------------
// Fp: Montgomery residues on N 64-bit limbs
// ------------------------------------------------------------------------------------------
template <class P>
struct Fp {
  static constexpr int N = P::N;
  static constexpr int LIMBS = P::N;
  uint64_t v[N];

  static Fp zero() { Fp r; memset(r.v, 0, sizeof r.v); return r; }
  static Fp one() { Fp r; memcpy(r.v, P::one(), sizeof r.v); return r; }
  static Fp from(const uint64_t* p) { Fp r; memcpy(r.v, p, sizeof r.v); return r; }
  void to(uint64_t* p) const { memcpy(p, v, sizeof v); }
  bool is_zero() const { uint64_t t = 0; for (int i = 0; i < N; i++) t |= v[i]; return t == 0; }
  bool operator==(const Fp& o) const { return memcmp(v, o.v, sizeof v) == 0; }
  bool operator!=(const Fp& o) const { return !(*this == o); }

  static bool geq_mod(const uint64_t* x) {
    for (int i = N - 1; i >= 0; i--) {
      if (x[i] > P::mod()[i]) return true;
      if (x[i] < P::mod()[i]) return false;
    }
    return true;
  }
  static void sub_mod(uint64_t* x) {
    uint64_t borrow = 0;
    for (int i = 0; i < N; i++) {
      u128 t = (u128)x[i] - P::mod()[i] - borrow;
      x[i] = (uint64_t)t;
      borrow = (uint64_t)(t >> 64) & 1;
    }
  }
  Fp operator+(const Fp& o) const {
    Fp r; uint64_t c = 0;
    for (int i = 0; i < N; i++) { u128 t = (u128)v[i] + o.v[i] + c; r.v[i] = (uint64_t)t; c = (uint64_t)(t >> 64); }
    if (c || geq_mod(r.v)) sub_mod(r.v);
    return r;
  }
  Fp operator-(const Fp& o) const {
    Fp r; uint64_t borrow = 0;
    for (int i = 0; i < N; i++) { u128 t = (u128)v[i] - o.v[i] - borrow; r.v[i] = (uint64_t)t; borrow = (uint64_t)(t >> 64) & 1; }
    if (borrow) { uint64_t c = 0; for (int i = 0; i < N; i++) { u128 t = (u128)r.v[i] + P::mod()[i] + c; r.v[i] = (uint64_t)t; c = (uint64_t)(t >> 64); } }
    return r;
  }
  Fp neg() const { return is_zero() ? *this : zero() - *this; }
  Fp dbl() const { return *this + *this; }
  // CIOS Montgomery product
  Fp operator*(const Fp& o) const {
    uint64_t t[N + 2];
    memset(t, 0, sizeof t);
    for (int i = 0; i < N; i++) {
      uint64_t c = 0;
      for (int j = 0; j < N; j++) { u128 x = (u128)v[j] * o.v[i] + t[j] + c; t[j] = (uint64_t)x; c = (uint64_t)(x >> 64); }
      u128 x = (u128)t[N] + c; t[N] = (uint64_t)x; t[N + 1] = (uint64_t)(x >> 64);
      uint64_t m = t[0] * P::INV;
      x = (u128)m * P::mod()[0] + t[0]; c = (uint64_t)(x >> 64);
      for (int j = 1; j < N; j++) { x = (u128)m * P::mod()[j] + t[j] + c; t[j - 1] = (uint64_t)x; c = (uint64_t)(x >> 64); }
      x = (u128)t[N] + c; t[N - 1] = (uint64_t)x; t[N] = t[N + 1] + (uint64_t)(x >> 64);
    }
    Fp r; memcpy(r.v, t, sizeof r.v);
    if (t[N] || geq_mod(r.v)) sub_mod(r.v);
    return r;
  }
  Fp sqr() const { return *this * *this; }
  Fp pow(const uint64_t* e, int nlimbs) const {
    Fp acc = one();
    for (int i = nlimbs * 64 - 1; i >= 0; i--) {
      acc = acc.sqr();
      if ((e[i >> 6] >> (i & 63)) & 1) acc = acc * *this;
    }
    return acc;
  }
  Fp inverse() const {
    uint64_t e[N]; memcpy(e, P::mod(), sizeof e);
    e[0] -= 2;  // p is odd and > 2: no borrow
    return pow(e, N);
  }
  Fp from_mont() const { Fp o = zero(); o.v[0] = 1; return *this * o; }
  Fp to_mont() const { return *this * from(P::r2()); }
};

template <class P>
struct Fp2 {
  typedef Fp<P> B;
  static constexpr int LIMBS = 2 * P::N;
  B c0, c1;
  static Fp2 zero() { return {B::zero(), B::zero()}; }
  static Fp2 one() { return {B::one(), B::zero()}; }
  static Fp2 from(const uint64_t* p) { return {B::from(p), B::from(p + P::N)}; }
  void to(uint64_t* p) const { c0.to(p); c1.to(p + P::N); }
  bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
  bool operator==(const Fp2& o) const { return c0 == o.c0 && c1 == o.c1; }
  bool operator!=(const Fp2& o) const { return !(*this == o); }
  Fp2 operator+(const Fp2& o) const { return {c0 + o.c0, c1 + o.c1}; }
  Fp2 operator-(const Fp2& o) const { return {c0 - o.c0, c1 - o.c1}; }
  Fp2 neg() const { return {c0.neg(), c1.neg()}; }
  Fp2 dbl() const { return {c0.dbl(), c1.dbl()}; }
  Fp2 operator*(const Fp2& o) const {  // u^2 = -1
    B t0 = c0 * o.c0, t1 = c1 * o.c1, t2 = (c0 + c1) * (o.c0 + o.c1);
    return {t0 - t1, t2 - t0 - t1};
  }
  Fp2 sqr() const { return *this * *this; }
  Fp2 inverse() const { B n = (c0.sqr() + c1.sqr()).inverse(); return {c0 * n, (c1 * n).neg()}; }
};

// ------------------------------------------------------------------------------------------
// Jacobian group (ark GroupProjective), a = 0
// ------------------------------------------------------------------------------------------
template <class F>
struct Aff { F x, y; bool inf; };

template <class F>
struct Jac {
  F x, y, z;
  static Jac zero() { return {F::zero(), F::one(), F::zero()}; }
  bool is_zero() const { return z.is_zero(); }

  void double_in_place() {  // dbl-2009-l
    if (is_zero()) return;
    F a = x.sqr(), b = y.sqr(), c = b.sqr();
    F d = ((x + b).sqr() - a - c).dbl();
    F e = a.dbl() + a;
    F f = e.sqr();
    z = (z * y).dbl();
    x = f - d.dbl();
    y = (d - x) * e - c.dbl().dbl().dbl();
  }
  void add_assign_mixed(const Aff<F>& o) {  // madd-2007-bl
    if (o.inf) return;
    if (is_zero()) { x = o.x; y = o.y; z = F::one(); return; }
    F z1z1 = z.sqr();
    F u2 = o.x * z1z1;
    F s2 = (o.y * z) * z1z1;
    if (x == u2 && y == s2) { double_in_place(); return; }
    F h = u2 - x;
    F hh = h.sqr();
    F i = hh.dbl().dbl();
    F j = h * i;
    F r = (s2 - y).dbl();
    F v = x * i;
    F x3 = r.sqr() - j - v.dbl();
    F y3 = r * (v - x3) - (y * j).dbl();
    F z3 = (z + h).sqr() - z1z1 - hh;
    x = x3; y = y3; z = z3;
  }
  void add_assign(const Jac& o) {  // add-2007-bl
    if (is_zero()) { *this = o; return; }
    if (o.is_zero()) return;
    F z1z1 = z.sqr(), z2z2 = o.z.sqr();
    F u1 = x * z2z2, u2 = o.x * z1z1;
    F s1 = y * o.z * z2z2, s2 = o.y * z * z1z1;
    if (u1 == u2 && s1 == s2) { double_in_place(); return; }
    F h = u2 - u1;
    F i = h.dbl().sqr();
    F j = h * i;
    F r = (s2 - s1).dbl();
    F v = u1 * i;
    F x3 = r.sqr() - j - v.dbl();
    F y3 = r * (v - x3) - (s1 * j).dbl();
    F z3 = ((z + o.z).sqr() - z1z1 - z2z2) * h;
    x = x3; y = y3; z = z3;
  }
  Aff<F> into_affine() const {
    if (is_zero()) return {F::zero(), F::zero(), true};
    F zi = z.inverse();
    F zi2 = zi.sqr();
    return {x * zi2, y * (zi2 * zi), false};
  }
};

// ------------------------------------------------------------------------------------------
// scalars: ark BigInteger256, canonical (non-Montgomery), 4 x u64 LE
// ------------------------------------------------------------------------------------------
struct Big256 { uint64_t v[4]; };
static inline bool big_is_zero(const Big256& s) { return (s.v[0] | s.v[1] | s.v[2] | s.v[3]) == 0; }
static inline bool big_is_one(const Big256& s) { return s.v[0] == 1 && (s.v[1] | s.v[2] | s.v[3]) == 0; }
// (s >> shift) % 2^c, c <= 32 (ark: divn(w_start) then as_ref()[0] % (1 << c))
static inline uint64_t big_window(const Big256& s, int shift, int c) {
  int limb = shift >> 6, off = shift & 63;
  uint64_t x = limb < 4 ? s.v[limb] >> off : 0;
  if (off && limb + 1 < 4) x |= s.v[limb + 1] << (64 - off);
  return x & ((1ull << c) - 1);
}

static int ark_window_bits(size_t size) {
  if (size < 32) return 3;
  int log2 = 0;
  while (((size_t)1 << log2) < size) log2++;
  return log2 * 69 / 100 + 2;
}

template <class F>
static std::vector<Aff<F>> load_bases(const uint64_t* bases, const uint8_t* inf, size_t n) {
  std::vector<Aff<F>> out(n);
  const int L = F::LIMBS;
  for (size_t i = 0; i < n; i++) {
    out[i].x = F::from(bases + i * 2 * L);
    out[i].y = F::from(bases + i * 2 * L + L);
    out[i].inf = inf ? ((inf[i >> 3] >> (i & 7)) & 1) : false;
  }
  return out;
}

template <class F>
static Jac<F> msm_ark(const Aff<F>* bases, const Big256* scalars, size_t size, int scalar_bits, int threads) {
  const int c = ark_window_bits(size);
  std::vector<int> starts;
  for (int w = 0; w < scalar_bits; w += c) starts.push_back(w);
  std::vector<Jac<F>> sums(starts.size());
  auto work = [&](size_t wi) {
    const int w_start = starts[wi];
    Jac<F> res = Jac<F>::zero();
    std::vector<Jac<F>> buckets(((size_t)1 << c) - 1, Jac<F>::zero());
    for (size_t i = 0; i < size; i++) {
      const Big256& s = scalars[i];
      if (big_is_zero(s)) continue;
      if (big_is_one(s)) {
        if (w_start == 0) res.add_assign_mixed(bases[i]);
      } else {
        uint64_t d = big_window(s, w_start, c);
        if (d != 0) buckets[d - 1].add_assign_mixed(bases[i]);
      }
    }
    Jac<F> running = Jac<F>::zero();
    for (size_t b = buckets.size(); b-- > 0;) {
      running.add_assign(buckets[b]);
      res.add_assign(running);
    }
    sums[wi] = res;
  };
  if (threads <= 1) {
    for (size_t wi = 0; wi < starts.size(); wi++) work(wi);
  } else {
    std::vector<std::thread> pool;
    size_t next = 0;
    // static round-robin of windows over `threads` workers (rayon would work-steal; the
    // windows are equal-cost for uniform scalars, so this is equivalent)
    for (int t = 0; t < threads; t++) {
      pool.emplace_back([&, t]() { for (size_t wi = t; wi < starts.size(); wi += threads) work(wi); });
    }
    (void)next;
    for (auto& th : pool) th.join();
  }
  Jac<F> lowest = sums[0];
  Jac<F> total = Jac<F>::zero();
  for (size_t wi = sums.size(); wi-- > 1;) {
    total.add_assign(sums[wi]);
    for (int k = 0; k < c; k++) total.double_in_place();
  }
  lowest.add_assign(total);
  return lowest;
}

template <class F>
static Jac<F> scalar_mul(const Aff<F>& p, const uint64_t* k, int nlimbs) {
  Jac<F> acc = Jac<F>::zero();
  for (int i = nlimbs * 64 - 1; i >= 0; i--) {
    acc.double_in_place();
    if ((k[i >> 6] >> (i & 63)) & 1) acc.add_assign_mixed(p);
  }
  return acc;
}

// Montgomery-trick batch normalisation of Jacobian points into packed affine limbs.
template <class F>
static void batch_to_affine(const std::vector<Jac<F>>& pts, uint64_t* out) {
  const int L = F::LIMBS;
  size_t n = pts.size();
  std::vector<F> pref(n);
  F acc = F::one();
  for (size_t i = 0; i < n; i++) { pref[i] = acc; if (!pts[i].is_zero()) acc = acc * pts[i].z; }
  F inv = acc.inverse();
  for (size_t i = n; i-- > 0;) {
    if (pts[i].is_zero()) { memset(out + i * 2 * L, 0, 2 * L * 8); continue; }
    F zi = inv * pref[i];
    inv = inv * pts[i].z;
    F zi2 = zi.sqr();
    (pts[i].x * zi2).to(out + i * 2 * L);
    (pts[i].y * (zi2 * zi)).to(out + i * 2 * L + L);
  }
}

template <class F, class C>
static Aff<F> generator() { return {F::from(C::gx()), F::from(C::gy()), false}; }

// ------------------------------------------------------------------------------------------
// NTT (ark Radix2EvaluationDomain, serial)
// ------------------------------------------------------------------------------------------
template <class P>
static void ntt_in_place(uint64_t* data, int log_n, bool inverse, bool coset) {
  typedef Fp<P> F;
  const size_t n = (size_t)1 << log_n;
  F* a = reinterpret_cast<F*>(data);
  // omega = ROOT^(2^(TWO_ADICITY - log_n))
  F w = F::from(inverse ? P::root_inv() : P::root());
  for (int i = 0; i < P::TWO_ADICITY - log_n; i++) w = w.sqr();
  if (coset && !inverse) {  // distribute powers of g
    F g = F::from(P::gen()), t = F::one();
    for (size_t i = 0; i < n; i++) { a[i] = a[i] * t; t = t * g; }
  }
  // bit reversal ("derange")
  for (size_t i = 1, j = 0; i < n; i++) {
    size_t bit = n >> 1;
    for (; j & bit; bit >>= 1) j ^= bit;
    j ^= bit;
    if (i < j) std::swap(a[i], a[j]);
  }
  std::vector<F> tw(n / 2 ? n / 2 : 1);
  for (size_t len = 2; len <= n; len <<= 1) {
    F wl = w;
    for (size_t k = len; k < n; k <<= 1) wl = wl.sqr();
    size_t half = len >> 1;
    tw[0] = F::one();
    for (size_t k = 1; k < half; k++) tw[k] = tw[k - 1] * wl;
    for (size_t s = 0; s < n; s += len)
      for (size_t k = 0; k < half; k++) {
        F u = a[s + k], v = a[s + k + half] * tw[k];
        a[s + k] = u + v;
        a[s + k + half] = u - v;
      }
  }
  if (inverse) {
    uint64_t nn[P::N]; memset(nn, 0, sizeof nn); nn[0] = n;
    F size_inv = F::from(nn).to_mont().inverse();
    if (coset) {
      F gi = F::from(P::gen_inv()), t = size_inv;
      for (size_t i = 0; i < n; i++) { a[i] = a[i] * t; t = t * gi; }
    } else {
      for (size_t i = 0; i < n; i++) a[i] = a[i] * size_inv;
    }
  }
}

// ------------------------------------------------------------------------------------------
// dispatch
// ------------------------------------------------------------------------------------------
typedef Fp<Bls12381Fq> Fq381;
typedef Fp2<Bls12381Fq> Fq381_2;
typedef Fp<Bn254Fq> Fq254;
typedef Fp2<Bn254Fq> Fq254_2;

#define DISPATCH_CURVE(curve, ...)                                    \
  switch (curve) {                                                    \
    case 0: { typedef Fq381 F; typedef Bls12381G1 C; __VA_ARGS__; break; }   \
    case 1: { typedef Fq381_2 F; typedef Bls12381G2 C; __VA_ARGS__; break; } \
    case 2: { typedef Fq254 F; typedef Bn254G1 C; __VA_ARGS__; break; }      \
    case 3: { typedef Fq254_2 F; typedef Bn254G2 C; __VA_ARGS__; break; }    \
    default: return -1;                                               \
  }

static int scalar_bits_of(int curve) { return curve < 2 ? 255 : 254; }

template <class F>
static void jac_out(const Jac<F>& j, uint64_t* out) {
  j.x.to(out); j.y.to(out + F::LIMBS); j.z.to(out + 2 * F::LIMBS);
}

extern "C" {

int oracle_window_bits(size_t n) { return ark_window_bits(n); }

// bases: n packed affine points (x||y, Montgomery u64 limbs); inf: optional bitset;
// scalars: n x 4 u64 canonical; out: Jacobian X||Y||Z Montgomery.
int oracle_msm(int curve, const uint64_t* bases, const uint8_t* inf, const uint64_t* scalars, size_t n,
               int threads, uint64_t* out_jac) {
  DISPATCH_CURVE(curve, {
    auto b = load_bases<F>(bases, inf, n);
    Jac<F> r = msm_ark<F>(b.data(), reinterpret_cast<const Big256*>(scalars), n, scalar_bits_of(curve), threads);
    jac_out(r, out_jac);
  });
  return 0;
}

// Jacobian (Montgomery limbs) -> affine x||y (Montgomery limbs); *is_inf = 1 for identity.
int oracle_to_affine(int curve, const uint64_t* jac, uint64_t* out_affine, int* is_inf) {
  DISPATCH_CURVE(curve, {
    Jac<F> j = {F::from(jac), F::from(jac + F::LIMBS), F::from(jac + 2 * F::LIMBS)};
    Aff<F> a = j.into_affine();
    *is_inf = a.inf;
    a.x.to(out_affine); a.y.to(out_affine + F::LIMBS);
  });
  return 0;
}

// out = [k]G as Jacobian; k = 4 u64 limbs canonical
int oracle_gen_mul(int curve, const uint64_t* k, uint64_t* out_jac) {
  DISPATCH_CURVE(curve, {
    Jac<F> r = scalar_mul<F>(generator<F, C>(), k, 4);
    jac_out(r, out_jac);
  });
  return 0;
}

// P_i = [start + i]G, i in [0, n), packed affine Montgomery (start + i >= 1)
int oracle_bases_seq(int curve, uint64_t start, size_t n, uint64_t* out) {
  DISPATCH_CURVE(curve, {
    Aff<F> g = generator<F, C>();
    uint64_t k[4] = {start, 0, 0, 0};
    Jac<F> cur = scalar_mul<F>(g, k, 1);
    const size_t CH = 4096;
    std::vector<Jac<F>> buf;
    for (size_t off = 0; off < n; off += CH) {
      size_t m = std::min(CH, n - off);
      buf.clear();
      for (size_t i = 0; i < m; i++) { buf.push_back(cur); cur.add_assign_mixed(g); }
      batch_to_affine<F>(buf, out + off * 2 * F::LIMBS);
    }
  });
  return 0;
}

// P_i = [d_i]G for 64-bit d_i
int oracle_bases_from_dlogs(int curve, const uint64_t* dlogs, size_t n, uint64_t* out) {
  DISPATCH_CURVE(curve, {
    Aff<F> g = generator<F, C>();
    std::vector<Jac<F>> buf(n);
    for (size_t i = 0; i < n; i++) buf[i] = scalar_mul<F>(g, dlogs + i, 1);
    batch_to_affine<F>(buf, out);
  });
  return 0;
}

// sum_i s_i * d_i mod r (s canonical 4 limbs, d 64-bit); field: 0 = BN254 Fr, 1 = BLS12-381 Fr
int oracle_dot_mod_r(int field, const uint64_t* scalars, const uint64_t* dlogs, size_t n, uint64_t* out) {
  auto run = [&](auto tag) {
    typedef decltype(tag) P;
    typedef Fp<P> F;
    F acc = F::zero();
    for (size_t i = 0; i < n; i++) {
      uint64_t d[4] = {dlogs[i], 0, 0, 0};
      // mont(s) * d (d canonical) = mont(s * d)
      F s = F::from(scalars + 4 * i).to_mont();
      acc = acc + s * F::from(d).to_mont();
    }
    acc.from_mont().to(out);
  };
  if (field == 0) run(Bn254Fr{}); else if (field == 1) run(Bls12381Fr{}); else return -1;
  return 0;
}

// in-place NTT over 2^log_n Montgomery elements, natural order in and out (ark semantics)
int oracle_ntt(int field, uint64_t* data, int log_n, int inverse, int coset) {
  if (field == 0) { if (log_n > Bn254Fr::TWO_ADICITY) return -2; ntt_in_place<Bn254Fr>(data, log_n, inverse, coset); }
  else if (field == 1) { if (log_n > Bls12381Fr::TWO_ADICITY) return -2; ntt_in_place<Bls12381Fr>(data, log_n, inverse, coset); }
  else return -1;
  return 0;
}

// elementwise Montgomery product (validates the 64-bit field code against the Python oracle)
int oracle_fp_mul(int field_id, const uint64_t* a, const uint64_t* b, uint64_t* out) {
  switch (field_id) {
    case 0: (Fq381::from(a) * Fq381::from(b)).to(out); break;
    case 1: (Fp<Bls12381Fr>::from(a) * Fp<Bls12381Fr>::from(b)).to(out); break;
    case 2: (Fq254::from(a) * Fq254::from(b)).to(out); break;
    case 3: (Fp<Bn254Fr>::from(a) * Fp<Bn254Fr>::from(b)).to(out); break;
    default: return -1;
  }
  return 0;
}

int oracle_hw_threads(void) { return (int)std::thread::hardware_concurrency(); }

}  // extern "C"
