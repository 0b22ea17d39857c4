// Prime-field arithmetic in Montgomery form, 32-bit limbs held in registers.
//
// Replaces ark_ff::Fp256 / Fp384 (ark-ff 0.3.0; re-exported through `pub mod ff` at
// /root/reference/plugins/arkworks/src/lib.rs:101-103).  Same value domain as ark: every
// element is the fully reduced Montgomery residue a*R mod p, R = 2^(32*N) = 2^(64*limbs64),
// so the little-endian u64 limbs ark keeps in memory can be loaded unchanged as 2x u32.
//
// mul() is a coarsely-integrated operand-scanning Montgomery product built from two
// interleaved accumulators ("aligned" and "offset by one limb"), so that every 32x32->64
// partial product lands on an even/odd register pair and the whole row is one carry chain:
// 4N+4 fma-pipe instructions per row, no carry fix-ups.  See DESIGN.md section "Field mul".
#pragma once
#include "ptx.cuh"

namespace ozl {

template <class P>
struct Fp {
  static constexpr int N = P::N;
  typedef P Params;
  uint32_t v[N];

  // ---- constants ---------------------------------------------------------------------------
  static OZL_DEV Fp zero() {
    Fp r;
#pragma unroll
    for (int i = 0; i < N; i++) r.v[i] = 0;
    return r;
  }
  static OZL_DEV Fp one() {  // R mod p
    Fp r;
#pragma unroll
    for (int i = 0; i < N; i++) r.v[i] = P::one()[i];
    return r;
  }
  static OZL_DEV Fp r2() {
    Fp r;
#pragma unroll
    for (int i = 0; i < N; i++) r.v[i] = P::r2()[i];
    return r;
  }
  static OZL_DEV Fp from_limbs(const uint32_t* c) {
    Fp r;
#pragma unroll
    for (int i = 0; i < N; i++) r.v[i] = c[i];
    return r;
  }

  // ---- predicates --------------------------------------------------------------------------
  OZL_DEV bool is_zero() const {
    uint32_t t = v[0];
#pragma unroll
    for (int i = 1; i < N; i++) t |= v[i];
    return t == 0;
  }
  friend OZL_DEV bool operator==(const Fp& a, const Fp& b) {
    uint32_t t = a.v[0] ^ b.v[0];
#pragma unroll
    for (int i = 1; i < N; i++) t |= a.v[i] ^ b.v[i];
    return t == 0;
  }
  friend OZL_DEV bool operator!=(const Fp& a, const Fp& b) { return !(a == b); }

  // ---- add / sub ---------------------------------------------------------------------------
  // r = (x >= p) ? x - p : x, for x < 2p
  static OZL_DEV void final_sub(uint32_t* x) {
    uint32_t t[N];
    t[0] = ptx::sub_cc(x[0], P::mod()[0]);
#pragma unroll
    for (int i = 1; i < N; i++) t[i] = ptx::subc_cc(x[i], P::mod()[i]);
    uint32_t borrow = ptx::subc(0, 0);  // 0xffffffff iff x < p
#pragma unroll
    for (int i = 0; i < N; i++) x[i] = borrow ? x[i] : t[i];
  }

  friend OZL_DEV Fp operator+(const Fp& a, const Fp& b) {
    Fp r;
    r.v[0] = ptx::add_cc(a.v[0], b.v[0]);
#pragma unroll
    for (int i = 1; i < N - 1; i++) r.v[i] = ptx::addc_cc(a.v[i], b.v[i]);
    r.v[N - 1] = ptx::addc(a.v[N - 1], b.v[N - 1]);  // a + b < 2p < 2^(32N): no carry out
    final_sub(r.v);
    return r;
  }

  friend OZL_DEV Fp operator-(const Fp& a, const Fp& b) {
    Fp r;
    r.v[0] = ptx::sub_cc(a.v[0], b.v[0]);
#pragma unroll
    for (int i = 1; i < N; i++) r.v[i] = ptx::subc_cc(a.v[i], b.v[i]);
    uint32_t mask = ptx::subc(0, 0);  // all ones iff a < b
    r.v[0] = ptx::add_cc(r.v[0], P::mod()[0] & mask);
#pragma unroll
    for (int i = 1; i < N - 1; i++) r.v[i] = ptx::addc_cc(r.v[i], P::mod()[i] & mask);
    r.v[N - 1] = ptx::addc(r.v[N - 1], P::mod()[N - 1] & mask);
    return r;
  }

  OZL_DEV Fp neg() const {
    if (is_zero()) return *this;
    Fp r;
    r.v[0] = ptx::sub_cc(P::mod()[0], v[0]);
#pragma unroll
    for (int i = 1; i < N - 1; i++) r.v[i] = ptx::subc_cc(P::mod()[i], v[i]);
    r.v[N - 1] = ptx::subc(P::mod()[N - 1], v[N - 1]);
    return r;
  }
  // conditional negate without divergence on the flag
  OZL_DEV Fp cneg(bool flag) const {
    Fp n = neg();
    Fp r;
#pragma unroll
    for (int i = 0; i < N; i++) r.v[i] = flag ? n.v[i] : v[i];
    return r;
  }

  OZL_DEV Fp dbl() const { return *this + *this; }

  // ---- Montgomery product ------------------------------------------------------------------
  // Row helpers.  X is the accumulator whose limb k sits at weight 2^(32k) ("aligned"),
  // Y the one whose limb k sits at weight 2^(32(k+1)) ("offset").
  static OZL_DEV void reduce_row(uint32_t* X, uint32_t* Y) {
    const uint32_t m = ptx::mul_lo(X[0], P::INV);
    Y[0] = ptx::mad_lo_cc(P::mod()[1], m, Y[0]);
    Y[1] = ptx::madc_hi_cc(P::mod()[1], m, Y[1]);
#pragma unroll
    for (int k = 2; k < N; k += 2) {
      Y[k] = ptx::madc_lo_cc(P::mod()[k + 1], m, Y[k]);
      Y[k + 1] = ptx::madc_hi_cc(P::mod()[k + 1], m, Y[k + 1]);
    }
    X[0] = ptx::mad_lo_cc(P::mod()[0], m, X[0]);
    X[1] = ptx::madc_hi_cc(P::mod()[0], m, X[1]);
#pragma unroll
    for (int j = 2; j < N; j += 2) {
      X[j] = ptx::madc_lo_cc(P::mod()[j], m, X[j]);
      X[j + 1] = ptx::madc_hi_cc(P::mod()[j], m, X[j + 1]);
    }
    Y[N - 1] = ptx::addc(Y[N - 1], 0);
  }

  // Divide the running value by 2^32 (X[0] == 0 on entry) and add a * bi.
  // On return the roles are swapped: Y is aligned, X is offset.
  static OZL_DEV void next_row(uint32_t* X, uint32_t* Y, const uint32_t* a, uint32_t bi) {
    Y[0] = ptx::add_cc(Y[0], X[1]);
#pragma unroll
    for (int k = 0; k < N; k += 2) {
      X[k] = ptx::madc_lo_cc(a[k + 1], bi, (k + 2 < N) ? X[k + 2] : 0u);
      X[k + 1] = ptx::madc_hi_cc(a[k + 1], bi, (k + 3 < N) ? X[k + 3] : 0u);
    }
    Y[0] = ptx::mad_lo_cc(a[0], bi, Y[0]);
    Y[1] = ptx::madc_hi_cc(a[0], bi, Y[1]);
#pragma unroll
    for (int j = 2; j < N; j += 2) {
      Y[j] = ptx::madc_lo_cc(a[j], bi, Y[j]);
      Y[j + 1] = ptx::madc_hi_cc(a[j], bi, Y[j + 1]);
    }
    X[N - 1] = ptx::addc(X[N - 1], 0);
  }

  friend OZL_DEV Fp operator*(const Fp& a, const Fp& b) {
    static_assert(N % 2 == 0, "even limb count required");
    uint32_t A[N], B[N];
    // row 0: pure products
    {
      const uint32_t b0 = b.v[0];
#pragma unroll
      for (int j = 0; j < N; j += 2) {
        A[j] = ptx::mul_lo(a.v[j], b0);
        A[j + 1] = ptx::mul_hi(a.v[j], b0);
        B[j] = ptx::mul_lo(a.v[j + 1], b0);
        B[j + 1] = ptx::mul_hi(a.v[j + 1], b0);
      }
      reduce_row(A, B);
    }
#pragma unroll
    for (int i = 1; i < N - 1; i += 2) {
      next_row(A, B, a.v, b.v[i]);
      reduce_row(B, A);
      next_row(B, A, a.v, b.v[i + 1]);
      reduce_row(A, B);
    }
    next_row(A, B, a.v, b.v[N - 1]);
    reduce_row(B, A);
    // value = A + (B >> 32), B[0] == 0
    Fp r;
    r.v[0] = ptx::add_cc(A[0], B[1]);
#pragma unroll
    for (int k = 1; k < N - 1; k++) r.v[k] = ptx::addc_cc(A[k], B[k + 1]);
    r.v[N - 1] = ptx::addc(A[N - 1], 0);
    final_sub(r.v);
    return r;
  }

  // Same shape as reduce_row with an arbitrary multiplicand c and multiplier d: adds c * d to the running
  // value held as (X aligned, Y offset).  Used by mul_add2 to put a second product into a CIOS row.
  static OZL_DEV void add_row(uint32_t* X, uint32_t* Y, const uint32_t* c, uint32_t d) {
    Y[0] = ptx::mad_lo_cc(c[1], d, Y[0]);
    Y[1] = ptx::madc_hi_cc(c[1], d, Y[1]);
#pragma unroll
    for (int k = 2; k < N; k += 2) {
      Y[k] = ptx::madc_lo_cc(c[k + 1], d, Y[k]);
      Y[k + 1] = ptx::madc_hi_cc(c[k + 1], d, Y[k + 1]);
    }
    X[0] = ptx::mad_lo_cc(c[0], d, X[0]);
    X[1] = ptx::madc_hi_cc(c[0], d, X[1]);
#pragma unroll
    for (int j = 2; j < N; j += 2) {
      X[j] = ptx::madc_lo_cc(c[j], d, X[j]);
      X[j + 1] = ptx::madc_hi_cc(c[j], d, X[j + 1]);
    }
    Y[N - 1] = ptx::addc(Y[N - 1], 0);
  }

  // (a * b + c * d) / R mod p with ONE Montgomery reduction: every CIOS row adds a * b_i, c * d_i and m * p.
  // 3 N^2 wide multiplies instead of the 4 N^2 of two products.  Needs 3 p (1 + 2^32) < 2^(32 (N + 1)), i.e.
  // p < 2^(32 N) / 3, so that the running value still fits the two accumulators (true for both base fields:
  // R / p = 9.6 for BLS12-381, 5.3 for BN254); the result is < p (2 p / R + 1) < 2 p, one final subtraction.
  // a * b - c * d is mul_add2(a, b, c.neg(), d).
  static OZL_DEV Fp mul_add2(const Fp& a, const Fp& b, const Fp& c, const Fp& d) {
    static_assert(N % 2 == 0, "even limb count required");
    static_assert(P::BITS <= 32 * N - 2, "mul_add2 needs p < R / 4 (not true for the BLS12-381 scalar field)");
    uint32_t A[N], B[N];
    {
      const uint32_t b0 = b.v[0];
#pragma unroll
      for (int j = 0; j < N; j += 2) {
        A[j] = ptx::mul_lo(a.v[j], b0);
        A[j + 1] = ptx::mul_hi(a.v[j], b0);
        B[j] = ptx::mul_lo(a.v[j + 1], b0);
        B[j + 1] = ptx::mul_hi(a.v[j + 1], b0);
      }
      add_row(A, B, c.v, d.v[0]);
      reduce_row(A, B);
    }
#pragma unroll
    for (int i = 1; i < N - 1; i += 2) {
      next_row(A, B, a.v, b.v[i]);
      add_row(B, A, c.v, d.v[i]);
      reduce_row(B, A);
      next_row(B, A, a.v, b.v[i + 1]);
      add_row(A, B, c.v, d.v[i + 1]);
      reduce_row(A, B);
    }
    next_row(A, B, a.v, b.v[N - 1]);
    add_row(B, A, c.v, d.v[N - 1]);
    reduce_row(B, A);
    Fp r;
    r.v[0] = ptx::add_cc(A[0], B[1]);
#pragma unroll
    for (int k = 1; k < N - 1; k++) r.v[k] = ptx::addc_cc(A[k], B[k + 1]);
    r.v[N - 1] = ptx::addc(A[N - 1], 0);
    final_sub(r.v);
    return r;
  }
  static OZL_DEV_NOINLINE Fp mul_add2_ni(Fp a, Fp b, Fp c, Fp d) { return mul_add2(a, b, c, d); }

  OZL_DEV Fp sqr() const { return *this * *this; }

  // ---- Montgomery squaring (separated operand scanning) --------------------------------------
  // Separated operand scanning: t = a^2 as 2 * (off-diagonal products) + diagonal, then N
  // reduction rows.  N(N-1)/2 + N + N^2 wide multiplies instead of 2 N^2 (222 vs 288 for N = 12);
  // the extra work is carry bookkeeping on the otherwise idle ALU pipe.
  //
  // Off-diagonal row i adds a_i * a_j (j > i) at limb i + j.  The products with j - i odd and those
  // with j - i even each form one carry chain over disjoint limb pairs.  The chain that stops one
  // limb lower runs first; its carry lands in limb i + N, which so far holds at most a carry bit,
  // and the carry of the second chain lands in limb i + N + 1, which is still zero.
  //
  // Measured on B200 inside k_accumulate: neutral for 8 limbs and 9 % SLOWER for 12 limbs (the 2N-limb
  // intermediate costs registers and the carry bookkeeping lengthens the dependent chains), so the hot
  // path keeps sqr() = mul(); this variant stays available and tested (tests/test_host_emu.py).
  OZL_DEV Fp sqr_sos() const {
    const uint32_t* a = v;
    uint32_t t[2 * N];
#pragma unroll
    for (int k = 0; k < 2 * N; k++) t[k] = 0;
#pragma unroll
    for (int i = 0; i < N - 1; i++) {
      // chain "x": j = i+1, i+3, ... ; chain "y": j = i+2, i+4, ...   (N even: x reaches j = N-1 iff i even)
      const int lower_start = (i % 2 == 0) ? i + 2 : i + 1;
      const int higher_start = (i % 2 == 0) ? i + 1 : i + 2;
#pragma unroll
      for (int pass = 0; pass < 2; pass++) {
        const int j0 = pass == 0 ? lower_start : higher_start;
        if (j0 < N) {
          int last = j0;
#pragma unroll
          for (int j = j0; j < N; j += 2) {
            const int s = i + j;
            t[s] = (j == j0) ? ptx::mad_lo_cc(a[i], a[j], t[s]) : ptx::madc_lo_cc(a[i], a[j], t[s]);
            t[s + 1] = ptx::madc_hi_cc(a[i], a[j], t[s + 1]);
            last = s + 1;
          }
          t[last + 1] = ptx::addc(t[last + 1], 0);
        }
      }
    }
    // t = 2 t  (2 T < a^2 < 2^(64N): no carry out)
    t[0] = ptx::add_cc(t[0], t[0]);
#pragma unroll
    for (int k = 1; k < 2 * N - 1; k++) t[k] = ptx::addc_cc(t[k], t[k]);
    t[2 * N - 1] = ptx::addc(t[2 * N - 1], t[2 * N - 1]);
    // t += sum a_i^2 2^(64 i)
#pragma unroll
    for (int i = 0; i < N; i++) {
      t[2 * i] = (i == 0) ? ptx::mad_lo_cc(a[i], a[i], t[2 * i]) : ptx::madc_lo_cc(a[i], a[i], t[2 * i]);
      t[2 * i + 1] = (i == N - 1) ? ptx::madc_hi(a[i], a[i], t[2 * i + 1]) : ptx::madc_hi_cc(a[i], a[i], t[2 * i + 1]);
    }
    // Montgomery reduction: row i clears limb i.  Carries out of a chain target limbs >= N, which never
    // feed a later multiplier m, so they are parked in cr[] (cr[k] -> limb N + k) and added once at the end.
    uint32_t cr[N + 1];
#pragma unroll
    for (int k = 0; k <= N; k++) cr[k] = 0;
#pragma unroll
    for (int i = 0; i < N; i++) {
      const uint32_t m = ptx::mul_lo(t[i], P::INV);
      t[i] = ptx::mad_lo_cc(P::mod()[0], m, t[i]);
      t[i + 1] = ptx::madc_hi_cc(P::mod()[0], m, t[i + 1]);
#pragma unroll
      for (int j = 2; j < N; j += 2) {
        t[i + j] = ptx::madc_lo_cc(P::mod()[j], m, t[i + j]);
        t[i + j + 1] = ptx::madc_hi_cc(P::mod()[j], m, t[i + j + 1]);
      }
      cr[i] = ptx::addc(cr[i], 0);          // even chain stops at limb i + N - 1
      t[i + 1] = ptx::mad_lo_cc(P::mod()[1], m, t[i + 1]);
      t[i + 2] = ptx::madc_hi_cc(P::mod()[1], m, t[i + 2]);
#pragma unroll
      for (int j = 3; j < N; j += 2) {
        t[i + j] = ptx::madc_lo_cc(P::mod()[j], m, t[i + j]);
        t[i + j + 1] = ptx::madc_hi_cc(P::mod()[j], m, t[i + j + 1]);
      }
      cr[i + 1] = ptx::addc(cr[i + 1], 0);  // odd chain stops at limb i + N
    }
    Fp r;
    r.v[0] = ptx::add_cc(t[N], cr[0]);
#pragma unroll
    for (int k = 1; k < N - 1; k++) r.v[k] = ptx::addc_cc(t[N + k], cr[k]);
    r.v[N - 1] = ptx::addc(t[2 * N - 1], cr[N - 1]);
    final_sub(r.v);
    return r;
  }

  // ---- unreduced product and stand-alone reduction for lazily reduced Fq2 arithmetic -----------------
  // t[0 .. 2N) = a * b (plain integers of N limbs, e.g. unreduced sums a0 + a1 < 2p), NO reduction.
  // Two accumulators like the CIOS rows above so that every mad.lo / madc.hi pair lands on an even-aligned register
  // pair and ptxas fuses it into one IMAD.WIDE: product a_j b_i goes to X[i + j] when i + j is even and to
  // Y[i + j - 1] (Y limb k has weight 2^(32 (k + 1))) when it is odd; t = X + (Y << 32).
  static OZL_DEV void mul_wide(const uint32_t* a, const uint32_t* b, uint32_t* t) {
    uint32_t X[2 * N], Y[2 * N];
#pragma unroll
    for (int k = 0; k < 2 * N; k++) { X[k] = 0; Y[k] = 0; }
#pragma unroll
    for (int i = 0; i < N; i++) {
      const int px = i & 1;                   // first j of the aligned chain
#pragma unroll
      for (int j = px; j < N; j += 2) {
        const int sx = i + j;
        X[sx] = (j == px) ? ptx::mad_lo_cc(a[j], b[i], X[sx]) : ptx::madc_lo_cc(a[j], b[i], X[sx]);
        X[sx + 1] = ptx::madc_hi_cc(a[j], b[i], X[sx + 1]);
      }
      // the chain ends at limb i + px + N - 1; the limb above holds at most one carry bit so far (no product of an
      // earlier row reaches it), and for the last row there is none: X <= a b < 2^(64 N)
      if (i + px + N < 2 * N) X[i + px + N] = ptx::addc(X[i + px + N], 0);
      const int py = 1 - px;                  // first j of the offset chain
#pragma unroll
      for (int j = py; j < N; j += 2) {
        const int sy = i + j - 1;
        Y[sy] = (j == py) ? ptx::mad_lo_cc(a[j], b[i], Y[sy]) : ptx::madc_lo_cc(a[j], b[i], Y[sy]);
        Y[sy + 1] = ptx::madc_hi_cc(a[j], b[i], Y[sy + 1]);
      }
      if (i + N - px < 2 * N) Y[i + N - px] = ptx::addc(Y[i + N - px], 0);
    }
    t[0] = X[0];
    t[1] = ptx::add_cc(X[1], Y[0]);
#pragma unroll
    for (int k = 2; k < 2 * N - 1; k++) t[k] = ptx::addc_cc(X[k], Y[k - 1]);
    t[2 * N - 1] = ptx::addc(X[2 * N - 1], Y[2 * N - 2]);
  }

  // Divide the running value (X aligned with X[0] == 0, Y offset) by 2^32 and add `hi` at limb N - 1; on return the
  // roles are swapped (Y aligned, X offset).  This is next_row() with the product row replaced by one injected limb.
  static OZL_DEV void shift_inject(uint32_t* X, uint32_t* Y, uint32_t hi) {
    Y[0] = ptx::add_cc(Y[0], X[1]);
#pragma unroll
    for (int k = 0; k < N - 2; k++) X[k] = ptx::addc_cc(X[k + 2], 0);
    X[N - 2] = ptx::addc_cc(hi, 0);
    X[N - 1] = ptx::addc(0, 0);
  }

  // Montgomery reduction of a 2N-limb value t < p 2^(32 N): t / 2^(32 N) mod p, fully reduced.  Same rows as the fused
  // multiplication (N^2 aligned wide multiplies); the high limbs of t enter one per row.
  static OZL_DEV Fp redc_cios(const uint32_t* t) {
    static_assert(N % 2 == 0, "even limb count required");
    uint32_t A[N], B[N];
#pragma unroll
    for (int k = 0; k < N; k++) { A[k] = t[k]; B[k] = 0; }
#pragma unroll
    for (int i = 0; i < N; i += 2) {
      reduce_row(A, B);
      shift_inject(A, B, t[N + i]);
      reduce_row(B, A);
      shift_inject(B, A, t[N + i + 1]);
    }
    // A aligned, B offset; the value is < 2 p < 2^(32 N), so B[N - 1] == 0
    Fp r;
    r.v[0] = A[0];
    r.v[1] = ptx::add_cc(A[1], B[0]);
#pragma unroll
    for (int k = 2; k < N - 1; k++) r.v[k] = ptx::addc_cc(A[k], B[k - 1]);
    r.v[N - 1] = ptx::addc(A[N - 1], B[N - 2]);
    final_sub(r.v);
    return r;
  }
  // x += y, x -= y over M limbs (no reduction; the callers' bounds exclude carries / borrows out)
  template <int M>
  static OZL_DEV void add_limbs(uint32_t* x, const uint32_t* y) {
    x[0] = ptx::add_cc(x[0], y[0]);
#pragma unroll
    for (int k = 1; k < M - 1; k++) x[k] = ptx::addc_cc(x[k], y[k]);
    x[M - 1] = ptx::addc(x[M - 1], y[M - 1]);
  }
  template <int M>
  static OZL_DEV void sub_limbs(uint32_t* x, const uint32_t* y) {
    x[0] = ptx::sub_cc(x[0], y[0]);
#pragma unroll
    for (int k = 1; k < M - 1; k++) x[k] = ptx::subc_cc(x[k], y[k]);
    x[M - 1] = ptx::subc(x[M - 1], y[M - 1]);
  }

  // ---- Karatsuba product + separated Montgomery reduction -------------------------------------
  // t[0 .. 2H) = x * y for H-limb operands, schoolbook: row i adds x_j * y_i at limb i + j as two carry
  // chains (even j, odd j) over disjoint limb pairs, like the off-diagonal rows of sqr_sos().
  template <int H>
  static OZL_DEV void mul_half(const uint32_t* x, const uint32_t* y, uint32_t* t) {
#pragma unroll
    for (int k = 0; k < 2 * H; k++) t[k] = 0;
#pragma unroll
    for (int i = 0; i < H; i++) {
      // even chain: j = 0, 2, ..: limbs (i + j, i + j + 1); stops at limb i + H - 1 (H even), carry into limb i + H
#pragma unroll
      for (int j = 0; j < H; j += 2) {
        t[i + j] = (j == 0) ? ptx::mad_lo_cc(x[j], y[i], t[i + j]) : ptx::madc_lo_cc(x[j], y[i], t[i + j]);
        t[i + j + 1] = ptx::madc_hi_cc(x[j], y[i], t[i + j + 1]);
      }
      t[i + H] = ptx::addc(t[i + H], 0);
      // odd chain: j = 1, 3, ..: limbs (i + j, i + j + 1); ends at limb i + H, whose carry out is impossible
      // for the last row pair only through limb i + H + 1 (zero or a carry bit so far)
#pragma unroll
      for (int j = 1; j < H; j += 2) {
        t[i + j] = (j == 1) ? ptx::mad_lo_cc(x[j], y[i], t[i + j]) : ptx::madc_lo_cc(x[j], y[i], t[i + j]);
        t[i + j + 1] = ptx::madc_hi_cc(x[j], y[i], t[i + j + 1]);
      }
      if (i + H + 1 < 2 * H) t[i + H + 1] = ptx::addc(t[i + H + 1], 0);
    }
  }

  // Montgomery reduction of a 2N-limb value t < p * 2^(32N): returns t / 2^(32N) mod p, fully reduced.
  // Row i clears limb i; carries out of a chain target limbs >= N, which never feed a later multiplier m,
  // so they are parked in cr[] (cr[k] -> limb N + k) and added once at the end.
  static OZL_DEV Fp redc_wide(uint32_t* t) {
    uint32_t cr[N + 1];
#pragma unroll
    for (int k = 0; k <= N; k++) cr[k] = 0;
#pragma unroll
    for (int i = 0; i < N; i++) {
      const uint32_t m = ptx::mul_lo(t[i], P::INV);
      t[i] = ptx::mad_lo_cc(P::mod()[0], m, t[i]);
      t[i + 1] = ptx::madc_hi_cc(P::mod()[0], m, t[i + 1]);
#pragma unroll
      for (int j = 2; j < N; j += 2) {
        t[i + j] = ptx::madc_lo_cc(P::mod()[j], m, t[i + j]);
        t[i + j + 1] = ptx::madc_hi_cc(P::mod()[j], m, t[i + j + 1]);
      }
      cr[i] = ptx::addc(cr[i], 0);          // even chain stops at limb i + N - 1
      t[i + 1] = ptx::mad_lo_cc(P::mod()[1], m, t[i + 1]);
      t[i + 2] = ptx::madc_hi_cc(P::mod()[1], m, t[i + 2]);
#pragma unroll
      for (int j = 3; j < N; j += 2) {
        t[i + j] = ptx::madc_lo_cc(P::mod()[j], m, t[i + j]);
        t[i + j + 1] = ptx::madc_hi_cc(P::mod()[j], m, t[i + j + 1]);
      }
      cr[i + 1] = ptx::addc(cr[i + 1], 0);  // odd chain stops at limb i + N
    }
    Fp r;
    r.v[0] = ptx::add_cc(t[N], cr[0]);
#pragma unroll
    for (int k = 1; k < N - 1; k++) r.v[k] = ptx::addc_cc(t[N + k], cr[k]);
    r.v[N - 1] = ptx::addc(t[2 * N - 1], cr[N - 1]);
    final_sub(r.v);
    return r;
  }

  // t[0 .. 2N) = a * b by one level of subtractive Karatsuba over halves of H = N / 2 limbs:
  //   a b = z0 + (z0 + z2 + (a_lo - a_hi)(b_hi - b_lo)) 2^(32H) + z2 2^(64H),  z0 = a_lo b_lo, z2 = a_hi b_hi
  // 3 H^2 wide multiplies instead of 4 H^2 (108 instead of 144 for N = 12); the differences, the signed
  // middle term and the recombination are additions on the otherwise idle ALU pipe.
  static OZL_DEV void mul_wide_kara(const Fp& a, const Fp& b, uint32_t* t) {
    constexpr int H = N / 2;
    static_assert(H % 2 == 0, "half length must be even");
    uint32_t da[H], db[H], z1[2 * H + 1], m[2 * H];
    // da = |a_lo - a_hi|, db = |b_hi - b_lo|, signs as all-ones masks
    da[0] = ptx::sub_cc(a.v[0], a.v[H]);
#pragma unroll
    for (int i = 1; i < H; i++) da[i] = ptx::subc_cc(a.v[i], a.v[H + i]);
    const uint32_t sa = ptx::subc(0, 0);
    db[0] = ptx::sub_cc(b.v[H], b.v[0]);
#pragma unroll
    for (int i = 1; i < H; i++) db[i] = ptx::subc_cc(b.v[H + i], b.v[i]);
    const uint32_t sb = ptx::subc(0, 0);
    // conditional negate: (x xor s) - s
    da[0] = ptx::sub_cc(da[0] ^ sa, sa);
#pragma unroll
    for (int i = 1; i < H; i++) da[i] = ptx::subc_cc(da[i] ^ sa, sa);
    db[0] = ptx::sub_cc(db[0] ^ sb, sb);
#pragma unroll
    for (int i = 1; i < H; i++) db[i] = ptx::subc_cc(db[i] ^ sb, sb);
    mul_half<H>(a.v, b.v, t);                 // z0 -> t[0 .. 2H)
    mul_half<H>(a.v + H, b.v + H, t + 2 * H); // z2 -> t[2H .. 4H)
    mul_half<H>(da, db, m);
    const uint32_t sm = sa ^ sb;              // all ones: the middle product enters negatively
    // z1 = z0 + z2 + (+/-) m   (2H + 1 limbs, non-negative)
    z1[0] = ptx::add_cc(t[0], t[2 * H]);
#pragma unroll
    for (int i = 1; i < 2 * H; i++) z1[i] = ptx::addc_cc(t[i], t[2 * H + i]);
    z1[2 * H] = ptx::addc(0, 0);
    // add (m xor sm) + (sm & 1), sign-extended: for sm = ~0 this is -m in two's complement over 2H + 1 limbs
    z1[0] = ptx::add_cc(z1[0], sm & 1u);
#pragma unroll
    for (int i = 1; i < 2 * H; i++) z1[i] = ptx::addc_cc(z1[i], 0);
    z1[2 * H] = ptx::addc(z1[2 * H], 0);
    z1[0] = ptx::add_cc(z1[0], m[0] ^ sm);
#pragma unroll
    for (int i = 1; i < 2 * H; i++) z1[i] = ptx::addc_cc(z1[i], m[i] ^ sm);
    z1[2 * H] = ptx::addc(z1[2 * H], sm);
    // t += z1 * 2^(32H)
    t[H] = ptx::add_cc(t[H], z1[0]);
#pragma unroll
    for (int i = 1; i <= 2 * H; i++) t[H + i] = ptx::addc_cc(t[H + i], z1[i]);
#pragma unroll
    for (int i = 3 * H + 1; i < 4 * H; i++) t[i] = ptx::addc_cc(t[i], 0);
  }

  // Montgomery product through Karatsuba + separated reduction; meant for the out-of-line body below
  // (the 2N-limb intermediate costs registers the inlined hot loop does not have).
  OZL_DEV Fp mul_kara(const Fp& b) const {
    uint32_t t[2 * N];
    mul_wide_kara(*this, b, t);
    return redc_wide(t);
  }
  static OZL_DEV_NOINLINE Fp mul_kara_ni(Fp a, Fp b) { return a.mul_kara(b); }

  // Out-of-line multiplier for the cold kernels: one ~5 KB copy instead of a ~6 KB inlined body
  // per call site, so bucket reduction / Horner / inversion stay resident in the instruction cache
  // (ncu: sm__icc_request_hit_rate 50 % -> the inlined versions were instruction-fetch bound).
  // Operands travel BY VALUE: the device ABI then keeps them in registers, whereas reference
  // parameters force the caller to spill both operands to its stack frame and the callee to reload
  // them (~70 local-memory operations around a ~330-instruction body).
  static OZL_DEV_NOINLINE Fp mul_ni(Fp a, Fp b) { return a * b; }
  static OZL_DEV Fp sqr_ni(const Fp& a) { return mul_ni(a, a); }
  // Dedicated squaring (separated operand scanning, 222 instead of 288 wide multiplies for N = 12) as its
  // own out-of-line body: inlined into k_accumulate it lost to mul() on register pressure, out of line the
  // pressure stays inside the callee.
  static OZL_DEV_NOINLINE Fp sqr_sos_ni(Fp a) { return a.sqr_sos(); }
  // Two independent products in one out-of-line body.  The cold kernels (bucket reduction, window
  // sums, Horner, proof assembly) run a handful of warps per SM, so a lone multiplication is bound by
  // the 4-cycle dependent-issue latency of its carry chains; ptxas interleaves the two chains here
  // (separate carry predicates), which nearly halves the time per product in that regime.
  struct Pair {
    Fp a, b;
  };
  static OZL_DEV_NOINLINE Pair mul2_ni(Fp a0, Fp b0, Fp a1, Fp b1) {
    Pair p;
    p.a = a0 * b0;
    p.b = a1 * b1;
    return p;
  }
  static OZL_DEV Pair sqr2_ni(const Fp& a0, const Fp& a1) { return mul2_ni(a0, a0, a1, a1); }

  // Leave / enter Montgomery form.
  OZL_DEV Fp from_mont() const {
    Fp o = zero();
    o.v[0] = 1;
    return *this * o;
  }
  OZL_DEV Fp to_mont() const { return *this * r2(); }

  // a^(p-2) by square-and-multiply over the bits of p - 2 (cold path: affine conversion).
  OZL_DEV_NOINLINE Fp inverse() const {
    uint32_t e[N];
    e[0] = ptx::sub_cc(P::mod()[0], 2);
#pragma unroll
    for (int i = 1; i < N; i++) e[i] = ptx::subc_cc(P::mod()[i], 0);
    Fp acc = one();
    for (int bit = P::BITS - 1; bit >= 0; bit--) {
      acc = sqr_ni(acc);
      if ((e[bit >> 5] >> (bit & 31)) & 1) acc = mul_ni(acc, *this);
    }
    return acc;
  }

  // memory <-> registers (16-byte vector accesses; pointers must be 16 B aligned)
  static OZL_DEV Fp load(const uint32_t* p) {
    Fp r;
#if defined(__CUDACC__)
    static_assert(N % 4 == 0, "limb count must be a multiple of 4");
#pragma unroll
    for (int i = 0; i < N; i += 4) {
      uint4 t = *reinterpret_cast<const uint4*>(p + i);
      r.v[i] = t.x; r.v[i + 1] = t.y; r.v[i + 2] = t.z; r.v[i + 3] = t.w;
    }
#else
    for (int i = 0; i < N; i++) r.v[i] = p[i];
#endif
    return r;
  }
  OZL_DEV void store(uint32_t* p) const {
#if defined(__CUDACC__)
#pragma unroll
    for (int i = 0; i < N; i += 4) {
      *reinterpret_cast<uint4*>(p + i) = make_uint4(v[i], v[i + 1], v[i + 2], v[i + 3]);
    }
#else
    for (int i = 0; i < N; i++) p[i] = v[i];
#endif
  }
};

// Quadratic extension Fq2 = Fq[u]/(u^2 + 1) (ark QuadExtField with NONRESIDUE = -1; both
// BLS12-381 and BN254 use it).  Memory layout c0 || c1, each a Montgomery Fq.
template <class P>
struct Fp2 {
  typedef Fp<P> Base;
  typedef P Params;
  static constexpr int N = 2 * P::N;
  Base c0, c1;

  static OZL_DEV Fp2 zero() { Fp2 r; r.c0 = Base::zero(); r.c1 = Base::zero(); return r; }
  static OZL_DEV Fp2 one() { Fp2 r; r.c0 = Base::one(); r.c1 = Base::zero(); return r; }
  static OZL_DEV Fp2 from_limbs(const uint32_t* c) {
    Fp2 r; r.c0 = Base::from_limbs(c); r.c1 = Base::from_limbs(c + P::N); return r;
  }
  OZL_DEV bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
  friend OZL_DEV bool operator==(const Fp2& a, const Fp2& b) { return a.c0 == b.c0 && a.c1 == b.c1; }
  friend OZL_DEV bool operator!=(const Fp2& a, const Fp2& b) { return !(a == b); }
  friend OZL_DEV Fp2 operator+(const Fp2& a, const Fp2& b) { Fp2 r; r.c0 = a.c0 + b.c0; r.c1 = a.c1 + b.c1; return r; }
  friend OZL_DEV Fp2 operator-(const Fp2& a, const Fp2& b) { Fp2 r; r.c0 = a.c0 - b.c0; r.c1 = a.c1 - b.c1; return r; }
  OZL_DEV Fp2 neg() const { Fp2 r; r.c0 = c0.neg(); r.c1 = c1.neg(); return r; }
  OZL_DEV Fp2 cneg(bool f) const { Fp2 r; r.c0 = c0.cneg(f); r.c1 = c1.cneg(f); return r; }
  OZL_DEV Fp2 dbl() const { Fp2 r; r.c0 = c0.dbl(); r.c1 = c1.dbl(); return r; }
  // Karatsuba: 3 base multiplications
  friend OZL_DEV Fp2 operator*(const Fp2& a, const Fp2& b) {
    Base t0 = a.c0 * b.c0;
    Base t1 = a.c1 * b.c1;
    Base t2 = (a.c0 + a.c1) * (b.c0 + b.c1);
    Fp2 r;
    r.c0 = t0 - t1;
    r.c1 = t2 - t0 - t1;
    return r;
  }
  static OZL_DEV_NOINLINE Fp2 mul_ni(Fp2 a, Fp2 b) {
    typename Base::Pair t = Base::mul2_ni(a.c0, b.c0, a.c1, b.c1);
    Base t2 = Base::mul_ni(a.c0 + a.c1, b.c0 + b.c1);
    Fp2 r;
    r.c0 = t.a - t.b;
    r.c1 = t2 - t.a - t.b;
    return r;
  }
  // Lazily reduced Karatsuba product: THREE unreduced base-field products and TWO Montgomery reductions, 5 N^2 wide
  // multiplies instead of the 6 N^2 of three full multiplications:
  //   t0 = a0 b0, t1 = a1 b1, t2 = (a0 + a1)(b0 + b1)  (sums unreduced: < 2p, so t2 < 4 p^2 < 2^(64 N))
  //   c1 = redc(t2 - t0 - t1)  = a0 b1 + a1 b0 < 2 p^2 < p R
  //   c0 = redc(t0 + p^2 - t1) in (0, 2 p^2)
  // Needs 2 p < R / 2, true for both base fields (p < R / 4).
  OZL_DEV Fp2 mul_lazy(const Fp2& o) const {
    constexpr int BN = P::N;
    static_assert(P::BITS <= 32 * BN - 2, "lazy reduction needs p < R / 4");
    uint32_t sa[BN], sb[BN], t0[2 * BN], t1[2 * BN], t2[2 * BN];
#pragma unroll
    for (int k = 0; k < BN; k++) { sa[k] = c0.v[k]; sb[k] = o.c0.v[k]; }
    Base::template add_limbs<BN>(sa, c1.v);
    Base::template add_limbs<BN>(sb, o.c1.v);
    Base::mul_wide(sa, sb, t2);
    Base::mul_wide(c0.v, o.c0.v, t0);
    Base::template sub_limbs<2 * BN>(t2, t0);
    Base::mul_wide(c1.v, o.c1.v, t1);
    Base::template sub_limbs<2 * BN>(t2, t1);
    Fp2 r;
    r.c1 = Base::redc_cios(t2);
    Base::template add_limbs<2 * BN>(t0, P::modsq());
    Base::template sub_limbs<2 * BN>(t0, t1);
    r.c0 = Base::redc_cios(t0);
    return r;
  }
  static OZL_DEV_NOINLINE Fp2 mul_lazy_ni(Fp2 a, Fp2 b) { return a.mul_lazy(b); }
  // a b - c d with the same three-product shape for both terms and still TWO reductions (8 N^2 wide multiplies
  // instead of 12 N^2): y3 = r (q - x3) - y p3 of every G2 addition.  The differences lie in (-2 p^2, 2 p^2); 2 p^2 is
  // added before the reduction (4 p^2 < p R needs p < R / 4).
  static OZL_DEV Fp2 mul_sub2_lazy(const Fp2& a, const Fp2& b, const Fp2& c, const Fp2& d) {
    constexpr int BN = P::N;
    static_assert(P::BITS <= 32 * BN - 2, "lazy reduction needs p < R / 4");
    uint32_t sa[BN], sb[BN], w[2 * BN], dd[2 * BN], u[2 * BN];
    // w = (a0 + a1)(b0 + b1) - a0 b0 - a1 b1 ; dd = a0 b0 - a1 b1 + 2 p^2
#pragma unroll
    for (int k = 0; k < BN; k++) { sa[k] = a.c0.v[k]; sb[k] = b.c0.v[k]; }
    Base::template add_limbs<BN>(sa, a.c1.v);
    Base::template add_limbs<BN>(sb, b.c1.v);
    Base::mul_wide(sa, sb, w);
    Base::mul_wide(a.c0.v, b.c0.v, dd);
    Base::template sub_limbs<2 * BN>(w, dd);
    Base::mul_wide(a.c1.v, b.c1.v, u);
    Base::template sub_limbs<2 * BN>(w, u);
    Base::template add_limbs<2 * BN>(dd, P::modsq());
    Base::template add_limbs<2 * BN>(dd, P::modsq());
    Base::template sub_limbs<2 * BN>(dd, u);
    // subtract the second product: w -= c0 d1 + c1 d0 (after adding 2 p^2) ; dd -= c0 d0 - c1 d1
    Base::template add_limbs<2 * BN>(w, P::modsq());
    Base::template add_limbs<2 * BN>(w, P::modsq());
    Base::mul_wide(c.c0.v, d.c0.v, u);
    Base::template sub_limbs<2 * BN>(dd, u);
    Base::template add_limbs<2 * BN>(w, u);
    Base::mul_wide(c.c1.v, d.c1.v, u);
    Base::template add_limbs<2 * BN>(dd, u);
    Base::template add_limbs<2 * BN>(w, u);
#pragma unroll
    for (int k = 0; k < BN; k++) { sa[k] = c.c0.v[k]; sb[k] = d.c0.v[k]; }
    Base::template add_limbs<BN>(sa, c.c1.v);
    Base::template add_limbs<BN>(sb, d.c1.v);
    Base::mul_wide(sa, sb, u);
    Base::template sub_limbs<2 * BN>(w, u);
    Fp2 r;
    r.c1 = Base::redc_cios(w);
    r.c0 = Base::redc_cios(dd);
    return r;
  }
  static OZL_DEV_NOINLINE Fp2 mul_sub2_lazy_ni(Fp2 a, Fp2 b, Fp2 c, Fp2 d) { return mul_sub2_lazy(a, b, c, d); }
  static OZL_DEV_NOINLINE Fp2 sqr_ni(Fp2 a) {
    typename Base::Pair t = Base::mul2_ni(a.c0 + a.c1, a.c0 - a.c1, a.c0, a.c1);
    Fp2 r;
    r.c0 = t.a;
    r.c1 = t.b.dbl();
    return r;
  }
  // two independent Fq2 products = six base products in three interleaved pairs
  struct Pair {
    Fp2 a, b;
  };
  static OZL_DEV_NOINLINE Pair mul2_ni(Fp2 x0, Fp2 y0, Fp2 x1, Fp2 y1) {
    typename Base::Pair p0 = Base::mul2_ni(x0.c0, y0.c0, x0.c1, y0.c1);
    typename Base::Pair p1 = Base::mul2_ni(x1.c0, y1.c0, x1.c1, y1.c1);
    typename Base::Pair p2 = Base::mul2_ni(x0.c0 + x0.c1, y0.c0 + y0.c1, x1.c0 + x1.c1, y1.c0 + y1.c1);
    Pair r;
    r.a.c0 = p0.a - p0.b;
    r.a.c1 = p2.a - p0.a - p0.b;
    r.b.c0 = p1.a - p1.b;
    r.b.c1 = p2.b - p1.a - p1.b;
    return r;
  }
  static OZL_DEV_NOINLINE Pair sqr2_ni(Fp2 a0, Fp2 a1) {
    typename Base::Pair s = Base::mul2_ni(a0.c0 + a0.c1, a0.c0 - a0.c1, a1.c0 + a1.c1, a1.c0 - a1.c1);
    typename Base::Pair m = Base::mul2_ni(a0.c0, a0.c1, a1.c0, a1.c1);
    Pair r;
    r.a.c0 = s.a;
    r.a.c1 = m.a.dbl();
    r.b.c0 = s.b;
    r.b.c1 = m.b.dbl();
    return r;
  }
  OZL_DEV Fp2 sqr_sos() const { return sqr(); }   // (test hook symmetry with Fp)
  static OZL_DEV Fp2 sqr_sos_ni(const Fp2& a) { return sqr_ni(a); }
  // a b + c d over Fq2 (no fused form: two products)
  static OZL_DEV Fp2 mul_add2(const Fp2& a, const Fp2& b, const Fp2& c, const Fp2& d) { return a * b + c * d; }
  static OZL_DEV Fp2 mul_add2_ni(const Fp2& a, const Fp2& b, const Fp2& c, const Fp2& d) { return mul_ni(a, b) + mul_ni(c, d); }
  OZL_DEV Fp2 mul_kara(const Fp2& b) const { return *this * b; }
  static OZL_DEV Fp2 mul_kara_ni(const Fp2& a, const Fp2& b) { return mul_ni(a, b); }
  // complex squaring: 2 base multiplications
  OZL_DEV Fp2 sqr() const {
    Base s = c0 + c1;
    Base d = c0 - c1;
    Base m = c0 * c1;
    Fp2 r;
    r.c0 = s * d;
    r.c1 = m.dbl();
    return r;
  }
  OZL_DEV Fp2 inverse() const {
    Base n = (Base::sqr_ni(c0) + Base::sqr_ni(c1)).inverse();
    Fp2 r;
    r.c0 = Base::mul_ni(c0, n);
    r.c1 = Base::mul_ni(c1, n).neg();
    return r;
  }
  static OZL_DEV Fp2 load(const uint32_t* p) { Fp2 r; r.c0 = Base::load(p); r.c1 = Base::load(p + P::N); return r; }
  OZL_DEV void store(uint32_t* p) const { c0.store(p); c1.store(p + P::N); }
};

}  // namespace ozl
