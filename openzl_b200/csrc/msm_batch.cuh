// Batched-affine pair tree: the first levels of bucket accumulation done with AFFINE additions whose
// inversions are shared by Montgomery's trick, so an addition costs ~6.5 field multiplications
// instead of the 10 of an XYZZ mixed addition (same group elements; ark's result is unchanged).
//
// Level l turns the per-bucket entry lists of level l-1 (m_g entries in bucket g) into lists of
// ceil(m_g / 2) entries: output j of bucket g = entry 2j + entry 2j+1 (or a copy of entry 2j when it
// has no partner).  Offsets of every level come from one scan each of ceil(count / 2^l).  After T
// levels the (8x shorter, for T = 3) lists go through the ordinary XYZZ slice accumulation.
//
// One thread owns k consecutive OUTPUT entries: a forward pass multiplies up the denominators
// d = x1 - x0 (prefix products parked in a [j][thread]-strided scratch so warps store coalesced),
// one Fermat inversion, then a backward pass peels the inverses off and finishes each addition:
//   forward   1 M                      per addition
//   inversion (BITS + popcount) M      per thread, amortised over k additions
//   backward  2 M (inverse peel) + 2 M + 1 S (lambda, x3, y3)
//
// EXPERIMENTAL, off by default (ozl_msm_set_batch_affine / OZL_MSM_BATCH).  Measured on B200 at
// 2^26 BLS12-381 G1: the levels run at 2.7e9 additions/s against 2.7e9 for
// the XYZZ kernel -- no gain.  Per addition the instruction count only drops from ~3950 to ~3300
// (6 multiplications + inversion share + 7 field subtractions + strided scratch traffic), the
// level-1 gathers are latency-bound at 12 warps/SM, and DRAM traffic rises from ~150 B to
// ~650 B per addition.  Splitting into forward / peel / finish kernels for occupancy, and L1/L2
// software prefetch, were both measured slower (more threads shorten the batches an inversion is
// amortised over; prefetches thrash).  Kept because it is bit-exact and documents the experiment.
// Exceptional pairs (equal x: doubling or cancellation; an operand that is itself a cancelled
// pair) take an out-of-line path; the point at infinity is encoded as (0, 0), which is on none of
// the four curves (b != 0).
#pragma once
#include "ec.cuh"

namespace ozl {

template <class F>
OZL_DEV void store_strided(uint32_t* p, size_t stride_u32, const F& f) {
  const uint32_t* s = reinterpret_cast<const uint32_t*>(&f);
#pragma unroll
  for (int i = 0; i < F::N; i += 4)
    *reinterpret_cast<uint4*>(p + (size_t)(i >> 2) * stride_u32) = make_uint4(s[i], s[i + 1], s[i + 2], s[i + 3]);
}
template <class F>
OZL_DEV F load_strided(const uint32_t* p, size_t stride_u32) {
  F f;
  uint32_t* d = reinterpret_cast<uint32_t*>(&f);
#pragma unroll
  for (int i = 0; i < F::N; i += 4) {
    const uint4 t = *reinterpret_cast<const uint4*>(p + (size_t)(i >> 2) * stride_u32);
    d[i] = t.x; d[i + 1] = t.y; d[i + 2] = t.z; d[i + 3] = t.w;
  }
  return f;
}

template <class F>
OZL_DEV bool affine_is_inf(const Affine<F>& p) { return p.x.is_zero() && p.y.is_zero(); }

// Classification of an exceptional pair.  Returns 0 = ordinary chord (d, num set), 1 = tangent
// (d = 2 y0, num = 3 x0^2), 2 = result is p0, 3 = result is p1, 4 = result is the identity.
// Inlined and fed copies so that the hot loop's operands never have their address taken.
template <class F>
OZL_DEV int pair_classify(const Affine<F> p0, const Affine<F> p1, F& d, F& num) {
  if (affine_is_inf(p0)) return 3;
  if (affine_is_inf(p1)) return 2;
  if (p0.x == p1.x) {
    if (p0.y == p1.y && !p0.y.is_zero()) {
      d = p0.y.dbl();
      F xx = F::sqr_ni(p0.x);
      num = xx.dbl() + xx;
      return 1;
    }
    return 4;
  }
  d = p1.x - p0.x;
  num = p1.y - p0.y;
  return 0;
}

// One output entry's operands: a0 / a1 index the source point array (level 1: through the sorted
// index array, sign in bit 31 negates y; later levels: the previous level's points).
struct PairDesc {
  uint32_t a0, a1;
  bool paired;
};

template <class F, bool FIRST>
struct PairSrc {
  const uint32_t* pts;
  const uint32_t* sorted;
  uint32_t in_base;
  static constexpr int AFF = 2 * F::N;
  OZL_DEV PairDesc describe(uint32_t i0, bool paired) const {
    PairDesc d;
    d.paired = paired;
    d.a0 = FIRST ? sorted[i0] : i0 - in_base;
    d.a1 = paired ? (FIRST ? sorted[i0 + 1] : i0 + 1 - in_base) : 0u;
    return d;
  }
  OZL_DEV const uint32_t* addr(uint32_t a) const { return pts + (size_t)(a & 0x7fffffffu) * AFF; }
  OZL_DEV F load_x(uint32_t a) const { return F::load(addr(a)); }
  OZL_DEV Affine<F> load(uint32_t a) const {
    Affine<F> p = load_affine<F>(addr(a));
    if (FIRST) p.y = p.y.cneg((a >> 31) != 0);
    return p;
  }
};

// Position of a thread in the bucket lists.  A warp's 32 lanes sit in 32 different buckets and some
// lane crosses a boundary in about every second iteration, so the boundaries of the NEIGHBOURING
// bucket are kept preloaded: a crossing is a register shuffle plus two loads that nobody waits for.
struct BucketWalker {
  const uint32_t* off_in;
  const uint32_t* off_out;
  uint32_t g, out_lo, out_hi, in_lo, in_hi;
  uint32_t nb_out, nb_in;   // forward: off_*[g + 2]; backward: off_*[g - 1]
  OZL_DEV void reload() {
    out_lo = off_out[g]; out_hi = off_out[g + 1];
    in_lo = off_in[g]; in_hi = off_in[g + 1];
  }
  OZL_DEV void locate(uint32_t o, uint32_t g_lo, uint32_t g_hi) {
    uint32_t lo = g_lo, hi = g_hi;   // invariant: off_out[lo] <= o < off_out[hi]
    while (hi - lo > 1) {
      const uint32_t mid = (lo + hi) >> 1;
      if (off_out[mid] <= o) lo = mid; else hi = mid;
    }
    g = lo;
    reload();
  }
  // offsets arrays have NB + 1 entries; g + 2 <= NB whenever a further bucket can be entered
  OZL_DEV void arm_forward(uint32_t g_hi) { nb_out = g + 2 <= g_hi ? off_out[g + 2] : out_hi; nb_in = g + 2 <= g_hi ? off_in[g + 2] : in_hi; }
  OZL_DEV void arm_backward() { nb_out = g ? off_out[g - 1] : 0u; nb_in = g ? off_in[g - 1] : 0u; }
  OZL_DEV void forward(uint32_t o, uint32_t g_hi) {
    while (o >= out_hi) {
      g++;
      out_lo = out_hi; in_lo = in_hi;
      out_hi = nb_out; in_hi = nb_in;
      arm_forward(g_hi);
    }
  }
  OZL_DEV void backward(uint32_t o) {
    while (o < out_lo) {
      g--;
      out_hi = out_lo; in_hi = in_lo;
      out_lo = nb_out; in_lo = nb_in;
      arm_backward();
    }
  }
  // input index of the first operand of output o, and whether it has a partner
  OZL_DEV uint32_t first(uint32_t o) const { return in_lo + 2u * (o - out_lo); }
  OZL_DEV bool paired(uint32_t i0) const { return i0 + 1u < in_hi; }
};

OZL_DEV void prefetch_l2(const void* p) {
#if defined(__CUDACC__)
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#endif
}

template <class F>
OZL_DEV F select(bool c, const F& a, const F& b) {   // c ? a : b, limb by limb (no address taken)
  F r;
  const uint32_t* pa = reinterpret_cast<const uint32_t*>(&a);
  const uint32_t* pb = reinterpret_cast<const uint32_t*>(&b);
  uint32_t* pr = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
  for (int i = 0; i < F::N; i++) pr[i] = c ? pa[i] : pb[i];
  return r;
}

// Operands of one output as they move through the software pipeline.
struct PairSlot {
  uint32_t a0, a1;
  bool live;      // inside this thread's range
  bool paired;    // has a partner (otherwise the output is a copy of operand 0)
};

// Both passes are software-pipelined by hand (ptxas does not pipeline loops of this size): the
// index loads run two outputs ahead and the operand loads one output ahead of the arithmetic, so
// the dependent gather index -> point is in flight while the previous output is being multiplied;
// even and odd outputs feed two independent product chains, and in the backward pass the inverse
// peel of output j-1 shares a basic block with the three-multiplication finish chain of output j.
template <class F, bool FIRST>
__global__ void __launch_bounds__(128, (F::N <= 8 ? 4 : (F::N <= 12 ? 3 : 2)))
k_pair_level(const uint32_t* __restrict__ pts, const uint32_t* __restrict__ sorted, const uint32_t* __restrict__ off_in,
             const uint32_t* __restrict__ off_out, uint32_t g_lo, uint32_t g_hi, uint32_t o_begin, uint32_t o_end,
             uint32_t in_base, uint32_t dst_base, uint32_t k, uint32_t* __restrict__ pre, uint32_t* __restrict__ dst) {
  constexpr int AFF = 2 * F::N;
  constexpr int PARTS = F::N / 4;
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t o0_64 = (uint64_t)o_begin + (uint64_t)t * k;
  if (o0_64 >= o_end) return;
  const uint32_t o0 = (uint32_t)o0_64;
  const uint32_t cnt = min(k, o_end - o0);
  const PairSrc<F, FIRST> src{pts, sorted, in_base};
  const size_t stride = (size_t)gridDim.x * blockDim.x * 4;   // u32 between the 16-byte parts of one element
  uint32_t* mypre = pre + (size_t)t * 4;
  BucketWalker w{off_in, off_out, 0, 0, 0, 0, 0, 0, 0};
  w.locate(o0, g_lo, g_hi);
  w.arm_forward(g_hi);

  // ---- forward: pre[j] = product of the earlier denominators of the chain (j & 1) ----------
  auto slot_fwd = [&](uint32_t j) {
    PairSlot s{0u, 0u, j < cnt, false};
    if (s.live) {
      const uint32_t o = o0 + j;
      w.forward(o, g_hi);
      const uint32_t i0 = w.first(o);
      s.paired = w.paired(i0);
      if (s.paired) {
        const PairDesc pd = src.describe(i0, true);
        s.a0 = pd.a0; s.a1 = pd.a1;
      }
    }
    return s;
  };
  F acc_p = F::one(), acc_q = F::one();               // chain of the current output / of the other parity
  {
    PairSlot cur = slot_fwd(0), nxt = slot_fwd(1);
    F x0 = F::zero(), x1 = F::zero();
    if (cur.paired) { x0 = src.load_x(cur.a0); x1 = src.load_x(cur.a1); }
    for (uint32_t j = 0; j < cnt; j++) {
      F nx0 = F::zero(), nx1 = F::zero();
      if (nxt.paired) { nx0 = src.load_x(nxt.a0); nx1 = src.load_x(nxt.a1); }
      const PairSlot nn = slot_fwd(j + 2);
      if (cur.paired) {
        F d = x1 - x0;
        bool use = true;
        if (d.is_zero() || (!FIRST && (x0.is_zero() || x1.is_zero()))) {
          F num;
          use = pair_classify<F>(src.load(cur.a0), src.load(cur.a1), d, num) <= 1;
        }
        if (use) {
          store_strided<F>(mypre + (size_t)j * PARTS * stride, stride, acc_p);
          acc_p = acc_p * d;
        }
      }
      const F tmp = acc_p; acc_p = acc_q; acc_q = tmp;
      cur = nxt; nxt = nn; x0 = nx0; x1 = nx1;
    }
  }
  // after cnt swaps acc_p is the chain of parity (cnt & 1); the last output cnt-1 has the other one
  const F tot = acc_p * acc_q;
  const F itot = tot.inverse();                       // product of non-zero factors
  F inv_p = itot * acc_p;                             // = 1 / acc_q : chain of output cnt-1
  F inv_q = itot * acc_q;                             // = 1 / acc_p : chain of output cnt-2

  // ---- backward ----------------------------------------------------------------------------
  auto slot_bwd = [&](uint32_t j, bool live) {
    PairSlot s{0u, 0u, live, false};
    if (live) {
      const uint32_t o = o0 + j;
      w.backward(o);
      const uint32_t i0 = w.first(o);
      s.paired = w.paired(i0);
      const PairDesc pd = src.describe(i0, s.paired);
      s.a0 = pd.a0; s.a1 = pd.a1;
    }
    return s;
  };
  // peel of one output: A = running inverse * prefix, running inverse *= d.  Returns false when the
  // output needs no inverse (copy, cancellation, operand at infinity).
  auto peel = [&](uint32_t j, const PairSlot& s, const F& x0, const F& x1, const F& pj, F& inv, F& A) {
    if (!s.paired) return;
    F d = x1 - x0;
    if (d.is_zero() || (!FIRST && (x0.is_zero() || x1.is_zero()))) {
      F num;
      if (pair_classify<F>(src.load(s.a0), src.load(s.a1), d, num) > 1) return;
    }
    A = inv * pj;
    inv = inv * d;
  };
  w.arm_backward();
  PairSlot cur = slot_bwd(cnt - 1, true);
  PairSlot nxt = slot_bwd(cnt - 2, cnt >= 2);
  F A = F::zero();
  {
    F x0 = F::zero(), x1 = F::zero(), pj = F::zero();
    if (cur.paired) {
      x0 = src.load_x(cur.a0); x1 = src.load_x(cur.a1);
      pj = load_strided<F>(mypre + (size_t)(cnt - 1) * PARTS * stride, stride);
    }
    peel(cnt - 1, cur, x0, x1, pj, inv_p, A);
    const F tmp = inv_p; inv_p = inv_q; inv_q = tmp;
  }
  for (uint32_t j = cnt; j-- > 0;) {
    // operand loads of output j-1 (its peel is at the bottom of this iteration)
    F nx0 = F::zero(), nx1 = F::zero(), npj = F::zero();
    if (nxt.paired) {
      nx0 = src.load_x(nxt.a0); nx1 = src.load_x(nxt.a1);
      npj = load_strided<F>(mypre + (size_t)(j - 1) * PARTS * stride, stride);
      prefetch_l2(src.addr(nxt.a0) + AFF - 4);        // tail of y, for the finish of output j-1
      prefetch_l2(src.addr(nxt.a1) + AFF - 4);
    } else if (nxt.live) {
      prefetch_l2(src.addr(nxt.a0));
      prefetch_l2(src.addr(nxt.a0) + AFF - 4);
    }
    const PairSlot nn = slot_bwd(j - 2, j >= 2);
    // finish output j
    {
      uint32_t* out = dst + (size_t)(o0 + j - dst_base) * AFF;
      const Affine<F> p0 = src.load(cur.a0);
      if (!cur.paired) {
        p0.x.store(out); p0.y.store(out + F::N);
      } else {
        const Affine<F> p1 = src.load(cur.a1);
        F d = p1.x - p0.x;
        F num = p1.y - p0.y;
        int kind = 0;
        if (d.is_zero() || (!FIRST && (p0.x.is_zero() || p1.x.is_zero()))) {
          F dd, nn2;
          kind = pair_classify<F>(p0, p1, dd, nn2);
          if (kind <= 1) num = nn2;
        }
        if (kind > 1) {
          Affine<F> r;
          if (kind == 2) r = p0;
          else if (kind == 3) r = p1;
          else { r.x = F::zero(); r.y = F::zero(); }
          r.x.store(out); r.y.store(out + F::N);
        } else {
          const F lam = num * A;
          const F x3 = lam.sqr() - p0.x - p1.x;
          const F y3 = lam * (p0.x - x3) - p0.y;
          x3.store(out); y3.store(out + F::N);
        }
      }
    }
    // peel output j-1
    F An = F::zero();
    if (nxt.live) peel(j - 1, nxt, nx0, nx1, npj, inv_p, An);
    const F tmp = inv_p; inv_p = inv_q; inv_q = tmp;
    A = An;
    cur = nxt; nxt = nn;
  }
}

// Chunk table: the bucket range is cut into Q pieces of about equal level-0 entry counts so the
// ping-pong buffers of the levels stay bounded.  tab[q] = first bucket of piece q (tab[Q] = NB);
// tab[(1 + l) * (Q + 1) + q] = level-l offset at that bucket, l = 0..T.
static __global__ void k_pair_chunk_table(const uint32_t* __restrict__ off0, const uint32_t* __restrict__ lvl_off, uint32_t NB,
                                          uint32_t Q, uint32_t T, uint32_t* __restrict__ tab) {
  const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q > Q) return;
  const uint32_t E = off0[NB];
  uint32_t g = NB;
  if (q < Q) {
    const uint32_t target = (uint32_t)(((uint64_t)E * q) / Q);
    uint32_t lo = 0, hi = NB;              // first g with off0[g] >= target
    while (lo < hi) {
      const uint32_t mid = (lo + hi) >> 1;
      if (off0[mid] >= target) hi = mid; else lo = mid + 1;
    }
    g = lo;
  }
  tab[q] = g;
  tab[(size_t)(Q + 1) + q] = off0[g];
  for (uint32_t l = 1; l <= T; l++) tab[(size_t)(1 + l) * (Q + 1) + q] = lvl_off[(size_t)(l - 1) * (NB + 1) + g];
}

}  // namespace ozl
