// Radix-2 NTT over the scalar field for sm_100a.
//
// Replaces ark_poly::Radix2EvaluationDomain::{fft,ifft,coset_fft,coset_ifft}_in_place
// (ark-poly 0.3.0, reached through `pub use poly` at
// /root/reference/plugins/arkworks/src/lib.rs:70-71; called seven times per proof by
// ark_groth16's R1CStoQAP::witness_map behind groth16.rs:454).  Same contract: natural order
// in, natural order out, X[k] = sum_j x[j] w^(jk); inverse multiplies by size_inv; coset
// variants scale by powers of the multiplicative generator before (forward) / after (inverse).
//
// Schedule: decimation-in-frequency, ceil(log n / 3) passes.  Each thread keeps a radix-8
// group (8 x 256-bit elements) in registers and performs three butterfly stages per pass, so
// every pass reads and writes each element once with 32-byte-per-thread, warp-contiguous
// accesses.  The last pass writes through the bit-reversal permutation, which makes the
// output natural-ordered without a separate permutation pass; the coset / size_inv scalings
// are fused into the first pass's loads / the last pass's stores.  Twiddles come from a
// device-resident table w^e, e < n/2, built once per (field, log n, direction).
#pragma once
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdlib>
#include "fp.cuh"
#include "ntt_ws.h"

namespace ozl {

static constexpr int NTT_LO_BITS = 12;  // two-level power table split for coset scaling

template <class P>
struct NttConsts {
  Fp<P> omega;      // w (or w^-1 for the inverse direction)
  Fp<P> size_inv;   // n^-1 (Montgomery)
  Fp<P> g;          // coset generator g (forward) or g^-1 (inverse)
  Fp<P> g_hi;       // g^(2^lo_bits)
  Fp<P> one;
};

template <class P>
__device__ __forceinline__ Fp<P> fp_pow_u32(const Fp<P>& b, uint32_t e) {
  Fp<P> acc = Fp<P>::one();
  for (int bit = 31; bit >= 0; bit--) {
    acc = acc.sqr();
    if ((e >> bit) & 1) acc = acc * b;
  }
  return acc;
}

template <class P>
__global__ void k_ntt_setup(int log_n, int inverse, NttConsts<P>* c) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  typedef Fp<P> F;
  F w = F::from_limbs(inverse ? P::root_inv() : P::root());
  for (int i = 0; i < P::TWO_ADICITY - log_n; i++) w = w.sqr();
  c->omega = w;
  F n = F::zero();
  n.v[0] = 1u << (log_n & 31);
  if (log_n >= 32) { n.v[0] = 0; n.v[1] = 1u << (log_n - 32); }
  c->size_inv = n.to_mont().inverse();
  F g = F::from_limbs(inverse ? P::gen_inv() : P::gen());
  c->g = g;
  const int lo = log_n < NTT_LO_BITS ? log_n : NTT_LO_BITS;
  F gh = g;
  for (int i = 0; i < lo; i++) gh = gh.sqr();
  c->g_hi = gh;
  c->one = F::one();
}

// out[i] = scale * base^i for i < count; every thread produces POW_RUN consecutive entries
static constexpr int POW_RUN = 32;
template <class P>
__global__ void __launch_bounds__(128)
k_build_powers(const Fp<P>* __restrict__ base_p, const Fp<P>* __restrict__ scale_p, uint32_t count, uint32_t* __restrict__ out) {
  typedef Fp<P> F;
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t first = (uint64_t)t * POW_RUN;
  if (first >= count) return;
  const F base = *base_p;
  F cur = fp_pow_u32<P>(base, (uint32_t)first) * (*scale_p);
  const uint32_t m = (uint32_t)min((uint64_t)POW_RUN, (uint64_t)count - first);
  for (uint32_t i = 0; i < m; i++) {
    cur.store(out + (first + i) * F::N);
    cur = cur * base;
  }
}

// One DIF pass of R stages starting at stage s.
//   pre  : multiply inputs by g^i           (coset forward, first pass only)
//   post : 0 none | 1 multiply outputs by *scale | 2 multiply outputs by ghi[k>>lo]*glo[k&mask]
//   last : write through the bit-reversal permutation (out must not alias in)
template <class P, int R>
__global__ void __launch_bounds__(128)
k_ntt_pass(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, const uint32_t* __restrict__ tw, int log_n, int s,
           int pre, int post, int last, const uint32_t* __restrict__ glo, const uint32_t* __restrict__ ghi,
           const Fp<P>* __restrict__ scale) {
  typedef Fp<P> F;
  constexpr int M = 1 << R;
  const uint32_t n = 1u << log_n;
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (n >> R)) return;
  const uint32_t q = n >> (s + R);          // smallest stride of this pass
  const uint32_t blk = t / q, j0 = t - blk * q;
  const uint32_t base = blk * (n >> s) + j0;
  const int lo_bits = log_n < NTT_LO_BITS ? log_n : NTT_LO_BITS;
  const uint32_t lo_mask = (1u << lo_bits) - 1u;

  F x[M];
#pragma unroll
  for (int m = 0; m < M; m++) {
    const uint32_t i = base + (uint32_t)m * q;
    x[m] = F::load(in + (size_t)i * F::N);
    if (pre) {
      F gp = F::load(ghi + (size_t)(i >> lo_bits) * F::N) * F::load(glo + (size_t)(i & lo_mask) * F::N);
      x[m] = x[m] * gp;
    }
  }
#pragma unroll
  for (int r = 0; r < R; r++) {
    const int half = 1 << (R - 1 - r);
#pragma unroll
    for (int m = 0; m < M; m++) {
      if (m & half) continue;
      const uint32_t e = (j0 + (uint32_t)(m & (half - 1)) * q) << (s + r);
      const F a = x[m], b = x[m + half];
      x[m] = a + b;
      if (e == 0) {
        x[m + half] = a - b;   // unit twiddle: every butterfly of the final stage, half of the one before
      } else {
        const F w = F::load(tw + (size_t)e * F::N);
        x[m + half] = (a - b) * w;
      }
    }
  }
#pragma unroll
  for (int m = 0; m < M; m++) {
    const uint32_t i = base + (uint32_t)m * q;
    uint32_t k = i;
    if (last) k = __brev(i) >> (32 - log_n);
    F v = x[m];
    if (post == 1) v = v * (*scale);
    else if (post == 2) v = v * (F::load(ghi + (size_t)(k >> lo_bits) * F::N) * F::load(glo + (size_t)(k & lo_mask) * F::N));
    v.store(out + (size_t)k * F::N);
  }
}

// ---------------------------------------------------------------------------------------------
// Shared-memory tile pass: R DIF stages per trip through HBM (R = 3, 6, 9 with radix-8 rounds; 4, 6, 8 with the
// shipped radix-4 rounds, see NTT_TILE4_* below -- the description that follows is the radix-8 geometry).
//
// A CTA of 256 threads owns a tile of 2048 elements = M x C, M = 2^R points of one sub-transform (stride
// q = n >> (s + R) apart in memory) times C = 2048 / M neighbouring sub-transforms ("columns": consecutive
// addresses while q >= C, consecutive blocks in the last pass where q = 1), so global traffic moves in
// runs of C x 32 bytes (or whole columns).  The tile lives in shared memory as two planes of 16-byte
// half-elements: eight lanes reading neighbouring columns then touch 128 contiguous bytes -> no bank
// conflicts.  The R stages run as R / 3 rounds; in a round every thread takes one radix-8 group out of
// the tile (elements ql = M >> (ls + 3) rows apart), does three stages in registers and puts it back.
//
// Code size is what bounded the register-only kernel above (ncu: sm__icc_request_hit_rate 69 %,
// stalled_no_instruction the second largest stall, with twelve inlined multiplier bodies per pass): here the
// stage loop is ROLLED.  Each iteration runs the four butterflies (x[i], x[i + 4]) -- four multiplier
// bodies in the whole kernel -- and then rotates the register file by one index bit
// (position b2 b1 b0 -> b1 b0 b2), which brings the next stage's partners to distance 4 again; three
// rotations are the identity, so a round leaves x[] in its original order.
// ---------------------------------------------------------------------------------------------
static constexpr int NTT_TILE_ELEMS = 2048;
static constexpr int NTT_TILE_THREADS = 256;
// rows of the tile are C + 1 half-elements apart in the last pass (q = 1), where the global phases walk a
// column: a stride of C * 16 bytes would put all eight lanes of a shared-memory phase on the same banks
// (ncu, unpadded: 5.9e7 bank conflicts in the last pass against 8.5e6 in the others)
static constexpr int NTT_TILE_SMEM_HALVES = NTT_TILE_ELEMS + 512;
static constexpr int NTT_TILE_SMEM_BYTES = 2 * NTT_TILE_SMEM_HALVES * 16;
// radix-4 variant (shipped; OZL_NTT_R4=0 selects the radix-8 one): 1024-element tiles, four elements per thread, two
// stages per round: 32 instead of 64 registers of data per thread, so four CTAs (32 warps) per SM instead of two (16)
static constexpr int NTT_TILE4_ELEMS = 1024;
static constexpr int NTT_TILE4_SMEM_HALVES = NTT_TILE4_ELEMS + 256;
static constexpr int NTT_TILE4_SMEM_BYTES = 2 * NTT_TILE4_SMEM_HALVES * 16;

// G = stages per round (3: radix-8 groups, 2: radix-4 groups); a CTA of THREADS threads owns THREADS << G elements.
// PAIRED: the products of a stage go through the paired out-of-line multiplier (two independent carry chains
// interleaved in one body) instead of inlined bodies (measured slower: 128 registers, spills).
template <class P, bool PAIRED = false, int G = 3, int THREADS = NTT_TILE_THREADS, int MINB = 2>
__global__ void __launch_bounds__(THREADS, MINB)
k_ntt_tile(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, const uint32_t* __restrict__ tw, int log_n, int s, int R,
           int pre, int post, int last, const uint32_t* __restrict__ glo, const uint32_t* __restrict__ ghi,
           const Fp<P>* __restrict__ scale) {
  typedef Fp<P> F;
  static_assert(F::N == 8, "tile layout assumes 32-byte elements");
  constexpr int NG = 1 << G, NH = NG / 2;
  constexpr uint32_t TILE = (uint32_t)THREADS << G;
  constexpr uint32_t HALVES = TILE + TILE / 4;
  extern __shared__ __align__(16) uint4 ntt_smem[];
  uint4* plane0 = ntt_smem;                            // low 16 bytes of every element
  uint4* plane1 = ntt_smem + HALVES;                   // high 16 bytes
  const uint32_t n = 1u << log_n;
  const uint32_t M = 1u << R, C = TILE >> R;
  const uint32_t q = n >> (s + R);                // memory stride between the points of one sub-transform
  const uint32_t u_base = blockIdx.x * C;         // first column: column u = blk * q + j0
  const int lo_bits = log_n < NTT_LO_BITS ? log_n : NTT_LO_BITS;
  const uint32_t lo_mask = (1u << lo_bits) - 1u;
  const uint32_t tid = threadIdx.x;
  auto col_base = [&](uint32_t c) -> uint32_t {   // memory index of point 0 of column c
    const uint32_t u = u_base + c;
    const uint32_t blk = u / q, j0 = u - blk * q;
    return blk * (n >> s) + j0;
  };
  // thread -> tile element for the global phases: columns fastest while they are adjacent in memory
  const bool col_major = q == 1;                  // last pass: a column is contiguous in memory, columns are M apart
  const uint32_t RS = C + (col_major ? 1u : 0u);  // row stride of the tile in shared memory
  // ---- load -------------------------------------------------------------------------------
  for (uint32_t t = tid; t < TILE; t += THREADS) {
    const uint32_t m = col_major ? (t & (M - 1)) : (t / C);
    const uint32_t c = col_major ? (t >> R) : (t & (C - 1));
    const uint32_t i = col_base(c) + m * q;
    F v = F::load(in + (size_t)i * F::N);
    if (pre) {
      const F gp = F::mul_ni(F::load(ghi + (size_t)(i >> lo_bits) * F::N), F::load(glo + (size_t)(i & lo_mask) * F::N));
      v = F::mul_ni(v, gp);
    }
    plane0[m * RS + c] = make_uint4(v.v[0], v.v[1], v.v[2], v.v[3]);
    plane1[m * RS + c] = make_uint4(v.v[4], v.v[5], v.v[6], v.v[7]);
  }
  __syncthreads();
  // ---- rounds of G stages -------------------------------------------------------------------
  {
    const uint32_t c = tid & (C - 1);
    const uint32_t rest = tid / C;                // (blkl, jl) of this thread's group: one group per thread per round
    const uint32_t u = u_base + c;
    const uint32_t j0 = u - (u / q) * q;
#pragma unroll 1
    for (int ls = 0; ls < R; ls += G) {
      const uint32_t ql = M >> (ls + G);
      const uint32_t blkl = rest / ql, jl = rest - blkl * ql;
      const uint32_t m0 = blkl * (M >> ls) + jl;
      // twiddle exponent of the butterfly at position i of stage r of this round (0 = unit twiddle).
      // Position i has been rotated left r times: the group index k it holds is i rotated right r times.
      auto exponent = [&](int r, int i) -> uint32_t {
        const uint32_t k = ((uint32_t)i >> r) | (((uint32_t)i << (G - r)) & (uint32_t)(NG - 1));
        const uint32_t kmask = ((uint32_t)NH >> r) - 1u;    // bits of k below the stage bit
        return ((jl + (k & kmask) * ql) * q + j0) << (s + ls + r);
      };
      // (Measured and rejected: prefetch.global.L1 of a stage's four twiddle lines one stage ahead -- 4.38 ms
      // against 4.27 ms at 2^24; the extra address arithmetic and LSU traffic cost more than the
      // long-scoreboard stalls they remove.)
      F x[NG];
#pragma unroll
      for (int k = 0; k < NG; k++) {
        const uint32_t a = (m0 + (uint32_t)k * ql) * RS + c;
        const uint4 l = plane0[a], h = plane1[a];
        x[k].v[0] = l.x; x[k].v[1] = l.y; x[k].v[2] = l.z; x[k].v[3] = l.w;
        x[k].v[4] = h.x; x[k].v[5] = h.y; x[k].v[6] = h.z; x[k].v[7] = h.w;
      }
#pragma unroll 1
      for (int r = 0; r < G; r++) {
        if (PAIRED && G == 3) {
          F w[NH], d[NH];
          bool any = false;
#pragma unroll
          for (int i = 0; i < NH; i++) {
            const uint32_t e = exponent(r, i);
            w[i] = F::load(tw + (size_t)e * F::N);     // tw[0] = 1
            any |= e != 0;
            const F a = x[i], b = x[i + NH];
            x[i] = a + b;
            d[i] = a - b;
          }
          if (any) {
#pragma unroll
            for (int i = 0; i < NH; i += 2) {
              const typename F::Pair p0 = F::mul2_ni(d[i], w[i], d[i + 1], w[i + 1]);
              x[NH + i] = p0.a; x[NH + i + 1] = p0.b;
            }
          } else {
#pragma unroll
            for (int i = 0; i < NH; i++) x[NH + i] = d[i];
          }
        } else {
#pragma unroll
          for (int i = 0; i < NH; i++) {
            const uint32_t e = exponent(r, i);
            const F a = x[i], b = x[i + NH];
            x[i] = a + b;
            if (e == 0) {
              x[i + NH] = a - b;
            } else {
              const F w = F::load(tw + (size_t)e * F::N);
              x[i + NH] = (a - b) * w;
            }
          }
        }
        // rotate the index bits left (b2 b1 b0 -> b1 b0 b2): the next stage's partners are NH apart again;
        // G rotations are the identity, so a round leaves x[] in its original order
        F y[NG];
#pragma unroll
        for (int i = 0; i < NG; i++) y[((i << 1) | (i >> (G - 1))) & (NG - 1)] = x[i];
#pragma unroll
        for (int i = 0; i < NG; i++) x[i] = y[i];
      }
#pragma unroll
      for (int k = 0; k < NG; k++) {
        const uint32_t a = (m0 + (uint32_t)k * ql) * RS + c;
        plane0[a] = make_uint4(x[k].v[0], x[k].v[1], x[k].v[2], x[k].v[3]);
        plane1[a] = make_uint4(x[k].v[4], x[k].v[5], x[k].v[6], x[k].v[7]);
      }
      __syncthreads();
    }
  }
  // ---- store ------------------------------------------------------------------------------
  for (uint32_t t = tid; t < TILE; t += THREADS) {
    const uint32_t m = col_major ? (t & (M - 1)) : (t / C);
    const uint32_t c = col_major ? (t >> R) : (t & (C - 1));
    const uint32_t i = col_base(c) + m * q;
    uint32_t k = i;
    if (last) k = __brev(i) >> (32 - log_n);
    const uint4 l = plane0[m * RS + c], h = plane1[m * RS + c];
    F v;
    v.v[0] = l.x; v.v[1] = l.y; v.v[2] = l.z; v.v[3] = l.w; v.v[4] = h.x; v.v[5] = h.y; v.v[6] = h.z; v.v[7] = h.w;
    if (post == 1) v = F::mul_ni(v, *scale);
    else if (post == 2) v = F::mul_ni(v, F::mul_ni(F::load(ghi + (size_t)(k >> lo_bits) * F::N), F::load(glo + (size_t)(k & lo_mask) * F::N)));
    v.store(out + (size_t)k * F::N);
  }
}

inline int ntt_ensure(void** p, size_t* cap, size_t bytes) {
  if (bytes <= *cap) return 0;
  if (*p) { cudaDeviceSynchronize(); cudaFree(*p); *p = nullptr; *cap = 0; }
  if (cudaMalloc(p, bytes) != cudaSuccess) return -4;
  *cap = bytes;
  return 0;
}

// returns 0 ok, -2 domain too large, -3 launch failure, -4 out of memory
template <class P>
int ntt_run(cudaStream_t st, NttWorkspace& ws, int field_id, uint32_t* d_data, uint32_t log_n_u, bool inverse, bool coset,
            int* launches) {
  typedef Fp<P> F;
  const int log_n = (int)log_n_u;
  if (log_n > P::TWO_ADICITY || log_n > 30) return -2;
  if (log_n == 0) return 0;  // size-1 domain: identity (size_inv = 1, g^0 = 1)
  const size_t n = (size_t)1 << log_n;
  const int dir = inverse ? 1 : 0;
  if (ntt_ensure(&ws.scratch, &ws.scratch_cap, n * F::N * 4)) return -4;
  if (ntt_ensure(&ws.tw[dir], &ws.tw_cap[dir], std::max<size_t>(n / 2, 1) * F::N * 4)) return -4;
  if (!ws.consts[dir] && cudaMalloc(&ws.consts[dir], 4096) != cudaSuccess) return -4;
  NttConsts<P>* consts = (NttConsts<P>*)ws.consts[dir];
  uint32_t* tw = (uint32_t*)ws.tw[dir];

  if (ws.key_field[dir] != field_id || ws.key_log_n[dir] != log_n) {
    k_ntt_setup<P><<<1, 32, 0, st>>>(log_n, inverse ? 1 : 0, consts);
    const uint32_t cnt = (uint32_t)(n / 2);
    const uint32_t threads = (cnt + POW_RUN - 1) / POW_RUN;
    k_build_powers<P><<<(threads + 127) / 128, 128, 0, st>>>(&consts->omega, &consts->one, cnt, tw);
    *launches += 2;
    ws.key_field[dir] = field_id; ws.key_log_n[dir] = log_n;
    ws.coset_key_field[dir] = -1;
  }
  const int lo_bits = log_n < NTT_LO_BITS ? log_n : NTT_LO_BITS;
  if (coset && (ws.coset_key_field[dir] != field_id || ws.coset_key_log_n[dir] != log_n)) {
    const uint32_t nlo = 1u << lo_bits, nhi = 1u << (log_n - lo_bits);
    if (ntt_ensure(&ws.glo[dir], &ws.glo_cap[dir], (size_t)nlo * F::N * 4)) return -4;
    if (ntt_ensure(&ws.ghi[dir], &ws.ghi_cap[dir], (size_t)nhi * F::N * 4)) return -4;
    k_build_powers<P><<<((nlo + POW_RUN - 1) / POW_RUN + 127) / 128, 128, 0, st>>>(&consts->g, &consts->one, nlo, (uint32_t*)ws.glo[dir]);
    // inverse coset: fold size_inv into the high table
    k_build_powers<P><<<((nhi + POW_RUN - 1) / POW_RUN + 127) / 128, 128, 0, st>>>(&consts->g_hi, inverse ? &consts->size_inv : &consts->one, nhi, (uint32_t*)ws.ghi[dir]);
    *launches += 2;
    ws.coset_key_field[dir] = field_id; ws.coset_key_log_n[dir] = log_n;
  }
  const uint32_t* glo = (const uint32_t*)ws.glo[dir];
  const uint32_t* ghi = (const uint32_t*)ws.ghi[dir];

  // Pass plan.  Register-only passes of 1..3 stages (k_ntt_pass) for the remainder and for small domains;
  // shared-memory tile passes (k_ntt_tile) for the rest: 2^24 = 8 + 8 + 8 (radix-4 rounds) or 9 + 9 + 6 (radix-8) instead of
  // eight radix-8 trips through HBM.  OZL_NTT_TILE=0 selects the register-only plan for comparison.
  static const bool kTile = []() { const char* e = getenv("OZL_NTT_TILE"); return !(e && e[0] == '0'); }();
  // OZL_NTT_R4: 0 = radix-8 rounds on 2048-element tiles (two CTAs per SM, 113 registers), 1 = radix-4 rounds on
  // 1024-element tiles with three CTAs per SM (68 registers), 2 = the same with four (64 registers; DEFAULT).
  // Measured at 2^24 BN254 Fr forward: 4.26 / 3.77 / 3.75 ms -- with half the data registers per thread twice the warps
  // are resident (32 per SM) and the multiplier pipe stays fed through the load / store / barrier phases of a tile.
  static const int kR4 = []() { const char* e = getenv("OZL_NTT_R4"); return e ? atoi(e) : 2; }();
  int radices[16], np = 0;
  bool tiled[16];
  const bool r4 = kR4 > 0 && kTile && log_n >= 10;   // radix-4 rounds on 1024-element tiles: passes of 4, 6 or 8 stages
  if (r4) {
    const int rem = log_n % 2;
    int L = log_n - rem;
    if (rem) { tiled[np] = false; radices[np++] = 1; }
    const int k = (L + 7) / 8;                // up to eight stages per pass: M <= 256 rows (padded rows fit), C >= 4 columns
    for (int i = 0; i < k; i++) {
      int R = ((L / 2 + (k - i) - 1) / (k - i)) * 2;
      tiled[np] = true; radices[np++] = R;
      L -= R;
    }
  } else {
    const int rem = log_n % 3;
    int L = log_n - rem;                       // multiple of 3
    const bool use_tiles = kTile && log_n >= 11;   // a tile holds 2048 elements
    if (rem) { tiled[np] = false; radices[np++] = rem; }
    if (use_tiles) {
      const int k = (L + 8) / 9;               // number of tile passes, as even as multiples of 3 allow
      for (int i = 0; i < k; i++) {
        int R = ((L / 3 + (k - i) - 1) / (k - i)) * 3;
        tiled[np] = true; radices[np++] = R;
        L -= R;
      }
    } else {
      for (int i = 0; i < L / 3; i++) { tiled[np] = false; radices[np++] = 3; }
    }
  }
  // per device, so not cached in a static: a process may hold contexts on several GPUs
  static const bool kPaired = []() { const char* e = getenv("OZL_NTT_PAIRED"); return e && e[0] == '1'; }();
  if (cudaFuncSetAttribute(k_ntt_tile<P, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, NTT_TILE_SMEM_BYTES) != cudaSuccess) return -3;
  if (cudaFuncSetAttribute(k_ntt_tile<P, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, NTT_TILE_SMEM_BYTES) != cudaSuccess) return -3;
  if (cudaFuncSetAttribute(k_ntt_tile<P, false, 2, 256, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, NTT_TILE4_SMEM_BYTES) != cudaSuccess) return -3;
  if (cudaFuncSetAttribute(k_ntt_tile<P, false, 2, 256, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, NTT_TILE4_SMEM_BYTES) != cudaSuccess) return -3;
  uint32_t* scratch = (uint32_t*)ws.scratch;
  int s = 0;
  for (int pi = 0; pi < np; pi++) {
    const int R = radices[pi];
    const bool first = pi == 0, lastp = pi == np - 1;
    const uint32_t* src = first ? d_data : scratch;
    uint32_t* dst = (lastp && np > 1) ? d_data : scratch;
    const int pre = (first && coset && !inverse) ? 1 : 0;
    int post = 0;
    if (lastp && inverse) post = coset ? 2 : 1;
    if (tiled[pi]) {
      const uint32_t blocks = (uint32_t)(n / NTT_TILE_ELEMS);
      if (r4 && kR4 == 2) k_ntt_tile<P, false, 2, 256, 4><<<(uint32_t)(n / NTT_TILE4_ELEMS), 256, NTT_TILE4_SMEM_BYTES, st>>>(src, dst, tw, log_n, s, R, pre, post, lastp, glo, ghi, &consts->size_inv);
      else if (r4) k_ntt_tile<P, false, 2, 256, 3><<<(uint32_t)(n / NTT_TILE4_ELEMS), 256, NTT_TILE4_SMEM_BYTES, st>>>(src, dst, tw, log_n, s, R, pre, post, lastp, glo, ghi, &consts->size_inv);
      else if (kPaired) k_ntt_tile<P, true><<<blocks, NTT_TILE_THREADS, NTT_TILE_SMEM_BYTES, st>>>(src, dst, tw, log_n, s, R, pre, post, lastp, glo, ghi, &consts->size_inv);
      else k_ntt_tile<P, false><<<blocks, NTT_TILE_THREADS, NTT_TILE_SMEM_BYTES, st>>>(src, dst, tw, log_n, s, R, pre, post, lastp, glo, ghi, &consts->size_inv);
    } else {
      const uint32_t threads = (uint32_t)(n >> R);
      static const int kBlock = []() { const char* e = getenv("OZL_NTT_BLOCK"); int v = e ? atoi(e) : 128; return (v == 32 || v == 64 || v == 128) ? v : 128; }();
      const uint32_t blocks = (threads + kBlock - 1) / kBlock;
      switch (R) {
        case 1: k_ntt_pass<P, 1><<<blocks, kBlock, 0, st>>>(src, dst, tw, log_n, s, pre, post, lastp, glo, ghi, &consts->size_inv); break;
        case 2: k_ntt_pass<P, 2><<<blocks, kBlock, 0, st>>>(src, dst, tw, log_n, s, pre, post, lastp, glo, ghi, &consts->size_inv); break;
        default: k_ntt_pass<P, 3><<<blocks, kBlock, 0, st>>>(src, dst, tw, log_n, s, pre, post, lastp, glo, ghi, &consts->size_inv); break;
      }
    }
    (*launches)++;
    s += R;
  }
  if (np == 1) {
    if (cudaMemcpyAsync(d_data, scratch, n * F::N * 4, cudaMemcpyDeviceToDevice, st) != cudaSuccess) return -3;
  }
  if (cudaGetLastError() != cudaSuccess) return -3;
  return 0;
}

}  // namespace ozl
