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

// Device tables are kept per direction (index 0 = forward, 1 = inverse) so a prover that
// alternates ifft / coset_fft / coset_ifft on one domain builds each table once.
struct NttWorkspace {
  void* scratch = nullptr; size_t scratch_cap = 0;
  void* tw[2] = {nullptr, nullptr}; size_t tw_cap[2] = {0, 0};
  void* glo[2] = {nullptr, nullptr}; size_t glo_cap[2] = {0, 0};
  void* ghi[2] = {nullptr, nullptr}; size_t ghi_cap[2] = {0, 0};
  void* consts[2] = {nullptr, nullptr};
  int key_field[2] = {-1, -1}, key_log_n[2] = {-1, -1};              // what tw/consts currently hold
  int coset_key_field[2] = {-1, -1}, coset_key_log_n[2] = {-1, -1};
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

  // pass plan: radix-8 passes, remainder first
  int radices[16], np = 0;
  {
    int rem = log_n % 3;
    if (rem) radices[np++] = rem;
    for (int i = 0; i < log_n / 3; i++) radices[np++] = 3;
  }
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
    const uint32_t threads = (uint32_t)(n >> R);
    static const int kBlock = []() { const char* e = getenv("OZL_NTT_BLOCK"); int v = e ? atoi(e) : 128; return (v == 32 || v == 64 || v == 128) ? v : 128; }();
    const uint32_t blocks = (threads + kBlock - 1) / kBlock;
    switch (R) {
      case 1: k_ntt_pass<P, 1><<<blocks, kBlock, 0, st>>>(src, dst, tw, log_n, s, pre, post, lastp, glo, ghi, &consts->size_inv); break;
      case 2: k_ntt_pass<P, 2><<<blocks, kBlock, 0, st>>>(src, dst, tw, log_n, s, pre, post, lastp, glo, ghi, &consts->size_inv); break;
      default: k_ntt_pass<P, 3><<<blocks, kBlock, 0, st>>>(src, dst, tw, log_n, s, pre, post, lastp, glo, ghi, &consts->size_inv); break;
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
