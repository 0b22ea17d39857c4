// Host-side runtime shared by every translation unit of libozl_b200: context, device workspace,
// bases registry, window planning, kernel sequencing and stage timing.  No CPU compute path.
#pragma once
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>   // header-only NVTX v3: no-ops unless a profiler is attached

#include <cmath>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <map>
#include <string>
#include <vector>

#include "../../include/ozl.h"
#include "params_gen.cuh"
#include "msm.cuh"
#include "msm_batch.cuh"
#include "ntt_ws.h"
#include "curve_ops.h"

using namespace ozl;
using namespace ozl_params;

namespace ozl_rt {

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
};

struct Bases {
  int curve = -1;
  size_t n = 0;
  uint32_t* d_pts = nullptr;   // factor * n * 2 * coord_u32: copy q holds 2^(pc * pWc * q) * P_i
  uint8_t* d_inf = nullptr;    // optional bitset
  int factor = 1;              // precomputed copies (ozl_msm_bases_precompute)
  int pc = 0, pWc = 0;         // window bits / windows per copy fixed when the copies were built
};

// Scratch buffers of one MSM in flight.  The context owns one; the Groth16 prover owns two more so
// that independent MSMs can run on separate streams and overlap their latency-bound tails.
struct MsmWorkspace {
  DevBuf counts, tile_sums, sorted, digits, chunk_out, window_out, misc;
  DevBuf offsets[8], partials[8];     // per point-range batch (MSM_MAX_BATCHES); [0] alone for a plain MSM
  // optional gate: the accumulation launches of an MSM on this workspace wait for this event (its digit extraction
  // and sort do not).  The Groth16 prover holds the side-stream accumulations back until the witness-map
  // transforms are through, because an accumulation kernel keeps every SM until its work runs out.
  cudaEvent_t accumulate_gate = nullptr;
  // yield: accumulation launches of this workspace use a grid of short-lived CTAs (one 32-slice batch per warp)
  // instead of a persistent one, so that concurrent streams interleave by priority at every CTA boundary
  bool yield_ctas = false;
  // Sort sharing.  Two MSMs over the SAME scalars whose bases have the same points at infinity and the same plan (the
  // b_g1 and b_g2 queries of a Groth16 key) need one digit extraction and one bucket sort between them: the producer's
  // workspace publishes its sort (publish_sort -> last_sort, ev_sorted), the consumer's names it in sort_from for the
  // duration of one MSM and skips its own digits / scan / scatter stages when the descriptions match.
  struct SortDesc {
    const uint32_t* scalars = nullptr;
    size_t n = 0, bases_n = 0;
    int c = 0, W = 0, Wc = 0;
    uint32_t L = 0;
    bool valid = false;
  } last_sort;
  bool publish_sort = false;
  cudaEvent_t ev_sorted = nullptr;
  const MsmWorkspace* sort_from = nullptr;
  // pipelined MSM (msm_run_batched): two high-priority streams that take the accumulation launches of alternate bucket
  // intervals, one event per interval ("this interval is sorted") and one per stream ("its accumulations are through")
  cudaStream_t acc_stream[2] = {nullptr, nullptr};
  cudaStream_t sort_stream = nullptr;     // greatest priority: the scatter launches of a pipelined batch
  cudaEvent_t ev_sort_ready = nullptr;    // "digits, counts and offsets of this batch are ready" (main stream -> sort stream)
  cudaEvent_t ev_interval[16] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
                                 nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t ev_acc_done[2] = {nullptr, nullptr};
  // batched-affine pair levels (msm_batch.cuh): per-level offsets, chunk table, ping-pong point
  // buffers, prefix-product scratch, and the last level's point lists
  DevBuf lvl_off, pair_tab, pair_a, pair_b, pair_pre, lvl_pts;
};

struct Groth16Pk;   // groth16.cu

struct Stage {
  std::string name;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  cudaStream_t stream = nullptr;
  int launches = 0;
};

}  // namespace ozl_rt
using namespace ozl_rt;

struct ozl_ctx {
  int device = 0;
  int sm_count = 148;
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;
  std::string last_error;
  std::map<uint32_t, Bases> bases;
  uint32_t next_handle = 1;
  int forced_c = 0;
  int batch_levels = -1;   // batched-affine pair levels: -1 = environment default (off), 0 = off, 1..8
  uint64_t launches = 0;
  bool timing = false;
  std::vector<Stage> stages;
  std::vector<Stage> event_pool;
  // workspace
  DevBuf scalars, out;
  MsmWorkspace ws;
  // double-buffered host->device pipeline of ozl_msm_submit
  cudaStream_t copy_stream = nullptr;
  DevBuf pipe_scalars[2], pipe_out[2];
  cudaEvent_t ev_h2d[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr};
  unsigned pipe_idx = 0;
  cudaEvent_t ev_batch[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // batch arrivals of a host-scalar MSM
  cudaEvent_t ev_prior = nullptr;   // "everything enqueued so far on the MSM's stream" (the staging buffer's last readers)
  NttWorkspace ntt_ws;
  // Groth16 proving keys resident on this context's device (groth16.cu); owned by the context
  std::map<uint32_t, ozl_rt::Groth16Pk*> pks;
  uint32_t next_pk = 1;
  void (*pk_deleter)(ozl_ctx*, ozl_rt::Groth16Pk*) = nullptr;
};

namespace ozl_rt {

#define CUDA_TRY(ctx, expr)                                                              \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) {                                                             \
      (ctx)->last_error = std::string(#expr) + ": " + cudaGetErrorString(_e);            \
      return _e == cudaErrorMemoryAllocation ? OZL_ERR_OOM : OZL_ERR_CUDA;               \
    }                                                                                    \
  } while (0)

#define LAUNCH_CHECK(ctx)                                                                \
  do {                                                                                   \
    (ctx)->launches++;                                                                   \
    if (!(ctx)->stages.empty() && (ctx)->timing) (ctx)->stages.back().launches++;        \
    cudaError_t _e = cudaGetLastError();                                                 \
    if (_e != cudaSuccess) {                                                             \
      (ctx)->last_error = std::string("kernel launch: ") + cudaGetErrorString(_e);       \
      return OZL_ERR_CUDA;                                                               \
    }                                                                                    \
  } while (0)

inline int ensure(ozl_ctx* ctx, DevBuf& b, size_t bytes) {
  if (bytes <= b.cap) return OZL_OK;
  if (b.p) {
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    CUDA_TRY(ctx, cudaFree(b.p));
    b.p = nullptr;
    b.cap = 0;
  }
  size_t want = bytes + bytes / 16 + 256;
  CUDA_TRY(ctx, cudaMalloc(&b.p, want));
  b.cap = want;
  return OZL_OK;
}

// Every pipeline stage is also an NVTX range (host-side span of the stage's launches), named like the
// stage timers; the Groth16 prover adds ranges named after ark-groth16's own timers (groth16.cu).
inline int stage_begin(ozl_ctx* ctx, const char* name, cudaStream_t st = nullptr, bool use_st = false) {
  nvtxRangePushA(name);
  if (!ctx->timing) return OZL_OK;
  Stage s;
  s.name = name;
  CUDA_TRY(ctx, cudaEventCreate(&s.e0));
  CUDA_TRY(ctx, cudaEventCreate(&s.e1));
  s.stream = use_st ? st : ctx->stream;
  CUDA_TRY(ctx, cudaEventRecord(s.e0, s.stream));
  ctx->stages.push_back(s);
  return OZL_OK;
}
inline int stage_end(ozl_ctx* ctx) {
  nvtxRangePop();
  if (!ctx->timing) return OZL_OK;
  CUDA_TRY(ctx, cudaEventRecord(ctx->stages.back().e1, ctx->stages.back().stream));
  return OZL_OK;
}
inline void stages_clear(ozl_ctx* ctx) {
  for (auto& s : ctx->stages) {
    if (s.e0) cudaEventDestroy(s.e0);
    if (s.e1) cudaEventDestroy(s.e1);
  }
  ctx->stages.clear();
}

#define STAGE(ctx, name)                          \
  do {                                            \
    int _r = stage_begin(ctx, name);              \
    if (_r) return _r;                            \
  } while (0)
#define STAGE_ON(ctx, name, st)                   \
  do {                                            \
    int _r = stage_begin(ctx, name, st, true);    \
    if (_r) return _r;                            \
  } while (0)
#define STAGE_END(ctx)                            \
  do {                                            \
    int _r = stage_end(ctx);                      \
    if (_r) return _r;                            \
  } while (0)

inline int coord_u32(int curve) {
  switch (curve) {
    case OZL_BLS12_381_G1: return 12;
    case OZL_BLS12_381_G2: return 24;
    case OZL_BN254_G1: return 8;
    case OZL_BN254_G2: return 16;
  }
  return 0;
}
inline int scalar_bits(int curve) { return (curve == OZL_BLS12_381_G1 || curve == OZL_BLS12_381_G2) ? 255 : 254; }

// Accumulation threads resident on the chip: 4 / 3 / 3 / 2 CTAs of 128 threads per SM for 8 / 12 / 16 / 24 limbs per coordinate.
inline uint32_t resident_acc_threads(int limbs_u32, int sm_count = 148) {
  return (uint32_t)sm_count * 128u * (limbs_u32 <= 8 ? 4u : (limbs_u32 <= 16 ? 3u : 2u));
}
// Entries per accumulation thread: 128 when there is enough work to fill the chip, shorter otherwise.
// OZL_MSM_WAVES=1 (experiment, measured and left off): pick the slice length that fills whole "waves" of resident
// threads (2^20 points x 16 windows on 12 limbs = 2.31 waves of 128-entry slices -> three waves of 104).  It does not
// pay: the kernel is bound by the multiplier pipe, not by latency, so the warps left in a partial last wave simply run
// faster -- 2^20: 5.66 -> 5.92 ms (more slices, more flushes), 2^22: 18.45 -> 17.84 ms, 2^24 and Groth16 unchanged.
inline uint32_t slice_len(uint64_t entries, uint32_t resident) {
  static const int kForceL = []() { const char* e = getenv("OZL_MSM_L"); return e ? atoi(e) : 0; }();
  if (kForceL >= 8) return (uint32_t)kForceL & ~7u;
  static const bool kWaves = []() { const char* e = getenv("OZL_MSM_WAVES"); return e && e[0] == '1'; }();
  uint64_t L;
  if (kWaves) {
    uint64_t waves = (entries + (uint64_t)128 * resident - 1) / ((uint64_t)128 * resident);
    if (waves < 1) waves = 1;
    L = (entries + waves * resident - 1) / (waves * resident);
    L = (L + 7) & ~(uint64_t)7;   // multiple of 8 entries: slices start 32-byte aligned (TMA needs 16)
  } else {
    L = (entries / ((uint64_t)148 * 384)) & ~(uint64_t)7;
  }
  if (L > 128) L = 128;
  if (L < 8) L = 8;
  return (uint32_t)L;
}

// Window width: minimise field multiplications  n*W*10 (mixed adds) + 2*Wc*B*95 + (n*W/128)*45
// (bucket reduction: two adds per bucket and one per slice partial, weighted by the efficiency
// measured for the reduce kernels on B200: c=20 beats c=22 at 2^26, c=16 beats c=13 at 2^20), where Wc = ceil(W / factor) bucket sets remain after base precomputation.
inline MsmPlan make_plan(int curve, size_t n, int forced_c, int fixed_Wc = 0, int factor = 1) {
  const int lambda = scalar_bits(curve);
  int best_c = 4;
  double best = 1e300;
  for (int c = 4; c <= 23; c++) {
    const int W = (lambda + 1 + c - 1) / c;
    const int Wc = (W + factor - 1) / factor;
    // The top window only holds lambda + 1 - c (W - 1) live bits.  When that is much less than c all
    // n points of that window fall into a handful of buckets: hot atomics in the sort and heavy
    // buckets in the reduction (measured: c = 19 is 20 % slower than c = 20 at 2^24).  Skip such widths.
    const int top_bits = lambda + 1 - c * (W - 1);
    if (c > 8 && top_bits < c - 8) continue;   // c - 8 admits c = 22 (14 live bits), measured best at 2^26 with one bucket set
    const double B = std::ldexp(1.0, c - 1);
    const double cost = (double)n * W * 10.0 + 2.0 * Wc * B * 95.0 + ((double)n * W / 128.0) * 45.0;
    if (cost < best) {
      best = cost;
      best_c = c;
    }
  }
  MsmPlan p;
  p.c = forced_c ? forced_c : best_c;
  p.W = (lambda + 1 + p.c - 1) / p.c;
  p.B = 1u << (p.c - 1);
  p.Wc = fixed_Wc ? fixed_Wc : p.W;
  p.NB = (uint32_t)p.Wc * p.B;
  p.L = slice_len((uint64_t)n * p.W, resident_acc_threads(coord_u32(curve)));   // 128 entries per accumulate thread when there is enough work to fill the chip, shorter otherwise
  uint32_t chunk = p.B / 4096;
  if (chunk < 4) chunk = 4;
  if (chunk > 16) chunk = 16;
  {
    static const int kForceChunk = []() { const char* e = getenv("OZL_MSM_CHUNK"); return e ? atoi(e) : 0; }();
    if (kForceChunk >= 1) chunk = (uint32_t)kForceChunk;
  }
  if (chunk > p.B) chunk = p.B;
  p.chunk = chunk;
  p.K = p.B / chunk;
  p.max_slots = p.NB + (uint32_t)(((uint64_t)n * p.W) / p.L) + 2;
  return p;
}

template <class Op>
int run_scan(ozl_ctx* ctx, MsmWorkspace& ws, cudaStream_t st, const uint32_t* in, uint32_t n, uint32_t* out, Op op) {
  const uint32_t tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
  int r = ensure(ctx, ws.tile_sums, (size_t)tiles * 4);
  if (r) return r;
  uint32_t* ts = (uint32_t*)ws.tile_sums.p;
  k_scan_tile_sums<<<tiles, SCAN_THREADS, 0, st>>>(in, n, ts, op);
  LAUNCH_CHECK(ctx);
  k_scan_tile_offsets<<<1, 1024, 0, st>>>(ts, tiles, out + n);
  LAUNCH_CHECK(ctx);
  k_scan_apply<<<tiles, SCAN_THREADS, 0, st>>>(in, n, ts, out, op);
  LAUNCH_CHECK(ctx);
  return OZL_OK;
}

// ---- batched-affine pair levels (kernels in msm_batch.cuh) ----------------------------------
struct BatchConfig {
  int levels;             // halving levels before the XYZZ accumulation (0 = off)
  uint32_t k;             // target additions per thread (one inversion each)
  uint64_t min_entries;   // below this many sorted entries the plain path is used
  double scratch_gb;      // bound on the per-chunk ping-pong + prefix scratch
};
inline BatchConfig batch_config() {
  static const BatchConfig cfg = []() {
    BatchConfig c{0, 1024, 1ull << 22, 12.0};
    if (const char* e = getenv("OZL_MSM_BATCH")) c.levels = atoi(e);
    if (const char* e = getenv("OZL_MSM_BATCH_K")) c.k = (uint32_t)atoi(e);
    if (const char* e = getenv("OZL_MSM_BATCH_MIN")) c.min_entries = strtoull(e, nullptr, 10);
    if (const char* e = getenv("OZL_MSM_BATCH_GB")) c.scratch_gb = atof(e);
    if (c.levels < 0) c.levels = 0;
    if (c.levels > 8) c.levels = 8;
    if (c.k < 16) c.k = 16;
    return c;
  }();
  return cfg;
}

template <class F>
int run_pair_levels(ozl_ctx* ctx, MsmWorkspace& ws, cudaStream_t st, const Bases& b, const MsmPlan& p, const BatchConfig& bc,
                    int T, size_t n, const uint32_t* sorted, const uint32_t* offsets, const uint32_t* lvl_off) {
  constexpr int AFF = 2 * F::N;
  int r;
  // chunks of about equal entry counts so that the level buffers fit the scratch budget
  const double E_bound = (double)n * p.W;
  const double bytes_per_entry = 0.5 * (AFF * 4 + F::N * 4) + 0.25 * AFF * 4;   // level-1 out + prefixes + level-2 out
  uint32_t Q = (uint32_t)std::ceil(E_bound * bytes_per_entry / (bc.scratch_gb * 1e9));
  if (Q < 1) Q = 1;
  if (Q > 4096) Q = 4096;
  const size_t tab_words = (size_t)(T + 2) * (Q + 1);
  if ((r = ensure(ctx, ws.pair_tab, tab_words * 4))) return r;
  uint32_t* d_tab = (uint32_t*)ws.pair_tab.p;
  k_pair_chunk_table<<<(Q + 1 + 127) / 128, 128, 0, st>>>(offsets, lvl_off, p.NB, Q, (uint32_t)T, d_tab);
  LAUNCH_CHECK(ctx);
  std::vector<uint32_t> tab(tab_words);
  CUDA_TRY(ctx, cudaMemcpyAsync(tab.data(), d_tab, tab_words * 4, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(ctx, cudaStreamSynchronize(st));
  auto off_at = [&](int l, uint32_t q) { return tab[(size_t)(1 + l) * (Q + 1) + q]; };

  // buffer sizes from the actual chunk extents
  size_t max_l1 = 0, max_l2 = 0;
  for (uint32_t q = 0; q < Q; q++) {
    max_l1 = std::max<size_t>(max_l1, off_at(1, q + 1) - off_at(1, q));
    if (T >= 2) max_l2 = std::max<size_t>(max_l2, off_at(2, q + 1) - off_at(2, q));
  }
  const size_t E_T = off_at(T, Q);
  // odd levels write buffer a, even levels buffer b; the last level writes lvl_pts
  const size_t a_entries = T >= 2 ? max_l1 : 0, b_entries = T >= 3 ? max_l2 : 0;
  if ((r = ensure(ctx, ws.pair_a, std::max<size_t>(a_entries, 1) * AFF * 4))) return r;
  if ((r = ensure(ctx, ws.pair_b, std::max<size_t>(b_entries, 1) * AFF * 4))) return r;
  if ((r = ensure(ctx, ws.lvl_pts, std::max<size_t>(E_T, 1) * AFF * 4))) return r;
  const int per_sm = (F::N <= 8 ? 4 : (F::N <= 12 ? 3 : 2)) * 128;
  const size_t resident = (size_t)ctx->sm_count * per_sm;
  // prefix scratch: k * T_total elements; T_total < n_out / k + 129 and k <= max(bc.k, 16) (see the k choice below)
  if ((r = ensure(ctx, ws.pair_pre, (max_l1 + (size_t)129 * std::max<uint32_t>(bc.k, 16) + 128) * F::N * 4))) return r;
  uint32_t* lvl_pts = (uint32_t*)ws.lvl_pts.p;
  uint32_t* pre = (uint32_t*)ws.pair_pre.p;

  for (uint32_t q = 0; q < Q; q++) {
    const uint32_t g_lo = tab[q], g_hi = tab[q + 1];
    if (g_hi == g_lo) continue;
    for (int l = 1; l <= T; l++) {
      const uint32_t o_begin = off_at(l, q), o_end = off_at(l, q + 1);
      if (o_end == o_begin) continue;
      const uint32_t n_out = o_end - o_begin;
      // additions per thread: whole waves of uniformly loaded threads, about bc.k each
      uint32_t waves = (uint32_t)((n_out + (uint64_t)bc.k * resident - 1) / ((uint64_t)bc.k * resident));
      uint32_t k = (uint32_t)((n_out + (uint64_t)waves * resident - 1) / ((uint64_t)waves * resident));
      if (k < 16) k = 16;
      const uint32_t nvt = (n_out + k - 1) / k;
      const uint32_t grid = (nvt + 127) / 128;
      const uint32_t* off_in = l == 1 ? offsets : lvl_off + (size_t)(l - 2) * (p.NB + 1);
      const uint32_t* off_out = lvl_off + (size_t)(l - 1) * (p.NB + 1);
      const uint32_t* src = l == 1 ? b.d_pts : (const uint32_t*)((l & 1) ? ws.pair_b.p : ws.pair_a.p);
      uint32_t* dst = l == T ? lvl_pts : (uint32_t*)((l & 1) ? ws.pair_a.p : ws.pair_b.p);
      const uint32_t dst_base = l == T ? 0u : o_begin;
      const uint32_t in_base = off_at(l - 1, q);
      if (l == 1)
        k_pair_level<F, true><<<grid, 128, 0, st>>>(src, sorted, off_in, off_out, g_lo, g_hi, o_begin, o_end, in_base, dst_base, k, pre, dst);
      else
        k_pair_level<F, false><<<grid, 128, 0, st>>>(src, sorted, off_in, off_out, g_lo, g_hi, o_begin, o_end, in_base, dst_base, k, pre, dst);
      LAUNCH_CHECK(ctx);
    }
  }
  return OZL_OK;
}

// One MSM = J >= 1 point-range BATCHES that share the bucket sets.  Every batch is digit-extracted,
// sorted and accumulated on its own (into its own region of slice partials); the bucket fold then sums a
// bucket's partials over all regions and the reduction runs once.  J = 1 is the plain pipeline.  J > 1
// exists for host scalars: batch j is accumulated while batch j+1 is still crossing PCIe, so the
// host->device copy of an `ozl_msm` call hides under the accumulation instead of preceding it
// (the copy is issued HERE, batch by batch, so that it also overlaps when the caller's memory is
// pageable and cudaMemcpyAsync blocks the host thread).
static constexpr int MSM_MAX_BATCHES = 8;
static constexpr int MSM_MAX_INTERVALS = 16;   // accumulation launches of one batch of the pipelined MSM
// ws.misc words: [0] slice counter of a whole-array accumulation, [8] input-error flags,
// [16 + j * MSM_MAX_INTERVALS + k] slice counter of interval k of batch j
static constexpr size_t MISC_BYTES = (16 + (size_t)MSM_MAX_BATCHES * MSM_MAX_INTERVALS) * 4;

// Pipelined MSM: streams and events of a workspace, created on first use.
inline int ensure_pipe(ozl_ctx* ctx, MsmWorkspace& ws) {
  if (ws.acc_stream[0]) return OZL_OK;
  int lo = 0, hi = 0;
  CUDA_TRY(ctx, cudaDeviceGetStreamPriorityRange(&lo, &hi));   // hi = numerically lowest = greatest priority
  static const int kPrio = []() { const char* e = getenv("OZL_PIPE_PRIO"); return e ? atoi(e) : 1; }();   // 0: all streams at the default priority
  // The SORT stream gets the greatest priority and its co-running launches a grid that needs only the registers the
  // accumulation leaves free, so they are placed at once; the accumulation streams keep the default (lowest) priority,
  // so the pending CTAs of the next interval never hold the sort back (measured the other way round: a pending
  // high-priority accumulation launch starved the scatter of the following interval for 60 ms).
  CUDA_TRY(ctx, cudaStreamCreateWithPriority(&ws.sort_stream, cudaStreamNonBlocking, kPrio ? hi : lo));
  CUDA_TRY(ctx, cudaEventCreateWithFlags(&ws.ev_sort_ready, cudaEventDisableTiming));
  for (int i = 0; i < 2; i++) {
    CUDA_TRY(ctx, cudaStreamCreateWithPriority(&ws.acc_stream[i], cudaStreamNonBlocking, lo));
    CUDA_TRY(ctx, cudaEventCreateWithFlags(&ws.ev_acc_done[i], cudaEventDisableTiming));
  }
  for (int k = 0; k < MSM_MAX_INTERVALS; k++) CUDA_TRY(ctx, cudaEventCreateWithFlags(&ws.ev_interval[k], cudaEventDisableTiming));
  return OZL_OK;
}

struct MsmBatches {
  int J = 1;
  size_t first[MSM_MAX_BATCHES] = {0};
  size_t count[MSM_MAX_BATCHES] = {0};
  // host source (optional): batch j is copied to d_scalars + first[j] on copy_stream, then awaited by the MSM's stream
  const uint64_t* h_scalars = nullptr;
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t* ev = nullptr;            // J events
};

template <class F>
int msm_run_batched(ozl_ctx* ctx, MsmWorkspace& ws, cudaStream_t st, const Bases& b, const uint32_t* d_scalars, size_t n,
                    const MsmBatches& mb, uint32_t* d_out) {
  constexpr int XY = 4 * F::N;
  const MsmPlan p = b.factor > 1 ? make_plan(b.curve, n, b.pc, b.pWc) : make_plan(b.curve, n, ctx->forced_c);
  if ((uint64_t)n * p.W >= 0xffffffffull || (uint64_t)b.n * b.factor >= 0x7fffffffull) {
    ctx->last_error = "msm: n * windows must stay below 2^32 and n * copies below 2^31 (32-bit sort entries)";
    return OZL_ERR_ARG;
  }
  // Batched-affine pair levels (msm_batch.cuh): T halving levels before the XYZZ accumulation (single batch only).
  BatchConfig bc = batch_config();
  if (ctx->batch_levels >= 0) {      // explicit request: honoured at every size
    bc.levels = ctx->batch_levels;
    bc.min_entries = 0;
  }
  const int T = (mb.J == 1 && n && (uint64_t)n * p.W >= bc.min_entries) ? bc.levels : 0;
  const int J = mb.J;
  size_t max_nj = 0;
  uint32_t Lj[MSM_MAX_BATCHES], heavy[MSM_MAX_BATCHES];
  for (int j = 0; j < J; j++) {
    max_nj = std::max(max_nj, mb.count[j]);
    Lj[j] = J == 1 ? p.L : slice_len((uint64_t)mb.count[j] * p.W, resident_acc_threads(F::N, ctx->sm_count));
    // a yielding workspace hands its SMs back every slice: 64-entry slices halve that interval (0.8 ms on G1, 1.1 ms on
    // G2 at the Groth16 sizes) for twice the slice partials -- measured 22.07 -> 21.79 ms per proof
    if (ws.yield_ctas && Lj[j] > 64) Lj[j] = 64;
  }
  int r;
  if ((r = ensure(ctx, ws.counts, (size_t)p.NB * 4))) return r;
  if ((r = ensure(ctx, ws.sorted, std::max<size_t>(max_nj * p.W, 1) * 4))) return r;
  if ((r = ensure(ctx, ws.digits, std::max<size_t>(max_nj * p.W, 1) * 4))) return r;
  for (int j = 0; j < J; j++) {
    const size_t slots = (size_t)p.NB + ((size_t)mb.count[j] * p.W) / Lj[j] + 2;
    if ((r = ensure(ctx, ws.offsets[j], ((size_t)p.NB + 1) * 4))) return r;
    if ((r = ensure(ctx, ws.partials[j], slots * XY * 4))) return r;
  }
  if ((r = ensure(ctx, ws.chunk_out, ((size_t)p.Wc * p.K + (size_t)p.Wc * 64) * XY * 4))) return r;
  if ((r = ensure(ctx, ws.window_out, (size_t)p.Wc * XY * 4))) return r;
  if ((r = ensure(ctx, ws.misc, MISC_BYTES))) return r;
  uint32_t* lvl_off = nullptr;
  if (T) {
    if ((r = ensure(ctx, ws.lvl_off, (size_t)T * ((size_t)p.NB + 1) * 4))) return r;
    lvl_off = (uint32_t*)ws.lvl_off.p;
  }

  uint32_t* counts = (uint32_t*)ws.counts.p;
  uint32_t* sorted = (uint32_t*)ws.sorted.p;
  uint32_t* digits = (uint32_t*)ws.digits.p;
  uint32_t* chunk_out = (uint32_t*)ws.chunk_out.p;
  uint32_t* window_out = (uint32_t*)ws.window_out.p;
  uint32_t* work_counter = (uint32_t*)ws.misc.p;     // [0] slice counter of the accumulation, [8] input-error flags
  const int grid_io = ctx->sm_count * 8;
  FoldRegions regions;
  regions.J = J;
  bool pipe_pending = false;   // accumulation launches outstanding on the side streams
  CUDA_TRY(ctx, cudaMemsetAsync(work_counter, 0, MISC_BYTES, st));

  for (int j = 0; j < J; j++) {
    const size_t first = mb.first[j], nj = mb.count[j];
    uint32_t* offsets = (uint32_t*)ws.offsets[j].p;
    uint32_t* partials = (uint32_t*)ws.partials[j].p;
    const uint32_t* sc = d_scalars + first * 8;
    bool reuse_sort = false;
    if (ws.sort_from && J == 1 && !T && !mb.h_scalars) {
      const MsmWorkspace::SortDesc& d = ws.sort_from->last_sort;
      reuse_sort = d.valid && ws.sort_from->ev_sorted && d.scalars == d_scalars && d.n == n && d.bases_n == b.n && d.c == p.c &&
                   d.W == p.W && d.Wc == p.Wc && d.L == Lj[0];
    }
    ws.last_sort.valid = false;
    if (mb.h_scalars) {
      if (nj) CUDA_TRY(ctx, cudaMemcpyAsync((void*)sc, mb.h_scalars + first * 4, nj * 32, cudaMemcpyHostToDevice, mb.copy_stream));
      CUDA_TRY(ctx, cudaEventRecord(mb.ev[j], mb.copy_stream));
      CUDA_TRY(ctx, cudaStreamWaitEvent(st, mb.ev[j], 0));
    }

    if (reuse_sort) {
      // the producer's offsets and sorted indices stand in for this MSM's own (its accumulation reads them only)
      CUDA_TRY(ctx, cudaStreamWaitEvent(st, ws.sort_from->ev_sorted, 0));
      offsets = (uint32_t*)ws.sort_from->offsets[0].p;
      sorted = (uint32_t*)ws.sort_from->sorted.p;
    } else {
    STAGE_ON(ctx, "digits_count", st);
    CUDA_TRY(ctx, cudaMemsetAsync(counts, 0, (size_t)p.NB * 4, st));
    if (j) CUDA_TRY(ctx, cudaMemsetAsync(work_counter, 0, 4, st));
    k_count<<<grid_io, 256, 0, st>>>(sc, b.d_inf, (uint32_t)first, (uint32_t)nj, p.c, p.W, p.Wc, p.B, counts, digits, work_counter + 8);
    LAUNCH_CHECK(ctx);
    STAGE_END(ctx);

    STAGE_ON(ctx, "scan", st);
    if ((r = run_scan(ctx, ws, st, counts, p.NB, offsets, ScanIdentity{1}))) return r;
    // level-l lists hold ceil(count / 2^l) entries per bucket (counts are consumed by the scatter below)
    for (int l = 1; l <= T; l++)
      if ((r = run_scan(ctx, ws, st, counts, p.NB, lvl_off + (size_t)(l - 1) * (p.NB + 1), ScanCeilDiv{1u << l}))) return r;
    STAGE_END(ctx);
    }   // !reuse_sort

    // ---- scatter + accumulate --------------------------------------------------------------------------------
    // The counting sort runs bucket-set-major, then one bucket RANGE of 2^18 buckets (8 MB of open 32-byte sectors)
    // at a time so that a bucket's region is completed while its sectors are still in L2 (measured at 2^26, c = 22,
    // 2^21 buckets: 34.2 / 27.9 / 26.9 / 23.3 ms with 1 / 2 / 4 / 8 ranges, re-reading the digits included), then the
    // windows that share the set.  After group (set s, range q) the bucket interval [s B + lo_q, s B + hi_q) of the
    // sorted array is final, and the intervals complete in increasing order.
    //
    // PIPELINED (large MSMs on the G1 curves): the accumulation of a finished interval starts at once, on one of two
    // high-priority side streams, while this stream goes on scattering the next interval.  The interval kernels are
    // one warp per CTA and capped at 160 (120) registers, which leaves 4096 registers per SM: the four scatter warps
    // per SM that fit there are latency-bound on L2 atomics and take next to nothing from the multiplier pipe the
    // accumulation is bound by.  Alternating streams let interval k + 1 take over SMs warp by warp as interval k runs
    // out of slices.  Only the first interval's sort stays exposed.
    static const int kForceParts = []() { const char* e = getenv("OZL_MSM_SCATTER_PARTS"); return e ? atoi(e) : 0; }();
    // OZL_MSM_PIPE: 0 = off (DEFAULT: measured slower on B200, see below), -1 = on from ~2^24 points, k > 0 = on at any size with at most k intervals
    static const int kPipe = []() { const char* e = getenv("OZL_MSM_PIPE"); return e ? atoi(e) : 0; }();
    static const int kPipeRegs = []() { const char* e = getenv("OZL_PIPE_REGS"); return e ? atoi(e) : 0; }();   // interval kernels capped 0 / 16 / 32 registers lower
    static const int kScatterU = []() { const char* e = getenv("OZL_SCATTER_U"); return e ? atoi(e) : 1; }();
    static const int kPipeStreams = []() { const char* e = getenv("OZL_PIPE_STREAMS"); return e ? atoi(e) : 1; }();
    static const int kPipeLayout = []() { const char* e = getenv("OZL_PIPE_LAYOUT"); return e ? atoi(e) : 1; }();   // 1 = doubling interval lengths, 0 = equal
    static const int kPipeSpread = []() { const char* e = getenv("OZL_PIPE_SPREAD"); return e ? atoi(e) : 1; }();   // 1 = one wide scatter CTA per SM
    static const int kPipeScatterCtas = []() { const char* e = getenv("OZL_PIPE_SCATTER_CTAS"); return e ? std::max(1, atoi(e)) : 4; }();   // co-resident 128-thread scatter CTAs per SM
    static const int kScatterGrid = []() { const char* e = getenv("OZL_SCATTER_GRID"); return e ? atoi(e) : 0; }();   // diagnostic: 128-thread CTAs per SM
    uint32_t parts = kForceParts > 0 ? (uint32_t)kForceParts : std::min<uint32_t>(16u, std::max<uint32_t>(1u, p.B >> 18));
    if (parts > p.B) parts = p.B;
    const uint32_t span = (p.B + parts - 1) / parts;
    const uint32_t groups = (uint32_t)p.Wc * parts;
    static const bool use_tma = []() { const char* e = getenv("OZL_ACC_TMA"); return !(e && e[0] == '0'); }();
    static const int acc_env = []() { const char* e = getenv("OZL_ACC_MODE"); return e ? atoi(e) : -1; }();
    // BN254 G1 (8 limbs, ~45 KB inlined) is 4 % faster inlined + fused; the G2 curves take the lazily reduced Fq2 products at
    // two CTAs per SM (mode 18 for BN254 G2, 8 for BLS12-381 G2: accumulate 9.50 -> 9.26 ms and 23.25 -> 22.31 ms at 2^20 --
    // 18 % fewer wide multiplies buy only 2-4 % because these kernels are bound by latency at 8 warps per SM, not by the pipe)
    static const int acc_env_g2 = []() { const char* e = getenv("OZL_ACC_MODE_G2"); return e ? atoi(e) : -1; }();   // G2 curves only (A/B inside a Groth16 proof)
    const int acc_mode = (F::N >= 16 && acc_env_g2 >= 0) ? acc_env_g2 : acc_env >= 0 ? acc_env : (F::N == 8 ? 6 : (F::N == 16 ? 18 : (F::N == 24 ? 8 : 5)));
    // intervals = accumulation launches of this batch: enough slices per launch for several full waves of the chip
    uint32_t intervals = 1;
    if (F::N <= 12 && !T && use_tma && acc_env < 0 && kPipe != 0 && groups > 1 && nj && !reuse_sort) {
      const uint64_t slices = (uint64_t)nj * p.W / Lj[j];
      const uint64_t resident = (uint64_t)ctx->sm_count * AccPipe<F>::WARPS_PER_SM * 32;
      intervals = std::min<uint32_t>(groups, kPipe > 0 ? (uint32_t)kPipe : 8u);
      if (kPipe < 0) intervals = (uint32_t)std::min<uint64_t>(intervals, slices / (4 * resident));   // OZL_MSM_PIPE=k forces it at any size (tests)
      if (intervals > MSM_MAX_INTERVALS) intervals = MSM_MAX_INTERVALS;
      if (intervals < 2) intervals = 1;
    }
    const bool pipe = intervals > 1;
    if (pipe && (r = ensure_pipe(ctx, ws))) return r;
    const cudaStream_t ss = pipe ? ws.sort_stream : st;   // stream of the scatter launches
    // interval k ends after group gend[k]: the first interval is ONE group (its sort is the exposed one)
    uint32_t gend[MSM_MAX_INTERVALS];
    if (pipe && kPipeLayout == 1) {
      // doubling: 1, 1, 2, 4, ... groups per interval.  Every launch boundary costs a drain bubble (the last slices of
      // a launch finish up to one slice = 2.4 ms apart), and a sort that co-runs at a quarter of its speed is
      // through long before the accumulation is, so late intervals can be long.
      uint32_t k = 0, e = 1;
      while (k + 1 < std::min<uint32_t>(intervals, MSM_MAX_INTERVALS) && e < groups) { gend[k++] = e; e = std::min(groups, e * 2); }
      gend[k++] = groups;
      intervals = k;
    } else {
      for (uint32_t k = 0; k < intervals; k++)
        gend[k] = intervals == 1 ? groups : (k == 0 ? 1u : 1u + (uint32_t)(((uint64_t)(groups - 1) * k) / (intervals - 1)));
    }
    uint32_t* partials_j = partials;
    // OZL_PIPE_TRACE=1 (diagnostic): timeline of the scatter groups and the interval launches, printed to stderr
    static const bool kTrace = []() { const char* e = getenv("OZL_PIPE_TRACE"); return e && e[0] == '1'; }();
    std::vector<cudaEvent_t> tr_g, tr_a0, tr_a1;
    cudaEvent_t tr_t0 = nullptr;
    auto tr_rec = [&](std::vector<cudaEvent_t>& v, cudaStream_t q) {
      cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, q); v.push_back(e);
    };
    if (kTrace && pipe) { cudaEventCreate(&tr_t0); cudaEventRecord(tr_t0, st); }   // before the sort stream's first launch
    auto launch_interval = [&](uint32_t k, uint32_t g_lo, uint32_t g_hi) -> int {
      cudaStream_t sa = ws.acc_stream[kPipeStreams == 1 ? 0 : (k & 1)];
      CUDA_TRY(ctx, cudaEventRecord(ws.ev_interval[k], ss));
      CUDA_TRY(ctx, cudaStreamWaitEvent(sa, ws.ev_interval[k], 0));
      if (k < 2 && ws.accumulate_gate) CUDA_TRY(ctx, cudaStreamWaitEvent(sa, ws.accumulate_gate, 0));
      if (k == 0) STAGE_ON(ctx, "accumulate", sa);
      if (kTrace) tr_rec(tr_a0, sa);
      uint32_t* counter = work_counter + 16 + (size_t)j * MSM_MAX_INTERVALS + k;
      const int grid = ctx->sm_count * AccPipe<F>::WARPS_PER_SM;
      if constexpr (F::N <= 12) {
        constexpr int R0 = AccPipe<F>::MAXREG;
        constexpr int M = F::N <= 8 ? 6 : 5;   // shipped product variant of the field (see acc_mode above)
        // the SM must be in its largest shared-memory configuration for the pinned scatter CTA to fit beside these warps
        static const bool carve_set = []() {
          cudaFuncSetAttribute(k_accumulate_tma_piece<F, M, R0>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
          cudaFuncSetAttribute(k_accumulate_tma_piece<F, M, R0 - 16>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
          cudaFuncSetAttribute(k_accumulate_tma_piece<F, M, R0 - 32>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
          return true;
        }();
        (void)carve_set;
        if (kPipeRegs == 2) k_accumulate_tma_piece<F, M, R0 - 32><<<grid, 32, 0, sa>>>(b.d_pts, sorted, offsets, p.NB, Lj[j], counter, partials_j, g_lo, g_hi);
        else if (kPipeRegs == 1) k_accumulate_tma_piece<F, M, R0 - 16><<<grid, 32, 0, sa>>>(b.d_pts, sorted, offsets, p.NB, Lj[j], counter, partials_j, g_lo, g_hi);
        else k_accumulate_tma_piece<F, M, R0><<<grid, 32, 0, sa>>>(b.d_pts, sorted, offsets, p.NB, Lj[j], counter, partials_j, g_lo, g_hi);
      } else {
        (void)counter; (void)grid; (void)g_lo; (void)g_hi;
        ctx->last_error = "msm: interval launches exist for the G1 curves only";
        return OZL_ERR_ARG;
      }
      LAUNCH_CHECK(ctx);
      if (kTrace) tr_rec(tr_a1, sa);
      return OZL_OK;
    };

    if (pipe_pending) {   // `sorted` is about to be overwritten: the previous batch's accumulations must be through
      CUDA_TRY(ctx, cudaStreamWaitEvent(st, ws.ev_acc_done[0], 0));
      CUDA_TRY(ctx, cudaStreamWaitEvent(st, ws.ev_acc_done[1], 0));
    }
    if (pipe) {
      CUDA_TRY(ctx, cudaEventRecord(ws.ev_sort_ready, st));
      CUDA_TRY(ctx, cudaStreamWaitEvent(ss, ws.ev_sort_ready, 0));
    }
    STAGE_ON(ctx, "scatter", ss);
    {
      uint32_t k = 0, g_lo = 0;
      for (uint32_t grp = 0; grp < groups && nj && !reuse_sort; grp++) {
        const uint32_t s_ = grp / parts, q = grp % parts;
        const uint32_t lo = q * span, hi = std::min<uint32_t>(p.B, lo + span);
        for (int w = (int)s_; w < p.W; w += p.Wc) {
          const uint32_t* dw = digits + (size_t)w * nj;
          const uint32_t idx0 = (uint32_t)((size_t)(w / p.Wc) * b.n + first);
          if (pipe || kScatterGrid > 0) {
            // 128-thread CTAs.  The first interval's groups run alone and take the whole chip; the others run under an
            // accumulation launch, in a grid of exactly the CTAs that fit into the registers it leaves free.
            if (pipe && k > 0 && kPipeSpread && kScatterGrid <= 0) {
              // ONE CTA of 128 x kPipeScatterCtas threads per SM, pinned there by a dynamic shared memory request that
              // a second copy cannot meet (the block scheduler otherwise packs sixteen small CTAs onto a quarter of the
              // SMs and the accumulation loses those SMs outright: measured zero-sum) but that fits beside the twelve
              // accumulation warps (12 x (8200 + 1024) + 118 KB + 1024 <= 228 KB).
              constexpr size_t kPin = 115968;   // two copies (+ 1 KB each) exceed the SM's 228 KB; one fits beside the accumulation's 12 x 9.25 KB
              static const bool attr_set = []() {
                cudaFuncSetAttribute(k_scatter_window<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPin);
                cudaFuncSetAttribute(k_scatter_window<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPin);
                cudaFuncSetAttribute(k_scatter_window<1>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
                cudaFuncSetAttribute(k_scatter_window<2>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
                return true;
              }();
              (void)attr_set;
              const int threads = 128 * std::min(4, kPipeScatterCtas);
              if (kScatterU == 2) k_scatter_window<2><<<ctx->sm_count, threads, kPin, ss>>>(dw, (uint32_t)nj, s_ * p.B, idx0, offsets, counts, sorted, lo, hi);
              else k_scatter_window<1><<<ctx->sm_count, threads, kPin, ss>>>(dw, (uint32_t)nj, s_ * p.B, idx0, offsets, counts, sorted, lo, hi);
              LAUNCH_CHECK(ctx);
              continue;
            }
            const int g128 = kScatterGrid > 0 ? ctx->sm_count * kScatterGrid : (k == 0 ? grid_io * 2 : ctx->sm_count * kPipeScatterCtas);
            if (kScatterU == 2) k_scatter_window<2><<<g128, 128, 0, ss>>>(dw, (uint32_t)nj, s_ * p.B, idx0, offsets, counts, sorted, lo, hi);
            else k_scatter_window<1><<<g128, 128, 0, ss>>>(dw, (uint32_t)nj, s_ * p.B, idx0, offsets, counts, sorted, lo, hi);
          } else {
            k_scatter_window<1><<<grid_io, 256, 0, ss>>>(dw, (uint32_t)nj, s_ * p.B, idx0, offsets, counts, sorted, lo, hi);
          }
          LAUNCH_CHECK(ctx);
        }
        if (kTrace && pipe) tr_rec(tr_g, ss);
        if (pipe && grp + 1 == gend[k]) {
          const uint32_t g_hi = grp + 1 == groups ? p.NB : s_ * p.B + hi;
          if (k == 0) STAGE_END(ctx);   // the "scatter" stage times the exposed part: the first interval's sort
          if ((r = launch_interval(k, g_lo, g_hi))) return r;
          g_lo = g_hi;
          k++;
        }
      }
    }
    if (!pipe) STAGE_END(ctx);
    if (ws.publish_sort && J == 1 && !T && !pipe && !reuse_sort && !mb.h_scalars) {
      if (!ws.ev_sorted) CUDA_TRY(ctx, cudaEventCreateWithFlags(&ws.ev_sorted, cudaEventDisableTiming));
      CUDA_TRY(ctx, cudaEventRecord(ws.ev_sorted, st));
      ws.last_sort.scalars = d_scalars; ws.last_sort.n = n; ws.last_sort.bases_n = b.n;
      ws.last_sort.c = p.c; ws.last_sort.W = p.W; ws.last_sort.Wc = p.Wc; ws.last_sort.L = Lj[0];
      ws.last_sort.valid = true;
    }
    // the next batch's digit extraction overwrites `digits` and `counts`: the main stream waits for the last scatter
    if (pipe) CUDA_TRY(ctx, cudaStreamWaitEvent(st, ws.ev_interval[intervals - 1], 0));

    const uint32_t* acc_offsets = offsets;   // offsets of the lists the XYZZ accumulation walks
    uint32_t acc_L = Lj[j];
    if (T) {
      STAGE_ON(ctx, "pair_levels", st);
      if ((r = run_pair_levels<F>(ctx, ws, st, b, p, bc, T, nj, sorted, offsets, lvl_off))) return r;
      STAGE_END(ctx);
      acc_offsets = lvl_off + (size_t)(T - 1) * (p.NB + 1);
      acc_L = slice_len(((uint64_t)nj * p.W) >> T, resident_acc_threads(F::N, ctx->sm_count));
      const size_t slots = (size_t)p.NB + ((((size_t)nj * p.W) >> T) + p.NB) / acc_L + 2;
      if ((r = ensure(ctx, ws.partials[j], slots * XY * 4))) return r;
      partials = (uint32_t*)ws.partials[j].p;
    }

    if (pipe) {
      // the stage ends when both side streams are through; the main stream only joins them where it has to
      // (before the next batch's scatter, above, and before the bucket reduction, below)
      for (int i = 0; i < 2; i++) CUDA_TRY(ctx, cudaEventRecord(ws.ev_acc_done[i], ws.acc_stream[i]));
      if (ctx->timing) {
        const int last_i = kPipeStreams == 1 ? 0 : (int)((intervals - 1) & 1);
        cudaStream_t last = ws.acc_stream[last_i];
        CUDA_TRY(ctx, cudaStreamWaitEvent(last, ws.ev_acc_done[last_i ^ 1], 0));
        nvtxRangePop();
        for (auto it = ctx->stages.rbegin(); it != ctx->stages.rend(); ++it)
          if (it->name == "accumulate") { CUDA_TRY(ctx, cudaEventRecord(it->e1, last)); break; }
        CUDA_TRY(ctx, cudaEventRecord(ws.ev_acc_done[last_i], last));
      } else {
        nvtxRangePop();
      }
      pipe_pending = true;
      if (kTrace) {
        cudaStreamSynchronize(ws.acc_stream[0]);
        cudaStreamSynchronize(ws.acc_stream[1]);
        cudaStreamSynchronize(st);
        auto at = [&](cudaEvent_t e) { float ms = 0; cudaEventElapsedTime(&ms, tr_t0, e); return ms; };
        fprintf(stderr, "[pipe trace] batch %d: %u groups, %u intervals, L %u\n  scatter group ends:", j, groups, intervals, Lj[j]);
        for (auto e : tr_g) fprintf(stderr, " %.2f", at(e));
        fprintf(stderr, "\n  interval [ready, end]:");
        for (size_t i = 0; i < tr_a0.size(); i++) fprintf(stderr, " [%.2f, %.2f]", at(tr_a0[i]), at(tr_a1[i]));
        fprintf(stderr, "\n");
        for (auto e : tr_g) cudaEventDestroy(e);
        for (auto e : tr_a0) cudaEventDestroy(e);
        for (auto e : tr_a1) cudaEventDestroy(e);
        cudaEventDestroy(tr_t0);
      }
    } else {
    if (ws.accumulate_gate) CUDA_TRY(ctx, cudaStreamWaitEvent(st, ws.accumulate_gate, 0));
    STAGE_ON(ctx, "accumulate", st);
    if (T) {
      k_accumulate<F, true><<<ctx->sm_count * 4, 128, 0, st>>>((const uint32_t*)ws.lvl_pts.p, nullptr, acc_offsets, p.NB, acc_L, work_counter, partials);
    } else {
      // grid: persistent (one CTA per resident slot, warps loop until the counter runs out) or, for a workspace that
      // yields, one CTA per four 32-slice batches, each warp taking exactly one
      static const int kYield = []() { const char* e = getenv("OZL_ACC_YIELD"); return e ? atoi(e) : -1; }();   // 0 / 1 override the workspace flag
      const bool yield = kYield >= 0 ? kYield != 0 : ws.yield_ctas;
      const uint64_t acc_slices = ((uint64_t)nj * p.W + acc_L - 1) / acc_L + 1;
      const uint32_t acc_grid = yield ? (uint32_t)std::max<uint64_t>(1, ((acc_slices + 31) / 32 + 3) / 4) : (uint32_t)ctx->sm_count * 4;
      const uint32_t acc_batches = yield ? 1u : 0xffffffffu;
      // TMA-staged index stream by default; OZL_ACC_TMA=0 selects the plain global-load variant
      // Field products of the hot loop (XYZZ::add_mixed_calls): 0 = all inlined, 1 = out-of-line mul (operands by
      // value), 2 = out-of-line paired mul, 3 = 1 + dedicated squaring, 4 = 3 with Karatsuba products, 5 = 3 with
      // y3 = r (q - x3) - y p3 as one fused dual product (single reduction), 6 = inlined + fused y3.
      // Measured on B200, accumulate stage at 2^26 BLS12-381 G1: 296.9 / 282.9 / 284.3 / 277.5 / 310.4 / 260.4 ms
      // for 0 .. 5: the inlined body (~100 KB of SASS) misses the instruction cache (ncu: icc hit rate 83.5 % ->
      // 99.998 %, fmaheavy 85.3 % -> 92.9 %), and the fused y3 saves N^2 of the 20 N^2 wide multiplies of an addition.
      if (use_tma && acc_mode == 1) k_accumulate_tma<F, 1><<<ctx->sm_count * 4, 128, 0, st>>>(b.d_pts, sorted, offsets, p.NB, acc_L, work_counter, partials);
      else if (use_tma && acc_mode == 2) k_accumulate_tma<F, 2><<<ctx->sm_count * 4, 128, 0, st>>>(b.d_pts, sorted, offsets, p.NB, acc_L, work_counter, partials);
      else if (use_tma && acc_mode == 3) k_accumulate_tma<F, 3><<<ctx->sm_count * 4, 128, 0, st>>>(b.d_pts, sorted, offsets, p.NB, acc_L, work_counter, partials);
      else if (use_tma && acc_mode == 13 && F::N == 12) k_accumulate_tma<F, 3, (F::N == 12 ? 4 : 2)><<<ctx->sm_count * 4, 128, 0, st>>>(b.d_pts, sorted, offsets, p.NB, acc_L, work_counter, partials);   // experiment: 4 CTAs per SM at 128 registers
      else if (use_tma && acc_mode == 6 && F::N <= 12) k_accumulate_tma<F, (F::N <= 12 ? 6 : 0)><<<acc_grid, 128, 0, st>>>(b.d_pts, sorted, offsets, p.NB, acc_L, work_counter, partials, acc_batches);
      else if (use_tma && acc_mode == 15 && F::N >= 16) {   // A/B: the other CTA count for the G2 curves (BN254 G2 ships 3, BLS12-381 G2 ships 2)
        if constexpr (F::N >= 16) k_accumulate_tma<F, 5, (F::N == 16 ? 2 : 3)><<<acc_grid, 128, 0, st>>>(b.d_pts, sorted, offsets, p.NB, acc_L, work_counter, partials, acc_batches);
      }
      else if (use_tma && (acc_mode == 8 || acc_mode == 18) && F::N >= 16) {   // lazily reduced Fq2 products (G2 curves); 18 = with the other CTA count
        if constexpr (F::N >= 16) {
          if (acc_mode == 8) k_accumulate_tma<F, 8><<<acc_grid, 128, 0, st>>>(b.d_pts, sorted, offsets, p.NB, acc_L, work_counter, partials, acc_batches);
          else k_accumulate_tma<F, 8, (F::N == 16 ? 2 : 3)><<<acc_grid, 128, 0, st>>>(b.d_pts, sorted, offsets, p.NB, acc_L, work_counter, partials, acc_batches);
        }
      }
      else if (use_tma && acc_mode == 7) k_accumulate_tma<F, 7><<<ctx->sm_count * 4, 128, 0, st>>>(b.d_pts, sorted, offsets, p.NB, acc_L, work_counter, partials);
      else if (use_tma && acc_mode == 5) k_accumulate_tma<F, 5><<<acc_grid, 128, 0, st>>>(b.d_pts, sorted, offsets, p.NB, acc_L, work_counter, partials, acc_batches);
      else if (use_tma && acc_mode == 4) k_accumulate_tma<F, 4><<<ctx->sm_count * 4, 128, 0, st>>>(b.d_pts, sorted, offsets, p.NB, acc_L, work_counter, partials);
      else if (use_tma) k_accumulate_tma<F><<<ctx->sm_count * 4, 128, 0, st>>>(b.d_pts, sorted, offsets, p.NB, acc_L, work_counter, partials);
      else k_accumulate<F><<<ctx->sm_count * 4, 128, 0, st>>>(b.d_pts, sorted, offsets, p.NB, acc_L, work_counter, partials);
    }
    LAUNCH_CHECK(ctx);
    STAGE_END(ctx);
    }

    // "heavy" = far more slice partials than the average bucket has (skew), not merely many
    const uint32_t acc_entries = (uint32_t)std::min<uint64_t>(T ? ((((uint64_t)nj * p.W) >> T) + p.NB) : (uint64_t)nj * p.W, 0xffffffffull);
    heavy[j] = std::max<uint32_t>(HEAVY_T_MIN, 4u * (acc_entries / acc_L / p.NB + 2u));
    regions.partials[j] = partials;
    regions.offsets[j] = acc_offsets;
    regions.L[j] = acc_L;
    regions.heavy_t[j] = heavy[j];
  }

  if (pipe_pending) {
    CUDA_TRY(ctx, cudaStreamWaitEvent(st, ws.ev_acc_done[0], 0));
    CUDA_TRY(ctx, cudaStreamWaitEvent(st, ws.ev_acc_done[1], 0));
  }
  STAGE_ON(ctx, "bucket_reduce", st);
  if (n) {
    // heavy-bucket collapse: 3 passes cover 2^30 partials per bucket; no-ops when nothing is heavy
    const dim3 hgrid((unsigned)std::min<uint32_t>((p.NB + 31) / 32, (uint32_t)ctx->sm_count * 16), HEAVY_GY);
    for (int j = 0; j < J; j++) {
      if (!mb.count[j] && J > 1) continue;
      uint32_t stride = 1;
      for (int pass = 0; pass < 3; pass++) {
        if ((uint64_t)mb.count[j] * p.W / regions.L[j] + 1 < stride) break;
        k_collapse_heavy<F><<<hgrid, 32, 0, st>>>(regions.partials[j], regions.offsets[j], p.NB, regions.L[j], stride, heavy[j]);
        LAUNCH_CHECK(ctx);
        stride *= HEAVY_GROUP;
      }
    }
    k_bucket_fold<F><<<(p.NB + 127) / 128, 128, 0, st>>>(regions, p.NB);
    LAUNCH_CHECK(ctx);
  }
  const uint32_t total_chunks = (uint32_t)p.Wc * p.K;
  k_bucket_reduce<F><<<(total_chunks + 127) / 128, 128, 0, st>>>(regions.partials[0], regions.offsets[0], regions.L[0], total_chunks, p.K, p.B, p.chunk,
                                                                  (J > 1 && n) ? 1 : 0, chunk_out);
  LAUNCH_CHECK(ctx);
  {
    const uint32_t Y = p.K >= 4096 ? 64 : 1;          // fan-out of the first summation launch
    uint32_t* stage = chunk_out + (size_t)p.Wc * p.K * XY;   // Y points per set, after the chunk results
    if (Y > 1) {
      k_window_sum<F><<<dim3(p.Wc, Y), 256, 0, st>>>(chunk_out, p.K, stage);
      LAUNCH_CHECK(ctx);
      k_window_sum<F><<<dim3(p.Wc, 1), 256, 0, st>>>(stage, Y, window_out);
    } else {
      k_window_sum<F><<<dim3(p.Wc, 1), 256, 0, st>>>(chunk_out, p.K, window_out);
    }
    LAUNCH_CHECK(ctx);
  }
  STAGE_END(ctx);

  STAGE_ON(ctx, "final", st);
  k_final<F><<<1, 32, 0, st>>>(window_out, p.Wc, p.c, d_out);
  LAUNCH_CHECK(ctx);
  STAGE_END(ctx);
  return OZL_OK;
}

template <class F>
int msm_run(ozl_ctx* ctx, MsmWorkspace& ws, cudaStream_t st, const Bases& b, const uint32_t* d_scalars, size_t n,
            uint32_t* d_out) {
  MsmBatches mb;
  mb.J = 1;
  mb.first[0] = 0;
  mb.count[0] = n;
  return msm_run_batched<F>(ctx, ws, st, b, d_scalars, n, mb, d_out);
}

inline const OzlCurveOps* curve_ops_for(int curve) {
  switch (curve) {
    case OZL_BLS12_381_G1: return &ozl_ops_bls12_381_g1;
    case OZL_BLS12_381_G2: return &ozl_ops_bls12_381_g2;
    case OZL_BN254_G1: return &ozl_ops_bn254_g1;
    case OZL_BN254_G2: return &ozl_ops_bn254_g2;
  }
  return nullptr;
}

// MSM over the first n bases of b with device scalars / device output, enqueued on ctx->stream.
inline int ozl_rt_msm(ozl_ctx* ctx, MsmWorkspace& ws, cudaStream_t st, const Bases& b, const uint32_t* d_scalars, size_t n,
                      uint32_t* d_out) {
  const OzlCurveOps* ops = curve_ops_for(b.curve);
  if (!ops || n > b.n) return OZL_ERR_ARG;
  return ops->msm(ctx, ws, st, b, d_scalars, n, d_out);
}

// Point-range batches for an MSM whose scalars start in HOST memory.  The first batch is small (its
// copy is the only part of the transfer nothing can hide), later ones grow no faster than the ratio
// of accumulation time to copy time: ~9:1 for page-locked memory on one GPU (56 GB/s against 2e8
// points/s), ~3.5:1 with eight GPUs pulling through one host, ~1.5:1 for pageable memory.
// OZL_MSM_H2D_SPLIT="f0,f1,..." overrides the fractions (the last batch takes the remainder).
inline void plan_host_batches(size_t n, bool pinned, MsmBatches& mb) {
  static const std::vector<double> forced = []() {
    std::vector<double> v;
    if (const char* e = getenv("OZL_MSM_H2D_SPLIT")) {
      const char* q = e;
      while (*q && (int)v.size() < MSM_MAX_BATCHES - 1) {
        char* end = nullptr;
        double f = strtod(q, &end);
        if (end == q) break;
        if (f > 0) v.push_back(f);
        q = *end ? end + 1 : end;
      }
    }
    return v;
  }();
  std::vector<double> fr;
  if (!forced.empty()) fr = forced;
  else if (n < ((size_t)1 << 22)) fr = {};
  else if (pinned) fr = {1.0 / 32, 7.0 / 32};
  else fr = {1.0 / 32, 1.5 / 32, 2.25 / 32, 3.4 / 32, 5.0 / 32, 7.6 / 32};
  mb.J = 0;
  size_t pos = 0;
  for (double f : fr) {
    if (mb.J >= MSM_MAX_BATCHES - 1) break;
    size_t k = ((size_t)(f * (double)n) + 1023) & ~(size_t)1023;
    if (k == 0 || pos + k >= n) break;
    mb.first[mb.J] = pos;
    mb.count[mb.J] = k;
    mb.J++;
    pos += k;
  }
  mb.first[mb.J] = pos;
  mb.count[mb.J] = n - pos;
  mb.J++;
}

// MSM with HOST scalars: copies batch by batch on the context's copy stream while earlier batches are
// being accumulated, result left in d_out (device).  *d_flags_out (device) = input-error flags.
inline int ozl_rt_msm_host(ozl_ctx* ctx, const Bases& b, const uint64_t* scalars, size_t n, uint32_t* d_out) {
  const OzlCurveOps* ops = curve_ops_for(b.curve);
  if (!ops || n > b.n) return OZL_ERR_ARG;
  int r;
  if ((r = ensure(ctx, ctx->scalars, std::max<size_t>(n, 1) * 32))) return r;
  if (!ctx->copy_stream) CUDA_TRY(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
  for (int k = 0; k < MSM_MAX_BATCHES; k++)
    if (!ctx->ev_batch[k]) CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->ev_batch[k], cudaEventDisableTiming));
  bool pinned = false;
  {
    cudaPointerAttributes at;
    if (n && cudaPointerGetAttributes(&at, scalars) == cudaSuccess) pinned = at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeManaged;
    else cudaGetLastError();
  }
  MsmBatches mb;
  plan_host_batches(n, pinned, mb);
  const bool batch_affine = ctx->batch_levels > 0 || (ctx->batch_levels < 0 && batch_config().levels > 0);
  if (batch_affine && mb.J > 1) {   // the experimental pair levels take one batch
    mb.J = 1; mb.first[0] = 0; mb.count[0] = n;
  }
  mb.h_scalars = scalars;
  mb.copy_stream = ctx->copy_stream;
  mb.ev = ctx->ev_batch;
  // the staging buffer may still be read by work enqueued earlier on the MSM's stream
  if (!ctx->ev_prior) CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->ev_prior, cudaEventDisableTiming));
  CUDA_TRY(ctx, cudaEventRecord(ctx->ev_prior, ctx->stream));
  CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_prior, 0));
  return ops->msm_batched(ctx, ctx->ws, ctx->stream, b, (const uint32_t*)ctx->scalars.p, n, mb, d_out);
}

// Reads the input-error flags the last MSM of this workspace left behind (device word) -- call after the
// MSM's stream has been synchronized together with the result copy.
inline const uint32_t* msm_err_flags(const MsmWorkspace& ws) { return ws.misc.p ? (const uint32_t*)ws.misc.p + 8 : nullptr; }

inline void free_workspace(MsmWorkspace& ws) {
  DevBuf* bufs[] = {&ws.counts, &ws.tile_sums, &ws.sorted, &ws.digits, &ws.chunk_out, &ws.window_out, &ws.misc,
                    &ws.lvl_off, &ws.pair_tab, &ws.pair_a, &ws.pair_b, &ws.pair_pre, &ws.lvl_pts,
                    &ws.offsets[0], &ws.offsets[1], &ws.offsets[2], &ws.offsets[3], &ws.offsets[4], &ws.offsets[5], &ws.offsets[6], &ws.offsets[7],
                    &ws.partials[0], &ws.partials[1], &ws.partials[2], &ws.partials[3], &ws.partials[4], &ws.partials[5], &ws.partials[6], &ws.partials[7]};
  for (DevBuf* b : bufs)
    if (b->p) { cudaFree(b->p); b->p = nullptr; b->cap = 0; }
  for (int i = 0; i < 2; i++) {
    if (ws.acc_stream[i]) { cudaStreamDestroy(ws.acc_stream[i]); ws.acc_stream[i] = nullptr; }
    if (ws.ev_acc_done[i]) { cudaEventDestroy(ws.ev_acc_done[i]); ws.ev_acc_done[i] = nullptr; }
  }
  for (auto& e : ws.ev_interval)
    if (e) { cudaEventDestroy(e); e = nullptr; }
  if (ws.sort_stream) { cudaStreamDestroy(ws.sort_stream); ws.sort_stream = nullptr; }
  if (ws.ev_sorted) { cudaEventDestroy(ws.ev_sorted); ws.ev_sorted = nullptr; }
  ws.last_sort.valid = false;
  if (ws.ev_sort_ready) { cudaEventDestroy(ws.ev_sort_ready); ws.ev_sort_ready = nullptr; }
}

}  // namespace ozl_rt
