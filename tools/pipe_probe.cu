// Pipe-rate probe for the "FP64 pipe co-issue" question: how many IMAD.WIDE,
// DFMA and IADD3 warp-instructions per clock one SM of this GPU sustains alone and interleaved.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/pipe_probe tools/pipe_probe.cu
// Each kernel runs `iters` rounds of 8 independent dependent-chains per thread, 32 warps per SM.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CHAINS 8

// MODE 0: per round and chain NI x IMAD.WIDE (64-bit accumulate, no carry), ND x DFMA, NA x IADD3.
// MODE 1: NI counts mad.lo.cc + madc.hi.cc PAIRS chained through the carry flag as in fp.cuh (ptxas fuses
//         each pair into one IMAD.WIDE.U32.X with a carry predicate in and out).
// Every instruction consumes its chain's previous result, so nothing is loop-invariant.
template <int NI, int ND, int NA, int MODE>
__global__ void __launch_bounds__(256) k_probe(uint64_t* out, int iters, uint32_t seed) {
  uint64_t acc[CHAINS];
  double d[CHAINS];
  uint32_t s[CHAINS], lo[CHAINS], hi[CHAINS];
  const uint32_t b = seed ^ 0x9e3779b9u;
  const double da = 1.0000001 + 1e-9 * threadIdx.x, db = 1e-30;
#pragma unroll
  for (int c = 0; c < CHAINS; c++) { acc[c] = c + threadIdx.x + seed; d[c] = 1.0 + c; s[c] = c + seed; lo[c] = c ^ seed; hi[c] = threadIdx.x; }
  for (int it = 0; it < iters; it++) {
    if (MODE == 0) {
#pragma unroll
      for (int k = 0; k < NI; k++)
#pragma unroll
        for (int c = 0; c < CHAINS; c++) {
          uint32_t l = (uint32_t)acc[c];
          asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[c]) : "r"(l), "r"(b));
        }
    } else {
      // one carry chain across the 8 "limb pairs", NI times: the row structure of the Montgomery product
#pragma unroll
      for (int k = 0; k < NI; k++) {
        asm volatile("mad.lo.cc.u32 %0, %1, %2, %0;" : "+r"(lo[0]) : "r"(hi[1]), "r"(b));
        asm volatile("madc.hi.cc.u32 %0, %1, %2, %0;" : "+r"(hi[0]) : "r"(hi[1]), "r"(b));
#pragma unroll
        for (int c = 1; c < CHAINS; c++) {
          asm volatile("madc.lo.cc.u32 %0, %1, %2, %0;" : "+r"(lo[c]) : "r"(hi[(c + 1) % CHAINS]), "r"(b));
          asm volatile("madc.hi.cc.u32 %0, %1, %2, %0;" : "+r"(hi[c]) : "r"(hi[(c + 1) % CHAINS]), "r"(b));
        }
      }
    }
#pragma unroll
    for (int k = 0; k < ND; k++)
#pragma unroll
      for (int c = 0; c < CHAINS; c++) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[c]) : "d"(da), "d"(db));
#pragma unroll
    for (int k = 0; k < NA; k++)
#pragma unroll
      for (int c = 0; c < CHAINS; c++) asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(s[c]) : "r"(s[(c + 1) % CHAINS]));
  }
  uint64_t r = 0;
#pragma unroll
  for (int c = 0; c < CHAINS; c++) r += acc[c] + (uint64_t)__double_as_longlong(d[c]) + s[c] + lo[c] + hi[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int NI, int ND, int NA, int MODE = 0>
static void run(const char* name, int sms, double ghz) {
  uint64_t* out;
  cudaMalloc(&out, (size_t)sms * 4 * 256 * 8);
  const int iters = 4096, blocks = sms * 4;
  k_probe<NI, ND, NA, MODE><<<blocks, 256>>>(out, 64, 1);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k_probe<NI, ND, NA, MODE><<<blocks, 256>>>(out, iters, 1);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double warps = (double)blocks * 8, per_round = (double)CHAINS;
  const double clk = ms * 1e-3 * ghz * 1e9;               // SM clocks elapsed
  auto rate = [&](int n) { return n ? warps * iters * per_round * n / sms / clk : 0.0; };   // warp-instr per clk per SM
  printf("{\"probe\": \"%s\", \"ms\": %.3f, \"imad_wide_warp_instr_per_clk_sm\": %.3f, \"dfma_warp_instr_per_clk_sm\": %.3f, "
         "\"iadd_warp_instr_per_clk_sm\": %.3f, \"total_issue_per_clk_sm\": %.3f}\n",
         name, ms, rate(NI), rate(ND), rate(NA), rate(NI) + rate(ND) + rate(NA));
  cudaFree(out);
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  const double ghz = p.clockRate * 1e-6;
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_ghz_nominal\": %.3f}\n", p.name, p.multiProcessorCount, ghz);
  const int sms = p.multiProcessorCount;
  run<4, 0, 0>("imad_wide only", sms, ghz);
  run<4, 0, 0, 1>("imad_wide.x carry chain only", sms, ghz);
  run<2, 2, 0, 1>("imad_wide.x + dfma 1:1", sms, ghz);
  run<1, 2, 0, 1>("imad_wide.x + dfma 1:2", sms, ghz);
  run<2, 0, 2, 1>("imad_wide.x + iadd 1:1", sms, ghz);
  run<1, 2, 2, 1>("imad_wide.x + dfma + iadd 1:2:2", sms, ghz);
  run<0, 4, 0>("dfma only", sms, ghz);
  run<0, 0, 4>("iadd only", sms, ghz);
  run<2, 2, 0>("imad_wide + dfma 1:1", sms, ghz);
  run<1, 2, 0>("imad_wide + dfma 1:2", sms, ghz);
  run<2, 1, 0>("imad_wide + dfma 2:1", sms, ghz);
  run<2, 0, 2>("imad_wide + iadd 1:1", sms, ghz);
  run<0, 2, 2>("dfma + iadd 1:1", sms, ghz);
  run<1, 2, 2>("imad_wide + dfma + iadd 1:2:2", sms, ghz);
  run<1, 2, 4>("imad_wide + dfma + iadd 1:2:4", sms, ghz);
  return 0;
}
