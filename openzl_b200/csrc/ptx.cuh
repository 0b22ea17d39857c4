// Carry-chain primitives for multi-limb integer arithmetic on sm_100a.
//
// Each wrapper is exactly one PTX instruction using the per-thread carry flag CC.CF
// (add.cc / addc / sub.cc / subc / mad.{lo,hi}.cc / madc.{lo,hi}.cc).  ptxas turns the
// carry flag into predicate registers and pairs mad.lo.cc + madc.hi.cc into
// IMAD.WIDE.U32(.X) on the fma pipe, which is what bounds every kernel in this library.
//
// When compiled WITHOUT nvcc and with -DOZL_HOST_EMU the same wrappers are emulated with a
// thread-local carry bit.  That build exists only so tests/ can exercise the limb-level
// algorithms of fp.cuh / ec.cuh with g++ on a machine that has no GPU; the product library
// never contains it.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define OZL_DEV __device__ __forceinline__
#define OZL_HOSTDEV __host__ __device__ __forceinline__
#define OZL_DEV_NOINLINE __device__ __noinline__
#else
#ifndef OZL_HOST_EMU
#error "ptx.cuh needs nvcc (or -DOZL_HOST_EMU for the g++ test build)"
#endif
#define OZL_DEV inline
#define OZL_HOSTDEV inline
#define OZL_DEV_NOINLINE inline
#endif

namespace ozl {
namespace ptx {

#if defined(__CUDACC__)

OZL_DEV uint32_t add_cc(uint32_t a, uint32_t b) {
  uint32_t r; asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
}
OZL_DEV uint32_t addc_cc(uint32_t a, uint32_t b) {
  uint32_t r; asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
}
OZL_DEV uint32_t addc(uint32_t a, uint32_t b) {
  uint32_t r; asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
}
OZL_DEV uint32_t sub_cc(uint32_t a, uint32_t b) {
  uint32_t r; asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
}
OZL_DEV uint32_t subc_cc(uint32_t a, uint32_t b) {
  uint32_t r; asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
}
OZL_DEV uint32_t subc(uint32_t a, uint32_t b) {
  uint32_t r; asm volatile("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
}
OZL_DEV uint32_t mul_lo(uint32_t a, uint32_t b) {
  uint32_t r; asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
}
OZL_DEV uint32_t mul_hi(uint32_t a, uint32_t b) {
  uint32_t r; asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
}
OZL_DEV uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r; asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r;
}
OZL_DEV uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r; asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r;
}
OZL_DEV uint32_t mad_hi_cc(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r; asm volatile("mad.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r;
}
OZL_DEV uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r; asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r;
}
OZL_DEV uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r; asm volatile("madc.hi.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r;
}

// ---- TMA (cp.async.bulk) + mbarrier: stage contiguous global slices into shared memory ------
OZL_DEV uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
OZL_DEV void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
OZL_DEV void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
OZL_DEV void mbar_arrive(uint64_t* bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_addr(bar)) : "memory");
}
OZL_DEV void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_addr(bar)), "r"(bytes)
               : "memory");
}
OZL_DEV void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_addr(bar)), "r"(parity)
      : "memory");
}
// 1-D bulk copy global -> shared (SASS: UBLKCP); bytes must be a multiple of 16, both addresses 16 B aligned
OZL_DEV void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_addr(bar))
               : "memory");
}
OZL_DEV void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }

// ---- FP64 primitives of the double-precision Montgomery product (fp64mul.cuh) -----------------
// Every value is an exact integer held in a double.  *_rz rounds toward zero (used ON PURPOSE to cut a sum
// at a chosen bit position); the *_x forms are exact by construction of their operands (the g++ build checks it).
OZL_DEV double fma_rz(double a, double b, double c) { return __fma_rz(a, b, c); }
OZL_DEV double add_rz(double a, double b) { return __dadd_rz(a, b); }
OZL_DEV double fma_x(double a, double b, double c) { return __fma_rn(a, b, c); }
OZL_DEV double add_x(double a, double b) { return __dadd_rn(a, b); }
OZL_DEV double mul_x(double a, double b) { return __dmul_rn(a, b); }
// 48-bit unsigned integer (hi16 : lo32) <-> double, through the 2^52 bias (no I2F / F2I conversions)
OZL_DEV double u48_to_double(uint32_t lo32, uint32_t hi16) { return __dadd_rn(__hiloint2double((int)(0x43300000u | hi16), (int)lo32), -4503599627370496.0); }
OZL_DEV void double_to_u48(double d, uint32_t& lo32, uint32_t& hi16) {
  const double b = __dadd_rn(d, 4503599627370496.0);
  lo32 = (uint32_t)__double2loint(b);
  hi16 = (uint32_t)__double2hiint(b) & 0xffffu;
}

#else  // ---- g++ emulation (tests only) ----------------------------------------------------

static thread_local uint32_t g_cf = 0;

inline uint32_t add_cc(uint32_t a, uint32_t b) {
  uint64_t t = (uint64_t)a + b; g_cf = (uint32_t)(t >> 32); return (uint32_t)t;
}
inline uint32_t addc_cc(uint32_t a, uint32_t b) {
  uint64_t t = (uint64_t)a + b + g_cf; g_cf = (uint32_t)(t >> 32); return (uint32_t)t;
}
inline uint32_t addc(uint32_t a, uint32_t b) { return a + b + g_cf; }
inline uint32_t sub_cc(uint32_t a, uint32_t b) {
  uint64_t t = (uint64_t)a - b; g_cf = (uint32_t)((t >> 32) & 1); return (uint32_t)t;
}
inline uint32_t subc_cc(uint32_t a, uint32_t b) {
  uint64_t t = (uint64_t)a - b - g_cf; g_cf = (uint32_t)((t >> 32) & 1); return (uint32_t)t;
}
inline uint32_t subc(uint32_t a, uint32_t b) { return a - b - g_cf; }
inline uint32_t mul_lo(uint32_t a, uint32_t b) { return a * b; }
inline uint32_t mul_hi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
inline uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { return add_cc(a * b, c); }
inline uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { return addc_cc(a * b, c); }
inline uint32_t mad_hi_cc(uint32_t a, uint32_t b, uint32_t c) { return add_cc(mul_hi(a, b), c); }
inline uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { return addc_cc(mul_hi(a, b), c); }
inline uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { return addc(mul_hi(a, b), c); }

// FP64 primitives, emulated EXACTLY with 128-bit integers (all operands are integer-valued doubles below
// 2^102); the *_x forms abort when their result would not be exactly representable, which is the property the
// device code relies on.
} }  // close namespaces for the includes
#include <cstdio>
#include <cstdlib>
namespace ozl { namespace ptx {
typedef __int128 emu_i128;
inline emu_i128 emu_to_int(double d) {
  if (d != (double)(emu_i128)d) { fprintf(stderr, "fp64 emu: non-integer operand %a\n", d); abort(); }
  return (emu_i128)d;
}
inline double emu_trunc53(emu_i128 v) {     // round toward zero to 53 significant bits
  const bool neg = v < 0;
  unsigned __int128 m = neg ? (unsigned __int128)(-v) : (unsigned __int128)v;
  int bits = 0;
  for (unsigned __int128 t = m; t; t >>= 1) bits++;
  if (bits > 53) m = (m >> (bits - 53)) << (bits - 53);
  const double r = (double)m;
  return neg ? -r : r;
}
inline double emu_exact(emu_i128 v, const char* what) {
  if (emu_trunc53(v) != (double)v || (emu_i128)(double)v != v) { fprintf(stderr, "fp64 emu: inexact %s\n", what); abort(); }
  return (double)v;
}
inline double fma_rz(double a, double b, double c) { return emu_trunc53(emu_to_int(a) * emu_to_int(b) + emu_to_int(c)); }
inline double add_rz(double a, double b) { return emu_trunc53(emu_to_int(a) + emu_to_int(b)); }
inline double fma_x(double a, double b, double c) { return emu_exact(emu_to_int(a) * emu_to_int(b) + emu_to_int(c), "fma"); }
inline double add_x(double a, double b) { return emu_exact(emu_to_int(a) + emu_to_int(b), "add"); }
inline double mul_x(double a, double b) {          // b is a power of two (possibly negative exponent)
  const double r = a * b;
  if (r != (double)(emu_i128)r) { fprintf(stderr, "fp64 emu: scaling left a fraction\n"); abort(); }
  return r;
}
inline double u48_to_double(uint32_t lo32, uint32_t hi16) { return (double)(((uint64_t)hi16 << 32) | lo32); }
inline void double_to_u48(double d, uint32_t& lo32, uint32_t& hi16) {
  const uint64_t v = (uint64_t)emu_to_int(d);
  if (v >> 48) { fprintf(stderr, "fp64 emu: limb above 48 bits\n"); abort(); }
  lo32 = (uint32_t)v;
  hi16 = (uint32_t)(v >> 32);
}

#endif

}  // namespace ptx
}  // namespace ozl
