// Goldilocks field p = 2^64 - 2^32 + 1 and its quadratic extension F[X]/(X^2 - 7) for device
// code. All values that leave a kernel are canonical (< p): the reference stores and hashes
// canonical u64s (plonky2_field GoldilocksField::to_canonical_u64; call sites
// src/starks/common/prover.rs:31-38).
#pragma once
#include "compat.cuh"

namespace gl {

static constexpr u64 P = 0xFFFFFFFF00000001ULL;
static constexpr u64 EPS = 0xFFFFFFFFULL;  // 2^64 mod p
static constexpr u64 COSET_SHIFT = 7;
static constexpr u64 ROOT_2_32 = 1753635133440165772ULL;  // POWER_OF_TWO_GENERATOR, order 2^32

PB_HD u64 add(u64 a, u64 b) {
  u64 s = a + b;
  s += (s < a) ? EPS : 0;  // a,b < p so the wrap happens at most once
  s -= (s >= P) ? P : 0;
  return s;
}
PB_HD u64 sub(u64 a, u64 b) {
  u64 d = a - b;
  return d - ((a < b) ? EPS : 0);
}
PB_HD u64 neg(u64 a) { return a ? P - a : 0; }
PB_HD u64 dbl(u64 a) { return add(a, a); }

PB_HD void mul_wide(u64 a, u64 b, u64& lo, u64& hi) {
#ifdef __CUDA_ARCH__
  lo = a * b;
  hi = __umul64hi(a, b);
#else
  u128 m = (u128)a * b;
  lo = (u64)m;
  hi = (u64)(m >> 64);
#endif
}
// hi * 2^64 + lo  mod p, canonical
PB_HD u64 reduce128(u64 lo, u64 hi) {
  u64 hi_hi = hi >> 32, hi_lo = hi & EPS;
  u64 t0 = lo - hi_hi;
  t0 -= (lo < hi_hi) ? EPS : 0;  // 2^96 == -1
  u64 t1 = (hi_lo << 32) - hi_lo;  // hi_lo * (2^32 - 1)
  u64 r = t0 + t1;
  r += (r < t0) ? EPS : 0;
  r -= (r >= P) ? P : 0;
  return r;
}
// (c2 * 2^128 + hi * 2^64 + lo) mod p with c2 < 2^32:  2^128 == -2^32 (mod p)
PB_HD u64 reduce160(u64 lo, u64 hi, u64 c2) {
  u64 r = reduce128(lo, hi);
  return sub(r, reduce128(c2 << 32, 0));
}
#if !PB_HOSTSIM
// Device forms. B200 integer pipes (tools/microbench/pipe_rates.cu): IMAD.WIDE.U32 runs at full FMA-pipe
// rate, mul.hi / IMAD.HI and carry-out IMADs at half rate, and compare-and-select sequences cost two ALU
// slots per 32-bit word - so products are four mul.wide.u32 and all carries are add.cc / addc chains.
// a * b for ANY u64 a, b; the result is some u64 congruent to a * b mod p (not necessarily < p).
__device__ __forceinline__ u64 mulw32(u32 a, u32 b) {
  u64 r;
  asm("mul.wide.u32 %0, %1, %2;" : "=l"(r) : "r"(a), "r"(b));
  return r;
}
// The 128-bit sum of the four products and the reduction
//   x0 + x1 2^32 + x2 2^64 + x3 2^96 == (x1:x0) - x3 - x2 + x2 2^32   (mod p),  one wrap of +-2^64 == +-(2^32 - 1)
// are written with 128-bit integers: nvcc turns them into 3-input IADD3 carry chains (21 instructions per
// multiplication against 25 for a hand-written add.cc / addc chain).
__device__ __forceinline__ u64 mul_lazy(u64 a, u64 b) {
  const u32 a0 = (u32)a, a1 = (u32)(a >> 32), b0 = (u32)b, b1 = (u32)(b >> 32);
  const u64 P = mulw32(a0, b0), Q = mulw32(a0, b1), R = mulw32(a1, b0), S = mulw32(a1, b1);
  const unsigned __int128 prod =
      (unsigned __int128)P + (((unsigned __int128)Q + R) << 32) + ((unsigned __int128)S << 64);
  const u64 lo = (u64)prod, hi = (u64)(prod >> 64);
  const u32 x2 = (u32)hi, x3 = (u32)(hi >> 32);
  // V in (-2^33, 2^65): r = V mod 2^64 and the wrap count w in {-1, 0, 1}; r + w (2^32 - 1) cannot wrap again
  const __int128 V = (__int128)lo - x3 - x2 + ((__int128)x2 << 32);
  const u64 r = (u64)V;
  const long long w = (long long)(V >> 64);
  return r + (u64)(w * 0xFFFFFFFFLL);
}
// a + b and a - b for ANY u64 a, b; results are arbitrary u64 representatives. A wrap by 2^64 is
// corrected by +-(2^32 - 1); the correction itself can wrap once more, never a third time.
__device__ __forceinline__ u64 add_lazy(u64 a, u64 b) {
  u32 r0, r1;
  asm("{\n\t"
      ".reg .u32 c;\n\t"
      "add.cc.u32   %0, %2, %4;\n\t"
      "addc.cc.u32  %1, %3, %5;\n\t"
      "addc.u32     c, 0, 0;\n\t"
      "sub.u32      c, 0, c;\n\t"
      "add.cc.u32   %0, %0, c;\n\t"
      "addc.cc.u32  %1, %1, 0;\n\t"
      "addc.u32     c, 0, 0;\n\t"
      "sub.u32      c, 0, c;\n\t"
      "add.cc.u32   %0, %0, c;\n\t"
      "addc.u32     %1, %1, 0;\n\t"
      "}"
      : "=&r"(r0), "=&r"(r1)
      : "r"((u32)a), "r"((u32)(a >> 32)), "r"((u32)b), "r"((u32)(b >> 32)));
  return ((u64)r1 << 32) | r0;
}
__device__ __forceinline__ u64 sub_lazy(u64 a, u64 b) {
  u32 r0, r1;
  asm("{\n\t"
      ".reg .u32 m;\n\t"
      "sub.cc.u32   %0, %2, %4;\n\t"
      "subc.cc.u32  %1, %3, %5;\n\t"
      "subc.u32     m, 0, 0;\n\t"    // borrow ? 0xffffffff : 0
      "sub.cc.u32   %0, %0, m;\n\t"
      "subc.cc.u32  %1, %1, 0;\n\t"
      "subc.u32     m, 0, 0;\n\t"
      "sub.cc.u32   %0, %0, m;\n\t"
      "subc.u32     %1, %1, 0;\n\t"
      "}"
      : "=&r"(r0), "=&r"(r1)
      : "r"((u32)a), "r"((u32)(a >> 32)), "r"((u32)b), "r"((u32)(b >> 32)));
  return ((u64)r1 << 32) | r0;
}
__device__ __forceinline__ u64 canonical(u64 a) { return a >= P ? a - P : a; }
#endif
PB_HD u64 mul(u64 a, u64 b) {
#if defined(__CUDA_ARCH__)
  const u64 r = mul_lazy(a, b);
  return r >= P ? r - P : r;
#else
  u64 lo, hi;
  mul_wide(a, b, lo, hi);
  return reduce128(lo, hi);
#endif
}
PB_HD u64 sqr(u64 a) { return mul(a, a); }
PB_HD u64 pow(u64 a, u64 e) {
  u64 r = 1;
  while (e) {
    if (e & 1) r = mul(r, a);
    a = mul(a, a);
    e >>= 1;
  }
  return r;
}
// a^(p-2); p - 2 = 0xFFFFFFFEFFFFFFFF. Addition chain: 2^32 - 1 block, then 32 more squarings.
PB_HD u64 inv(u64 a) {
  // t = a^(2^31 - 1)
  u64 t = a;
  for (int i = 0; i < 30; i++) t = mul(sqr(t), a);
  // a^(2^32 - 2) = t^2 ; we need exponent (2^32 - 2) * 2^32 + (2^32 - 1)
  u64 hi = sqr(t);           // a^(2^32 - 2)
  u64 lo = mul(hi, a);       // a^(2^32 - 1)
  u64 r = hi;
  for (int i = 0; i < 32; i++) r = sqr(r);
  return mul(r, lo);
}

// lazy accumulator for sums of products: 64x64 -> 128-bit terms accumulated in 160 bits
#if !PB_HOSTSIM
struct Acc {
  u32 w0, w1, w2, w3, w4;
  PB_HD Acc() : w0(0), w1(0), w2(0), w3(0), w4(0) {}
  // acc += a * b (any u64 a, b): 4 mul.wide.u32 + 13 carry-chain additions
  PB_HD void mac(u64 a, u64 b) {
#if defined(__CUDA_ARCH__)
    const u32 a0 = (u32)a, a1 = (u32)(a >> 32), b0 = (u32)b, b1 = (u32)(b >> 32);
    asm("{\n\t"
        ".reg .u64 P, Q, R, S; .reg .u32 p0, p1, q0, q1, r0, r1, s0, s1;\n\t"
        "mul.wide.u32 P, %5, %7;\n\t"
        "mul.wide.u32 Q, %5, %8;\n\t"
        "mul.wide.u32 R, %6, %7;\n\t"
        "mul.wide.u32 S, %6, %8;\n\t"
        "mov.b64 {p0, p1}, P; mov.b64 {q0, q1}, Q; mov.b64 {r0, r1}, R; mov.b64 {s0, s1}, S;\n\t"
        "add.cc.u32   %0, %0, p0;\n\t"
        "addc.cc.u32  %1, %1, p1;\n\t"
        "addc.cc.u32  %2, %2, s0;\n\t"
        "addc.cc.u32  %3, %3, s1;\n\t"
        "addc.u32     %4, %4, 0;\n\t"
        "add.cc.u32   %1, %1, q0;\n\t"
        "addc.cc.u32  %2, %2, q1;\n\t"
        "addc.cc.u32  %3, %3, 0;\n\t"
        "addc.u32     %4, %4, 0;\n\t"
        "add.cc.u32   %1, %1, r0;\n\t"
        "addc.cc.u32  %2, %2, r1;\n\t"
        "addc.cc.u32  %3, %3, 0;\n\t"
        "addc.u32     %4, %4, 0;\n\t"
        "}"
        : "+r"(w0), "+r"(w1), "+r"(w2), "+r"(w3), "+r"(w4)
        : "r"(a0), "r"(a1), "r"(b0), "r"(b1));
#else
    u64 l, h;
    mul_wide(a, b, l, h);
    add3(l, h);
#endif
  }
  PB_HD void add3(u64 l, u64 h) {  // host pass only (never on a hot path)
    u64 lo = ((u64)w1 << 32) | w0, hi = ((u64)w3 << 32) | w2;
    lo += l;
    u64 cy = lo < l;
    hi += cy;
    w4 += (hi < cy);
    hi += h;
    w4 += (hi < h);
    w0 = (u32)lo;
    w1 = (u32)(lo >> 32);
    w2 = (u32)hi;
    w3 = (u32)(hi >> 32);
  }
  PB_HD void addv(u64 v) {
#if defined(__CUDA_ARCH__)
    asm("add.cc.u32 %0, %0, %5;\n\taddc.cc.u32 %1, %1, %6;\n\taddc.cc.u32 %2, %2, 0;\n\t"
        "addc.cc.u32 %3, %3, 0;\n\taddc.u32 %4, %4, 0;"
        : "+r"(w0), "+r"(w1), "+r"(w2), "+r"(w3), "+r"(w4)
        : "r"((u32)v), "r"((u32)(v >> 32)));
#else
    add3(v, 0);
#endif
  }
  PB_HD u64 reduce() const {
    return reduce160(((u64)w1 << 32) | w0, ((u64)w3 << 32) | w2, w4);
  }
};
#else
struct Acc {
  u64 lo, hi;
  u32 c;
  PB_HD Acc() : lo(0), hi(0), c(0) {}
  PB_HD void mac(u64 a, u64 b) {
    u64 l, h;
    mul_wide(a, b, l, h);
    lo += l;
    u64 cy = lo < l;
    hi += cy;
    c += (hi < cy);
    hi += h;
    c += (hi < h);
  }
  PB_HD void addv(u64 v) {
    lo += v;
    u64 cy = lo < v;
    hi += cy;
    c += (hi < cy);
  }
  PB_HD u64 reduce() const { return reduce160(lo, hi, c); }
};
#endif

// ---- quadratic extension ---------------------------------------------------------------
struct E2 {
  u64 a, b;  // a + b X
};
PB_HD E2 e2(u64 a, u64 b) {
  E2 r;
  r.a = a;
  r.b = b;
  return r;
}
PB_HD E2 eadd(E2 x, E2 y) { return e2(add(x.a, y.a), add(x.b, y.b)); }
PB_HD E2 esub(E2 x, E2 y) { return e2(sub(x.a, y.a), sub(x.b, y.b)); }
PB_HD E2 emul(E2 x, E2 y) {
  u64 aa = mul(x.a, y.a), bb = mul(x.b, y.b);
  u64 ab = mul(x.a, y.b), ba = mul(x.b, y.a);
  return e2(add(aa, mul(7, bb)), add(ab, ba));
}
PB_HD E2 emul_base(E2 x, u64 s) { return e2(mul(x.a, s), mul(x.b, s)); }
PB_HD E2 einv(E2 x) {
  u64 norm = sub(sqr(x.a), mul(7, sqr(x.b)));
  u64 ni = inv(norm);
  return e2(mul(x.a, ni), mul(neg(x.b), ni));
}
PB_HD E2 epow(E2 x, u64 e) {
  E2 r = e2(1, 0);
  while (e) {
    if (e & 1) r = emul(r, x);
    x = emul(x, x);
    e >>= 1;
  }
  return r;
}

// host-side helpers (transcript-side scalar work: a handful of field ops per proof)
static inline u64 root_of_unity(unsigned log_n) {
  u64 g = ROOT_2_32;
  for (unsigned i = log_n; i < 32; i++) g = mul(g, g);
  return g;
}

PB_HD u32 brev32(u32 x, unsigned bits) {
#ifdef __CUDA_ARCH__
  return bits ? (__brev(x) >> (32 - bits)) : 0;
#else
  u32 r = 0;
  for (unsigned i = 0; i < bits; i++) r |= ((x >> i) & 1u) << (bits - 1 - i);
  return r;
#endif
}

}  // namespace gl
