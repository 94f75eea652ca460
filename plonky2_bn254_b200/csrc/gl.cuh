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
PB_HD u64 mul(u64 a, u64 b) {
  u64 lo, hi;
  mul_wide(a, b, lo, hi);
  return reduce128(lo, hi);
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
