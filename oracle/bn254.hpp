// ORACLE (test infrastructure, NOT product code): BN254 base field Fq / Fq2 arithmetic and
// the 16x16-bit limb conventions of the reference (src/starks/mod.rs:13-61,
// src/starks/modular/utils.rs:6-49, src/starks/utils.rs:12-17). The reference delegates the
// field arithmetic to ark-bn254 0.4.0 / ark-ff 0.4.2 (Cargo.lock:88-130, not vendored); the
// results are canonical integers in [0, p), so any correct modular arithmetic is bit-identical.
#pragma once
#include "gl.hpp"
#include <cstring>

namespace orc {

struct U256 {
  u64 v[4];
  bool operator==(const U256& o) const { return !memcmp(v, o.v, sizeof v); }
};

static const U256 BN_P = {{0x3c208c16d87cfd47ULL, 0x97816a916871ca8dULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL}};

static inline int u256_cmp(const U256& a, const U256& b) {
  for (int i = 3; i >= 0; i--) {
    if (a.v[i] < b.v[i]) return -1;
    if (a.v[i] > b.v[i]) return 1;
  }
  return 0;
}
static inline bool u256_is_zero(const U256& a) { return !(a.v[0] | a.v[1] | a.v[2] | a.v[3]); }
static inline u64 u256_add(U256& r, const U256& a, const U256& b) {
  u128 c = 0;
  for (int i = 0; i < 4; i++) {
    c += (u128)a.v[i] + b.v[i];
    r.v[i] = (u64)c;
    c >>= 64;
  }
  return (u64)c;
}
static inline u64 u256_sub(U256& r, const U256& a, const U256& b) {
  u64 borrow = 0;
  for (int i = 0; i < 4; i++) {
    u128 d = (u128)a.v[i] - b.v[i] - borrow;
    r.v[i] = (u64)d;
    borrow = (u64)(d >> 64) & 1;
  }
  return borrow;
}
static inline void u256_shr1(U256& a) {
  for (int i = 0; i < 3; i++) a.v[i] = (a.v[i] >> 1) | (a.v[i + 1] << 63);
  a.v[3] >>= 1;
}
static inline bool u256_bit(const U256& a, int i) { return (a.v[i >> 6] >> (i & 63)) & 1; }

// ---- Fq in canonical (non-Montgomery) form; mulmod via 512-bit product + Montgomery x2 ----
struct FqCtx {
  u64 n0inv;  // -p^-1 mod 2^64
  U256 r2;    // 2^512 mod p
  FqCtx() {
    u64 inv = 1;
    for (int i = 0; i < 6; i++) inv *= 2 - BN_P.v[0] * inv;
    n0inv = (u64)0 - inv;
    // r2 = 2^512 mod p by 512 modular doublings of 1
    U256 x = {{1, 0, 0, 0}};
    for (int i = 0; i < 512; i++) {
      U256 d;
      u64 c = u256_add(d, x, x);
      if (c || u256_cmp(d, BN_P) >= 0) u256_sub(d, d, BN_P);
      x = d;
    }
    r2 = x;
  }
};
static inline const FqCtx& fq_ctx() {
  static FqCtx c;
  return c;
}
// Montgomery product a*b*2^-256 mod p (CIOS)
static inline U256 fq_montmul(const U256& a, const U256& b) {
  const FqCtx& cx = fq_ctx();
  u64 t[6] = {0};
  for (int i = 0; i < 4; i++) {
    u128 c = 0;
    for (int j = 0; j < 4; j++) {
      c += (u128)a.v[j] * b.v[i] + t[j];
      t[j] = (u64)c;
      c >>= 64;
    }
    c += t[4];
    t[4] = (u64)c;
    t[5] = (u64)(c >> 64);
    u64 m = t[0] * cx.n0inv;
    c = (u128)m * BN_P.v[0] + t[0];
    c >>= 64;
    for (int j = 1; j < 4; j++) {
      c += (u128)m * BN_P.v[j] + t[j];
      t[j - 1] = (u64)c;
      c >>= 64;
    }
    c += t[4];
    t[3] = (u64)c;
    t[4] = t[5] + (u64)(c >> 64);
  }
  U256 r = {{t[0], t[1], t[2], t[3]}};
  if (t[4] || u256_cmp(r, BN_P) >= 0) u256_sub(r, r, BN_P);
  return r;
}
static inline U256 fq_mul(const U256& a, const U256& b) {
  // (a*b*R^-1) * R^2 * R^-1 = a*b
  return fq_montmul(fq_montmul(a, b), fq_ctx().r2);
}
static inline U256 fq_add(const U256& a, const U256& b) {
  U256 r;
  u64 c = u256_add(r, a, b);
  if (c || u256_cmp(r, BN_P) >= 0) u256_sub(r, r, BN_P);
  return r;
}
static inline U256 fq_sub(const U256& a, const U256& b) {
  U256 r;
  if (u256_sub(r, a, b)) u256_add(r, r, BN_P);
  return r;
}
static inline U256 fq_neg(const U256& a) {
  if (u256_is_zero(a)) return a;
  U256 r;
  u256_sub(r, BN_P, a);
  return r;
}
static inline U256 fq_from_u64(u64 x) { return U256{{x, 0, 0, 0}}; }
// binary extended Euclid; a in [1, p)
static inline U256 fq_inv(const U256& a) {
  assert(!u256_is_zero(a));
  U256 u = a, v = BN_P, x1 = {{1, 0, 0, 0}}, x2 = {{0, 0, 0, 0}};
  const U256 one = {{1, 0, 0, 0}};
  while (!(u == one) && !(v == one)) {
    while (!(u.v[0] & 1)) {
      u256_shr1(u);
      if (x1.v[0] & 1) u256_add(x1, x1, BN_P);  // x1 + p < 2^255, no overflow
      u256_shr1(x1);
    }
    while (!(v.v[0] & 1)) {
      u256_shr1(v);
      if (x2.v[0] & 1) u256_add(x2, x2, BN_P);
      u256_shr1(x2);
    }
    if (u256_cmp(u, v) >= 0) {
      u256_sub(u, u, v);
      x1 = fq_sub(x1, x2);
    } else {
      u256_sub(v, v, u);
      x2 = fq_sub(x2, x1);
    }
  }
  return (u == one) ? x1 : x2;
}
static inline U256 fq_pow(U256 b, const U256& e) {
  U256 r = {{1, 0, 0, 0}};
  for (int i = 0; i < 256; i++) {
    if (u256_bit(e, i)) r = fq_mul(r, b);
    b = fq_mul(b, b);
  }
  return r;
}

// ---- Fq2 = Fq[u]/(u^2 + 1) -----------------------------------------------------------------
struct Fq2 {
  U256 c0, c1;
  bool operator==(const Fq2& o) const { return c0 == o.c0 && c1 == o.c1; }
};
static inline Fq2 fq2_add(const Fq2& a, const Fq2& b) { return {fq_add(a.c0, b.c0), fq_add(a.c1, b.c1)}; }
static inline Fq2 fq2_sub(const Fq2& a, const Fq2& b) { return {fq_sub(a.c0, b.c0), fq_sub(a.c1, b.c1)}; }
static inline Fq2 fq2_mul(const Fq2& a, const Fq2& b) {
  return {fq_sub(fq_mul(a.c0, b.c0), fq_mul(a.c1, b.c1)), fq_add(fq_mul(a.c0, b.c1), fq_mul(a.c1, b.c0))};
}
static inline bool fq2_is_zero(const Fq2& a) { return u256_is_zero(a.c0) && u256_is_zero(a.c1); }
static inline Fq2 fq2_inv(const Fq2& a) {
  U256 norm = fq_add(fq_mul(a.c0, a.c0), fq_mul(a.c1, a.c1));
  U256 ni = fq_inv(norm);
  return {fq_mul(a.c0, ni), fq_mul(fq_neg(a.c1), ni)};
}
static inline Fq2 fq2_from_u64(u64 x) { return {fq_from_u64(x), fq_from_u64(0)}; }

// ---- 16-bit limb conventions (starks/mod.rs:13-61) ----------------------------------------
static const int N_LIMBS = 16, LIMB_BITS = 16;
static inline void u256_to_limbs(const U256& a, int64_t out[16]) {
  for (int i = 0; i < 16; i++) out[i] = (int64_t)((a.v[i / 4] >> (16 * (i % 4))) & 0xffff);
}
static inline U256 limbs_to_u256(const u64 in[16]) {
  U256 r = {{0, 0, 0, 0}};
  for (int i = 0; i < 16; i++) r.v[i / 4] |= (in[i] & 0xffff) << (16 * (i % 4));
  return r;
}

}  // namespace orc
