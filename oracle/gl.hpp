// ORACLE (test infrastructure, NOT product code): CPU restatement of the Goldilocks field
// and its quadratic extension as used by plonky2_field 0.2.2 (pinned in the reference's
// Cargo.lock:628-632; source not vendored under /root/reference).
//   field/src/goldilocks_field.rs        -> Fp  (p = 2^64 - 2^32 + 1, generator 7,
//                                                POWER_OF_TWO_GENERATOR = 1753635133440165772)
//   field/src/goldilocks_extensions.rs   -> Fp2 (F[X]/(X^2 - 7))
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use this.
#pragma once
#include <cstdint>
#include <cstddef>
#include <vector>
#include <cassert>

namespace orc {

typedef uint64_t u64;
typedef unsigned __int128 u128;

static const u64 GL_P = 0xFFFFFFFF00000001ULL;
static const u64 GL_EPS = 0xFFFFFFFFULL;  // 2^32 - 1 = 2^64 mod p

// all branch-free: the inputs are effectively random, so data-dependent branches mispredict
static inline u64 gl_add(u64 a, u64 b) {
  u64 s = a + b;
  s += ((u64)0 - (u64)(s < a)) & GL_EPS;  // wrapped: +2^64 == +EPS (mod p); cannot wrap twice since a,b < p
  s -= ((u64)0 - (u64)(s >= GL_P)) & GL_P;
  return s;
}
static inline u64 gl_sub(u64 a, u64 b) {
  u64 d = a - b;
  return d - (((u64)0 - (u64)(a < b)) & GL_EPS);  // borrow: -2^64 == -EPS, i.e. +p mod 2^64
}
static inline u64 gl_neg(u64 a) { return gl_sub(0, a); }
static inline u64 gl_reduce128(u128 x) {
  u64 lo = (u64)x, hi = (u64)(x >> 64);
  u64 hi_hi = hi >> 32, hi_lo = hi & GL_EPS;
  u64 t0 = lo - hi_hi;
  t0 -= ((u64)0 - (u64)(lo < hi_hi)) & GL_EPS;  // borrow: -2^64 == -EPS
  u64 t1 = hi_lo * GL_EPS;
  u64 r = t0 + t1;
  r += ((u64)0 - (u64)(r < t0)) & GL_EPS;
  r -= ((u64)0 - (u64)(r >= GL_P)) & GL_P;
  return r;
}
static inline u64 gl_mul(u64 a, u64 b) { return gl_reduce128((u128)a * b); }
static inline u64 gl_pow(u64 a, u64 e) {
  u64 r = 1;
  while (e) {
    if (e & 1) r = gl_mul(r, a);
    a = gl_mul(a, a);
    e >>= 1;
  }
  return r;
}
static inline u64 gl_inv(u64 a) {
  assert(a != 0);
  return gl_pow(a, GL_P - 2);
}
static inline u64 gl_from_i64(int64_t x) { return x >= 0 ? (u64)x % GL_P : GL_P - ((u64)(-x) % GL_P); }

// primitive_root_of_unity(n_log) = POWER_OF_TWO_GENERATOR^(2^(32 - n_log))
static inline u64 gl_root_of_unity(unsigned n_log) {
  assert(n_log <= 32);
  u64 g = 1753635133440165772ULL;
  for (unsigned i = n_log; i < 32; i++) g = gl_mul(g, g);
  return g;
}
static const u64 GL_COSET_SHIFT = 7;  // MULTIPLICATIVE_GROUP_GENERATOR == coset_shift()

// Montgomery's trick, like Field::batch_multiplicative_inverse (all inputs must be non-zero).
static inline std::vector<u64> gl_batch_inv(const std::vector<u64>& x) {
  size_t n = x.size();
  std::vector<u64> pre(n), out(n);
  u64 acc = 1;
  for (size_t i = 0; i < n; i++) {
    pre[i] = acc;
    acc = gl_mul(acc, x[i]);
  }
  u64 inv = gl_inv(acc);
  for (size_t i = n; i-- > 0;) {
    out[i] = gl_mul(inv, pre[i]);
    inv = gl_mul(inv, x[i]);
  }
  return out;
}

// ---- value-type wrappers so constraint code can be written once for F and F^2 -------------
struct Fp {
  u64 v;
  Fp() : v(0) {}
  explicit Fp(u64 x) : v(x) {}
  static Fp from_u64(u64 x) { return Fp(x % GL_P); }
  Fp operator+(Fp o) const { return Fp(gl_add(v, o.v)); }
  Fp operator-(Fp o) const { return Fp(gl_sub(v, o.v)); }
  Fp operator*(Fp o) const { return Fp(gl_mul(v, o.v)); }
  Fp operator-() const { return Fp(gl_neg(v)); }
  Fp& operator+=(Fp o) { v = gl_add(v, o.v); return *this; }
  Fp& operator-=(Fp o) { v = gl_sub(v, o.v); return *this; }
  Fp& operator*=(Fp o) { v = gl_mul(v, o.v); return *this; }
  bool operator==(Fp o) const { return v == o.v; }
  bool operator!=(Fp o) const { return v != o.v; }
  Fp inv() const { return Fp(gl_inv(v)); }
  bool is_zero() const { return v == 0; }
};

struct Fp2 {
  u64 c[2];
  Fp2() : c{0, 0} {}
  Fp2(u64 a, u64 b) : c{a, b} {}
  static Fp2 from_u64(u64 x) { return Fp2(x % GL_P, 0); }
  static Fp2 from_base(Fp x) { return Fp2(x.v, 0); }
  Fp2 operator+(Fp2 o) const { return Fp2(gl_add(c[0], o.c[0]), gl_add(c[1], o.c[1])); }
  Fp2 operator-(Fp2 o) const { return Fp2(gl_sub(c[0], o.c[0]), gl_sub(c[1], o.c[1])); }
  Fp2 operator-() const { return Fp2(gl_neg(c[0]), gl_neg(c[1])); }
  Fp2 operator*(Fp2 o) const {
    // (a0 + a1 X)(b0 + b1 X) mod X^2 - 7
    u64 a0b0 = gl_mul(c[0], o.c[0]), a1b1 = gl_mul(c[1], o.c[1]);
    u64 a0b1 = gl_mul(c[0], o.c[1]), a1b0 = gl_mul(c[1], o.c[0]);
    return Fp2(gl_add(a0b0, gl_mul(7, a1b1)), gl_add(a0b1, a1b0));
  }
  Fp2 scalar_mul(u64 s) const { return Fp2(gl_mul(c[0], s), gl_mul(c[1], s)); }
  Fp2& operator+=(Fp2 o) { *this = *this + o; return *this; }
  Fp2& operator-=(Fp2 o) { *this = *this - o; return *this; }
  Fp2& operator*=(Fp2 o) { *this = *this * o; return *this; }
  bool operator==(Fp2 o) const { return c[0] == o.c[0] && c[1] == o.c[1]; }
  bool operator!=(Fp2 o) const { return !(*this == o); }
  bool is_zero() const { return c[0] == 0 && c[1] == 0; }
  Fp2 inv() const {
    // 1/(a0 + a1 X) = (a0 - a1 X) / (a0^2 - 7 a1^2)
    u64 norm = gl_sub(gl_mul(c[0], c[0]), gl_mul(7, gl_mul(c[1], c[1])));
    u64 ni = gl_inv(norm);
    return Fp2(gl_mul(c[0], ni), gl_mul(gl_neg(c[1]), ni));
  }
  Fp2 pow(u64 e) const {
    Fp2 r(1, 0), b = *this;
    while (e) {
      if (e & 1) r = r * b;
      b = b * b;
      e >>= 1;
    }
    return r;
  }
  Fp2 exp_power_of_2(unsigned k) const {
    Fp2 r = *this;
    for (unsigned i = 0; i < k; i++) r = r * r;
    return r;
  }
};

static inline unsigned log2_strict(size_t n) {
  unsigned k = 0;
  while (((size_t)1 << k) < n) k++;
  assert(((size_t)1 << k) == n);
  return k;
}
static inline size_t reverse_bits(size_t x, unsigned bits) {
  size_t r = 0;
  for (unsigned i = 0; i < bits; i++) r |= ((x >> i) & 1) << (bits - 1 - i);
  return r;
}

}  // namespace orc
