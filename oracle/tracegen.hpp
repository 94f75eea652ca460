// ORACLE (test infrastructure, NOT product code): trace generation for the three STARKs of
// the reference, restated row by row in the reference's own (sequential) order:
//   modular witnesses    src/starks/modular/modulus_zero.rs:77-123, is_modulus_zero.rs:36-66,
//                        pol_utils.rs:207-246,339-363, utils.rs:6-49
//   G1 add               src/starks/curves/g1/add.rs:52-122
//   G2 add               src/starks/curves/g2/add.rs:59-130, g2/ext/{mul,add,sub,modulus_zero,
//                        is_modulus_zero}.rs
//   Fq mul               src/starks/fields/mul.rs:22-40
//   round flags          src/starks/common/round_flags.rs:21-44
//   row state machines   g1/scalar_mul_stark.rs:55-213, g2/scalar_mul_stark.rs:55-213,
//                        fields/exp_stark.rs:53-196
// Column maps follow the #[repr(C)] views (g1/scalar_mul_view.rs:32-49, g2/scalar_mul_view.rs:34-49,
// fields/exp_view.rs:33-48); all three share one skeleton parameterised by the register width L.
#pragma once
#include "bn254.hpp"
#include <stdexcept>
#include <string>

namespace orc {

enum Kind { KIND_G1 = 0, KIND_G2 = 1, KIND_FQ = 2 };
enum ErrCode { E_OK = 0, E_SCALAR_RANGE = 1, E_INFINITY = 2, E_NOT_CANONICAL = 3, E_INTERNAL = 4, E_BAD_ARG = 5 };
struct OracleError : std::runtime_error {
  int code;
  OracleError(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

static const int N_BITS = 256, PERIOD = 512, MZ_LEN = 80, IMZ_LEN = 96;

struct Layout {
  int kind, L, aux_len, width;
  int reg0, reg1, a, b, c, aux, bits, rf, ts, flag_op, flag_sq_nl, filter, freq, range_counter;
  int rc_lo, rc_hi;  // range-checked columns [rc_lo, rc_hi)
  int in_words;      // u64 words per input on the wire (s, then coordinates)
};
static inline Layout layout_for(int kind) {
  Layout l;
  l.kind = kind;
  l.L = kind == KIND_G1 ? 32 : kind == KIND_G2 ? 64 : 16;
  l.aux_len = kind == KIND_G1 ? 354 : kind == KIND_G2 ? 708 : 80;
  l.reg0 = 0;
  l.reg1 = l.L;
  l.a = 2 * l.L;
  l.b = 3 * l.L;
  l.c = 4 * l.L;
  l.aux = 5 * l.L;
  l.bits = l.aux + l.aux_len;
  l.rf = l.bits + N_BITS;
  l.ts = l.rf + 5;
  l.flag_op = l.ts + 1;
  l.flag_sq_nl = l.ts + 2;
  l.filter = l.ts + 3;
  l.freq = l.ts + 4;
  l.range_counter = l.ts + 5;
  l.width = l.ts + 6;
  l.rc_lo = 2 * l.L;
  l.rc_hi = l.bits;
  l.in_words = kind == KIND_G1 ? 20 : kind == KIND_G2 ? 36 : 8;
  return l;
}

// ---------------- limb polynomial helpers (pol_utils.rs) on i64 ------------------------------
static inline void pol_mul_wide(const int64_t a[16], const int64_t b[16], int64_t r[31]) {
  for (int i = 0; i < 31; i++) r[i] = 0;
  for (int i = 0; i < 16; i++)
    for (int j = 0; j < 16; j++) r[i + j] += a[i] * b[j];
}

struct BnLimbs {
  int64_t m[16];
  u64 p0inv16;  // p^-1 mod 2^16
  BnLimbs() {
    u256_to_limbs(BN_P, m);
    u64 inv = 1;
    for (int i = 0; i < 5; i++) inv = (inv * (2 - (u64)m[0] * inv)) & 0xffff;
    p0inv16 = inv;
  }
};
static inline const BnLimbs& bn_limbs() {
  static BnLimbs b;
  return b;
}

// modulus_zero.rs:77-123. input: 31 signed coefficients of a limb polynomial whose value at
// x = 2^16 is divisible by p. out[80] = is_quot_positive | quot_abs[17] | aux_lo[31] | aux_hi[31].
static inline void gen_modulus_zero(const int64_t input[31], u64 out[MZ_LEN]) {
  const BnLimbs& bl = bn_limbs();
  const int D = 40;  // 640-bit two's complement in base 2^16 (utils.rs:6-31 columns_to_bigint)
  int64_t dg[D];
  {
    int64_t carry = 0;
    for (int i = 0; i < D; i++) {
      int64_t t = carry + (i < 31 ? input[i] : 0);
      dg[i] = t & 0xffff;
      carry = t >> 16;  // arithmetic shift == floor division
    }
  }
  bool neg = dg[D - 1] >= 0x8000;
  if (neg) {  // |V|
    int64_t carry = 1;
    for (int i = 0; i < D; i++) {
      int64_t t = (0xffff - dg[i]) + carry;
      dg[i] = t & 0xffff;
      carry = t >> 16;
    }
  }
  // exact division |V| / p from the low end: q_i = r_i * p^-1 mod 2^16, r -= q_i p 2^(16 i)
  int64_t q[20];
  for (int i = 0; i < 20; i++) {
    int64_t qi = (int64_t)(((u64)dg[i] * bl.p0inv16) & 0xffff);
    q[i] = qi;
    int64_t carry = 0;
    for (int idx = i; idx < D; idx++) {
      int j = idx - i;
      int64_t t = dg[idx] - (j < 16 ? qi * bl.m[j] : 0) + carry;
      dg[idx] = t & 0xffff;
      carry = t >> 16;
    }
  }
  for (int i = 0; i < D; i++)
    if (dg[i] != 0) throw OracleError(E_INTERNAL, "modulus_zero: input not divisible by p");  // modulus_zero.rs:82
  for (int i = 17; i < 20; i++)
    if (q[i] != 0) throw OracleError(E_INTERNAL, "modulus_zero: quotient exceeds 17 limbs");  // utils.rs:34
  bool nonzero = false;
  for (int i = 0; i < 17; i++) nonzero |= q[i] != 0;
  out[0] = (nonzero && !neg) ? 1 : 0;  // Sign::Plus only (modulus_zero.rs:85-89)
  for (int i = 0; i < 17; i++) out[1 + i] = (u64)q[i];
  // constr = (input | 0) - quot_limbs (*) m    (modulus_zero.rs:93-96)
  int64_t constr[32];
  for (int i = 0; i < 32; i++) constr[i] = i < 31 ? input[i] : 0;
  for (int i = 0; i < 17; i++) {
    int64_t qs = neg ? -q[i] : q[i];
    for (int j = 0; j < 16; j++) constr[i + j] -= qs * bl.m[j];
  }
  // aux = constr / (x - 2^16)   (pol_utils.rs:339-363)
  int64_t aux[32];
  aux[0] = -(constr[0] >> 16);
  for (int d = 1; d < 31; d++) aux[d] = (aux[d - 1] - constr[d]) >> 16;
  aux[31] = 0;
  // the reference asserts aux[31] == 0, which (together with exact shifts) means the division
  // by (x - 2^16) is exact; check the exactness explicitly
  if (aux[30] - constr[31] != 0) throw OracleError(E_INTERNAL, "modulus_zero: aux[31] != 0");
  for (int i = 0; i < 31; i++) {
    int64_t c = aux[i] + ((int64_t)1 << 29);
    if (c < 0 || c > ((int64_t)1 << 30)) throw OracleError(E_INTERNAL, "modulus_zero: aux out of range");  // :103
    out[18 + i] = (u64)(c & 0xffff);
    out[49 + i] = (u64)((c >> 16) & 0xffff);
  }
}

// is_modulus_zero.rs:36-66. input: 16 signed limbs. out[96] = inv[16] | ModulusZeroAux[80].
static inline int gen_is_modulus_zero(const int64_t input[16], u64 out[IMZ_LEN]) {
  // r = value mod p, in [0, p)
  int64_t dg[20];
  int64_t carry = 0;
  for (int i = 0; i < 20; i++) {
    int64_t t = carry + (i < 16 ? input[i] : 0);
    dg[i] = t & 0xffff;
    carry = t >> 16;
  }
  bool neg = dg[19] >= 0x8000;
  if (neg) {
    int64_t c = 1;
    for (int i = 0; i < 20; i++) {
      int64_t t = (0xffff - dg[i]) + c;
      dg[i] = t & 0xffff;
      c = t >> 16;
    }
  }
  for (int i = 16; i < 20; i++)
    if (dg[i]) throw OracleError(E_INTERNAL, "is_modulus_zero: |input| >= 2^256");
  u64 d16[16];
  for (int i = 0; i < 16; i++) d16[i] = (u64)dg[i];
  U256 r = limbs_to_u256(d16);
  while (u256_cmp(r, BN_P) >= 0) u256_sub(r, r, BN_P);
  if (neg && !u256_is_zero(r)) u256_sub(r, BN_P, r);
  int is_zero = u256_is_zero(r);
  U256 inv = is_zero ? r : fq_inv(r);
  int64_t inv_l[16], diff[31];
  u256_to_limbs(inv, inv_l);
  pol_mul_wide(input, inv_l, diff);
  diff[0] += is_zero - 1;
  for (int i = 0; i < 16; i++) out[i] = (u64)inv_l[i];
  gen_modulus_zero(diff, out + 16);
  return is_zero;
}

// ---------------- G1 add (g1/add.rs:52-122) --------------------------------------------------
struct G1Pt {
  U256 x, y;
};
// aux[354] = is_x_eq | inv[16] | mz[80] | is_x_eq_filter | lambda[16] | lambda_aux[80] | x_aux[80] | y_aux[80]
static inline G1Pt gen_g1_add(const G1Pt& a, const G1Pt& b, u64 aux[354]) {
  int64_t ax[16], ay[16], bx[16], by[16];
  u256_to_limbs(a.x, ax);
  u256_to_limbs(a.y, ay);
  u256_to_limbs(b.x, bx);
  u256_to_limbs(b.y, by);
  int64_t dx[16];
  for (int i = 0; i < 16; i++) dx[i] = bx[i] - ax[i];
  int is_x_eq = gen_is_modulus_zero(dx, aux + 1);
  aux[0] = is_x_eq;
  U256 lambda;
  int64_t lam[16], diff[31], t0[31], t1[31];
  if (!is_x_eq) {
    lambda = fq_mul(fq_sub(b.y, a.y), fq_inv(fq_sub(b.x, a.x)));
    u256_to_limbs(lambda, lam);
    pol_mul_wide(lam, dx, diff);
    for (int i = 0; i < 16; i++) diff[i] -= by[i] - ay[i];
  } else {
    if (!(a.y == b.y)) throw OracleError(E_INFINITY, "g1 add: a = -b (point at infinity)");  // add.rs:76-78
    if (u256_is_zero(a.y)) throw OracleError(E_INFINITY, "g1 add: doubling a point with y = 0");
    U256 three_x2 = fq_mul(fq_from_u64(3), fq_mul(a.x, a.x));
    lambda = fq_mul(three_x2, fq_inv(fq_mul(fq_from_u64(2), a.y)));
    u256_to_limbs(lambda, lam);
    pol_mul_wide(ax, ax, t0);
    pol_mul_wide(lam, ay, t1);
    for (int i = 0; i < 31; i++) diff[i] = 2 * t1[i] - 3 * t0[i];
  }
  aux[97] = is_x_eq;  // is_x_eq_filter
  for (int i = 0; i < 16; i++) aux[98 + i] = (u64)lam[i];
  gen_modulus_zero(diff, aux + 114);
  G1Pt c;
  c.x = fq_sub(fq_sub(fq_mul(lambda, lambda), a.x), b.x);
  c.y = fq_sub(fq_mul(lambda, fq_sub(a.x, c.x)), a.y);
  int64_t cx[16], cy[16];
  u256_to_limbs(c.x, cx);
  u256_to_limbs(c.y, cy);
  // x: lambda^2 - (a.x + b.x + c.x)
  pol_mul_wide(lam, lam, diff);
  for (int i = 0; i < 16; i++) diff[i] -= ax[i] + bx[i] + cx[i];
  gen_modulus_zero(diff, aux + 194);
  // y: lambda (c.x - a.x) + c.y + a.y
  int64_t cxax[16];
  for (int i = 0; i < 16; i++) cxax[i] = cx[i] - ax[i];
  pol_mul_wide(lam, cxax, diff);
  for (int i = 0; i < 16; i++) diff[i] += cy[i] + ay[i];
  gen_modulus_zero(diff, aux + 274);
  return c;
}

// ---------------- G2 add (g2/add.rs:59-130) --------------------------------------------------
struct G2Pt {
  Fq2 x, y;
};
struct Ext16 {
  int64_t c0[16], c1[16];
};
struct Ext31 {
  int64_t c0[31], c1[31];
};
static inline Ext16 fq2_to_limbs(const Fq2& a) {
  Ext16 r;
  u256_to_limbs(a.c0, r.c0);
  u256_to_limbs(a.c1, r.c1);
  return r;
}
// g2/ext/mul.rs:14-32
static inline Ext31 mul_ext(const Ext16& x, const Ext16& y) {
  Ext31 r;
  int64_t t0[31], t1[31];
  pol_mul_wide(x.c0, y.c0, t0);
  pol_mul_wide(x.c1, y.c1, t1);
  for (int i = 0; i < 31; i++) r.c0[i] = t0[i] - t1[i];
  pol_mul_wide(x.c0, y.c1, t0);
  pol_mul_wide(x.c1, y.c0, t1);
  for (int i = 0; i < 31; i++) r.c1[i] = t0[i] + t1[i];
  return r;
}
static inline void gen_ext_modulus_zero(const Ext31& in, u64 out[160]) {
  gen_modulus_zero(in.c0, out);
  gen_modulus_zero(in.c1, out + 80);
}
// aux[708] = is_x_eq | is_c0_zero | is_c1_zero | c0_aux[96] | c1_aux[96] | is_x_eq_filter |
//            lambda[32] | lambda_aux[160] | x_aux[160] | y_aux[160]
static inline G2Pt gen_g2_add(const G2Pt& a, const G2Pt& b, u64 aux[708]) {
  Ext16 ax = fq2_to_limbs(a.x), ay = fq2_to_limbs(a.y), bx = fq2_to_limbs(b.x), by = fq2_to_limbs(b.y);
  Ext16 dx;
  for (int i = 0; i < 16; i++) {
    dx.c0[i] = bx.c0[i] - ax.c0[i];
    dx.c1[i] = bx.c1[i] - ax.c1[i];
  }
  int z0 = gen_is_modulus_zero(dx.c0, aux + 3);
  int z1 = gen_is_modulus_zero(dx.c1, aux + 99);
  int is_x_eq = z0 * z1;
  aux[0] = is_x_eq;
  aux[1] = z0;
  aux[2] = z1;
  Fq2 lambda;
  Ext16 lam;
  Ext31 diff;
  if (!is_x_eq) {
    lambda = fq2_mul(fq2_sub(b.y, a.y), fq2_inv(fq2_sub(b.x, a.x)));
    lam = fq2_to_limbs(lambda);
    diff = mul_ext(lam, dx);
    for (int i = 0; i < 16; i++) {
      diff.c0[i] -= by.c0[i] - ay.c0[i];
      diff.c1[i] -= by.c1[i] - ay.c1[i];
    }
  } else {
    if (!(a.y == b.y)) throw OracleError(E_INFINITY, "g2 add: a = -b (point at infinity)");
    if (fq2_is_zero(a.y)) throw OracleError(E_INFINITY, "g2 add: doubling a point with y = 0");
    Fq2 three_x2 = fq2_mul(fq2_from_u64(3), fq2_mul(a.x, a.x));
    lambda = fq2_mul(three_x2, fq2_inv(fq2_mul(fq2_from_u64(2), a.y)));
    lam = fq2_to_limbs(lambda);
    Ext31 xsq = mul_ext(ax, ax), ly = mul_ext(lam, ay);
    for (int i = 0; i < 31; i++) {
      diff.c0[i] = 2 * ly.c0[i] - 3 * xsq.c0[i];
      diff.c1[i] = 2 * ly.c1[i] - 3 * xsq.c1[i];
    }
  }
  aux[195] = is_x_eq;
  for (int i = 0; i < 16; i++) {
    aux[196 + i] = (u64)lam.c0[i];
    aux[212 + i] = (u64)lam.c1[i];
  }
  gen_ext_modulus_zero(diff, aux + 228);
  G2Pt c;
  c.x = fq2_sub(fq2_sub(fq2_mul(lambda, lambda), a.x), b.x);
  c.y = fq2_sub(fq2_mul(lambda, fq2_sub(a.x, c.x)), a.y);
  Ext16 cx = fq2_to_limbs(c.x), cy = fq2_to_limbs(c.y);
  diff = mul_ext(lam, lam);
  for (int i = 0; i < 16; i++) {
    diff.c0[i] -= ax.c0[i] + bx.c0[i] + cx.c0[i];
    diff.c1[i] -= ax.c1[i] + bx.c1[i] + cx.c1[i];
  }
  gen_ext_modulus_zero(diff, aux + 388);
  Ext16 cxax;
  for (int i = 0; i < 16; i++) {
    cxax.c0[i] = cx.c0[i] - ax.c0[i];
    cxax.c1[i] = cx.c1[i] - ax.c1[i];
  }
  diff = mul_ext(lam, cxax);
  for (int i = 0; i < 16; i++) {
    diff.c0[i] += cy.c0[i] + ay.c0[i];
    diff.c1[i] += cy.c1[i] + ay.c1[i];
  }
  gen_ext_modulus_zero(diff, aux + 548);
  return c;
}

// ---------------- Fq mul (fields/mul.rs:22-40) -----------------------------------------------
static inline U256 gen_fq_mul(const U256& a, const U256& b, u64 aux[80]) {
  U256 c = fq_mul(a, b);
  int64_t al[16], bl[16], cl[16], diff[31];
  u256_to_limbs(a, al);
  u256_to_limbs(b, bl);
  u256_to_limbs(c, cl);
  pol_mul_wide(al, bl, diff);
  for (int i = 0; i < 16; i++) diff[i] -= cl[i];
  gen_modulus_zero(diff, aux);
  return c;
}

// ---------------- round flags (common/round_flags.rs:21-44) ----------------------------------
struct RoundFlagTable {
  u64 t[PERIOD][5];
  RoundFlagTable() {
    for (int r = 0; r < PERIOD; r++) {
      u64 counter = (u64)r;
      u64 cprime = gl_sub(counter, (u64)(PERIOD - 1));
      t[r][0] = counter == 0;
      t[r][1] = cprime == 0;
      t[r][2] = counter;
      t[r][3] = counter ? gl_inv(counter) : 0;
      t[r][4] = cprime ? gl_inv(cprime) : 0;
    }
  }
};
static inline const RoundFlagTable& round_flags() {
  static RoundFlagTable t;
  return t;
}

// ---------------- the 512-row state machine, generic over the register type ------------------
static inline void put_u256(u64* dst, const U256& a) {
  int64_t l[16];
  u256_to_limbs(a, l);
  for (int i = 0; i < 16; i++) dst[i] = (u64)l[i];
}
struct RegG1 {
  typedef G1Pt T;
  static void put(u64* d, const T& p) {
    put_u256(d, p.x);
    put_u256(d + 16, p.y);
  }
  static T op(const T& a, const T& b, u64* aux) { return gen_g1_add(a, b, aux); }
};
struct RegG2 {
  typedef G2Pt T;
  static void put(u64* d, const T& p) {
    put_u256(d, p.x.c0);
    put_u256(d + 16, p.x.c1);
    put_u256(d + 32, p.y.c0);
    put_u256(d + 48, p.y.c1);
  }
  static T op(const T& a, const T& b, u64* aux) { return gen_g2_add(a, b, aux); }
};
struct RegFq {
  typedef U256 T;
  static void put(u64* d, const T& p) { put_u256(d, p); }
  static T op(const T& a, const T& b, u64* aux) { return gen_fq_mul(a, b, aux); }
};

// rows: row-major PERIOD x width, zero-initialised. `x` = base point / base, `start` = offset / one.
// Returns the result register of the last row (sum / product).
template <class R>
static typename R::T gen_one_set(const Layout& l, const U256& s, const typename R::T& x, const typename R::T& start,
                                 u64 timestamp, u64* rows) {
  typedef typename R::T T;
  const RoundFlagTable& rf = round_flags();
  bool bits[N_BITS];
  for (int i = 0; i < N_BITS; i++) bits[i] = u256_bit(s, i);
  T reg0 = x, reg1 = start;  // double/square, sum/product
  for (int r = 0; r < PERIOD; r++) {
    u64* row = rows + (size_t)r * l.width;
    bool adding = (r % 2 == 0);
    T a, b, c;
    if (adding) {
      // row 0: a = offset, b = x (first row);  row 2j: a = sum, b = double, bits rotate left
      a = reg1;
      b = reg0;
      if (r > 0) {
        bool b0 = bits[0];
        for (int i = 0; i + 1 < N_BITS; i++) bits[i] = bits[i + 1];
        bits[N_BITS - 1] = b0;
      }
      c = R::op(a, b, row + l.aux);
      if (bits[0]) reg1 = c;
    } else {
      a = reg0;
      b = reg0;
      c = R::op(a, b, row + l.aux);
      reg0 = c;
    }
    R::put(row + l.reg0, reg0);
    R::put(row + l.reg1, reg1);
    R::put(row + l.a, a);
    R::put(row + l.b, b);
    R::put(row + l.c, c);
    for (int i = 0; i < N_BITS; i++) row[l.bits + i] = bits[i];
    for (int i = 0; i < 5; i++) row[l.rf + i] = rf.t[r][i];
    row[l.ts] = timestamp;
    row[l.flag_op] = adding ? 1 : 0;
    row[l.flag_sq_nl] = adding ? 0 : (1 - rf.t[r][1]);
    row[l.filter] = 1;
  }
  return reg1;
}

static inline U256 read_u256(const u64* w, bool need_canonical) {
  U256 r = {{w[0], w[1], w[2], w[3]}};
  if (need_canonical && u256_cmp(r, BN_P) >= 0) throw OracleError(E_NOT_CANONICAL, "coordinate >= p");
  return r;
}

// generate_trace (g1/scalar_mul_stark.rs:55-87): returns column-major width x n_rows.
// `results` (optional) receives the last-row result register limbs per instance (L u64 each).
static inline std::vector<std::vector<u64>> generate_trace(int kind, const u64* inputs, const u64* timestamps,
                                                           size_t n_inputs, size_t min_rows,
                                                           std::vector<u64>* results = nullptr) {
  Layout l = layout_for(kind);
  size_t n = min_rows > n_inputs * PERIOD ? min_rows : n_inputs * PERIOD;
  size_t n_rows = 1;
  while (n_rows < n) n_rows <<= 1;
  std::vector<u64> rows(n_rows * (size_t)l.width, 0);
  if (results) results->assign(n_inputs * (size_t)l.L, 0);
  int err_code = 0;
  std::string err_msg;
#pragma omp parallel for schedule(dynamic, 1)
  for (size_t k = 0; k < n_inputs; k++) {
    try {
      const u64* w = inputs + k * l.in_words;
      U256 s = read_u256(w, false);
      u64* base = rows.data() + k * PERIOD * (size_t)l.width;
      u64* res = results ? results->data() + k * l.L : nullptr;
      if (kind == KIND_G1) {
        G1Pt x = {read_u256(w + 4, true), read_u256(w + 8, true)};
        G1Pt off = {read_u256(w + 12, true), read_u256(w + 16, true)};
        G1Pt out = gen_one_set<RegG1>(l, s, x, off, timestamps[k], base);
        if (res) RegG1::put(res, out);
      } else if (kind == KIND_G2) {
        G2Pt x = {{read_u256(w + 4, true), read_u256(w + 8, true)}, {read_u256(w + 12, true), read_u256(w + 16, true)}};
        G2Pt off = {{read_u256(w + 20, true), read_u256(w + 24, true)},
                    {read_u256(w + 28, true), read_u256(w + 32, true)}};
        G2Pt out = gen_one_set<RegG2>(l, s, x, off, timestamps[k], base);
        if (res) RegG2::put(res, out);
      } else {
        U256 x = read_u256(w + 4, true);
        U256 one = {{1, 0, 0, 0}};
        U256 out = gen_one_set<RegFq>(l, s, x, one, timestamps[k], base);
        if (res) RegFq::put(res, out);
      }
    } catch (OracleError& e) {
#pragma omp critical
      {
        err_code = e.code;
        err_msg = e.what();
      }
    }
  }
  if (err_code) throw OracleError(err_code, err_msg);
  // generate_range_checks (g1/scalar_mul_stark.rs:71-87)
  const size_t range_max = (size_t)1 << LIMB_BITS;
  if (n_rows < range_max) {
    // rows[x][FREQ] with x up to 65535 would be out of bounds in the reference (panic)
    for (size_t r = 0; r < n_rows; r++)
      for (int c = l.rc_lo; c < l.rc_hi; c++)
        if (rows[r * l.width + c] >= n_rows) throw OracleError(E_BAD_ARG, "trace shorter than the range table");
  }
  for (size_t r = 0; r < n_rows; r++) rows[r * l.width + l.range_counter] = r < range_max ? r : range_max - 1;
  std::vector<u64> freq(range_max, 0);
  for (size_t r = 0; r < n_rows; r++)
    for (int c = l.rc_lo; c < l.rc_hi; c++) {
      u64 v = rows[r * l.width + c];
      if (v >= range_max) throw OracleError(E_INTERNAL, "range-checked cell >= 2^16");  // :83
      freq[v]++;
    }
  for (size_t v = 0; v < range_max && v < n_rows; v++) rows[v * l.width + l.freq] = freq[v] % GL_P;
  // trace_rows_to_poly_values: transpose
  std::vector<std::vector<u64>> cols(l.width, std::vector<u64>(n_rows));
#pragma omp parallel for schedule(static)
  for (int c = 0; c < l.width; c++)
    for (size_t r = 0; r < n_rows; r++) cols[c][r] = rows[r * l.width + c];
  return cols;
}

}  // namespace orc
