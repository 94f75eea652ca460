// ORACLE (test infrastructure, NOT product code): constraint evaluation, written once for the
// base field (prover, quotient loop) and the extension field (verifier at zeta), emitted in
// EXACTLY the reference's order:
//   ConstraintConsumer        starky 0.4.0 constraint_consumer.rs (un-vendored)
//   eval_modulus_zero         src/starks/modular/modulus_zero.rs:163-198
//   eval_is_modulus_zero      src/starks/modular/is_modulus_zero.rs:69-84
//   eval_g1_add               src/starks/curves/g1/add.rs:125-185
//   eval_g2_add               src/starks/curves/g2/add.rs:133-196 (+ g2/ext/*.rs)
//   eval_fq_mul               src/starks/fields/mul.rs:43-57
//   eval_round_flags          src/starks/common/round_flags.rs:46-81
//   eval_packed_generic       g1/scalar_mul_stark.rs:226-339, g2/scalar_mul_stark.rs:226-338,
//                             fields/exp_stark.rs:208-327
#pragma once
#include "tracegen.hpp"

namespace orc {

template <class T>
struct Consumer {
  std::vector<T> alphas, accs;
  T z_last, l_first, l_last;
  size_t count = 0;
  Consumer(const std::vector<T>& al, T zl, T lf, T ll) : alphas(al), accs(al.size()), z_last(zl), l_first(lf), l_last(ll) {}
  void constraint(T c) {
    for (size_t i = 0; i < alphas.size(); i++) accs[i] = accs[i] * alphas[i] + c;
    count++;
  }
  void constraint_transition(T c) { constraint(c * z_last); }
  void constraint_first_row(T c) { constraint(c * l_first); }
  void constraint_last_row(T c) { constraint(c * l_last); }
};

template <class T>
static inline T K(u64 x) { return T::from_u64(x); }

template <class T>
static void pol_mul_wide_t(const T* a, const T* b, T* r) {  // 16 x 16 -> 31
  for (int i = 0; i < 31; i++) r[i] = T();
  for (int i = 0; i < 16; i++)
    for (int j = 0; j < 16; j++) r[i + j] = r[i + j] + a[i] * b[j];
}

// aux: 80 cells (is_quot_positive | quot_abs[17] | lo[31] | hi[31])
template <class T>
static void eval_modulus_zero(Consumer<T>& y, T filter, const T* input /*31*/, const T* aux) {
  const BnLimbs& bl = bn_limbs();
  T s = aux[0];
  y.constraint(filter * (s * s - s));
  T quot_sign = K<T>(2) * s - K<T>(1);
  T quot[17];
  for (int i = 0; i < 17; i++) quot[i] = quot_sign * aux[1 + i];
  T constr[32];
  for (int i = 0; i < 32; i++) constr[i] = T();
  for (int i = 0; i < 17; i++)
    for (int j = 0; j < 16; j++) constr[i + j] = constr[i + j] + quot[i] * K<T>((u64)bl.m[j]);
  T base = K<T>((u64)1 << 16), offset = K<T>((u64)1 << 29);
  T ap[32];
  for (int i = 0; i < 31; i++) ap[i] = (aux[18 + i] - offset) + base * aux[49 + i];
  ap[31] = T();
  // pol_adjoin_root(ap, base): (x - base) * ap(x)
  constr[0] = constr[0] + (-(base * ap[0]));
  for (int d = 1; d < 32; d++) constr[d] = constr[d] + (ap[d - 1] - base * ap[d]);
  for (int i = 0; i < 31; i++) constr[i] = constr[i] - input[i];
  for (int i = 0; i < 32; i++) y.constraint(filter * constr[i]);
}

// aux: 96 cells (inv[16] | ModulusZeroAux)
template <class T>
static void eval_is_modulus_zero(Consumer<T>& y, T filter, const T* input /*16*/, T is_zero, const T* aux) {
  T diff[31];
  pol_mul_wide_t(input, aux, diff);
  diff[0] = diff[0] + (is_zero - K<T>(1));
  eval_modulus_zero(y, filter, diff, aux + 16);
  for (int i = 0; i < 16; i++) y.constraint(filter * (input[i] * is_zero));
}

template <class T>
static void eval_g1_add(Consumer<T>& y, T filter, const T* a, const T* b, const T* c, const T* aux) {
  const T *ax = a, *ay = a + 16, *bx = b, *by = b + 16, *cx = c, *cy = c + 16;
  T dx[16];
  for (int i = 0; i < 16; i++) dx[i] = bx[i] - ax[i];
  T is_x_eq = aux[0];
  eval_is_modulus_zero(y, filter, dx, is_x_eq, aux + 1);
  T is_x_eq_filter = aux[97];
  y.constraint(filter * is_x_eq - is_x_eq_filter);
  T is_not_eq_filter = filter - is_x_eq_filter;
  const T* lam = aux + 98;
  T diff[31], t0[31], t1[31];
  // a.x != b.x
  pol_mul_wide_t(lam, dx, diff);
  for (int i = 0; i < 16; i++) diff[i] = diff[i] - (by[i] - ay[i]);
  eval_modulus_zero(y, is_not_eq_filter, diff, aux + 114);
  // a.x == b.x
  pol_mul_wide_t(ax, ax, t0);
  pol_mul_wide_t(lam, ay, t1);
  for (int i = 0; i < 31; i++) diff[i] = t1[i] * K<T>(2) - t0[i] * K<T>(3);
  eval_modulus_zero(y, is_x_eq_filter, diff, aux + 114);
  for (int i = 0; i < 16; i++) y.constraint(is_x_eq_filter * (ay[i] - by[i]));
  // x
  pol_mul_wide_t(lam, lam, diff);
  for (int i = 0; i < 16; i++) diff[i] = diff[i] - ((ax[i] + bx[i]) + cx[i]);
  eval_modulus_zero(y, filter, diff, aux + 194);
  // y
  T cxax[16];
  for (int i = 0; i < 16; i++) cxax[i] = cx[i] - ax[i];
  pol_mul_wide_t(lam, cxax, diff);
  for (int i = 0; i < 16; i++) diff[i] = diff[i] + (cy[i] + ay[i]);
  eval_modulus_zero(y, filter, diff, aux + 274);
}

template <class T>
struct ExtMul {
  T c0[31], c1[31];
};
template <class T>
static ExtMul<T> mul_ext_t(const T* x /*c0|c1*/, const T* yv) {
  ExtMul<T> r;
  T t0[31], t1[31];
  pol_mul_wide_t(x, yv, t0);
  pol_mul_wide_t(x + 16, yv + 16, t1);
  for (int i = 0; i < 31; i++) r.c0[i] = t0[i] - t1[i];
  pol_mul_wide_t(x, yv + 16, t0);
  pol_mul_wide_t(x + 16, yv, t1);
  for (int i = 0; i < 31; i++) r.c1[i] = t0[i] + t1[i];
  return r;
}
template <class T>
static void eval_ext_modulus_zero(Consumer<T>& y, T filter, const ExtMul<T>& in, const T* aux /*160*/) {
  eval_modulus_zero(y, filter, in.c0, aux);
  eval_modulus_zero(y, filter, in.c1, aux + 80);
}

template <class T>
static void eval_g2_add(Consumer<T>& y, T filter, const T* a, const T* b, const T* c, const T* aux) {
  const T *ax = a, *ay = a + 32, *bx = b, *by = b + 32, *cx = c, *cy = c + 32;  // each c0[16] | c1[16]
  T dx[32];
  for (int i = 0; i < 32; i++) dx[i] = bx[i] - ax[i];
  T is_x_eq = aux[0], is_c0_zero = aux[1], is_c1_zero = aux[2];
  // eval_is_ext_modulus_zero (g2/ext/is_modulus_zero.rs:48-73)
  y.constraint(filter * (is_c0_zero * is_c1_zero - is_x_eq));
  eval_is_modulus_zero(y, filter, dx, is_c0_zero, aux + 3);
  eval_is_modulus_zero(y, filter, dx + 16, is_c1_zero, aux + 99);
  T is_x_eq_filter = aux[195];
  y.constraint(filter * is_x_eq - is_x_eq_filter);
  T is_not_eq_filter = filter - is_x_eq_filter;
  const T* lam = aux + 196;
  ExtMul<T> diff = mul_ext_t(lam, dx);
  for (int i = 0; i < 16; i++) {
    diff.c0[i] = diff.c0[i] - (by[i] - ay[i]);
    diff.c1[i] = diff.c1[i] - (by[16 + i] - ay[16 + i]);
  }
  eval_ext_modulus_zero(y, is_not_eq_filter, diff, aux + 228);
  ExtMul<T> xsq = mul_ext_t(ax, ax), ly = mul_ext_t(lam, ay);
  for (int i = 0; i < 31; i++) {
    diff.c0[i] = K<T>(2) * ly.c0[i] - K<T>(3) * xsq.c0[i];
    diff.c1[i] = K<T>(2) * ly.c1[i] - K<T>(3) * xsq.c1[i];
  }
  eval_ext_modulus_zero(y, is_x_eq_filter, diff, aux + 228);
  for (int i = 0; i < 32; i++) y.constraint(is_x_eq_filter * (ay[i] - by[i]));
  diff = mul_ext_t(lam, lam);
  for (int i = 0; i < 16; i++) {
    diff.c0[i] = diff.c0[i] - ((ax[i] + bx[i]) + cx[i]);
    diff.c1[i] = diff.c1[i] - ((ax[16 + i] + bx[16 + i]) + cx[16 + i]);
  }
  eval_ext_modulus_zero(y, filter, diff, aux + 388);
  T cxax[32];
  for (int i = 0; i < 32; i++) cxax[i] = cx[i] - ax[i];
  diff = mul_ext_t(lam, cxax);
  for (int i = 0; i < 16; i++) {
    diff.c0[i] = diff.c0[i] + (cy[i] + ay[i]);
    diff.c1[i] = diff.c1[i] + (cy[16 + i] + ay[16 + i]);
  }
  eval_ext_modulus_zero(y, filter, diff, aux + 548);
}

template <class T>
static void eval_fq_mul(Consumer<T>& y, T filter, const T* a, const T* b, const T* c, const T* aux) {
  T diff[31];
  pol_mul_wide_t(a, b, diff);
  for (int i = 0; i < 16; i++) diff[i] = diff[i] - c[i];
  eval_modulus_zero(y, filter, diff, aux);
}

template <class T>
static void eval_round_flags(Consumer<T>& y, T filter, const T* rf, T next_counter) {
  T is_first = rf[0], is_last = rf[1], counter = rf[2], inv_counter = rf[3], inv_counter_prime = rf[4];
  T one = K<T>(1);
  T not_filter = one - filter;
  y.constraint(not_filter * is_first);
  y.constraint(not_filter * is_last);
  y.constraint(filter * (counter * inv_counter - (one - is_first)));
  y.constraint(filter * counter * is_first);
  T counter_prime = counter - K<T>(PERIOD - 1);
  y.constraint(filter * (counter_prime * inv_counter_prime - (one - is_last)));
  y.constraint(filter * counter_prime * is_last);
  y.constraint(filter * (one - is_last) * (next_counter - counter - one));
  y.constraint(filter * is_last * next_counter);
}

template <class T>
static inline void eval_eq_n(Consumer<T>& y, T filter, const T* a, const T* b, int n) {
  for (int i = 0; i < n; i++) y.constraint(filter * (a[i] - b[i]));
}

// eval_packed_generic for all three STARKs (they share the skeleton; see file header)
template <class T>
static void eval_stark(const Layout& l, const T* local, const T* next, Consumer<T>& y) {
  const int L = l.L;
  T one = K<T>(1);
  T filter = local[l.filter];
  const T* rf = local + l.rf;
  T is_first = rf[0];
  T is_not_last_round = filter - rf[1];
  T is_next_not_last_round = next[l.filter] - next[l.rf + 1];
  if (l.kind == KIND_G1)
    eval_g1_add(y, filter, local + l.a, local + l.b, local + l.c, local + l.aux);
  else if (l.kind == KIND_G2)
    eval_g2_add(y, filter, local + l.a, local + l.b, local + l.c, local + l.aux);
  else
    eval_fq_mul(y, filter, local + l.a, local + l.b, local + l.c, local + l.aux);
  // first round
  y.constraint(is_first * (local[l.flag_op] - one));
  eval_eq_n(y, is_first, local + l.reg0, local + l.b, L);
  T bit0 = local[l.bits];
  eval_eq_n(y, bit0 * is_first, local + l.reg1, local + l.c, L);
  eval_eq_n(y, (one - bit0) * is_first, local + l.reg1, local + l.a, L);
  if (l.kind == KIND_FQ) {
    // first round, a = 1 (fields/exp_stark.rs:255-260)
    for (int i = 0; i < 16; i++) y.constraint(is_first * (local[l.a + i] - (i == 0 ? one : T())));
  }
  // doubling/squaring step -> adding/multiplying step
  T f = local[l.flag_sq_nl];
  T nbit0 = next[l.bits];
  eval_eq_n(y, f, next + l.a, local + l.reg1, L);
  eval_eq_n(y, f, next + l.b, local + l.reg0, L);
  eval_eq_n(y, nbit0 * f, next + l.reg1, next + l.c, L);
  eval_eq_n(y, (one - nbit0) * f, next + l.reg1, next + l.a, L);
  eval_eq_n(y, f, next + l.reg0, local + l.reg0, L);
  y.constraint(f * (next[l.flag_op] - one));
  y.constraint(f * (next[l.flag_sq_nl] - T()));
  for (int i = 0; i < N_BITS; i++) y.constraint(f * (next[l.bits + i] - local[l.bits + (i + 1) % N_BITS]));
  // adding/multiplying step -> doubling/squaring step
  T g = local[l.flag_op];
  eval_eq_n(y, g, next + l.a, local + l.reg0, L);
  eval_eq_n(y, g, next + l.b, local + l.reg0, L);
  eval_eq_n(y, g, next + l.reg1, local + l.reg1, L);
  eval_eq_n(y, g, next + l.reg0, next + l.c, L);
  y.constraint(g * (next[l.flag_op] - T()));
  y.constraint(g * (next[l.flag_sq_nl] - is_next_not_last_round));
  for (int i = 0; i < N_BITS; i++) y.constraint(g * (next[l.bits + i] - local[l.bits + i]));
  // round flags, timestamp, filter
  eval_round_flags(y, filter, rf, next[l.rf + 2]);
  y.constraint(is_not_last_round * (next[l.ts] - local[l.ts]));
  y.constraint(is_not_last_round * (next[l.filter] - filter));
  // range counter
  T diff = next[l.range_counter] - local[l.range_counter];
  y.constraint_transition(diff * diff - diff);
  y.constraint_last_row(local[l.range_counter] - K<T>(((u64)1 << LIMB_BITS) - 1));
}

}  // namespace orc
