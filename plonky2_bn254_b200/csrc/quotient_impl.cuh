// K8 `quotient_eval_{g1,g2,fq}`: evaluates every constraint of the STARK at each point 7 * w^i of the
// quotient domain, folds them with the alpha challenges and divides by Z_H. One thread per point.
// Replaces, bit-exact, compute_quotient_polys of starky 0.4.0 prover.rs (un-vendored; reached from
// src/starks/common/prover.rs:55-65) together with the constraint emitters
//   eval_packed_generic     src/starks/curves/g1/scalar_mul_stark.rs:226-339,
//                           g2/scalar_mul_stark.rs:226-338, src/starks/fields/exp_stark.rs:208-327
//   eval_g1_add g1/add.rs:125-185, eval_g2_add g2/add.rs:133-196, eval_fq_mul fields/mul.rs:43-57
//   eval_modulus_zero modular/modulus_zero.rs:163-198, eval_is_modulus_zero is_modulus_zero.rs:69-84
//   eval_round_flags common/round_flags.rs:46-81
//   eval_packed_lookups_generic (starky lookup.rs), eval_cross_table_lookup_checks
//   (starky cross_table_lookup.rs)
//
// The reference folds constraints with a Horner recurrence acc = acc * alpha + c_k. All arithmetic
// is exact in F_p, so the algebraically equal form  sum_k alpha^(K-1-k) c_k  gives the identical
// canonical value; it is used here because (i) constraints that share a filter f contribute
// f * sum_k w_k d_k, one multiplication by f per group, and (ii) the inner sums are multiply-adds
// into a 160-bit lazy accumulator with one reduction per group instead of one per constraint.
// The constraint index k advances exactly in the reference's emission order.
// Bound: integer pipe (about 3 k lazy multiply-adds of limb-polynomial products plus ~1.6 k weighted
// terms per point for G1); HBM reads are (W + A) columns once plus the neighbouring "next" rows.
#pragma once
#include "quotient.h"

#define TG_P16_U64 {64839ULL, 55420ULL, 35862ULL, 15392ULL, 51853ULL, 26737ULL, 27281ULL, 38785ULL, 22621ULL, 33153ULL, 17846ULL, 47184ULL, 41001ULL, 57649ULL, 20082ULL, 12388ULL}

namespace quot {

template <int KIND>
struct LY {
  static constexpr int L = KIND == 0 ? 32 : KIND == 1 ? 64 : 16;
  static constexpr int AUXLEN = KIND == 0 ? 354 : KIND == 1 ? 708 : 80;
  static constexpr int reg0 = 0, reg1 = L, a = 2 * L, b = 3 * L, c = 4 * L, aux = 5 * L, bits = aux + AUXLEN,
                       rf = bits + 256, ts = rf + 5, flag_op = ts + 1, flag_sq = ts + 2, filter = ts + 3,
                       freq = ts + 4, rc = ts + 5, width = ts + 6, rc_lo = 2 * L, rc_hi = bits;
  static constexpr int NCOLS = rc_hi - rc_lo, NH = (NCOLS + 1) / 2;
  static constexpr int BASE_CONSTRAINTS = KIND == 0 ? 1111 : KIND == 1 ? 1693 : 770;
};

// Arithmetic inside the kernel works on lazy u64 representatives on the device (any u64 congruent to the value;
// gl::mul_lazy / add_lazy / sub_lazy and the 160-bit accumulator accept and produce them) and on canonical values
// in the host build; results are made canonical where a constraint group is closed (Emit::end_group).
PB_HD u64 q_add(u64 a, u64 b) {
#if defined(__CUDA_ARCH__)
  return gl::add_lazy(a, b);
#else
  return gl::add(a % gl::P, b % gl::P);
#endif
}
PB_HD u64 q_sub(u64 a, u64 b) {
#if defined(__CUDA_ARCH__)
  return gl::sub_lazy(a, b);
#else
  return gl::sub(a % gl::P, b % gl::P);
#endif
}
PB_HD u64 q_mul(u64 a, u64 b) {
#if defined(__CUDA_ARCH__)
  return gl::mul_lazy(a, b);
#else
  return gl::mul(a % gl::P, b % gl::P);
#endif
}
PB_HD u64 q_dbl(u64 a) { return q_add(a, a); }
PB_HD u64 q_sqr(u64 a) { return q_mul(a, a); }

struct Emit {
  const u64* w;
  int nch, k;
  gl::Acc acc[aux::MAXCH];
  u64 total[aux::MAXCH];
  u64* dbg = nullptr;
  int ng = 0;
  PB_HD Emit(const u64* w_, int nch_) : w(w_), nch(nch_), k(0) {
#pragma unroll
    for (int j = 0; j < aux::MAXCH; j++) total[j] = 0;
  }
  // constraint number k has value (group filter) * d
  PB_HD void term(u64 d) {
#pragma unroll
    for (int j = 0; j < aux::MAXCH; j++)
      if (j < nch) acc[j].mac(w[(size_t)k * nch + j], d);
    k++;
  }
  PB_HD void end_group(u64 f) {
#pragma unroll
    for (int j = 0; j < aux::MAXCH; j++)
      if (j < nch) {
        total[j] = gl::add(total[j], gl::mul(f, acc[j].reduce()));
        acc[j] = gl::Acc();
      }
    if (dbg) dbg[ng++] = total[0];
  }
};

#if PB_HOSTSIM
#define Q_NOINLINE static inline
#else
#define Q_NOINLINE static __host__ __device__ __noinline__
#endif

// The limb-polynomial helpers below are deliberately NOT inlined and keep real loops: a STARK has
// 5-22 of these products per row, and fully unrolled copies make the kernel tens of thousands of
// instructions long (and take ptxas tens of minutes) for no gain - they are multiply-add bound.

// 16 x 16 -> 31 product of limb polynomials over F_p (pol_mul_wide)
Q_NOINLINE void conv31(const u64* A, const u64* B, u64* out) {
#pragma unroll 1
  for (int k = 0; k < 31; k++) {
    gl::Acc s;
    const int lo = k > 15 ? k - 15 : 0, hi = k < 15 ? k : 15;
#pragma unroll 4
    for (int i = lo; i <= hi; i++) s.mac(A[i], B[k - i]);
    out[k] = s.reduce();
  }
}

// eval_modulus_zero: the 33 constraint values (before the filter) for aux cells at column auxcol
Q_NOINLINE void mz_values(const u64* tr, size_t stride, size_t i0, const u64* in, int auxcol, u64* out33) {
  const u64 P16[16] = TG_P16_U64;
  const u64* col = tr + (size_t)auxcol * stride + i0;
  u64 s = col[0];
  out33[0] = q_sub(q_sqr(s), s);
  u64 sign = q_sub(q_dbl(s), 1);
  u64 q[17];
#pragma unroll 1
  for (int i = 0; i < 17; i++) q[i] = q_mul(sign, col[(size_t)(1 + i) * stride]);
  const u64 base = (u64)1 << 16, offset = (u64)1 << 29;
  u64 ap_prev = 0;
#pragma unroll 1
  for (int k = 0; k < 32; k++) {
    gl::Acc s2;
    const int lo = k > 15 ? k - 15 : 0, hi = k < 16 ? k : 16;
#pragma unroll 4
    for (int i = lo; i <= hi; i++) s2.mac(q[i], P16[k - i]);
    u64 c = s2.reduce();
    u64 ap = 0;
    if (k < 31)
      ap = q_add(q_sub(col[(size_t)(18 + k) * stride], offset), q_mul(base, col[(size_t)(49 + k) * stride]));
    // (x - base) * ap(x): coefficient k is ap[k-1] - base * ap[k]
    c = q_add(c, q_sub(ap_prev, q_mul(base, ap)));
    if (k < 31) c = q_sub(c, in[k]);
    out33[1 + k] = c;
    ap_prev = ap;
  }
}

// All per-thread scratch arrays live in ONE function-scope object. They used to be block-scope locals
// of the (force-inlined) helpers; nvcc 12.9's stack colouring then overlapped two arrays that are live
// at the same time (a block-scope d0[31] and ext_conv's t[31] shared one slot in the G2 kernel), which
// made the G2 quotient wrong on the device only. One object has one lifetime, so nothing can overlap.
struct Scratch {
  u64 l0[16], l1[16], u0[16], u1[16];
  u64 c0[31], c1[31], d0[31], d1[31], t[31], v[33];
};

// The constraint system is evaluated in NPASS kernels over the same points (PASS = 0 .. NPASS-1), each adding its
// share of sum_k w_k c_k to the output; the last one divides by Z_H. One kernel for everything needed 104
// registers and a 2.4 KB scratch frame per thread and ran latency-bound at 24 % warps active (ncu,
// profiles/r1_kernels_final.md); the passes are smaller programs with higher occupancy.
//   PASS 0  addition / multiplication gadget, first part (is-zero test, slope)
//   PASS 1  addition gadget, second part (x and y modulus-zero checks); empty for fq_exp
//   PASS 2  register equalities, bit rotations, round flags, timestamp, range counter
//   PASS 3  logUp lookups, cross-table lookups, division by Z_H
static constexpr int NPASS = 4;

template <int KIND, int PASS>
struct QuotientK {
  Params p;
  typedef LY<KIND> Y;
  // constraints emitted before each pass (SURVEY.md Appendix D order)
  static constexpr int N_ADD1 = KIND == 0 ? 132 : KIND == 1 ? 264 : 33;
  static constexpr int N_ADD = KIND == 0 ? 198 : KIND == 1 ? 396 : 33;
  static constexpr int K0 = PASS == 0 ? 0 : PASS == 1 ? N_ADD1 : PASS == 2 ? N_ADD : Y::BASE_CONSTRAINTS;

  PB_HD u64 TL(size_t i0, int c) const { return p.tr[(size_t)c * p.tr_stride + i0]; }

  // eval_modulus_zero: 33 terms (the caller closes the group with the filter)
  PB_HD void mz(Emit& E, Scratch& S, size_t i0, const u64* in, int auxcol) const {
    mz_values(p.tr, p.tr_stride, i0, in, auxcol, S.v);
#pragma unroll 1
    for (int k = 0; k < 33; k++) E.term(S.v[k]);
  }

  // eval_is_modulus_zero: 49 terms under one filter. input columns: in_b - in_a (16 limbs)
  PB_HD void imz(Emit& E, Scratch& S, size_t i0, int col_b, int col_a, int is_zero_col, int auxcol) const {
    u64 *dx = S.u0, *iv = S.u1, *in = S.c0;
#pragma unroll 1
    for (int i = 0; i < 16; i++) {
      dx[i] = q_sub(TL(i0, col_b + i), TL(i0, col_a + i));
      iv[i] = TL(i0, auxcol + i);
    }
    conv31(dx, iv, in);
    u64 is_zero = TL(i0, is_zero_col);
    in[0] = q_add(in[0], q_sub(is_zero, 1));
    mz(E, S, i0, in, auxcol + 16);
#pragma unroll 1
    for (int i = 0; i < 16; i++) E.term(q_mul(dx[i], is_zero));
  }

  PB_HD void ld16(size_t i0, int col, u64 out[16]) const {
#pragma unroll 1
    for (int i = 0; i < 16; i++) out[i] = TL(i0, col + i);
  }

  PB_HD void add_g1(Emit& E, Scratch& S, size_t i0, u64 filter) const {
    const int A = Y::aux, a = Y::a, b = Y::b, c = Y::c;
    u64 *lam = S.l0, *t0 = S.u0, *in = S.c0, *in2 = S.d0;
    if (PASS == 1) {
      ld16(i0, A + 98, lam);
      add_g1_xy(E, S, i0, filter);
      return;
    }
    imz(E, S, i0, b, a, A, A + 1);
    E.end_group(filter);
    u64 is_x_eq = TL(i0, A), is_x_eq_filter = TL(i0, A + 97);
    E.term(q_sub(q_mul(filter, is_x_eq), is_x_eq_filter));
    E.end_group(1);
    ld16(i0, A + 98, lam);
    // a.x != b.x : lambda * dx - (b.y - a.y)
#pragma unroll 1
    for (int i = 0; i < 16; i++) t0[i] = q_sub(TL(i0, b + i), TL(i0, a + i));
    conv31(lam, t0, in);
#pragma unroll 1
    for (int i = 0; i < 16; i++) in[i] = q_sub(in[i], q_sub(TL(i0, b + 16 + i), TL(i0, a + 16 + i)));
    mz(E, S, i0, in, A + 114);
    E.end_group(q_sub(filter, is_x_eq_filter));
    // a.x == b.x : 2 lambda a.y - 3 a.x^2
    ld16(i0, a, t0);
    conv31(t0, t0, in2);
    ld16(i0, a + 16, t0);
    conv31(lam, t0, in);
#pragma unroll 1
    for (int i = 0; i < 31; i++) in[i] = q_sub(q_dbl(in[i]), q_add(q_dbl(in2[i]), in2[i]));
    mz(E, S, i0, in, A + 114);
#pragma unroll 1
    for (int i = 0; i < 16; i++) E.term(q_sub(TL(i0, a + 16 + i), TL(i0, b + 16 + i)));
    E.end_group(is_x_eq_filter);
  }
  PB_HD void add_g1_xy(Emit& E, Scratch& S, size_t i0, u64 filter) const {
    const int A = Y::aux, a = Y::a, b = Y::b, c = Y::c;
    u64 *lam = S.l0, *t0 = S.u0, *in = S.c0;
    // x : lambda^2 - (a.x + b.x + c.x)
    conv31(lam, lam, in);
#pragma unroll 1
    for (int i = 0; i < 16; i++)
      in[i] = q_sub(in[i], q_add(q_add(TL(i0, a + i), TL(i0, b + i)), TL(i0, c + i)));
    mz(E, S, i0, in, A + 194);
    // y : lambda (c.x - a.x) + c.y + a.y
#pragma unroll 1
    for (int i = 0; i < 16; i++) t0[i] = q_sub(TL(i0, c + i), TL(i0, a + i));
    conv31(lam, t0, in);
#pragma unroll 1
    for (int i = 0; i < 16; i++) in[i] = q_add(in[i], q_add(TL(i0, c + 16 + i), TL(i0, a + 16 + i)));
    mz(E, S, i0, in, A + 274);
    E.end_group(filter);
  }

  // (x * y) over Fq2 on limb polynomials: c0 = x0 y0 - x1 y1, c1 = x0 y1 + x1 y0
  PB_HD void ext_conv(Scratch& S, const u64 x0[16], const u64 x1[16], const u64 y0[16], const u64 y1[16], u64 c0[31],
                      u64 c1[31]) const {
    u64* t = S.t;
    conv31(x0, y0, c0);
    conv31(x1, y1, t);
#pragma unroll 1
    for (int i = 0; i < 31; i++) c0[i] = q_sub(c0[i], t[i]);
    conv31(x0, y1, c1);
    conv31(x1, y0, t);
#pragma unroll 1
    for (int i = 0; i < 31; i++) c1[i] = q_add(c1[i], t[i]);
  }

  PB_HD void add_g2(Emit& E, Scratch& S, size_t i0, u64 filter) const {
    const int A = Y::aux, a = Y::a, b = Y::b;  // points: x.c0 | x.c1 | y.c0 | y.c1
    if (PASS == 1) {
      ld16(i0, A + 196, S.l0);
      ld16(i0, A + 212, S.l1);
      add_g2_xy(E, S, i0, filter);
      return;
    }
    u64 is_x_eq = TL(i0, A), z0 = TL(i0, A + 1), z1 = TL(i0, A + 2);
    E.term(q_sub(q_mul(z0, z1), is_x_eq));
    imz(E, S, i0, b, a, A + 1, A + 3);
    imz(E, S, i0, b + 16, a + 16, A + 2, A + 99);
    E.end_group(filter);
    u64 is_x_eq_filter = TL(i0, A + 195);
    E.term(q_sub(q_mul(filter, is_x_eq), is_x_eq_filter));
    E.end_group(1);
    u64 *l0 = S.l0, *l1 = S.l1, *u0 = S.u0, *u1 = S.u1, *c0 = S.c0, *c1 = S.c1, *d0 = S.d0, *d1 = S.d1;
    ld16(i0, A + 196, l0);
    ld16(i0, A + 212, l1);
    // lambda * dx - dy
#pragma unroll 1
    for (int i = 0; i < 16; i++) {
      u0[i] = q_sub(TL(i0, b + i), TL(i0, a + i));
      u1[i] = q_sub(TL(i0, b + 16 + i), TL(i0, a + 16 + i));
    }
    ext_conv(S, l0, l1, u0, u1, c0, c1);
#pragma unroll 1
    for (int i = 0; i < 16; i++) {
      c0[i] = q_sub(c0[i], q_sub(TL(i0, b + 32 + i), TL(i0, a + 32 + i)));
      c1[i] = q_sub(c1[i], q_sub(TL(i0, b + 48 + i), TL(i0, a + 48 + i)));
    }
    mz(E, S, i0, c0, A + 228);
    mz(E, S, i0, c1, A + 308);
    E.end_group(q_sub(filter, is_x_eq_filter));
    // 2 lambda a.y - 3 a.x^2
    {
      ld16(i0, a, u0);
      ld16(i0, a + 16, u1);
      ext_conv(S, u0, u1, u0, u1, d0, d1);
      ld16(i0, a + 32, u0);
      ld16(i0, a + 48, u1);
      ext_conv(S, l0, l1, u0, u1, c0, c1);
#pragma unroll 1
      for (int i = 0; i < 31; i++) {
        c0[i] = q_sub(q_dbl(c0[i]), q_add(q_dbl(d0[i]), d0[i]));
        c1[i] = q_sub(q_dbl(c1[i]), q_add(q_dbl(d1[i]), d1[i]));
      }
    }
    mz(E, S, i0, c0, A + 228);
    mz(E, S, i0, c1, A + 308);
#pragma unroll 2
    for (int i = 0; i < 32; i++) E.term(q_sub(TL(i0, a + 32 + i), TL(i0, b + 32 + i)));
    E.end_group(is_x_eq_filter);
  }
  PB_HD void add_g2_xy(Emit& E, Scratch& S, size_t i0, u64 filter) const {
    const int A = Y::aux, a = Y::a, b = Y::b, c = Y::c;
    u64 *l0 = S.l0, *l1 = S.l1, *u0 = S.u0, *u1 = S.u1, *c0 = S.c0, *c1 = S.c1;
    // x
    ext_conv(S, l0, l1, l0, l1, c0, c1);
#pragma unroll 1
    for (int i = 0; i < 16; i++) {
      c0[i] = q_sub(c0[i], q_add(q_add(TL(i0, a + i), TL(i0, b + i)), TL(i0, c + i)));
      c1[i] = q_sub(c1[i], q_add(q_add(TL(i0, a + 16 + i), TL(i0, b + 16 + i)), TL(i0, c + 16 + i)));
    }
    mz(E, S, i0, c0, A + 388);
    mz(E, S, i0, c1, A + 468);
    // y
#pragma unroll 1
    for (int i = 0; i < 16; i++) {
      u0[i] = q_sub(TL(i0, c + i), TL(i0, a + i));
      u1[i] = q_sub(TL(i0, c + 16 + i), TL(i0, a + 16 + i));
    }
    ext_conv(S, l0, l1, u0, u1, c0, c1);
#pragma unroll 1
    for (int i = 0; i < 16; i++) {
      c0[i] = q_add(c0[i], q_add(TL(i0, c + 32 + i), TL(i0, a + 32 + i)));
      c1[i] = q_add(c1[i], q_add(TL(i0, c + 48 + i), TL(i0, a + 48 + i)));
    }
    mz(E, S, i0, c0, A + 548);
    mz(E, S, i0, c1, A + 628);
    E.end_group(filter);
  }

  PB_HD void mul_fq(Emit& E, Scratch& S, size_t i0, u64 filter) const {
    u64 *x = S.u0, *y = S.u1, *in = S.c0;
    ld16(i0, Y::a, x);
    ld16(i0, Y::b, y);
    conv31(x, y, in);
#pragma unroll 1
    for (int i = 0; i < 16; i++) in[i] = q_sub(in[i], TL(i0, Y::c + i));
    mz(E, S, i0, in, Y::aux);
    E.end_group(filter);
  }

  // n terms  X[i] - Y[i]  where X, Y are columns of the local (sel 0) or next (sel 1) row
  PB_HD void eq_terms(Emit& E, size_t ix, int colx, size_t iy, int coly, int n) const {
#pragma unroll 2
    for (int i = 0; i < n; i++) E.term(q_sub(TL(ix, colx + i), TL(iy, coly + i)));
  }

  PB_HD void operator()(size_t il) const {
    // il: index inside this call's block (the storage index), i: index in the quotient domain
    const size_t i = p.i_base + il;
    const size_t i_next = p.wrap ? ((il + 2) & (p.size - 1)) : il + 2;  // next_step = 2^quotient_degree_bits = 2
    const size_t i0 = il * p.step, i1 = i_next * p.step;
    const int nch = p.ch.nch;
    const int L = Y::L;
    Emit E(p.weights, nch);
    E.k = K0;
    if (p.dbg && i == p.dbg_point) E.dbg = p.dbg + 64 * PASS;
    const u64 filter = TL(i0, Y::filter);
    if (PASS <= 1) {
      Scratch S;
      if (KIND == 0)
        add_g1(E, S, i0, filter);
      else if (KIND == 1)
        add_g2(E, S, i0, filter);
      else if (PASS == 0)
        mul_fq(E, S, i0, filter);
      if (E.k != (PASS == 0 ? N_ADD1 : N_ADD)) *p.err = 1;
    } else {
      const u64 is_first = TL(i0, Y::rf), is_last = TL(i0, Y::rf + 1);
      const u64 x = gl::mul(gl::COSET_SHIFT, ntt::tpow(p.t.fwd_lo, p.t.fwd_hi, (u64)i << (ntt::LOG_M - p.log_size)));
      const u64 z_last = gl::sub(x, p.g_inv);
      const u64 z_h = p.zh[i & 1];
      const u64 l_last = gl::mul(z_h, gl::inv(gl::mul(p.n_field, gl::sub(gl::mul(p.g, x), 1))));
      if (PASS == 2) {
        // first round
        E.term(q_sub(TL(i0, Y::flag_op), 1));
        eq_terms(E, i0, Y::reg0, i0, Y::b, L);
        E.end_group(is_first);
        const u64 bit0 = TL(i0, Y::bits);
        eq_terms(E, i0, Y::reg1, i0, Y::c, L);
        E.end_group(q_mul(bit0, is_first));
        eq_terms(E, i0, Y::reg1, i0, Y::a, L);
        E.end_group(q_mul(q_sub(1, bit0), is_first));
        if (KIND == 2) {
#pragma unroll 1
          for (int k = 0; k < 16; k++) E.term(q_sub(TL(i0, Y::a + k), k == 0 ? 1 : 0));
          E.end_group(is_first);
        }
        // doubling / squaring step -> adding / multiplying step
        const u64 fs = TL(i0, Y::flag_sq), nbit0 = TL(i1, Y::bits);
        eq_terms(E, i1, Y::a, i0, Y::reg1, L);
        eq_terms(E, i1, Y::b, i0, Y::reg0, L);
        E.end_group(fs);
        eq_terms(E, i1, Y::reg1, i1, Y::c, L);
        E.end_group(q_mul(nbit0, fs));
        eq_terms(E, i1, Y::reg1, i1, Y::a, L);
        E.end_group(q_mul(q_sub(1, nbit0), fs));
        eq_terms(E, i1, Y::reg0, i0, Y::reg0, L);
        E.term(q_sub(TL(i1, Y::flag_op), 1));
        E.term(TL(i1, Y::flag_sq));
#pragma unroll 2
        for (int k = 0; k < 256; k++) E.term(q_sub(TL(i1, Y::bits + k), TL(i0, Y::bits + ((k + 1) & 255))));
        E.end_group(fs);
        // adding / multiplying step -> doubling / squaring step
        const u64 g = TL(i0, Y::flag_op);
        const u64 is_next_not_last = q_sub(TL(i1, Y::filter), TL(i1, Y::rf + 1));
        eq_terms(E, i1, Y::a, i0, Y::reg0, L);
        eq_terms(E, i1, Y::b, i0, Y::reg0, L);
        eq_terms(E, i1, Y::reg1, i0, Y::reg1, L);
        eq_terms(E, i1, Y::reg0, i1, Y::c, L);
        E.term(TL(i1, Y::flag_op));
        E.term(q_sub(TL(i1, Y::flag_sq), is_next_not_last));
        eq_terms(E, i1, Y::bits, i0, Y::bits, 256);
        E.end_group(g);
        // round flags (8 constraints, written out with their own factors)
        {
          const u64 counter = TL(i0, Y::rf + 2), inv_c = TL(i0, Y::rf + 3), inv_cp = TL(i0, Y::rf + 4);
          const u64 next_counter = TL(i1, Y::rf + 2);
          const u64 not_filter = q_sub(1, filter);
          E.term(q_mul(not_filter, is_first));
          E.term(q_mul(not_filter, is_last));
          E.term(q_mul(filter, q_sub(q_mul(counter, inv_c), q_sub(1, is_first))));
          E.term(q_mul(q_mul(filter, counter), is_first));
          const u64 cprime = q_sub(counter, (u64)(tg::PERIOD - 1));
          E.term(q_mul(filter, q_sub(q_mul(cprime, inv_cp), q_sub(1, is_last))));
          E.term(q_mul(q_mul(filter, cprime), is_last));
          E.term(q_mul(q_mul(filter, q_sub(1, is_last)), q_sub(q_sub(next_counter, counter), 1)));
          E.term(q_mul(q_mul(filter, is_last), next_counter));
          E.end_group(1);
        }
        // timestamp and filter continuity
        E.term(q_sub(TL(i1, Y::ts), TL(i0, Y::ts)));
        E.term(q_sub(TL(i1, Y::filter), filter));
        E.end_group(q_sub(filter, is_last));
        // range counter
        {
          const u64 rc = TL(i0, Y::rc), d = q_sub(TL(i1, Y::rc), rc);
          E.term(q_sub(q_sqr(d), d));
          E.end_group(z_last);
          E.term(q_sub(rc, 65535));
          E.end_group(l_last);
        }
        if (E.k != Y::BASE_CONSTRAINTS) *p.err = 1;
      } else {
        const u64 l_first = gl::mul(z_h, gl::inv(gl::mul(p.n_field, gl::sub(x, 1))));
        // logUp lookups (auxiliary columns: per challenge NH helpers then Z)
        const size_t a0 = i0, a1 = i1;
        for (int j = 0; j < nch; j++) {
          const u64 beta = p.ch.beta[j];
          const u64* hcol = p.ax + (size_t)(j * (Y::NH + 1)) * p.ax_stride;
          gl::Acc hsum;
          for (int k = 0; k < Y::NH; k++) {
            const u64 h = hcol[(size_t)k * p.ax_stride + a0];
            hsum.addv(h);
            const u64 c0 = q_add(TL(i0, Y::rc_lo + 2 * k), beta);
            if (2 * k + 1 < Y::NCOLS) {
              const u64 c1 = q_add(TL(i0, Y::rc_lo + 2 * k + 1), beta);
              E.term(q_sub(q_sub(q_mul(q_mul(c1, c0), h), c1), c0));
            } else {
              E.term(q_sub(q_mul(c0, h), 1));
            }
          }
          E.end_group(1);
          const u64 z = hcol[(size_t)Y::NH * p.ax_stride + a0], nz = hcol[(size_t)Y::NH * p.ax_stride + a1];
          E.term(z);
          E.end_group(l_first);
          const u64 hs = hsum.reduce();
          const u64 table = q_add(TL(i0, Y::rc), beta);
          const u64 yv = q_sub(q_mul(hs, table), TL(i0, Y::freq));
          E.term(q_sub(q_mul(q_sub(nz, z), table), yv));
          E.end_group(1);
        }
        // cross-table lookups: CTL-major, challenge-minor. comb = sum_k v_k beta^k + gamma as a lazy dot product with
        // the beta powers (p.bpow[j][k]); the 16 scalar limbs (le_bits sums) are computed once for all challenges.
        {
          const u64* zc = p.ax + (size_t)((Y::NH + 1) * nch) * p.ax_stride;
          u64 limbs[16];
#pragma unroll 1
          for (int k = 0; k < 16; k++) {
            gl::Acc s16;
#pragma unroll 4
            for (int b = 0; b < 16; b++) s16.mac(TL(i0, Y::bits + 16 * k + b), (u64)1 << b);
            limbs[k] = s16.reduce();
          }
          for (int c = 0; c < 2; c++) {
            const u64 f = c == 0 ? is_first : is_last;
            // looked columns: c = 0: x (b), offset (a, curves only), s limbs, timestamp;  c = 1: reg1, timestamp
            for (int j = 0; j < nch; j++) {
              const u64* bp = p.bpow + (size_t)j * p.bpow_stride;
              gl::Acc cs;
              int e = 0;
              if (c == 0) {
#pragma unroll 4
                for (int k = 0; k < L; k++) cs.mac(TL(i0, Y::b + k), bp[e++]);
                if (KIND != 2) {
#pragma unroll 4
                  for (int k = 0; k < L; k++) cs.mac(TL(i0, Y::a + k), bp[e++]);
                }
#pragma unroll 4
                for (int k = 0; k < 16; k++) cs.mac(limbs[k], bp[e++]);
              } else {
#pragma unroll 4
                for (int k = 0; k < L; k++) cs.mac(TL(i0, Y::reg1 + k), bp[e++]);
              }
              cs.mac(TL(i0, Y::ts), bp[e]);
              cs.addv(p.ch.gamma[j]);
              const u64 comb = cs.reduce();
              const u64 lz = zc[(size_t)(c * nch + j) * p.ax_stride + a0];
              const u64 nz = zc[(size_t)(c * nch + j) * p.ax_stride + a1];
              E.term(q_sub(q_mul(comb, lz), f));
              E.end_group(l_last);
              E.term(q_sub(q_mul(comb, q_sub(lz, nz)), f));
              E.end_group(z_last);
            }
          }
        }
        if (E.k != Y::BASE_CONSTRAINTS + (Y::NH + 2) * nch + 4 * nch) *p.err = 1;
      }
    }
    // accumulate this pass's share; the last pass divides by Z_H
#pragma unroll
    for (int j = 0; j < aux::MAXCH; j++)
      if (j < nch) {
        u64 v = E.total[j];
        u64* o = p.out + (size_t)j * p.out_stride + il;
        if (PASS > 0) v = gl::add(*o, v);
        if (PASS == NPASS - 1) v = gl::mul(v, p.zh_inv[i & 1]);
        *o = v;
      }
  }
};

}  // namespace quot
