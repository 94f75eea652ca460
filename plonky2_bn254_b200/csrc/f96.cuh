// Radix-2/4/8/16 Goldilocks DFT butterflies whose internal twiddles are powers of two.
//
// p = 2^64 - 2^32 + 1 divides 2^96 + 1, so 2 is a 192-th root of unity: with phi = 2^32 we have
// phi^2 = phi - 1 and phi^3 = -1 (mod p). The roots the NTT uses (plonky2_field 0.2.2
// GoldilocksField::primitive_root_of_unity, un-vendored; reached from
// src/starks/common/prover.rs:31-38 through PolynomialBatch::from_values) are
//   w_4 = 2^48, w_8 = 2^(24*13), w_16 = 2^(12*13), w_32 = 2^(6*13), w_64 = 2^(3*13)      (mod p)
// so a radix-R DFT, R <= 16, needs no field multiplication at all: only shifts, word rotations and
// additions, and its output for the standard root is the output for the root 2^(192/R) permuted by
// k -> 13 k mod R.
//
// Inside a butterfly values are signed 96-bit integers (three 32-bit words, two's complement) that are
// only congruent to the field element: an addition or subtraction is a 3-instruction carry chain with no
// modular correction, a multiplication by 2^e is "fold by phi^q" (8 instructions, result < 2^65.1 in
// magnitude) followed by a 3-instruction funnel shift by s < 32, e = 32 q + s; one reduction to a 64-bit
// representative per output. Magnitudes (checked by tools/f96_bounds.py): radix-16 outputs stay below
// 2^93.1, the format holds 2^95.
#pragma once
#include "gl.cuh"

namespace f96 {

struct V {
  u32 w0, w1, w2;  // value = w0 + w1 2^32 + (int32)w2 2^64
};

PB_HD V from64(u64 x) {
  V r;
  r.w0 = (u32)x;
  r.w1 = (u32)(x >> 32);
  r.w2 = 0;
  return r;
}

PB_HD V add(V a, V b) {
  V r;
#ifdef __CUDA_ARCH__
  asm("add.cc.u32 %0, %3, %6;\n\taddc.cc.u32 %1, %4, %7;\n\taddc.u32 %2, %5, %8;"
      : "=&r"(r.w0), "=&r"(r.w1), "=r"(r.w2)
      : "r"(a.w0), "r"(a.w1), "r"(a.w2), "r"(b.w0), "r"(b.w1), "r"(b.w2));
#else
  u64 s0 = (u64)a.w0 + b.w0;
  u64 s1 = (u64)a.w1 + b.w1 + (s0 >> 32);
  r.w0 = (u32)s0;
  r.w1 = (u32)s1;
  r.w2 = a.w2 + b.w2 + (u32)(s1 >> 32);
#endif
  return r;
}
PB_HD V sub(V a, V b) {
  V r;
#ifdef __CUDA_ARCH__
  asm("sub.cc.u32 %0, %3, %6;\n\tsubc.cc.u32 %1, %4, %7;\n\tsubc.u32 %2, %5, %8;"
      : "=&r"(r.w0), "=&r"(r.w1), "=r"(r.w2)
      : "r"(a.w0), "r"(a.w1), "r"(a.w2), "r"(b.w0), "r"(b.w1), "r"(b.w2));
#else
  u64 d0 = (u64)a.w0 - b.w0;
  u64 d1 = (u64)a.w1 - b.w1 - ((d0 >> 32) & 1);
  r.w0 = (u32)d0;
  r.w1 = (u32)d1;
  r.w2 = a.w2 - b.w2 - (u32)((d1 >> 32) & 1);
#endif
  return r;
}

// T == V * phi^Q (mod p), |T| < 2^65 + 2^34.  Uses phi^2 = phi - 1, phi^3 = -1:
//   Q = 0:  (w0 - w2) + (w1 + w2) phi
//   Q = 1: -(w1 + w2) + (w0 + w1) phi
//   Q = 2: -(w0 + w1) + (w0 - w2) phi
template <int Q>
PB_HD V fold(V v) {
  const i64 s2 = (i64)(int32_t)v.w2;
  i64 A, B;
  if (Q == 0) {
    A = (i64)v.w0 - s2;
    B = (i64)v.w1 + s2;
  } else if (Q == 1) {
    A = -((i64)v.w1 + s2);
    B = (i64)v.w0 + (i64)v.w1;
  } else {
    A = -((i64)v.w0 + (i64)v.w1);
    B = (i64)v.w0 - s2;
  }
  const i64 U = (A >> 32) + B;  // T = (A mod 2^32) + U 2^32
  V r;
  r.w0 = (u32)A;
  r.w1 = (u32)U;
  r.w2 = (u32)(U >> 32);
  return r;
}

template <int S>
PB_HD V shl(V v) {
  if (S == 0) return v;
  V r;
#ifdef __CUDA_ARCH__
  r.w2 = __funnelshift_l(v.w1, v.w2, S);
  r.w1 = __funnelshift_l(v.w0, v.w1, S);
  r.w0 = v.w0 << S;
#else
  r.w2 = (v.w2 << S) | (v.w1 >> ((32 - S) & 31));
  r.w1 = (v.w1 << S) | (v.w0 >> ((32 - S) & 31));
  r.w0 = v.w0 << S;
#endif
  return r;
}

// V * 2^E (mod p) for 0 <= E < 96, any |V| < 2^95; result magnitude < 2^(65.01 + E mod 32)
template <int E>
PB_HD V mulpow(V v) {
  static_assert(E >= 0 && E < 96, "exponent");
  if (E == 0) return v;
  return shl<E & 31>(fold<(E >> 5)>(v));
}

// a 64-bit representative of V (any u64 congruent to V mod p), |V| < 2^95
PB_HD u64 to64(V v) {
  const i64 s2 = (i64)(int32_t)v.w2;
  const i64 A = (i64)v.w0 - s2, B = (i64)v.w1 + s2;
  // Z = A + B 2^32 in (-2^63.6, 2^64.6): Z = r + k 2^64, k in {-1, 0, 1}; 2^64 == 2^32 - 1, and r + k (2^32 - 1)
  // cannot wrap a second time
  const __int128 Z = (__int128)A + ((__int128)B << 32);
  const u64 r = (u64)Z;
  const i64 k = (i64)(Z >> 64);
  return r + (u64)(k * (i64)0xFFFFFFFFLL);
}

// ---- in-register DFT with the root 2^(192 / R): y[k] = sum_j x[j] 2^((192 / R) j k) ------------------
// decimation in time: bit-reversed loads (compile-time register renaming), then LR stages of
// (a, b) -> (a + t b, a - t b) with t = 2^((192 >> s) i) for position i of a 2^s block: every exponent is < 96.
template <int LR>
struct Rev {
  static PB_HD constexpr int of(int j) {
    int r = 0;
    for (int b = 0; b < LR; b++) r |= ((j >> b) & 1) << (LR - 1 - b);
    return r;
  }
};

template <int LR, int S, int BLK, int I>
struct Bfly {
  static PB_HD void run(V* t) {
    constexpr int HALF = 1 << (S - 1);
    constexpr int E = (192 >> S) * I;
    const V a = t[BLK + I], b = mulpow<E>(t[BLK + I + HALF]);
    t[BLK + I] = add(a, b);
    t[BLK + I + HALF] = sub(a, b);
    if constexpr (I + 1 < HALF)
      Bfly<LR, S, BLK, I + 1>::run(t);
    else if constexpr (BLK + 2 * HALF < (1 << LR))
      Bfly<LR, S, BLK + 2 * HALF, 0>::run(t);
    else if constexpr (S < LR)
      Bfly<LR, S + 1, 0, 0>::run(t);
  }
};

template <int LR, int J>
struct Load {
  static PB_HD void run(V* t, const u64* x) {
    t[Rev<LR>::of(J)] = from64(x[J]);
    if constexpr (J + 1 < (1 << LR)) Load<LR, J + 1>::run(t, x);
  }
};

// output index map for the STANDARD roots: w_R = 2^((192 / R) 13), so
//   forward  y[k] = t[13 k mod R],   inverse (root w_R^-1)  y[k] = t[-13 k mod R]
template <int LR, bool INV>
struct OutIdx {
  static PB_HD constexpr int of(int k) {
    constexpr int R = 1 << LR;
    int u = (13 * k) % R;
    return INV ? (R - u) % R : u;
  }
};
template <int LR, bool INV, int K>
struct Store {
  static PB_HD void run(const V* t, u64* y) {
    y[K] = to64(t[OutIdx<LR, INV>::of(K)]);
    if constexpr (K + 1 < (1 << LR)) Store<LR, INV, K + 1>::run(t, y);
  }
};

// x[0..R) (any u64 representatives) -> y[k] = sum_j x[j] w_R^(+-jk), any u64 representatives. In place is fine.
template <int LR, bool INV>
PB_HD void dft(u64* x) {
  if constexpr (LR == 0) return;
  V t[1 << LR];
  Load<LR, 0>::run(t, x);
  Bfly<LR, 1, 0, 0>::run(t);
  Store<LR, INV, 0>::run(t, x);
}

}  // namespace f96
