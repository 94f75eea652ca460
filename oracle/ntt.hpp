// ORACLE (test infrastructure, NOT product code): FFT / iFFT / coset-FFT over Goldilocks as
// specified by plonky2_field 0.2.2 field/src/{fft.rs,polynomial/mod.rs} (un-vendored; call
// site common/prover.rs:31-38 via PolynomialBatch::from_values). Conventions:
//   fft(c)[j]  = sum_i c[i] w^(ij)        (natural order in and out, w = root_of_unity(log n))
//   ifft(v)[i] = n^-1 sum_j v[j] w^(-ij)
//   coset_fft(c, s)[j] = sum_i c[i] (s w^j)^i
// Plain iterative radix-2; the oracle favours obviousness over speed.
#pragma once
#include "gl.hpp"

namespace orc {

static inline void bit_reverse_permute(u64* a, size_t n) {
  unsigned lg = log2_strict(n);
  for (size_t i = 0; i < n; i++) {
    size_t j = reverse_bits(i, lg);
    if (i < j) {
      u64 t = a[i];
      a[i] = a[j];
      a[j] = t;
    }
  }
}

// in-place, natural -> natural, using root w of order n
static inline void ntt_inplace(u64* a, size_t n, u64 w) {
  if (n <= 1) return;
  bit_reverse_permute(a, n);
  unsigned lg = log2_strict(n);
  // twiddle table w^0..w^(n/2-1)
  std::vector<u64> tw(n / 2);
  {
    const size_t B = 1024;  // blocked so the table build is not one long dependency chain
    u64 wB = gl_pow(w, B);
    u64 start = 1;
    for (size_t b0 = 0; b0 < n / 2; b0 += B) {
      u64 x = start;
      for (size_t i = b0; i < b0 + B && i < n / 2; i++) {
        tw[i] = x;
        x = gl_mul(x, w);
      }
      start = gl_mul(start, wB);
    }
  }
  for (unsigned s = 1; s <= lg; s++) {
    size_t m = (size_t)1 << s, half = m >> 1, step = n / m;
    for (size_t k = 0; k < n; k += m)
      for (size_t j = 0; j < half; j++) {
        u64 t = gl_mul(tw[j * step], a[k + j + half]);
        u64 u = a[k + j];
        a[k + j] = gl_add(u, t);
        a[k + j + half] = gl_sub(u, t);
      }
  }
}

static inline void fft(std::vector<u64>& a) { ntt_inplace(a.data(), a.size(), gl_root_of_unity(log2_strict(a.size()))); }
static inline void ifft(std::vector<u64>& a) {
  size_t n = a.size();
  ntt_inplace(a.data(), n, gl_inv(gl_root_of_unity(log2_strict(n))));
  u64 ninv = gl_inv((u64)n % GL_P);
  for (auto& x : a) x = gl_mul(x, ninv);
}
static inline void coset_fft(std::vector<u64>& a, u64 shift) {
  u64 p = 1;
  for (auto& x : a) {
    x = gl_mul(x, p);
    p = gl_mul(p, shift);
  }
  fft(a);
}
static inline void coset_ifft(std::vector<u64>& a, u64 shift) {
  ifft(a);
  u64 si = gl_inv(shift), p = 1;
  for (auto& x : a) {
    x = gl_mul(x, p);
    p = gl_mul(p, si);
  }
}
// PolynomialValues::lde_onto_coset(rate_bits): ifft, zero-pad, coset_fft(shift 7), natural order
static inline std::vector<u64> lde_onto_coset(std::vector<u64> v, unsigned rate_bits) {
  ifft(v);
  v.resize(v.size() << rate_bits, 0);
  coset_fft(v, GL_COSET_SHIFT);
  return v;
}

// extension-field polynomials: components transform independently (roots and shift are in F)
static inline void coset_fft_ext(std::vector<Fp2>& a, u64 shift) {
  std::vector<u64> c0(a.size()), c1(a.size());
  for (size_t i = 0; i < a.size(); i++) {
    c0[i] = a[i].c[0];
    c1[i] = a[i].c[1];
  }
  coset_fft(c0, shift);
  coset_fft(c1, shift);
  for (size_t i = 0; i < a.size(); i++) a[i] = Fp2(c0[i], c1[i]);
}

// PolynomialCoeffs::eval at an extension point (Horner)
static inline Fp2 poly_eval_ext(const std::vector<u64>& c, Fp2 x) {
  Fp2 acc;
  for (size_t i = c.size(); i-- > 0;) acc = acc * x + Fp2(c[i], 0);
  return acc;
}
static inline u64 poly_eval_base(const std::vector<u64>& c, u64 x) {
  u64 acc = 0;
  for (size_t i = c.size(); i-- > 0;) acc = gl_add(gl_mul(acc, x), c[i]);
  return acc;
}

}  // namespace orc
