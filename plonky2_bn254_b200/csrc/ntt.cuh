// K3 `lde_batch`: Goldilocks low-degree extension of a column-major matrix.
// Replaces PolynomialBatch::from_values / from_coeffs of plonky2 0.2.2 fri/oracle.rs (un-vendored;
// call site src/starks/common/prover.rs:31-38): per column  iNTT(n) -> * 7^i -> zero-pad ->
// NTT(n * 2^r), natural-order output (value j = P(7 * w^j)).
//
// Decomposition (n = 2^L = 2^K1 * 2^Kc, N = n * 2^r), all in natural order at the interface:
//   pass A  (strided)    : top K1 inverse DIF stages on [2^K1 rows x 8 adjacent columns] tiles in
//                          shared memory + the four-step twiddle w^-(c * k1)
//   fused   (contiguous) : last Kc inverse DIF stages on a 2^Kc block, scale by 7^i / n, replicate
//                          x 2^r (the first r DIT stages of a zero-padded input are copies), the
//                          first Kc forward DIT stages on the 2^(Kc+r) block, four-step twiddle
//   pass D  (strided)    : last K1 forward DIT stages on [2^K1 rows x 8 columns] tiles
// The coefficient vector only ever exists in bit-reversed order inside the pipeline; it is never
// written out (openings and the FRI combination are computed from evaluations instead).
// Algorithmic bytes per column: 8 n (2 + 2^r)  (SURVEY.md 8d); bound: HBM.
#pragma once
#include "ntt_tables.cuh"

namespace ntt {

enum Mode { FROM_VALUES_LDE = 0, FROM_COEFFS_LDE = 1, INTT_COSET_NAT = 2 };

struct Plan {
  int L, r, Kc, K1;
};
static inline Plan make_plan(int L, int r) {
  Plan p;
  p.L = L;
  p.r = r;
  p.Kc = L < KC_MAX ? L : KC_MAX;
  if (L - p.Kc > K1_MAX) p.Kc = L - K1_MAX;
  p.K1 = L - p.Kc;
  if (p.Kc + r > 13 || L + r > LOG_M) throw Pb254Error(6, "ntt: size not supported");
  return p;
}

#if !PB_HOSTSIM
// ---------------------------------------------------------------------------------------------
// pass A: inverse DIF over the top K1 index bits; rows r = 0..2^K1-1 at stride C = 2^(L-K1).
static __global__ void __launch_bounds__(256) k_ntt_pass_a(const u64* __restrict__ in, u64* __restrict__ out, int L, int K1,
                                                    size_t in_stride, size_t out_stride, Tables t) {
  extern __shared__ u64 sm[];
  const int R = 1 << K1;
  const size_t C = (size_t)1 << (L - K1);
  u64* tw = sm + (size_t)R * TC;  // R/2 inverse twiddles of order R
  const int tid = threadIdx.x, nt = blockDim.x;
  const size_t c0 = (size_t)blockIdx.x * TC;
  const u64* src = in + (size_t)blockIdx.y * in_stride;
  u64* dst = out + (size_t)blockIdx.y * out_stride;
  for (int e = tid; e < R / 2; e += nt) tw[e] = tpow(t.inv_lo, t.inv_hi, (u64)e << (LOG_M - K1));
  for (int idx = tid; idx < R * TC; idx += nt) {
    int rr = idx / TC, c = idx % TC;
    sm[idx] = src[(size_t)rr * C + c0 + c];
  }
  __syncthreads();
  for (int s = 0; s < K1; s++) {
    const int half = R >> (s + 1);
    for (int b = tid; b < (R / 2) * TC; b += nt) {
      int c = b % TC, p = b / TC;
      int j = p & (half - 1), grp = p / half;
      int i0 = (grp * 2 * half + j) * TC + c, i1 = i0 + half * TC;
      u64 a = sm[i0], bb = sm[i1];
      sm[i0] = gl::add(a, bb);
      sm[i1] = gl::mul(gl::sub(a, bb), tw[j << s]);
    }
    __syncthreads();
  }
  for (int idx = tid; idx < R * TC; idx += nt) {
    int rr = idx / TC, c = idx % TC;
    u64 k1 = gl::brev32((u32)rr, K1);
    u64 e = (c0 + c) * k1;  // < 2^L
    dst[(size_t)rr * C + c0 + c] = gl::mul(sm[idx], tpow(t.inv_lo, t.inv_hi, e << (LOG_M - L)));
  }
}

// fused contiguous kernel, one 2^Kc block of one column per CTA
static __global__ void __launch_bounds__(256) k_ntt_fused(const u64* __restrict__ in, u64* __restrict__ out, int L, int Kc,
                                                   int r, size_t in_stride, size_t out_stride, Tables t, u64 ninv,
                                                   int mode) {
  extern __shared__ u64 sm[];
  const int C = 1 << Kc, Cp = C << r, K1 = L - Kc;
  u64* small = sm;                 // C
  u64* big = sm + C;               // Cp
  u64* twf = big + Cp;             // Cp/2 forward twiddles of order Cp
  u64* twi = twf + Cp / 2;         // C/2 inverse twiddles of order C
  const int tid = threadIdx.x, nt = blockDim.x;
  const size_t blk = blockIdx.x;
  const u64* src = in + (size_t)blockIdx.y * in_stride;
  u64* dst = out + (size_t)blockIdx.y * out_stride;
  if (mode != FROM_COEFFS_LDE)
    for (int e = tid; e < C / 2; e += nt) twi[e] = tpow(t.inv_lo, t.inv_hi, (u64)e << (LOG_M - Kc));
  if (mode != INTT_COSET_NAT)
    for (int e = tid; e < Cp / 2; e += nt) twf[e] = tpow(t.fwd_lo, t.fwd_hi, (u64)e << (LOG_M - Kc - r));
  if (mode == FROM_COEFFS_LDE) {
    for (int off = tid; off < C; off += nt) {
      u64 m = blk * C + off;
      small[off] = src[gl::brev32((u32)m, L)];
    }
  } else {
    for (int off = tid; off < C; off += nt) small[off] = src[blk * C + off];
  }
  __syncthreads();
  if (mode != FROM_COEFFS_LDE) {
    for (int s = 0; s < Kc; s++) {
      const int half = C >> (s + 1);
      for (int b = tid; b < C / 2; b += nt) {
        int j = b & (half - 1), grp = b / half;
        int i0 = grp * 2 * half + j, i1 = i0 + half;
        u64 a = small[i0], bb = small[i1];
        small[i0] = gl::add(a, bb);
        small[i1] = gl::mul(gl::sub(a, bb), twi[j << s]);
      }
      __syncthreads();
    }
  }
  if (mode == INTT_COSET_NAT) {
    for (int off = tid; off < C; off += nt) {
      u64 m = blk * C + off;
      u64 i = gl::brev32((u32)m, L);
      u64 sc = gl::mul(tpow(t.ish_lo, t.ish_hi, i), ninv);
      dst[i] = gl::mul(small[off], sc);
    }
    return;
  }
  for (int off = tid; off < C; off += nt) {
    u64 m = blk * C + off;
    u64 i = gl::brev32((u32)m, L);
    u64 sc = tpow(t.sh_lo, t.sh_hi, i);
    if (mode == FROM_VALUES_LDE) sc = gl::mul(sc, ninv);
    u64 v = gl::mul(small[off], sc);
    for (int q = 0; q < (1 << r); q++) big[(off << r) + q] = v;
  }
  __syncthreads();
  for (int s = r; s < Kc + r; s++) {
    const int half = 1 << s;
    for (int b = tid; b < Cp / 2; b += nt) {
      int j = b & (half - 1), grp = b >> s;
      int i0 = grp * 2 * half + j, i1 = i0 + half;
      u64 a = big[i0], x = gl::mul(big[i1], twf[j << (Kc + r - 1 - s)]);
      big[i0] = gl::add(a, x);
      big[i1] = gl::sub(a, x);
    }
    __syncthreads();
  }
  const u64 i1 = gl::brev32((u32)blk, K1);
  for (int off = tid; off < Cp; off += nt) {
    u64 v = big[off];
    if (K1 > 0) v = gl::mul(v, tpow(t.fwd_lo, t.fwd_hi, (i1 * off) << (LOG_M - L - r)));
    dst[blk * Cp + off] = v;
  }
}

// pass D: forward DIT over the block index (rows at stride Cp), in place
static __global__ void __launch_bounds__(256) k_ntt_pass_d(u64* __restrict__ data, int K1, size_t Cp, size_t stride,
                                                    Tables t) {
  extern __shared__ u64 sm[];
  const int R = 1 << K1;
  u64* tw = sm + (size_t)R * TC;
  const int tid = threadIdx.x, nt = blockDim.x;
  const size_t c0 = (size_t)blockIdx.x * TC;
  u64* col = data + (size_t)blockIdx.y * stride;
  for (int e = tid; e < R / 2; e += nt) tw[e] = tpow(t.fwd_lo, t.fwd_hi, (u64)e << (LOG_M - K1));
  for (int idx = tid; idx < R * TC; idx += nt) {
    int rr = idx / TC, c = idx % TC;
    sm[idx] = col[(size_t)rr * Cp + c0 + c];
  }
  __syncthreads();
  for (int s = 0; s < K1; s++) {
    const int half = 1 << s;
    for (int b = tid; b < (R / 2) * TC; b += nt) {
      int c = b % TC, p = b / TC;
      int j = p & (half - 1), grp = p >> s;
      int i0 = (grp * 2 * half + j) * TC + c, i1 = i0 + half * TC;
      u64 a = sm[i0], x = gl::mul(sm[i1], tw[j << (K1 - 1 - s)]);
      sm[i0] = gl::add(a, x);
      sm[i1] = gl::sub(a, x);
    }
    __syncthreads();
  }
  for (int idx = tid; idx < R * TC; idx += nt) {
    int rr = idx / TC, c = idx % TC;
    col[(size_t)rr * Cp + c0 + c] = sm[idx];
  }
}

static bool g_ntt_attr_set = false;
static inline void set_smem_attrs() {
  if (g_ntt_attr_set) return;
  PB_CUDA(cudaFuncSetAttribute(k_ntt_pass_a, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  PB_CUDA(cudaFuncSetAttribute(k_ntt_pass_d, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  PB_CUDA(cudaFuncSetAttribute(k_ntt_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  g_ntt_attr_set = true;
}
#else
// ---- hostsim stand-ins: the same mathematical transforms with a plain radix-2 NTT -------------
static inline void hs_ntt(std::vector<u64>& a, u64 w) {
  size_t n = a.size();
  unsigned lg = 0;
  while (((size_t)1 << lg) < n) lg++;
  for (size_t i = 0; i < n; i++) {
    size_t j = gl::brev32((u32)i, lg);
    if (i < j) std::swap(a[i], a[j]);
  }
  for (size_t m = 2; m <= n; m <<= 1) {
    u64 wm = gl::pow(w, n / m);
    for (size_t k = 0; k < n; k += m) {
      u64 x = 1;
      for (size_t j = 0; j < m / 2; j++) {
        u64 t = gl::mul(x, a[k + j + m / 2]), u = a[k + j];
        a[k + j] = gl::add(u, t);
        a[k + j + m / 2] = gl::sub(u, t);
        x = gl::mul(x, wm);
      }
    }
  }
}
#endif

// LDE of `ncols` columns. in: [col][in_stride] (n used), out: [col][out_stride] (N = n << r used).
// scratch: at least ncols * n words (unused when K1 == 0 or in FROM_COEFFS mode).
static inline void lde_columns(const TableSet& ts, const u64* in, size_t in_stride, u64* out, size_t out_stride,
                               u64* scratch, int ncols, int L, int r, int mode, pbStream s) {
  if (ncols == 0) return;
  const size_t n = (size_t)1 << L;
#if PB_HOSTSIM
  (void)ts;
  (void)scratch;
  (void)s;
  const size_t N = n << r;
#pragma omp parallel for schedule(dynamic, 1)
  for (int c = 0; c < ncols; c++) {
    std::vector<u64> v(in + (size_t)c * in_stride, in + (size_t)c * in_stride + n);
    if (mode == FROM_VALUES_LDE) {
      hs_ntt(v, gl::inv(gl::root_of_unity(L)));
      u64 ninv = gl::inv((u64)n);
      for (auto& x : v) x = gl::mul(x, ninv);
    }
    u64 p = 1;
    for (auto& x : v) {
      x = gl::mul(x, p);
      p = gl::mul(p, gl::COSET_SHIFT);
    }
    v.resize(N, 0);
    hs_ntt(v, gl::root_of_unity(L + r));
    memcpy(out + (size_t)c * out_stride, v.data(), N * 8);
  }
#else
  set_smem_attrs();
  Plan p = make_plan(L, r);
  const u64 ninv = gl::inv((u64)n % gl::P);
  const int C = 1 << p.Kc, Cp = C << r;
  const u64* fused_in = in;
  size_t fused_stride = in_stride;
  if (mode == FROM_VALUES_LDE && p.K1 > 0) {
    dim3 grid((unsigned)((n >> p.K1) / TC), ncols);
    size_t smem = ((size_t)(1 << p.K1) * TC + (1 << p.K1) / 2) * 8;
    k_ntt_pass_a<<<grid, 256, smem, s>>>(in, scratch, L, p.K1, in_stride, n, ts.t);
    g_pb_launches++;
    pb_check_last("ntt pass A");
    fused_in = scratch;
    fused_stride = n;
  }
  {
    dim3 grid((unsigned)(n >> p.Kc), ncols);
    size_t smem = ((size_t)C + Cp + Cp / 2 + C / 2) * 8;
    k_ntt_fused<<<grid, 256, smem, s>>>(fused_in, out, L, p.Kc, r, fused_stride, out_stride, ts.t, ninv, mode);
    g_pb_launches++;
    pb_check_last("ntt fused");
  }
  if (p.K1 > 0) {
    dim3 grid((unsigned)(Cp / TC), ncols);
    size_t smem = ((size_t)(1 << p.K1) * TC + (1 << p.K1) / 2) * 8;
    k_ntt_pass_d<<<grid, 256, smem, s>>>(out, p.K1, (size_t)Cp, out_stride, ts.t);
    g_pb_launches++;
    pb_check_last("ntt pass D");
  }
#endif
}

// coset iNTT (shift 7) of `ncols` columns of length n = 2^L: natural-order values on 7*<w> ->
// natural-order coefficients. out may not alias in. scratch: ncols * n words.
static inline void coset_intt_columns(const TableSet& ts, const u64* in, size_t in_stride, u64* out, size_t out_stride,
                                      u64* scratch, int ncols, int L, pbStream s) {
  const size_t n = (size_t)1 << L;
#if PB_HOSTSIM
  (void)ts;
  (void)scratch;
  (void)s;
  for (int c = 0; c < ncols; c++) {
    std::vector<u64> v(in + (size_t)c * in_stride, in + (size_t)c * in_stride + n);
    hs_ntt(v, gl::inv(gl::root_of_unity(L)));
    u64 sc = gl::inv((u64)n), si = gl::inv(gl::COSET_SHIFT);
    for (auto& x : v) {
      x = gl::mul(x, sc);
      sc = gl::mul(sc, si);
    }
    memcpy(out + (size_t)c * out_stride, v.data(), n * 8);
  }
#else
  set_smem_attrs();
  Plan p = make_plan(L, 0);
  const u64 ninv = gl::inv((u64)n % gl::P);
  const u64* fused_in = in;
  size_t fused_stride = in_stride;
  if (p.K1 > 0) {
    dim3 grid((unsigned)((n >> p.K1) / TC), ncols);
    size_t smem = ((size_t)(1 << p.K1) * TC + (1 << p.K1) / 2) * 8;
    k_ntt_pass_a<<<grid, 256, smem, s>>>(in, scratch, L, p.K1, in_stride, n, ts.t);
    g_pb_launches++;
    pb_check_last("intt pass A");
    fused_in = scratch;
    fused_stride = n;
  }
  dim3 grid((unsigned)(n >> p.Kc), ncols);
  const int C = 1 << p.Kc;
  size_t smem = ((size_t)C + C + C / 2 + C / 2) * 8;
  k_ntt_fused<<<grid, 256, smem, s>>>(fused_in, out, L, p.Kc, 0, fused_stride, out_stride, ts.t, ninv, INTT_COSET_NAT);
  g_pb_launches++;
  pb_check_last("intt fused");
#endif
}

}  // namespace ntt
