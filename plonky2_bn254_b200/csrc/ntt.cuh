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
// Split of the L index bits between the contiguous (fused) kernel and the strided passes. Measured on B200
// (tools/probe_commit.py, PB254_KC sweep): a 2^9 fused block with 2^10-row strided tiles is best at 2^19,
// and strided tiles beyond 2^10 rows (> 64 KB of shared memory, one CTA per SM) lose more than larger
// fused blocks cost; so Kc = max(9, L - 10). PB254_KC overrides it for experiments.
static inline Plan make_plan(int L, int r) {
  Plan p;
  p.L = L;
  p.r = r;
  static int forced = -1;
  if (forced < 0) {
    const char* e = getenv("PB254_KC");
    forced = e ? atoi(e) : 0;
  }
  p.Kc = forced > 0 ? forced : (L - 10 > 9 ? L - 10 : 9);
  if (p.Kc > L) p.Kc = L;
  if (p.Kc + r > 13) p.Kc = 13 - r;
  // the fused kernel keeps 2^Kc + 2^(Kc+r) padded words and both twiddle tables in shared memory (<= 200 KB)
  while (p.Kc > 3 && ((size_t)(17 << p.Kc) / 16 + (size_t)(17 << (p.Kc + r)) / 16 + 32 + (3 << (p.Kc + r)) / 2) * 8 > 200 * 1024)
    p.Kc--;
  if (L - p.Kc > K1_MAX) p.Kc = L - K1_MAX;
  p.K1 = L - p.Kc;
  if (p.Kc + r > 13 || L + r > LOG_M ||
      ((size_t)(17 << p.Kc) / 16 + (size_t)(17 << (p.Kc + r)) / 16 + 32 + (3 << (p.Kc + r)) / 2) * 8 > 200 * 1024)
    throw Pb254Error(6, "ntt: size not supported");
  return p;
}

#if !PB_HOSTSIM
// ---------------------------------------------------------------------------------------------
// Shared-memory tiles hold [row][column] with 2^LOGTC columns; one work item is a radix-Q butterfly
// (Q = 8, 4 or 2: three, two or one radix-2 stages) on Q rows of one column, done in registers with
// lazy u64 representatives (gl::mul_lazy / add_lazy / sub_lazy), so a 2^K-row transform costs
// ceil(K / 3) shared-memory round trips and barriers instead of K. Logical word i lives at
// i + (i >> 4): one pad word per 16 breaks the power-of-two strides of the last rounds (the Q rows of
// an item are then adjacent), which would otherwise be 16-way bank conflicts.
// Twiddles of order 2^k, k <= 13, come straight from the power tables: fwd_hi[e << (13 - k)].
#define NTT_PHYS(i) ((i) + ((i) >> 4))

template <int Q, bool DIF, int LOGTC>
__device__ __forceinline__ void tile_round(u64* __restrict__ sm, const u64* __restrict__ tw, int K, int s, int tid, int nt) {
  constexpr int LQ = Q == 8 ? 3 : Q == 4 ? 2 : 1;
  const int items = ((1 << K) >> LQ) << LOGTC;
  for (int t = tid; t < items; t += nt) {
    const int c = t & ((1 << LOGTC) - 1), u = t >> LOGTC;
    int base, step, j;
    if (DIF) {
      const int lstep = K - s - LQ;  // span = 2^(K - s) rows, the item's rows are span / Q apart
      step = 1 << lstep;
      j = u & (step - 1);
      base = ((u >> lstep) << (K - s)) + j;
    } else {
      step = 1 << s;
      j = u & (step - 1);
      base = ((u >> s) << (s + LQ)) + j;
    }
    u64 x[Q];
#pragma unroll
    for (int m = 0; m < Q; m++) x[m] = sm[NTT_PHYS(((base + m * step) << LOGTC) + c)];
#pragma unroll
    for (int k = 0; k < LQ; k++) {
      const int hb = DIF ? (Q >> (k + 1)) : (1 << k);
#pragma unroll
      for (int m = 0; m < Q; m++) {
        if (m & hb) continue;
        const int pos = m & (hb - 1);
        const u64 a = x[m], b2 = x[m + hb];
        if (DIF) {
          const u64 w = tw[(j + pos * step) << (s + k)];
          x[m] = gl::add_lazy(a, b2);
          x[m + hb] = gl::mul_lazy(gl::sub_lazy(a, b2), w);
        } else {
          const u64 w = tw[(j + pos * step) << (K - 1 - (s + k))];
          const u64 y = gl::mul_lazy(b2, w);
          x[m] = gl::add_lazy(a, y);
          x[m + hb] = gl::sub_lazy(a, y);
        }
      }
    }
#pragma unroll
    for (int m = 0; m < Q; m++) sm[NTT_PHYS(((base + m * step) << LOGTC) + c)] = x[m];
  }
}

// stages [s0, s1) of a 2^K-row transform on the tile, a barrier after every round
template <bool DIF, int LOGTC>
__device__ __forceinline__ void tile_stages(u64* sm, const u64* tw, int K, int s0, int s1, int tid, int nt) {
  int s = s0;
  while (s1 - s >= 3) {
    tile_round<8, DIF, LOGTC>(sm, tw, K, s, tid, nt);
    __syncthreads();
    s += 3;
  }
  if (s1 - s == 2) {
    tile_round<4, DIF, LOGTC>(sm, tw, K, s, tid, nt);
    __syncthreads();
  } else if (s1 - s == 1) {
    tile_round<2, DIF, LOGTC>(sm, tw, K, s, tid, nt);
    __syncthreads();
  }
}

static constexpr int LOG_TC = 3;  // TC = 8
static_assert((1 << LOG_TC) == TC, "tile width");
static inline size_t tile_smem_words(int K1) {
  size_t w = (size_t)(1 << K1) * TC;
  return w + (w >> 4) + 16 + (1 << K1) / 2;
}

// pass A: inverse DIF over the top K1 index bits; rows r = 0..2^K1-1 at stride C = 2^(L-K1).
static __global__ void __launch_bounds__(512) k_ntt_pass_a(const u64* __restrict__ in, u64* __restrict__ out, int L, int K1,
                                                    size_t in_stride, size_t out_stride, Tables t) {
  extern __shared__ u64 sm[];
  const int R = 1 << K1;
  const size_t C = (size_t)1 << (L - K1);
  u64* tw = sm + NTT_PHYS(R * TC) + 16;  // R/2 inverse twiddles of order R
  const int tid = threadIdx.x, nt = blockDim.x;
  const size_t c0 = (size_t)blockIdx.x * TC;
  const u64* src = in + (size_t)blockIdx.y * in_stride;
  u64* dst = out + (size_t)blockIdx.y * out_stride;
  for (int e = tid; e < R / 2; e += nt) tw[e] = t.inv_hi[(size_t)e << (LOG_T - K1)];
  for (int idx = tid; idx < R * TC; idx += nt) {
    int rr = idx >> LOG_TC, c = idx & (TC - 1);
    sm[NTT_PHYS(idx)] = src[(size_t)rr * C + c0 + c];
  }
  __syncthreads();
  tile_stages<true, LOG_TC>(sm, tw, K1, 0, K1, tid, nt);
  for (int idx = tid; idx < R * TC; idx += nt) {
    int rr = idx >> LOG_TC, c = idx & (TC - 1);
    u64 k1 = gl::brev32((u32)rr, K1);
    u64 e = (c0 + c) * k1;  // < 2^L
    // intermediate buffer: any u64 representative is fine for the next kernel
    dst[(size_t)rr * C + c0 + c] = gl::mul_lazy(sm[NTT_PHYS(idx)], tpow(t.inv_lo, t.inv_hi, e << (LOG_M - L)));
  }
}

// fused contiguous kernel, one 2^Kc block of one column per CTA
static __global__ void __launch_bounds__(512) k_ntt_fused(const u64* __restrict__ in, u64* __restrict__ out, int L, int Kc,
                                                   int r, size_t in_stride, size_t out_stride, Tables t, u64 ninv,
                                                   int mode) {
  extern __shared__ u64 sm[];
  const int C = 1 << Kc, Cp = C << r, K1 = L - Kc;
  u64* small = sm;                              // C (padded)
  u64* big = small + NTT_PHYS(C) + 16;          // Cp (padded)
  u64* twf = big + NTT_PHYS(Cp) + 16;           // Cp/2 forward twiddles of order Cp
  u64* twi = twf + Cp / 2;                      // C/2 inverse twiddles of order C
  const int tid = threadIdx.x, nt = blockDim.x;
  const size_t blk = blockIdx.x;
  const u64* src = in + (size_t)blockIdx.y * in_stride;
  u64* dst = out + (size_t)blockIdx.y * out_stride;
  if (mode != FROM_COEFFS_LDE)
    for (int e = tid; e < C / 2; e += nt) twi[e] = t.inv_hi[(size_t)e << (LOG_T - Kc)];
  if (mode != INTT_COSET_NAT)
    for (int e = tid; e < Cp / 2; e += nt) twf[e] = t.fwd_hi[(size_t)e << (LOG_T - Kc - r)];
  if (mode == FROM_COEFFS_LDE) {
    for (int off = tid; off < C; off += nt) {
      u64 m = blk * C + off;
      small[NTT_PHYS(off)] = src[gl::brev32((u32)m, L)];
    }
  } else {
    for (int off = tid; off < C; off += nt) small[NTT_PHYS(off)] = src[blk * C + off];
  }
  __syncthreads();
  if (mode != FROM_COEFFS_LDE) tile_stages<true, 0>(small, twi, Kc, 0, Kc, tid, nt);
  if (mode == INTT_COSET_NAT) {
    for (int off = tid; off < C; off += nt) {
      u64 m = blk * C + off;
      u64 i = gl::brev32((u32)m, L);
      u64 sc = gl::mul_lazy(tpow(t.ish_lo, t.ish_hi, i), ninv);
      dst[i] = gl::mul(small[NTT_PHYS(off)], sc);
    }
    return;
  }
  for (int off = tid; off < C; off += nt) {
    u64 m = blk * C + off;
    u64 i = gl::brev32((u32)m, L);
    u64 sc = tpow(t.sh_lo, t.sh_hi, i);
    if (mode == FROM_VALUES_LDE) sc = gl::mul_lazy(sc, ninv);
    u64 v = gl::mul_lazy(small[NTT_PHYS(off)], sc);
    for (int q = 0; q < (1 << r); q++) big[NTT_PHYS((off << r) + q)] = v;
  }
  __syncthreads();
  tile_stages<false, 0>(big, twf, Kc + r, r, Kc + r, tid, nt);
  const u64 i1 = gl::brev32((u32)blk, K1);
  for (int off = tid; off < Cp; off += nt) {
    u64 v = big[NTT_PHYS(off)];
    if (K1 > 0)
      v = gl::mul_lazy(v, tpow(t.fwd_lo, t.fwd_hi, (i1 * off) << (LOG_M - L - r)));  // pass D canonicalises
    else
      v = gl::canonical(v);
    dst[blk * Cp + off] = v;
  }
}

// pass D: forward DIT over the block index (rows at stride Cp), in place; writes canonical values
static __global__ void __launch_bounds__(512) k_ntt_pass_d(u64* __restrict__ data, int K1, size_t Cp, size_t stride,
                                                    Tables t) {
  extern __shared__ u64 sm[];
  const int R = 1 << K1;
  u64* tw = sm + NTT_PHYS(R * TC) + 16;
  const int tid = threadIdx.x, nt = blockDim.x;
  const size_t c0 = (size_t)blockIdx.x * TC;
  u64* col = data + (size_t)blockIdx.y * stride;
  for (int e = tid; e < R / 2; e += nt) tw[e] = t.fwd_hi[(size_t)e << (LOG_T - K1)];
  for (int idx = tid; idx < R * TC; idx += nt) {
    int rr = idx >> LOG_TC, c = idx & (TC - 1);
    sm[NTT_PHYS(idx)] = col[(size_t)rr * Cp + c0 + c];
  }
  __syncthreads();
  tile_stages<false, LOG_TC>(sm, tw, K1, 0, K1, tid, nt);
  for (int idx = tid; idx < R * TC; idx += nt) {
    int rr = idx >> LOG_TC, c = idx & (TC - 1);
    col[(size_t)rr * Cp + c0 + c] = gl::canonical(sm[NTT_PHYS(idx)]);
  }
}

static inline size_t fused_smem_words(int C, int Cp) {
  return (size_t)NTT_PHYS(C) + 16 + NTT_PHYS(Cp) + 16 + Cp / 2 + C / 2;
}
// one radix-8 item per thread and round where the tile is large enough
static inline unsigned tile_threads(int K1) {
  int items = ((1 << K1) * TC) >> 3;
  return items >= 512 ? 512u : items >= 64 ? (unsigned)items : 64u;
}
static inline unsigned fused_threads(int Cp) {  // one radix-8 item per thread in the forward rounds
  int t = Cp >> 3;
  return t >= 512 ? 512u : t >= 64 ? (unsigned)t : 64u;
}
// cudaFuncSetAttribute is per device: remember which devices of this process have been configured
// (contexts on different devices are created and used from different threads: the table is guarded)
static bool g_ntt_attr_set[64] = {};
static std::mutex g_ntt_attr_mutex;
static inline void set_smem_attrs() {
  int dev = 0;
  PB_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) dev = 63;
  std::lock_guard<std::mutex> lock(g_ntt_attr_mutex);
  if (g_ntt_attr_set[dev]) return;
  PB_CUDA(cudaFuncSetAttribute(k_ntt_pass_a, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  PB_CUDA(cudaFuncSetAttribute(k_ntt_pass_d, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  PB_CUDA(cudaFuncSetAttribute(k_ntt_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  g_ntt_attr_set[dev] = true;
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
    size_t smem = tile_smem_words(p.K1) * 8;
    k_ntt_pass_a<<<grid, tile_threads(p.K1), smem, s>>>(in, scratch, L, p.K1, in_stride, n, ts.t);
    g_pb_launches++;
    pb_check_last("ntt pass A");
    fused_in = scratch;
    fused_stride = n;
  }
  {
    dim3 grid((unsigned)(n >> p.Kc), ncols);
    size_t smem = fused_smem_words(C, Cp) * 8;
    k_ntt_fused<<<grid, fused_threads(Cp), smem, s>>>(fused_in, out, L, p.Kc, r, fused_stride, out_stride, ts.t, ninv, mode);
    g_pb_launches++;
    pb_check_last("ntt fused");
  }
  if (p.K1 > 0) {
    dim3 grid((unsigned)(Cp / TC), ncols);
    size_t smem = tile_smem_words(p.K1) * 8;
    k_ntt_pass_d<<<grid, tile_threads(p.K1), smem, s>>>(out, p.K1, (size_t)Cp, out_stride, ts.t);
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
    size_t smem = tile_smem_words(p.K1) * 8;
    k_ntt_pass_a<<<grid, tile_threads(p.K1), smem, s>>>(in, scratch, L, p.K1, in_stride, n, ts.t);
    g_pb_launches++;
    pb_check_last("intt pass A");
    fused_in = scratch;
    fused_stride = n;
  }
  dim3 grid((unsigned)(n >> p.Kc), ncols);
  const int C = 1 << p.Kc;
  size_t smem = fused_smem_words(C, C) * 8;
  k_ntt_fused<<<grid, fused_threads(C), smem, s>>>(fused_in, out, L, p.Kc, 0, fused_stride, out_stride, ts.t, ninv, INTT_COSET_NAT);
  g_pb_launches++;
  pb_check_last("intt fused");
#endif
}

}  // namespace ntt
