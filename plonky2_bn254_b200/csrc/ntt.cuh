// K3 `lde_batch`: Goldilocks low-degree extension of a column-major matrix.
// Replaces PolynomialBatch::from_values / from_coeffs of plonky2 0.2.2 fri/oracle.rs (un-vendored;
// call site src/starks/common/prover.rs:31-38): per column  iNTT(n) -> * 7^i -> zero-pad ->
// NTT(n * 2^r), natural-order output (value j = P(7 * w^j)).
//
// Decomposition (n = 2^L = 2^K1 * 2^Kc, N = n * 2^r), all in natural order at the interface:
//   pass A  (strided)    : top K1 inverse DIF stages on [2^K1 rows x 16 adjacent columns] tiles in shared
//                          memory; the four-step twiddle w^-(c * k1) rides on the last round
//   fused   (contiguous) : last Kc inverse DIF stages on a 2^Kc block; the scale 7^i / n rides on the last
//                          round, which also replicates x 2^r (the first r DIT stages of a zero-padded
//                          input are copies); then the first Kc forward DIT stages on the 2^(Kc+r) block
//   pass D  (strided)    : last K1 forward DIT stages on [2^K1 rows x 16 columns] tiles, the four-step
//                          twiddle rides on its first round; writes canonical values
// The coefficient vector only ever exists in bit-reversed order inside the pipeline; it is never
// written out (openings and the FRI combination are computed from evaluations instead).
//
// Round-2 kernels. A work item is a radix-16 (or 8 / 4 / 2) DFT of one thread, done in registers with the
// multiplication-free butterflies of f96.cuh (2 is a 192-th root of unity in Goldilocks: every internal
// twiddle of a radix <= 16 DFT is a power of two), so a round of four stages costs one shared-memory
// round trip, one generic multiplication per element (the twiddle between rounds, read from a table laid
// out [digit][position] so that a warp reads consecutive words) and ~37 shift/add instructions per
// element. Every multiplication that is not a butterfly twiddle (four-step twiddles, 7^i / n) comes from a
// per-size table in global memory that all columns share (it stays in L2) and replaces a round's
// twiddle, so each element is multiplied exactly once per round. Strided tiles are fetched by the bulk-
// copy engine (cp.async.bulk + mbarrier, one 128-byte row segment per thread), the contiguous block by
// 128-bit loads; the in-tile twiddle tables are staged into shared memory by one bulk copy per CTA.
// Algorithmic bytes per column: 8 n (2 + 2^r)  (SURVEY.md 8d).
#pragma once
#include "ntt_tables.cuh"
#include "f96.cuh"
#include <map>

namespace ntt {

enum Mode { FROM_VALUES_LDE = 0, FROM_COEFFS_LDE = 1, INTT_COSET_NAT = 2 };

PB_HD u64 mulz(u64 a, u64 b) {  // any representatives in, any representative out
#if defined(__CUDA_ARCH__)
  return gl::mul_lazy(a, b);
#else
  return gl::mul(a, b);
#endif
}
PB_HD u64 canon(u64 a) { return a >= gl::P ? a - gl::P : a; }

// ---- plan ------------------------------------------------------------------------------------------
// A transform part is a list of rounds; round i does lq[i] radix-2 stages in one radix-2^lq DFT per item, the
// items' elements are 2^lstep[i] slots apart; off[i] is the offset of its twiddle table (words) inside the
// part's table, laid out [digit - 1][position], position < 2^lstep (no table when lstep == 0).
static constexpr int MAX_ROUNDS = 3;
struct Rounds {
  int n, words;
  int lq[MAX_ROUNDS], lstep[MAX_ROUNDS], off[MAX_ROUNDS];
};
static inline void split_stages(int ns, int* sizes, int& nr) {
  nr = (ns + 3) / 4;
  for (int i = 0; i < nr; i++) sizes[i] = ns / nr + (i < ns % nr ? 1 : 0);
}
// inverse DIF over K index bits: natural order in, bit-reversed out
static inline Rounds rounds_dif(int K) {
  Rounds R = {};
  int sz[8];
  split_stages(K, sz, R.n);
  if (R.n > MAX_ROUNDS) throw Pb254Error(6, "ntt: size not supported");
  int s = 0;
  for (int i = 0; i < R.n; i++) {
    R.lq[i] = sz[i];
    R.lstep[i] = K - s - sz[i];
    R.off[i] = R.words;
    if (R.lstep[i] > 0) R.words += ((1 << sz[i]) - 1) << R.lstep[i];
    s += sz[i];
  }
  return R;
}
// forward DIT stages [s0, K): bit-reversed in, natural out
static inline Rounds rounds_dit(int K, int s0) {
  Rounds R = {};
  int sz[8];
  split_stages(K - s0, sz, R.n);
  if (R.n > MAX_ROUNDS) throw Pb254Error(6, "ntt: size not supported");
  int s = s0;
  for (int i = 0; i < R.n; i++) {
    R.lq[i] = sz[i];
    R.lstep[i] = s;
    R.off[i] = R.words;
    if (s > 0) R.words += ((1 << sz[i]) - 1) << s;
    s += sz[i];
  }
  return R;
}

#define NTT_PHYS(i) ((i) + ((i) >> 4))  // one pad word per 16: the stride-16 accesses of a step-1 round spread over banks

struct Plan {
  int L = 0, r = 0, Kc = 0, K1 = 0, logtc = 4;
  Rounds a, fi, ff, d;  // pass A, fused inverse, fused forward, pass D
  u64* dev = nullptr;   // one allocation: in-tile twiddles, four-step tables
  const u64 *tw_a = nullptr, *tw_f = nullptr, *tw_d = nullptr;
  const u64 *tw4_a = nullptr, *tw4_d = nullptr;
  u64* sc[3] = {nullptr, nullptr, nullptr};  // per mode, n words, block (bit-reversed) order, allocated on first use
  size_t fused_smem_words() const {
    const size_t C = (size_t)1 << Kc, Cp = C << r;
    return ((NTT_PHYS(C) + 17) & ~(size_t)1) + ((NTT_PHYS(Cp) + 17) & ~(size_t)1) + (size_t)fi.words + ff.words + 2;
  }
  size_t tile_smem_words() const { return ((size_t)1 << (K1 + logtc)) + a.words + d.words + 2; }
};

static inline bool k1_ok(int K1) { return K1 == 0 || K1 == 3 || K1 == 4 || (K1 >= 6 && K1 <= K1_MAX); }

// Split of the L index bits between the contiguous (fused) kernel and the strided passes: the fewest rounds
// (an inverse round touches n elements, a forward round n 2^r), 16-column strided tiles (K1 <= 10) and a fused
// block that leaves room for two CTAs per SM preferred. PB254_KC overrides Kc for experiments.
static inline void choose_split(int L, int r, int& Kc, int& K1) {
  static int forced = -1;
  if (forced < 0) {
    const char* e = getenv("PB254_KC");
    forced = e ? atoi(e) : 0;
  }
  long best = -1;
  for (int kc = (L < 12 ? L : 12); kc >= 0; kc--) {
    const int k1 = L - kc;
    if (!k1_ok(k1) || kc + r > LOG_T || L + r > LOG_M) continue;
    Plan p;
    p.Kc = kc;
    p.K1 = k1;
    p.r = r;
    p.fi = rounds_dif(kc);
    p.ff = rounds_dit(kc + r, r);
    if (p.fused_smem_words() * 8 > 200 * 1024) continue;
    const int ra = k1 ? (k1 + 3) / 4 : 0;
    long cost = (long)(p.fi.n + ra) * 16 + (long)((p.ff.n + ra) << r) * 16;
    if (k1 > 10) cost += 4;                                  // 8-column tiles
    if (p.fused_smem_words() * 8 > 110 * 1024) cost += 2;   // one CTA per SM
    if (forced > 0 && kc == forced) cost = 0;
    if (best < 0 || cost < best) {
      best = cost;
      Kc = kc;
      K1 = k1;
    }
  }
  if (best < 0) throw Pb254Error(6, "ntt: size not supported");
}

// ---- table generation (runs once per (L, r) and context) ---------------------------------------------
struct GenRoundTw {  // out[(k - 1) 2^lstep + j] = W^(+-(j k) 2^(26 - lstep - lq))
  u64* out;
  Tables t;
  int lq, lstep, inv;
  PB_HD void operator()(size_t i) const {
    const u64 j = i & (((size_t)1 << lstep) - 1), k = (i >> lstep) + 1;
    const u64 e = (j * k) << (LOG_M - lstep - lq);
    out[i] = canon(inv ? tpow(t.inv_lo, t.inv_hi, e) : tpow(t.fwd_lo, t.fwd_hi, e));
  }
};
struct GenFourStep {  // out[row 2^lc + c] = W^(+-(c brev_K1(row)) 2^(26 - bits))
  u64* out;
  Tables t;
  int K1, lc, bits, inv;
  PB_HD void operator()(size_t i) const {
    const u64 c = i & (((size_t)1 << lc) - 1), row = i >> lc;
    const u64 e = (c * gl::brev32((u32)row, K1)) << (LOG_M - bits);
    out[i] = canon(inv ? tpow(t.inv_lo, t.inv_hi, e) : tpow(t.fwd_lo, t.fwd_hi, e));
  }
};
struct GenScale {  // out[m] = 7^(+-brev_L(m)) * mult
  u64* out;
  Tables t;
  int L, inv;
  u64 mult;
  PB_HD void operator()(size_t m) const {
    const u64 i = gl::brev32((u32)m, L);
    const u64 s = inv ? tpow(t.ish_lo, t.ish_hi, i) : tpow(t.sh_lo, t.sh_hi, i);
    out[m] = gl::mul(canon(s), mult);
  }
};

static inline void gen_round_tables(const Tables& t, u64* dst, const Rounds& R, bool inv, pbStream s) {
  for (int i = 0; i < R.n; i++) {
    if (R.lstep[i] == 0) continue;
    GenRoundTw g = {dst + R.off[i], t, R.lq[i], R.lstep[i], inv ? 1 : 0};
    pb_launch("ntt tables", g, (size_t)((1 << R.lq[i]) - 1) << R.lstep[i], s);
  }
}

struct PlanCache {
  std::map<int, Plan> plans;
  Plan& get(const Tables& t, int L, int r, pbStream s) {
    const int key = L * 16 + r;
    auto it = plans.find(key);
    if (it != plans.end()) return it->second;
    Plan p;
    p.L = L;
    p.r = r;
    choose_split(L, r, p.Kc, p.K1);
    p.logtc = p.K1 > 10 ? 3 : 4;
    if (p.K1 && L - p.K1 < p.logtc) p.logtc = L - p.K1;  // never with L >= 16; keeps tiny sizes legal
    p.a = rounds_dif(p.K1);
    p.d = rounds_dit(p.K1, 0);
    p.fi = rounds_dif(p.Kc);
    p.ff = rounds_dit(p.Kc + r, r);
    const size_t n = (size_t)1 << L, N = n << r;
    auto up = [](size_t w) { return (w + 31) & ~(size_t)31; };
    const size_t o_a = 0, o_f = o_a + up(p.a.words), o_d = o_f + up(p.fi.words + p.ff.words), o_4a = o_d + up(p.d.words),
                 o_4d = o_4a + (p.K1 ? n : 0), total = o_4d + (p.K1 ? N : 0);
    p.dev = (u64*)pb_dev_alloc((total + 32) * 8);
    p.tw_a = p.dev + o_a;
    p.tw_f = p.dev + o_f;
    p.tw_d = p.dev + o_d;
    gen_round_tables(t, p.dev + o_a, p.a, true, s);
    gen_round_tables(t, p.dev + o_f, p.fi, true, s);
    gen_round_tables(t, p.dev + o_f + p.fi.words, p.ff, false, s);
    gen_round_tables(t, p.dev + o_d, p.d, false, s);
    if (p.K1) {
      p.tw4_a = p.dev + o_4a;
      p.tw4_d = p.dev + o_4d;
      GenFourStep ga = {p.dev + o_4a, t, p.K1, p.Kc, L, 1};
      pb_launch("ntt four-step inverse", ga, n, s);
      GenFourStep gd = {p.dev + o_4d, t, p.K1, p.Kc + r, L + r, 0};
      pb_launch("ntt four-step forward", gd, N, s);
    }
    return plans.emplace(key, p).first->second;
  }
  const u64* scale(Plan& p, const Tables& t, int mode, pbStream s) {
    if (p.sc[mode]) return p.sc[mode];
    const size_t n = (size_t)1 << p.L;
    p.sc[mode] = (u64*)pb_dev_alloc(n * 8);
    const u64 ninv = gl::inv((u64)n % gl::P);
    GenScale g = {p.sc[mode], t, p.L, mode == INTT_COSET_NAT ? 1 : 0, mode == FROM_COEFFS_LDE ? (u64)1 : ninv};
    pb_launch("ntt scale table", g, n, s);
    return p.sc[mode];
  }
  void destroy() {
    for (auto& kv : plans) {
      if (kv.second.dev) pb_dev_free(kv.second.dev);
      for (int m = 0; m < 3; m++)
        if (kv.second.sc[m]) pb_dev_free(kv.second.sc[m]);
    }
    plans.clear();
  }
};

static inline PlanCache& plan_cache(TableSet& ts) {
  if (!ts.plans) {
    ts.plans = new PlanCache;
    ts.plans_free = [](PlanCache* p) {
      p->destroy();
      delete p;
    };
  }
  return *ts.plans;
}

// ---- execution model glue ----------------------------------------------------------------------------
// The kernel bodies are written as phases over the CTA's threads so that the hostsim build can run them
// thread by thread between barriers (tests/test_hostsim_*: the real decomposition, tables and index maps are
// exercised without a GPU); nothing is carried in registers across a phase boundary.
#if PB_HOSTSIM
#define NTT_FOR_THREADS(nt) for (int tid = 0; tid < (int)(nt); tid++)
#define NTT_BARRIER() ((void)0)
#else
#define NTT_FOR_THREADS(nt) for (int tid = threadIdx.x, once_ = 1; once_; once_ = 0)
#define NTT_BARRIER() __syncthreads()
__device__ __forceinline__ u32 smem_u32(const void* p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(u64* bar, u32 count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(u64* bar, u32 bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// bulk-copy engine (TMA, 1-D): global -> shared, completion counted in bytes on the mbarrier
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, u32 bytes, u64* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(u64* bar, u32 parity) {
  u32 done = 0;
  while (!done)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done)
                 : "r"(smem_u32(bar)), "r"(parity)
                 : "memory");
}
#endif

template <int LQ>
PB_HD constexpr int brevq(int m) {
  int r = 0;
  for (int b = 0; b < LQ; b++) r |= ((m >> b) & 1) << (LQ - 1 - b);
  return r;
}

// One radix-2^LQ work item u of a round: its elements are the slots base + m 2^lstep. DIF (inverse): natural
// digit order in, DFT with the inverse root, element of output digit k multiplied by tw(k, ...) and stored at
// digit position brev(k). DIT (forward): the element at digit position m belongs to sub-transform q = brev(m),
// is multiplied by tw(q, ...), DFT with the forward root, output k stored at position k.
template <int LQ, bool DIF, class LD, class TW, class ST>
PB_D void radix_item(int u, int lstep, LD ld, TW tw, ST st) {
  constexpr int Q = 1 << LQ;
  const int j = u & ((1 << lstep) - 1);
  const int base = ((u >> lstep) << (lstep + LQ)) + j;
  u64 x[Q];
  if constexpr (DIF) {
#pragma unroll
    for (int m = 0; m < Q; m++) x[m] = ld(base + (m << lstep));
    f96::dft<LQ, true>(x);
#pragma unroll
    for (int k = 0; k < Q; k++) {
      const int slot = base + (brevq<LQ>(k) << lstep);
      st(slot, tw(k, j, slot, x[k]));
    }
  } else {
#pragma unroll
    for (int m = 0; m < Q; m++) {
      const int slot = base + (m << lstep);
      x[brevq<LQ>(m)] = tw(brevq<LQ>(m), j, slot, ld(slot));
    }
    f96::dft<LQ, false>(x);
#pragma unroll
    for (int k = 0; k < Q; k++) st(base + (k << lstep), x[k]);
  }
}

template <bool DIF, class LD, class TW, class ST>
PB_D void radix_item_lq(int lq, int u, int lstep, LD ld, TW tw, ST st) {
  switch (lq) {
    case 4: radix_item<4, DIF>(u, lstep, ld, tw, st); break;
    case 3: radix_item<3, DIF>(u, lstep, ld, tw, st); break;
    case 2: radix_item<2, DIF>(u, lstep, ld, tw, st); break;
    default: radix_item<1, DIF>(u, lstep, ld, tw, st); break;
  }
}
// the strided passes only ever use radix 16 and 8 (k1_ok)
template <bool DIF, class LD, class TW, class ST>
PB_D void radix_item_34(int lq, int u, int lstep, LD ld, TW tw, ST st) {
  if (lq == 4)
    radix_item<4, DIF>(u, lstep, ld, tw, st);
  else
    radix_item<3, DIF>(u, lstep, ld, tw, st);
}

// ---- strided passes ----------------------------------------------------------------------------------
struct StridedArgs {
  const u64* in;
  u64* out;
  size_t in_stride, out_stride;  // words between matrix columns
  int K1, lrs;                   // 2^K1 tile rows, 2^lrs words between rows
  Rounds R;
  const u64* tw;   // in-tile twiddles (R.words)
  const u64* tw4;  // four-step table, same [row][c] layout as one column
  // Row-block output of pass D (one proof across several GPUs, prover.cuh commit_sharded): row r of column `by` goes
  // to chunk r >> rb_log (the rank that will own the row) at out + chunk * rb_chunk_words + by * rb_col_stride +
  // (r mod 2^rb_log), and the first rb_halo rows of every chunk are repeated behind the previous chunk's rows (its
  // next-row halo, wrapping around) - the send buffer of the all-to-all is written by the transform itself instead
  // of by a pack copy. rb_log == 0: plain column-major output.
  int rb_log = 0;
  unsigned rb_chunks = 0;
  size_t rb_col_stride = 0, rb_chunk_words = 0, rb_halo = 0;
};
struct RowBlocks {
  u64* dst;
  int log_rows;         // rows per chunk = 2^log_rows
  unsigned chunks;
  size_t col_stride;    // words between columns inside a chunk (rows per chunk + halo)
  size_t chunk_words;   // words between chunks
  size_t halo;
};

// INV: pass A (DIF, inverse roots; last round multiplies by tw4 and stores to `out`, any representatives).
// !INV: pass D (DIT; first round multiplies by tw4, last round stores canonical values to `out`).
// RB: pass D writes row blocks (StridedArgs::rb_*); a template parameter so that the plain kernel is unchanged.
template <int LOGTC, bool INV, bool RB = false>
PB_D void strided_body(u64* sm, unsigned bx, unsigned by, int nt, const StridedArgs& a) {
  constexpr int TC = 1 << LOGTC;
  const int R = 1 << a.K1;
  u64* tile = sm;
  u64* tw = sm + ((size_t)R << LOGTC);
  const u64* src = a.in + (size_t)by * a.in_stride + (size_t)bx * TC;
  u64* dst = a.out + (size_t)by * a.out_stride + (size_t)bx * TC;
  const u64* tw4 = a.tw4 + (size_t)bx * TC;
  const int lrs = a.lrs;
#if PB_HOSTSIM
  for (int row = 0; row < R; row++) memcpy(tile + ((size_t)row << LOGTC), src + ((size_t)row << lrs), TC * 8);
  if (a.R.words) memcpy(tw, a.tw, (size_t)a.R.words * 8);
#else
  {
    u64* bar = tw + a.R.words + (a.R.words & 1);
    const int tid = threadIdx.x;
    if (tid == 0) {
      mbar_init(bar, 1);
      mbar_expect_tx(bar, (u32)(((size_t)R << LOGTC) * 8 + (size_t)a.R.words * 8));
    }
    __syncthreads();
    for (int row = tid; row < R; row += nt) bulk_g2s(tile + ((size_t)row << LOGTC), src + ((size_t)row << lrs), TC * 8, bar);
    if (tid == 0 && a.R.words) bulk_g2s(tw, a.tw, (u32)a.R.words * 8, bar);
    mbar_wait(bar, 0);
  }
#endif
  for (int ri = 0; ri < a.R.n; ri++) {
    const int lq = a.R.lq[ri], lstep = a.R.lstep[ri];
    const bool first = ri == 0, last = ri == a.R.n - 1;
    const u64* twp = tw + a.R.off[ri];
    const int items = (R >> lq) << LOGTC;
    NTT_FOR_THREADS(nt) {
      for (int t = tid; t < items; t += nt) {
        const int c = t & (TC - 1), u = t >> LOGTC;
        auto ld = [&](int slot) { return tile[(slot << LOGTC) + c]; };
        auto st_tile = [&](int slot, u64 v) { tile[(slot << LOGTC) + c] = v; };
        auto tw_tile = [&](int k, int j, int, u64 v) { return k == 0 ? v : mulz(v, twp[((k - 1) << lstep) + j]); };
        auto tw_four = [&](int, int, int slot, u64 v) { return mulz(v, tw4[((size_t)slot << lrs) + c]); };
        if constexpr (INV) {
          auto st_out = [&](int slot, u64 v) { dst[((size_t)slot << lrs) + c] = v; };
          if (last)
            radix_item_34<true>(lq, u, lstep, ld, tw_four, st_out);
          else
            radix_item_34<true>(lq, u, lstep, ld, tw_tile, st_tile);
        } else {
          auto st_out = [&](int slot, u64 v) {
            const u64 cv = canon(v);
            if constexpr (RB) {
              const size_t row = ((size_t)slot << lrs) + (size_t)bx * TC + c;
              const size_t q = row >> a.rb_log, off = row & (((size_t)1 << a.rb_log) - 1);
              u64* base = a.out + (size_t)by * a.rb_col_stride;
              base[q * a.rb_chunk_words + off] = cv;
              if (off < a.rb_halo)
                base[((q + a.rb_chunks - 1) % a.rb_chunks) * a.rb_chunk_words + ((size_t)1 << a.rb_log) + off] = cv;
            } else {
              dst[((size_t)slot << lrs) + c] = cv;
            }
          };
          if (first && last)
            radix_item_34<false>(lq, u, lstep, ld, tw_four, st_out);
          else if (first)
            radix_item_34<false>(lq, u, lstep, ld, tw_four, st_tile);
          else if (last)
            radix_item_34<false>(lq, u, lstep, ld, tw_tile, st_out);
          else
            radix_item_34<false>(lq, u, lstep, ld, tw_tile, st_tile);
        }
      }
    }
    if (!last) NTT_BARRIER();
  }
}

// ---- fused contiguous kernel ---------------------------------------------------------------------------
struct FusedArgs {
  const u64* in;
  u64* out;
  size_t in_stride, out_stride;
  int L, Kc, r, mode, canonical_out;
  Rounds fi, ff;
  const u64* tw;  // fi tables then ff tables
  const u64* sc;  // scale table of the mode, n words, block order
};

// One 2^Kc block (blockIdx.x) of one column (blockIdx.y).
PB_D void fused_body(u64* sm, unsigned bx, unsigned by, int nt, const FusedArgs& a) {
  const int C = 1 << a.Kc, r = a.r, Cp = C << r, L = a.L;
  u64* small = sm;
  u64* big = small + ((NTT_PHYS(C) + 17) & ~1);  // 16-byte aligned parts (bulk-copy destination)
  u64* tw = big + ((NTT_PHYS(Cp) + 17) & ~1);
  const u64* twi = tw;
  const u64* twf = tw + a.fi.words;
  const int tww = a.fi.words + a.ff.words;
  const size_t m0 = (size_t)bx << a.Kc;
  const u64* src = a.in + (size_t)by * a.in_stride;
  u64* dst = a.out + (size_t)by * a.out_stride;
  const u64* sc = a.sc + m0;
  const int mode = a.mode;
#if PB_HOSTSIM
  if (tww) memcpy(tw, a.tw, (size_t)tww * 8);
#else
  u64* bar = tw + tww + (tww & 1);
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_expect_tx(bar, (u32)tww * 8);
    if (tww) bulk_g2s(tw, a.tw, (u32)tww * 8, bar);
  }
#endif
  NTT_FOR_THREADS(nt) {
    if (mode == FROM_COEFFS_LDE) {
      for (int off = tid; off < C; off += nt) {
        const u64 v = mulz(src[gl::brev32((u32)(m0 + off), L)], sc[off]);
        for (int q = 0; q < (1 << r); q++) big[NTT_PHYS((off << r) + q)] = v;
      }
    } else if (C >= 2) {
#if PB_HOSTSIM
      for (int off = tid; off < C; off += nt) small[NTT_PHYS(off)] = src[m0 + off];
#else
      const ulonglong2* s2 = reinterpret_cast<const ulonglong2*>(src + m0);
      for (int i = tid; i < C / 2; i += nt) {
        const ulonglong2 v = s2[i];
        small[NTT_PHYS(2 * i)] = v.x;
        small[NTT_PHYS(2 * i + 1)] = v.y;
      }
#endif
    } else {
      if (tid == 0) small[0] = src[m0];
    }
  }
#if !PB_HOSTSIM
  __syncthreads();  // the mbarrier initialisation is visible to every waiter
  mbar_wait(bar, 0);
#endif
  NTT_BARRIER();
  if (mode != FROM_COEFFS_LDE) {
    if (a.fi.n == 0) {  // Kc == 0: a block is one coefficient
      NTT_FOR_THREADS(nt) {
        if (tid == 0) {
          const u64 v = mulz(small[0], sc[0]);
          if (mode == INTT_COSET_NAT)
            dst[gl::brev32((u32)m0, L)] = canon(v);
          else
            for (int q = 0; q < (1 << r); q++) big[NTT_PHYS(q)] = v;
        }
      }
      NTT_BARRIER();
    }
    for (int ri = 0; ri < a.fi.n; ri++) {
      const int lq = a.fi.lq[ri], lstep = a.fi.lstep[ri];
      const bool last = ri == a.fi.n - 1;
      const u64* twp = twi + a.fi.off[ri];
      const int items = C >> lq;
      NTT_FOR_THREADS(nt) {
        for (int u = tid; u < items; u += nt) {
          auto ld = [&](int slot) { return small[NTT_PHYS(slot)]; };
          auto st_tile = [&](int slot, u64 v) { small[NTT_PHYS(slot)] = v; };
          auto tw_tile = [&](int k, int j, int, u64 v) { return k == 0 ? v : mulz(v, twp[((k - 1) << lstep) + j]); };
          auto tw_scale = [&](int, int, int slot, u64 v) { return mulz(v, sc[slot]); };
          auto st_rep = [&](int slot, u64 v) {
            for (int q = 0; q < (1 << r); q++) big[NTT_PHYS((slot << r) + q)] = v;
          };
          auto st_nat = [&](int slot, u64 v) { dst[gl::brev32((u32)(m0 + slot), L)] = canon(v); };
          if (!last)
            radix_item_lq<true>(lq, u, lstep, ld, tw_tile, st_tile);
          else if (mode == INTT_COSET_NAT)
            radix_item_lq<true>(lq, u, lstep, ld, tw_scale, st_nat);
          else
            radix_item_lq<true>(lq, u, lstep, ld, tw_scale, st_rep);
        }
      }
      NTT_BARRIER();
    }
    if (mode == INTT_COSET_NAT) return;
  }
  u64* blk_out = dst + ((size_t)bx << (a.Kc + r));
  if (a.ff.n == 0) {  // Kc == 0: the block is 2^r copies
    NTT_FOR_THREADS(nt) {
      for (int off = tid; off < Cp; off += nt) blk_out[off] = a.canonical_out ? canon(big[NTT_PHYS(off)]) : big[NTT_PHYS(off)];
    }
    return;
  }
  for (int ri = 0; ri < a.ff.n; ri++) {
    const int lq = a.ff.lq[ri], lstep = a.ff.lstep[ri];
    const bool last = ri == a.ff.n - 1;
    const u64* twp = twf + a.ff.off[ri];
    const int items = Cp >> lq;
    const bool can = a.canonical_out != 0;
    NTT_FOR_THREADS(nt) {
      for (int u = tid; u < items; u += nt) {
        auto ld = [&](int slot) { return big[NTT_PHYS(slot)]; };
        auto st_tile = [&](int slot, u64 v) { big[NTT_PHYS(slot)] = v; };
        auto st_out = [&](int slot, u64 v) { blk_out[slot] = can ? canon(v) : v; };
        auto tw_tile = [&](int k, int j, int, u64 v) { return k == 0 ? v : mulz(v, twp[((k - 1) << lstep) + j]); };
        auto tw_none = [&](int, int, int, u64 v) { return v; };
        if (lstep == 0) {
          if (last)
            radix_item_lq<false>(lq, u, lstep, ld, tw_none, st_out);
          else
            radix_item_lq<false>(lq, u, lstep, ld, tw_none, st_tile);
        } else {
          if (last)
            radix_item_lq<false>(lq, u, lstep, ld, tw_tile, st_out);
          else
            radix_item_lq<false>(lq, u, lstep, ld, tw_tile, st_tile);
        }
      }
    }
    if (!last) NTT_BARRIER();
  }
}

#if !PB_HOSTSIM
#ifndef NTT_LB_CTAS
#define NTT_LB_CTAS 2
#endif
template <int LOGTC, bool INV, bool RB = false>
static __global__ void __launch_bounds__(256, NTT_LB_CTAS) k_ntt_strided(const __grid_constant__ StridedArgs a) {
  extern __shared__ __align__(16) u64 ntt_sm[];
  strided_body<LOGTC, INV, RB>(ntt_sm, blockIdx.x, blockIdx.y, (int)blockDim.x, a);
}
static __global__ void __launch_bounds__(256, NTT_LB_CTAS) k_ntt_fused(const __grid_constant__ FusedArgs a) {
  extern __shared__ __align__(16) u64 ntt_sm[];
  fused_body(ntt_sm, blockIdx.x, blockIdx.y, (int)blockDim.x, a);
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
  const int lim = 200 * 1024 + 64;
  PB_CUDA(cudaFuncSetAttribute(k_ntt_strided<3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
  PB_CUDA(cudaFuncSetAttribute(k_ntt_strided<3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
  PB_CUDA(cudaFuncSetAttribute(k_ntt_strided<4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
  PB_CUDA(cudaFuncSetAttribute(k_ntt_strided<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
  PB_CUDA(cudaFuncSetAttribute(k_ntt_strided<3, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
  PB_CUDA(cudaFuncSetAttribute(k_ntt_strided<4, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
  PB_CUDA(cudaFuncSetAttribute(k_ntt_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
  g_ntt_attr_set[dev] = true;
}
#endif

static inline unsigned threads_for(size_t items) { return items >= 256 ? 256u : items >= 32 ? (unsigned)items : 32u; }

template <bool INV, bool RB = false>
static inline void launch_strided(const Plan& p, const StridedArgs& a, unsigned tiles, int ncols, pbStream s) {
  const size_t smem = (((size_t)1 << (p.K1 + p.logtc)) + a.R.words + 4) * 8;
  const unsigned nt = threads_for(((size_t)1 << (p.K1 + p.logtc)) >> 4);
#if PB_HOSTSIM
  (void)s;
#pragma omp parallel
  {
    std::vector<u64> sm(smem / 8);
#pragma omp for collapse(2) schedule(dynamic, 1)
    for (int by = 0; by < ncols; by++)
      for (unsigned bx = 0; bx < tiles; bx++) {
        if (p.logtc == 4)
          strided_body<4, INV, RB>(sm.data(), bx, (unsigned)by, (int)nt, a);
        else if (p.logtc == 3)
          strided_body<3, INV, RB>(sm.data(), bx, (unsigned)by, (int)nt, a);
        else if (p.logtc == 2)
          strided_body<2, INV, RB>(sm.data(), bx, (unsigned)by, (int)nt, a);
        else if (p.logtc == 1)
          strided_body<1, INV, RB>(sm.data(), bx, (unsigned)by, (int)nt, a);
        else
          strided_body<0, INV, RB>(sm.data(), bx, (unsigned)by, (int)nt, a);
      }
  }
#else
  dim3 grid(tiles, (unsigned)ncols);
  if (p.logtc == 4)
    k_ntt_strided<4, INV, RB><<<grid, nt, smem, s>>>(a);
  else if (p.logtc == 3)
    k_ntt_strided<3, INV, RB><<<grid, nt, smem, s>>>(a);
  else
    throw Pb254Error(6, "ntt: size not supported");
  g_pb_launches++;
  pb_check_last(INV ? "ntt pass A" : "ntt pass D");
#endif
}

static inline void launch_fused(const Plan& p, const FusedArgs& a, int ncols, pbStream s) {
  const size_t smem = (p.fused_smem_words() + 4) * 8;
  const unsigned nt = threads_for(((size_t)1 << (p.Kc + p.r)) >> 4);
  const unsigned blocks = 1u << (p.L - p.Kc);
#if PB_HOSTSIM
  (void)s;
#pragma omp parallel
  {
    std::vector<u64> sm(smem / 8);
#pragma omp for collapse(2) schedule(dynamic, 1)
    for (int by = 0; by < ncols; by++)
      for (unsigned bx = 0; bx < blocks; bx++) fused_body(sm.data(), bx, (unsigned)by, (int)nt, a);
  }
#else
  dim3 grid(blocks, (unsigned)ncols);
  k_ntt_fused<<<grid, nt, smem, s>>>(a);
  g_pb_launches++;
  pb_check_last("ntt fused");
#endif
}

static inline void check_align(const void* p, size_t stride) {
  if (((uintptr_t)p & 15) || (stride & 1)) throw Pb254Error(6, "ntt: matrices must be 16-byte aligned with an even column stride");
}

// mode FROM_VALUES_LDE / FROM_COEFFS_LDE: LDE of `ncols` columns. in: [col][in_stride] (n used), out: [col][out_stride]
// (N = n << r used). mode INTT_COSET_NAT (r must be 0): natural-order values on 7 <w> -> natural-order coefficients.
// scratch: at least ncols * n words (unused when K1 == 0 or in FROM_COEFFS mode). out may not alias in.
// rb (LDE modes only): the last pass writes the result in row-block layout to rb->dst instead of column-major to
// `out` (which is then only the intermediate buffer); returns false - and ignores rb - when the transform is too small
// to have a strided last pass, the caller then repacks `out` itself.
static inline bool transform_columns(TableSet& ts, const u64* in, size_t in_stride, u64* out, size_t out_stride, u64* scratch,
                                     int ncols, int L, int r, int mode, pbStream s, const RowBlocks* rb = nullptr) {
  if (ncols == 0) return rb != nullptr;
  const size_t n = (size_t)1 << L;
#if !PB_HOSTSIM
  set_smem_attrs();
#endif
  check_align(in, in_stride);
  check_align(out, out_stride);
  PlanCache& pc = plan_cache(ts);
  Plan& p = pc.get(ts.t, L, r, s);
  const u64* sc = pc.scale(p, ts.t, mode, s);
  const u64* fused_in = in;
  size_t fused_stride = in_stride;
  if (mode != FROM_COEFFS_LDE && p.K1 > 0) {
    check_align(scratch, n);
    StridedArgs a = {in, scratch, in_stride, n, p.K1, p.Kc, p.a, p.tw_a, p.tw4_a};
    launch_strided<true>(p, a, (unsigned)(n >> (p.K1 + p.logtc)), ncols, s);
    fused_in = scratch;
    fused_stride = n;
  }
  {
    FusedArgs a = {fused_in, out, fused_stride, out_stride, L, p.Kc, r, mode, p.K1 == 0 ? 1 : 0, p.fi, p.ff, p.tw_f, sc};
    launch_fused(p, a, ncols, s);
  }
  if (mode != INTT_COSET_NAT && p.K1 > 0) {
    StridedArgs a = {out, out, out_stride, out_stride, p.K1, p.Kc + r, p.d, p.tw_d, p.tw4_d};
    if (rb) {
      a.out = rb->dst;
      a.rb_log = rb->log_rows;
      a.rb_chunks = rb->chunks;
      a.rb_col_stride = rb->col_stride;
      a.rb_chunk_words = rb->chunk_words;
      a.rb_halo = rb->halo;
      launch_strided<false, true>(p, a, (unsigned)((n << r) >> (p.K1 + p.logtc)), ncols, s);
      return true;
    }
    launch_strided<false>(p, a, (unsigned)((n << r) >> (p.K1 + p.logtc)), ncols, s);
    return false;
  }
  return false;
}

static inline void lde_columns(TableSet& ts, const u64* in, size_t in_stride, u64* out, size_t out_stride, u64* scratch,
                               int ncols, int L, int r, int mode, pbStream s) {
  transform_columns(ts, in, in_stride, out, out_stride, scratch, ncols, L, r, mode, s);
}
// LDE whose last pass writes row blocks (see StridedArgs::rb_*); false: `out` holds the plain LDE, nothing was written
// to rb->dst.
static inline bool lde_columns_row_blocks(TableSet& ts, const u64* in, size_t in_stride, u64* out, size_t out_stride,
                                          u64* scratch, int ncols, int L, int r, int mode, const RowBlocks& rb, pbStream s) {
  return transform_columns(ts, in, in_stride, out, out_stride, scratch, ncols, L, r, mode, s, &rb);
}

// coset iNTT (shift 7) of `ncols` columns of length n = 2^L: natural-order values on 7*<w> ->
// natural-order coefficients. out may not alias in. scratch: ncols * n words.
static inline void coset_intt_columns(TableSet& ts, const u64* in, size_t in_stride, u64* out, size_t out_stride, u64* scratch,
                                      int ncols, int L, pbStream s) {
  transform_columns(ts, in, in_stride, out, out_stride, scratch, ncols, L, 0, INTT_COSET_NAT, s);
}

}  // namespace ntt
