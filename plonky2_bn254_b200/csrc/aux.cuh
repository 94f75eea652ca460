// K6 `lookup_helpers` + K7 `ctl_z`: the auxiliary (logUp range-check + cross-table-lookup) columns.
// Replaces, bit-exact,
//   lookup_helper_columns      starky 0.4.0 lookup.rs   (reached via prove_with_commitment,
//                              src/starks/common/prover.rs:55-65; descriptor
//                              src/starks/curves/g1/scalar_mul_stark.rs:493-500)
//   get_ctl_data/partial_sums  starky 0.4.0 cross_table_lookup.rs (src/starks/common/prover.rs:46-52;
//                              descriptors g1/scalar_mul_ctl.rs:20-55, fields/exp_ctl.rs:18-51)
// Auxiliary matrix layout (column-major, A x n):  for each challenge j: [h_0 .. h_(H-1), Z_j], then
// the CTL Z columns in CTL-major / challenge-minor order.
//   h_k[i]   = 1/(f_2k[i] + beta_j) + 1/(f_2k+1[i] + beta_j)
//   Z_j[0]   = 0,  Z_j[i+1] = Z_j[i] + sum_k h_k[i] - freq[i] / (table[i] + beta_j)
//   Zctl[i]  = sum_{t >= i, filter[t] != 0} 1 / (sum_k v_k[t] beta^k + gamma)
// Every looked-up cell is < 2^16 (range-checked while filling the frequency column), so the
// inverses come from a 65536-entry table per challenge instead of one field inversion per cell.
// Bound: HBM (reads the 450 / 900 / 128 range-checked columns once, writes H columns per challenge).
#pragma once
#include <functional>
#include "context.cuh"
#include "tracegen.h"

namespace aux {

static constexpr int MAXCH = 4;

struct Challenges {
  int nch;
  u64 beta[MAXCH], gamma[MAXCH];
};

struct InvTableK {  // invt[j][v] = 1 / (v + beta_j)
  u64* invt;
  Challenges ch;
  PB_HD void operator()(size_t gid) const {
    size_t j = gid >> 16, v = gid & 0xffff;
    u64 d = gl::add((u64)v, ch.beta[j]);
    invt[gid] = d ? gl::inv(d) : 0;  // beta = -v has probability 2^-48; the verifier would reject
  }
};

// pb254_prove_trace only: the cells a generated trace has range-checked by construction (one thread per row)
struct RangePrecheckK {
  const u64* trace;
  size_t n;
  int rc_lo, rc_hi, table_col;
  int* err;
  PB_HD void operator()(size_t i) const {
    u64 bad = trace[(size_t)table_col * n + i] >> 16;
    for (int c = rc_lo; c < rc_hi; c++) bad |= trace[(size_t)c * n + i] >> 16;
    if (bad) *err = 1;
  }
};

struct HelpersK {  // one thread per row
  const u64* trace;
  u64* auxm;       // A x n
  u64* xs;         // nch x n row increments of Z
  const u64* invt;
  size_t n;
  int rc_lo, ncols, nh, freq_col, table_col, nch;
  PB_HD void operator()(size_t i) const {
    u64 rowsum[MAXCH];
#pragma unroll
    for (int j = 0; j < MAXCH; j++) rowsum[j] = 0;
    for (int k = 0; k < nh; k++) {
      u64 f0 = trace[(size_t)(rc_lo + 2 * k) * n + i] & 0xffff;
      bool pair = 2 * k + 1 < ncols;
      u64 f1 = pair ? (trace[(size_t)(rc_lo + 2 * k + 1) * n + i] & 0xffff) : 0;
#pragma unroll
      for (int j = 0; j < MAXCH; j++) {
        if (j < nch) {
          u64 h = invt[((size_t)j << 16) + f0];
          if (pair) h = gl::add(h, invt[((size_t)j << 16) + f1]);
          auxm[((size_t)j * (nh + 1) + k) * n + i] = h;
          rowsum[j] = gl::add(rowsum[j], h);
        }
      }
    }
    u64 fr = trace[(size_t)freq_col * n + i];
    u64 tb = trace[(size_t)table_col * n + i] & 0xffff;
#pragma unroll
    for (int j = 0; j < MAXCH; j++)
      if (j < nch) xs[(size_t)j * n + i] = gl::sub(rowsum[j], gl::mul(fr, invt[((size_t)j << 16) + tb]));
  }
};

// ---- modular prefix sums as three per-thread passes (chunk sums, scan of sums, chunk scan) -----
static constexpr int SCAN_CHUNK = 256;
struct ChunkSumK {
  const u64* in;  // ncol x n
  u64* sums;      // ncol x nchunks
  size_t n, nchunks;
  PB_HD void operator()(size_t gid) const {
    size_t col = gid / nchunks, ck = gid % nchunks;
    const u64* p = in + col * n + ck * SCAN_CHUNK;
    u64 s = 0;
    for (int i = 0; i < SCAN_CHUNK; i++) s = gl::add(s, p[i]);
    sums[gid] = s;
  }
};
struct SumsScanK {  // exclusive scan of the chunk sums of one column (forward or backward)
  u64* sums;
  size_t nchunks;
  int reverse;
  u64* totals;  // null, or [ncol]: the sum of the whole column (row-block form)
  PB_HD void operator()(size_t col) const {
    u64* p = sums + col * nchunks;
    u64 run = 0;
    for (size_t t = 0; t < nchunks; t++) {
      size_t ck = reverse ? nchunks - 1 - t : t;
      u64 v = p[ck];
      p[ck] = run;
      run = gl::add(run, v);
    }
    if (totals) totals[col] = run;
  }
};
// Row-block form (one proof across several GPUs): the matrix is the row block [rank n, (rank + 1) n) of the whole one;
// the running sums continue across ranks, so every chunk prefix gets the totals of the ranks before (forward) or
// after (reverse) this one. gather(send, recv, bytes) is the caller's all-gather.
struct BlockScan {
  int world, rank;
  std::function<void(const void*, void*, size_t)> gather;
  u64* totals;      // [max columns of one scan]
  u64* all_totals;  // [world][max columns]
};
struct CarryK {
  u64* sums;
  size_t nchunks;
  const u64* all_totals;
  int ncol, world, rank, reverse;
  PB_HD void operator()(size_t gid) const {
    const size_t col = gid / nchunks;
    u64 carry = 0;
    for (int p = 0; p < world; p++)
      if (reverse ? p > rank : p < rank) carry = gl::add(carry, all_totals[(size_t)p * ncol + col]);
    sums[gid] = gl::add(sums[gid], carry);
  }
};
// forward exclusive:  out[i] = sum_{t < i} in[t];   reverse inclusive: out[i] = sum_{t >= i} in[t]
struct ChunkScanK {
  const u64* in;
  const u64* sums;
  u64* out;          // column c written at out + out_cols[c] * n
  size_t n, nchunks;
  int reverse;
  int out_col0, out_col_stride;
  PB_HD void operator()(size_t gid) const {
    size_t col = gid / nchunks, ck = gid % nchunks;
    const u64* p = in + col * n + ck * SCAN_CHUNK;
    u64* o = out + (size_t)(out_col0 + (int)col * out_col_stride) * n + ck * SCAN_CHUNK;
    u64 run = sums[gid];
    if (!reverse) {
      for (int i = 0; i < SCAN_CHUNK; i++) {
        o[i] = run;
        run = gl::add(run, p[i]);
      }
    } else {
      for (int i = SCAN_CHUNK - 1; i >= 0; i--) {
        run = gl::add(run, p[i]);
        o[i] = run;
      }
    }
  }
};
static inline void scan_columns(const u64* in, u64* sums, u64* out, size_t n, int ncol, int reverse, int out_col0,
                                int out_col_stride, pbStream s, const BlockScan* bs = nullptr) {
  if (n % SCAN_CHUNK) throw Pb254Error(6, "scan: n must be a multiple of 256");
  size_t nchunks = n / SCAN_CHUNK;
  pb_launch("scan chunk sums", ChunkSumK{in, sums, n, nchunks}, (size_t)ncol * nchunks, s, 64);
  pb_launch("scan sums", SumsScanK{sums, nchunks, reverse, bs ? bs->totals : nullptr}, (size_t)ncol, s, 32);
  if (bs) {
    bs->gather(bs->totals, bs->all_totals, (size_t)ncol * 8);
    pb_launch("scan carry", CarryK{sums, nchunks, bs->all_totals, ncol, bs->world, bs->rank, reverse},
              (size_t)ncol * nchunks, s, 64);
  }
  pb_launch("scan chunks", ChunkScanK{in, sums, out, n, nchunks, reverse, out_col0, out_col_stride},
            (size_t)ncol * nchunks, s, 64);
}

// ---- CTL ------------------------------------------------------------------------------------
struct CtlDesc {       // device-resident flattened descriptors of the looked tables
  const int* term_col;  // column of term t
  const u64* term_coef;
  const int* col_start;  // [ncols + 1] into terms, per CTL: ctl_first_col[c] .. ctl_first_col[c+1]
  int ctl_first_col[3];
  int filter_col[2];
};
struct CtlTermsK {  // one thread per row: terms[c][j][i] = filter ? 1/comb : 0
  const u64* trace;
  u64* terms;  // (2 * nch) x n
  CtlDesc d;
  Challenges ch;
  size_t n;
  PB_HD void operator()(size_t i) const {
    for (int c = 0; c < 2; c++) {
      bool on = trace[(size_t)d.filter_col[c] * n + i] != 0;
      u64 comb[MAXCH];
#pragma unroll
      for (int j = 0; j < MAXCH; j++) comb[j] = 0;
      if (on) {
        for (int k = d.ctl_first_col[c + 1] - 1; k >= d.ctl_first_col[c]; k--) {
          u64 v = 0;
          for (int t = d.col_start[k]; t < d.col_start[k + 1]; t++)
            v = gl::add(v, gl::mul(trace[(size_t)d.term_col[t] * n + i], d.term_coef[t]));
#pragma unroll
          for (int j = 0; j < MAXCH; j++)
            if (j < ch.nch) comb[j] = gl::add(gl::mul(comb[j], ch.beta[j]), v);
        }
      }
#pragma unroll
      for (int j = 0; j < MAXCH; j++)
        if (j < ch.nch) {
          u64 t = 0;
          if (on) {
            u64 cb = gl::add(comb[j], ch.gamma[j]);
            t = cb ? gl::inv(cb) : 0;
          }
          terms[((size_t)c * ch.nch + j) * n + i] = t;
        }
    }
  }
};

struct HostCtl {
  std::vector<int> term_col, col_start;
  std::vector<u64> term_coef;
  int ctl_first_col[3];
  int filter_col[2];
  int ncols(int c) const { return ctl_first_col[c + 1] - ctl_first_col[c]; }
};
// g1_scalar_mul_ctl / g2_scalar_mul_ctl / fq_exp_ctl
static inline HostCtl host_ctl(const tg::Layout& l) {
  HostCtl h;
  auto begin_col = [&] { h.col_start.push_back((int)h.term_col.size()); };
  auto single = [&](int c) {
    begin_col();
    h.term_col.push_back(c);
    h.term_coef.push_back(1);
  };
  h.ctl_first_col[0] = 0;
  for (int i = 0; i < l.L; i++) single(l.b + i);  // x
  if (l.kind != 2)
    for (int i = 0; i < l.L; i++) single(l.a + i);  // offset
  for (int k = 0; k < 16; k++) {                    // s as 16 little-endian-bit limbs
    begin_col();
    for (int i = 0; i < 16; i++) {
      h.term_col.push_back(l.bits + 16 * k + i);
      h.term_coef.push_back((u64)1 << i);
    }
  }
  single(l.ts);
  h.ctl_first_col[1] = (int)h.col_start.size();
  for (int i = 0; i < l.L; i++) single(l.reg1 + i);  // sum / product
  single(l.ts);
  h.ctl_first_col[2] = (int)h.col_start.size();
  begin_col();
  h.filter_col[0] = l.rf + 0;
  h.filter_col[1] = l.rf + 1;
  return h;
}

static inline int num_helpers(const tg::Layout& l) { return (l.rc_hi - l.rc_lo + 1) / 2; }
static inline int num_aux(const tg::Layout& l, int nch) { return (num_helpers(l) + 1) * nch + 2 * nch; }

// Builds the A x n auxiliary matrix on the device. With `bs`: d_trace / d_aux are the row blocks (n rows each, stride n)
// of this rank, the running sums continue across the ranks (BlockScan).
static inline void build(Arena& ar, const tg::Layout& l, const u64* d_trace, size_t n, const Challenges& ch,
                         u64* d_aux, pbStream s, BlockScan* bs = nullptr) {
  const int nch = ch.nch, nh = num_helpers(l);
  u64* invt = ar.alloc_n<u64>((size_t)nch << 16);
  u64* xs = ar.alloc_n<u64>((size_t)2 * nch * n);  // reused for the CTL terms
  u64* sums = ar.alloc_n<u64>((size_t)2 * nch * (n / SCAN_CHUNK) + 16);
  if (bs) {
    bs->totals = ar.alloc_n<u64>((size_t)2 * nch);
    bs->all_totals = ar.alloc_n<u64>((size_t)bs->world * 2 * nch);
  }
  pb_launch("lookup inverse table", InvTableK{invt, ch}, (size_t)nch << 16, s, 128);
  pb_launch("lookup helpers",
            HelpersK{d_trace, d_aux, xs, invt, n, l.rc_lo, l.rc_hi - l.rc_lo, nh, l.freq, l.range_counter, nch}, n, s,
            128);
  // Z_j into column j * (nh + 1) + nh
  scan_columns(xs, sums, d_aux, n, nch, 0, nh, nh + 1, s, bs);
  HostCtl hc = host_ctl(l);
  int* d_tc = ar.alloc_n<int>(hc.term_col.size());
  u64* d_tk = ar.alloc_n<u64>(hc.term_coef.size());
  int* d_cs = ar.alloc_n<int>(hc.col_start.size());
  pb_h2d(d_tc, hc.term_col.data(), hc.term_col.size() * sizeof(int), s);
  pb_h2d(d_tk, hc.term_coef.data(), hc.term_coef.size() * sizeof(u64), s);
  pb_h2d(d_cs, hc.col_start.data(), hc.col_start.size() * sizeof(int), s);
  pb_sync(s);  // the host vectors above go out of scope
  CtlDesc d;
  d.term_col = d_tc;
  d.term_coef = d_tk;
  d.col_start = d_cs;
  for (int i = 0; i < 3; i++) d.ctl_first_col[i] = hc.ctl_first_col[i];
  d.filter_col[0] = hc.filter_col[0];
  d.filter_col[1] = hc.filter_col[1];
  pb_launch("ctl terms", CtlTermsK{d_trace, xs, d, ch, n}, n, s, 128);
  scan_columns(xs, sums, d_aux, n, 2 * nch, 1, (nh + 1) * nch, 1, s, bs);
}

}  // namespace aux
