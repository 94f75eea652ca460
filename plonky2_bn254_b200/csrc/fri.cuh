// K9 `open_at`, K10 `fri_combine`, K11 `fri_fold`, K12 `pow_grind`, K13 `query_gather`.
// Replaces, bit-exact (all arithmetic is exact in F_p / F_p^2, so evaluation-domain formulations
// give the same canonical values as the reference's coefficient-domain ones):
//   StarkOpeningSet::new                  starky 0.4.0 proof.rs
//   PolynomialBatch::prove_openings       plonky2 0.2.2 fri/oracle.rs
//   fri_committed_trees / fri_proof_of_work / fri_prover_query_rounds   plonky2 0.2.2 fri/prover.rs
// (un-vendored; reached from src/starks/common/prover.rs:55-65).
//
//   open_at      P(z) for a column given by its values on H = <w>:  barycentric form
//                P(z) = (z^n - 1)/n * sum_j v_j w^j / (z - w^j); the weight vector is shared by all
//                columns and the weights for g*z are the same vector rotated by one, so trace and
//                auxiliary coefficients are never materialised.
//   fri_combine  final(x) = ((F0(x) - F0(z))/(x - z) a^|b1| + (F1(x) - F1(gz))/(x - gz)) a^|b2|
//                           + (F2(x) - F2(1))/(x - 1),  F_b = sum_j a^j f_j, evaluated pointwise on the LDE
//                domain (the reference divides the coefficient vector by (X - z); same polynomial).
//   fri_fold     arity-16 fold on evaluations: for a coset {x0 h^m}, folded(x0^16) =
//                sum_m e_m * (1/16) sum_i (beta h^-m / x0)^i  (the reference folds coefficients and
//                re-evaluates by coset FFT).
// All of these are HBM-bound or tiny; fri_combine reads both committed LDE matrices once.
#pragma once
#include "context.cuh"
#include "poseidon.cuh"

namespace fri {

using gl::E2;

PB_HD void atomic_min_u64(u64* p, u64 v) {
#ifdef __CUDA_ARCH__
  atomicMin((unsigned long long*)p, (unsigned long long)v);
#else
  u64 cur = __atomic_load_n(p, __ATOMIC_RELAXED);
  while (v < cur && !__atomic_compare_exchange_n(p, &cur, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {
  }
#endif
}

// ---- K9 openings ---------------------------------------------------------------------------
struct BaryWeightsK {  // wz[j] = w^j / (zeta - w^j)
  E2* wz;
  E2 zeta;
  ntt::Tables t;
  int log_n;
  PB_HD void operator()(size_t j) const {
    u64 wj = ntt::tpow(t.fwd_lo, t.fwd_hi, (u64)j << (ntt::LOG_M - log_n));
    E2 den = gl::e2(gl::sub(zeta.a, wj), zeta.b);
    wz[j] = gl::emul_base(gl::einv(den), wj);
  }
};
struct PowTableK {  // pw[i] = zeta^i
  E2* pw;
  E2 zeta;
  PB_HD void operator()(size_t i) const { pw[i] = gl::epow(zeta, (u64)i); }
};
static constexpr int PARTS = 1024;  // row partitions per column: W x PARTS / 256 CTAs, several waves on 148 SMs
// partial[col][part] = { sum v w[i], sum v w[i-1] } over rows i = part (mod PARTS)
struct WeightedPartialK {
  const u64* vals;
  size_t stride, n;
  const E2* w;
  u64* partial;  // [col][PARTS][4]
  int with_next;
  PB_HD void operator()(size_t gid) const {
    size_t col = gid / PARTS, part = gid % PARTS;
    gl::Acc a0, a1, b0, b1;
    const u64* v = vals + col * stride;
    for (size_t i = part; i < n; i += PARTS) {
      u64 x = v[i];
      E2 w0 = w[i];
      a0.mac(x, w0.a);
      a1.mac(x, w0.b);
      if (with_next) {
        E2 w1 = w[(i + n - 1) & (n - 1)];
        b0.mac(x, w1.a);
        b1.mac(x, w1.b);
      }
    }
    u64* o = partial + gid * 4;
    o[0] = a0.reduce();
    o[1] = a1.reduce();
    o[2] = b0.reduce();
    o[3] = b1.reduce();
  }
};
// The 1024 partial sums of a column are added up in two steps (32 x 32) so that the second step is not one thread
// per column walking 4096 words: WeightedMidK leaves the sum of partials [32 g, 32 g + 32) in slot 32 g.
static constexpr int MID = 32;
struct WeightedMidK {
  u64* partial;
  PB_HD void operator()(size_t gid) const {  // gid = col * (PARTS / MID) + g
    u64* p = partial + gid * MID * 4;
    u64 s[4] = {0, 0, 0, 0};
    for (int q = 0; q < MID; q++)
      for (int k = 0; k < 4; k++) s[k] = gl::add(s[k], p[q * 4 + k]);
    for (int k = 0; k < 4; k++) p[k] = s[k];
  }
};
struct WeightedFinalK {
  const u64* partial;
  E2 scale;
  E2* out;       // [ncols]
  E2* out_next;  // [ncols] or null
  PB_HD void operator()(size_t col) const {
    u64 s[4] = {0, 0, 0, 0};
    for (int p = 0; p < PARTS; p += MID)
      for (int k = 0; k < 4; k++) s[k] = gl::add(s[k], partial[(col * PARTS + p) * 4 + k]);
    out[col] = gl::emul(gl::e2(s[0], s[1]), scale);
    if (out_next) out_next[col] = gl::emul(gl::e2(s[2], s[3]), scale);
  }
};
static inline void finish_openings(u64* partial, size_t ncols, E2 scale, E2* out, E2* out_next, pbStream s) {
  pb_launch("open mid", WeightedMidK{partial}, ncols * (PARTS / MID), s, 128);
  pb_launch("open fin", WeightedFinalK{partial, scale, out, out_next}, ncols, s, 64);
}

// ---- K10 combine -----------------------------------------------------------------------------
struct CombineK {
  const u64 *tr, *ax, *qt;
  size_t N;
  // Row-block form (prover.cuh, one proof across several GPUs): the matrices hold the rows [i_base, i_base + count) of
  // the LDE with their own column strides and out[il] (natural order) is the value of row i_base + il.
  // Whole-domain form: strides = N, i_base = 0, natural_out = 0 (bit-reversed output index).
  size_t tr_stride, ax_stride, qt_stride, i_base = 0;
  int natural_out = 0;
  int W, A, Q, nlk, nz;  // nlk = first CTL-Z column inside the auxiliary matrix, nz = 2 * num_challenges
  const E2* apow;        // alpha^c, c < W + A + Q
  E2 O0, O1, O2;         // sum_j alpha^j opening_j per batch
  E2 zeta, zeta_next, sh1, sh2;
  ntt::Tables t;
  int log_N;
  E2* out;  // bit-reversed order
  PB_HD void operator()(size_t il) const {
    const size_t i = i_base + il;
    const u64 x = gl::mul(gl::COSET_SHIFT, ntt::tpow(t.fwd_lo, t.fwd_hi, (u64)i << (ntt::LOG_M - log_N)));
    gl::Acc sa, sb, fa, fb;
#pragma unroll 4
    for (int c = 0; c < W; c++) {
      u64 v = tr[(size_t)c * tr_stride + il];
      E2 a = apow[c];
      sa.mac(a.a, v);
      sb.mac(a.b, v);
    }
#pragma unroll 4
    for (int c = 0; c < A; c++) {
      u64 v = ax[(size_t)c * ax_stride + il];
      E2 a = apow[W + c];
      sa.mac(a.a, v);
      sb.mac(a.b, v);
      if (c >= nlk) {
        E2 b = apow[c - nlk];
        fa.mac(b.a, v);
        fb.mac(b.b, v);
      }
    }
    E2 s_ta = gl::e2(sa.reduce(), sb.reduce());
    gl::Acc qa, qb;
    for (int q = 0; q < Q; q++) {
      u64 v = qt[(size_t)q * qt_stride + il];
      E2 a = apow[W + A + q];
      qa.mac(a.a, v);
      qb.mac(a.b, v);
    }
    E2 f0 = gl::eadd(s_ta, gl::e2(qa.reduce(), qb.reduce()));
    E2 f2 = gl::e2(fa.reduce(), fb.reduce());
    E2 t0 = gl::emul(gl::esub(f0, O0), gl::einv(gl::e2(gl::sub(x, zeta.a), gl::neg(zeta.b))));
    E2 t1 = gl::emul(gl::esub(s_ta, O1), gl::einv(gl::e2(gl::sub(x, zeta_next.a), gl::neg(zeta_next.b))));
    E2 t2 = gl::emul_base(gl::esub(f2, O2), gl::inv(gl::sub(x, 1)));
    E2 r = gl::eadd(gl::emul(gl::eadd(gl::emul(t0, sh1), t1), sh2), t2);
    out[natural_out ? il : (size_t)gl::brev32((u32)i, log_N)] = r;
  }
};
// natural order -> bit-reversed order (row-block form of the combination, after the all-gather)
struct BitReverseE2K {
  const E2* in;
  E2* out;
  int log_N;
  PB_HD void operator()(size_t i) const { out[gl::brev32((u32)i, log_N)] = in[i]; }
};

// ---- K11 fold ----------------------------------------------------------------------------------
struct FoldK {
  const E2* in;  // length 2^log_len, bit-reversed order
  E2* out;       // length 2^(log_len - arity_bits)
  E2 beta;
  u64 shift_inv;  // inverse of this layer's coset shift
  ntt::Tables t;
  int log_len, arity_bits;
  PB_HD void operator()(size_t l) const {
    const int arity = 1 << arity_bits;
    u64 x0inv = gl::mul(shift_inv, ntt::tpow(t.inv_lo, t.inv_hi, (u64)gl::brev32((u32)l, log_len - arity_bits)
                                                                     << (ntt::LOG_M - log_len)));
    E2 y = gl::emul_base(beta, x0inv);
    const u64 h_inv = ntt::tpow(t.inv_lo, t.inv_hi, (u64)1 << (ntt::LOG_M - arity_bits));
    u64 hm = 1;
    E2 acc = gl::e2(0, 0);
    const E2 one = gl::e2(1, 0);
    for (int m = 0; m < arity; m++) {
      E2 tm = gl::emul_base(y, hm);
      // sum_{i < arity} tm^i = prod_{k < arity_bits} (1 + tm^(2^k))
      E2 g = gl::eadd(one, tm), pw = tm;
      for (int k = 1; k < arity_bits; k++) {
        pw = gl::emul(pw, pw);
        g = gl::emul(g, gl::eadd(one, pw));
      }
      E2 e = in[(l << arity_bits) + gl::brev32((u32)m, arity_bits)];
      acc = gl::eadd(acc, gl::emul(e, g));
      hm = gl::mul(hm, h_inv);
    }
    out[l] = gl::emul_base(acc, gl::inv((u64)arity));
  }
};

// ---- K12 proof of work --------------------------------------------------------------------
struct PowK {
  u64 state[12];
  int pos, pow_bits;
  u64 base;
  u64* result;  // initialised to ~0
  PB_HD void operator()(size_t tix) const {
    u64 s[12];
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = state[i];
    u64 w = base + tix;
    // pos is uniform; write through a switch-free loop so `s` stays in registers
#pragma unroll
    for (int i = 0; i < 8; i++)
      if (i == pos) s[i] = w;
    poseidon::permute(s);
    if (pow_bits == 0 || (s[7] >> (64 - pow_bits)) == 0) atomic_min_u64(result, w);
  }
};

// ---- K13 query gather ----------------------------------------------------------------------
struct Section {
  int type;      // 0 leaf of a column-major LDE, 1 Merkle siblings, 2 contiguous row
  int off;       // word offset inside the per-query record
  int words;
  int shift;     // index >> shift selects the leaf of this tree
  int log_n;     // log2(number of leaves)
  const u64* ptr;
  size_t stride;
  // Row-block form: a type-0 section with rows > 0 holds only the LDE rows [row0, row0 + rows); other rows read as 0
  // (the owning rank supplies them, prover.cuh adds the ranks' records up). Sections with rows == 0 are whole.
  size_t row0 = 0, rows = 0;
};
static constexpr int MAX_SECTIONS = 32;
struct GatherK {
  Section sec[MAX_SECTIONS];
  int nsec, rec_words;
  const u64* indices;
  u64* out;
  int zero_whole = 0;  // row-block form: ranks other than 0 write 0 for the sections every rank holds in full
  PB_HD void operator()(size_t gid) const {
    size_t q = gid / rec_words;
    int w = (int)(gid % rec_words);
    int si = 0;
    for (int k = 1; k < nsec; k++)
      if (w >= sec[k].off) si = k;
    const Section& s = sec[si];
    int r = w - s.off;
    size_t idx = (size_t)(indices[q] >> s.shift);
    u64 v;
    if (s.type == 0 && s.rows) {
      const size_t row = gl::brev32((u32)idx, s.log_n);
      v = (row >= s.row0 && row < s.row0 + s.rows) ? s.ptr[(size_t)r * s.stride + (row - s.row0)] : 0;
    } else if (zero_whole) {
      v = 0;
    } else if (s.type == 0) {
      v = s.ptr[(size_t)r * s.stride + gl::brev32((u32)idx, s.log_n)];
    } else if (s.type == 1) {
      int lv = r >> 2;
      size_t node = (idx >> lv) ^ 1;
      size_t lvl_off = ((size_t)2 << s.log_n) - ((size_t)2 << (s.log_n - lv));
      v = s.ptr[(lvl_off + node) * 4 + (r & 3)];
    } else {
      v = s.ptr[idx * (size_t)s.words + r];
    }
    out[gid] = v;
  }
};

}  // namespace fri
