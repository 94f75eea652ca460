// ORACLE (test infrastructure, NOT product code): the starky prover and verifier behind
// src/starks/common/prover.rs:18-72 and src/starks/common/verifier.rs:32-98.
// The arithmetic lives in the un-vendored dependencies starky 0.4.0 / plonky2 0.2.2
// (InternetMaximalism/polygon-plonky2 @ eeb61ca9, reference Cargo.lock:568-570,613-616,
// 802-804); this file restates their published algorithms:
//   starky/src/prover.rs            prove_with_commitment, compute_quotient_polys
//   starky/src/lookup.rs            lookup_helper_columns, eval_packed_lookups_generic
//   starky/src/cross_table_lookup.rs get_ctl_data, partial_sums, eval_cross_table_lookup_checks
//   starky/src/proof.rs             StarkOpeningSet::{new,to_fri_openings}
//   starky/src/stark.rs             fri_instance
//   starky/src/{get_challenges,verifier}.rs
//   plonky2/src/fri/{oracle,prover,verifier,reduction_strategies,challenges}.rs
// PARITY UNPINNED against the Rust crate itself: the reference holds no golden vectors and
// cannot be built offline (no cargo). Pinned instead by: Poseidon KATs, verifier acceptance,
// and a slow Python model of the primitives (tests/).
#pragma once
#include "constraints.hpp"
#include "merkle.hpp"
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>

namespace orc {

struct StarkConfig {
  unsigned rate_bits = 1, cap_height = 4, num_challenges = 2, num_query_rounds = 84, pow_bits = 16, arity_bits = 4,
           final_poly_bits = 5;
};

static inline std::vector<unsigned> fri_arities(const StarkConfig& c, unsigned degree_bits) {
  // reduction_strategies.rs: ConstantArityBits(arity_bits, final_poly_bits)
  std::vector<unsigned> r;
  while (degree_bits > c.final_poly_bits && degree_bits + c.rate_bits >= c.cap_height + c.arity_bits) {
    r.push_back(c.arity_bits);
    degree_bits -= c.arity_bits;
  }
  return r;
}

// ---- lookup / CTL descriptors (g1/scalar_mul_stark.rs:493-500, g1/scalar_mul_ctl.rs:20-55,
//      fields/exp_ctl.rs:18-51) ---------------------------------------------------------------
struct LinCol {
  std::vector<std::pair<int, u64>> terms;  // sum coef * column
};
struct CtlDesc {
  std::vector<LinCol> cols;
  int filter_col;
};
static inline std::vector<CtlDesc> ctls_for(const Layout& l) {
  auto single = [](int c) { return LinCol{{{c, 1}}}; };
  std::vector<LinCol> s_limbs;
  for (int k = 0; k < 16; k++) {
    LinCol lc;
    for (int i = 0; i < 16; i++) lc.terms.push_back({l.bits + 16 * k + i, (u64)1 << i});  // Column::le_bits
    s_limbs.push_back(lc);
  }
  CtlDesc in, out;
  for (int i = 0; i < l.L; i++) in.cols.push_back(single(l.b + i));  // x
  if (l.kind != KIND_FQ)
    for (int i = 0; i < l.L; i++) in.cols.push_back(single(l.a + i));  // offset
  for (auto& c : s_limbs) in.cols.push_back(c);
  in.cols.push_back(single(l.ts));
  in.filter_col = l.rf + 0;  // is_first_round
  for (int i = 0; i < l.L; i++) out.cols.push_back(single(l.reg1 + i));  // sum / product
  out.cols.push_back(single(l.ts));
  out.filter_col = l.rf + 1;  // is_last_round
  return {in, out};
}
static inline int num_lookup_helpers(const Layout& l) { return (l.rc_hi - l.rc_lo + 1) / 2 + 1; }  // per challenge

template <class T>
static inline T eval_lincol(const LinCol& c, const T* row) {
  T acc;
  for (auto& t : c.terms) acc = acc + row[t.first] * T::from_u64(t.second);
  return acc;
}

// ---- proof object (field order: starky proof.rs / SURVEY C.7) -------------------------------
struct QueryStep {
  std::vector<Fp2> evals;
  std::vector<Hash4> siblings;
};
struct InitialTreeProof {
  std::vector<u64> leaf;
  std::vector<Hash4> siblings;
};
struct QueryRound {
  std::vector<InitialTreeProof> init;  // trace, aux, quotient
  std::vector<QueryStep> steps;
};
struct Proof {
  int kind = 0;
  unsigned degree_bits = 0;
  StarkConfig cfg;
  u64 init_challenger_state[12];
  std::vector<Hash4> trace_cap, aux_cap, quotient_cap;
  std::vector<Fp2> local_values, next_values, aux_polys, aux_polys_next;
  std::vector<u64> ctl_zs_first;
  std::vector<Fp2> quotient_polys;
  std::vector<std::vector<Hash4>> commit_caps;
  std::vector<QueryRound> queries;
  std::vector<Fp2> final_poly;
  u64 pow_witness = 0;
};

// intermediate artefacts kept for parity tests
struct ProveDebug {
  std::vector<u64> ctl_betas, ctl_gammas, alphas;
  Fp2 zeta, fri_alpha;
  std::vector<Fp2> fri_betas;
  std::vector<std::vector<u64>> aux_values;       // A columns x n
  std::vector<std::vector<u64>> quotient_chunks;  // 2*num_challenges x n (coefficients)
  std::vector<u64> query_indices;
};

static const u64 PROOF_MAGIC = 0x31465250343532ULL | ((u64)'B' << 56);  // "254PRF1" + 'B'

static inline std::vector<u64> serialize_proof(const Proof& p) {
  std::vector<u64> o;
  auto put_hashes = [&](const std::vector<Hash4>& v) {
    for (auto& h : v)
      for (int i = 0; i < 4; i++) o.push_back(h.e[i]);
  };
  auto put_ext = [&](const std::vector<Fp2>& v) {
    for (auto& e : v) {
      o.push_back(e.c[0]);
      o.push_back(e.c[1]);
    }
  };
  o.push_back(PROOF_MAGIC);
  o.push_back((u64)p.kind);
  o.push_back(p.degree_bits);
  o.push_back(p.cfg.rate_bits);
  o.push_back(p.cfg.cap_height);
  o.push_back(p.cfg.num_challenges);
  o.push_back(p.cfg.num_query_rounds);
  o.push_back(p.cfg.pow_bits);
  o.push_back(p.cfg.arity_bits);
  o.push_back(p.cfg.final_poly_bits);
  for (int i = 0; i < 12; i++) o.push_back(p.init_challenger_state[i]);
  put_hashes(p.trace_cap);
  put_hashes(p.aux_cap);
  put_hashes(p.quotient_cap);
  put_ext(p.local_values);
  put_ext(p.next_values);
  put_ext(p.aux_polys);
  put_ext(p.aux_polys_next);
  for (u64 x : p.ctl_zs_first) o.push_back(x);
  put_ext(p.quotient_polys);
  for (auto& c : p.commit_caps) put_hashes(c);
  for (auto& q : p.queries) {
    for (auto& it : q.init) {
      for (u64 x : it.leaf) o.push_back(x);
      put_hashes(it.siblings);
    }
    for (auto& st : q.steps) {
      put_ext(st.evals);
      put_hashes(st.siblings);
    }
  }
  put_ext(p.final_poly);
  o.push_back(p.pow_witness);
  return o;
}

static inline Proof deserialize_proof(const u64* w, size_t n_words) {
  size_t pos = 0;
  auto need = [&](size_t k) {
    if (pos + k > n_words) throw OracleError(E_BAD_ARG, "proof blob truncated");
  };
  auto get = [&]() {
    need(1);
    return w[pos++];
  };
  auto get_hashes = [&](size_t k) {
    std::vector<Hash4> v(k);
    need(4 * k);
    for (auto& h : v)
      for (int i = 0; i < 4; i++) h.e[i] = w[pos++];
    return v;
  };
  auto get_ext = [&](size_t k) {
    std::vector<Fp2> v(k);
    need(2 * k);
    for (auto& e : v) {
      e.c[0] = w[pos++];
      e.c[1] = w[pos++];
    }
    return v;
  };
  Proof p;
  if (get() != PROOF_MAGIC) throw OracleError(E_BAD_ARG, "bad proof magic");
  p.kind = (int)get();
  if (p.kind < 0 || p.kind > 2) throw OracleError(E_BAD_ARG, "bad kind");
  p.degree_bits = (unsigned)get();
  p.cfg.rate_bits = (unsigned)get();
  p.cfg.cap_height = (unsigned)get();
  p.cfg.num_challenges = (unsigned)get();
  p.cfg.num_query_rounds = (unsigned)get();
  p.cfg.pow_bits = (unsigned)get();
  p.cfg.arity_bits = (unsigned)get();
  p.cfg.final_poly_bits = (unsigned)get();
  if (p.degree_bits > 30 || p.cfg.rate_bits > 8 || p.cfg.cap_height > 16 || p.cfg.num_challenges > 8 ||
      p.cfg.num_query_rounds > 1024 || p.cfg.arity_bits == 0 || p.cfg.arity_bits > 8)
    throw OracleError(E_BAD_ARG, "bad proof header");
  for (int i = 0; i < 12; i++) p.init_challenger_state[i] = get();
  Layout l = layout_for(p.kind);
  size_t W = l.width, nch = p.cfg.num_challenges;
  size_t A = (size_t)num_lookup_helpers(l) * nch + 2 * nch, Q = 2 * nch;
  size_t cap = (size_t)1 << p.cfg.cap_height;
  p.trace_cap = get_hashes(cap);
  p.aux_cap = get_hashes(cap);
  p.quotient_cap = get_hashes(cap);
  p.local_values = get_ext(W);
  p.next_values = get_ext(W);
  p.aux_polys = get_ext(A);
  p.aux_polys_next = get_ext(A);
  for (size_t i = 0; i < 2 * nch; i++) p.ctl_zs_first.push_back(get());
  p.quotient_polys = get_ext(Q);
  std::vector<unsigned> ar = fri_arities(p.cfg, p.degree_bits);
  for (size_t i = 0; i < ar.size(); i++) p.commit_caps.push_back(get_hashes(cap));
  unsigned lde_bits = p.degree_bits + p.cfg.rate_bits;
  size_t widths[3] = {W, A, Q};
  for (unsigned q = 0; q < p.cfg.num_query_rounds; q++) {
    QueryRound qr;
    for (int t = 0; t < 3; t++) {
      InitialTreeProof it;
      need(widths[t]);
      it.leaf.assign(w + pos, w + pos + widths[t]);
      pos += widths[t];
      it.siblings = get_hashes(lde_bits - p.cfg.cap_height);
      qr.init.push_back(it);
    }
    unsigned bits = lde_bits;
    for (unsigned a : ar) {
      QueryStep st;
      st.evals = get_ext((size_t)1 << a);
      bits -= a;
      st.siblings = get_hashes(bits - p.cfg.cap_height);
      qr.steps.push_back(st);
    }
    p.queries.push_back(qr);
  }
  unsigned fin_bits = p.degree_bits;
  for (unsigned a : ar) fin_bits -= a;
  p.final_poly = get_ext((size_t)1 << fin_bits);
  p.pow_witness = get();
  if (pos != n_words) throw OracleError(E_BAD_ARG, "trailing words in proof blob");
  return p;
}

// ---- lookups (starky lookup.rs) --------------------------------------------------------------
// For one challenge beta: 225 (G1) helper columns h_k = 1/(f_2k + beta) + 1/(f_2k+1 + beta),
// then Z with Z[0] = 0, Z[i+1] = Z[i] + sum_k h_k[i] - freq[i]/(table[i] + beta).
static inline void lookup_columns(const Layout& l, const std::vector<std::vector<u64>>& trace, u64 beta,
                                  std::vector<std::vector<u64>>& out) {
  size_t n = trace[0].size();
  int ncols = l.rc_hi - l.rc_lo;
  int nh = (ncols + 1) / 2;
  size_t base = out.size();
  out.resize(base + nh + 1);
#pragma omp parallel for schedule(dynamic, 2)
  for (int k = 0; k < nh; k++) {
    std::vector<u64> acc(n, 0);
    for (int t = 0; t < 2 && 2 * k + t < ncols; t++) {
      const std::vector<u64>& col = trace[l.rc_lo + 2 * k + t];
      std::vector<u64> d(n);
      for (size_t i = 0; i < n; i++) d[i] = gl_add(col[i], beta);
      std::vector<u64> inv = gl_batch_inv(d);
      for (size_t i = 0; i < n; i++) acc[i] = gl_add(acc[i], inv[i]);
    }
    out[base + k] = std::move(acc);
  }
  std::vector<u64> tb(n);
  for (size_t i = 0; i < n; i++) tb[i] = gl_add(trace[l.range_counter][i], beta);
  std::vector<u64> tinv = gl_batch_inv(tb);
  std::vector<u64> z(n);
  z[0] = 0;
  for (size_t i = 0; i + 1 < n; i++) {
    u64 x = 0;
    for (int k = 0; k < nh; k++) x = gl_add(x, out[base + k][i]);
    x = gl_sub(x, gl_mul(trace[l.freq][i], tinv[i]));
    z[i + 1] = gl_add(z[i], x);
  }
  out[base + nh] = std::move(z);
}

// ---- CTL Z column (cross_table_lookup.rs partial_sums, single looked table, no helpers) ------
static inline std::vector<u64> ctl_z_column(const CtlDesc& d, const std::vector<std::vector<u64>>& trace, u64 beta,
                                            u64 gamma) {
  size_t n = trace[0].size();
  std::vector<u64> comb(n);
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < n; i++) {
    // GrandProductChallenge::combine = reduce_with_powers(values, beta) + gamma
    u64 acc = 0;
    for (size_t k = d.cols.size(); k-- > 0;) {
      u64 v = 0;
      for (auto& t : d.cols[k].terms) v = gl_add(v, gl_mul(trace[t.first][i], t.second));
      acc = gl_add(gl_mul(acc, beta), v);
    }
    comb[i] = gl_add(acc, gamma);
  }
  std::vector<u64> inv = gl_batch_inv(comb);
  std::vector<u64> z(n);
  u64 run = 0;
  for (size_t i = n; i-- > 0;) {
    if (trace[d.filter_col][i] != 0) run = gl_add(run, inv[i]);
    z[i] = run;
  }
  return z;
}

// ---- lookup + CTL constraint checks, generic over T (lookup.rs eval_packed_lookups_generic,
//      cross_table_lookup.rs eval_cross_table_lookup_checks) -----------------------------------
template <class T>
static void eval_lookups_and_ctls(const Layout& l, const std::vector<CtlDesc>& ctls, const T* local, const T* next,
                                  const T* aux_local, const T* aux_next, const std::vector<u64>& betas,
                                  const std::vector<u64>& gammas, Consumer<T>& y) {
  int ncols = l.rc_hi - l.rc_lo, nh = (ncols + 1) / 2, per = nh + 1;
  size_t nch = betas.size();
  for (size_t j = 0; j < nch; j++) {
    T ch = T::from_u64(betas[j]);
    const T* h = aux_local + j * per;
    for (int k = 0; k < nh; k++) {
      if (2 * k + 1 < ncols) {
        T c0 = local[l.rc_lo + 2 * k] + ch, c1 = local[l.rc_lo + 2 * k + 1] + ch;
        y.constraint(c1 * c0 * h[k] - c1 - c0);
      } else {
        T c0 = local[l.rc_lo + 2 * k] + ch;
        y.constraint(c0 * h[k] - T::from_u64(1));
      }
    }
    T z = h[nh], nz = aux_next[j * per + nh];
    T table = local[l.range_counter] + ch;
    T hs;
    for (int k = 0; k < nh; k++) hs = hs + h[k];
    T yv = hs * table - local[l.freq];
    y.constraint_first_row(z);
    y.constraint((nz - z) * table - yv);
  }
  size_t zbase = nch * per;
  for (size_t c = 0; c < ctls.size(); c++)
    for (size_t j = 0; j < nch; j++) {
      T beta = T::from_u64(betas[j]), gamma = T::from_u64(gammas[j]);
      T comb;
      for (size_t k = ctls[c].cols.size(); k-- > 0;) comb = comb * beta + eval_lincol(ctls[c].cols[k], local);
      comb = comb + gamma;
      T f = local[ctls[c].filter_col];
      T lz = aux_local[zbase + c * nch + j], nz = aux_next[zbase + c * nch + j];
      y.constraint_last_row(comb * lz - f);
      y.constraint_transition(comb * (lz - nz) - f);
    }
}

struct StageTimer {
  bool on;
  std::chrono::steady_clock::time_point t;
  StageTimer() : on(getenv("ORC_TIMING") != nullptr), t(std::chrono::steady_clock::now()) {}
  void lap(const char* what) {
    auto now = std::chrono::steady_clock::now();
    if (on) fprintf(stderr, "[oracle] %-28s %8.3f s\n", what, std::chrono::duration<double>(now - t).count());
    t = now;
  }
};

// ---- prover ----------------------------------------------------------------------------------
static inline Proof prove(int kind, const std::vector<std::vector<u64>>& trace, const StarkConfig& cfg,
                          ProveDebug* dbg = nullptr) {
  Layout l = layout_for(kind);
  if ((int)trace.size() != l.width) throw OracleError(E_BAD_ARG, "trace width mismatch");
  const size_t n = trace[0].size();
  const unsigned degree_bits = log2_strict(n);
  const unsigned rate_bits = cfg.rate_bits;
  const size_t nch = cfg.num_challenges;
  std::vector<unsigned> arities = fri_arities(cfg, degree_bits);
  std::vector<CtlDesc> ctls = ctls_for(l);

  Proof pf;
  pf.kind = kind;
  pf.degree_bits = degree_bits;
  pf.cfg = cfg;

  // common/prover.rs:31-44
  StageTimer tm;
  PolynomialBatch trace_c = PolynomialBatch::from_values(trace, rate_bits, cfg.cap_height);
  tm.lap("trace commit");
  Challenger ch;
  ch.observe_cap(trace_c.tree.cap());
  // get_ctl_data (:46-52): (beta, gamma) per challenge, then Z columns CTL-major
  std::vector<u64> betas(nch), gammas(nch);
  for (size_t j = 0; j < nch; j++) {
    betas[j] = ch.get_challenge();
    gammas[j] = ch.get_challenge();
  }
  std::vector<std::vector<u64>> ctl_zs;
  for (auto& d : ctls)
    for (size_t j = 0; j < nch; j++) ctl_zs.push_back(ctl_z_column(d, trace, betas[j], gammas[j]));
  ch.compact(pf.init_challenger_state);  // :54
  tm.lap("ctl z columns");

  // prove_with_commitment: lookup helper columns with challenges = the CTL betas
  std::vector<std::vector<u64>> aux;
  for (size_t j = 0; j < nch; j++) lookup_columns(l, trace, betas[j], aux);
  const size_t num_lookup_cols = aux.size();
  tm.lap("lookup columns");
  for (auto& z : ctl_zs) aux.push_back(z);
  const size_t A = aux.size();
  PolynomialBatch aux_c = PolynomialBatch::from_values(aux, rate_bits, cfg.cap_height);
  tm.lap("aux commit");
  ch.observe_cap(aux_c.tree.cap());
  std::vector<u64> alphas(nch);
  for (size_t j = 0; j < nch; j++) alphas[j] = ch.get_challenge();

  // compute_quotient_polys
  const unsigned qdb = 1;  // log2_ceil(quotient_degree_factor = constraint_degree - 1 = 2)
  if (qdb > rate_bits) throw OracleError(E_BAD_ARG, "rate_bits < quotient_degree_bits");
  const size_t step = (size_t)1 << (rate_bits - qdb), next_step = (size_t)1 << qdb;
  const size_t size = n << qdb;
  std::vector<u64> lag_first(n, 0), lag_last(n, 0);
  lag_first[0] = 1;
  lag_last[n - 1] = 1;
  lag_first = lde_onto_coset(lag_first, qdb);
  lag_last = lde_onto_coset(lag_last, qdb);
  // ZeroPolyOnCoset: Z_H(x_i) = 7^n * w_q^(i mod 2^qdb) - 1
  std::vector<u64> zh_inv((size_t)1 << qdb);
  {
    u64 g_pow_n = GL_COSET_SHIFT;
    for (unsigned i = 0; i < degree_bits; i++) g_pow_n = gl_mul(g_pow_n, g_pow_n);
    u64 wq = gl_root_of_unity(qdb), x = 1;
    for (size_t i = 0; i < zh_inv.size(); i++) {
      zh_inv[i] = gl_inv(gl_sub(gl_mul(g_pow_n, x), 1));
      x = gl_mul(x, wq);
    }
  }
  const u64 last = gl_inv(gl_root_of_unity(degree_bits));
  std::vector<u64> coset(size);
  {
    u64 w = gl_root_of_unity(degree_bits + qdb), x = GL_COSET_SHIFT;
    for (size_t i = 0; i < size; i++) {
      coset[i] = x;
      x = gl_mul(x, w);
    }
  }
  std::vector<std::vector<u64>> qvals(nch, std::vector<u64>(size));
  std::vector<Fp> alphas_t;
  for (u64 a : alphas) alphas_t.push_back(Fp(a));
  size_t n_constraints = 0;
#pragma omp parallel for schedule(dynamic, 64)
  for (size_t i = 0; i < size; i++) {
    size_t i_next = (i + next_step) % size;
    Consumer<Fp> y(alphas_t, Fp(gl_sub(coset[i], last)), Fp(lag_first[i]), Fp(lag_last[i]));
    const Fp* loc = reinterpret_cast<const Fp*>(trace_c.lde_row(i, step));
    const Fp* nxt = reinterpret_cast<const Fp*>(trace_c.lde_row(i_next, step));
    const Fp* aloc = reinterpret_cast<const Fp*>(aux_c.lde_row(i, step));
    const Fp* anxt = reinterpret_cast<const Fp*>(aux_c.lde_row(i_next, step));
    eval_stark(l, loc, nxt, y);
    eval_lookups_and_ctls(l, ctls, loc, nxt, aloc, anxt, betas, gammas, y);
    for (size_t j = 0; j < nch; j++) qvals[j][i] = gl_mul(y.accs[j].v, zh_inv[i % zh_inv.size()]);
    if (i == 0) n_constraints = y.count;
  }
  (void)n_constraints;
  tm.lap("quotient eval");
  std::vector<std::vector<u64>> chunks;
  for (size_t j = 0; j < nch; j++) {
    coset_ifft(qvals[j], GL_COSET_SHIFT);
    for (size_t c = 0; c < ((size_t)1 << qdb); c++)
      chunks.emplace_back(qvals[j].begin() + c * n, qvals[j].begin() + (c + 1) * n);
  }
  if (dbg) dbg->quotient_chunks = chunks;
  PolynomialBatch quot_c = PolynomialBatch::from_coeffs(chunks, rate_bits, cfg.cap_height);
  tm.lap("quotient ifft+commit");
  ch.observe_cap(quot_c.tree.cap());
  Fp2 zeta = ch.get_ext_challenge();
  const u64 g = gl_root_of_unity(degree_bits);
  if (zeta.exp_power_of_2(degree_bits) == Fp2(1, 0)) throw OracleError(E_INTERNAL, "opening point is in the subgroup");

  // StarkOpeningSet::new
  Fp2 zeta_next = zeta.scalar_mul(g);
  auto eval_all = [&](const PolynomialBatch& b, Fp2 z) {
    std::vector<Fp2> r(b.width());
#pragma omp parallel for schedule(dynamic, 4)
    for (size_t c = 0; c < b.width(); c++) r[c] = poly_eval_ext(b.polynomials[c], z);
    return r;
  };
  pf.local_values = eval_all(trace_c, zeta);
  pf.next_values = eval_all(trace_c, zeta_next);
  pf.aux_polys = eval_all(aux_c, zeta);
  pf.aux_polys_next = eval_all(aux_c, zeta_next);
  for (size_t c = num_lookup_cols; c < A; c++) pf.ctl_zs_first.push_back(poly_eval_base(aux_c.polynomials[c], 1));
  pf.quotient_polys = eval_all(quot_c, zeta);
  // observe_openings(to_fri_openings): [local|aux|quotient], [next|aux_next], [ctl_zs_first]
  for (auto& v : pf.local_values) ch.observe_ext(v);
  for (auto& v : pf.aux_polys) ch.observe_ext(v);
  for (auto& v : pf.quotient_polys) ch.observe_ext(v);
  for (auto& v : pf.next_values) ch.observe_ext(v);
  for (auto& v : pf.aux_polys_next) ch.observe_ext(v);
  for (u64 v : pf.ctl_zs_first) ch.observe_ext(Fp2(v, 0));

  tm.lap("openings");
  // PolynomialBatch::prove_openings
  Fp2 fri_alpha = ch.get_ext_challenge();
  struct Batch {
    Fp2 point;
    std::vector<const std::vector<u64>*> polys;
  };
  std::vector<Batch> batches(3);
  batches[0].point = zeta;
  for (auto& p : trace_c.polynomials) batches[0].polys.push_back(&p);
  for (auto& p : aux_c.polynomials) batches[0].polys.push_back(&p);
  for (auto& p : quot_c.polynomials) batches[0].polys.push_back(&p);
  batches[1].point = zeta_next;
  for (auto& p : trace_c.polynomials) batches[1].polys.push_back(&p);
  for (auto& p : aux_c.polynomials) batches[1].polys.push_back(&p);
  batches[2].point = Fp2(1, 0);
  for (size_t c = num_lookup_cols; c < A; c++) batches[2].polys.push_back(&aux_c.polynomials[c]);
  std::vector<Fp2> final_poly(n);
  for (auto& b : batches) {
    // composition = sum_j alpha^j f_j
    std::vector<Fp2> comp(n);
    std::vector<Fp2> pw(b.polys.size());
    Fp2 p(1, 0);
    for (size_t j = 0; j < pw.size(); j++) {
      pw[j] = p;
      p = p * fri_alpha;
    }
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) {
      Fp2 acc;
      for (size_t j = 0; j < b.polys.size(); j++) acc = acc + pw[j].scalar_mul((*b.polys[j])[i]);
      comp[i] = acc;
    }
    // divide_by_linear(point): synthetic division, drop the remainder, pad with a zero
    std::vector<Fp2> quot(n);
    Fp2 acc;
    for (size_t i = n; i-- > 0;) {
      acc = acc * b.point + comp[i];
      if (i > 0) quot[i - 1] = acc;
    }
    quot[n - 1] = Fp2();
    // final = final * alpha^{|batch|} + quot     (ReducingFactor::shift_poly)
    Fp2 sh = fri_alpha.pow(b.polys.size());
    for (size_t i = 0; i < n; i++) final_poly[i] = final_poly[i] * sh + quot[i];
  }
  const size_t N = n << rate_bits;
  std::vector<Fp2> coeffs = final_poly;
  coeffs.resize(N);
  std::vector<Fp2> values = coeffs;
  coset_fft_ext(values, GL_COSET_SHIFT);

  tm.lap("fri combine + lde");
  // fri_committed_trees
  std::vector<MerkleTree> fri_trees;
  std::vector<Fp2> fri_betas;
  u64 shift = GL_COSET_SHIFT;
  for (unsigned ab : arities) {
    size_t arity = (size_t)1 << ab;
    size_t len = values.size();
    unsigned lg = log2_strict(len);
    MerkleTree t;
    t.num_leaves = len / arity;
    t.leaf_len = 2 * arity;
    t.leaves.resize(2 * len);
    for (size_t i = 0; i < len; i++) {  // reverse_index_bits_in_place then flatten chunks
      size_t r = reverse_bits(i, lg);
      t.leaves[2 * r] = values[i].c[0];
      t.leaves[2 * r + 1] = values[i].c[1];
    }
    t.build(cfg.cap_height);
    ch.observe_cap(t.cap());
    pf.commit_caps.push_back(t.cap());
    fri_trees.push_back(std::move(t));
    Fp2 beta = ch.get_ext_challenge();
    fri_betas.push_back(beta);
    std::vector<Fp2> folded(coeffs.size() / arity);
    for (size_t i = 0; i < folded.size(); i++) {
      Fp2 acc;
      for (size_t k = arity; k-- > 0;) acc = acc * beta + coeffs[i * arity + k];
      folded[i] = acc;
    }
    coeffs = std::move(folded);
    shift = gl_pow(shift, arity);
    values = coeffs;
    coset_fft_ext(values, shift);
  }
  coeffs.resize(coeffs.size() >> rate_bits);
  pf.final_poly = coeffs;
  for (auto& c : coeffs) ch.observe_ext(c);

  tm.lap("fri commit phase");
  // fri_proof_of_work: minimal witness (the reference's rayon find_any is non-deterministic)
  {
    u64 base_state[12];
    memcpy(base_state, ch.state, sizeof base_state);
    size_t pos = ch.in_buf.size();
    for (size_t i = 0; i < pos; i++) base_state[i] = ch.in_buf[i];
    u64 found = 0;
    bool ok = false;
    for (u64 start = 0; !ok; start += 1 << 16) {
      u64 best = ~(u64)0;
#pragma omp parallel for schedule(static) reduction(min : best)
      for (u64 w = start; w < start + (1 << 16); w++) {
        u64 s[12];
        memcpy(s, base_state, sizeof s);
        s[pos] = w;
        poseidon_permute(s);
        if (cfg.pow_bits == 0 || (s[7] >> (64 - cfg.pow_bits)) == 0)
          if (w < best) best = w;
      }
      if (best != ~(u64)0) {
        found = best;
        ok = true;
      }
    }
    pf.pow_witness = found;
    ch.observe_element(found);
    u64 resp = ch.get_challenge();
    if (cfg.pow_bits && (resp >> (64 - cfg.pow_bits)) != 0) throw OracleError(E_INTERNAL, "pow check failed");
  }

  tm.lap("pow");
  // fri_prover_query_rounds
  const PolynomialBatch* init[3] = {&trace_c, &aux_c, &quot_c};
  for (unsigned q = 0; q < cfg.num_query_rounds; q++) {
    size_t x_index = (size_t)(ch.get_challenge() % N);
    if (dbg) dbg->query_indices.push_back(x_index);
    QueryRound qr;
    for (int t = 0; t < 3; t++) {
      InitialTreeProof it;
      const MerkleTree& tr = init[t]->tree;
      it.leaf.assign(tr.leaf(x_index), tr.leaf(x_index) + tr.leaf_len);
      it.siblings = tr.prove(x_index);
      qr.init.push_back(std::move(it));
    }
    for (size_t i = 0; i < fri_trees.size(); i++) {
      unsigned ab = arities[i];
      const MerkleTree& tr = fri_trees[i];
      size_t li = x_index >> ab;
      QueryStep st;
      for (size_t k = 0; k < tr.leaf_len / 2; k++) st.evals.push_back(Fp2(tr.leaf(li)[2 * k], tr.leaf(li)[2 * k + 1]));
      st.siblings = tr.prove(li);
      qr.steps.push_back(std::move(st));
      x_index >>= ab;
    }
    pf.queries.push_back(std::move(qr));
  }
  tm.lap("queries");
  pf.trace_cap = trace_c.tree.cap();
  pf.aux_cap = aux_c.tree.cap();
  pf.quotient_cap = quot_c.tree.cap();
  if (dbg) {
    dbg->ctl_betas = betas;
    dbg->ctl_gammas = gammas;
    dbg->alphas = alphas;
    dbg->zeta = zeta;
    dbg->fri_alpha = fri_alpha;
    dbg->fri_betas = fri_betas;
    dbg->aux_values = aux;
  }
  return pf;
}

// ---- verifier (common/verifier.rs:32-98 + starky verifier.rs + plonky2 fri/verifier.rs) ------
// extra_looking[c] = list of public tuples for CTL c (g1/scalar_mul_ctl.rs:57-80).
static inline void verify(const Proof& pf, const std::vector<std::vector<std::vector<u64>>>& extra_looking) {
  Layout l = layout_for(pf.kind);
  const StarkConfig& cfg = pf.cfg;
  const size_t nch = cfg.num_challenges;
  const unsigned degree_bits = pf.degree_bits;
  std::vector<CtlDesc> ctls = ctls_for(l);
  std::vector<unsigned> arities = fri_arities(cfg, degree_bits);
  auto fail = [](const char* m) { throw OracleError(E_INTERNAL, std::string("verify: ") + m); };
  const size_t W = l.width, per = num_lookup_helpers(l), num_lookup_cols = per * nch, A = num_lookup_cols + 2 * nch,
               Q = 2 * nch;
  if (pf.local_values.size() != W || pf.next_values.size() != W || pf.aux_polys.size() != A ||
      pf.aux_polys_next.size() != A || pf.quotient_polys.size() != Q || pf.ctl_zs_first.size() != 2 * nch)
    fail("opening set shape");
  if (pf.commit_caps.size() != arities.size() || pf.queries.size() != cfg.num_query_rounds) fail("fri shape");

  Challenger ch;
  ch.observe_cap(pf.trace_cap);
  std::vector<u64> betas(nch), gammas(nch);
  for (size_t j = 0; j < nch; j++) {
    betas[j] = ch.get_challenge();
    gammas[j] = ch.get_challenge();
  }
  u64 st[12];
  ch.compact(st);
  if (memcmp(st, pf.init_challenger_state, sizeof st)) fail("init_challenger_state mismatch");
  // get_challenges(ignore_trace_cap = true)
  ch.observe_cap(pf.aux_cap);
  std::vector<u64> alphas(nch);
  for (size_t j = 0; j < nch; j++) alphas[j] = ch.get_challenge();
  ch.observe_cap(pf.quotient_cap);
  Fp2 zeta = ch.get_ext_challenge();
  for (auto& v : pf.local_values) ch.observe_ext(v);
  for (auto& v : pf.aux_polys) ch.observe_ext(v);
  for (auto& v : pf.quotient_polys) ch.observe_ext(v);
  for (auto& v : pf.next_values) ch.observe_ext(v);
  for (auto& v : pf.aux_polys_next) ch.observe_ext(v);
  for (u64 v : pf.ctl_zs_first) ch.observe_ext(Fp2(v, 0));
  Fp2 fri_alpha = ch.get_ext_challenge();
  std::vector<Fp2> fri_betas;
  for (auto& cap : pf.commit_caps) {
    ch.observe_cap(cap);
    fri_betas.push_back(ch.get_ext_challenge());
  }
  for (auto& c : pf.final_poly) ch.observe_ext(c);
  ch.observe_element(pf.pow_witness);
  u64 pow_resp = ch.get_challenge();
  const size_t N = (size_t)1 << (degree_bits + cfg.rate_bits);
  std::vector<size_t> q_idx(cfg.num_query_rounds);
  for (auto& q : q_idx) q = (size_t)(ch.get_challenge() % N);

  // constraint check at zeta
  const u64 g = gl_root_of_unity(degree_bits);
  Fp2 one(1, 0);
  Fp2 zeta_pow_n = zeta.exp_power_of_2(degree_bits);
  Fp2 z_h = zeta_pow_n - one;
  u64 nf = ((u64)1 << degree_bits) % GL_P;
  Fp2 l_first = z_h * ((zeta - one).scalar_mul(nf)).inv();
  Fp2 l_last = z_h * ((zeta.scalar_mul(g) - one).scalar_mul(nf)).inv();
  Fp2 z_last = zeta - Fp2(gl_inv(g), 0);
  std::vector<Fp2> al;
  for (u64 a : alphas) al.push_back(Fp2(a, 0));
  Consumer<Fp2> y(al, z_last, l_first, l_last);
  eval_stark(l, pf.local_values.data(), pf.next_values.data(), y);
  eval_lookups_and_ctls(l, ctls, pf.local_values.data(), pf.next_values.data(), pf.aux_polys.data(),
                        pf.aux_polys_next.data(), betas, gammas, y);
  for (size_t j = 0; j < nch; j++) {
    Fp2 t = pf.quotient_polys[2 * j] + pf.quotient_polys[2 * j + 1] * zeta_pow_n;  // reduce_with_powers(chunk, zeta^n)
    if (y.accs[j] != z_h * t) fail("quotient identity at zeta");
  }

  // FRI (plonky2 fri/verifier.rs)
  if (cfg.pow_bits && (pow_resp >> (64 - cfg.pow_bits)) != 0) fail("proof of work");
  unsigned fin_bits = degree_bits;
  for (unsigned a : arities) fin_bits -= a;
  if (pf.final_poly.size() != ((size_t)1 << fin_bits)) fail("final poly length");
  Fp2 zeta_next = zeta.scalar_mul(g);
  // precomputed reduced openings: sum_j alpha^j opening_j per batch
  auto reduce = [&](const std::vector<Fp2>& v) {
    Fp2 acc;
    for (size_t i = v.size(); i-- > 0;) acc = acc * fri_alpha + v[i];
    return acc;
  };
  std::vector<Fp2> b0 = pf.local_values, b1 = pf.next_values, b2;
  b0.insert(b0.end(), pf.aux_polys.begin(), pf.aux_polys.end());
  b0.insert(b0.end(), pf.quotient_polys.begin(), pf.quotient_polys.end());
  b1.insert(b1.end(), pf.aux_polys_next.begin(), pf.aux_polys_next.end());
  for (u64 v : pf.ctl_zs_first) b2.push_back(Fp2(v, 0));
  Fp2 red_open[3] = {reduce(b0), reduce(b1), reduce(b2)};
  Fp2 points[3] = {zeta, zeta_next, one};
  const unsigned lde_bits = degree_bits + cfg.rate_bits;
  const std::vector<Hash4>* caps[3] = {&pf.trace_cap, &pf.aux_cap, &pf.quotient_cap};
  size_t widths[3] = {W, A, Q};
  for (size_t qi = 0; qi < q_idx.size(); qi++) {
    const QueryRound& qr = pf.queries[qi];
    size_t x_index = q_idx[qi];
    if (qr.init.size() != 3 || qr.steps.size() != arities.size()) fail("query shape");
    for (int t = 0; t < 3; t++) {
      if (qr.init[t].leaf.size() != widths[t]) fail("initial leaf width");
      if (!merkle_verify(qr.init[t].leaf.data(), widths[t], x_index, *caps[t], qr.init[t].siblings))
        fail("initial Merkle proof");
    }
    // subgroup_x = 7 * w^(rev(x_index))
    u64 sx = gl_mul(GL_COSET_SHIFT, gl_pow(gl_root_of_unity(lde_bits), reverse_bits(x_index, lde_bits)));
    Fp2 subgroup_x(sx, 0);
    // fri_combine_initial
    std::vector<Fp2> e0, e1, e2;
    for (u64 v : qr.init[0].leaf) e0.push_back(Fp2(v, 0));
    e1 = e0;
    for (u64 v : qr.init[1].leaf) {
      e0.push_back(Fp2(v, 0));
      e1.push_back(Fp2(v, 0));
    }
    for (u64 v : qr.init[2].leaf) e0.push_back(Fp2(v, 0));
    for (size_t c = num_lookup_cols; c < A; c++) e2.push_back(Fp2(qr.init[1].leaf[c], 0));
    std::vector<Fp2>* ev[3] = {&e0, &e1, &e2};
    Fp2 sum;
    for (int b = 0; b < 3; b++) {
      Fp2 red = reduce(*ev[b]);
      Fp2 num = red - red_open[b];
      Fp2 den = subgroup_x - points[b];
      sum = sum * fri_alpha.pow(ev[b]->size()) + num * den.inv();
    }
    Fp2 old_eval = sum;
    unsigned bits = lde_bits;
    for (size_t i = 0; i < arities.size(); i++) {
      unsigned ab = arities[i];
      size_t arity = (size_t)1 << ab;
      const QueryStep& stp = qr.steps[i];
      if (stp.evals.size() != arity) fail("step evals");
      size_t x_in_coset = x_index & (arity - 1);
      size_t coset_index = x_index >> ab;
      if (stp.evals[x_in_coset] != old_eval) fail("FRI consistency");
      // compute_evaluation: interpolate the arity points and evaluate at beta
      u64 gg = gl_root_of_unity(ab);
      size_t rev_x = reverse_bits(x_in_coset, ab);
      u64 coset_start = gl_mul(sx, gl_pow(gl_inv(gg), rev_x));
      std::vector<Fp2> evs(arity);
      for (size_t k = 0; k < arity; k++) evs[reverse_bits(k, ab)] = stp.evals[k];  // reverse_index_bits
      // Lagrange interpolation over points coset_start * gg^k
      Fp2 beta = fri_betas[i];
      Fp2 res;
      for (size_t k = 0; k < arity; k++) {
        u64 xk = gl_mul(coset_start, gl_pow(gg, k));
        Fp2 numr(1, 0);
        u64 den = 1;
        for (size_t m = 0; m < arity; m++) {
          if (m == k) continue;
          u64 xm = gl_mul(coset_start, gl_pow(gg, m));
          numr = numr * (beta - Fp2(xm, 0));
          den = gl_mul(den, gl_sub(xk, xm));
        }
        res = res + evs[k] * numr.scalar_mul(gl_inv(den));
      }
      old_eval = res;
      std::vector<u64> flat;
      for (auto& e : stp.evals) {
        flat.push_back(e.c[0]);
        flat.push_back(e.c[1]);
      }
      if (!merkle_verify(flat.data(), flat.size(), coset_index, pf.commit_caps[i], stp.siblings))
        fail("FRI layer Merkle proof");
      sx = gl_pow(sx, arity);
      x_index = coset_index;
      bits -= ab;
    }
    (void)bits;
    // final poly evaluation at subgroup_x
    Fp2 acc;
    for (size_t i = pf.final_poly.size(); i-- > 0;) acc = acc * Fp2(sx, 0) + pf.final_poly[i];
    if (acc != old_eval) fail("final polynomial evaluation");
  }

  // CTL: Z(1) == sum over public tuples of 1/combine (common/ctl_values.rs:28-47,
  // verify_cross_table_lookups with no looking tables)
  if (extra_looking.size() != ctls.size()) fail("extra looking values shape");
  for (size_t c = 0; c < ctls.size(); c++)
    for (size_t j = 0; j < nch; j++) {
      u64 sum = 0;
      for (auto& tuple : extra_looking[c]) {
        if (tuple.size() != ctls[c].cols.size()) fail("extra looking tuple width");
        u64 acc = 0;
        for (size_t k = tuple.size(); k-- > 0;) acc = gl_add(gl_mul(acc, betas[j]), tuple[k]);
        sum = gl_add(sum, gl_inv(gl_add(acc, gammas[j])));
      }
      if (sum != pf.ctl_zs_first[c * nch + j]) fail("cross-table lookup sum");
    }
}

}  // namespace orc
