// ORACLE (test infrastructure, NOT product code): C entry points used by tests/, smoke() and
// bench.py's cpu_baseline / --impl reference legs through ctypes. Nothing in the product
// (plonky2_bn254_b200/) links or loads this library.
#include "stark.hpp"
#include <chrono>
#include <cstdio>
#include <omp.h>

using namespace orc;

namespace {
thread_local std::string g_err;
int fail(const OracleError& e) {
  g_err = e.what();
  return e.code ? e.code : E_INTERNAL;
}

// native (trace-independent) result s*x + offset / x^s, as the reference computes it with
// arkworks in g1_generate_ctl_values (g1/scalar_mul_ctl.rs:57-80)
struct OptG1 {
  bool inf;
  G1Pt p;
};
OptG1 g1_add_native(const OptG1& a, const OptG1& b) {
  if (a.inf) return b;
  if (b.inf) return a;
  U256 lambda;
  if (a.p.x == b.p.x) {
    if (!(a.p.y == b.p.y) || u256_is_zero(a.p.y)) return {true, {}};
    lambda = fq_mul(fq_mul(fq_from_u64(3), fq_mul(a.p.x, a.p.x)), fq_inv(fq_add(a.p.y, a.p.y)));
  } else {
    lambda = fq_mul(fq_sub(b.p.y, a.p.y), fq_inv(fq_sub(b.p.x, a.p.x)));
  }
  G1Pt c;
  c.x = fq_sub(fq_sub(fq_mul(lambda, lambda), a.p.x), b.p.x);
  c.y = fq_sub(fq_mul(lambda, fq_sub(a.p.x, c.x)), a.p.y);
  return {false, c};
}
struct OptG2 {
  bool inf;
  G2Pt p;
};
OptG2 g2_add_native(const OptG2& a, const OptG2& b) {
  if (a.inf) return b;
  if (b.inf) return a;
  Fq2 lambda;
  if (a.p.x == b.p.x) {
    if (!(a.p.y == b.p.y) || fq2_is_zero(a.p.y)) return {true, {}};
    lambda = fq2_mul(fq2_mul(fq2_from_u64(3), fq2_mul(a.p.x, a.p.x)), fq2_inv(fq2_add(a.p.y, a.p.y)));
  } else {
    lambda = fq2_mul(fq2_sub(b.p.y, a.p.y), fq2_inv(fq2_sub(b.p.x, a.p.x)));
  }
  G2Pt c;
  c.x = fq2_sub(fq2_sub(fq2_mul(lambda, lambda), a.p.x), b.p.x);
  c.y = fq2_sub(fq2_mul(lambda, fq2_sub(a.p.x, c.x)), a.p.y);
  return {false, c};
}

void native_result(int kind, const u64* w, u64* out /* L limbs */) {
  U256 s = read_u256(w, false);
  if (kind == KIND_G1) {
    OptG1 acc = {true, {}}, base = {false, {read_u256(w + 4, true), read_u256(w + 8, true)}};
    for (int i = 255; i >= 0; i--) {
      acc = g1_add_native(acc, acc);
      if (u256_bit(s, i)) acc = g1_add_native(acc, base);
    }
    OptG1 off = {false, {read_u256(w + 12, true), read_u256(w + 16, true)}};
    acc = g1_add_native(acc, off);
    if (acc.inf) throw OracleError(E_INFINITY, "native result is the point at infinity");
    RegG1::put(out, acc.p);
  } else if (kind == KIND_G2) {
    OptG2 acc = {true, {}};
    OptG2 base = {false, {{read_u256(w + 4, true), read_u256(w + 8, true)}, {read_u256(w + 12, true), read_u256(w + 16, true)}}};
    for (int i = 255; i >= 0; i--) {
      acc = g2_add_native(acc, acc);
      if (u256_bit(s, i)) acc = g2_add_native(acc, base);
    }
    OptG2 off = {false, {{read_u256(w + 20, true), read_u256(w + 24, true)}, {read_u256(w + 28, true), read_u256(w + 32, true)}}};
    acc = g2_add_native(acc, off);
    if (acc.inf) throw OracleError(E_INFINITY, "native result is the point at infinity");
    RegG2::put(out, acc.p);
  } else {
    U256 r = fq_pow(read_u256(w + 4, true), s);
    RegFq::put(out, r);
  }
}

// extra looking values (g1/scalar_mul_ctl.rs:57-80, fields/exp_ctl.rs:53-75)
std::vector<std::vector<std::vector<u64>>> ctl_values(int kind, const u64* inputs, const u64* ts, size_t k) {
  Layout l = layout_for(kind);
  std::vector<std::vector<std::vector<u64>>> e(2);
  e[0].resize(k);
  e[1].resize(k);
#pragma omp parallel for schedule(dynamic, 1)
  for (size_t i = 0; i < k; i++) {
    const u64* w = inputs + i * l.in_words;
    std::vector<u64> in, out(l.L);
    int ncoord = (l.in_words - 4) / 4;  // 4 (G1), 8 (G2), 1 (Fq)
    int per_pt = kind == KIND_FQ ? 1 : ncoord / 2;
    // x limbs, then offset limbs (none for Fq), then s limbs, then timestamp
    for (int c = 0; c < ncoord; c++) {
      (void)per_pt;
      u64 tmp[16];
      put_u256(tmp, read_u256(w + 4 + 4 * c, true));
      in.insert(in.end(), tmp, tmp + 16);
    }
    u64 sl[16];
    put_u256(sl, read_u256(w, false));
    in.insert(in.end(), sl, sl + 16);
    in.push_back(ts[i]);
    native_result(kind, w, out.data());
    out.push_back(ts[i]);
    e[0][i] = in;
    e[1][i] = out;
  }
  return e;
}

struct ProofHandle {
  Proof proof;
  ProveDebug dbg;
  std::vector<u64> blob;
};
}  // namespace

extern "C" {

const char* orc_last_error() { return g_err.c_str(); }
int orc_num_threads() { return omp_get_max_threads(); }
void orc_set_num_threads(int n) { omp_set_num_threads(n); }

void orc_poseidon_round_constants(u64* out360) { memcpy(out360, poseidon_consts().rc, sizeof(u64) * 360); }
void orc_poseidon_permute(u64* state12) { poseidon_permute(state12); }
void orc_hash_no_pad(const u64* in, size_t n, u64* out4) {
  Hash4 h = hash_no_pad(in, n);
  memcpy(out4, h.e, sizeof h.e);
}
void orc_two_to_one(const u64* l, const u64* r, u64* out4) {
  Hash4 a, b;
  memcpy(a.e, l, 32);
  memcpy(b.e, r, 32);
  Hash4 h = two_to_one(a, b);
  memcpy(out4, h.e, 32);
}
u64 orc_gl_mul(u64 a, u64 b) { return gl_mul(a, b); }
u64 orc_gl_inv(u64 a) { return gl_inv(a); }
u64 orc_gl_root_of_unity(unsigned k) { return gl_root_of_unity(k); }

void orc_fft(u64* a, size_t n) {
  std::vector<u64> v(a, a + n);
  fft(v);
  memcpy(a, v.data(), n * 8);
}
void orc_ifft(u64* a, size_t n) {
  std::vector<u64> v(a, a + n);
  ifft(v);
  memcpy(a, v.data(), n * 8);
}
// column-major cols x n values -> coefficients (cols x n) and natural-order LDE (cols x n<<rate_bits)
void orc_lde_batch(const u64* values, size_t cols, size_t n, unsigned rate_bits, u64* coeffs_out, u64* lde_out) {
#pragma omp parallel for schedule(dynamic, 1)
  for (size_t c = 0; c < cols; c++) {
    std::vector<u64> v(values + c * n, values + (c + 1) * n);
    ifft(v);
    if (coeffs_out) memcpy(coeffs_out + c * n, v.data(), n * 8);
    v.resize(n << rate_bits, 0);
    coset_fft(v, GL_COSET_SHIFT);
    if (lde_out) memcpy(lde_out + c * (n << rate_bits), v.data(), (n << rate_bits) * 8);
  }
}
// PolynomialBatch::from_values on a column-major matrix: cap (2^cap_height x 4) and, optionally,
// all digest levels concatenated bottom-up (leaf digests first).
int orc_commit(const u64* values, size_t cols, size_t n, unsigned rate_bits, unsigned cap_height, int from_coeffs,
               u64* cap_out, u64* digests_out) {
  try {
    std::vector<std::vector<u64>> v(cols);
    for (size_t c = 0; c < cols; c++) v[c].assign(values + c * n, values + (c + 1) * n);
    PolynomialBatch b = from_coeffs ? PolynomialBatch::from_coeffs(v, rate_bits, cap_height)
                                    : PolynomialBatch::from_values(v, rate_bits, cap_height);
    memcpy(cap_out, b.tree.cap().data(), b.tree.cap().size() * 32);
    if (digests_out) {
      size_t off = 0;
      for (auto& lv : b.tree.levels) {
        memcpy(digests_out + off, lv.data(), lv.size() * 32);
        off += lv.size() * 4;
      }
    }
    return 0;
  } catch (OracleError& e) {
    return fail(e);
  }
}
// Same commitment as orc_commit (from_values), evaluated in column chunks of SPONGE_RATE so that only one chunk of
// the LDE and one sponge state per row are alive: the cap of a matrix whose LDE does not fit in host memory
// (BASELINE config 4: 427 columns x 2^24 LDE rows). hash_no_pad absorbs 8 elements per permutation, so the state of
// row j after chunk c equals the state of the one-shot hash after the same 8 (c + 1) elements.
int orc_commit_streamed(const u64* values, size_t cols, size_t n, unsigned rate_bits, unsigned cap_height, u64* cap_out) {
  try {
    if (cols <= 4) throw OracleError(E_INTERNAL, "streamed commit: rows of <= 4 elements are not hashed");
    const size_t N = n << rate_bits;
    const unsigned lg = log2_strict(N);
    std::vector<u64> state(N * 12, 0);
    // one chunk = a multiple of SPONGE_RATE columns, at least one column per thread for the LDE phase
    const size_t per = (size_t)SPONGE_RATE * (((size_t)omp_get_max_threads() + SPONGE_RATE - 1) / SPONGE_RATE);
    for (size_t c0 = 0; c0 < cols; c0 += per) {
      const size_t len = cols - c0 < per ? cols - c0 : per;
      std::vector<std::vector<u64>> lde(len);
#pragma omp parallel for schedule(dynamic, 1)
      for (size_t k = 0; k < len; k++)
        lde[k] = lde_onto_coset(std::vector<u64>(values + (c0 + k) * n, values + (c0 + k + 1) * n), rate_bits);
#pragma omp parallel for schedule(static)
      for (size_t j = 0; j < N; j++) {
        u64* s = &state[j * 12];
        for (size_t k0 = 0; k0 < len; k0 += SPONGE_RATE) {
          const size_t m = len - k0 < (size_t)SPONGE_RATE ? len - k0 : (size_t)SPONGE_RATE;
          for (size_t k = 0; k < m; k++) s[k] = lde[k0 + k][j];
          poseidon_permute(s);
        }
      }
    }
    std::vector<Hash4> level(N);
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < N; i++) memcpy(level[i].e, &state[reverse_bits(i, lg) * 12], 32);
    std::vector<u64>().swap(state);
    for (unsigned l = 0; l < lg - cap_height; l++) {
      size_t m = level.size() / 2;
      std::vector<Hash4> nxt(m);
#pragma omp parallel for schedule(static)
      for (size_t i = 0; i < m; i++) nxt[i] = two_to_one(level[2 * i], level[2 * i + 1]);
      level.swap(nxt);
    }
    memcpy(cap_out, level.data(), level.size() * 32);
    return 0;
  } catch (OracleError& e) {
    return fail(e);
  }
}
// Merkle tree over given row-major leaves
int orc_merkle(const u64* leaves, size_t num_leaves, size_t leaf_len, unsigned cap_height, u64* cap_out) {
  MerkleTree t;
  t.num_leaves = num_leaves;
  t.leaf_len = leaf_len;
  t.leaves.assign(leaves, leaves + num_leaves * leaf_len);
  t.build(cap_height);
  memcpy(cap_out, t.cap().data(), t.cap().size() * 32);
  return 0;
}

int orc_width(int kind) { return layout_for(kind).width; }
int orc_in_words(int kind) { return layout_for(kind).in_words; }
int orc_reg_len(int kind) { return layout_for(kind).L; }
int orc_num_aux(int kind, unsigned num_challenges) {
  return (num_lookup_helpers(layout_for(kind)) + 2) * (int)num_challenges;
}
size_t orc_trace_rows(size_t n_inputs, size_t min_rows) {
  size_t n = min_rows > n_inputs * PERIOD ? min_rows : n_inputs * PERIOD, r = 1;
  while (r < n) r <<= 1;
  return r;
}

// modular witnesses, exposed for the Python big-int cross-check
int orc_gen_modulus_zero(const int64_t* input31, u64* out80) {
  try {
    gen_modulus_zero(input31, out80);
    return 0;
  } catch (OracleError& e) {
    return fail(e);
  }
}
int orc_gen_is_modulus_zero(const int64_t* input16, u64* out96, int* is_zero) {
  try {
    *is_zero = gen_is_modulus_zero(input16, out96);
    return 0;
  } catch (OracleError& e) {
    return fail(e);
  }
}
int orc_native_result(int kind, const u64* input, u64* out_limbs) {
  try {
    native_result(kind, input, out_limbs);
    return 0;
  } catch (OracleError& e) {
    return fail(e);
  }
}

// generate_trace: cols_out is width x orc_trace_rows(...) column-major; results (optional) n_inputs x L
int orc_generate_trace(int kind, const u64* inputs, const u64* timestamps, size_t n_inputs, size_t min_rows,
                       u64* cols_out, u64* results_out) {
  try {
    std::vector<u64> res;
    auto cols = generate_trace(kind, inputs, timestamps, n_inputs, min_rows, results_out ? &res : nullptr);
    size_t n = cols[0].size();
    for (size_t c = 0; c < cols.size(); c++) memcpy(cols_out + c * n, cols[c].data(), n * 8);
    if (results_out) memcpy(results_out, res.data(), res.size() * 8);
    return 0;
  } catch (OracleError& e) {
    return fail(e);
  }
}

static StarkConfig cfg_from(const unsigned* c) {
  StarkConfig s;
  if (c) {
    s.rate_bits = c[0];
    s.cap_height = c[1];
    s.num_challenges = c[2];
    s.num_query_rounds = c[3];
    s.pow_bits = c[4];
    s.arity_bits = c[5];
    s.final_poly_bits = c[6];
  }
  return s;
}

// prove from a column-major trace. cfg7 = {rate_bits, cap_height, num_challenges, num_query_rounds,
// pow_bits, arity_bits, final_poly_bits} or NULL for standard_fast_config.
int orc_prove(int kind, const u64* trace_cols, size_t n_rows, const unsigned* cfg7, int keep_debug, void** handle) {
  try {
    Layout l = layout_for(kind);
    std::vector<std::vector<u64>> tr(l.width);
    for (int c = 0; c < l.width; c++) tr[c].assign(trace_cols + (size_t)c * n_rows, trace_cols + (size_t)(c + 1) * n_rows);
    ProofHandle* h = new ProofHandle();
    try {
      h->proof = prove(kind, tr, cfg_from(cfg7), keep_debug ? &h->dbg : nullptr);
    } catch (...) {
      delete h;
      throw;
    }
    h->blob = serialize_proof(h->proof);
    *handle = h;
    return 0;
  } catch (OracleError& e) {
    return fail(e);
  }
}
// trace generation + prove in one call (what run_once does, generators/g1/stark_proof.rs:154-163)
int orc_prove_inputs(int kind, const u64* inputs, const u64* timestamps, size_t n_inputs, size_t min_rows,
                     const unsigned* cfg7, int keep_debug, void** handle, double* t_trace_s, double* t_prove_s) {
  try {
    auto t0 = std::chrono::steady_clock::now();
    auto tr = generate_trace(kind, inputs, timestamps, n_inputs, min_rows);
    auto t1 = std::chrono::steady_clock::now();
    ProofHandle* h = new ProofHandle();
    try {
      h->proof = prove(kind, tr, cfg_from(cfg7), keep_debug ? &h->dbg : nullptr);
    } catch (...) {
      delete h;
      throw;
    }
    auto t2 = std::chrono::steady_clock::now();
    h->blob = serialize_proof(h->proof);
    *handle = h;
    if (t_trace_s) *t_trace_s = std::chrono::duration<double>(t1 - t0).count();
    if (t_prove_s) *t_prove_s = std::chrono::duration<double>(t2 - t1).count();
    return 0;
  } catch (OracleError& e) {
    return fail(e);
  }
}
void orc_proof_free(void* handle) { delete (ProofHandle*)handle; }
size_t orc_proof_words(void* handle) { return ((ProofHandle*)handle)->blob.size(); }
void orc_proof_copy(void* handle, u64* out) {
  auto& b = ((ProofHandle*)handle)->blob;
  memcpy(out, b.data(), b.size() * 8);
}
// debug artefacts: which = 0 aux values (A x n), 1 quotient chunks (2*nch x n), 2 challenges
// [betas(nch), gammas(nch), alphas(nch), zeta(2), fri_alpha(2), fri_betas(2 each)], 3 query indices
size_t orc_proof_debug_words(void* handle, int which) {
  ProveDebug& d = ((ProofHandle*)handle)->dbg;
  if (which == 0) return d.aux_values.empty() ? 0 : d.aux_values.size() * d.aux_values[0].size();
  if (which == 1) return d.quotient_chunks.empty() ? 0 : d.quotient_chunks.size() * d.quotient_chunks[0].size();
  if (which == 2) return d.ctl_betas.size() + d.ctl_gammas.size() + d.alphas.size() + 4 + 2 * d.fri_betas.size();
  if (which == 3) return d.query_indices.size();
  return 0;
}
void orc_proof_debug_copy(void* handle, int which, u64* out) {
  ProveDebug& d = ((ProofHandle*)handle)->dbg;
  size_t o = 0;
  if (which == 0)
    for (auto& c : d.aux_values) {
      memcpy(out + o, c.data(), c.size() * 8);
      o += c.size();
    }
  if (which == 1)
    for (auto& c : d.quotient_chunks) {
      memcpy(out + o, c.data(), c.size() * 8);
      o += c.size();
    }
  if (which == 2) {
    for (u64 x : d.ctl_betas) out[o++] = x;
    for (u64 x : d.ctl_gammas) out[o++] = x;
    for (u64 x : d.alphas) out[o++] = x;
    out[o++] = d.zeta.c[0];
    out[o++] = d.zeta.c[1];
    out[o++] = d.fri_alpha.c[0];
    out[o++] = d.fri_alpha.c[1];
    for (auto& b : d.fri_betas) {
      out[o++] = b.c[0];
      out[o++] = b.c[1];
    }
  }
  if (which == 3)
    for (u64 x : d.query_indices) out[o++] = x;
}

// verify a serialized proof against the public inputs of the batch (extra looking values are
// recomputed natively from the inputs, as run_once does before calling verify)
int orc_verify(const u64* blob, size_t n_words, const u64* inputs, const u64* timestamps, size_t n_inputs) {
  try {
    Proof p = deserialize_proof(blob, n_words);
    auto extra = ctl_values(p.kind, inputs, timestamps, n_inputs);
    verify(p, extra);
    return 0;
  } catch (OracleError& e) {
    return fail(e);
  }
}

// constraint evaluation of one (local, next) row pair over the base field, for kernel parity
// tests: returns the per-alpha accumulators before division by Z_H.
int orc_eval_constraints_base(int kind, const u64* local, const u64* next, const u64* aux_local, const u64* aux_next,
                              const u64* betas, const u64* gammas, const u64* alphas, unsigned nch, u64 z_last,
                              u64 l_first, u64 l_last, u64* acc_out) {
  Layout l = layout_for(kind);
  auto ctls = ctls_for(l);
  std::vector<Fp> al;
  for (unsigned j = 0; j < nch; j++) al.push_back(Fp(alphas[j]));
  Consumer<Fp> y(al, Fp(z_last), Fp(l_first), Fp(l_last));
  eval_stark(l, (const Fp*)local, (const Fp*)next, y);
  std::vector<u64> b(betas, betas + nch), g(gammas, gammas + nch);
  eval_lookups_and_ctls(l, ctls, (const Fp*)local, (const Fp*)next, (const Fp*)aux_local, (const Fp*)aux_next, b, g, y);
  for (unsigned j = 0; j < nch; j++) acc_out[j] = y.accs[j].v;
  return (int)y.count;
}

}  // extern "C"
