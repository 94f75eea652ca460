// Field map of the serialized StarkProofWithMetadata (the blob prover.cuh writes): the one place that knows where a
// field lives; used by the verifier and exported as pb254_proof_parse for the reference-side consumer
// (set_stark_proof_target, src/generators/g1/stark_proof.rs:173-178, which walks the StarkProof struct field by
// field; field order of starky 0.4.0 proof.rs / plonky2 0.2.2 fri/proof.rs). Host only.
#pragma once
#include "../../include/pb254.h"
#include "tracegen.h"
#include "aux.cuh"
#include <vector>

namespace proofview {

static const u64 MAGIC = 0x31465250343532ULL | ((u64)'B' << 56);

static inline std::vector<unsigned> fri_arities(const pb254_config& c, unsigned degree_bits) {
  std::vector<unsigned> r;
  while (degree_bits > c.final_poly_bits && degree_bits + c.rate_bits >= c.cap_height + c.arity_bits) {
    r.push_back(c.arity_bits);
    degree_bits -= c.arity_bits;
  }
  return r;
}

// Throws Pb254Error(PB254_E_BAD_ARG) with the reason when the words are not a well-formed blob.
static inline void parse(const u64* blob, size_t words, pb254_proof_layout& o) {
  auto bad = [](const char* m) { throw Pb254Error(PB254_E_BAD_ARG, std::string("proof blob: ") + m); };
  if (!blob || words < 22 || blob[0] != MAGIC) bad("not a pb254 proof");
  memset(&o, 0, sizeof o);
  if (blob[1] > 2) bad("unknown STARK kind");
  if (blob[2] < 8 || blob[2] > 26) bad("degree_bits out of range");
  for (int i = 0; i < 7; i++)
    if (blob[3 + i] > 4096) bad("StarkConfig field out of range");
  o.kind = (uint32_t)blob[1];
  o.degree_bits = (uint32_t)blob[2];
  o.config.rate_bits = (uint32_t)blob[3];
  o.config.cap_height = (uint32_t)blob[4];
  o.config.num_challenges = (uint32_t)blob[5];
  o.config.num_query_rounds = (uint32_t)blob[6];
  o.config.pow_bits = (uint32_t)blob[7];
  o.config.arity_bits = (uint32_t)blob[8];
  o.config.final_poly_bits = (uint32_t)blob[9];
  const pb254_config& c = o.config;
  if (c.rate_bits < 1 || c.rate_bits > 3 || c.num_challenges < 1 || c.num_challenges > (unsigned)aux::MAXCH || c.arity_bits < 1 ||
      c.arity_bits > 4 || c.cap_height > 16 || c.final_poly_bits > 16 || c.num_query_rounds < 1)
    bad("StarkConfig field out of range");
  const tg::Layout l = tg::layout_for((int)o.kind);
  const int nch = (int)c.num_challenges, logN = (int)o.degree_bits + (int)c.rate_bits, cap_h = (int)c.cap_height;
  if (cap_h > logN) bad("cap height");
  o.trace_width = (uint32_t)l.width;
  o.aux_width = (uint32_t)aux::num_aux(l, nch);
  o.quotient_width = 2 * c.num_challenges;
  o.num_ctl_zs = 2 * c.num_challenges;
  const std::vector<unsigned> ar = fri_arities(c, o.degree_bits);
  if (ar.size() > PB254_MAX_FRI_LAYERS) bad("too many FRI layers");
  o.num_fri_layers = (uint32_t)ar.size();
  for (size_t i = 0; i < ar.size(); i++) o.fri_arity_bits[i] = ar[i];
  o.cap_words = (u64)4 << cap_h;
  const u64 W = o.trace_width, A = o.aux_width, Q = o.quotient_width;
  u64 pos = 10;
  auto take = [&](u64 n) {
    const u64 at = pos;
    pos += n;
    return at;
  };
  o.init_challenger_state = take(12);
  o.trace_cap = take(o.cap_words);
  o.auxiliary_polys_cap = take(o.cap_words);
  o.quotient_polys_cap = take(o.cap_words);
  o.local_values = take(2 * W);
  o.next_values = take(2 * W);
  o.auxiliary_polys = take(2 * A);
  o.auxiliary_polys_next = take(2 * A);
  o.ctl_zs_first = take(o.num_ctl_zs);
  o.quotient_polys = take(2 * Q);
  o.commit_phase_merkle_caps = take(ar.size() * o.cap_words);
  // one query record
  o.initial_path_words = (uint32_t)(4 * (logN - cap_h));
  uint32_t q = 0;
  auto qtake = [&](uint32_t n) {
    const uint32_t at = q;
    q += n;
    return at;
  };
  o.q_trace_leaf = qtake((uint32_t)W);
  o.q_trace_path = qtake(o.initial_path_words);
  o.q_aux_leaf = qtake((uint32_t)A);
  o.q_aux_path = qtake(o.initial_path_words);
  o.q_quotient_leaf = qtake((uint32_t)Q);
  o.q_quotient_path = qtake(o.initial_path_words);
  int ll = logN;
  for (size_t i = 0; i < ar.size(); i++) {
    ll -= (int)ar[i];
    if (ll < cap_h) bad("FRI layer smaller than the cap");
    o.q_step_evals_words[i] = 2u << ar[i];
    o.q_step_evals[i] = qtake(o.q_step_evals_words[i]);
    o.q_step_path_words[i] = (uint32_t)(4 * (ll - cap_h));
    o.q_step_path[i] = qtake(o.q_step_path_words[i]);
  }
  o.query_words = q;
  o.query_round_proofs = take((u64)c.num_query_rounds * q);
  int log_final = (int)o.degree_bits;
  for (unsigned ab : ar) log_final -= (int)ab;
  o.final_poly_words = (u64)2 << log_final;
  o.final_poly = take(o.final_poly_words);
  o.pow_witness = take(1);
  o.words = pos;
  if (pos != words) bad("length does not match the header");
}

}  // namespace proofview
