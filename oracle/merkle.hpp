// ORACLE (test infrastructure, NOT product code): Poseidon Merkle tree with cap and the
// PolynomialBatch commitment, as in plonky2 0.2.2 plonky2/src/hash/merkle_tree.rs and
// plonky2/src/fri/oracle.rs (un-vendored; call site common/prover.rs:31-38).
//   leaves[rev(j)] = LDE row j;  digest(leaf) = hash_or_noop(row);  parent = two_to_one(l, r)
//   cap = the level with 2^cap_height nodes;  prove(i) = siblings bottom-up, below the cap.
#pragma once
#include "poseidon.hpp"
#include "ntt.hpp"

namespace orc {

struct MerkleTree {
  size_t num_leaves = 0, leaf_len = 0;
  std::vector<u64> leaves;                 // row-major num_leaves x leaf_len
  std::vector<std::vector<Hash4>> levels;  // levels[0] = leaf digests ... levels.back() = cap
  const u64* leaf(size_t i) const { return &leaves[i * leaf_len]; }
  const std::vector<Hash4>& cap() const { return levels.back(); }

  void build(unsigned cap_height) {
    unsigned lg = log2_strict(num_leaves);
    assert(cap_height <= lg);
    levels.clear();
    levels.emplace_back(num_leaves);
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < num_leaves; i++) levels[0][i] = hash_or_noop(leaf(i), leaf_len);
    for (unsigned l = 0; l < lg - cap_height; l++) {
      size_t m = levels[l].size() / 2;
      std::vector<Hash4> nxt(m);
#pragma omp parallel for schedule(static)
      for (size_t i = 0; i < m; i++) nxt[i] = two_to_one(levels[l][2 * i], levels[l][2 * i + 1]);
      levels.push_back(std::move(nxt));
    }
  }
  std::vector<Hash4> prove(size_t idx) const {
    std::vector<Hash4> sib;
    for (size_t l = 0; l + 1 < levels.size(); l++) sib.push_back(levels[l][(idx >> l) ^ 1]);
    return sib;
  }
};

// merkle_proofs.rs: verify_merkle_proof_to_cap
static inline bool merkle_verify(const u64* leaf, size_t leaf_len, size_t idx, const std::vector<Hash4>& cap,
                                 const std::vector<Hash4>& siblings) {
  Hash4 cur = hash_or_noop(leaf, leaf_len);
  for (auto& s : siblings) {
    cur = (idx & 1) ? two_to_one(s, cur) : two_to_one(cur, s);
    idx >>= 1;
  }
  return idx < cap.size() && cur == cap[idx];
}

struct PolynomialBatch {
  size_t degree = 0;
  unsigned rate_bits = 0;
  std::vector<std::vector<u64>> polynomials;  // coefficient form, each of length `degree`
  MerkleTree tree;

  size_t lde_size() const { return degree << rate_bits; }
  size_t width() const { return polynomials.size(); }
  // get_lde_values(index, step): row (index*step) of the natural-order LDE
  const u64* lde_row(size_t index, size_t step) const {
    return tree.leaf(reverse_bits(index * step, log2_strict(lde_size())));
  }

  // from_coeffs: zero-pad, coset FFT (shift 7), transpose, bit-reverse rows, Merkle
  static PolynomialBatch from_coeffs(std::vector<std::vector<u64>> coeffs, unsigned rate_bits, unsigned cap_height) {
    PolynomialBatch b;
    b.degree = coeffs[0].size();
    b.rate_bits = rate_bits;
    size_t W = coeffs.size(), N = b.degree << rate_bits;
    unsigned lg = log2_strict(N);
    b.tree.num_leaves = N;
    b.tree.leaf_len = W;
    b.tree.leaves.assign(N * W, 0);
    std::vector<std::vector<u64>> lde(W);
#pragma omp parallel for schedule(dynamic, 4)
    for (size_t c = 0; c < W; c++) {
      lde[c] = coeffs[c];
      lde[c].resize(N, 0);
      coset_fft(lde[c], GL_COSET_SHIFT);
    }
    // transpose to rows and bit-reverse the row index: leaves[i] = LDE row rev(i)
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < N; i++) {
      size_t j = reverse_bits(i, lg);
      u64* row = &b.tree.leaves[i * W];
      for (size_t c = 0; c < W; c++) row[c] = lde[c][j];
    }
    b.polynomials = std::move(coeffs);
    b.tree.build(cap_height);
    return b;
  }
  // from_values: iFFT each column first
  static PolynomialBatch from_values(const std::vector<std::vector<u64>>& values, unsigned rate_bits,
                                     unsigned cap_height) {
    std::vector<std::vector<u64>> coeffs(values.size());
#pragma omp parallel for schedule(dynamic, 4)
    for (size_t c = 0; c < values.size(); c++) {
      coeffs[c] = values[c];
      ifft(coeffs[c]);
    }
    return from_coeffs(std::move(coeffs), rate_bits, cap_height);
  }
};

}  // namespace orc
