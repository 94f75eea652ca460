// ORACLE (test infrastructure, NOT product code): Poseidon-Goldilocks, width 12, as in
// plonky2 0.2.2 plonky2/src/hash/{poseidon.rs,poseidon_goldilocks.rs,hashing.rs} (un-vendored
// dependency, pinned in the reference's Cargo.lock:613-616). Naive round form:
//   for r in 0..30: add RC[12r+i]; S-box x^7 (all lanes in rounds 0..3 and 26..29, lane 0
//   otherwise); MDS out[r] = sum_i s[(i+r)%12]*CIRC[i] + s[r]*DIAG[r].
// Round constants are REGENERATED from the published recipe (ChaCha8Rng::seed_from_u64(0),
// gen_range(0..p)) and pinned by the two upstream known-answer vectors in tests/.
#pragma once
#include "gl.hpp"
#include <cstring>

namespace orc {

static const int SPONGE_WIDTH = 12, SPONGE_RATE = 8, HALF_FULL = 4, N_PARTIAL = 22;
static const int N_ROUNDS = 30;
static const u64 MDS_CIRC[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};
static const u64 MDS_DIAG[12] = {8, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};

struct ChaCha8 {
  uint32_t key[8];
  uint64_t counter = 0;
  uint32_t buf[16];
  int idx = 16;
  static uint32_t rotl(uint32_t x, int k) { return (x << k) | (x >> (32 - k)); }
  explicit ChaCha8(uint64_t seed) {
    // rand_core::SeedableRng::seed_from_u64 — PCG32 expansion of the u64 into a 32-byte key
    uint64_t state = seed;
    for (int i = 0; i < 8; i++) {
      state = state * 6364136223846793005ULL + 11634580027462260723ULL;
      uint32_t xorshifted = (uint32_t)(((state >> 18) ^ state) >> 27);
      uint32_t rot = (uint32_t)(state >> 59);
      key[i] = (xorshifted >> rot) | (xorshifted << ((32 - rot) & 31));
    }
  }
  void block() {
    uint32_t s[16] = {0x61707865, 0x3320646e, 0x79622d32, 0x6b206574};
    for (int i = 0; i < 8; i++) s[4 + i] = key[i];
    s[12] = (uint32_t)counter;
    s[13] = (uint32_t)(counter >> 32);
    s[14] = 0;
    s[15] = 0;
    uint32_t x[16];
    memcpy(x, s, sizeof x);
#define QR(a, b, c, d)                                   \
  x[a] += x[b]; x[d] = rotl(x[d] ^ x[a], 16);            \
  x[c] += x[d]; x[b] = rotl(x[b] ^ x[c], 12);            \
  x[a] += x[b]; x[d] = rotl(x[d] ^ x[a], 8);             \
  x[c] += x[d]; x[b] = rotl(x[b] ^ x[c], 7);
    for (int r = 0; r < 4; r++) {  // 8 rounds = 4 double rounds
      QR(0, 4, 8, 12) QR(1, 5, 9, 13) QR(2, 6, 10, 14) QR(3, 7, 11, 15)
      QR(0, 5, 10, 15) QR(1, 6, 11, 12) QR(2, 7, 8, 13) QR(3, 4, 9, 14)
    }
#undef QR
    for (int i = 0; i < 16; i++) buf[i] = x[i] + s[i];
    counter++;
    idx = 0;
  }
  uint32_t next_u32() {
    if (idx == 16) block();
    return buf[idx++];
  }
  uint64_t next_u64() {
    uint64_t lo = next_u32();
    uint64_t hi = next_u32();
    return lo | (hi << 32);
  }
  // rand 0.8 UniformInt<u64>::sample_single(0, range)
  uint64_t gen_range(uint64_t range) {
    unsigned lz = __builtin_clzll(range);
    uint64_t zone = (range << lz) - 1;
    for (;;) {
      u128 m = (u128)next_u64() * range;
      if ((u64)m <= zone) return (u64)(m >> 64);
    }
  }
};

struct PoseidonConsts {
  u64 rc[N_ROUNDS * SPONGE_WIDTH];
  PoseidonConsts() {
    ChaCha8 rng(0);
    for (int i = 0; i < N_ROUNDS * SPONGE_WIDTH; i++) rc[i] = rng.gen_range(GL_P);
  }
};
static inline const PoseidonConsts& poseidon_consts() {
  static PoseidonConsts c;
  return c;
}

static inline u64 sbox7(u64 x) {
  u64 x2 = gl_mul(x, x), x4 = gl_mul(x2, x2), x3 = gl_mul(x, x2);
  return gl_mul(x3, x4);
}

static inline void mds_layer(u64 s[12]) {
  // entries are < 2^6, so split each lane into 32-bit halves and accumulate in plain u64:
  // sum_lo, sum_hi < 12 * 41 * 2^32 < 2^42; result = sum_lo + 2^32 * sum_hi (< 2^75) reduced once
  u64 lo[24], hi[24];
  for (int i = 0; i < 12; i++) {
    lo[i] = lo[i + 12] = s[i] & GL_EPS;
    hi[i] = hi[i + 12] = s[i] >> 32;
  }
  for (int r = 0; r < 12; r++) {
    u64 al = 0, ah = 0;
#pragma GCC unroll 12
    for (int i = 0; i < 12; i++) {
      al += lo[i + r] * MDS_CIRC[i];
      ah += hi[i + r] * MDS_CIRC[i];
    }
    al += lo[r] * MDS_DIAG[r];
    ah += hi[r] * MDS_DIAG[r];
    s[r] = gl_reduce128((u128)al + ((u128)ah << 32));
  }
}

static inline void poseidon_permute(u64 s[12]) {
  const u64* rc = poseidon_consts().rc;
  for (int r = 0; r < N_ROUNDS; r++) {
    for (int i = 0; i < 12; i++) s[i] = gl_add(s[i], rc[r * 12 + i]);
    if (r < HALF_FULL || r >= HALF_FULL + N_PARTIAL) {
      for (int i = 0; i < 12; i++) s[i] = sbox7(s[i]);
    } else {
      s[0] = sbox7(s[0]);
    }
    mds_layer(s);
  }
}

struct Hash4 {
  u64 e[4];
  bool operator==(const Hash4& o) const { return !memcmp(e, o.e, sizeof e); }
};

// hashing.rs: hash_n_to_m_no_pad (overwrite-mode sponge, rate 8), 4 outputs
static inline Hash4 hash_no_pad(const u64* in, size_t n) {
  u64 s[12] = {0};
  for (size_t off = 0; off < n; off += SPONGE_RATE) {
    size_t len = n - off < (size_t)SPONGE_RATE ? n - off : (size_t)SPONGE_RATE;
    for (size_t i = 0; i < len; i++) s[i] = in[off + i];
    poseidon_permute(s);
  }
  Hash4 h;
  for (int i = 0; i < 4; i++) h.e[i] = s[i];
  return h;
}
// config.rs: Hasher::hash_or_noop — inputs of <= 4 elements are zero-padded, not hashed
static inline Hash4 hash_or_noop(const u64* in, size_t n) {
  if (n <= 4) {
    Hash4 h = {{0, 0, 0, 0}};
    for (size_t i = 0; i < n; i++) h.e[i] = in[i];
    return h;
  }
  return hash_no_pad(in, n);
}
// hashing.rs: compress (two_to_one)
static inline Hash4 two_to_one(const Hash4& l, const Hash4& r) {
  u64 s[12] = {0};
  for (int i = 0; i < 4; i++) {
    s[i] = l.e[i];
    s[4 + i] = r.e[i];
  }
  poseidon_permute(s);
  Hash4 h;
  for (int i = 0; i < 4; i++) h.e[i] = s[i];
  return h;
}

// iop/challenger.rs: Challenger<F, PoseidonHash> (duplex sponge, overwrite mode)
struct Challenger {
  u64 state[12];
  std::vector<u64> in_buf, out_buf;
  Challenger() { memset(state, 0, sizeof state); }
  void duplexing() {
    assert(in_buf.size() <= (size_t)SPONGE_RATE);
    for (size_t i = 0; i < in_buf.size(); i++) state[i] = in_buf[i];
    in_buf.clear();
    poseidon_permute(state);
    out_buf.assign(state, state + SPONGE_RATE);
  }
  void observe_element(u64 x) {
    out_buf.clear();
    in_buf.push_back(x);
    if (in_buf.size() == (size_t)SPONGE_RATE) duplexing();
  }
  void observe_elements(const u64* x, size_t n) {
    for (size_t i = 0; i < n; i++) observe_element(x[i]);
  }
  void observe_hash(const Hash4& h) { observe_elements(h.e, 4); }
  void observe_cap(const std::vector<Hash4>& cap) {
    for (auto& h : cap) observe_hash(h);
  }
  void observe_ext(const Fp2& x) {
    observe_element(x.c[0]);
    observe_element(x.c[1]);
  }
  u64 get_challenge() {
    if (!in_buf.empty() || out_buf.empty()) duplexing();
    u64 r = out_buf.back();
    out_buf.pop_back();
    return r;
  }
  Fp2 get_ext_challenge() {
    u64 a = get_challenge();
    u64 b = get_challenge();
    return Fp2(a, b);
  }
  // compact(): flush pending inputs, drop outputs, return the sponge state
  void compact(u64 out[12]) {
    if (!in_buf.empty()) duplexing();
    out_buf.clear();
    memcpy(out, state, sizeof state);
  }
};

}  // namespace orc
