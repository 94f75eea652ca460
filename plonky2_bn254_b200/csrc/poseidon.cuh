// Poseidon-Goldilocks permutation (width 12, rate 8, x^7, 4 + 22 + 4 rounds) for device code,
// plus the sponge helpers the commitment scheme uses. Replaces plonky2 0.2.2
// hash/poseidon.rs::Poseidon::poseidon, hashing.rs::{hash_n_to_m_no_pad, compress} (un-vendored
// dependency; reached from src/starks/common/prover.rs:31-38 through PolynomialBatch::from_values).
//
// One permutation per thread, state in registers. The MDS layer uses the fact that the matrix
// entries are < 2^6: lanes are split into 32-bit halves and the two half-sums (< 2^42) are
// recombined and reduced once per lane, so a layer costs 2*12*12 32-bit multiply-adds plus 12
// short reductions instead of 144 full field multiplications.
#pragma once
#include "gl.cuh"

namespace poseidon {

static constexpr int WIDTH = 12, RATE = 8, HALF_FULL = 4, N_PARTIAL = 22, N_ROUNDS = 30;

#if PB_HOSTSIM
static const u64 RC[N_ROUNDS * WIDTH] = {
#include "poseidon_constants.inc"
};
#define PB_RC(i) (poseidon::RC[(i)])
#else
// host copy (transcript) and constant-memory copy (kernels)
static const u64 RC_HOST[N_ROUNDS * WIDTH] = {
#include "poseidon_constants.inc"
};
static __constant__ u64 RC_DEV[N_ROUNDS * WIDTH] = {
#include "poseidon_constants.inc"
};
#ifdef __CUDA_ARCH__
#define PB_RC(i) (poseidon::RC_DEV[(i)])
#else
#define PB_RC(i) (poseidon::RC_HOST[(i)])
#endif
#endif

PB_HD u64 sbox(u64 x) {
  u64 x2 = gl::sqr(x), x4 = gl::sqr(x2), x3 = gl::mul(x, x2);
  return gl::mul(x3, x4);
}

// value = lo + 2^32 * hi with lo, hi < 2^43  ->  canonical residue
PB_HD u64 reduce_split(u64 lo, u64 hi) {
  u64 l = lo + (hi << 32);
  u64 carry = l < lo;
  u64 top = (hi >> 32) + carry;  // < 2^12: value = top * 2^64 + l
  u64 t1 = (top << 32) - top;    // top * (2^32 - 1)
  u64 r = l + t1;
  r += (r < l) ? gl::EPS : 0;
  r -= (r >= gl::P) ? gl::P : 0;
  return r;
}

PB_HD void mds(u64 s[12]) {
  // MDS_MATRIX_CIRC = [17,15,41,16,2,28,13,13,39,18,34,20], MDS_MATRIX_DIAG = [8,0,...]
  u64 lo[12], hi[12];
#pragma unroll
  for (int i = 0; i < 12; i++) {
    lo[i] = s[i] & gl::EPS;
    hi[i] = s[i] >> 32;
  }
  constexpr u64 C[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};
#pragma unroll
  for (int r = 0; r < 12; r++) {
    u64 al = 0, ah = 0;
#pragma unroll
    for (int i = 0; i < 12; i++) {
      al += lo[(i + r) % 12] * C[i];
      ah += hi[(i + r) % 12] * C[i];
    }
    if (r == 0) {
      al += lo[0] * 8;
      ah += hi[0] * 8;
    }
    s[r] = reduce_split(al, ah);
  }
}

PB_HD void permute(u64 s[12]) {
#pragma unroll 1
  for (int r = 0; r < HALF_FULL; r++) {
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = sbox(gl::add(s[i], PB_RC(r * 12 + i)));
    mds(s);
  }
#pragma unroll 1
  for (int r = HALF_FULL; r < HALF_FULL + N_PARTIAL; r++) {
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = gl::add(s[i], PB_RC(r * 12 + i));
    s[0] = sbox(s[0]);
    mds(s);
  }
#pragma unroll 1
  for (int r = HALF_FULL + N_PARTIAL; r < N_ROUNDS; r++) {
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = sbox(gl::add(s[i], PB_RC(r * 12 + i)));
    mds(s);
  }
}

struct Digest {
  u64 e[4];
};

// compress / two_to_one: permute(l | r | 0^4)[0..4]
PB_HD Digest two_to_one(const Digest& l, const Digest& r) {
  u64 s[12];
#pragma unroll
  for (int i = 0; i < 4; i++) {
    s[i] = l.e[i];
    s[4 + i] = r.e[i];
    s[8 + i] = 0;
  }
  permute(s);
  Digest d;
#pragma unroll
  for (int i = 0; i < 4; i++) d.e[i] = s[i];
  return d;
}

}  // namespace poseidon
