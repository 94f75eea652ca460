// Poseidon-Goldilocks permutation (width 12, rate 8, x^7, 4 + 22 + 4 rounds) for device code,
// plus the sponge helpers the commitment scheme uses. Replaces plonky2 0.2.2
// hash/poseidon.rs::Poseidon::poseidon, hashing.rs::{hash_n_to_m_no_pad, compress} (un-vendored
// dependency; reached from src/starks/common/prover.rs:31-38 through PolynomialBatch::from_values).
//
// One permutation per thread, state in registers. Two forms with identical results:
//   * permute_generic  canonical arithmetic everywhere (host transcript, hostsim build);
//   * permute (device) "lazy" arithmetic tuned for the B200 integer pipes (see below); inputs must be
//     canonical, outputs are canonical.
//
// Device form. The kernel is bound by instruction issue on the FMA (IMAD) and ALU pipes, so the design
// minimises instructions per permutation (ncu: 47 k thread-instructions with canonical arithmetic) and
// keeps the code small enough for the instruction cache:
//   - lanes live as arbitrary u64 representatives of their residue (no conditional subtractions);
//   - a field multiplication is 4 IMAD.WIDE.U32 plus a carry-chain reduction written in PTX
//     (2^64 = 2^32 - 1, 2^96 = -1 mod p): add.cc / addc, no compare-and-select;
//   - the MDS layer uses that all matrix entries are < 2^6: lanes are cut into 16-bit pieces and two
//     (piece x entry) products are accumulated per dp2a instruction (IDP.2A, 288 per layer); with
//     IMAD.WIDE on 32-bit halves ptxas emits one IMAD.WIDE plus one 64-bit carry-chain addition per
//     product, and those IADD3/IADD3.X were 60 % of all stall samples (profiles/r1_poseidon_compact.txt);
//   - the optimised-partial-round factorisation of the reference implementation is NOT used: with
//     6-bit matrix entries the dense layer (288 IMAD.WIDE) is as cheap on this machine as the sparse
//     layer with 64-bit constants (22 full multiplications) - see DESIGN.md.
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
  constexpr u64 C[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};
#if !defined(__CUDA_ARCH__)
  // host (Fiat-Shamir transcript: ~700 permutations per proof on the critical path): 128-bit accumulation,
  // one reduction per lane
  u64 t[24];
  for (int i = 0; i < 12; i++) t[i] = t[i + 12] = s[i];
  for (int r = 0; r < 12; r++) {
    u128 acc = r == 0 ? (u128)t[0] * 8 : (u128)0;
    for (int i = 0; i < 12; i++) acc += (u128)t[i + r] * C[i];
    s[r] = gl::reduce128((u64)acc, (u64)(acc >> 64));
  }
#else
  u64 lo[12], hi[12];
#pragma unroll
  for (int i = 0; i < 12; i++) {
    lo[i] = s[i] & gl::EPS;
    hi[i] = s[i] >> 32;
  }
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
#endif
}

PB_HD void permute_generic(u64 s[12]) {
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

#if !PB_HOSTSIM
// ---- device form: lazy representatives, PTX carry chains, compact code -----------------------------
// Piece table: constant k of (round r, lane i) as four u32 RC2[(r*12 + i)*2 ..] = {k0, k1, k2, k3}, its 16-bit
// pieces (block r = 30 is all zero), followed by the 12 round-0 constants as plain u64. Kernels that hash a
// lot stage it into shared memory (RC2_WORDS u64s, 16-byte aligned).
static constexpr int RC2_WORDS = (N_ROUNDS + 1) * WIDTH * 2 + WIDTH;
static __constant__ __align__(16) u64 RC2_DEV[RC2_WORDS] = {
#include "poseidon_constants_split.inc"
};

namespace lazy {

// gl::mul_lazy with the final wrap correction w (2^32 - 1), w in {-1, 0, 1}, built on the ALU pipe
// (((w >> 1) << 32) | (u32)(-w)) instead of an IMAD.WIDE: in this kernel the FMA pipe is the bottleneck
// (85 % busy against 64 % for the ALU pipe, profiles/r1_kernels_final.md).
__device__ __forceinline__ u64 mul(u64 a, u64 b) {
  const u32 a0 = (u32)a, a1 = (u32)(a >> 32), b0 = (u32)b, b1 = (u32)(b >> 32);
  const u64 P = gl::mulw32(a0, b0), Q = gl::mulw32(a0, b1), R = gl::mulw32(a1, b0), S = gl::mulw32(a1, b1);
  const unsigned __int128 prod =
      (unsigned __int128)P + (((unsigned __int128)Q + R) << 32) + ((unsigned __int128)S << 64);
  const u64 lo = (u64)prod, hi = (u64)(prod >> 64);
  const u32 x2 = (u32)hi, x3 = (u32)(hi >> 32);
  const __int128 V = (__int128)lo - x3 - x2 + ((__int128)x2 << 32);
  const u64 r = (u64)V;
  const int w = (int)(long long)(V >> 64);
  u32 adj_lo, adj_hi;
  asm("neg.s32 %0, %2;\n\tshr.s32 %1, %2, 1;" : "=r"(adj_lo), "=r"(adj_hi) : "r"(w));
  return r + (((u64)adj_hi << 32) | adj_lo);
}

__device__ __forceinline__ u64 sbox(u64 x) {
  const u64 x2 = mul(x, x), x4 = mul(x2, x2), x3 = mul(x, x2);
  return mul(x3, x4);
}

// s <- MDS s + (next round's constants); rc2 = 24 words of the split table.
// MDS on 16-bit pieces with dp2a: lanes are cut into four 16-bit pieces; pieces of the same weight of two
// neighbouring lanes are packed into one register, and one dp2a adds two (piece x 6-bit entry) products to
// a 32-bit accumulator: 6 dp2a per (output lane, weight) instead of 12 IMAD.WIDE + 12 64-bit additions.
template <int R, int K>
struct Coef {
  static constexpr u32 C[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};
  static constexpr u32 m(int r, int i) { return C[(i - r + 12) % 12] + ((r == 0 && i == 0) ? 8u : 0u); }
  static constexpr u32 value = m(R, 2 * K) | (m(R, 2 * K + 1) << 8);
};
template <int R, int K>
__device__ __forceinline__ void dp_row(u32 acc[4], const u32 (*X)[4]) {
#pragma unroll
  for (int q = 0; q < 4; q++)
    asm("dp2a.lo.u32.u32 %0, %1, %2, %0;" : "+r"(acc[q]) : "r"(X[K][q]), "r"(Coef<R, K>::value));
  if constexpr (K + 1 < 6) dp_row<R, K + 1>(acc, X);
}
template <int R>
__device__ __forceinline__ void mds_rows(u64 s[12], const u32 (*X)[4], const u64* rc2) {
  // the next round's constant enters as the initial value of the four piece accumulators
  const uint4 k = *reinterpret_cast<const uint4*>(rc2 + 2 * R);
  u32 acc[4] = {k.x, k.y, k.z, k.w};
  dp_row<R, 0>(acc, X);
  // value = acc0 + acc1 2^16 + acc2 2^32 + acc3 2^48 (acc < 2^25) -> u64 representative. The layer is bound by
  // the FMA pipe (288 IDP), so this part uses the ALU pipe only: byte permutes for the 16-bit shifts (ptxas
  // turns shifts by 16 into IMAD.SHL / half-rate IMAD.HI) and carry chains.
  const u32 t1 = __byte_perm(acc[1], 0, 0x1044), u1 = __byte_perm(acc[1], 0, 0x4432);
  const u32 t3 = __byte_perm(acc[3], 0, 0x1044), u3 = __byte_perm(acc[3], 0, 0x4432);
  u32 lo, hi;
  asm("{\n\t"
      ".reg .u32 w2, c;\n\t"
      "add.cc.u32   %0, %2, %3;\n\t"   // w0 = acc0 + (acc1 << 16)
      "addc.u32     %1, %4, %5;\n\t"   // w1 = (acc1 >> 16) + acc2 + carry      (< 2^26)
      "add.cc.u32   %1, %1, %6;\n\t"   // w1 += acc3 << 16
      "addc.u32     w2, %7, 0;\n\t"    // w2 = (acc3 >> 16) + carry             (< 2^10)
      "add.cc.u32   %1, %1, w2;\n\t"   // + w2 2^32 ...
      "addc.u32     c, 0, 0;\n\t"
      "sub.cc.u32   %0, %0, w2;\n\t"   // ... - w2        (w2 2^64 == w2 (2^32 - 1)); cannot borrow out
      "subc.u32     %1, %1, 0;\n\t"
      "sub.u32      c, 0, c;\n\t"      // wrapped by 2^64: add 2^32 - 1 (the wrapped value is < 2^42)
      "add.cc.u32   %0, %0, c;\n\t"
      "addc.u32     %1, %1, 0;\n\t"
      "}"
      : "=&r"(lo), "=&r"(hi)
      : "r"(acc[0]), "r"(t1), "r"(u1), "r"(acc[2]), "r"(t3), "r"(u3));
  s[R] = ((u64)hi << 32) | lo;
  if constexpr (R + 1 < 12) mds_rows<R + 1>(s, X, rc2);
}
__device__ __forceinline__ void mds_rc(u64 s[12], const u64* rc2) {
  u32 X[6][4];
#pragma unroll
  for (int k = 0; k < 6; k++) {
    const u32 a0 = (u32)s[2 * k], a1 = (u32)(s[2 * k] >> 32), b0 = (u32)s[2 * k + 1], b1 = (u32)(s[2 * k + 1] >> 32);
    X[k][0] = __byte_perm(a0, b0, 0x5410);
    X[k][1] = __byte_perm(a0, b0, 0x7632);
    X[k][2] = __byte_perm(a1, b1, 0x5410);
    X[k][3] = __byte_perm(a1, b1, 0x7632);
  }
  mds_rows<0>(s, X, rc2);
}
// The whole permutation has one copy of the 12-lane S-box code (three lanes at a time, lanes rotated through
// fixed registers) and two copies of the MDS code: ~1.7 k instructions, so it stays inside the 32 KB L1.5
// instruction cache. (Straight-line code per round type was 12.6 k
// instructions and ran at a 51 % instruction-cache hit rate - ncu, profiles/r1_*.)
__device__ __forceinline__ void permute(u64 s[12], const u64* rc2) {
#pragma unroll
  for (int i = 0; i < 12; i++) {  // inputs are canonical: s + k wraps at most once
    const u64 k = rc2[N_ROUNDS * 24 + 24 + i];
    u64 t = s[i] + k;
    if (t < k) t += gl::EPS;
    s[i] = t;
  }
  // two passes over { 4 full rounds; in the first pass also the 22 partial rounds }: one copy of the 12-lane
  // S-box code, and the partial rounds as their own straight-line body, so that the scheduler overlaps the serial
  // S-box chain of lane 0 with the dp2a products of the other lanes
#pragma unroll 1
  for (int phase = 0; phase < 2; phase++) {
#pragma unroll 1
    for (int r4 = 0; r4 < HALF_FULL; r4++) {
      const int r = phase * (HALF_FULL + N_PARTIAL) + r4;
#pragma unroll 1
      for (int j = 0; j < 2; j++) {  // six lanes per iteration, then the two halves swap places
        u64 t[6];
#pragma unroll
        for (int i = 0; i < 6; i++) t[i] = sbox(s[i]);
#pragma unroll
        for (int i = 0; i < 6; i++) {
          s[i] = s[i + 6];
          s[i + 6] = t[i];
        }
      }
      mds_rc(s, rc2 + (r + 1) * 24);
    }
    if (phase == 0) {
#pragma unroll 1
      for (int r = HALF_FULL; r < HALF_FULL + N_PARTIAL; r++) {
        s[0] = sbox(s[0]);
        mds_rc(s, rc2 + (r + 1) * 24);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 12; i++) s[i] = s[i] >= gl::P ? s[i] - gl::P : s[i];
}

}  // namespace lazy
#endif

PB_HD void permute(u64 s[12]) {
#if defined(__CUDA_ARCH__)
  lazy::permute(s, RC2_DEV);
#else
  permute_generic(s);
#endif
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
