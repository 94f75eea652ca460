// K4 `leaf_hash` + K5 `merkle_levels`: Poseidon Merkle tree with cap over a column-major LDE.
// Replaces MerkleTree::new of plonky2 0.2.2 hash/merkle_tree.rs as used by PolynomialBatch
// (un-vendored; call site src/starks/common/prover.rs:31-38):
//   leaf i = LDE row rev(i);  digest = hash_or_noop(row) (rows of <= 4 elements are NOT hashed);
//   parent = two_to_one(left, right);  cap = level with 2^cap_height nodes.
// Layout: the LDE stays column-major [col][N] in natural row order; thread j reads element j of
// every column (a warp reads 256 contiguous bytes per column, fully coalesced, no transpose) and
// writes its 32-byte digest at the bit-reversed position. Digest levels are stored bottom-up in
// one array: level 0 (N digests) | level 1 (N/2) | ... | cap.
// Algorithmic bytes: 8 W N (read) + 32 (2N - 2^cap_height) (digests); the leaf kernel is bound by
// the integer pipe (ceil(W/8) permutations per row), not by HBM.
#pragma once
#include "poseidon.cuh"
#include "poseidon_tc.cuh"

namespace merkle {

using poseidon::Digest;

struct LeafHashK {
  const u64* lde;
  size_t stride;
  int W, log_n;
  Digest* out;
  int natural = 0;  // 1: digest of row j is stored at j (row-sharded hashing), 0: at bit_reverse(j)
  PB_HD void operator()(size_t j) const {
    u64 s[12];
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = 0;
    if (W <= 4) {  // hash_or_noop
      for (int c = 0; c < W; c++) s[c] = lde[(size_t)c * stride + j];
    } else {
      int c = 0;
      for (; c + 8 <= W; c += 8) {
#pragma unroll
        for (int k = 0; k < 8; k++) s[k] = lde[(size_t)(c + k) * stride + j];
        poseidon::permute(s);
      }
      if (c < W) {
        for (int k = 0; c + k < W; k++) s[k] = lde[(size_t)(c + k) * stride + j];
        poseidon::permute(s);
      }
    }
    Digest d;
#pragma unroll
    for (int i = 0; i < 4; i++) d.e[i] = s[i];
    out[natural ? j : (size_t)gl::brev32((u32)j, log_n)] = d;
  }
};

// leaves stored row-major and already in leaf order (FRI layers): leaf i = rows[i*len .. +len)
struct RowHashK {
  const u64* rows;
  int len;
  Digest* out;
  PB_HD void operator()(size_t i) const {
    u64 s[12];
#pragma unroll
    for (int k = 0; k < 12; k++) s[k] = 0;
    const u64* row = rows + i * (size_t)len;
    if (len <= 4) {
      for (int c = 0; c < len; c++) s[c] = row[c];
    } else {
      for (int c = 0; c < len; c += 8) {
        for (int k = 0; k < 8 && c + k < len; k++) s[k] = row[c + k];
        poseidon::permute(s);
      }
    }
    Digest d;
#pragma unroll
    for (int k = 0; k < 4; k++) d.e[k] = s[k];
    out[i] = d;
  }
};

struct LevelK {
  const Digest* child;
  Digest* parent;
  PB_HD void operator()(size_t i) const { parent[i] = poseidon::two_to_one(child[2 * i], child[2 * i + 1]); }
};

#if !PB_HOSTSIM
// ---- hot kernels: split round-constant table in shared memory, grid sized to the machine -----------
__device__ __forceinline__ void stage_rc2(u64* rc2) {
  for (int i = threadIdx.x; i < poseidon::RC2_WORDS; i += blockDim.x) rc2[i] = poseidon::RC2_DEV[i];
  __syncthreads();
}

// One thread per LDE row, the MDS layers on the tensor cores (poseidon_tc.cuh): one CTA = 128 rows = the 128 rows of the
// u8 x u8 -> s32 product. Every thread takes part in every permutation (the CTA-wide barrier and the TMEM loads are
// collective), rows past the end hash row 0 and drop the result.
__global__ void __launch_bounds__(poseidon::tc::CTA, poseidon::tc::CTAS_PER_SM)
    k_leaf_hash_tc(const u64* __restrict__ lde, size_t stride, int W, int log_n, Digest* __restrict__ out, size_t rows,
                   int natural) {
  extern __shared__ unsigned char tc_dyn[];
  __shared__ u64 tc_bar;
  __shared__ u32 tc_slot[2];
  poseidon::tc::Ctx ctx;
  poseidon::tc::setup(ctx, tc_dyn, &tc_bar, tc_slot);
  const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = j < rows;
  const u64* p = lde + (live ? j : 0);
  u64 s[12];
#pragma unroll
  for (int i = 0; i < 12; i++) s[i] = 0;
  const int chunks = (W + 7) / 8;
#pragma unroll 1
  for (int c = 0; c < chunks; c++) {
    const int len = W - 8 * c;
    const u64* q = p + (size_t)c * 8 * stride;
#pragma unroll
    for (int k = 0; k < 8; k++)
      if (k < len) s[k] = q[(size_t)k * stride];
    poseidon::tc::permute(s, ctx);
  }
  if (live) {
    Digest d;
#pragma unroll
    for (int i = 0; i < 4; i++) d.e[i] = s[i];
    out[natural ? j : (size_t)gl::brev32((u32)j, log_n)] = d;
  }
  poseidon::tc::teardown(ctx);
}

__global__ void __launch_bounds__(128, 6) k_level(const Digest* __restrict__ child, Digest* __restrict__ parent, size_t n) {
  __shared__ __align__(16) u64 rc2[poseidon::RC2_WORDS];
  stage_rc2(rc2);
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  u64 s[12];
  const Digest l = child[2 * i], r = child[2 * i + 1];
#pragma unroll
  for (int k = 0; k < 4; k++) {
    s[k] = l.e[k];
    s[4 + k] = r.e[k];
    s[8 + k] = 0;
  }
  poseidon::lazy::permute(s, rc2);
  Digest d;
#pragma unroll
  for (int k = 0; k < 4; k++) d.e[k] = s[k];
  parent[i] = d;
}
#endif

static inline size_t tree_digests(int log_n, int cap_height) {
  size_t total = 0;
  for (int l = log_n; l >= cap_height; l--) total += (size_t)1 << l;
  return total;
}
// offset (in digests) of level `lvl` (0 = leaves)
static inline size_t level_offset(int log_n, int lvl) {
  size_t off = 0;
  for (int l = 0; l < lvl; l++) off += (size_t)1 << (log_n - l);
  return off;
}

// parent[i] = two_to_one(child[2 i], child[2 i + 1]) for i < n
static inline void launch_level(const Digest* child, Digest* parent, size_t n, pbStream s) {
#if PB_HOSTSIM
  LevelK k{child, parent};
  pb_launch("merkle level", k, n, s, 128);
#else
  k_level<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(child, parent, n);
  g_pb_launches++;
  pb_check_last("merkle level");
#endif
}
static inline void build_levels(Digest* digests, int log_n, int cap_height, pbStream s) {
  if (cap_height > log_n) throw Pb254Error(6, "merkle: cap_height > log2(leaves)");
  size_t off = 0;
  for (int l = 0; l < log_n - cap_height; l++) {
    size_t m = (size_t)1 << (log_n - l);
    launch_level(digests + off, digests + off + m, m / 2, s);
    off += m;
  }
}

// digests: tree_digests(log_n, cap_height) entries; the cap is the last 2^cap_height of them
static inline void build_from_lde(const u64* lde, size_t stride, int W, int log_n, int cap_height, Digest* digests,
                                  pbStream s) {
#if PB_HOSTSIM
  LeafHashK k{lde, stride, W, log_n, digests};
  pb_launch("leaf hash", k, (size_t)1 << log_n, s, 128);
#else
  if (W <= 4) {  // hash_or_noop: no permutation, the generic functor kernel is fine
    LeafHashK k{lde, stride, W, log_n, digests};
    pb_launch("leaf copy", k, (size_t)1 << log_n, s, 128);
  } else {
    const size_t n = (size_t)1 << log_n;
    k_leaf_hash_tc<<<(unsigned)((n + 127) / 128), poseidon::tc::CTA, poseidon::tc::SMEM_BYTES, s>>>(lde, stride, W, log_n,
                                                                                                  digests, n, 0);
    g_pb_launches++;
    pb_check_last("leaf hash");
  }
#endif
  build_levels(digests, log_n, cap_height, s);
}
static inline void build_from_rows(const u64* rows, int len, int log_n, int cap_height, Digest* digests, pbStream s) {
  RowHashK k{rows, len, digests};
  pb_launch("row hash", k, (size_t)1 << log_n, s, 128);
  build_levels(digests, log_n, cap_height, s);
}


// ---- row-sharded pieces for the oversized-trace mode -------------------------------------
// digests of `rows` rows of a column-major matrix [W][stride], natural order: out[i] = H(row i)
static inline void hash_rows_natural(const u64* m, size_t stride, int W, size_t rows, Digest* out, pbStream s) {
#if PB_HOSTSIM
  LeafHashK k{m, stride, W, 0, out, 1};
  pb_launch("leaf hash (rows)", k, rows, s, 128);
#else
  if (W <= 4) {
    LeafHashK k{m, stride, W, 0, out, 1};
    pb_launch("leaf copy (rows)", k, rows, s, 128);
  } else {
    k_leaf_hash_tc<<<(unsigned)((rows + 127) / 128), poseidon::tc::CTA, poseidon::tc::SMEM_BYTES, s>>>(m, stride, W, 0, out,
                                                                                                     rows, 1);
    g_pb_launches++;
    pb_check_last("leaf hash (rows)");
  }
#endif
}
// leaves[q] = all[bit_reverse(first + q, log_N)] for q < 2^log_sub: the leaves of one subtree in tree order
struct SubtreeLeavesK {
  const Digest* all;
  Digest* leaves;
  size_t first;
  int log_N;
  PB_HD void operator()(size_t q) const { leaves[q] = all[gl::brev32((u32)(first + q), log_N)]; }
};

}  // namespace merkle
