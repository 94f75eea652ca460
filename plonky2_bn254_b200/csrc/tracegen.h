// K1 `tracegen_{g1,g2,fq}` + K2 `range_check_fill`: batched trace generation on the device.
// Replaces (bit-exact on every cell):
//   generate_trace / generate_one_set / generate_first_row / generate_transition
//        src/starks/curves/g1/scalar_mul_stark.rs:55-213, g2/scalar_mul_stark.rs:55-213,
//        src/starks/fields/exp_stark.rs:53-196
//   generate_g1_add src/starks/curves/g1/add.rs:52-122, generate_g2_add g2/add.rs:59-130,
//   generate_fq_mul src/starks/fields/mul.rs:22-40
//   generate_modulus_zero src/starks/modular/modulus_zero.rs:77-123
//   generate_is_modulus_zero src/starks/modular/is_modulus_zero.rs:36-66
//   generate_round_flags src/starks/common/round_flags.rs:21-44
//   generate_range_checks g1/scalar_mul_stark.rs:71-87
//
// The reference walks the 512 rows of an instance sequentially with one or two field inversions
// per row. Here (result-identical):
//   1. chains   one thread per instance: the doubling chain D_j = 2^j x and the running-sum chain
//               T_j = S_(j-1) + D_j in Jacobian coordinates, each normalised to affine with one
//               Montgomery-trick inversion;
//   2. dens     one thread per row: the slope denominator (b.x - a.x, or 2 a.y when a.x = b.x);
//   3. batchinv one thread per 16 denominators;
//   4. rows     one thread per row (512 lanes = 16 warps per scalar-mul): slope, result point,
//               every limb-polynomial witness, flags and bits, written column-major (coalesced).
// Row r of an instance (j = r >> 1): even r is an adding row a = S_(j-1), b = D_j, c = T_j; odd r is a
// doubling row a = b = D_j, c = D_(j+1); S_j = bit_j ? T_j : S_(j-1), S_(-1) = offset.
#pragma once
#include "compat.cuh"
#include "context.cuh"

namespace tg {

static constexpr int PERIOD = 512, NBITS = 256;
// device error word: the maximum wins, so the root cause outranks the inconsistencies it triggers
static constexpr int ERR_INTERNAL = 1, ERR_INFINITY = 2, ERR_NOT_CANONICAL = 3, ERR_NOT_ON_CURVE = 4;

struct Layout {
  int kind, L, aux_len, width;
  int reg0, reg1, a, b, c, aux, bits, rf, ts, flag_op, flag_sq_nl, filter, freq, range_counter;
  int rc_lo, rc_hi, in_words;
};
static inline Layout layout_for(int kind) {
  Layout l;
  l.kind = kind;
  l.L = kind == 0 ? 32 : kind == 1 ? 64 : 16;
  l.aux_len = kind == 0 ? 354 : kind == 1 ? 708 : 80;
  l.reg0 = 0;
  l.reg1 = l.L;
  l.a = 2 * l.L;
  l.b = 3 * l.L;
  l.c = 4 * l.L;
  l.aux = 5 * l.L;
  l.bits = l.aux + l.aux_len;
  l.rf = l.bits + NBITS;
  l.ts = l.rf + 5;
  l.flag_op = l.ts + 1;
  l.flag_sq_nl = l.ts + 2;
  l.filter = l.ts + 3;
  l.freq = l.ts + 4;
  l.range_counter = l.ts + 5;
  l.width = l.ts + 6;
  l.rc_lo = 2 * l.L;
  l.rc_hi = l.bits;
  l.in_words = kind == 0 ? 20 : kind == 1 ? 36 : 8;
  return l;
}

// device bytes needed besides the trace itself
static inline size_t scratch_bytes(int kind, size_t K) {
  size_t fe = kind == 1 ? 64 : 32;  // coordinate size
  if (kind == 2) return (257 + 256) * K * 32 + 256 * K * 2 + PERIOD * 5 * 8 + 8192;
  size_t aff = 2 * fe, jac = 3 * fe;
  size_t nden = kind == 1 ? 3 : 1;
  return (257 + 256 + 1) * K * aff + 2 * 256 * K * jac + 256 * K * fe + 256 * K * 2 + nden * PERIOD * K * 32 +
         PERIOD * 5 * 8 + 16384;
}


// Fills the column-major device trace (width x n_rows); errors land in *d_err (ERR_* codes).
// row0 != 0: the matrix is the row block [row0, row0 + n_rows) of a larger trace (one proof across several GPUs):
// the range counter continues at row0; the frequency column then holds the histogram of THIS block's cells in its
// first 65536 rows, to be added up across the blocks by the caller (pb254.cu).
void generate(Arena& ar, int kind, const u64* d_inputs, const u64* d_ts, size_t K, size_t n_rows, u64* d_trace,
              int* d_err, pbStream s, size_t row0 = 0);

}  // namespace tg
