// K8 `quotient_eval_{g1,g2,fq}` interface. See quotient_impl.cuh.
#pragma once
#include "context.cuh"
#include "aux.cuh"

namespace quot {

struct Params {
  const u64* tr;      // trace LDE, column-major, stride tr_stride
  size_t tr_stride;
  const u64* ax;      // auxiliary LDE, column-major
  size_t ax_stride;
  u64* out;           // nch x size, natural order: combined constraints / Z_H at 7 * w^i
  const u64* weights; // [K][nch]: alpha_j^(K-1-k)
  const u64* bpow;    // [nch][bpow_stride]: beta_j^k, k < 2 L + 17 (CTL combinations)
  int bpow_stride;
  size_t size;        // quotient domain size = 2 n
  int log_size;
  // Row-block form (one proof across several GPUs, prover.cuh): this call evaluates the `count` points
  // [i_base, i_base + count); tr / ax hold the LDE rows of exactly these points followed by the 2 * step rows of the
  // next-row halo (no wrap-around inside the block), out is [nch][out_stride] with point i_base at index 0.
  // Whole-domain form: i_base = 0, count = size, out_stride = size, wrap = 1.
  size_t i_base, count, out_stride;
  int wrap;
  size_t step;        // LDE index of quotient point i is i * step  (2^(rate_bits - 1))
  ntt::Tables t;
  u64 g;              // root of unity of order n
  u64 g_inv;          // "last" = g^-1
  u64 n_field;        // n as a field element
  u64 zh[2];          // Z_H at even / odd points: 7^n * (+-1) - 1
  u64 zh_inv[2];
  aux::Challenges ch;
  int* err;           // set to 1 if the emitted constraint count disagrees with the weight table
  u64* dbg;           // keep_debug only: running total[0] after every constraint group of point dbg_point
  size_t dbg_point;
};

// constraints per (local, next) row pair incl. lookups and CTLs for `nch` challenges
int num_constraints(int kind, int nch);

void run_g1(const Params& p, pbStream s);
void run_g2(const Params& p, pbStream s);
void run_fq(const Params& p, pbStream s);

}  // namespace quot
