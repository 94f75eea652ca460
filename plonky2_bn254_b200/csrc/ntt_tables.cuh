// Two-level power tables (roots of unity of order 2^26 and powers of the coset shift 7) shared by
// the NTT, quotient and FRI kernels.
#pragma once
#include "gl.cuh"
#include <vector>

namespace ntt {

static constexpr int LOG_T = 13;          // two-level power tables: x^e = lo[e & 8191] * hi[e >> 13]
static constexpr int T = 1 << LOG_T;
static constexpr int LOG_M = 2 * LOG_T;   // master root has order 2^26
static constexpr int K1_MAX = 11;

struct Tables {
  const u64 *fwd_lo, *fwd_hi;  // W^e,   W = root of unity of order 2^26
  const u64 *inv_lo, *inv_hi;  // W^-e
  const u64 *sh_lo, *sh_hi;    // 7^e
  const u64 *ish_lo, *ish_hi;  // 7^-e
};

PB_HD u64 tpow(const u64* lo, const u64* hi, u64 e) { return gl::mul(lo[e & (T - 1)], hi[e >> LOG_T]); }

static inline void host_build_table(u64 base, std::vector<u64>& lo, std::vector<u64>& hi) {
  lo.resize(T);
  hi.resize(T);
  lo[0] = 1;
  for (int i = 1; i < T; i++) lo[i] = gl::mul(lo[i - 1], base);
  u64 step = gl::mul(lo[T - 1], base);
  hi[0] = 1;
  for (int i = 1; i < T; i++) hi[i] = gl::mul(hi[i - 1], step);
}

struct PlanCache;  // per-size NTT plans and their tables (ntt.cuh), created on first use

struct TableSet {
  u64* dev = nullptr;  // 8 * T words
  Tables t;
  PlanCache* plans = nullptr;
  void (*plans_free)(PlanCache*) = nullptr;
  void init(pbStream s) {
    std::vector<u64> all(8 * T), lo, hi;
    u64 W = gl::root_of_unity(LOG_M);
    u64 bases[4] = {W, gl::inv(W), gl::COSET_SHIFT, gl::inv(gl::COSET_SHIFT)};
    for (int k = 0; k < 4; k++) {
      host_build_table(bases[k], lo, hi);
      memcpy(&all[(2 * k) * T], lo.data(), T * 8);
      memcpy(&all[(2 * k + 1) * T], hi.data(), T * 8);
    }
    dev = (u64*)pb_dev_alloc(8 * T * 8);
    pb_h2d(dev, all.data(), 8 * T * 8, s);
    pb_sync(s);
    t.fwd_lo = dev;
    t.fwd_hi = dev + T;
    t.inv_lo = dev + 2 * T;
    t.inv_hi = dev + 3 * T;
    t.sh_lo = dev + 4 * T;
    t.sh_hi = dev + 5 * T;
    t.ish_lo = dev + 6 * T;
    t.ish_hi = dev + 7 * T;
  }
  void destroy() {
    if (plans && plans_free) plans_free(plans);
    plans = nullptr;
    if (dev) pb_dev_free(dev);
    dev = nullptr;
  }
};

}  // namespace ntt
