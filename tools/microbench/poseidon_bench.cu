// Micro-benchmark (GPU): Poseidon permutation variants and integer pipe rates. Build: see Makefile.
#include "../../plonky2_bn254_b200/csrc/compat.cuh"
namespace lazy {
static constexpr u64 EPS = 0xFFFFFFFFULL, P = 0xFFFFFFFF00000001ULL;
#ifdef MUL_I128
__device__ __forceinline__ u64 mulw(u32 a, u32 b) { u64 r; asm("mul.wide.u32 %0, %1, %2;" : "=l"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ u64 mul(u64 a, u64 b) {
  const u32 a0 = (u32)a, a1 = (u32)(a >> 32), b0 = (u32)b, b1 = (u32)(b >> 32);
  const u64 P = mulw(a0, b0), Q = mulw(a0, b1), R = mulw(a1, b0), S = mulw(a1, b1);
  unsigned __int128 prod = (unsigned __int128)P + (((unsigned __int128)Q + R) << 32) + ((unsigned __int128)S << 64);
  const u64 lo = (u64)prod, hi = (u64)(prod >> 64);
  const u32 x2 = (u32)hi, x3 = (u32)(hi >> 32);
  __int128 V = (__int128)lo - x3 - x2 + ((__int128)x2 << 32);
  const u64 r = (u64)V;
  const long long w = (long long)(V >> 64);
#ifdef ADJ_ALU
  // w in {-1, 0, 1}: w (2^32 - 1) = ((w >> 1) << 32) | (u32)(-w), built on the ALU pipe (the FMA pipe is the bottleneck)
  u32 adj_lo, adj_hi;
  asm("neg.s32 %0, %2;\n\tshr.s32 %1, %2, 1;" : "=r"(adj_lo), "=r"(adj_hi) : "r"((int)w));
  return r + (((u64)adj_hi << 32) | adj_lo);
#else
  return r + (u64)(w * 0xFFFFFFFFLL);
#endif
}
#else
__device__ __forceinline__ u64 mul(u64 a, u64 b) {
  const u32 a0 = (u32)a, a1 = (u32)(a >> 32), b0 = (u32)b, b1 = (u32)(b >> 32);
  u32 r0, r1;
  asm("{\n\t"
      ".reg .u32 x0, x1, x2, x3, m, l, h, c;\n\t"
      ".reg .u64 P, Q, R, S; .reg .u32 p1, q0, q1, s0, s1, t0, t1;\n\t"
      "mul.wide.u32 P, %2, %4;\n\t"
      "mul.wide.u32 Q, %2, %5;\n\t"
      "mul.wide.u32 R, %3, %4;\n\t"
      "mul.wide.u32 S, %3, %5;\n\t"
      "mov.b64 {x0, p1}, P; mov.b64 {q0, q1}, Q; mov.b64 {t0, t1}, R; mov.b64 {s0, s1}, S;\n\t"
      "add.cc.u32   x1, p1, q0;\n\t"
      "addc.cc.u32  x2, q1, s0;\n\t"
      "addc.u32     x3, s1, 0;\n\t"
      "add.cc.u32   x1, x1, t0;\n\t"
      "addc.cc.u32  x2, x2, t1;\n\t"
      "addc.u32     x3, x3, 0;\n\t"
      "sub.cc.u32   %0, x0, x3;\n\t"
      "subc.cc.u32  %1, x1, 0;\n\t"
      "subc.u32     m, 0, 0;\n\t"
      "sub.cc.u32   %0, %0, m;\n\t"
      "subc.u32     %1, %1, 0;\n\t"
      "sub.cc.u32   l, 0, x2;\n\t"
      "subc.u32     h, x2, 0;\n\t"
      "add.cc.u32   %0, %0, l;\n\t"
      "addc.cc.u32  %1, %1, h;\n\t"
      "addc.u32     c, 0, 0;\n\t"
      "sub.u32      c, 0, c;\n\t"
      "add.cc.u32   %0, %0, c;\n\t"
      "addc.u32     %1, %1, 0;\n\t"
      "}"
      : "=&r"(r0), "=&r"(r1)
      : "r"(a0), "r"(a1), "r"(b0), "r"(b1));
  return ((u64)r1 << 32) | r0;
}
#endif
__device__ __forceinline__ u64 sbox(u64 x) {
  const u64 x2 = mul(x, x), x4 = mul(x2, x2), x3 = mul(x, x2);
  return mul(x3, x4);
}
__device__ __forceinline__ u64 reduce_split(u64 al, u64 ah) {
  const u32 h0 = (u32)ah, h1 = (u32)(ah >> 32);
  u64 t = al;
  asm("mad.wide.u32 %0, %1, 0xffffffff, %0;" : "+l"(t) : "r"(h1));
  u32 r0 = (u32)t, r1 = (u32)(t >> 32);
  asm("{\n\t"
      ".reg .u32 c;\n\t"
      "add.cc.u32   %1, %1, %2;\n\t"
      "addc.u32     c, 0, 0;\n\t"
      "sub.u32      c, 0, c;\n\t"
      "add.cc.u32   %0, %0, c;\n\t"
      "addc.u32     %1, %1, 0;\n\t"
      "}"
      : "+r"(r0), "+r"(r1)
      : "r"(h0));
  return ((u64)r1 << 32) | r0;
}
// rc2: [12][2] u64 = (lo half, hi half) of the next round's constants
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
#ifdef RC_PIECES
  // the next round's constant enters as the initial value of the four piece accumulators
  const uint4 k = *reinterpret_cast<const uint4*>(rc2 + 2 * R);
  u32 acc[4] = {k.x, k.y, k.z, k.w};
  dp_row<R, 0>(acc, X);
#ifdef ALU_RECOMB
  {
    // value = acc0 + acc1 2^16 + acc2 2^32 + acc3 2^48 (acc < 2^25) -> u64 representative, ALU pipe only
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
        "sub.cc.u32   %0, %0, w2;\n\t"   // ... - w2        (w2 2^64 == w2 (2^32 - 1))
        "subc.u32     %1, %1, 0;\n\t"
        "sub.u32      c, 0, c;\n\t"      // wrapped by 2^64: add 2^32 - 1
        "add.cc.u32   %0, %0, c;\n\t"
        "addc.u32     %1, %1, 0;\n\t"
        "}"
        : "=&r"(lo), "=&r"(hi)
        : "r"(acc[0]), "r"(t1), "r"(u1), "r"(acc[2]), "r"(t3), "r"(u3));
    s[R] = ((u64)hi << 32) | lo;
  }
  if constexpr (R + 1 < 12) mds_rows<R + 1>(s, X, rc2);
  return;
#endif
  const u64 al = (u64)acc[0] + ((u64)acc[1] << 16), ah = (u64)acc[2] + ((u64)acc[3] << 16);
#else
  u32 acc[4] = {0, 0, 0, 0};
  dp_row<R, 0>(acc, X);
  u64 al = rc2[2 * R], ah = rc2[2 * R + 1];
  asm("mad.wide.u32 %0, %1, 1, %0;" : "+l"(al) : "r"(acc[0]));
  asm("mad.wide.u32 %0, %1, 65536, %0;" : "+l"(al) : "r"(acc[1]));
  asm("mad.wide.u32 %0, %1, 1, %0;" : "+l"(ah) : "r"(acc[2]));
  asm("mad.wide.u32 %0, %1, 65536, %0;" : "+l"(ah) : "r"(acc[3]));
#endif
  s[R] = reduce_split(al, ah);
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
__device__ __forceinline__ void permute(u64 s[12], const u64* rc2) {
  // rc2[0] block holds round 0's constants: s += rc (canonical inputs)
#pragma unroll
  for (int i = 0; i < 12; i++) {
#ifdef RC_PIECES
    u64 k = rc2[31 * 24 + i];
#else
    u64 k = rc2[2 * i] | (rc2[2 * i + 1] << 32);
#endif
    u64 t = s[i] + k;
    if (t < k) t += EPS;
    s[i] = t;
  }
#ifdef SPLIT_LOOPS
  // partial rounds as their own straight-line body so that the scheduler can overlap the serial S-box chain of
  // lane 0 with the dp2a products of the other lanes
#pragma unroll 1
  for (int phase = 0; phase < 2; phase++) {
#pragma unroll 1
    for (int r4 = 0; r4 < 4; r4++) {
      const int r = phase * 26 + r4;
#ifdef SBOX6
#pragma unroll 1
      for (int j = 0; j < 2; j++) {
        u64 t[6];
#pragma unroll
        for (int i = 0; i < 6; i++) t[i] = sbox(s[i]);
#pragma unroll
        for (int i = 0; i < 6; i++) { s[i] = s[i + 6]; s[i + 6] = t[i]; }
      }
#else
#ifdef UNROLL_SBOX
#pragma unroll
#else
#pragma unroll 1
#endif
      for (int j = 0; j < 4; j++) {
        const u64 t0 = sbox(s[0]), t1 = sbox(s[1]), t2 = sbox(s[2]);
#pragma unroll
        for (int i = 0; i < 9; i++) s[i] = s[i + 3];
        s[9] = t0; s[10] = t1; s[11] = t2;
      }
#endif
      mds_rc(s, rc2 + (r + 1) * 24);
    }
    if (phase == 0) {
#pragma unroll 1
      for (int r = 4; r < 26; r++) {
        s[0] = sbox(s[0]);
        mds_rc(s, rc2 + (r + 1) * 24);
      }
    }
  }
#else
#pragma unroll 1
  for (int r = 0; r < 30; r++) {
    if (r < 4 || r >= 26) {
#pragma unroll 1
      for (int j = 0; j < 4; j++) {
        const u64 t0 = sbox(s[0]), t1 = sbox(s[1]), t2 = sbox(s[2]);
#pragma unroll
        for (int i = 0; i < 9; i++) s[i] = s[i + 3];
        s[9] = t0; s[10] = t1; s[11] = t2;
      }
    } else {
      s[0] = sbox(s[0]);
    }
    mds_rc(s, rc2 + (r + 1) * 24);
  }
#endif
#pragma unroll
  for (int i = 0; i < 12; i++) s[i] = s[i] >= P ? s[i] - P : s[i];
}
}

#include "../../plonky2_bn254_b200/csrc/poseidon.cuh"
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>

#ifndef MINB
#define MINB 1
#endif
__global__ void __launch_bounds__(128, MINB) k_compact(const u64* in, u64* out, const u64* rc2g, int reps) {
  __shared__ __align__(16) u64 rc2[31 * 24 + 12];
  for (int i = threadIdx.x; i < 31 * 24 + 12; i += blockDim.x) rc2[i] = rc2g[i];
  __syncthreads();
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  u64 s[12];
  for (int k = 0; k < 12; k++) s[k] = in[i * 12 + k];
  for (int r = 0; r < reps; r++) lazy::permute(s, rc2);
  for (int k = 0; k < 12; k++) out[i * 12 + k] = s[k];
}
__global__ void __launch_bounds__(128) k_generic(const u64* in, u64* out, int reps) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  u64 s[12];
  for (int k = 0; k < 12; k++) s[k] = in[i * 12 + k];
  for (int r = 0; r < reps; r++) poseidon::permute_generic(s);
  for (int k = 0; k < 12; k++) out[i * 12 + k] = s[k];
}
// pipe-rate probes: N independent chains per thread
template <int MODE>
__global__ void k_rate(u32* out, int iters) {
  u32 a = threadIdx.x * 2654435761u + 1, b = blockIdx.x * 40503u + 7;
  u64 acc[8];
  for (int j = 0; j < 8; j++) acc[j] = a + j;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int j = 0; j < 8; j++) {
      if (MODE == 0) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[j]) : "r"(a), "r"(b));
      if (MODE == 1) { u32 lo = (u32)acc[j]; asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(lo) : "r"(a), "r"(b)); acc[j] = lo; }
      if (MODE == 2) { u32 lo = (u32)acc[j]; asm volatile("add.u32 %0, %0, %1;" : "+r"(lo) : "r"(b)); acc[j] = lo; }
      if (MODE == 3) { u64 t; asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(t) : "r"((u32)acc[j]), "r"(b)); acc[j] = t; }
    }
  }
  u64 x = 0;
  for (int j = 0; j < 8; j++) x ^= acc[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = (u32)x ^ (u32)(x >> 32);
}

static float time_ms(cudaEvent_t a, cudaEvent_t b) { float ms; cudaEventElapsedTime(&ms, a, b); return ms; }

int main() {
  const size_t n = (size_t)148 * 8 * 128 * 8;  // states
  const int reps = 16;
  std::vector<u64> h(n * 12);
  u64 x = 88172645463325252ULL;
  for (auto& v : h) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; v = x % 0xFFFFFFFF00000001ULL; }
  for (int k = 0; k < 12; k++) { h[k] = 0; h[12 + k] = k; h[24 + k] = 0xFFFFFFFF00000000ULL; }
  std::vector<u64> rc2(31 * 24 + 12, 0);
#ifdef RC_PIECES
  for (int i = 0; i < 360; i++) { u64 k = poseidon::RC_HOST[i]; rc2[2 * i] = (k & 0xFFFF) | (((k >> 16) & 0xFFFF) << 32); rc2[2 * i + 1] = ((k >> 32) & 0xFFFF) | ((k >> 48) << 32); }
  for (int i = 0; i < 12; i++) rc2[31 * 24 + i] = poseidon::RC_HOST[i];
#else
  for (int i = 0; i < 360; i++) { rc2[2 * i] = poseidon::RC_HOST[i] & 0xFFFFFFFFULL; rc2[2 * i + 1] = poseidon::RC_HOST[i] >> 32; }
#endif
  u64 *din, *dout, *dout2, *drc;
  cudaMalloc(&din, n * 96); cudaMalloc(&dout, n * 96); cudaMalloc(&dout2, n * 96); cudaMalloc(&drc, rc2.size() * 8);
  cudaMemcpy(din, h.data(), n * 96, cudaMemcpyHostToDevice);
  cudaMemcpy(drc, rc2.data(), rc2.size() * 8, cudaMemcpyHostToDevice);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  // correctness: one permutation each
  k_generic<<<n / 128, 128>>>(din, dout, 1);
  k_compact<<<n / 128, 128>>>(din, dout2, drc, 1);
  std::vector<u64> a(n * 12), b(n * 12);
  cudaMemcpy(a.data(), dout, n * 96, cudaMemcpyDeviceToHost);
  cudaMemcpy(b.data(), dout2, n * 96, cudaMemcpyDeviceToHost);
  size_t bad = 0; for (size_t i = 0; i < n * 12; i++) bad += a[i] != b[i];
  printf("KAT0 %016llx (want 3c18a9786cb0b359)  KAT1 %016llx (want d64e1e3efc5b8e9e)  mismatches generic vs compact: %zu  err=%s\n",
         (unsigned long long)a[0], (unsigned long long)a[12], bad, cudaGetErrorString(cudaGetLastError()));
  for (int pass = 0; pass < 2; pass++) {
    cudaEventRecord(e0); k_generic<<<n / 128, 128>>>(din, dout, reps); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float tg = time_ms(e0, e1);
    cudaEventRecord(e0); k_compact<<<n / 128, 128>>>(din, dout2, drc, reps); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float tc = time_ms(e0, e1);
    printf("generic %.3f ms = %.3f Gperm/s   compact %.3f ms = %.3f Gperm/s\n", tg, n * reps / tg * 1e-6, tc, n * reps / tc * 1e-6);
  }
  u32* dr; cudaMalloc(&dr, 148 * 16 * 256 * 4);
  const int iters = 4096;
  const double ops = 148.0 * 16 * 256 * iters * 8;
  float t[4];
  for (int pass = 0; pass < 2; pass++) {
    cudaEventRecord(e0); k_rate<0><<<148 * 16, 256>>>(dr, iters); cudaEventRecord(e1); cudaEventSynchronize(e1); t[0] = time_ms(e0, e1);
    cudaEventRecord(e0); k_rate<1><<<148 * 16, 256>>>(dr, iters); cudaEventRecord(e1); cudaEventSynchronize(e1); t[1] = time_ms(e0, e1);
    cudaEventRecord(e0); k_rate<2><<<148 * 16, 256>>>(dr, iters); cudaEventRecord(e1); cudaEventSynchronize(e1); t[2] = time_ms(e0, e1);
    cudaEventRecord(e0); k_rate<3><<<148 * 16, 256>>>(dr, iters); cudaEventRecord(e1); cudaEventSynchronize(e1); t[3] = time_ms(e0, e1);
  }
  const char* names[4] = {"mad.wide.u32 (acc chain)", "mad.lo.u32", "add.u32", "mul.wide.u32"};
  for (int m = 0; m < 4; m++) printf("%-26s %.3f ms  %.2f Tops/s  %.1f lanes/clk/SM @1.965GHz\n", names[m], t[m], ops / t[m] * 1e-9, ops / t[m] * 1e3 / 148 / 1.965e9);
  return 0;
}
