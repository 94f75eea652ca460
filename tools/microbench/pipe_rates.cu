// Micro-benchmark (GPU): issue rate of the integer instructions the Goldilocks kernels are made of.
// Each thread runs 8 independent dependency chains of one PTX pattern; prints lanes/clk/SM at 1.965 GHz.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
typedef unsigned int u32;

template <int MODE>
__global__ void k_rate(u32* out, int iters, u32 seed) {
  u32 a = threadIdx.x * 2654435761u + seed, b = blockIdx.x * 40503u + 7 + seed;
  u32 lo[8], hi[8];
  for (int j = 0; j < 8; j++) { lo[j] = a + j; hi[j] = b ^ j; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int j = 0; j < 8; j++) {
      if (MODE == 0) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(lo[j]) : "r"(a), "r"(b));
      if (MODE == 1) asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(lo[j]) : "r"(a), "r"(b));
      if (MODE == 2) { u64 t = ((u64)hi[j] << 32) | lo[j]; asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(t) : "r"(a), "r"(b)); lo[j] = (u32)t; hi[j] = (u32)(t >> 32); }
      if (MODE == 3) asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.u32 %1, %2, %3, %1;" : "+r"(lo[j]), "+r"(hi[j]) : "r"(a), "r"(b));
      if (MODE == 4) asm volatile("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, %3;" : "+r"(lo[j]), "+r"(hi[j]) : "r"(a), "r"(b));
      if (MODE == 5) asm volatile("add.u32 %0, %0, %1;" : "+r"(lo[j]) : "r"(b));
      if (MODE == 6) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(lo[j]) : "r"(a), "r"(b));
      if (MODE == 7) asm volatile("shf.l.wrap.b32 %0, %0, %1, 5;" : "+r"(lo[j]) : "r"(hi[j]));
      if (MODE == 8) asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(lo[j]) : "r"(b));
      if (MODE == 9) { asm volatile("{.reg .pred p; setp.lt.u32 p, %0, %1; selp.u32 %0, %2, %0, p;}" : "+r"(lo[j]) : "r"(a), "r"(b)); }
      if (MODE == 11) asm volatile("dp2a.lo.u32.u32 %0, %1, %2, %0;" : "+r"(lo[j]) : "r"(a), "r"(b));
      if (MODE == 12) asm volatile("dp4a.u32.u32 %0, %1, %2, %0;" : "+r"(lo[j]) : "r"(a), "r"(b));
      if (MODE == 13) asm volatile("prmt.b32 %0, %0, %1, 0x5410;" : "+r"(lo[j]) : "r"(b));
      if (MODE == 14) { asm volatile("dp2a.lo.u32.u32 %0, %1, %2, %0;" : "+r"(lo[j]) : "r"(a), "r"(b)); asm volatile("add.cc.u32 %0, %0, %1;\n\taddc.u32 %0, %0, %1;" : "+r"(hi[j]) : "r"(b)); }
      if (MODE == 10) asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.cc.u32 %1, %2, %3, %1;\n\taddc.u32 %0, %0, 0;" : "+r"(lo[j]), "+r"(hi[j]) : "r"(a), "r"(b));
    }
  }
  u32 x = 0;
  for (int j = 0; j < 8; j++) x ^= lo[j] ^ hi[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}

template <int MODE>
static void run(const char* name, int ptx_per_iter, u32* d, cudaEvent_t e0, cudaEvent_t e1) {
  const int iters = 4096;
  float best = 1e9;
  for (int pass = 0; pass < 3; pass++) {
    cudaEventRecord(e0);
    k_rate<MODE><<<148 * 16, 256>>>(d, iters, pass);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  const double pat = 148.0 * 16 * 256 * iters * 8;
  printf("%-44s %.3f ms  %.1f patterns/clk/SM  (%d PTX instr per pattern)\n", name, best, pat / best * 1e3 / 148 / 1.965e9,
         ptx_per_iter);
}

int main() {
  u32* d;
  cudaMalloc(&d, 148 * 16 * 256 * 4);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  run<0>("mad.lo.u32 (IMAD)", 1, d, e0, e1);
  run<1>("mad.hi.u32 (IMAD.HI)", 1, d, e0, e1);
  run<2>("mad.wide.u32 (IMAD.WIDE)", 1, d, e0, e1);
  run<3>("mad.lo.cc + madc.hi", 2, d, e0, e1);
  run<10>("mad.lo.cc + madc.hi.cc + addc", 3, d, e0, e1);
  run<4>("add.cc + addc (IADD3 + IADD3.X)", 2, d, e0, e1);
  run<5>("add.u32", 1, d, e0, e1);
  run<6>("lop3", 1, d, e0, e1);
  run<7>("shf.l.wrap", 1, d, e0, e1);
  run<8>("mul.hi.u32", 1, d, e0, e1);
  run<11>("dp2a.lo.u32.u32 (IDP.2A)", 1, d, e0, e1);
  run<12>("dp4a.u32.u32 (IDP.4A)", 1, d, e0, e1);
  run<13>("prmt", 1, d, e0, e1);
  run<14>("dp2a + (add.cc, addc) interleaved", 3, d, e0, e1);
  run<9>("setp + selp", 2, d, e0, e1);
  printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
