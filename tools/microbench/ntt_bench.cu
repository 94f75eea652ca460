// Micro-benchmark (GPU): the LDE kernels of csrc/ntt.cuh on a slab of columns, timed with CUDA events.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I plonky2_bn254_b200/csrc [-DNTT_LB_CTAS=3] \
//        tools/microbench/ntt_bench.cu -o gpurun_out/ntt_bench && gpurun_out/ntt_bench [cols] [log_n] [r]
#include "ntt.cuh"
std::atomic<unsigned long long> g_pb_launches;
int main(int argc, char** argv) {
  const int cols = argc > 1 ? atoi(argv[1]) : 200, L = argc > 2 ? atoi(argv[2]) : 19, r = argc > 3 ? atoi(argv[3]) : 1;
  const size_t n = (size_t)1 << L, N = n << r;
  cudaStream_t s;
  cudaStreamCreate(&s);
  ntt::TableSet ts;
  ts.init(s);
  u64 *in, *out, *scratch;
  cudaMalloc(&in, cols * n * 8);
  cudaMalloc(&out, cols * N * 8);
  cudaMalloc(&scratch, cols * n * 8);
  cudaMemset(in, 0x5a, cols * n * 8);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e9;
  for (int it = 0; it < 6; it++) {
    cudaEventRecord(e0, s);
    ntt::lde_columns(ts, in, n, out, N, scratch, cols, L, r, ntt::FROM_VALUES_LDE, s);
    cudaEventRecord(e1, s);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (it && ms < best) best = ms;
  }
  u64 chk[4];
  cudaMemcpy(chk, out + N / 3, 32, cudaMemcpyDeviceToHost);
  printf("cols=%d L=%d r=%d: %.3f ms  %.1f GB/s algorithmic  chk=%016llx err=%s\n", cols, L, r, best,
         8.0 * cols * n * (2 + (1 << r)) / best / 1e6, (unsigned long long)(chk[0] ^ chk[3]), cudaGetErrorString(cudaGetLastError()));
  return 0;
}
