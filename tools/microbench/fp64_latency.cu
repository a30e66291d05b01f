// DFMA throughput vs (independent chains per thread) x (warps per SM): how much ILP x TLP the
// FP64 pipe of a B200 SM needs.  nvcc -arch=sm_100a -O3 fp64_latency.cu -o fp64_latency
#include <cstdio>
#include <cuda_runtime.h>

template <int CHAINS>
__global__ void chain_kernel(int iters, double* sink) {
  double a[CHAINS];
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) a[c] = 1.0 + 1e-9 * (threadIdx.x + c);
  const double m = 0.999999999, k = 1e-12;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) a[c] = fma(a[c], m, k);
  }
  double s = 0;
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) s += a[c];
  if (s == 1234.5) *sink = s;
}

template <int CHAINS>
void run(int warps_per_sm, double* sink) {
  int threads = 32 * (warps_per_sm >= 4 ? 4 : warps_per_sm);
  int blocks_per_sm = warps_per_sm / (threads / 32);
  int iters = 1 << 16;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  chain_kernel<CHAINS><<<148 * blocks_per_sm, threads>>>(iters / 4, sink);
  cudaEventRecord(e0);
  chain_kernel<CHAINS><<<148 * blocks_per_sm, threads>>>(iters, sink);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double fma_per_s = (double)iters * CHAINS * threads * 148.0 * blocks_per_sm / (ms * 1e-3);
  // per SM per clock at 1.965 GHz nominal
  printf("chains=%d warps/SM=%2d : %7.2f TFLOP/s  (%.1f DFMA lanes/clk/SM @1.965GHz)\n", CHAINS, warps_per_sm,
         2 * fma_per_s / 1e12, fma_per_s / 148.0 / 1.965e9);
}

int main() {
  double* sink; cudaMalloc(&sink, 8);
  int ws[] = {4, 8, 12, 16, 32};
  for (int w : ws) { run<1>(w, sink); run<2>(w, sink); run<4>(w, sink); run<8>(w, sink); }
  return 0;
}
