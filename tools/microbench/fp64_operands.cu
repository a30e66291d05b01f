// Does DFMA throughput depend on operand register diversity (RF banking / reuse cache)?
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(int iters, double* sink, double seed) {
  double a[8], b[8], d[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) { a[c] = 1.0 + 1e-9 * (threadIdx.x + c); b[c] = 0.999999 + seed * c; d[c] = 1e-12 + seed * c; }
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      if (MODE == 0) a[c] = fma(a[c], b[0], d[0]);          // shared multiplier/addend
      if (MODE == 1) a[c] = fma(a[c], b[c], d[c]);          // three distinct operands per DFMA
      if (MODE == 2) a[c] = fma(b[c], d[(c + 1) & 7], a[c]);  // accumulate form, rotating operands
      if (MODE == 3) a[c] = a[c] * b[c];                    // DMUL
      if (MODE == 4) a[c] = a[c] + b[c];                    // DADD
    }
  }
  double s = 0;
#pragma unroll
  for (int c = 0; c < 8; ++c) s += a[c] + b[c] + d[c];
  if (s == 1234.5) *sink = s;
}

template <int MODE> void run(const char* name, double* sink) {
  int iters = 1 << 15;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<148 * 4, 128>>>(iters / 4, sink, 0.0);
  cudaEventRecord(e0);
  k<MODE><<<148 * 4, 128>>>(iters, sink, 0.0);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double ops = (double)iters * 8 * 128 * 148.0 * 4 / (ms * 1e-3);
  printf("%-40s %6.1f lanes/clk/SM @1.965GHz\n", name, ops / 148.0 / 1.965e9);
}

int main() {
  double* sink; cudaMalloc(&sink, 8);
  run<0>("DFMA a=fma(a,m,k) shared m,k", sink);
  run<1>("DFMA a=fma(a,b_c,d_c) distinct", sink);
  run<2>("DFMA a=fma(b_c,d_c',a) accumulate", sink);
  run<3>("DMUL", sink);
  run<4>("DADD", sink);
  return 0;
}
