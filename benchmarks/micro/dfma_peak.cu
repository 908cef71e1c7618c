// FP64 FMA peak micro-benchmark variants (roofline denominator sanity check).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o dfma_peak dfma_peak.cu && ./dfma_peak
#include <cstdio>
#include <cuda_runtime.h>
template <int CH, bool REGOPS>
__global__ void __launch_bounds__(256) k(double* out, int iters, double a, double b) {
  double x[CH];
  double aa = a + threadIdx.x * 1e-12, bb = b + threadIdx.x * 1e-13;
#pragma unroll
  for (int i = 0; i < CH; ++i) x[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i) x[i] = REGOPS ? fma(x[i], aa, bb) : fma(x[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < CH; ++i) s += x[i];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int CH, bool REGOPS>
void run(const char* name, int blocks_per_sm) {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int blocks = p.multiProcessorCount * blocks_per_sm, threads = 256, iters = 1 << 14;
  double* out; cudaMalloc(&out, (size_t)blocks * threads * 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  double best = 0;
  for (int r = 0; r < 5; ++r) {
    cudaEventRecord(e0); k<CH, REGOPS><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double tf = 2.0 * CH * iters * (double)blocks * threads / (ms * 1e-3) / 1e12; if (r && tf > best) best = tf;
  }
  printf("%-28s blocks/SM=%d  %.2f TFLOP/s\n", name, blocks_per_sm, best);
  cudaFree(out);
}
int main() {
  run<8, false>("8 chains, const operands", 8);
  run<8, true>("8 chains, reg operands", 8);
  run<16, false>("16 chains, const operands", 4);
  run<16, true>("16 chains, reg operands", 4);
  run<4, true>("4 chains, reg operands", 8);
  run<16, true>("16 chains, reg operands", 2);
  run<32, true>("32 chains, reg operands", 2);
  return 0;
}
