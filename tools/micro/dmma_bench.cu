// Throughput of mma.sync.m8n8k4.f64 on one GPU: `chains` independent accumulator chains per warp, `warps` warps per SM.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_bench dmma_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int CH>
__global__ void k(double* out, int iters, double a0, double b0) {
  double c[CH][2];
#pragma unroll
  for (int j = 0; j < CH; j++) c[j][0] = c[j][1] = 0.0;
  double a = a0 + threadIdx.x, b = b0;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int j = 0; j < CH; j++) dmma(c[j][0], c[j][1], a, b);
  }
  double s = 0;
#pragma unroll
  for (int j = 0; j < CH; j++) s += c[j][0] + c[j][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int CH>
void run(int warps_per_sm, int sms, double* out) {
  const int iters = 20000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<CH><<<sms, 32 * warps_per_sm>>>(out, 100, 1.0, 1e-9);
  cudaEventRecord(e0);
  k<CH><<<sms, 32 * warps_per_sm>>>(out, iters, 1.0, 1e-9);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double n = (double)sms * warps_per_sm * iters * CH;
  printf("chains %2d warps/SM %2d : %.2f ms  %.1f GDMMA/s  %.2f TFLOP/s  (%.2f ns per DMMA per SM)\n", CH, warps_per_sm, ms, n / ms / 1e6, n * 512 / ms / 1e9,
         ms * 1e6 / (n / sms));
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  double* out; cudaMalloc(&out, 148 * 1024 * 8 * 2);
  printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
  for (int w : {1, 4, 8, 16, 32}) { run<1>(w, p.multiProcessorCount, out); run<5>(w, p.multiProcessorCount, out); run<10>(w, p.multiProcessorCount, out); }
  return 0;
}
