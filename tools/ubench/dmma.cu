// Microbenchmark: FP64 tensor-core (DMMA, mma.sync.m8n8k4.f64) against the DFMA pipe on B200.
// Each warp keeps NACC independent 8x8 accumulator tiles and issues back-to-back mma.sync on them.
#include <cstdio>
#include <cuda_runtime.h>

template <int NACC>
__global__ void k_dmma(double* out, int iters, double a0, double b0) {
  double c[NACC][2];
  for (int i = 0; i < NACC; ++i) c[i][0] = c[i][1] = 0.0;
  double a = a0 + threadIdx.x * 1e-9, b = b0 - threadIdx.x * 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0.0;
  for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void k_dfma(double* out, int iters, double a0, double b0) {
  double c[NACC];
  for (int i = 0; i < NACC; ++i) c[i] = i;
  double a = a0 + threadIdx.x * 1e-9, b = b0 - threadIdx.x * 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) c[i] = fma(a, c[i], b);
  }
  double s = 0.0;
  for (int i = 0; i < NACC; ++i) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double* out;
  cudaMalloc(&out, sizeof(double) * sms * 8 * 1024);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int iters = 20000;
  for (int warps : {4, 8, 16, 32}) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      k_dmma<8><<<sms, warps * 32>>>(out, iters, 1.0000001, 0.9999999);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      double flop = 2.0 * 8 * 8 * 4 * 8.0 * iters * warps * sms;
      if (rep) printf("DMMA m8n8k4: %2d warps/SM x 8 tiles: %.3f ms  %.2f TFLOP/s\n", warps, ms, flop / ms / 1e9);
      cudaEventRecord(e0);
      k_dfma<8><<<sms, warps * 32>>>(out, iters, 1.0000001, 0.9999999);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      cudaEventElapsedTime(&ms, e0, e1);
      flop = 2.0 * 32 * 8.0 * iters * warps * sms;
      if (rep) printf("DFMA       : %2d warps/SM x 8 chains: %.3f ms  %.2f TFLOP/s\n", warps, ms, flop / ms / 1e9);
    }
  }
  printf("cuda status: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
