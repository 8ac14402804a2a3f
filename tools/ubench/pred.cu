// Microbenchmark: does a predicated-off DFMA occupy the FP64 pipe on B200?  8 independent chains per thread, every
// fma under a runtime predicate (inline PTX; checked in SASS: @P DFMA, no select).  on = all predicates true,
// off = all false, half = every second chain predicated off.
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_pred(double* out, int iters, double a0, double b0, unsigned mask) {
  double c[8];
  for (int i = 0; i < 8; ++i) c[i] = i;
  double a = a0 + threadIdx.x * 1e-9, b = b0 - threadIdx.x * 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; i += 4)
      asm volatile("{.reg .pred p; setp.eq.u32 p, %6, 0; @p bra.uni SKIP;\n\t"
                   "fma.rn.f64 %0, %4, %0, %5; fma.rn.f64 %1, %4, %1, %5; fma.rn.f64 %2, %4, %2, %5; fma.rn.f64 %3, %4, %3, %5;\n\t"
                   "SKIP: }"
                   : "+d"(c[i]), "+d"(c[i + 1]), "+d"(c[i + 2]), "+d"(c[i + 3]) : "d"(a), "d"(b), "r"(mask & (1u << i)));
  }
  double s = 0.0;
  for (int i = 0; i < 8; ++i) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double* out;
  cudaMalloc(&out, sizeof(double) * sms * 1024);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int iters = 20000;
  for (int warps : {8, 16, 32}) {
    for (unsigned mask : {0xffu, 0x0fu, 0x00u}) {
      float ms = 0;
      for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        k_pred<<<sms, warps * 32>>>(out, iters, 1.0000001, 0.9999999, mask);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
      }
      printf("%2d warps/SM, predicate mask %02x: %.3f ms  (%.2f cycles per warp-level DFMA slot per SMSP at 1.965 GHz)\n", warps, mask, ms,
             ms * 1e-3 * 1.965e9 / (8.0 * iters * warps / 4.0));
    }
  }
  return 0;
}
