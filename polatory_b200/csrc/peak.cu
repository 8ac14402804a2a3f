// FP64 FMA peak microbenchmark ("not in MEASURED_PEAKS.json -- measure with a
// DFMA-chain microbenchmark first").  8 independent DFMA chains per thread, enough resident
// warps to saturate the FP64 pipe of every SM; timed with CUDA events.
#include <cuda_runtime.h>

#include "../../include/polatory_b200.h"

namespace {
__global__ void __launch_bounds__(256) k_dfma_chain(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
    x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
    x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}
}  // namespace

extern "C" int plt_measure_fp64_peak(double* tflops) {
  if (!tflops) return PLT_ERR_INVALID;
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
    return PLT_ERR_CUDA;
  const int blocks = sms * 8, threads = 256, iters = 1 << 15;
  double* buf = nullptr;
  if (cudaMalloc(&buf, sizeof(double) * blocks * threads) != cudaSuccess) return PLT_ERR_CUDA;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  double best = 0.0;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0);
    k_dfma_chain<<<blocks, threads>>>(buf, iters, 0.999999, 1e-9);
    cudaEventRecord(e1);
    if (cudaEventSynchronize(e1) != cudaSuccess) { cudaFree(buf); return PLT_ERR_CUDA; }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    double flops = 2.0 * 8.0 * iters * static_cast<double>(blocks) * threads;
    double tf = flops / (ms * 1e-3) / 1e12;
    if (rep > 0 && tf > best) best = tf;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(buf);
  *tflops = best;
  return PLT_OK;
}
