// Microbenchmark: operand-delivery paths for the Hadamard M2L inner loop on sm_100a.
// Per-SM throughput (cycles per warp-level instruction) of LDS.128 / LDS.64 / tcgen05.ld from TMEM,
// alone, mixed, and against DFMA; and the cost of predicated-off DFMAs.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ void lds128(double& a, double& b, uint32_t addr) {
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(a), "=d"(b) : "r"(addr));
}
__device__ __forceinline__ void lds64(double& a, uint32_t addr) {
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(a) : "r"(addr));
}
__device__ __forceinline__ void ldtm2(uint32_t& a, uint32_t& b, uint32_t taddr) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(a), "=r"(b) : "r"(taddr));
}
__device__ __forceinline__ void ldtm4(uint32_t& a, uint32_t& b, uint32_t& c, uint32_t& d, uint32_t taddr) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(taddr));
}
__device__ __forceinline__ void tm_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

template <int MODE>
__global__ void __launch_bounds__(512, 1) k_bench(int iters, int pred_on, double seed, long long* cycles, double* sink) {
  extern __shared__ double sm[];
  __shared__ uint32_t s_tmem;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 16384; i += blockDim.x) sm[i] = seed * i;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&s_tmem)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tbase = s_tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  // fill this warp's lane quarter (every warp of the same quarter writes the same data)
  for (int c = 0; c < 512; c += 2) {
    uint32_t v0 = lane * 1000 + c, v1 = v0 + 1;
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(tbase + c), "r"(v0), "r"(v1));
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");

  const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(sm);
  double acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = seed + j;
  uint32_t isink = 0;
  const bool p = pred_on != 0;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const uint32_t so = sbase + ((it * 1024) & 0xffff);  // wanders over 64 KB + 8*... (stay inside 128 KB)
    const uint32_t to = tbase + ((it * 32) & 0xff);
    if constexpr (MODE == 0) {  // 8 x LDS.128
#pragma unroll
      for (int j = 0; j < 8; ++j) { double a, b; lds128(a, b, so + j * 512 + lane * 16); acc[j] += a + b; }
    } else if constexpr (MODE == 1) {  // 8 x LDS.64
#pragma unroll
      for (int j = 0; j < 8; ++j) { double a; lds64(a, so + j * 256 + lane * 8); acc[j] += a; }
    } else if constexpr (MODE == 2) {  // 8 x LDTM.x2
      uint32_t r[16];
#pragma unroll
      for (int j = 0; j < 8; ++j) ldtm2(r[2 * j], r[2 * j + 1], to + j * 8);
      tm_wait_ld();
#pragma unroll
      for (int j = 0; j < 16; ++j) isink ^= r[j];
    } else if constexpr (MODE == 3) {  // 8 x LDTM.x4
      uint32_t r[32];
#pragma unroll
      for (int j = 0; j < 8; ++j) ldtm4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3], to + j * 8);
      tm_wait_ld();
#pragma unroll
      for (int j = 0; j < 32; ++j) isink ^= r[j];
    } else if constexpr (MODE == 4) {  // 8 x (LDS.64 + LDTM.x2)
      uint32_t r[16];
#pragma unroll
      for (int j = 0; j < 8; ++j) ldtm2(r[2 * j], r[2 * j + 1], to + j * 8);
#pragma unroll
      for (int j = 0; j < 8; ++j) { double a; lds64(a, so + j * 256 + lane * 8); acc[j] += a; }
      tm_wait_ld();
#pragma unroll
      for (int j = 0; j < 16; ++j) isink ^= r[j];
    } else if constexpr (MODE == 5) {  // 32 DFMA, predicated by p
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int j = 0; j < 8; ++j) if (p) acc[j] = fma(acc[j], 1.0000001, seed);
    } else if constexpr (MODE == 6) {  // 8 x (LDS.128 + 4 DFMA): the tiled kernel's ratio
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        double a, b; lds128(a, b, so + j * 512 + lane * 16);
        double x = acc[j], y = acc[(j + 4) & 7];
        x = fma(a, seed, x); x = fma(-b, seed, x); y = fma(a, seed, y); y = fma(b, seed, y);
        acc[j] = x; acc[(j + 4) & 7] = y;
      }
    } else if constexpr (MODE == 7) {  // 8 x (LDS.64 + LDTM.x2 + 4 DFMA)
      uint32_t r[16];
#pragma unroll
      for (int j = 0; j < 8; ++j) ldtm2(r[2 * j], r[2 * j + 1], to + j * 8);
      double a[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) lds64(a[j], so + j * 256 + lane * 8);
      tm_wait_ld();
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        double b = __hiloint2double(r[2 * j + 1], r[2 * j]);
        double x = acc[j], y = acc[(j + 4) & 7];
        x = fma(a[j], seed, x); x = fma(-b, seed, x); y = fma(a[j], seed, y); y = fma(b, seed, y);
        acc[j] = x; acc[(j + 4) & 7] = y;
      }
    } else if constexpr (MODE == 8) {  // 8 x (LDTM.x4 + 4 DFMA): everything from TMEM
      uint32_t r[32];
#pragma unroll
      for (int j = 0; j < 8; ++j) ldtm4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3], to + j * 8);
      tm_wait_ld();
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        double a = __hiloint2double(r[4 * j + 1], r[4 * j]), b = __hiloint2double(r[4 * j + 3], r[4 * j + 2]);
        double x = acc[j], y = acc[(j + 4) & 7];
        x = fma(a, seed, x); x = fma(-b, seed, x); y = fma(a, seed, y); y = fma(b, seed, y);
        acc[j] = x; acc[(j + 4) & 7] = y;
      }
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) s += acc[j];
  if (s == 12345.678 || isink == 0xdeadbeef) sink[threadIdx.x] = s + isink;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(s_tmem));
}

template <int MODE>
int run(const char* name, int warps, int pred_on, int ops_per_iter) {
  const int iters = 4000, grid = 148;
  long long* d_cycles; double* d_sink;
  CK(cudaMalloc(&d_cycles, grid * sizeof(long long)));
  CK(cudaMalloc(&d_sink, 512 * sizeof(double)));
  CK(cudaFuncSetAttribute(k_bench<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * 8));
  for (int rep = 0; rep < 2; ++rep) k_bench<MODE><<<grid, warps * 32, 16384 * 8>>>(iters, pred_on, 1e-9, d_cycles, d_sink);
  CK(cudaDeviceSynchronize());
  long long h[148];
  CK(cudaMemcpy(h, d_cycles, sizeof(h), cudaMemcpyDeviceToHost));
  double avg = 0;
  for (int i = 0; i < grid; ++i) avg += h[i];
  avg /= grid;
  printf("%-44s warps=%2d pred=%d  cycles/iter/warp=%8.2f  SM-cycles per warp-op=%6.3f\n", name, warps, pred_on,
         avg / iters, avg / iters / ops_per_iter / warps);
  cudaFree(d_cycles); cudaFree(d_sink);
  return 0;
}

int main() {
  for (int w : {4, 8, 16}) {
    run<0>("8x LDS.128 (512 B/warp)", w, 1, 8);
    run<1>("8x LDS.64  (256 B/warp)", w, 1, 8);
    run<2>("8x tcgen05.ld 32x32b.x2 (256 B/warp)", w, 1, 8);
    run<3>("8x tcgen05.ld 32x32b.x4 (512 B/warp)", w, 1, 8);
    run<4>("8x (LDS.64 + tcgen05.ld.x2)", w, 1, 8);
    run<5>("32x DFMA predicated ON", w, 1, 32);
    run<5>("32x DFMA predicated OFF", w, 0, 32);
    run<6>("8x (LDS.128 + 4 DFMA)  [per cFMA]", w, 1, 8);
    run<7>("8x (LDS.64 + LDTM.x2 + 4 DFMA) [per cFMA]", w, 1, 8);
    run<8>("8x (LDTM.x4 + 4 DFMA) [per cFMA]", w, 1, 8);
  }
  return 0;
}
