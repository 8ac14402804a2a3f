// Shared host/device plumbing: error type, stream-ordered device buffers, launch helpers.
#pragma once

#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <utility>

#include "../../include/polatory_b200.h"

namespace plt {

struct Error : std::runtime_error {
  int status;
  Error(int s, const std::string& m) : std::runtime_error(m), status(s) {}
};

#define PLT_CUDA(expr)                                                                     \
  do {                                                                                     \
    cudaError_t e__ = (expr);                                                              \
    if (e__ != cudaSuccess)                                                                \
      throw ::plt::Error(PLT_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__)); \
  } while (0)

#define PLT_REQUIRE(cond, msg)                                   \
  do {                                                           \
    if (!(cond)) throw ::plt::Error(PLT_ERR_INVALID, (msg));     \
  } while (0)

// Launch bookkeeping: every kernel launch of the library goes through PLT_LAUNCH so that
// plt_eval_launch_count() is a count, not an estimate.
struct LaunchCounter {
  int64_t n = 0;
};

#define PLT_LAUNCH(ctr, kernel, grid, block, smem, stream, ...)              \
  do {                                                                       \
    kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);              \
    (ctr).n++;                                                               \
    PLT_CUDA(cudaGetLastError());                                            \
  } while (0)

// Stream-ordered RAII device buffer (cudaMallocAsync pool: reuse without device syncs).
template <class T>
class DevBuf {
 public:
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  DevBuf(DevBuf&& o) noexcept { swap(o); }
  DevBuf& operator=(DevBuf&& o) noexcept {
    if (this != &o) {
      release();
      swap(o);
    }
    return *this;
  }
  ~DevBuf() { release(); }

  void alloc(size_t n, cudaStream_t s) {
    if (n <= cap_ && ptr_) {
      n_ = n;
      stream_ = s;
      return;
    }
    release();
    stream_ = s;
    if (n > 0) {
      void* p = nullptr;
      PLT_CUDA(cudaMallocAsync(&p, n * sizeof(T), s));
      ptr_ = static_cast<T*>(p);
    }
    n_ = cap_ = n;
  }
  void zero(cudaStream_t s) {
    if (n_) PLT_CUDA(cudaMemsetAsync(ptr_, 0, n_ * sizeof(T), s));
  }
  void fill_byte(int v, cudaStream_t s) {
    if (n_) PLT_CUDA(cudaMemsetAsync(ptr_, v, n_ * sizeof(T), s));
  }
  void release() {
    if (ptr_) cudaFreeAsync(ptr_, stream_);
    ptr_ = nullptr;
    n_ = cap_ = 0;
  }
  T* get() const { return ptr_; }
  size_t size() const { return n_; }
  explicit operator bool() const { return ptr_ != nullptr; }

 private:
  void swap(DevBuf& o) {
    std::swap(ptr_, o.ptr_);
    std::swap(n_, o.n_);
    std::swap(cap_, o.cap_);
    std::swap(stream_, o.stream_);
  }
  T* ptr_ = nullptr;
  size_t n_ = 0, cap_ = 0;
  cudaStream_t stream_ = nullptr;
};

inline int ceil_div(int64_t a, int64_t b) { return static_cast<int>((a + b - 1) / b); }

constexpr int kMaxDim = 3;
constexpr int kNumSM = 148;  // B200

}  // namespace plt
