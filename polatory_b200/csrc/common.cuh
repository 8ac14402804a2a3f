// Shared host/device plumbing: error type, stream-ordered device buffers, launch helpers.
#pragma once

#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <algorithm>
#include <utility>
#include <vector>

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

// Grow-only bump arena for the temporaries of one evaluate(): after the first call with a given
// problem shape the steady state performs no allocation at all (no cudaMalloc, no pool
// traffic); reset() rewinds it.  Blocks are 256-byte aligned.  Growth adds a block; the next
// reset() coalesces the blocks into one (the only place that synchronises the device).
class Arena {
 public:
  Arena() = default;
  Arena(const Arena&) = delete;
  Arena& operator=(const Arena&) = delete;
  ~Arena() {
    for (auto& b : blocks_) cudaFree(b.ptr);
  }
  void reset() {
    if (blocks_.size() > 1) {
      size_t total = 0;
      for (auto& b : blocks_) total += b.cap;
      PLT_CUDA(cudaDeviceSynchronize());
      for (auto& b : blocks_) cudaFree(b.ptr);
      blocks_.clear();
      add_block(total);
    }
    for (auto& b : blocks_) b.used = 0;
    high_ = 0;
  }
  template <class T>
  T* take(size_t n) {
    size_t bytes = (n * sizeof(T) + 255) & ~size_t{255};
    if (bytes == 0) bytes = 256;
    for (auto& b : blocks_) {
      if (b.used + bytes <= b.cap) {
        T* p = reinterpret_cast<T*>(static_cast<char*>(b.ptr) + b.used);
        b.used += bytes;
        return p;
      }
    }
    add_block(std::max(bytes, kMinBlock));
    blocks_.back().used = bytes;
    return static_cast<T*>(blocks_.back().ptr);
  }
  // Snapshot / restore of the bump pointers (nested temporaries, e.g. the accuracy search).
  std::vector<size_t> mark() const {
    std::vector<size_t> m;
    for (auto& b : blocks_) m.push_back(b.used);
    return m;
  }
  void rewind(const std::vector<size_t>& m) {
    for (size_t i = 0; i < blocks_.size(); ++i) blocks_[i].used = i < m.size() ? m[i] : 0;
  }
  size_t capacity() const {
    size_t t = 0;
    for (auto& b : blocks_) t += b.cap;
    return t;
  }

 private:
  static constexpr size_t kMinBlock = size_t{64} << 20;
  struct Block {
    void* ptr;
    size_t cap, used;
  };
  void add_block(size_t cap) {
    void* p = nullptr;
    PLT_CUDA(cudaMalloc(&p, cap));
    blocks_.push_back(Block{p, cap, 0});
  }
  std::vector<Block> blocks_;
  size_t high_ = 0;
};

inline int ceil_div(int64_t a, int64_t b) { return static_cast<int>((a + b - 1) / b); }

constexpr int kMaxDim = 3;
constexpr int kNumSM = 148;  // B200 (compile-time bound for fixed-size reduction grids)

// SM count of the current device (persistent grids are sized from it), queried once.
inline int num_sm() {
  static const int n = [] {
    int dev = 0, v = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) {
      cudaGetLastError();
      return kNumSM;
    }
    return v;
  }();
  return n;
}

}  // namespace plt
