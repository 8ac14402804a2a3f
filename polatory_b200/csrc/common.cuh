// Shared host/device plumbing: error type, stream-ordered device buffers, launch helpers.
#pragma once

#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <algorithm>
#include <utility>
#include <map>
#include <mutex>
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

size_t release_cached_arena_blocks();  // ArenaBlockCache::release_all (below)

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
      if (cudaMallocAsync(&p, n * sizeof(T), s) != cudaSuccess) {  // out of memory: drop the cached arena blocks, retry
        cudaGetLastError();
        release_cached_arena_blocks();
        PLT_CUDA(cudaMallocAsync(&p, n * sizeof(T), s));
      }
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

// Process-wide cache of arena blocks.  An evaluator's arena settles at ONE block of its steady-state size; a fit creates
// and destroys a few dozen evaluators (matvec and residual operators, the level transfers of the preconditioner), and
// the next fit -- or the next sampler over another model -- asks for exactly the same sizes again.  cudaMalloc / cudaFree
// of multi-GB blocks cost 10 - 100 ms each and synchronise the device, so released blocks are kept here (up to a
// quarter of the device memory, plt_set_cached_memory_limit) and handed out again when a request of about that size
// comes (at most 25 % larger than asked).  plt_release_cached_memory() returns everything to the driver; an allocation
// failure does the same before it retries.
class ArenaBlockCache {
 public:
  static ArenaBlockCache& get() {
    static ArenaBlockCache* c = new ArenaBlockCache();  // never destroyed: the CUDA context may be gone by then
    return *c;
  }
  void* take(size_t want, size_t& cap) {
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(mu_);
    for (auto it = free_.lower_bound(want); it != free_.end() && it->first <= want + want / 4; ++it) {
      if (it->second.dev != dev) continue;
      void* p = it->second.ptr;
      cap = it->first;
      total_ -= cap;
      free_.erase(it);
      return p;
    }
    return nullptr;
  }
  // The caller has synchronised the device: no kernel still uses the block.
  void give(void* p, size_t cap) {
    int dev = 0;
    cudaGetDevice(&dev);
    {
      std::lock_guard<std::mutex> lock(mu_);
      if (limit_ == kUnset) {
        size_t free_b = 0, total_b = 0;
        limit_ = cudaMemGetInfo(&free_b, &total_b) == cudaSuccess ? total_b / 4 : 0;
        cudaGetLastError();
      }
      if (total_ + cap <= limit_) {
        free_.emplace(cap, Entry{p, dev});
        total_ += cap;
        return;
      }
    }
    cudaFree(p);
  }
  size_t release_all() {
    std::multimap<size_t, Entry> drop;
    size_t bytes = 0;
    {
      std::lock_guard<std::mutex> lock(mu_);
      drop.swap(free_);
      bytes = total_;
      total_ = 0;
    }
    for (auto& e : drop) cudaFree(e.second.ptr);
    cudaGetLastError();
    return bytes;
  }
  void set_limit(size_t bytes) {
    {
      std::lock_guard<std::mutex> lock(mu_);
      limit_ = bytes;
    }
    if (cached() > bytes) release_all();
  }
  size_t cached() {
    std::lock_guard<std::mutex> lock(mu_);
    return total_;
  }

 private:
  static constexpr size_t kUnset = ~size_t{0};
  struct Entry {
    void* ptr;
    int dev;
  };
  std::mutex mu_;
  std::multimap<size_t, Entry> free_;
  size_t total_ = 0, limit_ = kUnset;
};

inline size_t release_cached_arena_blocks() { return ArenaBlockCache::get().release_all(); }

// Grow-only bump arena for the temporaries of one evaluate(): after the first call with a given
// problem shape the steady state performs no allocation at all (no cudaMalloc, no pool
// traffic); reset() rewinds it.  Blocks are 256-byte aligned.  Growth adds a block; the next
// reset() coalesces the blocks into one (the only place that synchronises the device).  Blocks come from and go back to
// the process-wide ArenaBlockCache.
class Arena {
 public:
  Arena() = default;
  Arena(const Arena&) = delete;
  Arena& operator=(const Arena&) = delete;
  ~Arena() {
    if (blocks_.empty()) return;
    cudaDeviceSynchronize();  // (errors ignored: at interpreter exit the context may be gone)
    cudaGetLastError();
    for (auto& b : blocks_) ArenaBlockCache::get().give(b.ptr, b.cap);
  }
  void reset() {
    if (blocks_.size() > 1) {
      size_t total = 0;
      for (auto& b : blocks_) total += b.cap;
      PLT_CUDA(cudaDeviceSynchronize());
      for (auto& b : blocks_) ArenaBlockCache::get().give(b.ptr, b.cap);
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
    size_t got = cap;
    void* p = ArenaBlockCache::get().take(cap, got);
    if (!p) {
      got = cap;
      if (cudaMalloc(&p, cap) != cudaSuccess) {  // out of memory: give the cached blocks back to the driver first
        cudaGetLastError();
        ArenaBlockCache::get().release_all();
        PLT_CUDA(cudaMalloc(&p, cap));
      }
    }
    blocks_.push_back(Block{p, got, 0});
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
