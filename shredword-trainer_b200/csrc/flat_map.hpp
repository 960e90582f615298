// flat_map.hpp -- minimal open-addressing u64 -> V map for the host control path (pair versions, phantom pairs).
// One 16-byte cell per entry (key, value, used) so a lookup touches one cache line; prefetch() lets the caller overlap
// the miss with other work (the heap replay does ~110 lookups per merge into a table of millions of pairs).
#pragma once
#include <sys/mman.h>

#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace shred {

inline uint64_t mix64(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
  return x;
}

// Growable zero-initialised array of trivially copyable T in 2 MB aligned memory with transparent huge pages requested:
// the per-pair version array is tens of MB and read at random, so 4 KB pages would miss the TLB on every access.
template <class T>
class HugeArray {
 public:
  ~HugeArray() { std::free(d_); }
  size_t size() const { return n_; }
  void clear() { n_ = 0; }
  T& operator[](size_t i) { return d_[i]; }
  const T* data() const { return d_; }
  void ensure(size_t n) {  // size() >= n afterwards, new elements are zero
    if (n <= n_) return;
    if (n > cap_) {
      size_t nc = cap_ ? cap_ : 4096;
      while (nc < n) nc *= 2;
      const size_t bytes = (nc * sizeof(T) + (2u << 20) - 1) & ~static_cast<size_t>((2u << 20) - 1);
      void* p = nullptr;
      if (posix_memalign(&p, 2u << 20, bytes) != 0 || !p) std::abort();
#ifdef MADV_HUGEPAGE
      madvise(p, bytes, MADV_HUGEPAGE);
#endif
      if (n_) std::memcpy(p, d_, n_ * sizeof(T));
      std::free(d_);
      d_ = static_cast<T*>(p);
      cap_ = bytes / sizeof(T);
    }
    std::memset(static_cast<void*>(d_ + n_), 0, (n - n_) * sizeof(T));
    n_ = n;
  }

 private:
  T* d_ = nullptr;
  size_t n_ = 0, cap_ = 0;
};

template <class V>
class FlatMap {
  struct Cell { uint64_t key; V val; uint32_t used; };

 public:
  void clear() { cells_.clear(); n_ = 0; }
  size_t size() const { return n_; }
  void prefetch(uint64_t k) const {
    if (!cells_.empty()) __builtin_prefetch(&cells_[mix64(k) & (cells_.size() - 1)]);
  }
  // get-or-create, value-initialised
  V& operator[](uint64_t k) {
    if ((n_ + 1) * 2 > cells_.size()) grow();
    size_t m = cells_.size() - 1, s = mix64(k) & m;
    while (cells_[s].used) { if (cells_[s].key == k) return cells_[s].val; s = (s + 1) & m; }
    cells_[s].used = 1; cells_[s].key = k; cells_[s].val = V(); ++n_;
    return cells_[s].val;
  }
  V* find(uint64_t k) {
    if (cells_.empty()) return nullptr;
    size_t m = cells_.size() - 1, s = mix64(k) & m;
    while (cells_[s].used) { if (cells_[s].key == k) return &cells_[s].val; s = (s + 1) & m; }
    return nullptr;
  }

 private:
  void grow() {
    size_t nc = cells_.empty() ? 1024 : cells_.size() * 2;
    std::vector<Cell> old;
    old.swap(cells_);
    cells_.assign(nc, Cell{0, V(), 0});
    size_t m = nc - 1;
    for (const Cell& c : old) if (c.used) {
      size_t s = mix64(c.key) & m;
      while (cells_[s].used) s = (s + 1) & m;
      cells_[s] = c;
    }
  }
  std::vector<Cell> cells_;
  size_t n_ = 0;
};

}  // namespace shred
