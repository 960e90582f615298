// flat_map.hpp -- minimal open-addressing u64 -> V map for the host control path (pair versions, phantom pairs).
#pragma once
#include <cstdint>
#include <cstdlib>
#include <vector>

namespace shred {

inline uint64_t mix64(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
  return x;
}

template <class V>
class FlatMap {
 public:
  void clear() { keys_.clear(); vals_.clear(); used_.clear(); n_ = 0; }
  size_t size() const { return n_; }
  // get-or-create, value-initialised
  V& operator[](uint64_t k) {
    if ((n_ + 1) * 2 > keys_.size()) grow();
    size_t m = keys_.size() - 1, s = mix64(k) & m;
    while (used_[s]) { if (keys_[s] == k) return vals_[s]; s = (s + 1) & m; }
    used_[s] = 1; keys_[s] = k; vals_[s] = V(); ++n_;
    return vals_[s];
  }
  V* find(uint64_t k) {
    if (keys_.empty()) return nullptr;
    size_t m = keys_.size() - 1, s = mix64(k) & m;
    while (used_[s]) { if (keys_[s] == k) return &vals_[s]; s = (s + 1) & m; }
    return nullptr;
  }

 private:
  void grow() {
    size_t nc = keys_.empty() ? 1024 : keys_.size() * 2;
    std::vector<uint64_t> ok; std::vector<V> ov; std::vector<uint8_t> ou;
    ok.swap(keys_); ov.swap(vals_); ou.swap(used_);
    keys_.assign(nc, 0); vals_.assign(nc, V()); used_.assign(nc, 0);
    size_t m = nc - 1;
    for (size_t i = 0; i < ok.size(); i++) if (ou[i]) {
      size_t s = mix64(ok[i]) & m;
      while (used_[s]) s = (s + 1) & m;
      used_[s] = 1; keys_[s] = ok[i]; vals_[s] = ov[i];
    }
  }
  std::vector<uint64_t> keys_;
  std::vector<V> vals_;
  std::vector<uint8_t> used_;
  size_t n_ = 0;
};

}  // namespace shred
