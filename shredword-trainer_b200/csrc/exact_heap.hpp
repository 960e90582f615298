// exact_heap.hpp -- the replay heap.
//
// The reference picks the next merge with an array binary max-heap that compares `freq` only
// (reference shredword/csrc/bpe/heap.cpp:53-114), so which of several equal-frequency pairs wins is decided by
// the heap's structural history.  Bit-exact merge lists therefore require replaying the same sift rules on the same
// push/pop sequence.  Entries use the reference's 24-byte layout (heap.h:17-21) so that
// Trainer.heap.data can be read by C consumers of the reference ABI.
#pragma once
#include <cstdint>
#include <cstdlib>
#include <cstring>

#include "../../include/shred_abi.h"

namespace shred {

class ExactHeap {
 public:
  ~ExactHeap() { std::free(d_); }
  void clear() { n_ = 0; }
  size_t size() const { return n_; }
  size_t capacity() const { return cap_; }
  bool empty() const { return n_ == 0; }
  BPEHeapEntry* data() { return d_; }
  const BPEHeapEntry& top() const { return d_[0]; }

  // heap.cpp:70-79: append, then swap upwards while the parent's freq is strictly smaller.
  void push(PairKey key, uint64_t freq, uint32_t version) {
    if (n_ == cap_) {
      cap_ = cap_ ? cap_ * 2 : 4096;  // bpe.h:19 MIN_HEAP_SIZE, doubling as heap.cpp:59-68
      d_ = static_cast<BPEHeapEntry*>(std::realloc(d_, cap_ * sizeof(BPEHeapEntry)));
      if (!d_) { std::abort(); }
    }
    size_t i = n_++;
    BPEHeapEntry x;
    std::memset(&x, 0, sizeof x);
    x.key = key; x.freq = freq; x.version = version;
    while (i > 0) {
      size_t p = (i - 1) >> 1;
      if (d_[p].freq >= freq) break;
      d_[i] = d_[p];
      i = p;
    }
    d_[i] = x;
    ++pushes;
  }

  // heap.cpp:97-111: last entry to the root, then swap with the left child if it is strictly larger, with the right
  // child if it is strictly larger than the better of the two, until neither is.
  BPEHeapEntry pop() {
    BPEHeapEntry top = d_[0];
    BPEHeapEntry x = d_[--n_];
    size_t i = 0;
    for (;;) {
      // the walk is a chain of dependent cache misses once it leaves the hot top levels: pull in the eight
      // great-grandchildren (contiguous, 192 bytes) while the next two levels are being compared
      const size_t g = 8 * i + 7;
      if (g < n_) { __builtin_prefetch(d_ + g); __builtin_prefetch(d_ + g + 3); __builtin_prefetch(d_ + g + 6); __builtin_prefetch(d_ + g + 7); }
      size_t l = 2 * i + 1, r = l + 1, best = i;
      uint64_t bf = x.freq;
      if (l < n_ && d_[l].freq > bf) { best = l; bf = d_[l].freq; }
      if (r < n_ && d_[r].freq > bf) { best = r; }
      if (best == i) break;
      d_[i] = d_[best];
      i = best;
    }
    if (n_) d_[i] = x;
    ++pops;
    return top;
  }

  uint64_t pushes = 0, pops = 0;

 private:
  BPEHeapEntry* d_ = nullptr;
  size_t n_ = 0, cap_ = 0;
};

}  // namespace shred
